"""Run one representative epilogue-bound launch in isolation (for ncu). mode: c3 | d0 | c33"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch


def ACT():
    from pix2latent_b200 import native
    return native.act_dtype()
from test_conv_gemm_gpu import run_conv, pack_w, nhwc
import os
from pix2latent_b200 import _lib as _L
_L.lib()  # applies P2L_OPTS
modes = (sys.argv[1] if len(sys.argv) > 1 else "c3").split(",")
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"; torch.manual_seed(0)
N = 18
for mode in modes:
  if True:
    if mode == "c3":    # block 11 conv_3: 1x1 64->128 @256^2, +bias +upsampled skip, raw + bn/relu act
        H = W = 256; Cin, Cout = 64, 128
        x = torch.randn(N, H, W, Cin, device=dev).to(ACT())
        w = (torch.randn(Cout, Cin, 1, 1, device=dev) / 8).to(ACT())
        skip = torch.randn(N, H // 2, W // 2, 2 * Cout, device=dev).to(ACT())
        a = torch.randn(N, Cout, device=dev); s = torch.randn(N, Cout, device=dev); bias = torch.randn(Cout, device=dev)
        raw = torch.empty(N, H, W, Cout, device=dev, dtype=ACT()); act = torch.empty_like(raw)
        kw = dict(A=x, A_N=N, A_H=H, A_W=W, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=Cout, kh=1, kw=1, NI=N, H=H, W=W, BN=128, mode=0,
                  bias=bias, resid=skip, resid_C=2 * Cout, resid_shift=1, raw=raw, raw_C=Cout, aff_a=a, aff_s=s, aff_stride=Cout,
                  relu=1, act=act, act_C=Cout)
    elif mode == "d0":  # block 10 conv_0 dgrad: 1x1 64->256 @128^2, saved + stats + gain + skip gradient
        H = W = 128; C, Cout = 64, 256
        g = torch.randn(N, H, W, C, device=dev).to(ACT())
        w = (torch.randn(Cout, C, 1, 1, device=dev) / 8).to(ACT())
        saved = torch.relu(torch.randn(N, H, W, Cout, device=dev)).to(ACT())
        addin = torch.randn(N, H, W, Cout, device=dev).to(ACT())
        a = torch.randn(N, Cout, device=dev); st0 = torch.zeros(N, Cout, device=dev); st1 = torch.zeros(N, Cout, device=dev)
        out = torch.empty(N, H, W, Cout, device=dev, dtype=ACT())
        kw = dict(A=g, A_N=N, A_H=H, A_W=W, A_C=C, Cin=C, B=pack_w(w), Cout=Cout, kh=1, kw=1, NI=N, H=H, W=W, BN=128, mode=1,
                  saved=saved, saved_C=Cout, stat0=st0, stat1=st1, stat_stride=Cout, aff_a=a, aff_stride=Cout, addin=addin,
                  addin_C=Cout, addin_climit=Cout, addin_pool=0, dx=out, dx_C=Cout)
    elif mode in ("d33s", "d33"):  # block 11 conv_2 / conv_1 dgrad: 3x3 64->64 @256^2, (saved + stats + gain | plain)
        H = W = 256; Cin = Cout = 64
        g = torch.randn(N, H, W, Cin, device=dev).to(ACT())
        w = (torch.randn(Cout, Cin, 3, 3, device=dev) / 24).to(ACT())
        saved = torch.relu(torch.randn(N, H, W, Cout, device=dev)).to(ACT())
        a = torch.randn(N, Cout, device=dev); st0 = torch.zeros(N, Cout, device=dev); st1 = torch.zeros(N, Cout, device=dev)
        out = torch.empty(N, H, W, Cout, device=dev, dtype=ACT())
        kw = dict(A=g, A_N=N, A_H=H, A_W=W, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=Cout, kh=3, kw=3, pad_h=1, pad_w=1, NI=N, H=H, W=W,
                  BN=64, mode=1, dx=out, dx_C=Cout)
        if mode == "d33s":
            kw.update(saved=saved, saved_C=Cout, stat0=st0, stat1=st1, stat_stride=Cout, aff_a=a, aff_stride=Cout)
    elif mode == "c3a":   # block 11 conv_3 as the last block runs it: activated output only (nothing reads its raw output)
        H = W = 256; Cin, Cout = 64, 128
        x = torch.randn(N, H, W, Cin, device=dev).to(ACT())
        w = (torch.randn(Cout, Cin, 1, 1, device=dev) / 8).to(ACT())
        skip = torch.randn(N, H // 2, W // 2, 2 * Cout, device=dev).to(ACT())
        a = torch.randn(N, Cout, device=dev); s = torch.randn(N, Cout, device=dev); bias = torch.randn(Cout, device=dev)
        act = torch.empty(N, H, W, Cout, device=dev, dtype=ACT())
        kw = dict(A=x, A_N=N, A_H=H, A_W=W, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=Cout, kh=1, kw=1, NI=N, H=H, W=W, BN=128, mode=0,
                  bias=bias, resid=skip, resid_C=2 * Cout, resid_shift=1, aff_a=a, aff_s=s, aff_stride=Cout,
                  relu=1, act=act, act_C=Cout)
    elif mode == "lo3":  # block 0 conv_1: 3x3 512->512 @4x4 (M = 288 pixels: 24 tiles of N = 64, K = 4608), bn/relu act
        H = W = 4; Cin = Cout = 512
        x = torch.randn(N, H, W, Cin, device=dev).to(ACT())
        w = (torch.randn(Cout, Cin, 3, 3, device=dev) / 68).to(ACT())
        a = torch.randn(N, Cout, device=dev); s = torch.randn(N, Cout, device=dev); bias = torch.randn(Cout, device=dev)
        act = torch.empty(N, H, W, Cout, device=dev, dtype=ACT())
        kw = dict(A=x, A_N=N, A_H=H, A_W=W, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=Cout, kh=3, kw=3, pad_h=1, pad_w=1, NI=N, H=H, W=W,
                  BN=64, mode=0, bias=bias, aff_a=a, aff_s=s, aff_stride=Cout, relu=1, act=act, act_C=Cout)
    elif mode == "lo1":  # block 0 conv_0: 1x1 2048->512 @4x4 (24 tiles, K = 2048)
        H = W = 4; Cin, Cout = 2048, 512
        x = torch.randn(N, H, W, Cin, device=dev).to(ACT())
        w = (torch.randn(Cout, Cin, 1, 1, device=dev) / 45).to(ACT())
        a = torch.randn(N, Cout, device=dev); s = torch.randn(N, Cout, device=dev); bias = torch.randn(Cout, device=dev)
        act = torch.empty(N, H, W, Cout, device=dev, dtype=ACT())
        kw = dict(A=x, A_N=N, A_H=H, A_W=W, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=Cout, kh=1, kw=1, NI=N, H=H, W=W,
                  BN=64, mode=0, bias=bias, aff_a=a, aff_s=s, aff_stride=Cout, relu=1, act=act, act_C=Cout)
    else:               # block 11 conv_1: 3x3 64->64 @256^2 with bn/relu act
        H = W = 256; Cin = Cout = 64
        x = torch.randn(N, H, W, Cin, device=dev).to(ACT())
        w = (torch.randn(Cout, Cin, 3, 3, device=dev) / 24).to(ACT())
        a = torch.randn(N, Cout, device=dev); s = torch.randn(N, Cout, device=dev); bias = torch.randn(Cout, device=dev)
        act = torch.empty(N, H, W, Cout, device=dev, dtype=ACT())
        kw = dict(A=x, A_N=N, A_H=H, A_W=W, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=Cout, kh=3, kw=3, pad_h=1, pad_w=1, NI=N, H=H, W=W,
                  BN=64, mode=0, bias=bias, aff_a=a, aff_s=s, aff_stride=Cout, relu=1, act=act, act_C=Cout)
    for _ in range(reps):
        run_conv(**kw)
    print("done", mode)
