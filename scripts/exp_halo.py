"""GPU experiment: which descriptor form makes the halo-patch 3x3 kernel correct, and what it buys."""
import sys, time
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch


def ACT():
    from pix2latent_b200 import native
    return native.act_dtype()
import torch.nn.functional as F
from pix2latent_b200 import _lib
from test_conv_gemm_gpu import run_conv, pack_w, nhwc, rel_err

def conv_case(N, H, W, Cin, Cout, BN, k=3, time_it=False, extra=None):
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(N, Cin, H, W, device=dev).to(ACT())
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).to(ACT())
    bias = torch.randn(Cout, device=dev)
    xa = nhwc(x); wp = pack_w(w)
    out = torch.zeros(N, H, W, Cout, device=dev, dtype=ACT())
    kw = dict(A=xa, A_N=N, A_H=H, A_W=W, A_C=Cin, a_c0=0, Cin=Cin, B=wp, Cout=Cout, kh=k, kw=k, pad_h=k // 2, pad_w=k // 2,
              NI=N, H=H, W=W, BN=BN, mode=0, bias=bias, raw=out, raw_C=Cout)
    if extra: kw.update(extra(N, H, W, Cout, dev))
    run_conv(**kw)
    err = None
    if N * H * W * Cout < 3e8:
        ref = F.conv2d(x.float(), w.float(), bias, padding=k // 2)
        err = rel_err(out.permute(0, 3, 1, 2), ref)
    t = None
    if time_it:
        for _ in range(3): run_conv(**kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run_conv(**kw)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 10
    return err, t

shapes = [(2, 32, 32, 128, 128, 128), (1, 64, 64, 64, 64, 64), (3, 16, 16, 64, 256, 256), (2, 20, 24, 64, 64, 64), (2, 15, 15, 192, 128, 64)]
for halo, bo in [(0, 0), (10, 0), (10, 1), (16, 0), (16, 1)]:
    _lib.set_option("halo", halo); _lib.set_option("halo_bo", bo)
    errs = []
    for s in shapes:
        try:
            e, _ = conv_case(*s)
            errs.append("%.1e" % e)
        except Exception as ex:
            errs.append("EXC " + str(ex)[:60])
    print("halo=%d bo=%d  rel errs: %s" % (halo, bo, errs), flush=True)

def flops(N, H, W, Cin, Cout, k): return 2.0 * N * H * W * Cin * Cout * k * k
big = [(18, 256, 256, 64, 64, 64, 3), (18, 128, 128, 128, 128, 128, 3), (18, 64, 64, 256, 256, 256, 3), (18, 64, 64, 256, 256, 128, 3),
       (18, 256, 256, 64, 128, 128, 1), (18, 128, 128, 512, 128, 128, 1)]
for halo, bo in [(0, 0), (10, 0), (16, 0)]:
    _lib.set_option("halo", halo); _lib.set_option("halo_bo", bo)
    for (N, H, W, Cin, Cout, BN, k) in big:
        if k == 1 and halo != 0: continue
        e, t = conv_case(N, H, W, Cin, Cout, BN, k, time_it=True)
        print("halo=%d  N%d %dx%d %d->%d k%d BN%d: %.1f us  %.0f TFLOP/s" % (halo, N, H, W, Cin, Cout, k, BN, t * 1e3, flops(N, H, W, Cin, Cout, k) / t / 1e9), flush=True)

# epilogue-heavy 1x1 (conv_3 of block 11): resid + raw + affine act
def extra(N, H, W, Cout, dev):
    skip = torch.randn(N, H // 2, W // 2, 2 * Cout, device=dev).to(ACT())
    a = torch.randn(N, Cout, device=dev); s = torch.randn(N, Cout, device=dev)
    act = torch.zeros(N, H, W, Cout, device=dev, dtype=ACT())
    return dict(resid=skip, resid_C=2 * Cout, resid_shift=1, aff_a=a, aff_s=s, aff_stride=Cout, relu=1, act=act, act_C=Cout)
_lib.set_option("halo", 0)
e, t = conv_case(18, 256, 256, 64, 128, 128, 1, time_it=True, extra=extra)
print("block11 conv_3 (1x1 64->128 +resid +raw +act): %.1f us ; bytes ~%.0f MB -> %.0f GB/s" % (t * 1e3, 18 * 65536 * (64 + 128 + 256 + 32) * 2 / 1e6, 18 * 65536 * (64 + 128 + 256 + 32) * 2 / t / 1e6))
