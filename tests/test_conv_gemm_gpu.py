"""Parity of the tcgen05 implicit-GEMM convolution (pix2latent_b200/csrc/conv_gemm.cuh) against
plain PyTorch fp32 ops on the same bf16-rounded inputs. Every epilogue feature the BigGAN /
LPIPS path uses is exercised here in isolation (per-kernel parity)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def ACT():
    from pix2latent_b200 import native
    return native.act_dtype()


def _lib():
    from pix2latent_b200 import _lib
    return _lib


def pack_w(w):
    """torch conv weight [Cout, Cin, kh, kw] -> bf16 [Cout, kh*kw*Cin] (tap-major, channel-minor)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(ACT())


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def run_conv(**kw):
    L = _lib()
    a = L.ConvArgs()
    keep = []
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            keep.append(v)
            setattr(a, k, v.data_ptr())
        else:
            setattr(a, k, v)
    if a.alpha == 0:
        a.alpha = 1.0
    L.check(L.lib().p2l_debug_conv(a, L.current_stream()))
    torch.cuda.synchronize()


def rel_err(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


CASES_FWD = [
    # N, H, W, Cin, Cout, k, BN
    (2, 16, 16, 64, 64, 1, 64),
    (3, 32, 32, 128, 128, 3, 128),
    (1, 256, 256, 64, 64, 3, 64),
    (18, 4, 4, 256, 128, 3, 128),
    (5, 8, 8, 128, 256, 3, 256),
    (2, 15, 15, 192, 128, 3, 64),
    (2, 31, 31, 64, 192, 5, 64),
    (2, 64, 64, 512, 128, 3, 128),
    (3, 64, 64, 64, 256, 1, 256),
]


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,BN", CASES_FWD)
def test_conv_fwd_raw(N, H, W, Cin, Cout, k, BN):
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(N, Cin, H, W, device=dev).to(ACT())
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).to(ACT())
    bias = torch.randn(Cout, device=dev)
    ref = F.conv2d(x.float(), w.float(), bias, padding=k // 2)
    xa = nhwc(x)
    out = torch.zeros(N, H, W, Cout, device=dev, dtype=ACT())
    out32 = torch.zeros(N, H, W, Cout, device=dev, dtype=torch.float32)
    run_conv(A=xa, A_N=N, A_H=H, A_W=W, A_C=Cin, a_c0=0, Cin=Cin, B=pack_w(w), Cout=Cout, kh=k, kw=k,
             pad_h=k // 2, pad_w=k // 2, NI=N, H=H, W=W, BN=BN, mode=0, bias=bias,
             raw=out, raw_C=Cout, raw_f32=out32, raw_f32_C=Cout)
    assert rel_err(out32.permute(0, 3, 1, 2), ref) < 2e-3
    assert rel_err(out.permute(0, 3, 1, 2), ref) < 6e-3


@pytest.mark.parametrize("N,H,W,Cin,Cout,BN", [
    (3, 16, 16, 128, 64, 64),
    (18, 64, 64, 64, 128, 128),    # 576 tiles: several per persistent CTA, both epilogue groups, slab rings wrap
    (9, 64, 64, 256, 192, 64),     # 3 column tiles of 64
])
def test_conv_fwd_affine_relu_resid_up(N, H, W, Cin, Cout, BN):
    """conv_0/conv_3-style epilogue: bias, skip add (channel-sliced, nearest-upsampled), raw write,
    per-sample affine + relu, 2x replicated activated write."""
    torch.manual_seed(1)
    dev = "cuda"
    x = torch.randn(N, Cin, H, W, device=dev).to(ACT())
    w = (torch.randn(Cout, Cin, 1, 1, device=dev) / Cin ** 0.5).to(ACT())
    bias = torch.randn(Cout, device=dev)
    skip = torch.randn(N, 2 * Cout, H // 2, W // 2, device=dev).to(ACT())  # first Cout channels used
    a = torch.randn(N, Cout, device=dev)
    s = torch.randn(N, Cout, device=dev)
    v = F.conv2d(x.float(), w.float(), bias) + F.interpolate(skip[:, :Cout].float(), scale_factor=2, mode="nearest")
    y = torch.relu(a[:, :, None, None] * v + s[:, :, None, None])
    y_up = F.interpolate(y, scale_factor=2, mode="nearest")
    raw = torch.zeros(N, H, W, Cout, device=dev, dtype=ACT())
    act = torch.zeros(N, 2 * H, 2 * W, Cout, device=dev, dtype=ACT())
    act_lo = torch.zeros(N, H, W, Cout, device=dev, dtype=ACT())
    run_conv(A=nhwc(x), A_N=N, A_H=H, A_W=W, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=Cout, kh=1, kw=1,
             NI=N, H=H, W=W, BN=BN, mode=0, bias=bias, resid=nhwc(skip), resid_C=2 * Cout, resid_shift=1,
             raw=raw, raw_C=Cout, aff_a=a, aff_s=s, aff_stride=Cout, relu=1,
             act=act, act_C=Cout, act_up=1, act_lo=act_lo)
    assert rel_err(raw.permute(0, 3, 1, 2), v) < 6e-3
    assert rel_err(act_lo.permute(0, 3, 1, 2), y) < 8e-3
    assert rel_err(act.permute(0, 3, 1, 2), y_up) < 8e-3


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("N,H,W,C,BN", [(2, 256, 256, 64, 64), (3, 64, 64, 128, 128), (5, 40, 24, 64, 64)])
def test_conv3x3_paths_agree(mode, N, H, W, C, BN):
    """The three 3x3 main loops — per-tap A loads (halo_mode 0), halo patches with the weight matrix
    resident in shared memory (1, 64-channel layers) and halo patches with a weight ring (2) — against
    torch, forward (bias + affine + relu, raw and act) and backward (mask, statistics, gain)."""
    from pix2latent_b200 import _lib as L
    torch.manual_seed(7)
    dev = "cuda"
    x = torch.randn(N, C, H, W, device=dev).to(ACT())
    w = (torch.randn(C, C, 3, 3, device=dev) / (C * 9) ** 0.5).to(ACT())
    bias = torch.randn(C, device=dev)
    a = torch.randn(N, C, device=dev)
    s = torch.randn(N, C, device=dev)
    v = F.conv2d(x.float(), w.float(), bias, padding=1)
    y = torch.relu(a[:, :, None, None] * v + s[:, :, None, None])
    saved = y.to(ACT())
    acc = F.conv2d(x.float(), w.float(), padding=1)
    dpre = acc * (saved.float() > 0)
    L.set_option("halo_mode", mode)
    try:
        raw = torch.zeros(N, H, W, C, device=dev, dtype=ACT())
        act = torch.zeros(N, H, W, C, device=dev, dtype=ACT())
        run_conv(A=nhwc(x), A_N=N, A_H=H, A_W=W, A_C=C, Cin=C, B=pack_w(w), Cout=C, kh=3, kw=3, pad_h=1, pad_w=1,
                 NI=N, H=H, W=W, BN=BN, mode=0, bias=bias, raw=raw, raw_C=C, aff_a=a, aff_s=s, aff_stride=C, relu=1,
                 act=act, act_C=C)
        st0 = torch.zeros(N, C, device=dev)
        st1 = torch.zeros(N, C, device=dev)
        dx = torch.zeros(N, H, W, C, device=dev, dtype=ACT())
        run_conv(A=nhwc(x), A_N=N, A_H=H, A_W=W, A_C=C, Cin=C, B=pack_w(w), Cout=C, kh=3, kw=3, pad_h=1, pad_w=1,
                 NI=N, H=H, W=W, BN=BN, mode=1, saved=nhwc(saved), saved_C=C, stat0=st0, stat1=st1, stat_stride=C,
                 aff_a=a, aff_stride=C, dx=dx, dx_C=C)
    finally:
        L.set_option("halo_mode", 0)
    assert rel_err(raw.permute(0, 3, 1, 2), v) < 6e-3
    assert rel_err(act.permute(0, 3, 1, 2), y) < 8e-3
    assert rel_err(dx.permute(0, 3, 1, 2), dpre * a[:, :, None, None]) < 6e-3
    assert rel_err(st0, dpre.sum((2, 3))) < 2e-3
    assert rel_err(st1, (dpre * saved.float()).sum((2, 3))) < 2e-3


@pytest.mark.parametrize("halo_rgb", [0, 1])
@pytest.mark.parametrize("N,H,W", [(2, 64, 64), (3, 256, 256)])
def test_conv_fwd_rgb_tanh(halo_rgb, N, H, W):
    from pix2latent_b200 import _lib as L
    torch.manual_seed(2)
    dev = "cuda"
    Cin = 128
    x = torch.randn(N, Cin, H, W, device=dev).to(ACT())
    w = (torch.randn(3, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5).to(ACT())
    bias = torch.randn(3, device=dev) * 0.1
    ref = torch.tanh(F.conv2d(x.float(), w.float(), bias, padding=1))
    img = torch.zeros(N, 3, H, W, device=dev)
    L.set_option("halo_rgb", halo_rgb)
    try:
        run_conv(A=nhwc(x), A_N=N, A_H=H, A_W=W, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=3, kh=3, kw=3, pad_h=1, pad_w=1,
                 NI=N, H=H, W=W, BN=16, mode=0, bias=bias, img_nchw=img)
    finally:
        L.set_option("halo_rgb", 0)
    assert (img - ref).abs().max().item() < 5e-3


def test_gemm_batched_b_fp32_out():
    """attention logits: S[b] = theta[b] (4096x64) @ phi[b]^T (1024x64)."""
    torch.manual_seed(3)
    dev = "cuda"
    b, H, W, d, nk = 3, 64, 64, 64, 1024
    theta = torch.randn(b, H, W, d, device=dev).to(ACT())
    phi = torch.randn(b, nk, d, device=dev).to(ACT())
    ref = torch.einsum("bqd,bkd->bqk", theta.float().reshape(b, H * W, d), phi.float())
    S = torch.zeros(b, H * W, nk, device=dev)
    run_conv(A=theta, A_N=b, A_H=H, A_W=W, A_C=d, Cin=d, B=phi, Cout=nk, B_batch=b, kh=1, kw=1,
             NI=b, H=H, W=W, BN=128, mode=0, raw_f32=S, raw_f32_C=nk)
    assert rel_err(S, ref) < 2e-3


@pytest.mark.parametrize("N,H,W,C,Cout,k,BN,pool", [
    (3, 16, 16, 128, 128, 3, 128, 0),
    (2, 32, 32, 64, 256, 1, 128, 1),
    (18, 4, 4, 128, 64, 3, 64, 0),
    (5, 8, 8, 64, 128, 1, 128, 0),
    (1, 128, 128, 64, 64, 3, 64, 0),
    (18, 64, 64, 64, 256, 1, 128, 0),   # 1152 tiles, saved + skip-gradient slabs through the input rings
    (9, 64, 64, 128, 128, 1, 64, 0),
    (7, 32, 32, 512, 128, 1, 128, 0),
])
def test_conv_bwd(N, H, W, C, Cout, k, BN, pool):
    """dgrad-style launch: A = upstream gradient [N,H,W,C], B = transposed/flipped weights
    [Cout(=forward Cin), k*k*C]; epilogue = relu mask from the saved activation, BN-affine
    gradient sums, multiply by the affine gain, add the skip gradient."""
    torch.manual_seed(4)
    dev = "cuda"
    g = torch.randn(N, C, H, W, device=dev).to(ACT())
    w = (torch.randn(Cout, C, k, k, device=dev) / (C * k * k) ** 0.5).to(ACT())
    saved = torch.relu(torch.randn(N, Cout, H, W, device=dev)).to(ACT())
    a = torch.randn(N, Cout, device=dev)
    if pool:
        addin = torch.randn(N, Cout // 2, 2 * H, 2 * W, device=dev).to(ACT())
        add_ref = F.avg_pool2d(addin.float(), 2) * 4
    else:
        addin = torch.randn(N, Cout // 2, H, W, device=dev).to(ACT())
        add_ref = addin.float()
    acc = F.conv2d(g.float(), w.float(), padding=k // 2)
    dpre = acc * (saved.float() > 0)
    s0 = dpre.sum((2, 3))
    s1 = (dpre * saved.float()).sum((2, 3))
    dx = dpre * a[:, :, None, None]
    dx[:, :Cout // 2] += add_ref
    st0 = torch.zeros(N, Cout, device=dev)
    st1 = torch.zeros(N, Cout, device=dev)
    out = torch.zeros(N, H, W, Cout, device=dev, dtype=ACT())
    out32 = torch.zeros(N, H, W, Cout, device=dev)
    run_conv(A=nhwc(g), A_N=N, A_H=H, A_W=W, A_C=C, Cin=C, B=pack_w(w), Cout=Cout, kh=k, kw=k,
             pad_h=k // 2, pad_w=k // 2, NI=N, H=H, W=W, BN=BN, mode=1,
             saved=nhwc(saved), saved_C=Cout, stat0=st0, stat1=st1, stat_stride=Cout,
             aff_a=a, aff_stride=Cout, addin=nhwc(addin), addin_C=Cout // 2, addin_climit=Cout // 2,
             addin_pool=pool, dx=out, dx_C=Cout, dx_f32=out32, dx_f32_C=Cout)
    assert rel_err(out32.permute(0, 3, 1, 2), dx) < 2e-3
    assert rel_err(out.permute(0, 3, 1, 2), dx) < 6e-3
    assert rel_err(st0, s0) < 2e-3
    assert rel_err(st1, s1) < 2e-3
