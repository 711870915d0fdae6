"""Kernel-selection options of the tensor-core path, each against the other setting on the same inputs (all through the
C-ABI): the row-wise softmax fusions of the attention GEMM epilogues ("attn_fused"), the transposed epilogue outputs
that replace the attention backward's transposes ("attn_emit_t"), and the serpentine tile order ("serpentine": every
other launch walks its tiles backwards so that it starts on what its producer left in L2). Kernel level
(p2l_debug_conv) against torch fp32, then model level: image and latent gradients of the reduced generator with the
option on / off."""
import contextlib
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)


@contextlib.contextmanager
def options(**kw):
    """Set kernel options for the duration of the block and restore what was there (defaults may be on or off)."""
    from pix2latent_b200 import _lib
    old = {k: _lib.get_option(k) for k in kw}
    try:
        for k, v in kw.items():
            _lib.set_option(k, v)
        yield
    finally:
        for k, v in old.items():
            _lib.set_option(k, v)


def test_two_pass_softmax_and_fused_ds():
    from pix2latent_b200 import native
    from test_conv_gemm_gpu import run_conv
    dt = native.act_dtype()
    torch.manual_seed(0)
    b, H, dq, Nk, dv = 2, 32, 64, 256, 128          # Nq = 1024 queries, 256 keys
    Nq = H * H
    theta = (torch.randn(b, Nq, dq, device="cuda") * 0.5).to(dt)
    phi = (torch.randn(b, Nk, dq, device="cuda") * 0.5).to(dt)
    S_ref = torch.bmm(theta.float(), phi.float().transpose(1, 2))
    P_ref = torch.softmax(S_ref, dim=-1)
    BN = 128
    nt = Nk // BN
    rowstat = torch.zeros(b * Nq * nt * 2, device="cuda")
    P = torch.empty(b, Nq, Nk, device="cuda", dtype=dt)
    PT = torch.zeros(b, Nk, Nq, device="cuda", dtype=dt)
    common = dict(A=theta, A_N=b, A_H=H, A_W=H, A_C=dq, Cin=dq, B=phi, Cout=Nk, B_batch=b, kh=1, kw=1, NI=b, H=H, W=H, BN=BN, mode=0)
    run_conv(rowstat=rowstat, rowstat_nt=nt, **common)
    run_conv(rowstat_in=rowstat, rowstat_nt=nt, raw=P, raw_C=Nk, outT=PT, outT_c0=0, outT_c1=Nk, **common)
    err = (P.float() - P_ref).abs().max().item()
    print("two-pass softmax max abs err %.2e" % err)
    assert err < 2e-3
    assert (P.float().sum(-1) - 1).abs().max().item() < 5e-3
    assert torch.equal(PT, P.transpose(1, 2).contiguous()), "transposed epilogue output != transpose of the main output"
    # dS = P o (dO g^T - D), D = rowsum(dO o O), O = P g
    g = (torch.randn(b, Nk, dv, device="cuda") * 0.5).to(dt)
    dO = (torch.randn(b, Nq, dv, device="cuda") * 0.5).to(dt)
    O = torch.bmm(P.float(), g.float())
    D = (dO.float() * O).sum(-1).contiguous()
    dP = torch.bmm(dO.float(), g.float().transpose(1, 2))
    dS_ref = P.float() * (dP - D[..., None])
    dS = torch.empty(b, Nq, Nk, device="cuda", dtype=dt)
    dST = torch.zeros(b, Nk, Nq, device="cuda", dtype=dt)
    run_conv(A=dO, A_N=b, A_H=H, A_W=H, A_C=dv, Cin=dv, B=g, Cout=Nk, B_batch=b, kh=1, kw=1, NI=b, H=H, W=H, BN=BN, mode=0,
             rowsub=D.view(-1), mulin=P, mulin_C=Nk, raw=dS, raw_C=Nk, outT=dST, outT_c0=0, outT_c1=Nk)
    rel = ((dS.float() - dS_ref).norm() / dS_ref.norm()).item()
    print("fused dS rel err %.2e" % rel)
    assert rel < 5e-3
    assert torch.equal(dST, dS.transpose(1, 2).contiguous())


@pytest.mark.parametrize("mode,BN,Cout,c0,c1", [(0, 128, 384, 0, 64), (1, 128, 256, 0, 256), (0, 64, 128, 32, 96)])
def test_transposed_output_channel_range(mode, BN, Cout, c0, c1):
    """outT for a channel sub-range, forward (raw, TMA-store path) and backward (dx) epilogues."""
    from pix2latent_b200 import native
    from test_conv_gemm_gpu import run_conv, pack_w
    dt = native.act_dtype()
    torch.manual_seed(1)
    N, H, Cin = 3, 32, 128
    x = torch.randn(N, H, H, Cin, device="cuda").to(dt)
    w = (torch.randn(Cout, Cin, 1, 1, device="cuda") / Cin ** 0.5).to(dt)
    main = torch.zeros(N, H, H, Cout, device="cuda", dtype=dt)
    outT = torch.full((N, c1 - c0, H * H), 7.0, device="cuda", dtype=dt)
    kw = dict(raw=main, raw_C=Cout) if mode == 0 else dict(dx=main, dx_C=Cout)
    run_conv(A=x, A_N=N, A_H=H, A_W=H, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=Cout, kh=1, kw=1, NI=N, H=H, W=H, BN=BN, mode=mode,
             outT=outT, outT_c0=c0, outT_c1=c1, **kw)
    assert torch.equal(outT, main.view(N, H * H, Cout)[:, :, c0:c1].transpose(1, 2).contiguous())


def _model_io(cfg, orc, b=3, seed=3):
    torch.manual_seed(seed)
    z = torch.fmod(torch.randn(b, 128), 2.0).cuda()
    c = orc.get_class_embedding(3).repeat(b, 1).cuda()
    dimg = torch.randn(b, 3, cfg.output_dim, cfg.output_dim, device="cuda") * 1e-2
    return z, c, dimg


def _run_model(cfg, orc, z, c, dimg):
    import test_step_gpu as ts
    from pix2latent_b200.model import BigGAN
    model = BigGAN(config=ts._product_cfg(cfg), state_dict=orc.state_dict())  # plans are built under the options in force
    img = model.native.forward(z, c)
    dz, dc = model.native.backward(z.shape[0], dimg)
    torch.cuda.synchronize()
    return img.clone(), dz.clone(), dc.clone()


def _cmp(name, a, b_):
    (i0, z0, c0), (i1, z1, c1) = a, b_
    rel = ((i1 - i0).norm() / i0.norm()).item()
    cz = torch.nn.functional.cosine_similarity(z0.flatten().double(), z1.flatten().double(), dim=0).item()
    cc = torch.nn.functional.cosine_similarity(c0.flatten().double(), c1.flatten().double(), dim=0).item()
    print("%s: image rel diff %.2e, cos dz %.6f, cos dc %.6f" % (name, rel, cz, cc))
    return rel, cz, cc


@pytest.fixture(scope="module")
def problem():
    import make_golden as mg
    cfg, orc, target, weight = mg.problem()
    assert cfg.attention_layer_position >= 0
    return cfg, orc


def test_generator_fused_attention(problem):
    cfg, orc = problem
    io = _model_io(cfg, orc)
    with options(attn_fused=0, attn_emit_t=0):
        base = _run_model(cfg, orc, *io)
    with options(attn_fused=1, attn_emit_t=0):
        fused = _run_model(cfg, orc, *io)
    with options(attn_fused=1, attn_emit_t=1):
        emit = _run_model(cfg, orc, *io)
    rel, cz, cc = _cmp("fused attention vs materialised logits", base, fused)
    assert rel < 2e-3 and cz > 0.999 and cc > 0.999
    # the transposed operands are the same 16-bit values the transpose kernels produced: identical results
    for u, v in zip(fused, emit):
        assert torch.equal(u, v), "transposed epilogue outputs changed the result"


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,H,Cin,Cout,k,BN", [(3, 64, 64, 128, 1, 128), (2, 256, 64, 64, 3, 64), (18, 4, 512, 512, 3, 64), (5, 40, 128, 192, 3, 64)])
def test_reverse_tile_order_is_exact(mode, N, H, Cin, Cout, k, BN):
    """tile_reverse changes the order in which a launch visits its tiles, not what a tile computes: outputs and the
    (per-tile, fixed-order) BN-gradient sums are bit-identical."""
    from pix2latent_b200 import native
    from test_conv_gemm_gpu import run_conv, pack_w
    dt = native.act_dtype()
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(N, H, H, Cin, device=dev).to(dt)
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).to(dt)
    a = torch.rand(N, Cout, device=dev) + 0.5
    s = torch.randn(N, Cout, device=dev) * 0.1
    bias = torch.randn(Cout, device=dev) * 0.1
    saved = torch.relu(torch.randn(N, H, H, Cout, device=dev)).to(dt)
    out = {}
    for rev in (0, 1):
        common = dict(A=x, A_N=N, A_H=H, A_W=H, A_C=Cin, Cin=Cin, B=pack_w(w), Cout=Cout, kh=k, kw=k, pad_h=k // 2, pad_w=k // 2,
                      NI=N, H=H, W=H, BN=BN, mode=mode, tile_reverse=rev)
        if mode == 0:
            raw = torch.zeros(N, H, H, Cout, device=dev, dtype=dt)
            act = torch.zeros_like(raw)
            run_conv(bias=bias, aff_a=a, aff_s=s, aff_stride=Cout, relu=1, raw=raw, raw_C=Cout, act=act, act_C=Cout, **common)
            out[rev] = (raw, act)
        else:
            dx = torch.zeros(N, H, H, Cout, device=dev, dtype=dt)
            st0, st1 = torch.zeros(N, Cout, device=dev), torch.zeros(N, Cout, device=dev)
            run_conv(saved=saved, saved_C=Cout, stat0=st0, stat1=st1, stat_stride=Cout, aff_a=a, aff_stride=Cout, dx=dx, dx_C=Cout,
                     **common)
            out[rev] = (dx, st0, st1)
    for u, v in zip(out[0], out[1]):
        assert torch.equal(u, v)


def test_generator_serpentine_is_exact(problem):
    cfg, orc = problem
    io = _model_io(cfg, orc, b=5)
    with options(serpentine=0):
        base = _run_model(cfg, orc, *io)
    with options(serpentine=1):
        serp = _run_model(cfg, orc, *io)
    for u, v in zip(base, serp):
        assert torch.equal(u, v)
