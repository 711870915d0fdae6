"""StyleGAN2 at the benchmarked configurations (BASELINE.json configs[2] and [4]) against the oracle's rosinality
generator (oracle/stylegan2.py, fp32 on the GPU, TF32 off) with the projection loss INCLUDING the LPIPS term:

  * LSUN-cars 512x512, 9 candidates (one chunk of /root/reference examples/invert_stylegan2_cars_cma.py:110's population
    of 22), loss restricted to rows 64:-64 (weight = loss_mask, examples/invert_stylegan2_cars_basincma.py:39-42):
    teacher-forced replay of an oracle Adam trajectory (z, lr 0.05, Clamp(2)) with the SAME per-layer noise on both
    sides — per-step loss, dz at step 0, final image and final LPIPS term;
  * FFHQ 1024x1024, 2 candidates: one forward + backward (the 32-channel top level runs zero-padded to 64 channels).

Reference path: /root/reference pix2latent/model/stylegan2.py:110-119 (forward_z, clamp) under
pix2latent/optimizer/closure.py:51-58. Tolerances (16-bit operands, fp32 accumulation): per-step |dloss| <= 2e-3 (1 + |loss|),
dz cosine >= 0.99, final LPIPS |delta| <= 1e-3, image mean-abs <= 5e-3."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-300)).item()


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _target(res, band):
    g = torch.Generator().manual_seed(1)
    t = torch.tanh(0.5 * torch.randn(1, 3, res, res, generator=g))
    t = torch.nn.functional.avg_pool2d(t, 8)
    t = torch.nn.functional.interpolate(t, size=(res, res), mode="bilinear", align_corners=False)[0]
    w = torch.zeros(3, res, res)
    if band:
        w[:, res // 8:-(res // 8), :] = 1.0
    else:
        w[:] = 1.0
    return t.cuda(), w.cuda()


def _world(size):
    from oracle import lpips as olp, stylegan2 as osg
    from pix2latent_b200.native import NativeLPIPS
    from pix2latent_b200.model.stylegan2 import StyleGAN2
    from test_biggan_gpu import lpips_native_state
    _setup()
    orc = osg.make_stylegan2(size, None, seed=0).cuda()
    for p in orc.parameters():
        p.requires_grad_(False)
    lp = olp.make_lpips("alex", seed=0).cuda()
    model = StyleGAN2(state_dict=orc.model.state_dict(), size=size)
    nl = NativeLPIPS("alex", lpips_native_state(lp))
    return orc, lp, model, nl


def test_cars512_replay_with_lpips():
    from oracle import lpips as olp
    from pix2latent_b200.native import sg2_step
    orc, lp, model, nl = _world(512)
    ref_loss = olp.ProjectionLoss(lpips_module=lp)
    ref_per = olp.PerceptualLoss(lpips_module=lp)
    b, steps = 9, 8
    target, weight = _target(512, band=True)
    tgt = nl.make_target(target, weight, weight, 1, 1.0, 10.0)        # weight = loss_mask, as the example registers them
    tgt_per = nl.make_target(target, weight, weight, 1, 0.0, 1.0)
    T, Wt = target[None].expand(b, -1, -1, -1), weight[None].expand(b, -1, -1, -1)
    torch.manual_seed(2)
    z = torch.fmod(torch.randn(b, 512), 2.0).cuda().requires_grad_(True)
    opt = torch.optim.Adam([z], lr=0.05)
    gen = torch.Generator(device="cuda").manual_seed(3)
    errs = []
    for k in range(steps + 1):
        with torch.no_grad():
            z.clamp_(-2, 2)
        noise = [torch.randn(s, device="cuda", generator=gen) for s in orc.model.noise_shapes(b)]
        l_nat, dz, img = sg2_step(model.native, nl, tgt, z.detach(), noise, True, 1.0 / b)
        opt.zero_grad()
        ref_img = orc(z, noise)
        l_ref = ref_loss(ref_img, T, Wt, Wt)
        l_ref.mean().backward()
        err = ((l_nat - l_ref.detach()).abs() / (1 + l_ref.detach().abs())).max().item()
        errs.append(err)
        if k == 0:
            c0 = cos(dz, z.grad)
            print("cars-512 step 0: cos dz %.5f  |dz| ratio %.4f" % (c0, (dz.norm() / z.grad.norm()).item()))
            assert c0 >= 0.9995   # measured 1.00000
        if k == steps:
            break
        opt.step()
    print("cars-512 per-step max |dloss|/(1+|loss|):", " ".join("%.1e" % e for e in errs))
    assert max(errs) <= 4e-4   # measured 1.7e-4
    with torch.no_grad():
        ref_p = ref_per(ref_img, T, Wt, Wt)
    nat_p = tgt_per.loss_forward(img, False)
    d = (img - ref_img.detach()).abs()
    print("cars-512 final image: max-abs %.3e mean-abs %.3e; final LPIPS max |d| %.2e (values %.4f .. %.4f)"
          % (d.max().item(), d.mean().item(), (nat_p - ref_p).abs().max().item(), ref_p.min().item(), ref_p.max().item()))
    assert (nat_p - ref_p).abs().max().item() <= 1e-4   # north star: 1e-3; measured 1.4e-5
    assert d.mean().item() <= 6e-4   # measured 2.6e-4


def test_ffhq1024_forward_backward():
    from oracle import lpips as olp
    from pix2latent_b200.native import sg2_step
    orc, lp, model, nl = _world(1024)
    assert model.channels[1024] == 64            # the 32-channel level runs zero-padded
    ref_loss = olp.ProjectionLoss(lpips_module=lp)
    b = 2
    target, weight = _target(1024, band=False)
    tgt = nl.make_target(target, weight, None, 1, 1.0, 10.0)
    torch.manual_seed(4)
    z = torch.fmod(torch.randn(b, 512), 2.0).cuda().requires_grad_(True)
    noise = [torch.randn(s, device="cuda") for s in orc.model.noise_shapes(b)]
    l_nat, dz, img = sg2_step(model.native, nl, tgt, z.detach(), noise, True, 1.0 / b)
    ref_img = orc(z, noise)
    l_ref = ref_loss(ref_img, target[None].expand(b, -1, -1, -1), weight[None].expand(b, -1, -1, -1))
    l_ref.mean().backward()
    d = (img - ref_img.detach()).abs()
    print("ffhq-1024: loss native %s oracle %s; image max-abs %.3e mean-abs %.3e; cos dz %.5f |dz| ratio %.4f"
          % (l_nat.tolist(), l_ref.tolist(), d.max().item(), d.mean().item(), cos(dz, z.grad), (dz.norm() / z.grad.norm()).item()))
    assert ((l_nat - l_ref.detach()).abs() / (1 + l_ref.detach().abs())).max().item() <= 2e-4   # measured 5e-5
    assert cos(dz, z.grad) >= 0.9995 and d.mean().item() <= 6e-4   # measured 0.99999, 2.8e-4
