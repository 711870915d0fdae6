"""Device-resident inner loop on the GPU,
through the C-ABI:
  * the Adam kernel against torch.optim.Adam with one param group per latent tensor
    (variable_manager.py:231-238) on the same gradients — fp32 round-off level;
  * a fused run of K steps against the SAME kernels driven step by step (p2l_biggan_step + p2l_adam_update, i.e. the
    product's per-step path): the step is bitwise reproducible (no atomics anywhere on the path), so the two runs are
    IDENTICAL — losses, tracked latents, final latents, image;
  * CUDA-graph replay against plain launches of the same loop: identical;
  * two fused calls of 3 steps against one of 6 (the Adam state carries over): identical;
  * the product API (GradientOptimizer / BasinCMAOptimizer) with and without the fused path: identical."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)


@pytest.fixture(scope="module")
def world():
    import test_step_gpu as ts
    ts._setup()
    import make_golden as mg
    from oracle import lpips as olp
    from pix2latent_b200.loss_functions import ProjectionLoss
    from pix2latent_b200.model import BigGAN
    cfg, orc, target, weight = mg.problem()
    orc = orc.cuda()
    lp = olp.make_lpips("alex", seed=0).cuda()
    model = BigGAN(config=ts._product_cfg(cfg), state_dict=orc.state_dict())
    loss = ProjectionLoss(lpips_state_dict=ts._lpips_state(lp))
    return cfg, orc, model, loss, target.cuda(), weight.cuda()


def test_adam_kernel_matches_torch_adam():
    from pix2latent_b200 import native
    torch.manual_seed(0)
    b, zd, cd = 7, 128, 128
    z = torch.randn(b, zd, device="cuda")
    c = torch.randn(b, cd, device="cuda") * 0.1
    zt = [z[i].clone().requires_grad_(True) for i in range(b)]
    ct = [c[i].clone().requires_grad_(True) for i in range(b)]
    opt = torch.optim.Adam([{"params": t, "lr": 0.05} for t in zt] + [{"params": t, "lr": 0.01} for t in ct])
    cfg = native.adam_config(0.05, 0.01)
    state = native.AdamState(b, zd, cd, z.device)
    for k in range(12):
        dz = torch.randn(b, zd, device="cuda") * (10.0 ** (-(k % 4)))
        dc = torch.randn(b, cd, device="cuda") * 1e-3
        for i in range(b):
            zt[i].grad = dz[i].clone()
            ct[i].grad = dc[i].clone()
        opt.step()
        native.adam_update(z, c, dz, dc, cfg, state)
    torch.cuda.synchronize()
    assert state.step_count() == 12
    assert (z - torch.stack(zt).detach()).abs().max().item() < 2e-6
    assert (c - torch.stack(ct).detach()).abs().max().item() < 2e-6
    mz, vz, mc, vc = state.moments()
    assert torch.allclose(mz[3], opt.state[zt[3]]["exp_avg"], rtol=1e-4, atol=1e-8)
    assert torch.allclose(vc[5], opt.state[ct[5]]["exp_avg_sq"], rtol=1e-4, atol=1e-12)


def _vm_cma(model, target, weight):
    import make_golden as mg
    import pix2latent_b200.distribution as dist
    import pix2latent_b200.utils.function_hooks as hook
    from pix2latent_b200 import VariableManager
    vm = VariableManager(device="cuda")
    mg.register(vm, hook, dist, model, target, weight, True)
    return vm


def _start(orc, b, seed):
    torch.manual_seed(seed)
    z = (torch.fmod(torch.randn(b, 128), 2.0) * 1.2).cuda()  # some entries beyond the clamp bound
    c = orc.get_class_embedding(3).repeat(b, 1).clone().cuda()
    return z.contiguous(), c.contiguous()


def test_fused_loop_equals_step_by_step(world):
    from pix2latent_b200 import native
    cfg, orc, model, loss, target, weight = world
    b, K, trunc = 5, 6, 2.0
    tgt = loss.prepared_target(target, weight)
    dloss = torch.tensor([1 / 3, 1 / 3, 1 / 3, 1 / 2, 1 / 2], device="cuda")
    z0, c0 = _start(orc, b, 31)
    # ---- step by step: the per-step product path (closure._step_native) spelled out: Clamp hook, fused step, Adam kernel
    z, c = z0.clone(), c0.clone()
    st = native.AdamState(b, 128, 128, z.device)
    acfg = native.adam_config(0.05, 0.01)
    ref_losses, ref_z = [], []
    for k in range(K):
        ref_z.append(z.clone())
        z.clamp_(-trunc, trunc)
        l, dz, dc, img_ref = native.biggan_step(model.native, loss.native_lpips(), tgt, z, c, True, 1.0, dloss=dloss)
        native.adam_update(z, c, dz, dc, acfg, st)
        ref_losses.append(l.clone())
    ref_losses = torch.stack(ref_losses)
    img_ref = img_ref.clone()
    # ---- fused, with and without the graph
    for use_graph in (False, True):
        zf, cf = z0.clone(), c0.clone()
        r = native.biggan_optimize(model.native, loss.native_lpips(), tgt, zf, cf, K, native.adam_config(0.05, 0.01, clamp_z=trunc),
                                   dloss=dloss, track=True, use_graph=use_graph)
        torch.cuda.synchronize()
        assert r["graph"] == use_graph, "CUDA-graph capture of the step was refused" if use_graph else "?"
        assert r["state"].step_count() == K
        what = "graph replay" if use_graph else "plain launches"
        assert torch.equal(r["loss"], ref_losses), what
        assert torch.equal(r["z_hist"], torch.stack(ref_z)), what   # tracked inputs are recorded before the hook of each step
        assert torch.equal(zf, z) and torch.equal(cf, c), what
        assert torch.equal(r["img"], img_ref), what                  # the returned image is the last forward's
        assert torch.equal(r["state"].mv, st.mv), what


def test_fused_step_matches_torch_adam_path(world):
    """The same run with torch.optim.Adam over 2n per-sample param groups (what the reference's optimizer is): agrees to
    Adam's own round-off over the first steps (the two differ in the ORDER of Adam's fp32 operations only)."""
    from pix2latent_b200 import native
    cfg, orc, model, loss, target, weight = world
    b, trunc = 4, 2.0
    tgt = loss.prepared_target(target, weight)
    z0, c0 = _start(orc, b, 33)
    zt = [z0[i].clone().requires_grad_(True) for i in range(b)]
    ct = [c0[i].clone().requires_grad_(True) for i in range(b)]
    opt = torch.optim.Adam([{"params": t, "lr": 0.05} for t in zt] + [{"params": t, "lr": 0.01} for t in ct])
    z, c = z0.clone(), c0.clone()
    r = native.biggan_optimize(model.native, loss.native_lpips(), tgt, z, c, 2, native.adam_config(0.05, 0.01, clamp_z=trunc),
                               grad_scale=0.5, track=True, use_graph=False)
    for k in range(2):
        for t in zt:
            t.data.clamp_(-trunc, trunc)
        l, dz, dc, _ = native.biggan_step(model.native, loss.native_lpips(), tgt, torch.stack(zt).detach(), torch.stack(ct).detach(),
                                          True, 0.5)
        if k == 0:
            assert torch.equal(l, r["loss"][0])
        for i in range(b):
            zt[i].grad, ct[i].grad = dz[i], dc[i]
        opt.step()
        if k == 0:
            # first update: |dz| = lr up to Adam's round-off wherever the gradient is not at noise level
            d = (r["z_hist"][1] - torch.stack(zt).detach()).abs()
            assert (d > 1e-4).float().mean().item() < 0.01 and d.mean().item() < 1e-3


def test_fused_loop_state_carries_over(world):
    """two fused calls of 3 steps == one of 6 (Adam moments and step count live in the state), bit for bit."""
    from pix2latent_b200 import native
    cfg, orc, model, loss, target, weight = world
    tgt = loss.prepared_target(target, weight)
    z0, c0 = _start(orc, 3, 32)
    cfgA = native.adam_config(0.05, 0.01, clamp_z=2.0)
    za, ca = z0.clone(), c0.clone()
    ra = native.biggan_optimize(model.native, loss.native_lpips(), tgt, za, ca, 6, cfgA, grad_scale=1 / 3, use_graph=False)
    zb, cb = z0.clone(), c0.clone()
    r1 = native.biggan_optimize(model.native, loss.native_lpips(), tgt, zb, cb, 3, cfgA, grad_scale=1 / 3, use_graph=False)
    r2 = native.biggan_optimize(model.native, loss.native_lpips(), tgt, zb, cb, 3, cfgA, state=r1["state"], grad_scale=1 / 3,
                                use_graph=True, track=True)
    torch.cuda.synchronize()
    assert r2["state"].step_count() == 6
    assert torch.equal(torch.cat([r1["loss"], r2["loss"]]), ra["loss"])
    assert torch.equal(za, zb) and torch.equal(ca, cb)
    assert torch.equal(r2["state"].mv, ra["state"].mv)


def test_product_api_fused_vs_per_step(world):
    """GradientOptimizer and BasinCMAOptimizer: the SAME trajectory with the device-resident loop on and off."""
    import test_step_gpu as ts
    from pix2latent_b200.optimizer import BasinCMAOptimizer, GradientOptimizer
    cfg, orc, model, loss, target, weight = world
    out = {}
    for fused in (False, True):
        torch.manual_seed(40)
        opt = GradientOptimizer(model, ts._vm(model, target, weight, "cuda"), loss, max_batch_size=2)
        opt.fuse_inner_loop = fused
        v, outs, losses = opt.optimize(num_samples=5, grad_steps=6)
        assert opt.fused_calls == (1 if fused else 0)
        adam_state = [v.opt.state[t]["exp_avg"].clone() for t in v.input.z.data]
        assert all(int(v.opt.state[t]["step"]) == 6 for t in v.input.z.data)   # handed back to the torch optimizer
        out[fused] = (torch.stack(v.input.z.data).detach().clone(), np.array(losses[0][1]["loss"]), outs[0].clone(),
                      [t.clone() for t in opt.tracked["z"]], adam_state)
    (z0, l0, o0, t0, a0), (z1, l1, o1, t1, a1) = out[False], out[True]
    print("GradientOptimizer fused vs per-step: loss", l0, l1)
    assert np.array_equal(l0, l1)
    assert torch.equal(z0, z1) and torch.equal(o0, o1)
    assert len(t0) == len(t1) == 6 and all(torch.equal(u, w) for u, w in zip(t0, t1))
    assert all(torch.equal(u, w) for u, w in zip(a0, a1))
    res = {}
    for fused in (False, True):
        torch.manual_seed(41)
        np.random.seed(41)
        opt = BasinCMAOptimizer(model, _vm_cma(model, target, weight), loss, max_batch_size=9)
        opt.cma_seed = 7
        opt.fuse_inner_loop = fused
        v, outs, losses = opt.optimize(meta_steps=2, grad_steps=3, last_grad_steps=4)
        assert opt.fused_calls == (3 if fused else 0)
        res[fused] = np.array(losses[0][1]["loss"])
    print("BasinCMA fused vs per-step: final losses", res[False], res[True])
    assert np.array_equal(res[False], res[True])
