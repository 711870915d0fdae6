"""Bitwise reproducibility of the inner step: no kernel on the path accumulates with atomics
— BN-gradient sums, loss sums and the StyleGAN2 style / demodulation reductions are per-block partials summed in a
fixed order — so
  * the same step run twice gives identical bits (loss, dz, dc, image), also from a freshly built model / target;
  * a candidate's result does not depend on which other candidates share its launch: a batch evaluated whole equals the
    same candidates evaluated in two separate calls (what candidate sharding over GPUs does), bit for bit."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)


@pytest.fixture(scope="module")
def world():
    import make_golden as mg
    import test_step_gpu as ts
    from oracle import lpips as olp
    ts._setup()
    cfg, orc, target, weight = mg.problem()
    lp = olp.make_lpips("alex", seed=0)
    return cfg, orc.cuda(), ts._lpips_state(lp.cuda()), target.cuda(), weight.cuda()


def _build(world):
    import test_step_gpu as ts
    from pix2latent_b200.loss_functions import ProjectionLoss
    from pix2latent_b200.model import BigGAN
    cfg, orc, lp_sd, target, weight = world
    model = BigGAN(config=ts._product_cfg(cfg), state_dict=orc.state_dict())
    loss = ProjectionLoss(lpips_state_dict=dict(lp_sd))   # a new dict: a new native LPIPS with fresh plans
    return model, loss, loss.prepared_target(target.clone(), weight.clone())


def _inputs(orc, b, seed=5):
    torch.manual_seed(seed)
    z = torch.fmod(torch.randn(b, 128), 2.0).cuda()
    c = (orc.get_class_embedding(3).repeat(b, 1) + 0.01 * torch.randn(b, 128, device="cuda")).contiguous()
    dloss = (torch.rand(b, device="cuda") + 0.5) / b
    return z, c, dloss


def test_same_step_twice_same_bits(world):
    from pix2latent_b200 import native
    orc = world[1]
    z, c, dloss = _inputs(orc, 6)
    outs = []
    for fresh in (True, False, True):
        if fresh:
            model, loss, tgt = _build(world)
        r = native.biggan_step(model.native, loss.native_lpips(), tgt, z, c, True, 1.0, dloss=dloss)
        torch.cuda.synchronize()
        outs.append([t.clone() for t in r])
    for other in outs[1:]:
        for name, u, v in zip(("loss", "dz", "dc", "img"), outs[0], other):
            assert torch.equal(u, v), name


@pytest.mark.parametrize("split", [(2, 4), (1, 5), (3, 3), (9, 9), (18, 2)])
def test_candidate_does_not_depend_on_its_batch(world, split):
    from pix2latent_b200 import native
    orc = world[1]
    model, loss, tgt = _build(world)
    z, c, dloss = _inputs(orc, sum(split), seed=6)
    whole = [t.clone() for t in native.biggan_step(model.native, loss.native_lpips(), tgt, z, c, True, 1.0, dloss=dloss)]
    lo = 0
    for n in split:
        part = native.biggan_step(model.native, loss.native_lpips(), tgt, z[lo:lo + n].contiguous(), c[lo:lo + n].contiguous(),
                                  True, 1.0, dloss=dloss[lo:lo + n].contiguous())
        for name, u, v in zip(("loss", "dz", "dc", "img"), whole, part):
            assert torch.equal(u[lo:lo + n], v), "%s of candidates [%d, %d)" % (name, lo, lo + n)
        lo += n


def test_stylegan2_step_twice_same_bits():
    from oracle import lpips as olp, stylegan2 as osg
    from pix2latent_b200.native import NativeLPIPS, NativeStyleGAN2, sg2_step
    from test_biggan_gpu import lpips_native_state
    orc = osg.make_stylegan2(64, {4: 128, 8: 128, 16: 64, 32: 64, 64: 64}, seed=0).cuda()
    nat = NativeStyleGAN2(64, {4: 128, 8: 128, 16: 64, 32: 64, 64: 64}, orc.model.state_dict())
    nl = NativeLPIPS("alex", lpips_native_state(olp.make_lpips("alex", seed=0).cuda()))
    torch.manual_seed(3)
    b = 3
    z = torch.randn(b, 512, device="cuda")
    noise = [torch.randn(s, device="cuda") for s in orc.model.noise_shapes(b)]
    target = torch.tanh(torch.randn(3, 64, 64, device="cuda"))
    tgt = nl.make_target(target, None, None, 1, 1.0, 10.0)
    a = [t.clone() for t in sg2_step(nat, nl, tgt, z, noise, True, 1.0 / b)]
    b_ = sg2_step(nat, nl, tgt, z, noise, True, 1.0 / b)
    for name, u, v in zip(("loss", "dz", "img"), a, b_):
        assert torch.equal(u, v), name


@pytest.mark.parametrize("split", [(2, 3), (1, 4), (4, 1)])
def test_stylegan2_candidate_does_not_depend_on_its_batch(split):
    """What lets closure._step_native_sg2 run the reference's chunks (9/9/4, ...) as ONE physical batch with per-candidate
    1/b_chunk scales: loss, dz and image of a candidate are the same bits whatever batch it is evaluated in."""
    from oracle import lpips as olp, stylegan2 as osg
    from pix2latent_b200.native import NativeLPIPS, NativeStyleGAN2, sg2_step
    from test_biggan_gpu import lpips_native_state
    ch = {4: 128, 8: 128, 16: 64, 32: 64, 64: 64}
    orc = osg.make_stylegan2(64, ch, seed=0).cuda()
    nat = NativeStyleGAN2(64, ch, orc.model.state_dict())
    nl = NativeLPIPS("alex", lpips_native_state(olp.make_lpips("alex", seed=0).cuda()))
    torch.manual_seed(4)
    b = sum(split)
    z = torch.randn(b, 512, device="cuda")
    noise = [torch.randn(s, device="cuda") for s in orc.model.noise_shapes(b)]
    dloss = torch.tensor([1.0 / split[0]] * split[0] + [1.0 / split[1]] * split[1], device="cuda")
    tgt = nl.make_target(torch.tanh(torch.randn(3, 64, 64, device="cuda")), None, None, 1, 1.0, 10.0)
    whole = [t.clone() for t in sg2_step(nat, nl, tgt, z, noise, True, 1.0, dloss=dloss)]
    lo = 0
    for n in split:
        part = sg2_step(nat, nl, tgt, z[lo:lo + n].contiguous(), [t[lo:lo + n].contiguous() for t in noise], True, 1.0,
                        dloss=dloss[lo:lo + n].contiguous())
        for name, u, v in zip(("loss", "dz", "img"), whole, part):
            assert torch.equal(u[lo:lo + n], v), "%s of candidates [%d, %d)" % (name, lo, lo + n)
        # and the reference's form of the same chunk: scalar 1/b_chunk scale
        ref = sg2_step(nat, nl, tgt, z[lo:lo + n].contiguous(), [t[lo:lo + n].contiguous() for t in noise], True, 1.0 / n)
        assert torch.equal(ref[1], part[1]) and torch.equal(ref[0], part[0])
        lo += n
