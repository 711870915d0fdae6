"""Product API with the native StyleGAN2 model (the reference's invert_stylegan2_cars_* setup at a
reduced size): replay of recorded hook / noise draws through the product's closure.step versus the
oracle step; optimizer loop under nn.DataParallel wrapping as the examples do
(examples/invert_stylegan2_cars_basincma.py:50-51, 61-96)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _lpips_state(lp):
    sd = {}
    for k, sl in enumerate(lp.net.slices):
        for name, mod in sl.named_children():
            if hasattr(mod, "weight"):
                sd["net.slice%d.%s.weight" % (k + 1, name)] = mod.weight
                sd["net.slice%d.%s.bias" % (k + 1, name)] = mod.bias
    for k, l in enumerate(lp.lins):
        sd["lin%d.weight" % k] = l
    return sd


@pytest.fixture(scope="module")
def world():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from oracle import lpips as olp, stylegan2 as osg
    from pix2latent_b200.loss_functions import ProjectionLoss
    from pix2latent_b200.model.stylegan2 import StyleGAN2
    orc = osg.make_stylegan2(32, osg.TINY_CHANNELS, seed=0).cuda()
    lp = olp.make_lpips("alex", seed=0).cuda()
    ref_loss = olp.ProjectionLoss(lpips_module=lp)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = StyleGAN2(state_dict=orc.model.state_dict(), size=32, channels=dict(osg.TINY_CHANNELS))
    loss = ProjectionLoss(lpips_state_dict=_lpips_state(lp))
    g = torch.Generator().manual_seed(3)
    target = torch.tanh(torch.randn(3, 32, 32, generator=g)).cuda()
    weight = torch.ones(3, 32, 32).cuda()
    mask = torch.zeros(3, 32, 32)
    mask[:, 4:-4, :] = 1
    return orc, ref_loss, model, loss, target, weight, mask.cuda()


def _vm(target, weight, mask):
    from pix2latent_b200 import VariableManager
    import pix2latent_b200.distribution as dist
    import pix2latent_b200.utils.function_hooks as hook
    vm = VariableManager(device="cuda")
    vm.register(variable_name="z", shape=(512,), distribution=dist.TruncatedNormalModulo(), var_type="input",
                learning_rate=0.05, hook_fn=hook.Compose(hook.NormalPerturb(sigma=0.05), hook.Clamp(trunc=2.0)))
    vm.register(variable_name="target", shape=(3, 32, 32), requires_grad=False, default=target, var_type="output")
    vm.register(variable_name="weight", shape=(3, 32, 32), requires_grad=False, default=weight, var_type="output")
    vm.register(variable_name="loss_mask", shape=(3, 32, 32), requires_grad=False, default=mask, var_type="output")
    return vm


class _Replay(nn.Module):
    """Oracle generator fed with the recorded noise tensors, chunk by chunk."""

    def __init__(self, orc, tape):
        super().__init__()
        self.orc, self.tape = orc, tape

    def forward(self, z):
        return self.orc(z, self.tape.pop(0))


def test_replay_step_matches_oracle(world):
    from oracle import closure as oc
    from pix2latent_b200.optimizer.closure import _native_sg2_pair, step
    orc, ref_loss, model, loss, target, weight, mask = world
    n, chunk, steps = 5, 2, 3
    g = torch.Generator(device="cuda").manual_seed(11)
    tape = [[torch.randn(s, device="cuda", generator=g) for s in orc.model.noise_shapes(bsz)]
            for _ in range(steps) for bsz in (2, 2, 1)]
    tape_a, tape_b = list(tape), list(tape)
    model.draw_noise = lambda b, device: tape_a.pop(0)
    torch.manual_seed(31)
    v_nat = _vm(target, weight, mask).initialize(n)
    assert _native_sg2_pair(model, v_nat, loss)
    torch.manual_seed(31)
    v_ref = _vm(target, weight, mask).initialize(n)
    ref_model = _Replay(orc, tape_b)
    for k in range(steps):
        torch.manual_seed(100 + k)   # the NormalPerturb hook draws
        _, l_nat, _ = step(model, v_nat, loss, optimize=True, max_batch_size=chunk)
        torch.manual_seed(100 + k)
        _, l_ref, _ = oc.step(ref_model, v_ref, ref_loss, optimize=True, max_batch_size=chunk)
        l_nat, l_ref = np.array(l_nat), np.array(l_ref)
        print("step", k, l_nat, l_ref)
        assert np.abs(l_nat - l_ref).max() < 3e-2 * (1 + np.abs(l_ref).max())
    zn, zr = torch.stack(v_nat.input.z.data), torch.stack(v_ref.input.z.data)
    assert (zn - zr).abs().mean().item() < 0.05
    del model.draw_noise


def test_optimizers_with_dataparallel_wrapper(world):
    from pix2latent_b200.optimizer import CMAOptimizer, GradientOptimizer
    orc, ref_loss, model, loss, target, weight, mask = world
    wrapped = nn.DataParallel(model)
    torch.manual_seed(5)
    opt = GradientOptimizer(wrapped, _vm(target, weight, mask), loss, max_batch_size=3)
    variables, outs, l = opt.optimize(num_samples=4, grad_steps=3)
    assert len(l[0][1]["loss"]) == 4 and outs[0].shape[0] == 3 and np.isfinite(l[0][1]["loss"]).all()
    vm = _vm(target, weight, mask)
    vm.edit_variable("z", {"grad_free": True})
    copt = CMAOptimizer(wrapped, vm, loss, max_batch_size=9)
    copt.cma_seed = 1
    variables, outs, l = copt.optimize(meta_steps=2, grad_steps=2)
    assert copt.num_samples == 22 and len(l[0][1]["loss"]) == 22  # 4 + floor(3 ln 512)
