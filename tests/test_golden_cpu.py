"""CPU checks against vectors produced by the REAL reference code (tests/golden/make_golden.py,
tests/golden/reference_cpu.npz):
  (a) the oracle restatement (oracle/closure.py, oracle/lpips.py losses) reproduces them;
  (b) the product's host code (VariableManager, closure.step autograd path, hooks, distribution,
      GradientOptimizer / BasinCMAOptimizer / CMAOptimizer) reproduces them when driven with the
      same model and loss callables.
Nothing here touches the GPU; the generator/LPIPS used are the oracle modules on CPU."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_golden as mg  # noqa: E402  (only its problem()/register() helpers; no reference import)

GOLD = np.load(os.path.join(HERE, "golden", "reference_cpu.npz"))
TOL = dict(rtol=2e-5, atol=2e-6)


@pytest.fixture(scope="module")
def prob():
    from oracle import lpips as olp
    cfg, model, target, weight = mg.problem()
    loss_fn = olp.ProjectionLoss(lpips_module=olp.make_lpips("alex", seed=0))
    return cfg, model, target, weight, loss_fn


def test_oracle_losses_match_reference(prob):
    from oracle import lpips as olp
    cfg, model, target, weight, loss_fn = prob
    img = torch.from_numpy(GOLD["loss_img"])
    mask = torch.zeros(3, 128, 128)
    mask[:, 16:-16] = 1
    t2, w2, m2 = (x[None].expand(2, -1, -1, -1) for x in (target, weight, mask))
    np.testing.assert_allclose(loss_fn(img, t2, w2).numpy(), GOLD["loss_w"], **TOL)
    np.testing.assert_allclose(loss_fn(img, t2, w2, m2).numpy(), GOLD["loss_wm"], **TOL)
    np.testing.assert_allclose(loss_fn(img, t2).view(2, -1).mean(1).numpy(), GOLD["loss_none"], **TOL)
    np.testing.assert_allclose(olp.ReconstructionLoss()(img, t2, w2).numpy(), GOLD["rec_w"], **TOL)
    np.testing.assert_allclose(loss_fn.ploss_fn(img, t2, w2).numpy(), GOLD["per_w"], **TOL)


def _spec(model, target, weight, grad_free, hook, dist):
    import torch.optim as optim
    return {
        "z": dict(shape=(128,), var_type="input", requires_grad=True, default=None,
                  distribution=dist.TruncatedNormalModulo(), optimizer=optim.Adam, learning_rate=0.05,
                  hook_fn=hook.Clamp(2.0), grad_free=grad_free),
        "c": dict(shape=(128,), var_type="input", requires_grad=True, default=model.get_class_embedding(3)[0],
                  distribution=None, optimizer=optim.Adam, learning_rate=0.01, hook_fn=None, grad_free=False),
        "target": dict(shape=(3, 128, 128), var_type="output", requires_grad=False, default=target,
                       distribution=None, optimizer=optim.Adam, learning_rate=0.05, hook_fn=None, grad_free=False),
        "weight": dict(shape=(3, 128, 128), var_type="output", requires_grad=False, default=weight,
                       distribution=None, optimizer=optim.Adam, learning_rate=0.05, hook_fn=None, grad_free=False),
    }


def test_oracle_step_matches_reference(prob):
    from oracle import closure as oc
    import pix2latent_b200.distribution as dist
    import pix2latent_b200.utils.function_hooks as hook
    cfg, model, target, weight, loss_fn = prob
    torch.manual_seed(21)
    variables = oc.initialize(_spec(model, target, weight, False, hook, dist), 3)
    np.testing.assert_array_equal(torch.stack(variables.input.z.data).detach().numpy(), GOLD["step_z0"])
    losses = []
    for _ in range(3):
        _, l, _ = oc.step(model, variables, loss_fn, optimize=True, max_batch_size=2)
        losses.append(np.array(l))
    o, l, _ = oc.step(model, variables, loss_fn, optimize=False, max_batch_size=2)
    losses.append(np.array(l))
    np.testing.assert_allclose(np.stack(losses), GOLD["step_losses"], **TOL)
    np.testing.assert_allclose(torch.stack(variables.input.z.data).detach().numpy(), GOLD["step_z"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(torch.stack(variables.input.c.data).detach().numpy(), GOLD["step_c"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(o.mean((1, 2, 3)).numpy(), GOLD["step_out_mean"], rtol=1e-4, atol=1e-6)


def _vm(model, target, weight, grad_free):
    from pix2latent_b200 import VariableManager
    import pix2latent_b200.distribution as dist
    import pix2latent_b200.utils.function_hooks as hook
    vm = VariableManager(device="cpu")
    mg.register(vm, hook, dist, model, target, weight, grad_free)
    return vm


def test_product_step_matches_reference(prob):
    from pix2latent_b200.optimizer.closure import step
    cfg, model, target, weight, loss_fn = prob
    torch.manual_seed(21)
    variables = _vm(model, target, weight, False).initialize(3)
    np.testing.assert_array_equal(torch.stack(variables.input.z.data).detach().numpy(), GOLD["step_z0"])
    losses = []
    for _ in range(3):
        _, l, _ = step(model, variables, loss_fn, optimize=True, max_batch_size=2)
        losses.append(np.array(l))
    o, l, misc = step(model, variables, loss_fn, optimize=False, max_batch_size=2)
    losses.append(np.array(l))
    assert misc == {} and o.shape == (3, 3, 128, 128) and len(l) == 3
    np.testing.assert_allclose(np.stack(losses), GOLD["step_losses"], **TOL)
    np.testing.assert_allclose(torch.stack(variables.input.z.data).detach().numpy(), GOLD["step_z"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(torch.stack(variables.input.c.data).detach().numpy(), GOLD["step_c"], rtol=1e-4, atol=1e-5)


def test_product_gradient_optimizer_matches_reference(prob):
    from pix2latent_b200.optimizer import GradientOptimizer
    cfg, model, target, weight, loss_fn = prob
    torch.manual_seed(22)
    opt = GradientOptimizer(model, _vm(model, target, weight, False), loss_fn, max_batch_size=2)
    variables, outs, loss = opt.optimize(num_samples=3, grad_steps=3)
    assert loss[0][0] == 3 and len(outs) == 1 and outs[0].shape[0] == 3
    np.testing.assert_allclose(np.array(loss[0][1]["loss"]), GOLD["grad_loss"], **TOL)
    np.testing.assert_allclose(torch.stack(variables.input.z.data).detach().numpy(), GOLD["grad_z"], rtol=1e-4, atol=1e-5)
    assert len(opt.tracked["z"]) == 3 and opt.tracked["z"][0].shape == (3, 128)


def test_product_basincma_matches_reference(prob):
    from pix2latent_b200.optimizer import BasinCMAOptimizer
    cfg, model, target, weight, loss_fn = prob
    torch.manual_seed(23)
    opt = BasinCMAOptimizer(model, _vm(model, target, weight, True), loss_fn, max_batch_size=9)
    opt.cma_seed = mg.CMA_SEED
    variables, outs, loss = opt.optimize(meta_steps=2, grad_steps=2, last_grad_steps=2)
    assert opt.num_samples == 18  # 4 + floor(3 ln 128)
    assert loss[0][0] == int(GOLD["basin_total_steps"])
    assert len(opt.tracked["z"]) == int(GOLD["basin_tracked_len"])
    np.testing.assert_allclose(np.array(loss[0][1]["loss"]), GOLD["basin_loss"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(torch.stack(variables.input.z.data).detach().numpy(), GOLD["basin_z"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(np.array(list(opt.cma_optimizers.values())[0].mean()), GOLD["basin_cma_mean"],
                               rtol=1e-3, atol=1e-4)


def test_product_cma_matches_reference(prob):
    from pix2latent_b200.optimizer import CMAOptimizer
    cfg, model, target, weight, loss_fn = prob
    torch.manual_seed(24)
    opt = CMAOptimizer(model, _vm(model, target, weight, True), loss_fn, max_batch_size=9)
    opt.cma_seed = mg.CMA_SEED
    variables, outs, loss = opt.optimize(meta_steps=2, grad_steps=2)
    np.testing.assert_allclose(np.array(loss[0][1]["loss"]), GOLD["cma_loss"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(torch.stack(variables.input.z.data).detach().numpy(), GOLD["cma_z"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(np.array(list(opt.cma_optimizers.values())[0].mean()), GOLD["cma_mean"], rtol=1e-3, atol=1e-4)
