"""Generate golden vectors by running the REAL reference code (/root/reference/pix2latent) on CPU.

The reference cannot be imported as shipped in this image: its third-party arithmetic
(``pytorch_pretrained_biggan``, ``lpips``) and host libraries (``cma``, ``nevergrad``,
``easydict``) are not installed and there is no network, and it hard-codes
``.cuda()`` (F7). This script therefore
  * stubs the missing modules: ``lpips.LPIPS`` -> oracle/lpips.py, ``cma`` -> the package's minimal
    CMA-ES with a fixed seed, ``easydict`` -> an attribute dict, ``nevergrad`` /
    ``pytorch_pretrained_biggan`` -> empty modules;
  * makes ``Tensor.cuda()`` / ``Module.cuda()`` the identity;
  * then imports the reference package UNMODIFIED and runs its own ``VariableManager``,
    ``closure.step``, ``ProjectionLoss``, ``GradientOptimizer``, ``BasinCMAOptimizer`` and
    ``CMAOptimizer`` with the oracle generator (oracle/biggan.py, tiny128 config) as the model.

What the vectors pin (byte-for-byte reference code): mini-batch chunking and the 1/b gradient
scale, per-sample leaves + per-tensor Adam param groups, hook order, the loss algebra
(sum(|t-o|W)/sum(W) + 10 * sum(map W)/sum(W)), and the three optimizer loops incl. the Baldwinian
``tell``. What they do NOT pin: the third-party generator / LPIPS arithmetic itself (restated in
oracle/, "parity unpinned" — no published vectors exist).

Run here (CPU container):  python tests/golden/make_golden.py   -> tests/golden/reference_cpu.npz
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import biggan as obg  # noqa: E402
from oracle import lpips as olp  # noqa: E402
from pix2latent_b200.optimizer import _minicma  # noqa: E402

CMA_SEED = 123


def install_stubs():
    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            d = dict(d or {}, **kw)
            for k, v in d.items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __setattr__ = __setitem__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    m = types.ModuleType("easydict"); m.EasyDict = EasyDict; sys.modules["easydict"] = m
    m = types.ModuleType("cma")
    m.CMAEvolutionStrategy = lambda mu, sigma, opts=None: _minicma.CMAEvolutionStrategy(
        mu, sigma, dict(opts or {}, seed=CMA_SEED))
    sys.modules["cma"] = m
    m = types.ModuleType("nevergrad")
    m.optimizers = types.SimpleNamespace(registry={})
    sys.modules["nevergrad"] = m
    sys.modules["pytorch_pretrained_biggan"] = types.ModuleType("pytorch_pretrained_biggan")
    m = types.ModuleType("lpips")
    m.LPIPS = lambda net="alex", spatial=True, **kw: olp.make_lpips(net, seed=0, spatial=spatial)
    sys.modules["lpips"] = m
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


def problem(seed=0):
    """Synthetic inversion problem shared with tests/test_golden_cpu.py."""
    cfg = obg.BigGANConfig.tiny128()
    model = obg.make_biggan(cfg, seed=0)
    for p in model.parameters():
        p.requires_grad_(False)  # keeps the CPU run fast; weight grads do not affect dz/dc
    g = torch.Generator().manual_seed(11)
    target = torch.tanh(torch.randn(3, 128, 128, generator=g))
    weight = torch.full((3, 128, 128), 0.3)
    weight[:, 32:96, 32:96] = 1.0
    return cfg, model, target, weight


def register(vm, hook, dist, model, target, weight, grad_free):
    vm.register(variable_name="z", shape=(128,), grad_free=grad_free,
                distribution=dist.TruncatedNormalModulo(sigma=1.0, trunc=2.0), var_type="input",
                learning_rate=0.05, hook_fn=hook.Clamp(2.0))
    vm.register(variable_name="c", shape=(128,), default=model.get_class_embedding(3)[0], var_type="input",
                learning_rate=0.01)
    vm.register(variable_name="target", shape=(3, 128, 128), requires_grad=False, default=target, var_type="output")
    vm.register(variable_name="weight", shape=(3, 128, 128), requires_grad=False, default=weight, var_type="output")


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import pix2latent  # the reference package, unmodified
    from pix2latent import VariableManager
    from pix2latent.optimizer import BasinCMAOptimizer, CMAOptimizer, GradientOptimizer
    from pix2latent.optimizer.closure import step as ref_step
    import pix2latent.loss_functions as LF
    import pix2latent.utils.function_hooks as hook
    import pix2latent.distribution as dist
    assert pix2latent.__file__.startswith(REF)

    cfg, model, target, weight = problem()
    loss_fn = LF.ProjectionLoss()
    out = {}

    # ---- (1) closure.step: 3 samples, chunks of 2 -> (2, 1); 3 optimise steps + 1 eval-only step
    torch.manual_seed(21)
    vm = VariableManager()
    register(vm, hook, dist, model, target, weight, grad_free=False)
    variables = vm.initialize(3)
    out["step_z0"] = torch.stack(variables.input.z.data).detach().numpy().copy()
    losses = []
    for _ in range(3):
        _, l, _ = ref_step(model, variables, loss_fn, optimize=True, max_batch_size=2)
        losses.append(np.array(l))
    o, l, _ = ref_step(model, variables, loss_fn, optimize=False, max_batch_size=2)
    losses.append(np.array(l))
    out["step_losses"] = np.stack(losses)
    out["step_z"] = torch.stack(variables.input.z.data).detach().numpy()
    out["step_c"] = torch.stack(variables.input.c.data).detach().numpy()
    out["step_out_mean"] = o.mean((1, 2, 3)).numpy()

    # ---- (2) loss algebra on fixed images (with / without weight, with mask)
    g = torch.Generator().manual_seed(5)
    img = torch.tanh(torch.randn(2, 3, 128, 128, generator=g))
    mask = torch.zeros(3, 128, 128); mask[:, 16:-16] = 1
    t2, w2, m2 = (x[None].expand(2, -1, -1, -1) for x in (target, weight, mask))
    out["loss_img"] = img.numpy()
    out["loss_w"] = loss_fn(img, t2, w2).numpy()
    out["loss_wm"] = loss_fn(img, t2, w2, m2).numpy()
    out["loss_none"] = loss_fn(img, t2).view(2, -1).mean(1).numpy()
    out["rec_w"] = LF.ReconstructionLoss()(img, t2, w2).numpy()
    out["per_w"] = loss_fn.ploss_fn(img, t2, w2).numpy()

    # ---- (3) GradientOptimizer.optimize
    torch.manual_seed(22)
    vm = VariableManager()
    register(vm, hook, dist, model, target, weight, grad_free=False)
    opt = GradientOptimizer(model, vm, loss_fn, max_batch_size=2)
    variables, outs, loss = opt.optimize(num_samples=3, grad_steps=3)
    out["grad_loss"] = np.array(loss[0][1]["loss"])
    out["grad_z"] = torch.stack(variables.input.z.data).detach().numpy()

    # ---- (4) BasinCMAOptimizer.optimize (population 18 from the CMA default, chunks of 9)
    torch.manual_seed(23)
    vm = VariableManager()
    register(vm, hook, dist, model, target, weight, grad_free=True)
    opt = BasinCMAOptimizer(model, vm, loss_fn, max_batch_size=9)
    variables, outs, loss = opt.optimize(meta_steps=2, grad_steps=2, last_grad_steps=2)
    out["basin_loss"] = np.array(loss[0][1]["loss"])
    out["basin_total_steps"] = np.array(loss[0][0])
    out["basin_z"] = torch.stack(variables.input.z.data).detach().numpy()
    out["basin_cma_mean"] = np.array(list(opt.cma_optimizers.values())[0].mean())
    out["basin_tracked_len"] = np.array(len(opt.tracked["z"]))

    # ---- (5) CMAOptimizer.optimize
    torch.manual_seed(24)
    vm = VariableManager()
    register(vm, hook, dist, model, target, weight, grad_free=True)
    opt = CMAOptimizer(model, vm, loss_fn, max_batch_size=9)
    variables, outs, loss = opt.optimize(meta_steps=2, grad_steps=2)
    out["cma_loss"] = np.array(loss[0][1]["loss"])
    out["cma_z"] = torch.stack(variables.input.z.data).detach().numpy()
    out["cma_mean"] = np.array(list(opt.cma_optimizers.values())[0].mean())

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cpu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
