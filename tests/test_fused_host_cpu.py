"""Host logic of the device-resident inner loop on CPU: eligibility rules, and that
a run through ``_BaseOptimizer.grad_steps`` with the native call replaced by its oracle
(oracle/inner_loop.py) leaves the per-sample leaves, the torch optimizer state, the tracked inputs and
the reported losses exactly where the per-step path (closure.step + torch.optim.Adam) leaves them —
including a hand-over back to the per-step path afterwards."""
import numpy as np
import pytest
import torch
import torch.optim as optim

from oracle import inner_loop as oil


class ToyModel(torch.nn.Module):
    """img[b, 3, 2, 2] from (z, c): smooth and non-separable, so Adam trajectories are non-trivial."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.A = torch.nn.Parameter(torch.randn(12, 6, generator=g) * 0.7, requires_grad=False)
        self.B = torch.nn.Parameter(torch.randn(12, 4, generator=g) * 0.7, requires_grad=False)
        self.native = object()

    def forward(self, z, c):
        return torch.tanh(z @ self.A.T + (c @ self.B.T) * (1 + 0.1 * z.sum(1, keepdim=True))).view(-1, 3, 2, 2)


class ToyLoss:
    def __call__(self, out, target, weight=None, loss_mask=None):
        w = weight if weight is not None else torch.ones_like(target)
        return ((out - target).abs() * w).flatten(1).sum(1) / w.flatten(1).sum(1) + 0.3 * ((out - target) ** 2).flatten(1).mean(1)

    def prepared_target(self, target, weight, mask):
        return (target, weight)

    def native_lpips(self):
        return None


def _make(n, hook, seed=3):
    from pix2latent_b200 import VariableManager
    torch.manual_seed(seed)
    vm = VariableManager(device="cpu")
    vm.register("z", (6,), "input", learning_rate=0.05, hook_fn=hook)
    vm.register("c", (4,), "input", default=torch.randn(4) * 0.3, learning_rate=0.01)
    vm.register("target", (3, 2, 2), "output", requires_grad=False, default=torch.tanh(torch.randn(3, 2, 2)))
    vm.register("weight", (3, 2, 2), "output", requires_grad=False, default=torch.rand(3, 2, 2) + 0.2)
    torch.manual_seed(seed + 1)
    return vm


def _patch(monkeypatch, model, loss_fn):
    """Route native.biggan_optimize to the oracle and declare the toy pair 'native'."""
    from pix2latent_b200 import native
    from pix2latent_b200.optimizer import closure

    def fake_pair(m, vars, lf):
        return getattr(lf, "_fuse_ok", True) and set(vars.input.keys()) == {"z", "c"}

    def fake_optimize(gen, lp, tgt, z, c, steps, cfg, state=None, dloss=None, grad_scale=1.0, track=False,
                      want_img=True, use_graph=True):
        target, weight = tgt
        b = z.shape[0]

        def step_fn(zz, cc):
            img = model(z=zz, c=cc)
            return loss_fn(img, target[None].expand(b, -1, -1, -1), weight[None].expand(b, -1, -1, -1)), img

        mz, vz, mc, vc = state.moments()
        st = dict(m_z=mz, v_z=vz, m_c=mc, v_c=vc, step=state.step_count())
        r = oil.run(step_fn, z, c, steps, cfg.lr_z, cfg.lr_c, (cfg.beta1, cfg.beta2), cfg.eps, cfg.clamp_z, cfg.clamp_c,
                    dloss=dloss, grad_scale=grad_scale, state=st, track=track)
        state.counters[0] = st["step"]
        return {"loss": r["loss"], "z_hist": r["z_hist"], "c_hist": r["c_hist"], "img": r["img"], "state": state,
                "graph": False}

    monkeypatch.setattr(closure, "_native_pair", fake_pair)
    monkeypatch.setattr(native, "biggan_optimize", fake_optimize)


def _run(fused, monkeypatch, n=5, steps=7, extra=2, hook="clamp", max_batch_size=2):
    from pix2latent_b200.optimizer.base_optimizer import _BaseOptimizer
    from pix2latent_b200.optimizer import closure
    from pix2latent_b200.utils import function_hooks as hk
    model, loss_fn = ToyModel(), ToyLoss()
    h = {"clamp": hk.Clamp(0.6), "none": None, "compose": hk.Compose(hk.Clamp(0.9), hk.Clamp(0.6))}[hook]
    vm = _make(n, h)
    opt = _BaseOptimizer(model, vm, loss_fn, max_batch_size=max_batch_size)
    opt.fuse_inner_loop = fused
    _patch(monkeypatch, model, loss_fn)
    if not fused:
        monkeypatch.setattr(closure, "_native_pair", lambda *a: False)
    variables = vm.initialize(n)
    seen = []
    opt.grad_steps(variables, steps, lambda j: seen.append(j))
    assert seen == list(range(steps))
    first = dict(z=torch.stack(variables.input.z.data).detach().clone(), loss=np.array(opt.loss, dtype=np.float64),
                 out=opt.out.detach().clone(), fused_calls=opt.fused_calls)
    # hand over to the per-step path: the torch optimizer must carry on from the same moments / step count
    monkeypatch.setattr(closure, "_native_pair", lambda *a: False)
    for _ in range(extra):
        opt.step(variables, optimize=True)
    return first, torch.stack(variables.input.z.data).detach(), torch.stack(variables.input.c.data).detach(), opt


@pytest.mark.parametrize("hook", ["clamp", "none", "compose"])
def test_fused_run_equals_per_step_path(monkeypatch, hook):
    a, za, ca, oa = _run(False, monkeypatch, hook=hook)
    b, zb, cb, ob = _run(True, monkeypatch, hook=hook)
    assert a["fused_calls"] == 0 and b["fused_calls"] == 1
    assert torch.allclose(a["z"], b["z"], atol=2e-6), (a["z"] - b["z"]).abs().max()
    assert np.allclose(a["loss"], b["loss"], atol=2e-6)
    assert torch.allclose(a["out"], b["out"], atol=2e-6)
    # after two more per-step updates (Adam state handed back to torch)
    assert torch.allclose(za, zb, atol=5e-6) and torch.allclose(ca, cb, atol=5e-6)
    # tracked inputs: one entry per step, recorded before that step's hooks
    assert len(oa.tracked["z"]) == len(ob.tracked["z"]) == 9
    for ta, tb in zip(oa.tracked["z"], ob.tracked["z"]):
        assert torch.allclose(ta, tb, atol=5e-6)
    for ta, tb in zip(oa.tracked["c"], ob.tracked["c"]):
        assert torch.allclose(ta, tb, atol=5e-6)


def test_fused_run_continues_existing_adam_state(monkeypatch):
    """per-step, then fused, then per-step: the fused run imports the torch optimizer's moments."""
    from pix2latent_b200.optimizer.base_optimizer import _BaseOptimizer
    from pix2latent_b200.optimizer import closure
    from pix2latent_b200.utils import function_hooks as hk
    res = []
    for fused in (False, True):
        model, loss_fn = ToyModel(), ToyLoss()
        vm = _make(4, hk.Clamp(0.6))
        opt = _BaseOptimizer(model, vm, loss_fn, max_batch_size=3)
        opt.fuse_inner_loop = fused
        _patch(monkeypatch, model, loss_fn)
        variables = vm.initialize(4)
        monkeypatch.setattr(closure, "_native_pair", lambda *a: False)
        for _ in range(3):
            opt.step(variables, optimize=True)
        if fused:
            _patch(monkeypatch, model, loss_fn)
        opt.grad_steps(variables, 5)
        assert opt.fused_calls == (1 if fused else 0)
        res.append(torch.stack(variables.input.z.data).detach().clone())
    assert torch.allclose(res[0], res[1], atol=5e-6), (res[0] - res[1]).abs().max()


def test_fusable_rules(monkeypatch):
    from pix2latent_b200.optimizer.base_optimizer import _BaseOptimizer, _clamp_of
    from pix2latent_b200.utils import function_hooks as hk
    assert _clamp_of(None) == 0.0 and _clamp_of(hk.Clamp(2.0)) == 2.0
    assert _clamp_of(hk.NormalPerturb(0.1)) is False
    assert _clamp_of(hk.Compose(hk.NormalPerturb(0.1), hk.Clamp(2.0))) is False  # draws from the RNG: per-step path
    model, loss_fn = ToyModel(), ToyLoss()
    _patch(monkeypatch, model, loss_fn)

    def plan(vm, n_steps=5, **kw):
        opt = _BaseOptimizer(model, vm, loss_fn, **kw)
        return opt, opt._fusable(vm.initialize(3), n_steps)

    vm = _make(3, hk.Clamp(0.5))
    opt, p = plan(vm)
    assert p is not None and p["lr_z"] == 0.05 and p["lr_c"] == 0.01 and p["clamp_z"] == 0.5 and p["clamp_c"] == 0.0
    assert p["betas"] == (0.9, 0.999) and p["eps"] == 1e-8 and p["step0"] == 0
    assert plan(vm, n_steps=1)[1] is None                      # a single step gains nothing
    assert plan(vm, log=True)[1] is None                       # per-step collages need every image
    o2 = _BaseOptimizer(model, vm, loss_fn)
    o2.register_transform(lambda t, p: t, "t", "target")
    assert o2._fusable(vm.initialize(3), 5) is None            # transform search: per-candidate targets
    vm2 = _make(3, hk.NormalPerturb(0.1))
    assert plan(vm2)[1] is None
    vm3 = _make(3, None)
    vm3.edit_variable("weight", {"optimizer": optim.SGD})       # optimizer class of the LAST spec
    assert plan(vm3)[1] is None
    vm4 = _make(3, None)
    vm4.edit_variable("c", {"requires_grad": False})           # frozen class vector: lr 0
    _, p4 = plan(vm4)
    assert p4 is not None and p4["lr_c"] == 0.0 and p4["lr_z"] == 0.05
    o5 = _BaseOptimizer(model, vm, loss_fn)
    o5.fuse_inner_loop = False
    assert o5._fusable(vm.initialize(3), 5) is None
