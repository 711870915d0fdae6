"""Host-side behaviour of the kept API: VariableManager, hooks,
distributions, attribute dict, the minimal CMA-ES."""
import numpy as np
import pytest
import torch
import torch.optim as optim


def test_variable_manager_contract(capsys):
    from pix2latent_b200 import VariableManager
    vm = VariableManager(device="cpu")
    assert vm.register("z", (4,), "input", learning_rate=0.05) is True
    assert vm.register("z", (4,), "input") is False  # prints, does not raise (variable_manager.py:126-128)
    assert "already exists" in capsys.readouterr().out
    with pytest.raises(AssertionError):  # default/shape mismatch is an assert (:130-133)
        vm.register("bad", (3,), "input", default=torch.zeros(4))
    assert vm.register("c", (2,), "input", default=torch.ones(2), learning_rate=0.01, optimizer=optim.SGD)
    assert vm.register("t", (1, 2, 2), "output", requires_grad=False, default=torch.zeros(1, 2, 2))
    assert vm.edit_variable("nope", {}) is False
    assert vm.edit_variable("z", {"unknown": 1}) is False
    assert vm.edit_variable("z", {"learning_rate": 0.1}) is True
    v = vm.initialize(3)
    assert v.num_samples == 3 and len(v.input.z.data) == 3 and v.input.z.data[0].shape == (4,)
    # one param group per per-sample tensor, variable-specific lr; optimizer class of the LAST spec
    assert isinstance(v.opt, optim.Adam)  # 't' was registered last with the default Adam
    lrs = [g["lr"] for g in v.opt.param_groups]
    assert lrs == [0.1] * 3 + [0.01] * 3
    assert all(t.requires_grad for t in v.input.z.data) and not v.output.t.data[0].requires_grad
    # defaults are cloned per sample
    assert v.input.c.data[0].data_ptr() != v.input.c.data[1].data_ptr()
    vm.unregister("t", "missing")
    assert "t" not in vm.variable_info


def test_split_vars_shares_optimizer():
    from pix2latent_b200 import VariableManager
    from pix2latent_b200.variable_manager import split_vars
    vm = VariableManager(device="cpu")
    vm.register("z", (2,), "input")
    v = vm.initialize(5)
    parts = split_vars(v, 2)
    assert [p.num_samples for p in parts] == [2, 2, 1]
    assert all(p.opt is v.opt for p in parts)
    assert parts[1].input.z.data[0] is v.input.z.data[2]


def test_hooks_and_distribution():
    from pix2latent_b200 import distribution as dist
    from pix2latent_b200.utils import function_hooks as hook
    torch.manual_seed(0)
    x = dist.TruncatedNormalModulo(sigma=5.0, trunc=0.1)(1000, (8,))  # args ignored: sigma=1, trunc=2
    assert x.abs().max() < 2.0 and x.std() > 0.7
    ts = [torch.full((4,), 3.0), torch.full((4,), -5.0)]
    hook.Clamp(2.0)(ts)
    assert ts[0].tolist() == [2.0] * 4 and ts[1].tolist() == [-2.0] * 4
    torch.manual_seed(1)
    a = [torch.zeros(3), torch.zeros(3)]
    hook.NormalPerturb(0.5)(a)
    torch.manual_seed(1)
    ref = [0.5 * torch.randn(3), 0.5 * torch.randn(3)]  # one randn_like per tensor, in order
    assert torch.equal(a[0], ref[0]) and torch.equal(a[1], ref[1])
    b = [torch.tensor([1.0, 2.0, 3.0, 6.0])]
    hook.Normalize()(b)
    assert abs(b[0].mean().item()) < 1e-6 and abs(b[0].std().item() - 1) < 1e-6
    calls = []
    hook.Compose(lambda v: calls.append("a"), lambda v: calls.append("b"))([])
    assert calls == ["a", "b"]
    s = hook.ScheduledNormalPerturb(sigma=0.1, max_step=3)
    y = [torch.zeros(1000)]
    s(y); s(y); s(y)
    assert s.t == 3
    assert dist.normal(2.0)(4, (3,)).shape == (4, 3)
    assert dist.truncated_clamp_normal(1.0, 0.5)(100, (2,)).abs().max() <= 0.5


def test_save_variables_roundtrip(tmp_path):
    from pix2latent_b200 import VariableManager, save_variables
    vm = VariableManager(device="cpu")
    vm.register("z", (2,), "input")
    v = vm.initialize(2)
    v.loss = [[3, {"loss": [0.1, 0.2]}]]
    p = str(tmp_path / "vars.npy")
    save_variables(p, v)
    d = np.load(p, allow_pickle=True).item()
    assert "opt" not in d and len(d["input"]["z"]["data"]) == 2 and d["loss"][0][0] == 3


def test_minicma_defaults_and_convergence():
    from pix2latent_b200.optimizer._minicma import CMAEvolutionStrategy
    assert CMAEvolutionStrategy(np.zeros(128), 1.0).sp.popsize == 18   # README.md:74 of the reference
    assert CMAEvolutionStrategy(np.zeros(512), 1.0).sp.popsize == 22
    es = CMAEvolutionStrategy(np.full(6, 3.0), 1.0, {"seed": 1})
    for _ in range(120):
        X = es.ask()
        es.tell(X, [float(np.sum(x ** 2)) for x in X])
    assert np.sum(es.mean ** 2) < 1e-6


def test_cma_wrapper_scalar_hack():
    from pix2latent_b200.optimizer.base_cma_optimizer import CMA
    c = CMA([0.0], sigma=1.0, seed=3)
    x = c.ask()
    assert x.shape[1] == 1 and c.is_scalar
    c.tell(x, list(np.abs(x[:, 0])))
    assert np.asarray(c.mean()).shape == (1,)


def test_c1_plumbing_shape_on_cpu():
    """BASELINE.json configs[0]: invert_biggan_adam.py, num_samples=1, 256x256 — the reference's own CPU-runnable
    case (shortened to 2 gradient steps): the product's GradientOptimizer driving the full-size oracle generator and
    oracle ProjectionLoss on CPU reproduces the oracle closure's trajectory (oracle/closure.py, itself pinned to the
    real reference's closure.step by tests/test_golden_cpu.py)."""
    import torch.optim as optim
    from oracle import biggan as obg, closure as oc, lpips as olp
    from pix2latent_b200 import VariableManager
    from pix2latent_b200.optimizer import GradientOptimizer
    import pix2latent_b200.distribution as dist
    import pix2latent_b200.utils.function_hooks as hook
    model = obg.make_biggan(obg.BigGANConfig.deep256(), seed=0, calibrate=False)
    for p in model.parameters():
        p.requires_grad_(False)
    loss_fn = olp.ProjectionLoss(lpips_module=olp.make_lpips("alex", seed=0))
    g = torch.Generator().manual_seed(3)
    target = torch.tanh(0.5 * torch.randn(3, 256, 256, generator=g))
    weight = torch.full((3, 256, 256), 0.3)
    weight[:, 64:192, 64:192] = 1.0
    c0 = model.get_class_embedding(153)[0]

    def vm():
        m = VariableManager(device="cpu")
        m.register("z", (128,), "input", distribution=dist.TruncatedNormalModulo(sigma=1.0, trunc=2.0), learning_rate=0.05,
                   hook_fn=hook.Clamp(2.0))
        m.register("c", (128,), "input", default=c0, learning_rate=0.01)
        m.register("target", (3, 256, 256), "output", requires_grad=False, default=target)
        m.register("weight", (3, 256, 256), "output", requires_grad=False, default=weight)
        return m

    torch.manual_seed(7)
    opt = GradientOptimizer(model, vm(), loss_fn, max_batch_size=9)
    variables, outs, loss = opt.optimize(num_samples=1, grad_steps=2)
    assert outs[0].shape == (3, 256, 256) and loss[0][0] == 2 and len(loss[0][1]["loss"]) == 1
    # the same two steps through the oracle's restatement of closure.step
    torch.manual_seed(7)
    spec = {k: dict(v) for k, v in vm().variable_info.items()}
    ref_vars = oc.initialize(spec, 1)
    for _ in range(2):
        _, l_ref, _ = oc.step(model, ref_vars, loss_fn, optimize=True, max_batch_size=9)
    np.testing.assert_allclose(np.array(loss[0][1]["loss"]), np.array(l_ref), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(torch.stack(variables.input.z.data).detach().numpy(),
                               torch.stack(ref_vars.input.z.data).detach().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("script", ["invert_biggan_basincma.py", "invert_biggan_with_transform.py",
                                    "invert_stylegan2_cars_basincma.py"])
def test_reference_examples_import_through_the_alias(script):
    """The reference's own example scripts resolve every `pix2latent.*` import against this package (argument
    parsing happens after the imports, so `--help` exercises exactly the import block). Needs the reference checkout."""
    import os
    import subprocess
    import sys
    path = os.path.join("/root/reference/examples", script)
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root)
    r = subprocess.run([sys.executable, path, "--help"], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")
    assert r.returncode == 0, r.stderr[-2000:]
    assert "usage:" in r.stdout
