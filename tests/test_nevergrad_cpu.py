"""Nevergrad search loops (reference: pix2latent/optimizer/{ng,hybrid_ng,base_ng}_optimizer.py) on CPU with the
offline ask / tell stand-in (pix2latent_b200/optimizer/_mining.py; the real package is not installed here):
one optimizer for the single grad_free variable, num_samples asks per meta-iteration, one tell per candidate with its
refined loss, the population size free (BASELINE.json configs[4] uses 64)."""
import numpy as np
import pytest
import torch


def _problem():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg
    from oracle import lpips as olp
    cfg, model, target, weight = mg.problem()
    loss_fn = olp.ProjectionLoss(lpips_module=olp.make_lpips("alex", seed=0))
    return mg, model, target, weight, loss_fn


def _vm(mg, model, target, weight):
    from pix2latent_b200 import VariableManager
    import pix2latent_b200.distribution as dist
    import pix2latent_b200.utils.function_hooks as hook
    vm = VariableManager(device="cpu")
    mg.register(vm, hook, dist, model, target, weight, True)
    return vm


def test_stand_in_ask_tell_buffering():
    from pix2latent_b200.optimizer import _mining as ng
    ng._CMA.seed = 3
    opt = ng.optimizers.registry["CMA"](parametrization=ng.p.Array(init=np.zeros(6)), budget=100)
    pop = opt.es.sp.popsize  # 4 + floor(3 ln 6) = 9
    mean0 = opt.es.mean.copy()
    cands = [opt.ask() for _ in range(pop + 2)]          # more asks than one population: a second draw of the same distribution
    assert cands[0].args[0].shape == (6,) and opt.num_ask == pop + 2
    for c in cands[:pop - 1]:
        opt.tell(c, float(np.sum(c.args[0] ** 2)))
    assert np.array_equal(opt.es.mean, mean0)            # distribution moves only after popsize tells
    opt.tell(cands[pop - 1], float(np.sum(cands[pop - 1].args[0] ** 2)))
    assert not np.array_equal(opt.es.mean, mean0) and opt.es.countiter == 1
    assert opt.provide_recommendation().args[0].shape == (6,)
    # minimises a quadratic
    for _ in range(60):
        cs = [opt.ask() for _ in range(pop)]
        for c in cs:
            opt.tell(c, float(np.sum((c.args[0] - 1.0) ** 2)))
    assert np.abs(opt.es.mean - 1.0).max() < 0.1


def test_hybrid_nevergrad_loop():
    from pix2latent_b200.optimizer import HybridNevergradOptimizer, _mining
    mg, model, target, weight, loss_fn = _problem()
    _mining._CMA.seed = 11
    torch.manual_seed(5)
    opt = HybridNevergradOptimizer("CMA", model, _vm(mg, model, target, weight), loss_fn, max_batch_size=3)
    variables, outs, loss = opt.optimize(num_samples=5, meta_steps=2, grad_steps=2, last_grad_steps=3)
    ngo = list(opt.ng_optimizers.values())[0]
    assert ngo.num_ask == 15 and ngo.num_tell == 10       # 3 draws of 5, tells after the first two
    assert loss[0][0] == 2 * 2 + 3 and len(loss[0][1]["loss"]) == 5 and np.isfinite(loss[0][1]["loss"]).all()
    assert len(opt.tracked["z"]) == 2 * (2 + 1) + 3       # grad steps + the eval-only step of every ng_update
    assert outs[0].shape[0] == 3
    with pytest.raises(AssertionError):
        HybridNevergradOptimizer("NoSuchMethod", model, _vm(mg, model, target, weight), loss_fn)


def test_nevergrad_only_loop_improves():
    from pix2latent_b200.optimizer import NevergradOptimizer, _mining
    mg, model, target, weight, loss_fn = _problem()
    _mining._CMA.seed = 12
    torch.manual_seed(6)
    opt = NevergradOptimizer("CMA", model, _vm(mg, model, target, weight), loss_fn, max_batch_size=9)
    variables, outs, loss = opt.optimize(num_samples=18, meta_steps=3, grad_steps=1)
    ngo = list(opt.ng_optimizers.values())[0]
    assert ngo.num_ask == 4 * 18 and ngo.num_tell == 3 * 18 and ngo.es.countiter == 3
    assert len(loss[0][1]["loss"]) == 18
