"""Nevergrad plumbing shared by NevergradOptimizer / HybridNevergradOptimizer (reference:
pix2latent/optimizer/base_ng_optimizer.py:10-171). Host side.

The reference file cannot run as shipped (it uses ``cprint`` and ``CMA`` without importing them
and calls ``initialize(num_seeds=...)``); the intended behaviour is implemented:
one Nevergrad optimizer for the single ``grad_free`` variable, ``num_samples`` asks per
meta-iteration, one ``tell`` per candidate with its (refined) loss."""
import numpy as np
import torch

from .. import parallel
from ..utils.image import binarize
from ..utils.misc import cprint


def _ng():
    try:
        import nevergrad as ng
        return ng
    except ImportError:
        # offline image: the minimal ask / tell stand-in (CMA, RandomSearch); production uses the real package
        from . import _mining
        return _mining


class _BaseNevergradOptimizer():

    def __init__(self, method):
        ng = _ng()
        self.method = method
        self.valid_methods = [x[0] for x in ng.optimizers.registry.items()]
        self.sequential_methods = ["SQPCMA", "chainCMAPowell", "Powell"]  # not exhaustive
        self.is_sequential = self.method in self.sequential_methods
        if self.is_sequential:
            cprint("{} is a sequential method. batch size is set to 1".format(self.method), "y")
        assert self.method in self.valid_methods, "unknown nevergrad method: {}".format(self.method)
        self.ng_optimizers = {}
        self._sampled = {}

    @torch.no_grad()
    def setup_ng(self, var_manager, budget):
        ng = _ng()
        for name, spec in var_manager.variable_info.items():
            gf = spec["grad_free"]
            if gf is False:
                continue
            mu = gf[0] if (type(gf) == tuple and gf[0] is not None) else np.zeros(spec["shape"])
            param = ng.p.Array(init=np.asarray(mu, dtype=np.float64))
            self.ng_optimizers[(spec["var_type"], name)] = \
                ng.optimizers.registry[self.method](parametrization=param, budget=budget)
        assert len(self.ng_optimizers.keys()) == 1, \
            "currently only a single input variable can be optimized via " + \
            "Nevergrad but got: {}".format(self.ng_optimizers.keys())

    @torch.no_grad()
    def ng_init(self, var_manager, num_samples):
        if self.is_sequential:
            num_samples = 1
        variables = var_manager.initialize(num_samples=num_samples)
        rank, size = parallel.world()
        for (var_type, name), opt in self.ng_optimizers.items():
            asked = [opt.ask() for _ in range(num_samples)] if rank == 0 else None
            values = np.concatenate([np.asarray(x.args[0])[None] for x in asked]) if rank == 0 else None
            if size > 1:
                values = parallel.broadcast_array(values)
            slots = variables[var_type][name].data
            for i, d in enumerate(values):
                slots[i].data = torch.Tensor(d).data.type_as(slots[i].data)
            self._sampled[(var_type, name)] = asked
        return variables

    @torch.no_grad()
    def ng_update(self, variables, loss=None, inverted_loss=False):
        for key, opt in self.ng_optimizers.items():
            asked = self._sampled[key]
            if loss is None:
                out, loss, _ = self.step(variables, optimize=False)
                loss = self.gathered_loss()
            if inverted_loss and hasattr(variables, "transform"):
                info = self.var_manager.variable_info
                target = info["target"]["default"].unsqueeze(0).type_as(out)
                weight = info["weight"]["default"].unsqueeze(0).type_as(out)
                t_fn = self.transform_fns["target"]["fn"]
                # (candidate sharding: `out` holds this rank's candidates only)
                rank, size = parallel.world()
                n_all = variables.num_samples
                lo, hi = parallel.shard_bounds(n_all, rank, size) if size > 1 else (0, n_all)
                out = t_fn(out, torch.stack(variables.transform.t.data[lo:hi]), invert=True)
                loss = self.loss_fn(out, target, binarize(weight)).cpu().detach().numpy()
                if size > 1:
                    loss = np.array(parallel.allgather_losses(loss, n_all))
            if parallel.world()[0] == 0:
                for cand, l in zip(asked, loss):
                    opt.tell(cand, float(l))
