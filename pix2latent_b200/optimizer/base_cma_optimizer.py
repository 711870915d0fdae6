"""CMA-ES plumbing shared by CMAOptimizer / BasinCMAOptimizer (reference:
pix2latent/optimizer/base_cma_optimizer.py:9-215). Host side; the population it asks for is
evaluated by ``closure.step``.

Kept behaviours: one CMA instance per ``grad_free`` variable and exactly one such variable
(assert, :64-66); population size dictated by CMA (:60); ``cma_init`` re-initialises every
variable (fresh Adam state) and overwrites the asked variable's ``.data`` with the float64
samples cast to the tensor's type (:79-87); ``cma_update`` re-evaluates the CURRENT (possibly
gradient-refined) variables with an eval-only step but tells CMA the ORIGINAL samples (:115-140).
"""
import numpy as np
import torch

from .. import parallel
from ..utils.image import binarize
from ..utils.misc import HiddenPrints, cprint


def _cma_module():
    try:
        import cma
        return cma
    except ImportError:
        from . import _minicma
        return _minicma


class CMA():
    """Wrapper over ``cma.CMAEvolutionStrategy`` (1-d problems are padded to 2-d with the
    covariance adaptation off, as the reference does, base_cma_optimizer.py:170-173)."""

    def __init__(self, mu=128 * [0], sigma=1.0, seed=None, popsize=None):
        options = {}
        if seed is not None:
            options["seed"] = seed
        if popsize is not None:
            options["popsize"] = int(popsize)
        self.is_scalar = False
        if len(mu) == 1:
            mu = list(mu) * 2
            options["CMA_on"] = 0
            self.is_scalar = True
        with HiddenPrints():
            self.cma = _cma_module().CMAEvolutionStrategy(mu, sigma, options)

    def batch_size(self):
        return self.cma.sp.popsize

    def ask(self, batch_size=None):
        x = np.array(self.cma.ask(batch_size))
        if self.is_scalar:
            self._x = x
            self._x_proxy = x[:, :1]
            return self._x_proxy
        return x

    def tell(self, x, y):
        if self.is_scalar:
            assert x is self._x_proxy
            return self.cma.tell(self._x, y)
        return self.cma.tell(x, y)

    def mean(self):
        x = self.cma.mean
        return x[:1] if self.is_scalar else x


class _BaseCMAOptimizer():
    """Mixin used together with _BaseOptimizer."""

    def __init__(self):
        self.num_samples = -1
        self.cma_optimizers = {}
        self._sampled = {}
        self.cma_seed = None  # not in the reference (its CMA is never seeded); opt-in
        self.cma_popsize = None  # opt-in: population size other than PyCMA's default (BASELINE configs[3]: 144 over 8 GPUs)

    @torch.no_grad()
    def setup_cma(self, var_manager):
        for name, spec in var_manager.variable_info.items():
            gf = spec["grad_free"]
            if gf is False:
                continue
            mu, sigma = (gf if type(gf) == tuple else (None, None))
            mu = np.zeros(spec["shape"]) if mu is None else mu
            sigma = 1.0 if sigma is None else sigma
            opt = CMA(mu, sigma=sigma, seed=self.cma_seed, popsize=self.cma_popsize)
            self.cma_optimizers[(spec["var_type"], name)] = opt
            self.num_samples = max(self.num_samples, opt.batch_size())
        cprint("(cma-es) number of samples: {}".format(self.num_samples), "y")
        assert len(self.cma_optimizers.keys()) == 1, \
            "currently only a single input variable can be optimized via CMA " + \
            "but got: {}".format(self.cma_optimizers.keys())

    @torch.no_grad()
    def cma_init(self, var_manager):
        variables = var_manager.initialize(num_samples=self.num_samples)
        rank, size = parallel.world()
        for (var_type, name), opt in self.cma_optimizers.items():
            asked = opt.ask() if rank == 0 else None
            if size > 1:
                asked = parallel.broadcast_array(asked)  # rank 0 owns the search state
            slots = variables[var_type][name].data
            for i, d in enumerate(asked):
                slots[i].data = torch.Tensor(d).data.type_as(slots[i].data)
            self._sampled[(var_type, name)] = asked
        return variables

    @torch.no_grad()
    def cma_update(self, variables, loss=None, inverted_loss=False):
        for key, opt in self.cma_optimizers.items():
            asked = self._sampled[key]
            if loss is None:
                out, loss, _ = self.step(variables, optimize=False)
                loss = self.gathered_loss()  # the one data-path collective: N scalar losses
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
                opt.tell(asked, loss)
        return loss
