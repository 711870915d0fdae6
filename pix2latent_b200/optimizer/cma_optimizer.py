"""Pure CMA-ES search followed by optional gradient fine-tuning (reference:
pix2latent/optimizer/cma_optimizer.py:12-93)."""
from .base_cma_optimizer import _BaseCMAOptimizer
from .base_optimizer import _BaseOptimizer


class CMAOptimizer(_BaseOptimizer, _BaseCMAOptimizer):

    def __init__(self, *args, **kwargs):
        _BaseOptimizer.__init__(self, *args, **kwargs)
        _BaseCMAOptimizer.__init__(self)

    def optimize(self, meta_steps, grad_steps=0, pbar=None, num_samples=None):
        """
        Args
            meta_steps (int): CMA updates (each = one eval-only pass + the re-evaluation inside
                cma_update, i.e. two forward passes per meta step, as in the reference)
            grad_steps (int): gradient updates applied to a final CMA draw [Default: 0]
            num_samples: must be None (PyCMA fixes the population size)
        """
        assert num_samples == None, "PyCMA optimizer has fixed sample size"
        self.setup_cma(self.var_manager)
        self._start_run()
        total_steps = meta_steps + grad_steps
        i = 0
        for _ in range(meta_steps):
            variables = self._variables = self.cma_init(self.var_manager)
            self.step(variables, optimize=False, transform=False)
            i += 1
            self._maybe_log(variables, i, grad_steps)
            self.cma_update(variables, inverted_loss=True)
            self._progress(i, total_steps, i, pbar)
        variables = self._variables = self.cma_init(self.var_manager)
        for j in range(grad_steps):
            self.step(variables, optimize=True, transform=(j == 0))
            i += 1
            self._after_step(i, total_steps, log_at=i + 1, log_last=grad_steps, pbar=pbar)
        return self._finish(variables, total_steps)
