"""BasinCMA: outer CMA-ES over z, inner gradient descent from every asked sample, CMA told the
ORIGINAL samples with the REFINED losses (reference: pix2latent/optimizer/basincma_optimizer.py
:12-83). BASELINE.json configs[1]/[3] are this loop."""
from .base_cma_optimizer import _BaseCMAOptimizer
from .base_optimizer import _BaseOptimizer


class BasinCMAOptimizer(_BaseOptimizer, _BaseCMAOptimizer):

    def __init__(self, *args, **kwargs):
        _BaseOptimizer.__init__(self, *args, **kwargs)
        _BaseCMAOptimizer.__init__(self)

    def optimize(self, meta_steps, grad_steps, last_grad_steps=300, pbar=None, num_samples=None):
        """
        Args
            meta_steps (int): CMA updates
            grad_steps (int): gradient updates per CMA update
            last_grad_steps (int): gradient updates applied to the final CMA draw
            num_samples: must be None (PyCMA fixes the population size)
        """
        assert num_samples == None, "PyCMA optimizer has fixed sample size"
        self.setup_cma(self.var_manager)
        self._start_run()
        total_steps = meta_steps * grad_steps + last_grad_steps
        i = 0
        for meta_iter in range(meta_steps + 1):
            last = meta_iter == meta_steps
            variables = self._variables = self.cma_init(self.var_manager)
            n_inner = last_grad_steps if last else grad_steps

            def on_step(j, i0=i):
                self._after_step(i0 + j + 1, total_steps, log_at=i0 + j + 2, log_last=grad_steps, pbar=pbar)

            self.grad_steps(variables, n_inner, on_step)
            i += n_inner
            if not last:
                self.cma_update(variables, inverted_loss=True)
        return self._finish(variables, total_steps)
