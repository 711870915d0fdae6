"""Hybrid Nevergrad: outer gradient-free search, inner gradient descent (reference:
pix2latent/optimizer/hybrid_ng_optimizer.py:12-81). The population size is free here, which is
what BASELINE.json configs[4] (population 64) uses."""
from .base_ng_optimizer import _BaseNevergradOptimizer
from .base_optimizer import _BaseOptimizer


class HybridNevergradOptimizer(_BaseOptimizer, _BaseNevergradOptimizer):

    def __init__(self, method, *args, **kwargs):
        _BaseOptimizer.__init__(self, *args, **kwargs)
        _BaseNevergradOptimizer.__init__(self, method=method)

    def optimize(self, num_samples, meta_steps, grad_steps, last_grad_steps=300, pbar=None):
        """
        Args
            num_samples (int): candidates per Nevergrad update
            meta_steps (int): Nevergrad updates
            grad_steps (int): gradient updates per Nevergrad update
            last_grad_steps (int): gradient updates applied to the final draw
        """
        self._start_run()
        total_steps = meta_steps * grad_steps + last_grad_steps
        self.setup_ng(self.var_manager, budget=meta_steps * grad_steps)
        i = 0
        for meta_iter in range(meta_steps + 1):
            last = meta_iter == meta_steps
            variables = self._variables = self.ng_init(self.var_manager, num_samples)
            n_inner = last_grad_steps if last else grad_steps

            def on_step(j, i0=i):
                self._after_step(i0 + j + 1, total_steps, log_at=i0 + j + 2, log_last=grad_steps, pbar=pbar)

            self.grad_steps(variables, n_inner, on_step)
            i += n_inner
            if not last:
                self.ng_update(variables, inverted_loss=True)
        return self._finish(variables, total_steps)
