"""Nevergrad search followed by optional gradient fine-tuning (reference:
pix2latent/optimizer/ng_optimizer.py:14-91)."""
from .base_ng_optimizer import _BaseNevergradOptimizer
from .base_optimizer import _BaseOptimizer


class NevergradOptimizer(_BaseOptimizer, _BaseNevergradOptimizer):

    def __init__(self, method, *args, **kwargs):
        _BaseOptimizer.__init__(self, *args, **kwargs)
        _BaseNevergradOptimizer.__init__(self, method=method)

    def optimize(self, num_samples, meta_steps, grad_steps=0, pbar=None):
        """
        Args
            num_samples (int): candidates per Nevergrad update
            meta_steps (int): Nevergrad updates
            grad_steps (int): gradient updates applied to a final draw [Default: 0]
        """
        self.setup_ng(self.var_manager, budget=meta_steps)
        self._start_run()
        total_steps = meta_steps + grad_steps
        i = 0
        for _ in range(meta_steps):
            variables = self._variables = self.ng_init(self.var_manager, num_samples)
            self.step(variables, optimize=False, transform=False)
            i += 1
            self._maybe_log(variables, i, grad_steps)
            self.ng_update(variables, inverted_loss=True)
            self._progress(i, total_steps, i, pbar)
        variables = self._variables = self.ng_init(self.var_manager, num_samples)
        for j in range(grad_steps):
            self.step(variables, optimize=True, transform=(j == 0))
            i += 1
            self._after_step(i, total_steps, log_at=i + 1, log_last=grad_steps, pbar=pbar)
        return self._finish(variables, total_steps)
