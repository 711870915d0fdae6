"""Plain gradient descent with the optimizer declared in the VariableManager (reference:
pix2latent/optimizer/gradient_optimizer.py:11-56)."""
from .base_optimizer import _BaseOptimizer


class GradientOptimizer(_BaseOptimizer):

    def __init__(self, *args, **kwargs):
        _BaseOptimizer.__init__(self, *args, **kwargs)

    def optimize(self, num_samples, grad_steps, pbar=None):
        """
        Args
            num_samples (int): samples optimised in parallel
            grad_steps (int): gradient updates
            pbar: optional progress object with ``.progress(fraction)``
        """
        self._start_run()
        variables = self._variables = self.var_manager.initialize(num_samples=num_samples)
        # the reference reports progress i/grad_steps before incrementing
        self.grad_steps(variables, grad_steps,
                        lambda i: self._after_step(i, grad_steps, log_at=i + 1, log_last=grad_steps, pbar=pbar))
        return self._finish(variables, grad_steps)
