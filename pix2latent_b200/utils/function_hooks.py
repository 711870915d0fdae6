"""In-place pre-forward hooks on the per-sample latent tensors (reference:
pix2latent/utils/function_hooks.py). They run inside the step before EVERY forward, including
eval-only steps (closure.py:42-44), and they draw from torch's global RNG one tensor at a time —
the RNG order is part of the behaviour."""
import math

import torch


class Clamp():
    """clamp_ every tensor to [-trunc, trunc]."""

    def __init__(self, trunc):
        self.trunc = trunc

    def __call__(self, vars):
        for v in vars:
            v.data.clamp_(-self.trunc, self.trunc)


class Normalize():
    """Shift/scale every tensor to zero mean, unit (unbiased) std — StyleGAN2 latent normalisation.
    ``mu``/``std`` are accepted and unused, as in the reference (function_hooks.py:40-50)."""

    def __init__(self, mu=0., std=1.):
        self.mu = mu
        self.std = std

    def __call__(self, vars):
        for v in vars:
            m, s = v.mean(), v.std()
            v.data.add_(-m).div_(s)


class NormalPerturb():
    """v += sigma * N(0, I), one randn_like per tensor."""

    def __init__(self, sigma=0.1):
        self.sigma = sigma

    def __call__(self, vars):
        for v in vars:
            v.data.add_(self.sigma * torch.randn_like(v))


class ScheduledNormalPerturb():
    """Perturbation decaying from sigma to 0 over max_step calls:
    strength = (sigma * max(0, 1 - t/(max_step-1))) ** 2 (the reference hard-codes pow=2 and
    forgets to import math, function_hooks.py:91,98 — behaviour kept, import fixed)."""

    def __init__(self, sigma=0.1, max_step=500, pow=2):
        self.sigma = sigma
        self.max_step = max_step
        self.t = 0
        self.pow = 2

    def __call__(self, vars):
        for v in vars:
            p = self.t / (float(self.max_step) - 1)
            strength = math.pow(self.sigma * max(0, 1 - p), self.pow)
            v.data.add_(strength * torch.randn_like(v))
        self.t += 1


class Compose():
    """Apply hooks in order."""

    def __init__(self, *hook_fns):
        self.hook_fns = hook_fns

    def __call__(self, vars):
        for fn in self.hook_fns:
            fn(vars)
