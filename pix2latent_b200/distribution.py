"""Latent initialisation distributions (host side; same call signatures as the reference's
pix2latent/distribution.py:5-78: ``dist(num_samples, shape) -> Tensor[num_samples, *shape]``)."""
import torch


class TruncatedNormalModulo():
    """N(mu, I) wrapped into (-trunc, trunc) by float-modulo.

    Reference quirk kept on purpose (distribution.py:27-28): the constructor IGNORES its
    ``sigma`` and ``trunc`` arguments and always uses sigma=1.0, trunc=2.0, so
    ``TruncatedNormalModulo(sigma=1.0, trunc=args.truncate)`` in the examples samples from the
    same distribution whatever ``args.truncate`` is."""

    def __init__(self, mu=0., sigma=1., trunc=2.):
        self.mu = mu if type(mu) in [int, float] else mu.detach().cpu()
        self.sigma = 1.0
        self.trunc = 2.0

    def __call__(self, num_samples, shape):
        with torch.no_grad():
            x = self.sigma * torch.randn((num_samples, *shape))
            return torch.fmod(x + self.mu, self.trunc)


def truncated_clamp_normal(sigma=1.0, trunc=2.0):
    """sigma * N(0, I) clamped to [-trunc, trunc]. (The reference version, distribution.py:41-58,
    raises NameError when called — `samples`/`_clamp` are undefined; this one does what its
    docstring says.)"""
    def _dist_fn(num_samples, shape):
        with torch.no_grad():
            return (sigma * torch.randn((num_samples, *shape))).clamp_(-trunc, trunc)
    return _dist_fn


def normal(sigma=1.0):
    """sigma * N(0, I)."""
    def _dist_fn(num_samples, shape):
        with torch.no_grad():
            return sigma * torch.randn((num_samples, *shape))
    return _dist_fn
