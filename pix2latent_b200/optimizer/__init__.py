from .gradient_optimizer import GradientOptimizer
from .basincma_optimizer import BasinCMAOptimizer
from .cma_optimizer import CMAOptimizer


def __getattr__(name):
    # Nevergrad-based optimizers import `nevergrad` lazily so the package imports without it
    if name == "NevergradOptimizer":
        from .ng_optimizer import NevergradOptimizer
        return NevergradOptimizer
    if name == "HybridNevergradOptimizer":
        from .hybrid_ng_optimizer import HybridNevergradOptimizer
        return HybridNevergradOptimizer
    raise AttributeError(name)
