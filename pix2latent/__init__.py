"""`pix2latent` import name for the reference's example scripts: every `pix2latent.*` module
resolves to the corresponding `pix2latent_b200.*` module (same API, native sm_100a hot path), so
`examples/invert_*.py` of the reference run unmodified against this package."""
import importlib
import sys

import pix2latent_b200 as _impl
from pix2latent_b200 import VariableManager, distribution, save_variables  # noqa: F401

__version__ = _impl.__version__
_ALIASES = ["distribution", "variable_manager", "loss_functions", "optimizer", "optimizer.closure",
            "optimizer.base_optimizer", "optimizer.gradient_optimizer", "optimizer.base_cma_optimizer",
            "optimizer.cma_optimizer", "optimizer.basincma_optimizer", "model", "model.biggan", "model.stylegan2", "utils",
            "utils.function_hooks", "utils.misc", "utils.image"]
for _name in _ALIASES:
    sys.modules["pix2latent." + _name] = importlib.import_module("pix2latent_b200." + _name)


def __getattr__(name):
    mod = importlib.import_module("pix2latent_b200." + name)
    sys.modules["pix2latent." + name] = mod
    return mod
