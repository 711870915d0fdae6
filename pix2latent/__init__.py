"""`pix2latent` import name for the reference's example scripts: every `pix2latent.*` module
resolves to the corresponding `pix2latent_b200.*` module (same API, native sm_100a hot path), so
`examples/invert_*.py` of the reference run unmodified against this package."""
import importlib
import importlib.abc
import importlib.util
import sys

import pix2latent_b200 as _impl
from pix2latent_b200 import VariableManager, distribution, save_variables  # noqa: F401

__version__ = _impl.__version__


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        mod = importlib.import_module(self.target)  # the very same module object
        self._spec = mod.__spec__
        return mod

    def exec_module(self, module):
        module.__spec__ = self._spec  # the import machinery stamped the alias spec on the shared module


class _AliasFinder(importlib.abc.MetaPathFinder):
    """`import pix2latent.a.b` -> the module object of `pix2latent_b200.a.b`."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("pix2latent."):
            return None
        real = "pix2latent_b200." + fullname[len("pix2latent."):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real))


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())


def __getattr__(name):
    try:
        return importlib.import_module("pix2latent." + name)
    except ImportError as e:
        raise AttributeError(name) from e
