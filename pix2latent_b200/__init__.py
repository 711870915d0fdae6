"""pix2latent_b200 — the latent-inversion inner loop of pix2latent, built for B200 (sm_100a).

Public surface = the reference package's (pix2latent/__init__.py): ``VariableManager``,
``save_variables``, ``distribution``; sub-packages ``optimizer``, ``model``, ``loss_functions``,
``utils``. The hot path (generator forward, projection loss, backward to the latent) runs in the
C-ABI library ``libp2l.so`` (include/p2l.h); there is no other backend.
"""
from . import distribution
from .variable_manager import VariableManager, save_variables

__version__ = "0.1.0"
__all__ = ["optimizer", "utils", "model", "loss_functions", "distribution", "VariableManager",
           "save_variables"]
