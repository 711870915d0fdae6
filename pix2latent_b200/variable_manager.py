"""Variable registry -> per-sample leaf tensors + one optimizer (host side of the boundary).

Mirrors the behaviour of /root/reference pix2latent/variable_manager.py (``VariableManager``
:69-240, ``split_vars`` :16-46, ``save_variables`` :49-65) including the quirks that change
numerics:
  * every sample of every variable is its own leaf tensor, one param-group per tensor with the
    variable's learning rate (:231-235);
  * the optimizer class is the one of the LAST registered spec (:238);
  * ``split_vars`` chunks share the single optimizer (:41);
  * ``register`` / ``edit_variable`` report problems by printing and returning False (:126-128,
    :182-190), shape mismatch of ``default`` is an assert (:130-133).
Differences on purpose: tensors go to ``device`` (default: CUDA when present) instead of a
hard-coded ``.cuda()`` (:217), and the attribute-dict is a local class (easydict is not a
dependency).
"""
import pprint

import numpy as np
import torch
import torch.optim as optim

from . import distribution as dist


class AttrDict(dict):
    """dict with attribute access, applied recursively to nested dicts (easydict semantics)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        src = dict(d or {})
        src.update(kw)
        for k, v in src.items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __delattr__(self, k):
        del self[k]


_META_KEYS = ("opt", "num_samples", "shard")  # "shard": (lo, hi, N) of a candidate-sharded view (parallel.shard_vars)


def split_vars(vars, size):
    """Chunks of at most ``size`` samples; every chunk refers to the same optimizer."""
    n = vars.num_samples
    chunks = []
    for lo in range(0, max(n, 1), size):
        part = {}
        count = 0
        for var_type, group in vars.items():
            if var_type in _META_KEYS:
                continue
            part[var_type] = {}
            for name, entry in group.items():
                piece = entry.data[lo:lo + size]
                count = len(piece)
                part[var_type][name] = {"data": piece, "hook_fn": entry.hook_fn}
        part["opt"] = vars.opt
        part["num_samples"] = count
        if "shard" in vars:
            part["shard"] = vars["shard"]
        chunks.append(AttrDict(part))
    return chunks


def save_variables(save_path, variables):
    """np.save of the variable dict with tensors moved to the CPU (vars.npy of the examples)."""
    out = {}
    for var_type, group in variables.items():
        if var_type == "opt":
            continue  # the reference intends to drop the optimizer (its `del` is a no-op bug)
        if not isinstance(group, dict):
            out[var_type] = group
            continue
        g = {}
        for name, entry in group.items():
            if isinstance(entry, dict) and "data" in entry:
                e = dict(entry)
                e["data"] = [t.detach().cpu() if torch.is_tensor(t) else t for t in entry["data"]]
                g[name] = e
            else:
                g[name] = entry
        out[var_type] = g
    np.save(save_path, out, allow_pickle=True)


def default_device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


class VariableManager():
    """Creates the variables an inversion optimises (same surface as the reference class)."""

    def __init__(self, device=None):
        self.variable_info = {}
        self.device = torch.device(device) if device is not None else None

    def __str__(self):
        return "<Variable Manager>\n{}".format(pprint.pformat(self.variable_info))

    def register(self, variable_name, shape, var_type, requires_grad=True, default=None,
                 distribution=dist.TruncatedNormalModulo(sigma=1.0, trunc=2.0), optimizer=optim.Adam,
                 learning_rate=0.05, hook_fn=None, grad_free=False):
        """Add a variable spec. ``var_type`` is 'input' (model kwargs), 'output' (loss kwargs)
        or 'transform'. See the reference docstring (variable_manager.py:95-124) for the fields."""
        if variable_name in self.variable_info:
            print("variable `{}`` already exists.".format(variable_name))
            return False
        if default is not None:
            assert tuple(default.size()) == tuple(shape), \
                "default and shape must match but got {} vs {}".format(list(default.size()), shape)
        self.variable_info[variable_name] = dict(
            shape=shape, var_type=var_type, requires_grad=requires_grad, default=default,
            distribution=distribution, optimizer=optimizer, learning_rate=learning_rate,
            hook_fn=hook_fn, grad_free=grad_free)
        return True

    def unregister(self, *variable_names):
        for name in variable_names:
            if name in self.variable_info:
                del self.variable_info[name]
            else:
                print("no variable named {}".format(name))

    def edit_variable(self, variable_name, replace_dict):
        if variable_name not in self.variable_info:
            print("variable `{}` does not exist".format(variable_name))
            return False
        spec = self.variable_info[variable_name]
        for k, v in replace_dict.items():
            if k not in spec:
                print("variable `{}` has no attribute {}".format(k, v))
                return False
            spec[k] = v
        return True

    @torch.no_grad()
    def initialize(self, num_samples):
        """Fresh per-sample leaves and a fresh optimizer (hence fresh Adam moments) — the
        reference calls this at every CMA / Nevergrad meta-iteration (base_cma_optimizer.py:79)."""
        device = self.device or default_device()
        groups, params = {}, []
        spec = None
        for name, spec in self.variable_info.items():
            if spec["default"] is not None:
                src = [spec["default"]] * num_samples
            else:
                src = list(spec["distribution"](num_samples, spec["shape"]))
            data = [t.detach().clone().to(device).requires_grad_(False) for t in src]
            groups.setdefault(spec["var_type"], {})[name] = dict(
                data=data, hook_fn=spec["hook_fn"], grad_free=spec["grad_free"],
                requires_grad=spec["requires_grad"])
            if spec["requires_grad"]:
                for t in data:
                    params.append({"params": t.requires_grad_(True), "lr": spec["learning_rate"]})
        assert spec is not None, "no variables registered"
        groups["opt"] = spec["optimizer"](params)
        groups["num_samples"] = num_samples
        return AttrDict(groups)
