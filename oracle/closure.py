"""ORACLE (test infrastructure) — CPU restatement of the reference's evaluation step and variable
handling, device-parametrised (the reference hard-codes ``.cuda()``).

Follows /root/reference pix2latent/optimizer/closure.py:6-79 (``step``),
pix2latent/variable_manager.py:16-46 (``split_vars``) and :196-240 (``initialize``), under the
installed torch's semantics (SURVEY.md §8c "version quirk": ``zero_grad(set_to_none=True)``;
``Adam.step(closure)`` runs the closure under enable_grad and skips params whose grad is None).

PINNED: tests/golden/make_golden.py runs the REAL reference functions (imported from
/root/reference with the missing third-party modules stubbed) on the same inputs and
tests/test_golden_cpu.py checks this restatement — and the product's host code — against those
vectors.
"""
import numpy as np
import torch


class Bag(dict):
    """attribute-access dict (stand-in for easydict, which the reference uses)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def initialize(variable_info, num_samples, device="cpu"):
    """variable_manager.py:196-240 on an explicit device."""
    groups, params, spec = {}, [], None
    with torch.no_grad():
        for name, spec in variable_info.items():
            if spec["default"] is not None:
                data = num_samples * [spec["default"]]
            else:
                data = list(spec["distribution"](num_samples, spec["shape"]))
            data = [d.detach().clone().to(device).requires_grad_(False) for d in data]
            groups.setdefault(spec["var_type"], Bag())[name] = Bag(
                data=data, hook_fn=spec["hook_fn"], grad_free=spec["grad_free"], requires_grad=spec["requires_grad"])
            if not spec["requires_grad"]:
                continue
            for d in data:
                params.append({"params": d.requires_grad_(True), "lr": spec["learning_rate"]})
    out = Bag(groups)
    out["opt"] = spec["optimizer"](params)
    out["num_samples"] = num_samples
    return out


def split_vars(vars, size):
    """variable_manager.py:16-46."""
    n_splits = int(np.ceil(vars.num_samples / float(size)))
    chunks = []
    for i in range(n_splits):
        sub = Bag()
        n = 0
        for var_type, var_dict in vars.items():
            if var_type in ["opt", "num_samples"]:
                continue
            sub[var_type] = Bag()
            for var_name, var_data in var_dict.items():
                data = var_data.data[i * size:(i + 1) * size]
                n = len(data)
                sub[var_type][var_name] = Bag(data=data, hook_fn=var_data.hook_fn)
        sub["opt"] = vars.opt
        sub["num_samples"] = n
        chunks.append(sub)
    return chunks


def step(model, vars, loss_fn, optimize=True, max_batch_size=9):
    """closure.py:6-79. Returns (stacked outs, list of per-sample losses, {})."""
    outs, indiv_losses = [], []
    for _vars in split_vars(vars, size=max_batch_size):
        box = {}

        def closure():
            b_sz = _vars.num_samples
            target_args = {k: torch.stack(v.data) for k, v in _vars.output.items()}
            if optimize:
                _vars.opt.zero_grad()
            for _, var_dict in _vars.input.items():  # (1) hooks, closure.py:42-44
                if var_dict.hook_fn is not None:
                    var_dict.hook_fn(var_dict.data)
            input_args = {k: torch.stack(v.data) for k, v in _vars.input.items()}
            out = model(**input_args)  # (2) closure.py:51
            loss = loss_fn(out, **target_args).view(b_sz, -1).mean(1)  # (3) closure.py:55
            if optimize:
                loss.mean().backward()  # closure.py:58: every sample's grad carries 1/b_sz
            box["out"], box["loss"] = out, loss.detach().cpu().numpy()

        if optimize:  # (4) closure.py:64-71
            _vars.opt.step(closure)
        else:
            with torch.no_grad():
                _vars.opt.step(closure)
        if optimize:
            _vars.opt.zero_grad()
        outs.extend(box["out"].detach())
        indiv_losses.extend(box["loss"])
    return torch.stack(outs), indiv_losses, {}
