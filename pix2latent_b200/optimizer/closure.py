"""The inner evaluation step — the drop-in boundary.

Behavioural contract = /root/reference pix2latent/optimizer/closure.py:6-79:
  for each chunk of <= max_batch_size samples (chunks share ONE optimizer):
      zero_grad; run the input hooks in place (also on eval-only steps); stack the per-sample
      leaves; out = model(**inputs); loss = loss_fn(out, **targets).view(b,-1).mean(1);
      loss.mean().backward()  (=> every sample's gradient is scaled by 1/b_chunk);
      optimizer.step (only the leaves that received a gradient move); zero_grad
  return (stacked outputs [N,...], list of N per-sample losses, {})

Two executions of that contract:
  * ``_step_autograd`` — any nn.Module / loss, through torch autograd (the native BigGAN and the
    native losses take part as autograd.Functions);
  * ``_step_native`` — when model and loss are the library's own, the whole
    generator-forward -> loss -> backward-to-latent chain is ONE C-ABI call per step
    (p2l_biggan_step), with the 1/b_chunk scales passed per sample, so the population is not
    split into mini-batches on a 180 GB device (the reference's chunking is a memory workaround;
    results per candidate do not depend on it).
"""
import torch

from ..variable_manager import split_vars


def _unwrap(model):
    return model.module if isinstance(model, torch.nn.DataParallel) else model


def _native_pair(model, vars, loss_fn):
    from ..loss_functions import _NativeLoss
    from ..model.biggan import BigGAN
    m = _unwrap(model)
    if not (isinstance(m, BigGAN) and isinstance(loss_fn, _NativeLoss) and m.native is not None):
        return False
    if any(k not in ("input", "output", "opt", "num_samples") for k in vars.keys()):
        return False  # e.g. 'transform' variables: per-sample targets -> autograd path
    if set(vars.input.keys()) != {"z", "c"}:
        return False
    outs = set(vars.output.keys()) if "output" in vars else set()
    return "target" in outs and outs <= {"target", "weight", "loss_mask"}


def _native_targets_pair(model, vars, loss_fn):
    """BigGAN + native loss with PER-CANDIDATE targets: a 'transform' variable group is present, i.e.
    base_optimizer.apply_transform replaced every sample's target / weight (transform search)."""
    from ..loss_functions import _NativeLoss
    from ..model.biggan import BigGAN
    m = _unwrap(model)
    if not (isinstance(m, BigGAN) and isinstance(loss_fn, _NativeLoss) and m.native is not None):
        return False
    if "transform" not in vars.keys():
        return False
    if any(k not in ("input", "output", "transform", "opt", "num_samples") for k in vars.keys()):
        return False
    if set(vars.input.keys()) != {"z", "c"}:
        return False
    outs = set(vars.output.keys()) if "output" in vars else set()
    return "target" in outs and outs <= {"target", "weight", "loss_mask"}


def _run_hooks(group):
    for _, var in group.items():
        if var.hook_fn is not None:
            var.hook_fn(var.data)


def _step_autograd(model, vars, loss_fn, optimize, max_batch_size):
    outs, losses = [], []
    for chunk in split_vars(vars, size=max_batch_size):
        state = {}

        def closure():
            b = chunk.num_samples
            targets = {k: torch.stack(v.data) for k, v in chunk.output.items()} if "output" in chunk else {}
            if optimize:
                chunk.opt.zero_grad()
            _run_hooks(chunk.input)
            inputs = {k: torch.stack(v.data) for k, v in chunk.input.items()}
            out = model(**inputs)
            loss = loss_fn(out, **targets).view(b, -1).mean(1)
            if optimize:
                loss.mean().backward()
            state["out"] = out
            state["loss"] = loss.detach().cpu().numpy()

        if optimize:
            chunk.opt.step(closure)
            chunk.opt.zero_grad()
        else:
            with torch.no_grad():
                chunk.opt.step(closure)
        outs.extend(state["out"].detach())
        losses.extend(state["loss"])
    return torch.stack(outs), losses, {}


def _step_native(model, vars, loss_fn, optimize, max_batch_size):
    from .. import native
    m = _unwrap(model)
    chunks = split_vars(vars, size=max_batch_size)
    # hooks chunk by chunk in the reference's order (they consume torch's global RNG)
    for chunk in chunks:
        _run_hooks(chunk.input)
    z_list, c_list = vars.input.z.data, vars.input.c.data
    n = len(z_list)
    with torch.no_grad():
        z = torch.stack(z_list)
        c = torch.stack(c_list)
    out_vars = vars.output
    first = {k: v.data[0] for k, v in out_vars.items()}
    tgt = loss_fn.prepared_target(first["target"], first.get("weight"), first.get("loss_mask"))
    # d(mean over the chunk)/d loss_i = 1 / chunk size, per sample
    dloss = torch.cat([torch.full((ch.num_samples,), 1.0 / ch.num_samples) for ch in chunks]).to(z.device)
    loss, dz, dc, img = native.biggan_step(m.native, loss_fn.native_lpips(), tgt, z, c, want_grad=optimize,
                                           grad_scale=1.0, dloss=dloss)
    if optimize:
        opt = vars.opt
        opt.zero_grad()
        for i in range(n):
            if z_list[i].requires_grad:
                z_list[i].grad = dz[i]
            if c_list[i].requires_grad:
                c_list[i].grad = dc[i]
        opt.step()
        opt.zero_grad()
    return img, list(loss.cpu().numpy()), {}


def _step_native_targets(model, vars, loss_fn, optimize, max_batch_size):
    """As ``_step_native`` with one cached NativeTarget per candidate (p2l_biggan_step_targets)."""
    from .. import native
    m = _unwrap(model)
    chunks = split_vars(vars, size=max_batch_size)
    for chunk in chunks:
        _run_hooks(chunk.input)
    z_list, c_list = vars.input.z.data, vars.input.c.data
    n = len(z_list)
    with torch.no_grad():
        z = torch.stack(z_list)
        c = torch.stack(c_list)
    o = vars.output
    tgts = loss_fn.prepared_targets(o.target.data, o.weight.data if "weight" in o else None,
                                    o.loss_mask.data if "loss_mask" in o else None)
    dloss = torch.cat([torch.full((ch.num_samples,), 1.0 / ch.num_samples) for ch in chunks]).to(z.device)
    loss, dz, dc, img = native.biggan_step_targets(m.native, loss_fn.native_lpips(), tgts, z, c, want_grad=optimize,
                                                   grad_scale=1.0, dloss=dloss)
    if optimize:
        opt = vars.opt
        opt.zero_grad()
        for i in range(n):
            if z_list[i].requires_grad:
                z_list[i].grad = dz[i]
            if c_list[i].requires_grad:
                c_list[i].grad = dc[i]
        opt.step()
        opt.zero_grad()
    return img, list(loss.cpu().numpy()), {}


def _native_sg2_pair(model, vars, loss_fn):
    from ..loss_functions import _NativeLoss
    from ..model.stylegan2 import StyleGAN2
    m = _unwrap(model)
    if not (isinstance(m, StyleGAN2) and isinstance(loss_fn, _NativeLoss) and m.native is not None and m.search == "z"):
        return False  # w / w+ search: the differentiable model / loss calls of the autograd path
    if any(k not in ("input", "output", "opt", "num_samples") for k in vars.keys()):
        return False
    if set(vars.input.keys()) != {"z"}:
        return False
    outs = set(vars.output.keys()) if "output" in vars else set()
    return "target" in outs and outs <= {"target", "weight", "loss_mask"}


def _step_native_sg2(model, vars, loss_fn, optimize, max_batch_size):
    """StyleGAN2 (z search): chunk by chunk as the reference does, because hooks (NormalPerturb) and the
    per-layer noise both draw from torch's RNG between chunks — the draw ORDER is part of the behaviour."""
    from .. import native
    m = _unwrap(model)
    first = {k: v.data[0] for k, v in vars.output.items()}
    tgt = loss_fn.prepared_target(first["target"], first.get("weight"), first.get("loss_mask"))
    outs, losses = [], []
    for chunk in split_vars(vars, size=max_batch_size):
        if optimize:
            chunk.opt.zero_grad()
        _run_hooks(chunk.input)
        z_list = chunk.input.z.data
        with torch.no_grad():
            z = torch.stack(z_list)
        noises = m.draw_noise(z.shape[0], z.device)
        loss, dz, img = native.sg2_step(m.native, loss_fn.native_lpips(), tgt, z, noises, want_grad=optimize,
                                        grad_scale=1.0 / chunk.num_samples)
        if optimize:
            for i, t in enumerate(z_list):
                if t.requires_grad:
                    t.grad = dz[i]
            chunk.opt.step()
            chunk.opt.zero_grad()
        outs.extend(img)
        losses.extend(loss.cpu().numpy())
    return torch.stack(outs), losses, {}


def step(model, vars, loss_fn, optimize=True, max_batch_size=9):
    """One evaluation (and, with ``optimize``, one gradient update) of every sample in ``vars``.

    Returns ``(outs [N,3,H,W], indiv_losses list[N], {})`` as the reference does."""
    if _native_pair(model, vars, loss_fn):
        return _step_native(model, vars, loss_fn, optimize, max_batch_size)
    if _native_targets_pair(model, vars, loss_fn):
        return _step_native_targets(model, vars, loss_fn, optimize, max_batch_size)
    if _native_sg2_pair(model, vars, loss_fn):
        return _step_native_sg2(model, vars, loss_fn, optimize, max_batch_size)
    return _step_autograd(model, vars, loss_fn, optimize, max_batch_size)
