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
import os
from collections.abc import Sequence

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
    if any(k not in ("input", "output", "opt", "num_samples", "shard") for k in vars.keys()):
        return False  # e.g. 'transform' variables: per-sample targets
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
    if any(k not in ("input", "output", "transform", "opt", "num_samples", "shard") for k in vars.keys()):
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
    # nn.DataParallel around the library's own models is a pass-through (their native handle is bound to one device;
    # the reference wraps StyleGAN2 that way, examples/invert_stylegan2_cars_basincma.py:51)
    if isinstance(model, torch.nn.DataParallel) and getattr(model.module, "native", None) is not None:
        model = model.module
    flush_native_adam(vars.opt)  # a torch-side step follows: opt.state must be current
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


class LazyLosses(Sequence):
    """Per-sample losses of a step as the list of floats the reference returns (closure.py:79) — but the device ->
    host copy happens on FIRST USE, not inside the step: a step no longer ends with a stream synchronisation
    (closure.py:60), so the launches of the next step queue up behind the running one."""

    def __init__(self, dev_tensor):
        self._dev = dev_tensor
        self._host = None

    def device_tensor(self):
        return self._dev

    def _get(self):
        if self._host is None:
            self._host = list(self._dev.detach().cpu().numpy())
        return self._host

    def __len__(self):
        return int(self._dev.shape[0])

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())

    def __repr__(self):
        return repr(self._get())

    def __array__(self, dtype=None, copy=None):
        import numpy as np
        a = np.asarray(self._get())
        return a if dtype is None else a.astype(dtype)


def chunk_scales(vars, max_batch_size, device):
    """d(mean over the chunk)/d loss_i = 1 / chunk size (closure.py:58), per sample of ``vars``. Under candidate
    sharding (``vars.shard`` = (lo, hi, N), parallel.shard_vars) the chunks are those of the WHOLE population, so every
    candidate gets the scale it has in the single-process run."""
    n = vars.num_samples
    lo, hi, total = vars["shard"] if "shard" in vars else (0, n, n)
    key = (lo, hi, total, max_batch_size, str(device))
    hit = _scale_cache.get(key)
    if hit is None:
        sizes = [min(max_batch_size, total - (g // max_batch_size) * max_batch_size) for g in range(lo, hi)]
        hit = torch.tensor([1.0 / s for s in sizes], dtype=torch.float32).to(device)
        if len(_scale_cache) > 64:
            _scale_cache.clear()
        _scale_cache[key] = hit
    return hit


_scale_cache = {}


def adam_plan(opt, z_list, c_list):
    """Hyper-parameters of a plain ``torch.optim.Adam`` over exactly the per-sample z / c leaves (one lr per variable,
    shared betas / eps, no weight decay / amsgrad / ...), or None. ``step0`` / ``stateful``: the torch optimizer's own
    state (fresh everywhere, or stepped the same number of times everywhere)."""
    if type(opt) is not torch.optim.Adam:
        return None
    group_of = {}
    for g in opt.param_groups:
        if (g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False)
                or g.get("capturable", False) or g.get("differentiable", False) or g.get("fused", None)):
            return None
        if g.get("decoupled_weight_decay", False):
            return None
        for p in g["params"]:
            group_of[id(p)] = g
    hp = None
    lrs = []
    for lst in (z_list, c_list):
        lr = None
        for t in lst:
            g = group_of.get(id(t))
            if g is None:
                if t.requires_grad:
                    return None  # a trainable leaf the optimizer does not know
                this = 0.0
            else:
                if not t.requires_grad:
                    return None
                this = float(g["lr"])
                key = (tuple(float(x) for x in g["betas"]), float(g["eps"]))
                if hp is None:
                    hp = key
                elif hp != key:
                    return None
            if lr is None:
                lr = this
            elif lr != this:
                return None
        lrs.append(lr)
    if hp is None:
        return None
    steps = set()
    for lst in (z_list, c_list):
        for t in lst:
            if id(t) in group_of:
                st = opt.state.get(t, None)
                steps.add(int(st["step"]) if st else 0)
    if len(steps) > 1:
        return None
    return dict(lr_z=lrs[0], lr_c=lrs[1], betas=hp[0], eps=hp[1], step0=steps.pop() if steps else 0,
                stateful=set(group_of.keys()), n_owned=len(group_of))


def native_adam(opt, z_list, c_list):
    """Device-resident Adam state (native.AdamState: moments [n, dim] + step counter) shared by the per-step path
    (p2l_adam_update) and the fused loop (p2l_biggan_optimize) for the leaves ``z_list`` / ``c_list`` of ``opt``;
    None when ``opt`` is not a plain Adam. While it exists it is the truth; ``flush_native_adam`` writes it back into
    ``opt.state`` (torch's per-tensor entries) — done at the end of every optimize() and before any torch-side step."""
    from .. import native
    key = (tuple(id(t) for t in z_list), tuple(id(t) for t in c_list))
    cache = getattr(opt, "_p2l_adam", None)
    if cache is not None and cache["key"] != key:
        flush_native_adam(opt)
        cache = None
    plan = adam_plan(opt, z_list, c_list)
    if plan is None:
        return None
    if cache is None:
        n = len(z_list)
        zd, cd = z_list[0].numel(), c_list[0].numel()
        state = native.AdamState(n, zd, cd, z_list[0].device, step=plan["step0"])
        if plan["step0"] > 0:
            mz, vz, mc, vc = state.moments()
            with torch.no_grad():
                for i in range(n):
                    for t, mm, vv in ((z_list[i], mz, vz), (c_list[i], mc, vc)):
                        st = opt.state.get(t, None)
                        if st:
                            mm[i].copy_(st["exp_avg"].view(-1))
                            vv[i].copy_(st["exp_avg_sq"].view(-1))
        cache = dict(key=key, state=state, z_list=list(z_list), c_list=list(c_list), steps=plan["step0"],
                     stateful=plan["stateful"])
        opt._p2l_adam = cache
    cache["plan"] = plan
    return cache


@torch.no_grad()
def flush_native_adam(opt):
    """Write the native Adam state back into ``opt.state`` and drop it (see ``native_adam``)."""
    cache = getattr(opt, "_p2l_adam", None)
    if cache is None:
        return
    mz, vz, mc, vc = cache["state"].moments()
    t_now = float(cache["steps"])
    for i in range(len(cache["z_list"])):
        for t, mm, vv in ((cache["z_list"][i], mz, vz), (cache["c_list"][i], mc, vc)):
            if id(t) in cache["stateful"] and t_now > 0:
                opt.state[t] = {"step": torch.tensor(t_now), "exp_avg": mm[i].clone().view_as(t),
                                "exp_avg_sq": vv[i].clone().view_as(t)}
    del opt._p2l_adam


def _uniform_outputs(loss_fn, out_vars):
    """True when every sample's target / weight / loss_mask equals sample 0's — the condition under which ONE prepared
    target serves the whole population (what VariableManager.initialize produces from a registered default). The
    reference stacks every sample's own tensors (closure.py:33-34), so anything else goes through the per-candidate
    target path. Checked once per set of tensors (storage + version), not per step."""
    key = tuple((name, tuple((t.data_ptr(), t._version) for t in v.data)) for name, v in out_vars.items())
    cache = loss_fn.__dict__.setdefault("_uniform_cache", {})
    hit = cache.get(key)
    if hit is None:
        hit = True
        for _, v in out_vars.items():
            first = v.data[0]
            for t in v.data[1:]:
                if t is first or (t.data_ptr() == first.data_ptr() and t.shape == first.shape):
                    continue
                if t.shape != first.shape or not torch.equal(t, first):
                    hit = False
                    break
            if not hit:
                break
        if len(cache) > 16:
            cache.clear()
        cache[key] = hit
    return hit


def _apply_update(vars, z, c, dz, dc, z_list, c_list):
    """``opt.step()`` of closure.py:65 for the latent leaves: the device-resident Adam when the optimizer is a plain
    Adam (one launch for the whole population), torch's optimizer over the 2n per-sample groups otherwise."""
    from .. import native
    opt = vars.opt
    ad = native_adam(opt, z_list, c_list)
    if ad is not None:
        p = ad["plan"]
        cfg = native.adam_config(p["lr_z"], p["lr_c"], p["betas"], p["eps"])
        native.adam_update(z, c, dz, dc, cfg, ad["state"])
        ad["steps"] += 1
        with torch.no_grad():
            torch._foreach_copy_([t.data.view(-1) for t in z_list], list(z.unbind(0)))
            torch._foreach_copy_([t.data.view(-1) for t in c_list], list(c.unbind(0)))
        return
    opt.zero_grad()
    for i in range(len(z_list)):
        if z_list[i].requires_grad:
            z_list[i].grad = dz[i].view_as(z_list[i])
        if c_list[i].requires_grad:
            c_list[i].grad = dc[i].view_as(c_list[i])
    opt.step()
    opt.zero_grad()


def _step_native(model, vars, loss_fn, optimize, max_batch_size):
    from .. import native
    m = _unwrap(model)
    chunks = split_vars(vars, size=max_batch_size)
    # hooks chunk by chunk in the reference's order (they consume torch's global RNG)
    for chunk in chunks:
        _run_hooks(chunk.input)
    z_list, c_list = vars.input.z.data, vars.input.c.data
    with torch.no_grad():
        z = torch.stack(z_list).float().contiguous()
        c = torch.stack(c_list).float().contiguous()
    first = {k: v.data[0] for k, v in vars.output.items()}
    tgt = loss_fn.prepared_target(first["target"], first.get("weight"), first.get("loss_mask"))
    dloss = chunk_scales(vars, max_batch_size, z.device)
    loss, dz, dc, img = native.biggan_step(m.native, loss_fn.native_lpips(), tgt, z, c, want_grad=optimize,
                                           grad_scale=1.0, dloss=dloss)
    if optimize:
        _apply_update(vars, z, c, dz, dc, z_list, c_list)
    return img, LazyLosses(loss), {}


def _step_native_targets(model, vars, loss_fn, optimize, max_batch_size):
    """As ``_step_native`` with one cached NativeTarget per candidate (p2l_biggan_step_targets)."""
    from .. import native
    m = _unwrap(model)
    chunks = split_vars(vars, size=max_batch_size)
    for chunk in chunks:
        _run_hooks(chunk.input)
    z_list, c_list = vars.input.z.data, vars.input.c.data
    with torch.no_grad():
        z = torch.stack(z_list).float().contiguous()
        c = torch.stack(c_list).float().contiguous()
    o = vars.output
    tgts = loss_fn.prepared_targets(o.target.data, o.weight.data if "weight" in o else None,
                                    o.loss_mask.data if "loss_mask" in o else None)
    dloss = chunk_scales(vars, max_batch_size, z.device)
    loss, dz, dc, img = native.biggan_step_targets(m.native, loss_fn.native_lpips(), tgts, z, c, want_grad=optimize,
                                                   grad_scale=1.0, dloss=dloss)
    if optimize:
        _apply_update(vars, z, c, dz, dc, z_list, c_list)
    return img, LazyLosses(loss), {}


def _native_sg2_pair(model, vars, loss_fn):
    from ..loss_functions import _NativeLoss
    from ..model.stylegan2 import StyleGAN2
    m = _unwrap(model)
    if not (isinstance(m, StyleGAN2) and isinstance(loss_fn, _NativeLoss) and m.native is not None and m.search == "z"):
        return False  # w / w+ search: the differentiable model / loss calls of the autograd path
    if any(k not in ("input", "output", "opt", "num_samples", "shard") for k in vars.keys()):
        return False
    if set(vars.input.keys()) != {"z"}:
        return False
    outs = set(vars.output.keys()) if "output" in vars else set()
    return "target" in outs and outs <= {"target", "weight", "loss_mask"}


# candidates per physical StyleGAN2 launch sequence (the reference's max_batch_size only sets the 1/b_chunk gradient scales
# and the order of the RNG draws here; a candidate's result does not depend on the batch it is evaluated in)
SG2_PHYS_BATCH = int(os.environ.get("P2L_SG2_PHYS_BATCH", "24"))


def _step_native_sg2(model, vars, loss_fn, optimize, max_batch_size):
    """StyleGAN2 (z search). The reference works chunk by chunk (closure.py:32-66): hooks (NormalPerturb) and the per-layer
    noise both draw from torch's RNG between chunks, and ``loss.mean().backward()`` scales a chunk's gradients by 1/b_chunk.
    Here the hooks run and the noise is drawn chunk by chunk IN THAT ORDER, the per-candidate scales follow the reference's
    chunking (``chunk_scales``), and the whole population goes through the native step in physical batches of up to
    ``SG2_PHYS_BATCH`` candidates — bitwise the same per candidate (tests/test_determinism_gpu.py), fewer launches."""
    from .. import native
    m = _unwrap(model)
    first = {k: v.data[0] for k, v in vars.output.items()}
    tgt = loss_fn.prepared_target(first["target"], first.get("weight"), first.get("loss_mask"))
    if optimize:
        vars.opt.zero_grad()
    z_list = vars.input.z.data
    dev = z_list[0].device
    parts = []
    for chunk in split_vars(vars, size=max_batch_size):
        _run_hooks(chunk.input)
        parts.append(m.draw_noise(chunk.num_samples, dev))
    noises = [torch.cat([p[l] for p in parts]) if len(parts) > 1 else parts[0][l] for l in range(len(parts[0]))]
    with torch.no_grad():
        z = torch.stack(z_list)
    dloss = chunk_scales(vars, max_batch_size, z.device)
    n = z.shape[0]
    outs, losses, grads = [], [], []
    for lo in range(0, n, SG2_PHYS_BATCH):
        hi = min(n, lo + SG2_PHYS_BATCH)
        loss, dz, img = native.sg2_step(m.native, loss_fn.native_lpips(), tgt, z[lo:hi], [t[lo:hi] for t in noises],
                                        want_grad=optimize, grad_scale=1.0, dloss=dloss[lo:hi])
        outs.append(img)
        losses.append(loss)
        grads.append(dz)
    if optimize:
        dz = torch.cat(grads) if len(grads) > 1 else grads[0]
        for i, t in enumerate(z_list):
            if t.requires_grad:
                t.grad = dz[i]
        vars.opt.step()
        vars.opt.zero_grad()
    return (torch.cat(outs) if len(outs) > 1 else outs[0]), list((torch.cat(losses) if len(losses) > 1 else losses[0]).cpu().numpy()), {}


def step(model, vars, loss_fn, optimize=True, max_batch_size=9):
    """One evaluation (and, with ``optimize``, one gradient update) of every sample in ``vars``.

    Returns ``(outs [N,3,H,W], indiv_losses list[N], {})`` as the reference does."""
    if _native_pair(model, vars, loss_fn):
        if _uniform_outputs(loss_fn, vars.output):
            return _step_native(model, vars, loss_fn, optimize, max_batch_size)
        return _step_native_targets(model, vars, loss_fn, optimize, max_batch_size)  # per-sample targets / weights
    if _native_targets_pair(model, vars, loss_fn):
        return _step_native_targets(model, vars, loss_fn, optimize, max_batch_size)
    if _native_sg2_pair(model, vars, loss_fn):
        return _step_native_sg2(model, vars, loss_fn, optimize, max_batch_size)
    return _step_autograd(model, vars, loss_fn, optimize, max_batch_size)
