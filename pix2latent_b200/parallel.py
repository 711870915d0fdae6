"""Candidate sharding across the GPUs of one box.

The population is embarrassingly parallel: candidates interact only through the host-side
ask/tell of CMA / Nevergrad and through the 1/b_chunk factor of ``loss.mean()``, which depends on
the chunk SIZE only. One process per GPU (torchrun); rank r owns candidates
[r*N/R, (r+1)*N/R); rank 0 owns the search state:

    ask   : rank 0 asks, ONE broadcast of z[N, dim]
    steps : each rank runs closure.step on its own shard — no communication
    tell  : ONE all_gather of the per-candidate scalar losses (N/R floats per rank), rank 0 tells

Backends: NCCL over NVLink on GPUs, gloo on CPU (tests). The reference's only multi-GPU path,
nn.DataParallel around the StyleGAN2 module (examples/invert_stylegan2_cars_basincma.py:51),
re-broadcasts ~120 MB of weights every forward and is not reproduced.
"""
import numpy as np
import torch
import torch.distributed as dist

from .variable_manager import AttrDict


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, size):
    """Contiguous, as-even-as-possible split; the first n % size ranks get one more."""
    base, extra = divmod(n, size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_vars(vars, rank, size):
    """View of ``vars`` restricted to this rank's candidates (same tensors, same optimizer)."""
    lo, hi = shard_bounds(vars.num_samples, rank, size)
    out = {}
    for var_type, group in vars.items():
        if var_type in ("opt", "num_samples", "shard"):
            continue
        out[var_type] = {}
        for name, entry in group.items():
            e = dict(entry)
            e["data"] = entry.data[lo:hi]
            out[var_type][name] = e
    out["opt"] = vars.opt
    out["num_samples"] = hi - lo
    out["shard"] = (lo, hi, vars.num_samples)  # closure.chunk_scales: the 1/b_chunk scales follow the GLOBAL chunking
    return AttrDict(out)


def _comm_device(t=None):
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_array(arr, shape=None, dtype=np.float64):
    """numpy array from rank 0 to every rank."""
    rank, size = world()
    if size == 1:
        return arr
    dev = _comm_device()
    if rank == 0:
        t = torch.as_tensor(np.ascontiguousarray(arr, dtype=dtype)).to(dev)
        meta = torch.tensor(list(t.shape) + [-1] * (4 - t.dim()), dtype=torch.long, device=dev)
    else:
        meta = torch.empty(4, dtype=torch.long, device=dev)
    dist.broadcast(meta, 0)
    shp = [int(x) for x in meta.tolist() if x >= 0]
    if rank != 0:
        t = torch.empty(shp, dtype=torch.as_tensor(np.zeros(1, dtype=dtype)).dtype, device=dev)
    dist.broadcast(t, 0)
    return t.cpu().numpy()


def allgather_losses(local, n_total):
    """Concatenate per-candidate losses of all ranks in candidate order -> list of n_total floats.
    The single data-path collective of the search loop (a few hundred bytes). On NCCL the losses stay on the device:
    the step's loss tensor (closure.LazyLosses) is the send buffer, ONE all_gather, ONE device -> host copy."""
    rank, size = world()
    if size == 1:
        return list(local)
    dev = _comm_device()
    counts = [shard_bounds(n_total, r, size) for r in range(size)]
    width = max(hi - lo for lo, hi in counts)
    src = local.device_tensor() if hasattr(local, "device_tensor") else None
    if src is not None and src.device == dev:
        loc = src.detach().float()
    else:
        loc = torch.as_tensor(np.asarray(local, dtype=np.float32)).to(dev)
    if loc.numel() == width:
        buf = loc.contiguous()
    else:
        buf = torch.zeros(width, dtype=torch.float32, device=dev)
        buf[: loc.numel()] = loc
    out = torch.empty(size * width, dtype=torch.float32, device=dev)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out, buf)
    else:
        dist.all_gather(list(out.view(size, width).unbind(0)), buf)
    host = out.view(size, width).cpu().numpy()
    res = []
    for r, (lo, hi) in enumerate(counts):
        res.extend(host[r, : hi - lo].tolist())
    return [np.float32(x) for x in res]


def allgather_rows(local, n_total):
    """Gather row-sharded tensors [n_local, ...] into [n_total, ...] on every rank."""
    rank, size = world()
    if size == 1:
        return local
    dev = _comm_device()
    counts = [shard_bounds(n_total, r, size) for r in range(size)]
    width = max(hi - lo for lo, hi in counts)
    src = local.detach().to(dev)
    buf = torch.zeros((width,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
    buf[: src.shape[0]] = src
    out = [torch.empty_like(buf) for _ in range(size)]
    dist.all_gather(out, buf)
    return torch.cat([out[r][: hi - lo] for r, (lo, hi) in enumerate(counts)]).to(local.device)
