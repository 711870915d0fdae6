"""Shared host loop machinery of all optimizers (reference: pix2latent/optimizer/base_optimizer.py
:9-141): holds model / loss / variable manager, runs the optional target transforms, tracks the
per-step inputs on the CPU and delegates the evaluation to ``closure.step``."""
import os
import time

import numpy as np
import torch

from .. import parallel
from ..utils.image import to_grid, to_image
from ..utils.misc import progress_print
from .closure import step


# Device-resident inner loop: runs of gradient steps go through ONE C-ABI call
# (p2l_biggan_optimize: Clamp hook + fused step + per-candidate Adam, CUDA-graph replayed) when the
# run is expressible there — see _BaseOptimizer._fusable. P2L_FUSE_INNER_LOOP=0 turns it off.
FUSE_INNER_LOOP = os.environ.get("P2L_FUSE_INNER_LOOP", "1") != "0"


def _clamp_of(hook_fn):
    """Bound of a hook that is None / Clamp / Compose of Clamps; False when it is anything else."""
    from ..utils import function_hooks as hk
    if hook_fn is None:
        return 0.0
    if type(hook_fn) is hk.Clamp:
        return float(hook_fn.trunc) if hook_fn.trunc > 0 else False
    if type(hook_fn) is hk.Compose and len(hook_fn.hook_fns) > 0:
        bounds = [_clamp_of(h) for h in hook_fn.hook_fns]
        if any(b is False or b == 0.0 for b in bounds):
            return False
        return min(bounds)
    return False


class _BaseOptimizer():

    def __init__(self, model, var_manager, loss_fn, max_batch_size=9, log=False, track_variables=True, **kwargs):
        """
        Args
            model: the generator to invert (callable with the registered input names as kwargs)
            var_manager (VariableManager)
            loss_fn: loss(out, **registered output variables) -> per-sample loss
            max_batch_size (int): mini-batch size of the evaluation (the 1/b gradient scale of
                ``loss.mean().backward()`` follows it)
        """
        self.max_batch_size = max_batch_size
        self.model = model.eval()
        self.var_manager = var_manager
        self.loss_fn = loss_fn
        self.transform_fns = {}
        self.log = log
        self.log_iter = 5
        self.show_iter = 50
        self.log_resize_factor = None
        self.track_variables = track_variables
        self.tracked = {}
        self.fuse_inner_loop = FUSE_INNER_LOOP
        self.fused_calls = 0  # runs of gradient steps that went through the device-resident loop

    def register_benchmark(self, benchmark):
        self.bm = benchmark

    def register_transform(self, transform_fn, tranform_var_name, target_var_name):
        """Before optimising, ``transform_fn(target_var, transform_var)`` replaces the target variable."""
        self.transform_fns[target_var_name] = {
            "fn": transform_fn, "transform_param": tranform_var_name, "target_var": target_var_name}

    def apply_transform(self, variables, transform_dict):
        info = self.var_manager.variable_info
        src, dst = transform_dict["transform_param"], transform_dict["target_var"]
        src_data = torch.stack(variables[info[src]["var_type"]][src].data)
        dst_list = variables[info[dst]["var_type"]][dst].data
        new = list(transform_dict["fn"](torch.stack(dst_list), src_data))
        for i, t in enumerate(new):
            dst_list[i].data = t.data
        if hasattr(self.loss_fn, "invalidate_targets"):
            self.loss_fn.invalidate_targets()  # prepared targets are keyed by storage: drop what the old tensors cached

    def step(self, variables, optimize=True, transform=False):
        if transform and len(self.transform_fns) > 0:
            for td in self.transform_fns.values():
                self.apply_transform(variables, td)
        if self.track_variables:
            self.track(variables)
        rank, size = parallel.world()
        # one process per GPU: evaluate this rank's candidates only; nothing is communicated here
        local = parallel.shard_vars(variables, rank, size) if size > 1 else variables
        self.out, self.loss, self.other = step(
            self.model, local, loss_fn=self.loss_fn, optimize=optimize, max_batch_size=self.max_batch_size)
        self._n_total = variables.num_samples
        return self.out, self.loss, self.other

    # ---- runs of gradient steps --------------------------------------------------------------
    def grad_steps(self, variables, n_steps, on_step=None):
        """``n_steps`` gradient updates of ``variables`` — the loop body every optimizer of the reference
        repeats (`self.step(variables, optimize=True, transform=(j == 0))`, e.g. basincma_optimizer.py:60-66);
        ``on_step(j)`` runs after update j (logging / progress). Uses the device-resident loop when the
        run is expressible there, the per-step path otherwise; both produce the same trajectory."""
        plan = self._fusable(variables, n_steps)
        if plan is not None:
            self._fused_steps(variables, n_steps, plan)
            for j in range(n_steps):
                if on_step is not None:
                    on_step(j)
            return
        for j in range(n_steps):
            self.step(variables, optimize=True, transform=(j == 0))
            if on_step is not None:
                on_step(j)

    def _fusable(self, variables, n_steps):
        """Hyper-parameters of the fused run, or None when this run must go step by step: needs the
        library's own BigGAN + loss, plain torch.optim.Adam over the z / c leaves (one lr per variable, no
        weight decay / amsgrad), hooks that are None or Clamp, no per-step logging and no transform."""
        from .closure import _native_pair
        if not self.fuse_inner_loop or n_steps < 2 or self.log or len(self.transform_fns) > 0:
            return None
        if not _native_pair(self.model, variables, self.loss_fn):
            return None
        from .closure import _uniform_outputs
        if not _uniform_outputs(self.loss_fn, variables.output):
            return None  # per-sample targets / weights: the per-step path with one target per candidate
        from .closure import adam_plan
        z_list, c_list = variables.input.z.data, variables.input.c.data
        plan = adam_plan(variables.opt, z_list, c_list)
        if plan is None:
            return None
        if plan["n_owned"] != sum(1 for lst in (z_list, c_list) for t in lst if id(t) in plan["stateful"]):
            return None  # the optimizer also owns parameters other than z / c
        clamps = [_clamp_of(variables.input.z.hook_fn), _clamp_of(variables.input.c.hook_fn)]
        if any(c is False for c in clamps):
            return None
        plan["clamp_z"], plan["clamp_c"] = clamps
        return plan

    @torch.no_grad()
    def _fused_steps(self, variables, n_steps, plan):
        from .. import native
        from .closure import LazyLosses, _unwrap, chunk_scales, native_adam
        rank, size = parallel.world()
        local = parallel.shard_vars(variables, rank, size) if size > 1 else variables
        m = _unwrap(self.model)
        z_list, c_list = local.input.z.data, local.input.c.data
        n = len(z_list)
        z = torch.stack(z_list).float().contiguous()
        c = torch.stack(c_list).float().contiguous()
        dev = z.device
        first = {k: v.data[0] for k, v in local.output.items()}
        tgt = self.loss_fn.prepared_target(first["target"], first.get("weight"), first.get("loss_mask"))
        dloss = chunk_scales(local, self.max_batch_size, dev)
        # the Adam state (moments [n, dim], step count) lives on the device and is shared with the per-step path
        ad = native_adam(variables.opt, z_list, c_list)
        cfg = native.adam_config(plan["lr_z"], plan["lr_c"], plan["betas"], plan["eps"], plan["clamp_z"], plan["clamp_c"])
        res = native.biggan_optimize(m.native, self.loss_fn.native_lpips(), tgt, z, c, n_steps, cfg, state=ad["state"],
                                     dloss=dloss, grad_scale=1.0, track=self.track_variables)
        ad["steps"] += n_steps
        torch._foreach_copy_([t.data.view(-1) for t in z_list], list(z.unbind(0)))
        torch._foreach_copy_([t.data.view(-1) for t in c_list], list(c.unbind(0)))
        if self.track_variables:
            lo, hi = parallel.shard_bounds(variables.num_samples, rank, size) if size > 1 else (0, n)
            for name, hist in (("z", res["z_hist"]), ("c", res["c_hist"])):
                hist = hist.cpu()
                if size > 1:
                    base = torch.stack(variables.input[name].data).cpu()
                for j in range(n_steps):
                    if size > 1:
                        full = base.clone()
                        full[lo:hi] = hist[j]
                    else:
                        full = hist[j].clone()
                    self.tracked.setdefault(name, []).append(full)
        self.out = res["img"]
        self.loss = LazyLosses(res["loss"][-1])
        self.loss_history = res["loss"]  # [n_steps, n_local] device tensor of every step's losses
        self.other = {}
        self._n_total = variables.num_samples
        self.fused_calls += 1
        self.fused_graph = res["graph"]

    def gathered_loss(self):
        """Per-candidate losses of the whole population (one all_gather of scalars when sharded)."""
        return parallel.allgather_losses(self.loss, self._n_total)

    def track(self, variables):
        for name, var in variables.input.items():
            self.tracked.setdefault(name, []).append(torch.stack(var.data).cpu().detach().clone())

    def optimize(self):
        raise NotImplementedError

    def log_result(self, variables, step_iter):
        if hasattr(self, "bm"):
            res = self.bm.evaluate(self.out, variables.output.target.data[0].unsqueeze(0),
                                   variables.output.weight.data[0].unsqueeze(0))
        else:
            res = {"loss": np.array(self.gathered_loss())}
        self.losses.append([step_iter, res])
        collage = to_image(to_grid(self.out.cpu()), cv2_format=False)
        if self.log_resize_factor is not None:
            import cv2
            collage = cv2.resize(np.array(collage, dtype=np.uint8), None, fx=self.log_resize_factor,
                                 fy=self.log_resize_factor, interpolation=cv2.INTER_AREA)
        self.outs.append(collage)

    # ---- helpers shared by the concrete loops ------------------------------------------------
    def _start_run(self):
        self.losses, self.outs = [], []
        if hasattr(self.loss_fn, "invalidate_targets"):
            # a new run: targets / weights may have been rewritten in place through .data since the last one (such writes do
            # not bump the tensor version the cache is keyed on); re-preparing the target costs one LPIPS forward
            self.loss_fn.invalidate_targets()
        self._t_mark = time.time()

    def _maybe_log(self, variables, log_at, log_last):
        if self.log and ((log_at % self.log_iter == 0) or (log_at == log_last)):
            self.log_result(variables, log_at)

    def _progress(self, i, total_steps, shown, pbar):
        if pbar is not None:
            pbar.progress(i / total_steps)
        elif shown % self.show_iter == 0:
            progress_print("optimize", shown, total_steps, "c", (time.time() - self._t_mark) / self.show_iter)
            self._t_mark = time.time()

    def _after_step(self, i, total_steps, log_at, log_last, pbar):
        """Logging / progress after a gradient step (the reference's conditions)."""
        self._maybe_log(self._variables, log_at, log_last)
        self._progress(i, total_steps, log_at, pbar)

    def sync_inputs(self, variables):
        """Candidate sharding: every rank refines only its shard; bring the current latents of ALL candidates to every
        rank (an all_gather of N x dim floats — KBs). No-op in a single process."""
        rank, size = parallel.world()
        if size <= 1:
            return
        n = variables.num_samples
        lo, hi = parallel.shard_bounds(n, rank, size)
        with torch.no_grad():
            for var in variables.input.values():
                full = parallel.allgather_rows(torch.stack(var.data[lo:hi]), n)
                for i, t in enumerate(var.data):
                    t.data.copy_(full[i])

    def _finish(self, variables, total_steps):
        from .closure import flush_native_adam
        flush_native_adam(variables.opt)  # torch's optimizer state is current again when optimize() returns
        rank, size = parallel.world()
        if size > 1:
            # final state of every shard to every rank: latents (KBs), losses, images
            n = variables.num_samples
            self.sync_inputs(variables)
            self.loss = self.gathered_loss()
            self.out = parallel.allgather_rows(self.out, n)
        if self.log:
            return variables, self.outs, self.losses
        grid = to_grid(torch.stack(list(self.out.cpu().detach())))
        return variables, [grid], [[total_steps, {"loss": list(self.loss)}]]
