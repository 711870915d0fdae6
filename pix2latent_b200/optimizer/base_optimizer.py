"""Shared host loop machinery of all optimizers (reference: pix2latent/optimizer/base_optimizer.py
:9-141): holds model / loss / variable manager, runs the optional target transforms, tracks the
per-step inputs on the CPU and delegates the evaluation to ``closure.step``."""
import time

import numpy as np
import torch

from .. import parallel
from ..utils.image import to_grid, to_image
from ..utils.misc import progress_print
from .closure import step


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

    def _finish(self, variables, total_steps):
        rank, size = parallel.world()
        if size > 1:
            # final state of every shard to every rank: latents (KBs), losses, images
            n = variables.num_samples
            lo, hi = parallel.shard_bounds(n, rank, size)
            with torch.no_grad():
                for var in variables.input.values():
                    full = parallel.allgather_rows(torch.stack(var.data[lo:hi]), n)
                    for i, t in enumerate(var.data):
                        t.data.copy_(full[i])
            self.loss = self.gathered_loss()
            self.out = parallel.allgather_rows(self.out, n)
        if self.log:
            return variables, self.outs, self.losses
        grid = to_grid(torch.stack(list(self.out.cpu().detach())))
        return variables, [grid], [[total_steps, {"loss": self.loss}]]
