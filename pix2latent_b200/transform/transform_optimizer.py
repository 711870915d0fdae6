"""BasinCMA over a target transformation (reference: pix2latent/transform/transform_optimizer.py:21-255):
the outer CMA searches the transformation parameter ``t`` (a 'transform' variable), every meta-iteration
re-transforms each candidate's target / weight with its own ``t`` (base_optimizer.apply_transform on the first
inner step) and refines (z, c) by gradient descent; CMA is told the loss measured in the UN-transformed frame
(base_cma_optimizer.cma_update(inverted_loss=True)). "Variable propagation" restarts the latents of the next
meta-iteration around a moving average of the best latents so far.

Execution: per-candidate targets go through ONE native call per step (closure._step_native_targets ->
p2l_biggan_step_targets); the resampling is p2l_affine_resample."""
import time

import numpy as np
import torch

from ..optimizer.base_cma_optimizer import _BaseCMAOptimizer
from ..optimizer.base_optimizer import _BaseOptimizer
from ..utils.image import to_grid, to_image
from ..utils.misc import progress_print


class TransformBasinCMAOptimizer(_BaseOptimizer, _BaseCMAOptimizer):

    def __init__(self, *args, **kwargs):
        _BaseOptimizer.__init__(self, *args, **kwargs)
        _BaseCMAOptimizer.__init__(self)
        self.variables_to_propagate = []

    @torch.no_grad()
    def vis_transform(self, variables):
        target = torch.stack(variables.output.target.data)
        weight = torch.stack(variables.output.weight.data)
        im = to_image(to_grid((target * weight).cpu()), cv2_format=False)
        if self.log_resize_factor is not None:
            import cv2
            im = cv2.resize(np.array(im, dtype=np.uint8), None, fx=self.log_resize_factor, fy=self.log_resize_factor,
                            interpolation=cv2.INTER_AREA)
        self.transform_outs.append(im)

    def set_variable_propagation(self, variable_name):
        """propagate this input variable from one meta-iteration to the next"""
        if variable_name in self.variables_to_propagate:
            print("variable {} already exists".format(variable_name))
            return
        self.variables_to_propagate.append(variable_name)

    def del_variable_propagation(self, variable_name):
        # (the reference tests `in` where it means `not in`, transform_optimizer.py:68-70; intended behaviour here)
        if variable_name not in self.variables_to_propagate:
            print("variable {} is not propagated".format(variable_name))
            return
        self.variables_to_propagate.remove(variable_name)

    def _propagated(self, variables):
        for name in self.variables_to_propagate:
            if name not in variables.input:
                raise RuntimeError("variable propagation is set for {} but no such variable was found".format(name))
            entry = variables.input[name]
            if name not in self.vp_means:
                self.vp_means[name] = torch.stack(entry.data).mean(0)
            yield name, entry

    @torch.no_grad()
    def update_propagation_variable_statistic(self, variables, ema_beta=0.5):
        """moving average towards the seed that currently performs best (ema_beta = 1 forgets the past)"""
        from .. import parallel
        losses = self.loss
        if parallel.world()[1] > 1:
            # candidates are sharded: the best seed may live on another rank
            self.sync_inputs(variables)
            losses = self.gathered_loss()
        for name, entry in self._propagated(variables):
            best = entry.data[int(np.argmin(losses))]
            self.vp_means[name] = (1.0 - ema_beta) * self.vp_means[name] + ema_beta * best

    @torch.no_grad()
    def propagate_variable(self, variables, curr_iter, total_iter, magnitude=1.0, renormalize=True):
        """resample the propagated variables around their moving average; the noise shrinks linearly with the
        progress; ``renormalize`` standardises every sample to zero mean / unit std"""
        sigma = magnitude * (1 - (curr_iter / float(total_iter)))
        for name, entry in self._propagated(variables):
            for i in range(len(entry.data)):
                fresh = (self.vp_means[name] + sigma * torch.randn_like(entry.data[i])).data
                if renormalize:
                    fresh = (fresh - fresh.mean()) / fresh.std()
                entry.data[i].data = fresh

    def get_candidate(self):
        return self._candidate

    # ---- the search ---------------------------------------------------------------------------------------------
    def _reset(self):
        self.losses, self.outs, self.transform_outs = [], [], []
        self._best_loss, self._candidate = 999, None
        self.vp_means = {}
        self.transform_tracked = []

    def _refine(self, variables, n_steps, done, total_steps, grad_steps, pbar, t_mark):
        """n_steps gradient updates of (z, c) against this meta-iteration's transformed targets; returns (steps done, t_mark)"""
        for j in range(n_steps):
            self.step(variables, optimize=True, transform=(j == 0))   # j == 0: every candidate's target / weight is resampled
            done += 1
            if self.log:
                if j == 0:
                    self.vis_transform(variables)
                if done % self.log_iter == 0 or done == grad_steps:
                    self.log_result(variables, done)
            if pbar is not None:
                pbar.progress(done / total_steps)
            elif done % self.show_iter == 0:
                progress_print("optimize", done, total_steps, "c", (time.time() - t_mark) / self.show_iter)
                t_mark = time.time()
        return done, t_mark

    def _remember_best(self, variables, loss):
        best = int(np.argmin(loss))
        if loss[best] < self._best_loss:
            self._candidate = variables.transform.t.data[best].cpu().detach()
            self._best_loss = loss[best]

    def optimize(self, meta_steps, grad_steps, last_grad_steps=None, pbar=None):
        """
        Args
            meta_steps (int): CMA updates of the transformation parameter
            grad_steps (int): gradient updates of the latents per CMA update
            last_grad_steps (int): gradient updates of the last meta-iteration (default: grad_steps)
        Returns (variables, (outs, transformed targets, best candidate's target), losses)
        """
        self.setup_cma(self.var_manager)
        self._reset()
        last_grad_steps = grad_steps if last_grad_steps is None else last_grad_steps
        total_steps = (meta_steps - 1) * grad_steps + last_grad_steps
        done, t_mark, loss = 0, time.time(), None
        for meta_iter in range(meta_steps):
            final = meta_iter == meta_steps - 1
            variables = self._variables = self.cma_init(self.var_manager)
            if meta_iter > 0:
                self.propagate_variable(variables, meta_iter, meta_steps)
            self.transform_tracked.append(torch.stack(variables.transform.t.data).cpu().detach().clone())
            done, t_mark = self._refine(variables, last_grad_steps if final else grad_steps, done, total_steps, grad_steps,
                                        pbar, t_mark)
            if not final:
                loss = np.asarray(self.cma_update(variables, inverted_loss=True))
            elif loss is None:
                loss = np.asarray(self.loss)  # a single meta-iteration: the reference would stop on an unbound name here
            # (after the last meta-iteration `loss` still holds the previous CMA update's losses, as in the reference)
            self.update_propagation_variable_statistic(variables)
            self._remember_best(variables, loss)
        candidate_out = variables.output.target.data[int(np.argmin(loss))]
        if self.log:
            return variables, (self.outs, self.transform_outs, candidate_out), self.losses
        from .. import parallel
        if parallel.world()[1] > 1:
            self.sync_inputs(variables)
            self.loss = self.gathered_loss()
            self.out = parallel.allgather_rows(self.out, variables.num_samples)
        transform_target = to_grid(torch.stack(variables.output.target.data).cpu())
        transform_out = to_grid(torch.stack(list(self.out.cpu().detach())))
        return variables, ([transform_out], [transform_target], candidate_out), list(self.loss)
