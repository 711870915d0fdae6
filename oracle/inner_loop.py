"""ORACLE (test infrastructure) — CPU restatement of the device-resident inner loop
(p2l_biggan_optimize, pix2latent_b200/csrc/optim.cu): ``steps`` repetitions of what the reference
does per mini-batch and step, for a whole population at once.

Follows /root/reference:
  pix2latent/optimizer/base_optimizer.py:94-97,105-106  track(): inputs cloned BEFORE the step's hooks
  pix2latent/optimizer/closure.py:42-44                  hooks in place before the forward
  pix2latent/utils/function_hooks.py:24-27               Clamp: clamp_(-trunc, trunc)
  pix2latent/optimizer/closure.py:51-58                  out = model(..); loss = loss_fn(..).view(b,-1).mean(1);
                                                         loss.mean().backward()  (per-sample scale = dloss)
  pix2latent/optimizer/closure.py:65                     opt.step(): torch.optim.Adam, one param group per
                                                         latent tensor (variable_manager.py:231-238)
and the installed torch's Adam arithmetic (torch/optim/adam.py _single_tensor_adam: lerp_,
mul_/addcmul_, bias corrections in double, addcdiv_).

PINNED: tests/test_fused_host_cpu.py checks this restatement against torch.optim.Adam driven by the
product's per-step path (itself checked against the real reference's closure.step golden vectors in
tests/test_golden_cpu.py) on the same inputs.
"""
import math

import torch


def adam_update(p, g, m, v, t, lr, beta1, beta2, eps):
    """One Adam update of p (in place) with gradient g; t = step number (1-based)."""
    m.add_((g - m) * (1.0 - beta1))
    v.mul_(beta2).add_(g * g * (1.0 - beta2))
    bc1 = 1.0 - math.pow(beta1, t)
    bc2 = 1.0 - math.pow(beta2, t)
    step_size = lr / bc1
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p.sub_(step_size * (m / denom))


def run(step_fn, z, c, steps, lr_z, lr_c, betas=(0.9, 0.999), eps=1e-8, clamp_z=0.0, clamp_c=0.0,
        dloss=None, grad_scale=1.0, state=None, track=False):
    """``step_fn(z, c) -> (loss[b], img)`` differentiable in z, c. z [b,zd], c [b,cd] are updated in
    place. state: dict(m_z, v_z, m_c, v_c, step) or None (fresh). Returns dict(loss [steps,b],
    z_hist, c_hist, img, state)."""
    b = z.shape[0]
    if state is None:
        state = dict(m_z=torch.zeros_like(z), v_z=torch.zeros_like(z), m_c=torch.zeros_like(c),
                     v_c=torch.zeros_like(c), step=0)
    up = torch.full((b,), float(grad_scale), dtype=z.dtype) if dloss is None else dloss.to(z.dtype) * grad_scale
    losses, zh, ch, img = [], [], [], None
    for _ in range(steps):
        if track:
            zh.append(z.detach().clone())
            ch.append(c.detach().clone())
        with torch.no_grad():
            if clamp_z > 0:
                z.clamp_(-clamp_z, clamp_z)
            if clamp_c > 0:
                c.clamp_(-clamp_c, clamp_c)
        with torch.enable_grad():
            zz = z.detach().clone().requires_grad_(True)
            cc = c.detach().clone().requires_grad_(True)
            loss, img = step_fn(zz, cc)
            (loss * up).sum().backward()
        losses.append(loss.detach().clone())
        state["step"] += 1
        with torch.no_grad():
            adam_update(z, zz.grad, state["m_z"], state["v_z"], state["step"], lr_z, betas[0], betas[1], eps)
            adam_update(c, cc.grad, state["m_c"], state["v_c"], state["step"], lr_c, betas[0], betas[1], eps)
    return dict(loss=torch.stack(losses) if losses else torch.zeros(0, b), z_hist=torch.stack(zh) if zh else None,
                c_hist=torch.stack(ch) if ch else None, img=None if img is None else img.detach(), state=state)
