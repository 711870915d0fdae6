"""Projection / reconstruction / perceptual losses behind pix2latent's loss API, executed by the
native sm_100a library.

Same surface as /root/reference pix2latent/loss_functions.py: ``ProjectionLoss(lpips_net='alex',
beta=10)``, ``ReconstructionLoss(loss_type)``, ``PerceptualLoss(net, use_gpu)``, each called as
``loss(output, target, weight=None, loss_mask=None)`` and differentiable w.r.t. ``output``; the
small free functions (l1/l2, masked losses) are plain torch.

What differs from the reference's execution (not its results): the LPIPS branch of the constant
target is evaluated once per (target, weight, mask) and cached (`_TargetCache`), and the five
bilinear upsamples + weighted sums are folded into the feature-resolution distance kernel.
"""
import warnings

import torch
from torch import nn

from . import native
from .model import synth


def l1_loss(out, target):
    """|x - y|"""
    return torch.abs(target - out)


def l2_loss(out, target):
    """(x - y)^2"""
    return (target - out) ** 2


def _expand_like(t, ref):
    return t.repeat(ref.size(0), 1, 1, 1) if t.size(0) == 1 else t


def masked_l1_loss(out, target, mask):
    mask, target = _expand_like(mask, out), _expand_like(target, out)
    return torch.sum(l1_loss(out, target) * mask, [1, 2, 3]) / torch.sum(mask, [1, 2, 3])


def masked_l2_loss(out, target, mask):
    mask, target = _expand_like(mask, out), _expand_like(target, out)
    return torch.sum(l2_loss(out, target) * mask, [1, 2, 3]) / torch.sum(mask, [1, 2, 3])


def invertibility_loss(ims, target_transform, transform_params, mask=None):
    """MSE(ims, T^-1(T(ims)))"""
    if ims.size(0) == 1:
        ims = ims.repeat(len(transform_params), 1, 1, 1)
    inverted = target_transform(target_transform(ims, transform_params), transform_params, invert=True)
    if mask is None:
        return torch.mean((ims - inverted) ** 2, [1, 2, 3])
    return masked_l2_loss(ims, inverted, mask)


def weight_regularization(orig_model, curr_model, reg="l1", weight_dict=None):
    orig = orig_model.state_dict()
    total = 0.0
    for name, p in curr_model.named_parameters():
        if "bn" in name:
            continue
        d = p - orig[name]
        if reg == "l1":
            v = d.abs().mean()
        elif reg == "l2":
            v = (d ** 2).mean()
        elif reg == "inf":
            v = d.abs().max()
        else:
            raise ValueError(reg)
        total = total + (1.0 if weight_dict is None else weight_dict[name]) * v
    return total


# ------------------------------------------------------------------------------------ native part
_lpips_cache = {}


def _lpips_from_package(net):
    try:
        import lpips  # noqa: F401
    except ImportError:
        return None
    from .utils.misc import HiddenPrints
    with HiddenPrints():
        return synth.lpips_state_from_package(lpips.LPIPS(net=net, spatial=True).state_dict())


def get_native_lpips(net="alex", state_dict=None, seed=0, allow_synthetic=False):
    """One NativeLPIPS per (net, device, weights id). Weights (model/weights.py): ``state_dict`` if given; else a
    ``torch.save``d LPIPS state dict at ``$P2L_LPIPS_CKPT``; else the ``lpips`` package when importable; else — ONLY with
    ``allow_synthetic`` / ``P2L_ALLOW_SYNTHETIC=1`` — seeded synthetic weights; otherwise ``MissingWeights``."""
    import os
    from .model import weights
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    key = (net, dev, id(state_dict) if state_dict is not None else ("default", seed))
    if key not in _lpips_cache:
        if state_dict is None:
            state_dict, _ = weights.resolve("LPIPS(%s)" % net, None, [os.environ.get("P2L_LPIPS_CKPT")],
                                            lambda: _lpips_from_package(net), lambda: synth.lpips_state_dict(net, seed),
                                            allow_synthetic, post=synth.lpips_state_from_package)
        sd = {k: v.cuda() for k, v in state_dict.items()}
        _lpips_cache[key] = native.NativeLPIPS(net, sd)
    return _lpips_cache[key]


class _TargetCache:
    """(target, weight, mask) -> NativeTarget, keyed by tensor identity + version counters."""

    def __init__(self, lp, rec_type, rec_weight, per_weight, capacity=4):
        self.lp, self.rec_type, self.rec_weight, self.per_weight = lp, rec_type, rec_weight, per_weight
        self.capacity = capacity
        self.entries = []  # (key, NativeTarget, refs that keep the id()s alive)

    @staticmethod
    def _key(*ts):
        return tuple(None if t is None else (t.data_ptr(), t._version, tuple(t.shape)) for t in ts)

    def get(self, target, weight, mask):
        key = self._key(target, weight, mask)
        for k, tgt, _ in self.entries:
            if k == key:
                return tgt
        tgt = self.lp.make_target(target, weight, mask, self.rec_type, self.rec_weight, self.per_weight)
        # detached aliases pin the STORAGE the key's pointers refer to (a leaf whose `.data` is rebound later,
        # base_optimizer.py apply_transform, would otherwise let the allocator reuse the address under a live key)
        self.entries.append((key, tgt, tuple(None if t is None else t.detach() for t in (target, weight, mask))))
        if len(self.entries) > self.capacity:
            self.entries.pop(0)
        return tgt


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, tgt, want_grad):
        ctx.tgt = tgt
        ctx.b = output.shape[0]
        return tgt.loss_forward(output.detach(), want_grad)

    @staticmethod
    def backward(ctx, dloss):
        return ctx.tgt.loss_backward(ctx.b, dloss.contiguous()), None, None


def _same_rows(t):
    return t is None or t.size(0) == 1 or bool((t[1:] == t[:1]).all())


class _NativeLoss(nn.Module):
    """Common machinery: rec_weight * pixel term + per_weight * LPIPS term, natively."""

    def __init__(self, net, rec_type, rec_weight, per_weight, lpips_state_dict=None, allow_synthetic=False):
        super().__init__()
        self._net, self._rec_type = net, rec_type
        self._rec_weight, self._per_weight = rec_weight, per_weight
        self._lpips_state = lpips_state_dict
        # a pure pixel loss never runs the perceptual net: its (unused) weights may be anything
        self._allow_synthetic = allow_synthetic or per_weight == 0.0
        self._cache = None

    def native_lpips(self):
        return get_native_lpips(self._net, self._lpips_state, allow_synthetic=self._allow_synthetic)

    def target_cache(self):
        if self._cache is None:
            self._cache = _TargetCache(self.native_lpips(), self._rec_type, self._rec_weight, self._per_weight)
        return self._cache

    def invalidate_targets(self):
        """Drop every cached prepared target. The caches are keyed by (storage pointer, version counter, shape); a
        buffer rewritten through ``.data`` (``t.data.copy_(...)``) keeps all three, so whoever edits a registered
        target / weight / loss_mask IN PLACE calls this (``_BaseOptimizer.apply_transform`` does)."""
        self._cache = None
        self._row_cache = None
        self.__dict__.pop("_uniform_cache", None)

    def prepared_target(self, target, weight=None, loss_mask=None):
        """target/weight/loss_mask: single [3,H,W] tensors -> cached NativeTarget."""
        return self.target_cache().get(target, weight, loss_mask)

    def prepared_targets(self, targets, weights=None, loss_masks=None):
        """Per-candidate targets (lists of [3,H,W] tensors; transform search): one cached NativeTarget per
        row, rows with identical storage share one."""
        if getattr(self, "_row_cache", None) is None:
            self._row_cache = _TargetCache(self.native_lpips(), self._rec_type, self._rec_weight, self._per_weight,
                                           capacity=96)
        n = len(targets)
        return [self._row_cache.get(targets[i], None if weights is None else weights[i],
                                    None if loss_masks is None else loss_masks[i]) for i in range(n)]

    def __call__(self, output, target, weight=None, loss_mask=None):
        if not output.is_cuda:
            raise RuntimeError("pix2latent_b200 losses run on a CUDA (sm_100a) device only")
        b = output.size(0)
        want_grad = torch.is_grad_enabled() and output.requires_grad
        shared = _same_rows(target) and _same_rows(weight) and _same_rows(loss_mask)
        if shared:
            tgt = self.prepared_target(target[0], None if weight is None else weight[0],
                                       None if loss_mask is None else loss_mask[0])
            loss = _LossFn.apply(output, tgt, want_grad)
        else:
            # per-sample targets (e.g. after a spatial transform of the target): one row at a time
            rows = []
            for i in range(b):
                tgt = self.lp_target_uncached(target[i], None if weight is None else weight[i],
                                              None if loss_mask is None else loss_mask[i])
                rows.append(_LossFn.apply(output[i:i + 1], tgt, want_grad))
            loss = torch.cat(rows)
        if weight is None and self._returns_map_without_weight:
            # the reference returns the un-reduced map here and closure.py:55 takes its mean;
            # the mean is what the native reduction with W = 1 already is. Shape it so that
            # `.view(b, -1).mean(1)` is the identity.
            return loss.view(b, 1)
        return loss

    _returns_map_without_weight = True

    def lp_target_uncached(self, target, weight, mask):
        return self.native_lpips().make_target(target, weight, mask, self._rec_type, self._rec_weight,
                                               self._per_weight)


class ReconstructionLoss(_NativeLoss):
    """Weighted L1 / L2 pixel loss: sum(|t - o| * W) / sum(W) per sample."""

    def __init__(self, loss_type="l1"):
        if loss_type in ["l1", 1]:
            rt = 1
        elif loss_type in ["l2", 2]:
            rt = 2
        else:
            raise ValueError("Unknown loss_type {}".format(loss_type))
        super().__init__("alex", rt, 1.0, 0.0)
        self.loss_fn = l1_loss if rt == 1 else l2_loss


class PerceptualLoss(_NativeLoss):
    """LPIPS(net, spatial=True) with spatial weighting: sum(map * W) / sum(W) per sample."""

    def __init__(self, net="vgg", use_gpu=True, lpips_state_dict=None, allow_synthetic=False):
        super().__init__(net, 1, 0.0, 1.0, lpips_state_dict, allow_synthetic)


class ProjectionLoss(_NativeLoss):
    """The paper's default: weighted L1 + beta * weighted LPIPS."""

    def __init__(self, lpips_net="alex", beta=10, lpips_state_dict=None, allow_synthetic=False):
        super().__init__(lpips_net, 1, 1.0, float(beta), lpips_state_dict, allow_synthetic)
        self.beta = beta
