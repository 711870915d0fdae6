"""Pre-alignment statistics and transform composition (reference:
pix2latent/transform/transform_utils.py:53-184). Host arithmetic on a handful of numbers."""
import numpy as np
import torch

from ..utils.image import binarize

# BigGAN's typical object box, fractions of the image: centre of mass (y, x) and extent (h, w)
_BIGGAN_CENTER = (137 / 255., 127 / 255.)
_BIGGAN_EXTENT = (213 / 255., 210 / 255.)


def get_biggan_stats():
    """precomputed BigGAN object statistics: (centre of mass, object size)"""
    return list(_BIGGAN_CENTER), list(_BIGGAN_EXTENT)


def bbox_from_mask(mask):
    """(top, left, bottom, right) of the non-zero region of a [c,h,w] mask (last non-zero row / column, inclusive);
    the whole range when the mask is empty."""
    assert mask.dim() == 3, "expected 3d tensor but got {}".format(mask.dim())
    plane = mask.mean(0)
    out = []
    for axis_sum, full in ((plane.sum(1), mask.size(1)), (plane.sum(0), mask.size(2))):
        hit = torch.nonzero(axis_sum != 0).flatten()
        out.append((int(hit[0]), int(hit[-1])) if hit.numel() else (0, full))
    (top, bottom), (left, right) = out
    return top, left, bottom, right


def compute_stat_from_mask(mask):
    """Binary mask [c,h,w] -> ((cy, cx), (h, w)) of its bounding box as fractions of the image."""
    top, left, bottom, right = bbox_from_mask(mask)
    H, W = mask.size(1), mask.size(2)
    h, w = bottom - top, right - left
    return ((top + h // 2) / H, (left + w // 2) / W), (h / H, w / W)


def convert_to_t(src_center, src_size, dst_center, dst_size):
    """[s, tx, ty] that takes an object of (src_center, src_size) to (dst_center, dst_size): the scale follows the
    object's larger side, the shift is in affine_grid's [-1, 1] units, x first."""
    src_c, src_s = np.asarray(src_center, dtype=np.float64), np.asarray(src_size, dtype=np.float64)
    dst_c, dst_s = np.asarray(dst_center, dtype=np.float64), np.asarray(dst_size, dtype=np.float64)
    side = int(np.argmax(src_s))
    shift_yx = 2.0 * (src_c - dst_c)
    return torch.tensor([src_s[side] / dst_s[side], shift_yx[1], shift_yx[0]]).float()


def compute_pre_alignment(weight):
    """Initial transformation parameter that moves the masked object onto BigGAN's typical object box."""
    src = compute_stat_from_mask(binarize(weight))
    return convert_to_t(src[0], src[1], *get_biggan_stats()).numpy()


class ComposeTransform():
    """Chain of transform functions. Entries are functions or (function, weight) pairs; every function owns a slice of
    the parameter vector, rescaled around that function's default: ``weight * (t - t_default) + t_default``."""

    def __init__(self, transform_list):
        assert type(transform_list) == list
        self.transform_list = [list(e) if isinstance(e, (tuple, list)) else [e, 1.0] for e in transform_list]
        self._t = [np.asarray(fn.t, dtype=np.float32) for fn, _ in self.transform_list]

    def get_param(self, as_tensor=False):
        return torch.Tensor(np.concatenate(self._t)) if as_tensor else self._t

    def get_opt_param(self):
        return np.concatenate([fn.get_opt_param() for fn, _ in self.transform_list])

    def reweight(self, t, weight, t_mean):
        return t_mean + weight * (t - t_mean)

    def __call__(self, ims, t, invert=False, only_spatial=False):
        if t.size(0) == 1:
            t = t.repeat(ims.size(0), 1)
        offset = 0
        for (fn, weight), default in zip(self.transform_list, self._t):
            width = len(fn.t)
            if fn.is_spatial or not only_spatial:
                center = torch.from_numpy(default).type_as(t)
                ims = fn(ims, self.reweight(t[:, offset:offset + width], weight, center), invert=invert)
            offset += width
        return ims

    def __str__(self):
        return "<ComposeTransform\n\t{}\n>".format("\n\t".join(str(fn) for fn, _ in self.transform_list))
