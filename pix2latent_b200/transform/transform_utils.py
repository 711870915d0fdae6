"""Pre-alignment statistics and transform composition (reference:
pix2latent/transform/transform_utils.py:53-184). Host arithmetic on a handful of numbers."""
import numpy as np
import torch

from ..utils.image import binarize


def compute_pre_alignment(weight):
    """Initial [s, tx, ty] that moves the masked object onto BigGAN's typical object box."""
    dst_center, dst_size = get_biggan_stats()
    src_center, src_size = compute_stat_from_mask(binarize(weight))
    t = convert_to_t(src_center, src_size, dst_center, dst_size)
    return t.numpy()


def convert_to_t(src_center, src_size, dst_center, dst_size):
    """Transformation parameter taking an object of (src_center, src_size) to (dst_center, dst_size);
    the scale follows the object's larger side, the shift is in affine_grid's [-1, 1] units (x, y)."""
    src_center, src_size = np.array(src_center), np.array(src_size)
    dst_center, dst_size = np.array(dst_center), np.array(dst_size)
    scale_idx = np.argmax(src_size).squeeze()
    s = (src_size / dst_size)[scale_idx]
    dxy = (src_center - dst_center) * 2.
    t = np.array([s, *dxy[::-1]])
    return torch.from_numpy(t).float()


def get_biggan_stats():
    """precomputed BigGAN object statistics: (centre of mass, object size), fractions of the image"""
    center_of_mass = [137 / 255., 127 / 255.]
    object_size = [213 / 255., 210 / 255.]
    return center_of_mass, object_size


def compute_stat_from_mask(mask):
    """Binary mask [c,h,w] -> ((cy, cx), (h, w)) of its bounding box as fractions of the image."""
    st_h, st_w, en_h, en_w = bbox_from_mask(mask)
    obj_size = obj_h, obj_w = en_h - st_h, en_w - st_w
    obj_center = (st_h + obj_h // 2, st_w + obj_w // 2)
    obj_size = (obj_size[0] / mask.size(1), obj_size[1] / mask.size(2))
    obj_center = (obj_center[0] / mask.size(1), obj_center[1] / mask.size(2))
    return obj_center, obj_size


def bbox_from_mask(mask):
    assert len(list(mask.size())) == 3, "expected 3d tensor but got {}".format(len(list(mask.size())))
    rows = (mask.mean(0).sum(1) != 0).nonzero()
    cols = (mask.mean(0).sum(0) != 0).nonzero()
    if rows.numel() > 0:
        tlc_h, brc_h = rows[0].item(), rows[-1].item()
    else:
        tlc_h, brc_h = 0, mask.size(1)  # whole range when the mask is empty
    if cols.numel() > 0:
        tlc_w, brc_w = cols[0].item(), cols[-1].item()
    else:
        tlc_w, brc_w = 0, mask.size(2)
    return tlc_h, tlc_w, brc_h, brc_w


class ComposeTransform():
    """Chain of transform functions, each optionally with a weight that rescales its slice of the parameter
    vector around that function's default (``weight * (t - t_default) + t_default``)."""

    def __init__(self, transform_list):
        assert type(transform_list) == list
        self.transform_list = []
        for t_fn in transform_list:
            if type(t_fn) in [tuple, list]:
                self.transform_list.append(t_fn)
            else:
                self.transform_list.append([t_fn, 1.0])
        self._t = [np.asarray(x[0].t, dtype=np.float32) for x in self.transform_list]

    def get_param(self, as_tensor=False):
        if as_tensor:
            return torch.Tensor(np.concatenate(self._t))
        return self._t

    def get_opt_param(self):
        return np.concatenate([x[0].get_opt_param() for x in self.transform_list])

    def reweight(self, t, weight, t_mean):
        return (weight * (t - t_mean)) + t_mean

    def __call__(self, ims, t, invert=False, only_spatial=False):
        if t.size(0) == 1:
            t = t.repeat(ims.size(0), 1)
        t_i = 0
        for i, (fn, fn_weight) in enumerate(self.transform_list):
            t_sz = len(fn.t)
            if (only_spatial and fn.is_spatial) or not only_spatial:
                t_param = t[:, t_i:t_i + t_sz]
                t_mu = torch.from_numpy(self._t[i]).type_as(t_param)
                t_param = self.reweight(t_param, fn_weight, t_mu)
                ims = fn(ims, t_param, invert=invert)
            t_i += t_sz
        return ims

    def __str__(self):
        return "<ComposeTransform\n\t{}\n>".format("\n\t".join([f[0].__str__() for f in self.transform_list]))
