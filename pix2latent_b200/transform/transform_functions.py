"""Colour transformations of the target (reference: pix2latent/transform/transform_functions.py:12-120:
Hue / Brightness / Gamma / Saturation / Contrast, each a one-parameter torchvision colour op with a clamp
range and an inverse-parameter rule). As in the reference they are NOT differentiable and run through
PIL on the host: they are applied once per CMA meta-iteration to a handful of images, off the hot path.
The result returns to the device of the input (the reference hard-codes ``.cuda()``)."""
import numpy as np
import torch

# name -> (torchvision.transforms.functional op, default t, (t_min, t_max), inverse-parameter rule)
_RECIPROCAL = "reciprocal"
_NEGATE = "negate"
_COLOR_OPS = {
    "Hue": ("adjust_hue", 0.0, (-0.5, 0.5), _NEGATE),
    "Brightness": ("adjust_brightness", 1.0, (0.667, 1.5), _RECIPROCAL),
    "Gamma": ("adjust_gamma", 1.0, (0.667, 1.5), _RECIPROCAL),
    "Saturation": ("adjust_saturation", 1.0, (0.667, 1.5), _RECIPROCAL),
    "Contrast": ("adjust_contrast", 1.0, (0.667, 1.5), _RECIPROCAL),
}


def _inverse_param(rule, t):
    return -t if rule == _NEGATE else 1.0 / t


class ColorTransform(object):
    """``fn(pil_image, t)`` applied image by image with ``t`` clamped to ``t_range``."""

    is_spatial = False

    def __init__(self, fn, t=(1,), t_range=(0.667, 1.5), t_inv_fn=None, optimize=True):
        lo, hi = t_range
        assert hi > lo, "t_range should be increasing"
        self.fn, self.t_inv_fn = fn, t_inv_fn
        self.t = np.asarray(t, dtype=np.float32)
        self.t_min, self.t_max = lo, hi
        self.optimize = optimize

    def get_opt_param(self):
        return self.t if self.optimize else []

    def apply(self, ims, t, invert=False):
        import torchvision.transforms.functional as TVF
        assert ims.size(0) == t.size(0) and t.size(1) == 1
        if invert:
            t = self.t_inv_fn(t)
        t = t.clamp(self.t_min, self.t_max)
        unit = (ims.detach().cpu() + 1.0) / 2.0  # [-1, 1] -> [0, 1] for PIL
        done = [TVF.to_tensor(self.fn(TVF.to_pil_image(im), float(ti))) for im, ti in zip(unit, t)]
        return (2.0 * (torch.stack(done) - 0.5)).float().to(ims.device)

    __call__ = apply

    def __str__(self):
        return "ColorTransform: {}".format(self.fn)


def _make(name):
    op, default, (lo, hi), rule = _COLOR_OPS[name]
    eps = 1e-6 if name == "Hue" else 0.0  # adjust_hue rejects the end points

    def __init__(self, t=None, t_min=lo, t_max=hi):
        import torchvision.transforms.functional as TVF
        ColorTransform.__init__(self, fn=getattr(TVF, op), t=[default] if t is None else t,
                                t_range=(t_min + eps, t_max - eps), t_inv_fn=lambda x: _inverse_param(rule, x))

    return type(name + "Transform", (ColorTransform,), {"__init__": __init__, "__doc__": "torchvision %s" % op})


HueTransform = _make("Hue")
BrightnessTransform = _make("Brightness")
GammaTransform = _make("Gamma")
SaturationTransform = _make("Saturation")
ContrastTransform = _make("Contrast")
