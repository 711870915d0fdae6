"""Scale + translation of the target / weight images (reference:
pix2latent/transform/spatial_transform.py:11-108). The resampling — ``F.grid_sample(ims,
F.affine_grid(theta, ims.size()))`` there — runs in the native library (p2l_affine_resample); there is
no CPU path."""
import numpy as np
import torch

from .. import native
from .base_transform import TransformTemplate
from .transform_utils import compute_pre_alignment


def _theta(t, invert):
    """[b,3] = (s, tx, ty) -> affine matrices [b,2,3] (spatial_transform.py:79-83, 99-103)."""
    theta = torch.zeros(t.size(0), 2, 3, dtype=t.dtype, device=t.device)
    if not invert:
        theta[:, 0, 0] = t[:, 0]
        theta[:, 1, 1] = t[:, 0]
        theta[:, :, 2] = t[:, 1:]
    else:
        theta[:, 0, 0] = 1.0 / t[:, 0]
        theta[:, 1, 1] = 1.0 / t[:, 0]
        theta[:, :, 2] = -(t[:, 1:] / t[:, :1])
    return theta


class SpatialTransform(TransformTemplate):
    """Transformation parameter ``[s, t_x, t_y]``; the search variable is ``delta_t`` with
    ``t = default_t + sensitivity * delta_t``."""

    def __init__(self, t=[1., 0., 0.], identity_t=[1., 0., 0.], pre_align=None, sensitivity=0.1):
        """
        Args:
            identity_t (list): identity parameter, centre of the search
            pre_align (image): if not None, a binary mask image used to compute the initial alignment
            sensitivity (float): scale of delta_t
        """
        self.identity_t = np.array(identity_t, dtype=np.float32)
        self.is_spatial = True
        self.sensitivity = sensitivity
        self.t = t
        if pre_align is not None:
            self.t = compute_pre_alignment(pre_align)
        self._t = torch.Tensor(self.t)

    def __call__(self, ims, delta_t, invert=False):
        t = self._t.type_as(ims) + (self.sensitivity * delta_t.type_as(ims))
        if invert:
            return self.invert_transform(ims, t)
        return self.transform(ims, t)

    def get_default_param(self, as_tensor=True):
        if as_tensor:
            return self._t
        return self.t

    def get_identity_param(self, as_tensor=True):
        # (the reference's version reads an undefined `as_tensor`, spatial_transform.py:62-65)
        if as_tensor:
            return torch.Tensor(self.identity_t)
        return self.identity_t

    @staticmethod
    def _resample(ims, theta):
        if not ims.is_cuda:
            raise RuntimeError("SpatialTransform resamples on a CUDA (sm_100a) device only; got a CPU tensor")
        return native.affine_resample(ims, theta.to(ims.device))

    def transform(self, ims, t):
        """ims [b,c,h,w], t [b,3] -> transformed images."""
        return self._resample(ims, _theta(t, invert=False))

    def invert_transform(self, ims, t):
        """Inverse of ``transform`` with the same t (up to resampling loss)."""
        return self._resample(ims, _theta(t, invert=True))

    def __str__(self):
        return "SpatialTransform: {}".format(super().__str__())
