"""ORACLE (test infrastructure) — CPU restatement of the spatial target transformation.

Follows /root/reference pix2latent/transform/spatial_transform.py:43-104 (``SpatialTransform.__call__``,
``transform``, ``invert_transform``) with ``F.affine_grid`` / ``F.grid_sample`` (torch defaults: bilinear,
zero padding, align_corners=False) written out as explicit index arithmetic — the same formulas the CUDA
kernel (pix2latent_b200/csrc/kernels.cu affine_resample_kernel) evaluates.

PINNED: tests/golden/make_golden_transform.py runs the REAL reference SpatialTransform on CPU;
tests/test_transform_cpu.py checks this restatement against those vectors (tests/golden/reference_transform_cpu.npz).
"""
import numpy as np
import torch


def affine_resample(src, theta):
    """src [b|1,C,H,W], theta [b,2,3] (torch tensors) -> [b,C,H,W] float32."""
    s = src.detach().cpu().numpy().astype(np.float32)
    th = theta.detach().cpu().numpy().astype(np.float32)
    b = th.shape[0]
    _, C, H, W = s.shape
    out = np.zeros((b, C, H, W), dtype=np.float32)
    x = ((2.0 * np.arange(W, dtype=np.float32) + 1.0) / W - 1.0)[None, :]  # affine_grid, align_corners=False
    y = ((2.0 * np.arange(H, dtype=np.float32) + 1.0) / H - 1.0)[:, None]
    for n in range(b):
        im = s[0] if s.shape[0] == 1 else s[n]
        gx = th[n, 0, 0] * x + th[n, 0, 1] * y + th[n, 0, 2]
        gy = th[n, 1, 0] * x + th[n, 1, 1] * y + th[n, 1, 2]
        ix = ((gx + 1.0) * W - 1.0) * 0.5  # grid_sample un-normalisation, align_corners=False
        iy = ((gy + 1.0) * H - 1.0) * 0.5
        x0 = np.floor(ix).astype(np.int64)
        y0 = np.floor(iy).astype(np.int64)
        ax, ay = ix - x0, iy - y0
        for dy, dx, wgt in ((0, 0, (1 - ax) * (1 - ay)), (0, 1, ax * (1 - ay)), (1, 0, (1 - ax) * ay), (1, 1, ax * ay)):
            xx, yy = x0 + dx, y0 + dy
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)  # zero padding
            v = im[:, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
            out[n] += np.where(ok, wgt, 0.0).astype(np.float32)[None] * v
    return torch.from_numpy(out)


def theta_of(t, invert=False):
    """(s, tx, ty) rows -> [b,2,3] (spatial_transform.py:79-83 forward, :99-103 inverse)."""
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


class SpatialTransform:
    """spatial_transform.py:11-66 with the resampling above (CPU)."""

    is_spatial = True

    def __init__(self, t=(1., 0., 0.), sensitivity=0.1):
        self.t = list(t)
        self._t = torch.Tensor(self.t)
        self.sensitivity = sensitivity

    def __call__(self, ims, delta_t, invert=False):
        t = self._t.type_as(ims) + self.sensitivity * delta_t
        return affine_resample(ims, theta_of(t, invert)).type_as(ims)

    def get_default_param(self, as_tensor=True):
        return self._t if as_tensor else self.t


class TorchSpatialTransform(SpatialTransform):
    """Same interface with torch's own ``F.affine_grid`` + ``F.grid_sample`` — bit-identical to what the
    reference executes, so host-logic comparisons (tests/test_transform_cpu.py (c)) carry no resampling noise
    into the chaotic Adam trajectories."""

    def __call__(self, ims, delta_t, invert=False):
        import torch.nn.functional as F
        t = self._t.type_as(ims) + self.sensitivity * delta_t
        theta = theta_of(t, invert).type_as(ims)
        return F.grid_sample(ims, F.affine_grid(theta, ims.size(), align_corners=False), align_corners=False)
