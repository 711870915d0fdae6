"""Partial pin of the third-party StyleGAN2 arithmetic: the FIR resampling op (rosinality op/upfirdn2d.py, restated in
oracle/stylegan2.py::upfirdn2d and re-implemented in CUDA by k_sg_post_fwd / k_sg_torgb_fwd) against an implementation
this repo did not write — scipy.signal.upfirdn (polyphase up-FIR-down) applied along both axes — for the two uses the
generator makes of it: `Upsample([1,3,3,1], factor 2)` of the RGB skip and `Blur([1,3,3,1], pad=(1,1))` after a stride-2
transposed convolution. Plus the separable-kernel identity the CUDA kernels rely on (the 4x4 FIR is k (x) k)."""
import numpy as np
import torch


def _scipy_upfirdn2d(x, k1, up, pad0, pad1):
    """Full up-FIR along both axes with scipy, then the crop rosinality's (pad0, pad1) zero padding corresponds to."""
    from scipy.signal import upfirdn
    kw = len(k1)
    h, w = x.shape[-2:]

    def axis(a, ax, n):
        full = upfirdn(k1, a, up=up, axis=ax)                 # length (n-1)*up + kw
        if full.shape[ax] < n * up + kw - 1:                   # the zero-inserted signal's trailing zeros
            padw = [(0, 0)] * a.ndim
            padw[ax] = (0, n * up + kw - 1 - full.shape[ax])
            full = np.pad(full, padw)
        start, length = kw - 1 - pad0, n * up + pad0 + pad1 - kw + 1
        return np.take(full, np.arange(start, start + length), axis=ax)

    return axis(axis(x, -2, h), -1, w)


def test_upsample_skip_matches_scipy():
    from oracle import stylegan2 as osg
    torch.manual_seed(0)
    x = torch.randn(2, 3, 9, 7, dtype=torch.float64)
    up = osg.Upsample([1, 3, 3, 1], factor=2)
    y = up(x).numpy()
    k1 = np.array([1, 3, 3, 1], dtype=np.float64)
    k1 = k1 / k1.sum() * 2                                       # (k (x) k / sum) * factor^2 = (2 k/sum) (x) (2 k/sum)
    ref = _scipy_upfirdn2d(x.numpy(), k1, 2, up.pad[0], up.pad[1])
    assert y.shape == (2, 3, 18, 14) == ref.shape
    np.testing.assert_allclose(y, ref, rtol=1e-12, atol=1e-12)
    # a constant image stays constant under the up-sampler (unit DC gain), away from the zero-padded border
    c = up(torch.ones(1, 1, 8, 8, dtype=torch.float64))[0, 0, 2:-2, 2:-2]
    assert torch.allclose(c, torch.ones_like(c))


def test_blur_after_transposed_conv_matches_scipy():
    from oracle import stylegan2 as osg
    torch.manual_seed(1)
    x = torch.randn(1, 4, 17, 17, dtype=torch.float64)           # (2H+1)^2 grid a stride-2 conv_transpose produces
    blur = osg.Blur([1, 3, 3, 1], pad=(1, 1), upsample_factor=2)
    y = blur(x).numpy()
    k1 = np.array([1, 3, 3, 1], dtype=np.float64)
    k1 = k1 / k1.sum() * 2
    ref = _scipy_upfirdn2d(x.numpy(), k1, 1, 1, 1)
    assert y.shape == (1, 4, 16, 16) == ref.shape
    np.testing.assert_allclose(y, ref, rtol=1e-12, atol=1e-12)


def test_fir_kernel_is_separable():
    from oracle import stylegan2 as osg
    k2 = osg.make_kernel([1, 3, 3, 1]).double() * 4
    k1 = torch.tensor([0.25, 0.75, 0.75, 0.25], dtype=torch.float64)   # kFir in sg2_kernels.cu
    assert torch.allclose(k2, k1[:, None] * k1[None, :])
