"""StyleGAN2 generator behind pix2latent's model API, executed by the native sm_100a library.

Same surface as /root/reference pix2latent/model/stylegan2.py:66-138: ``StyleGAN2(model='cars' |
'ffhq', search='z')``, ``__call__(z, noises=None)`` -> ``[b,3,R,R]`` clamped to [-1,1],
``.noise_shape``, ``.reshape_noise``, tolerant of ``nn.DataParallel`` wrapping (the examples wrap
it, examples/invert_stylegan2_cars_basincma.py:51; here that wrapper is bypassed — candidates shard
across processes instead, pix2latent_b200/parallel.py).

Per-layer noise: rosinality draws fresh N(0,1) per layer per forward; this wrapper draws it with
torch's CUDA generator in the same layer order and hands the tensors to the library, so a run can
be replayed with explicit ``noises``. ``search='w+'`` (forward_w) is not built yet
(SURVEY.md §8f N3).
"""
import warnings

import torch
import torch.nn as nn

from .. import native
from . import synth

CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}
IM_DIM = {"cars": 512, "ffhq": 1024}


class _SG2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, model, noises):
        ctx.model, ctx.b = model, z.shape[0]
        return model.native.forward(z, noises)

    @staticmethod
    def backward(ctx, dimg):
        return ctx.model.native.backward(ctx.b, dimg.contiguous()), None, None


class StyleGAN2(nn.Module):
    def __init__(self, model="cars", search="z", state_dict=None, size=None, channels=None, seed=0):
        super().__init__()
        if search != "z":
            raise NotImplementedError("StyleGAN2(search=%r): only the z search of the reference examples is built" % search)
        self.im_res = int(size or IM_DIM[model])
        self.channels = dict(channels or CHANNELS)
        if any(self.channels[2 ** i] % 64 for i in range(2, self.im_res.bit_length())):
            raise NotImplementedError("feature widths must be multiples of 64 (ffhq-1024's 32-channel top level is not built yet)")
        if state_dict is None:
            warnings.warn("StyleGAN2: no checkpoint reachable offline; using seeded random-init weights (seed=%d)" % seed)
            state_dict = synth.stylegan2_state_dict(self.im_res, self.channels, seed)
        self._state = state_dict
        self.search = search
        self.native = None
        log_size = self.im_res.bit_length() - 1
        self.num_layers = (log_size - 2) * 2 + 1
        self.noise_shape = [[1, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)] for i in range(self.num_layers)]
        if torch.cuda.is_available():
            self._build()

    def _build(self):
        sd = {k: v.cuda() for k, v in self._state.items()}
        self.native = native.NativeStyleGAN2(self.im_res, self.channels, sd)

    def cuda(self, device=None):
        if self.native is None:
            if not torch.cuda.is_available():
                raise RuntimeError("StyleGAN2.cuda(): no CUDA device; pix2latent_b200 has no CPU path")
            self._build()
        return self

    def eval(self):
        return self

    def train(self, mode=True):
        return self

    def draw_noise(self, b, device):
        """Fresh per-layer noise in the reference's order (one normal_() per NoiseInjection)."""
        return [torch.randn(b, 1, s[2], s[3], device=device) for s in self.noise_shape]

    def forward(self, z, noises=None, truncation=1.0):
        if self.native is None:
            raise RuntimeError("StyleGAN2: native sm_100a model not built (no CUDA device). There is no CPU fallback.")
        if noises is None:
            noises = self.draw_noise(z.shape[0], z.device)
        if torch.is_grad_enabled() and z.requires_grad:
            return _SG2Fn.apply(z, self, noises)
        return self.native.forward(z, noises)

    def reshape_noise(self, z):  # stylegan2.py:128-138
        st, out = 0, []
        for d in self.noise_shape:
            en = st + d[-2] * d[-1]
            out.append(z[:, st:en].reshape(-1, 1, d[-2], d[-1]))
            st = en
        assert z.size(1) == en
        return out
