"""StyleGAN2 generator behind pix2latent's model API, executed by the native sm_100a library.

Same surface as /root/reference pix2latent/model/stylegan2.py:66-138: ``StyleGAN2(model='cars' |
'ffhq', search='z')``, ``__call__(z, noises=None)`` -> ``[b,3,R,R]`` clamped to [-1,1],
``.noise_shape``, ``.reshape_noise``, tolerant of ``nn.DataParallel`` wrapping (the examples wrap
it, examples/invert_stylegan2_cars_basincma.py:51; here that wrapper is bypassed — candidates shard
across processes instead, pix2latent_b200/parallel.py).

Per-layer noise: rosinality draws fresh N(0,1) per layer per forward; this wrapper draws it with
torch's CUDA generator in the same layer order and hands the tensors to the library, so a run can
be replayed with explicit ``noises``.

``search='w+'`` (stylegan2.py:97-104, 122-138): ``__call__(z, noises)`` takes the latent in W ([b,512]) or W+
([b,n_latent,512]) and the per-layer noise as ONE flat tensor [b, sum(r*r)] (``reshape_noise`` splits it); both
are differentiable inputs (p2l_sg2_forward_w / p2l_sg2_backward_w); ``latent_mean`` / ``latent_std`` are the
statistics of style(N(0,I)) over 4096 samples as the reference computes them.
"""
import os

import torch
import torch.nn as nn

from .. import native
from . import synth

CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}
IM_DIM = {"cars": 512, "ffhq": 1024}


def pad_channels_to_64(sd, size, channels):
    """Zero-pad every level whose width is not a multiple of 64 (rosinality key layout)."""
    ch = dict(channels)
    log_size = size.bit_length() - 1
    need = {r: ch[r] for r in (2 ** i for i in range(2, log_size + 1)) if ch[r] % 64}
    if not need:
        return sd, ch
    sd = {k: v.clone() for k, v in sd.items()}
    newc = {r: ((c + 63) // 64) * 64 for r, c in need.items()}

    def pad(t, dim, n):
        shape = list(t.shape)
        shape[dim] = n - t.shape[dim]
        return torch.cat([t, torch.zeros(shape, dtype=t.dtype, device=t.device)], dim)

    def fix_styled(pre, cin_res, cout_res):
        if cout_res in newc:
            sd[pre + ".conv.weight"] = pad(sd[pre + ".conv.weight"], 1, newc[cout_res])
            sd[pre + ".activate.bias"] = pad(sd[pre + ".activate.bias"], 0, newc[cout_res])
        if cin_res in newc:
            # the equalised-lr scale is 1/sqrt(fan_in): keep scale*W unchanged under the wider fan-in
            sd[pre + ".conv.weight"] = pad(sd[pre + ".conv.weight"], 2, newc[cin_res]) * (newc[cin_res] / need[cin_res]) ** 0.5
            sd[pre + ".conv.modulation.weight"] = pad(sd[pre + ".conv.modulation.weight"], 0, newc[cin_res])
            sd[pre + ".conv.modulation.bias"] = pad(sd[pre + ".conv.modulation.bias"], 0, newc[cin_res])

    def fix_rgb(pre, res):
        if res in newc:
            sd[pre + ".conv.weight"] = pad(sd[pre + ".conv.weight"], 2, newc[res]) * (newc[res] / need[res]) ** 0.5
            sd[pre + ".conv.modulation.weight"] = pad(sd[pre + ".conv.modulation.weight"], 0, newc[res])
            sd[pre + ".conv.modulation.bias"] = pad(sd[pre + ".conv.modulation.bias"], 0, newc[res])

    if 4 in newc:
        sd["input.input"] = pad(sd["input.input"], 1, newc[4])
    fix_styled("conv1", 4, 4)
    fix_rgb("to_rgb1", 4)
    for i in range(3, log_size + 1):
        r, rp = 2 ** i, 2 ** (i - 1)
        fix_styled("convs.%d" % (2 * (i - 3)), rp, r)
        fix_styled("convs.%d" % (2 * (i - 3) + 1), r, r)
        fix_rgb("to_rgbs.%d" % (i - 3), r)
    ch.update(newc)
    return sd, ch


class _SG2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, model, noises):
        ctx.model, ctx.b = model, z.shape[0]
        return model.native.forward(z, noises)

    @staticmethod
    def backward(ctx, dimg):
        return ctx.model.native.backward(ctx.b, dimg.contiguous()), None, None


class _SG2WFn(torch.autograd.Function):
    """image = G(w | w+, noise) with the mapping network skipped; backward to the latent and the flat noise."""

    @staticmethod
    def forward(ctx, w, noise_flat, model):
        b = w.shape[0]
        noises = None if noise_flat is None else model.reshape_noise(noise_flat)
        ctx.model, ctx.b, ctx.w_dim, ctx.has_noise = model, b, w.dim(), noise_flat is not None
        return native.sg2_forward_w(model.native, w, noises)

    @staticmethod
    def backward(ctx, dimg):
        want_noise = ctx.has_noise and ctx.needs_input_grad[1]
        dlat, dn = native.sg2_backward_w(ctx.model.native, ctx.b, dimg.contiguous(), want_noise_grad=want_noise)
        dw = dlat.sum(1) if ctx.w_dim == 2 else dlat  # a plain w feeds every row
        dnoise = torch.cat([t.reshape(ctx.b, -1) for t in dn], 1) if want_noise else None
        return dw, dnoise, None


CKPT_FILES = {"cars": "stylegan2-car-config-f.pt", "ffhq": "stylegan2-ffhq-config-f.pt"}  # stylegan2.py:53-61


class StyleGAN2(nn.Module):
    """Weights (model/weights.py): ``state_dict`` (rosinality ``g_ema`` keys) if given; else the rosinality checkpoint
    file — ``checkpoint=``, ``$P2L_STYLEGAN2_CKPT`` (a file, or a directory holding the reference's file names), or
    ``stylegan2-pytorch/<file>`` next to this module (where the reference keeps it, stylegan2.py:9,53-61); else — ONLY
    with ``allow_synthetic=True`` / ``P2L_ALLOW_SYNTHETIC=1`` — seeded synthetic weights; otherwise ``MissingWeights``."""

    def __init__(self, model="cars", search="z", state_dict=None, size=None, channels=None, seed=0, checkpoint=None,
                 allow_synthetic=False):
        super().__init__()
        if search not in ("z", "w+"):
            raise ValueError("StyleGAN2(search=%r): expected 'z' or 'w+'" % search)
        self.im_res = int(size or IM_DIM[model])
        self.channels = dict(channels or CHANNELS)
        from . import weights
        env = os.environ.get("P2L_STYLEGAN2_CKPT")
        fname = CKPT_FILES.get(model)
        cands = [checkpoint, env if env and os.path.isfile(env) else None,
                 os.path.join(env, fname) if env and fname and os.path.isdir(env) else None,
                 os.path.join(os.path.dirname(os.path.abspath(__file__)), "stylegan2-pytorch", fname) if fname else None]
        state_dict, self.weights_source = weights.resolve(
            "StyleGAN2(%s)" % model, state_dict, cands, None,
            lambda: synth.stylegan2_state_dict(self.im_res, self.channels, seed), allow_synthetic)
        # the tcgen05 path tiles channels by 64: narrower levels (ffhq-1024's 32-channel top level) are
        # zero-padded to 64 — functionally exact, the padded channels have zero weights on both sides
        self._state, self.channels = pad_channels_to_64(state_dict, self.im_res, self.channels)
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
        self.n_latent = self.native.n_latent
        if self.search == "w+":
            self._latent_statistics()

    @torch.no_grad()
    def _latent_statistics(self, n_mean_latent=4096):
        """stylegan2.py:99-104: mean and (scalar) std of style(z), z ~ N(0, I), 4096 samples."""
        z = torch.randn(n_mean_latent, 512, device="cuda")
        w = torch.cat([native.sg2_style(self.native, z[i:i + 512]) for i in range(0, n_mean_latent, 512)])
        self.latent_mean = w.mean(0)
        self.latent_std = ((w - self.latent_mean).pow(2).sum() / n_mean_latent) ** 0.5

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
        if self.search == "w+":
            return self.forward_w(z, noises)
        if noises is None:
            noises = self.draw_noise(z.shape[0], z.device)
        if torch.is_grad_enabled() and z.requires_grad:
            return _SG2Fn.apply(z, self, noises)
        return self.native.forward(z, noises)

    def forward_w(self, z, noises, truncation=1.0):  # stylegan2.py:122-125
        """z: latent in W [b,512] or W+ [b,n_latent,512]; noises: flat [b, sum(r*r)] (``reshape_noise`` layout).
        ``noises=None`` draws fresh per-layer noise (the reference requires the tensor)."""
        if noises is None:
            noises = torch.cat([n.reshape(z.shape[0], -1) for n in self.draw_noise(z.shape[0], z.device)], 1)
        if torch.is_grad_enabled() and (z.requires_grad or noises.requires_grad):
            return _SG2WFn.apply(z, noises, self)
        return native.sg2_forward_w(self.native, z, self.reshape_noise(noises))

    def reshape_noise(self, z):  # stylegan2.py:128-138
        st, out = 0, []
        for d in self.noise_shape:
            en = st + d[-2] * d[-1]
            out.append(z[:, st:en].reshape(-1, 1, d[-2], d[-1]))
            st = en
        assert z.size(1) == en
        return out
