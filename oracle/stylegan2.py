"""ORACLE (test infrastructure) — CPU/torch restatement of the StyleGAN2 generator that
pix2latent's wrapper runs (reference: pix2latent/model/stylegan2.py:66-138: ``Generator(im_res,
512, 8, channel_multiplier=2)``, ``forward_z`` :116-119 = ``model([z], truncation=1.0, ...)[0]
.clamp_(-1, 1)``).

The arithmetic lives in ``rosinality/stylegan2-pytorch`` (``model.py``, ``op/fused_act.py``,
``op/upfirdn2d.py``), which pix2latent git-clones at run time from an UNPINNED HEAD
(stylegan2.py:24-25) and which is not present in this image; this file restates the published
code (pure-torch branches of the two custom ops) — SURVEY.md Appendix A.3. PARITY UNPINNED (no
vectors exist). Module and parameter names follow the rosinality ``g_ema`` state dict so official
checkpoints load.

Noise: rosinality draws fresh N(0,1) per layer per call (NoiseInjection with noise=None); for a
replayable oracle ``forward`` takes the per-layer noise tensors explicitly; with
``noise=None`` it draws them in the reference's order with ``torch.randn``.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=2 ** 0.5):  # op/fused_act.py (CPU branch)
    rest = [1] * (x.ndim - bias.ndim - 1)
    return F.leaky_relu(x + bias.view(1, bias.shape[0], *rest), negative_slope=negative_slope) * scale


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):  # op/upfirdn2d.py upfirdn2d_native
    pad0, pad1 = pad
    _, channel, in_h, in_w = x.shape
    x = x.reshape(-1, in_h, in_w, 1)
    _, in_h, in_w, minor = x.shape
    kh, kw = kernel.shape
    out = x.view(-1, in_h, 1, in_w, 1, minor)
    out = F.pad(out, [0, 0, 0, up - 1, 0, 0, 0, up - 1])
    out = out.view(-1, in_h * up, in_w * up, minor)
    out = F.pad(out, [0, 0, max(pad0, 0), max(pad1, 0), max(pad0, 0), max(pad1, 0)])
    out = out[:, max(-pad0, 0): out.shape[1] - max(-pad1, 0), max(-pad0, 0): out.shape[2] - max(-pad1, 0), :]
    out = out.permute(0, 3, 1, 2)
    out = out.reshape([-1, 1, in_h * up + pad0 + pad1, in_w * up + pad0 + pad1])
    w = torch.flip(kernel, [0, 1]).view(1, 1, kh, kw)
    out = F.conv2d(out, w)
    out = out.reshape(-1, minor, in_h * up + pad0 + pad1 - kh + 1, in_w * up + pad0 + pad1 - kw + 1)
    out = out.permute(0, 2, 3, 1)
    out = out[:, ::down, ::down, :]
    out_h = (in_h * up + pad0 + pad1 - kh + down) // down
    out_w = (in_w * up + pad0 + pad1 - kw + down) // down
    return out.view(-1, channel, out_h, out_w)


def make_kernel(k):
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


class PixelNorm(nn.Module):
    def forward(self, x):
        return x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


class Upsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel) * (factor ** 2))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, x):
        return upfirdn2d(x, self.kernel.to(x.dtype), up=self.factor, down=1, pad=self.pad)


class Blur(nn.Module):
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, x):
        return upfirdn2d(x, self.kernel.to(x.dtype), pad=self.pad)


class EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, x):
        if self.activation:
            return fused_leaky_relu(F.linear(x, self.weight * self.scale), self.bias * self.lr_mul)
        return F.linear(x, self.weight * self.scale, bias=self.bias * self.lr_mul)


class ModulatedConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size, self.in_channel, self.out_channel, self.upsample = kernel_size, in_channel, out_channel, upsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def forward(self, x, style):
        batch, in_channel, height, width = x.shape
        style = self.modulation(style).view(batch, 1, in_channel, 1, 1)
        weight = self.scale * self.weight * style
        if self.demodulate:
            demod = torch.rsqrt(weight.pow(2).sum([2, 3, 4]) + 1e-8)
            weight = weight * demod.view(batch, self.out_channel, 1, 1, 1)
        weight = weight.view(batch * self.out_channel, in_channel, self.kernel_size, self.kernel_size)
        if self.upsample:
            x = x.view(1, batch * in_channel, height, width)
            weight = weight.view(batch, self.out_channel, in_channel, self.kernel_size, self.kernel_size)
            weight = weight.transpose(1, 2).reshape(batch * in_channel, self.out_channel, self.kernel_size, self.kernel_size)
            out = F.conv_transpose2d(x, weight, padding=0, stride=2, groups=batch)
            _, _, height, width = out.shape
            out = out.view(batch, self.out_channel, height, width)
            return self.blur(out)
        x = x.view(1, batch * in_channel, height, width)
        out = F.conv2d(x, weight, padding=self.padding, groups=batch)
        _, _, height, width = out.shape
        return out.view(batch, self.out_channel, height, width)


class NoiseInjection(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, x):
        return self.input.repeat(x.shape[0], 1, 1, 1)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope, self.scale = negative_slope, scale

    def forward(self, x):
        return fused_leaky_relu(x, self.bias, self.negative_slope, self.scale)


class StyledConv(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample, blur_kernel=blur_kernel)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, x, style, noise=None):
        return self.activate(self.noise(self.conv(x, style), noise=noise))


class ToRGB(nn.Module):
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, x, style, skip=None):
        out = self.conv(x, style) + self.bias
        if skip is not None:
            out = out + self.upsample(skip)
        return out


DEFAULT_CHANNELS = lambda cm: {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm,
                               512: 32 * cm, 1024: 16 * cm}


class Generator(nn.Module):
    """rosinality model.py Generator. ``channels`` may be overridden for reduced test configs."""

    def __init__(self, size, style_dim=512, n_mlp=8, channel_multiplier=2, blur_kernel=(1, 3, 3, 1), lr_mlp=0.01,
                 channels=None):
        super().__init__()
        self.size, self.style_dim = size, style_dim
        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        self.channels = channels or DEFAULT_CHANNELS(channel_multiplier)
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs, self.to_rgbs, self.noises = nn.ModuleList(), nn.ModuleList(), nn.Module()
        in_channel = self.channels[4]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer("noise_%d" % layer_idx, torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2

    def noise_shapes(self, batch):
        return [(batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i in range(self.num_layers)]

    def forward(self, styles, input_is_latent=False, noise=None):
        """truncation = 1 (the only value pix2latent passes); ``noise``: list of num_layers tensors or None."""
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if noise is None:
            noise = [None] * self.num_layers
        latent = styles[0]
        if latent.ndim < 3:
            latent = latent.unsqueeze(1).repeat(1, self.n_latent, 1)
        out = self.input(latent)
        out = self.conv1(out, latent[:, 0], noise=noise[0])
        skip = self.to_rgb1(out, latent[:, 1])
        i = 1
        for conv1, conv2, noise1, noise2, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2], noise[2::2], self.to_rgbs):
            out = conv1(out, latent[:, i], noise=noise1)
            out = conv2(out, latent[:, i + 1], noise=noise2)
            skip = to_rgb(out, latent[:, i + 2], skip)
            i += 2
        return skip, None


class StyleGAN2Oracle(nn.Module):
    """pix2latent/model/stylegan2.py:66-138 (search='z') with explicit noise and a device argument."""

    def __init__(self, size=512, channels=None):
        super().__init__()
        self.im_res = size
        self.model = Generator(size, 512, 8, channel_multiplier=2, channels=channels)
        self.noise_shape = [list(getattr(self.model.noises, "noise_%d" % i).size()) for i in range(self.model.num_layers)]
        self.search = "z"
        self.eval()

    def forward(self, z, noises=None):  # stylegan2.py:110-119
        out = self.model([z], noise=noises)[0]
        return out.clamp(-1.0, 1.0)


TINY_CHANNELS = {4: 128, 8: 128, 16: 64, 32: 64}


@torch.no_grad()
def init_random_(m: StyleGAN2Oracle, seed=0):
    """Seeded synthetic weights in the regime of a trained network: N(0,1) weights under the
    equalised-lr scaling (as rosinality initialises them), non-zero noise strengths and biases."""
    g = torch.Generator().manual_seed(2000 + seed)
    for name, p in m.model.named_parameters():
        if name.endswith("noise.weight"):
            p.copy_(0.1 * torch.randn(p.shape, generator=g))
        elif name.endswith("modulation.bias"):
            p.fill_(1.0)
        elif name.endswith("modulation.weight"):
            p.copy_(torch.randn(p.shape, generator=g))
        elif name.startswith("style.") and name.endswith(".weight"):
            p.copy_(torch.randn(p.shape, generator=g) / 0.01)
        elif name.startswith("style.") and name.endswith(".bias"):
            p.copy_(torch.randn(p.shape, generator=g))
        elif name.endswith("activate.bias") or name.endswith("to_rgb1.bias") or ".bias" in name:
            p.copy_(0.1 * torch.randn(p.shape, generator=g))
        elif "to_rgb" in name and name.endswith("conv.weight"):
            p.copy_(0.25 * torch.randn(p.shape, generator=g))  # keeps the summed skip image in range
        else:
            p.copy_(torch.randn(p.shape, generator=g))
    return m


def make_stylegan2(size=512, channels=None, seed=0, dtype=torch.float32):
    st = torch.random.get_rng_state()
    m = StyleGAN2Oracle(size, channels)
    torch.random.set_rng_state(st)
    init_random_(m, seed)
    return m.to(dtype).eval()
