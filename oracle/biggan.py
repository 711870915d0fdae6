"""ORACLE (test infrastructure, not product code) — CPU/torch restatement of the generator that
pix2latent's BigGAN wrapper runs (reference: pix2latent/model/biggan.py:23-58).

The arithmetic itself lives in the third-party package ``pytorch_pretrained_biggan>=0.1.1``
(requirements.txt:9), which is NOT present under /root/reference nor installable here, so this
file restates that package's published ``model.py`` (Generator / GenBlock / BigGANBatchNorm /
SelfAttn) and ``config.py`` (biggan-deep-256) — see SURVEY.md Appendix A.1.  PARITY UNPINNED for
the third-party arithmetic: the reference ships no golden vectors; the in-tree
call sites that ARE pinned against the real reference code are listed in
tests/golden/make_golden.py.

Module / parameter names follow the HF package after pix2latent strips spectral norm
(pix2latent/utils/misc.py:150-157: ``weight_orig`` -> ``weight``), so an official checkpoint's
state dict loads into ``BigGANOracle`` unchanged.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline legs may import this.
"""
import math
from dataclasses import dataclass, field
from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class BigGANConfig:
    """pytorch_pretrained_biggan/config.py (biggan-deep-256 values as defaults)."""
    output_dim: int = 256
    z_dim: int = 128
    class_embed_dim: int = 128
    channel_width: int = 128
    num_classes: int = 1000
    layers: List[Tuple[bool, int, int]] = field(default_factory=lambda: [
        (False, 16, 16), (True, 16, 16), (False, 16, 16), (True, 16, 8), (False, 8, 8), (True, 8, 8),
        (False, 8, 8), (True, 8, 4), (False, 4, 4), (True, 4, 2), (False, 2, 2), (True, 2, 1)])
    attention_layer_position: int = 8
    eps: float = 1e-4
    n_stats: int = 51

    @staticmethod
    def deep256():
        return BigGANConfig()

    @staticmethod
    def tiny128():
        """Reduced config for fast parity tests: every layer type of biggan-deep-256 (up / same
        blocks, channel-dropping skips, attention, rgb head), 128x128 output, attention at 32x32.
        Channel counts stay multiples of 64 (what the tcgen05 path tiles on)."""
        return BigGANConfig(output_dim=128, num_classes=16,
                            layers=[(True, 4, 4), (True, 4, 4), (True, 4, 4), (False, 4, 4),
                                    (True, 4, 2), (True, 2, 1)],
                            attention_layer_position=3)

    @property
    def first_channels(self):
        return self.layers[0][1] * self.channel_width


class BigGANBatchNorm(nn.Module):
    """HF model.py BigGANBatchNorm: eval-only BN with 51 truncation-indexed stat rows and (when
    conditional) gain/offset predicted from the 256-d condition vector."""

    def __init__(self, num_features, condition_vector_dim=None, n_stats=51, eps=1e-4, conditional=True):
        super().__init__()
        self.num_features, self.eps, self.conditional = num_features, eps, conditional
        self.register_buffer("running_means", torch.zeros(n_stats, num_features))
        self.register_buffer("running_vars", torch.ones(n_stats, num_features))
        self.step_size = 1.0 / (n_stats - 1)
        if conditional:
            self.scale = nn.Linear(condition_vector_dim, num_features, bias=False)
            self.offset = nn.Linear(condition_vector_dim, num_features, bias=False)
        else:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))

    def stats(self, truncation):
        coef, start_idx = math.modf(truncation / self.step_size)
        start_idx = int(start_idx)
        if coef != 0.0:
            mean = self.running_means[start_idx] * coef + self.running_means[start_idx + 1] * (1 - coef)
            var = self.running_vars[start_idx] * coef + self.running_vars[start_idx + 1] * (1 - coef)
        else:
            mean, var = self.running_means[start_idx], self.running_vars[start_idx]
        return mean, var

    def forward(self, x, truncation, condition_vector=None):
        mean, var = self.stats(truncation)
        if self.conditional:
            mean = mean[None, :, None, None]
            var = var[None, :, None, None]
            weight = 1 + self.scale(condition_vector)[:, :, None, None]
            bias = self.offset(condition_vector)[:, :, None, None]
            return (x - mean) / torch.sqrt(var + self.eps) * weight + bias
        return F.batch_norm(x, mean, var, self.weight, self.bias, training=False, momentum=0.0, eps=self.eps)


class GenBlock(nn.Module):
    """HF model.py GenBlock (bottleneck residual block, reduction 4)."""

    def __init__(self, in_size, out_size, condition_vector_dim, reduction_factor=4, up_sample=False,
                 n_stats=51, eps=1e-4):
        super().__init__()
        self.up_sample = up_sample
        self.drop_channels = in_size != out_size
        mid = in_size // reduction_factor
        self.bn_0 = BigGANBatchNorm(in_size, condition_vector_dim, n_stats, eps, True)
        self.conv_0 = nn.Conv2d(in_size, mid, 1)
        self.bn_1 = BigGANBatchNorm(mid, condition_vector_dim, n_stats, eps, True)
        self.conv_1 = nn.Conv2d(mid, mid, 3, padding=1)
        self.bn_2 = BigGANBatchNorm(mid, condition_vector_dim, n_stats, eps, True)
        self.conv_2 = nn.Conv2d(mid, mid, 3, padding=1)
        self.bn_3 = BigGANBatchNorm(mid, condition_vector_dim, n_stats, eps, True)
        self.conv_3 = nn.Conv2d(mid, out_size, 1)
        self.relu = nn.ReLU()

    def forward(self, x, cond_vector, truncation):
        x0 = x
        x = self.conv_0(self.relu(self.bn_0(x, truncation, cond_vector)))
        x = self.relu(self.bn_1(x, truncation, cond_vector))
        if self.up_sample:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = self.conv_1(x)
        x = self.conv_2(self.relu(self.bn_2(x, truncation, cond_vector)))
        x = self.conv_3(self.relu(self.bn_3(x, truncation, cond_vector)))
        if self.drop_channels:
            x0 = x0[:, : x0.shape[1] // 2]
        if self.up_sample:
            x0 = F.interpolate(x0, scale_factor=2, mode="nearest")
        return x + x0


class SelfAttn(nn.Module):
    """HF model.py SelfAttn (SAGAN attention with 2x2 max-pooled keys/values)."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.snconv1x1_theta = nn.Conv2d(in_channels, in_channels // 8, 1, bias=False)
        self.snconv1x1_phi = nn.Conv2d(in_channels, in_channels // 8, 1, bias=False)
        self.snconv1x1_g = nn.Conv2d(in_channels, in_channels // 2, 1, bias=False)
        self.snconv1x1_o_conv = nn.Conv2d(in_channels // 2, in_channels, 1, bias=False)
        self.maxpool = nn.MaxPool2d(2, stride=2, padding=0)
        self.softmax = nn.Softmax(dim=-1)
        self.gamma = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        _, ch, h, w = x.size()
        theta = self.snconv1x1_theta(x).view(-1, ch // 8, h * w)
        phi = self.maxpool(self.snconv1x1_phi(x)).view(-1, ch // 8, h * w // 4)
        attn = self.softmax(torch.bmm(theta.permute(0, 2, 1), phi))
        g = self.maxpool(self.snconv1x1_g(x)).view(-1, ch // 2, h * w // 4)
        attn_g = torch.bmm(g, attn.permute(0, 2, 1)).view(-1, ch // 2, h, w)
        return x + self.gamma * self.snconv1x1_o_conv(attn_g)


class Generator(nn.Module):
    """HF model.py Generator."""

    def __init__(self, config: BigGANConfig):
        super().__init__()
        self.config = config
        ch = config.channel_width
        cdim = config.z_dim * 2
        self.gen_z = nn.Linear(cdim, 4 * 4 * config.layers[0][1] * ch)
        layers = []
        for i, (up, cin, cout) in enumerate(config.layers):
            if i == config.attention_layer_position:
                layers.append(SelfAttn(ch * cin))
            layers.append(GenBlock(ch * cin, ch * cout, cdim, up_sample=up, n_stats=config.n_stats, eps=config.eps))
        self.layers = nn.ModuleList(layers)
        self.bn = BigGANBatchNorm(ch * config.layers[-1][2], n_stats=config.n_stats, eps=config.eps, conditional=False)
        self.relu = nn.ReLU()
        c_last = ch * config.layers[-1][2]
        self.conv_to_rgb = nn.Conv2d(c_last, c_last, 3, padding=1)
        self.tanh = nn.Tanh()

    def forward(self, cond_vector, truncation):
        z = self.gen_z(cond_vector)
        # TF (NHWC) weight convention -> NCHW
        z = z.view(-1, 4, 4, self.config.first_channels).permute(0, 3, 1, 2).contiguous()
        for layer in self.layers:
            z = layer(z, cond_vector, truncation) if isinstance(layer, GenBlock) else layer(z)
        z = self.conv_to_rgb(self.relu(self.bn(z, truncation)))
        return self.tanh(z[:, :3])


class BigGANOracle(nn.Module):
    """pix2latent/model/biggan.py:15-58 with a device argument instead of hard-coded .cuda()."""

    def __init__(self, config: BigGANConfig = None):
        super().__init__()
        self.config = config or BigGANConfig.deep256()
        self.embeddings = nn.Linear(self.config.num_classes, self.config.class_embed_dim, bias=False)
        self.generator = Generator(self.config)
        self.eval()

    def get_class_embedding(self, cls):  # biggan.py:37-47
        with torch.no_grad():
            w = self.embeddings.weight
            if type(cls) == int:
                c = torch.zeros(1, self.config.num_classes, dtype=w.dtype, device=w.device)
                c[:, cls] = 1
            elif len(cls.size()) == 2:
                c = cls
            else:
                raise ValueError
            return self.embeddings(c)

    def forward(self, z=None, c=None, truncation=1.0):  # biggan.py:50-58
        assert 0 < truncation <= 1
        assert len(z.size()) == 2, "expected z to be 2D"
        assert len(c.size()) == 2, "expected c to be 2D"
        assert c.size(1) == self.config.class_embed_dim, \
            "expected c to have dim (?, 128) but got {}".format(c.size())
        return self.generator(torch.cat((z, c), dim=1), truncation)


# ----------------------------------------------------------------------------- synthetic weights
@torch.no_grad()
def init_random_(model: BigGANOracle, seed=0, calibrate=True):
    """Deterministic random-init weights (no network => no pretrained checkpoint).

    conv/linear ~ N(0, gain/fan_in); BN tables perturbed; with ``calibrate`` the row-50 BN
    statistics are then set from a 4-sample forward pass so that every BN sees O(1) inputs, the
    regime the trained network operates in (keeps 60 layers of bf16 arithmetic well-scaled)."""
    g = torch.Generator().manual_seed(seed)

    def randn(*shape):
        return torch.randn(*shape, generator=g)

    for name, m in model.named_modules():
        if isinstance(m, nn.Conv2d):
            fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
            if name.endswith("conv_to_rgb"):
                gain = 0.3
            elif name.endswith(("conv_3", "snconv1x1_o_conv")):
                gain = 0.5
            else:
                gain = 1.0 if "snconv" in name else 2.0
            m.weight.copy_(randn(*m.weight.shape) * (gain / fan_in) ** 0.5)
            if m.bias is not None:
                m.bias.copy_(randn(*m.bias.shape) * 0.05)
        elif isinstance(m, nn.Linear):
            std = 0.03 if name.endswith((".scale", ".offset")) else (1.0 / m.in_features) ** 0.5
            m.weight.copy_(randn(*m.weight.shape) * std)
            if m.bias is not None:
                m.bias.copy_(randn(*m.bias.shape) * 0.05)
        elif isinstance(m, BigGANBatchNorm):
            m.running_means.copy_(randn(*m.running_means.shape) * 0.1)
            m.running_vars.copy_(1.0 + 0.2 * torch.rand(m.running_vars.shape, generator=g))
            if not m.conditional:
                m.weight.copy_(1.0 + 0.1 * randn(*m.weight.shape))
                m.bias.copy_(0.1 * randn(*m.bias.shape))
        elif isinstance(m, SelfAttn):
            m.gamma.fill_(0.5)
    if calibrate:
        zc = torch.fmod(randn(4, model.config.z_dim), 2.0)
        cc = model.embeddings.weight[:, :4].t().contiguous()
        hooks = []

        def pre_hook(mod, args):
            x = args[0]
            row = int(round(1.0 / mod.step_size))
            mod.running_means[row] = x.mean((0, 2, 3))
            mod.running_vars[row] = x.var((0, 2, 3), unbiased=False) + 1e-3

        for m in model.modules():
            if isinstance(m, BigGANBatchNorm):
                hooks.append(m.register_forward_pre_hook(pre_hook))
        model(z=zc.to(model.embeddings.weight.dtype), c=cc)
        for h in hooks:
            h.remove()
    return model


def make_biggan(config: BigGANConfig = None, seed=0, dtype=torch.float32, calibrate=True):
    torch_state = torch.random.get_rng_state()
    m = BigGANOracle(config)
    torch.random.set_rng_state(torch_state)  # constructing nn modules consumed RNG; undo
    init_random_(m, seed=seed, calibrate=calibrate)
    for p in m.parameters():
        p.requires_grad_(True)  # the reference leaves generator params trainable
    return m.to(dtype).eval()
