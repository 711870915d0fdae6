"""Seeded synthetic weights for the generator / perceptual nets.

There is no network in the build environment, so the official checkpoints
(pytorch_pretrained_biggan's S3 files, lpips' linear layers, rosinality's StyleGAN2 .pt) cannot
be fetched. These factories produce state dicts with the OFFICIAL key names and shapes, so the
same loaders accept real checkpoints when they are available. Plain tensor code:
no nn.Module, nothing from oracle/.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

import torch


@dataclass
class BigGANConfig:
    """pytorch_pretrained_biggan/config.py::BigGANConfig (biggan-deep-256 defaults)."""
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


def biggan_state_dict(config: BigGANConfig = None, seed=0):
    """Random-init BigGAN-deep weights under pix2latent's BigGAN module keys (after
    remove_spectral_norm): 'embeddings.weight', 'generator.gen_z.*', 'generator.layers.{i}.*',
    'generator.bn.*', 'generator.conv_to_rgb.*'."""
    cfg = config or BigGANConfig()
    g = torch.Generator().manual_seed(seed)
    sd = {}
    ch, cdim = cfg.channel_width, cfg.z_dim * 2

    def randn(*s):
        return torch.randn(*s, generator=g)

    def conv(name, co, ci, k, gain, bias=True):
        sd[name + ".weight"] = randn(co, ci, k, k) * (gain / (ci * k * k)) ** 0.5
        if bias:
            sd[name + ".bias"] = randn(co) * 0.05

    def bn(name, c, conditional=True):
        sd[name + ".running_means"] = randn(cfg.n_stats, c) * 0.1
        sd[name + ".running_vars"] = 1.0 + 0.2 * torch.rand(cfg.n_stats, c, generator=g)
        if conditional:
            sd[name + ".scale.weight"] = randn(c, cdim) * 0.03
            sd[name + ".offset.weight"] = randn(c, cdim) * 0.03
        else:
            sd[name + ".weight"] = 1.0 + 0.1 * randn(c)
            sd[name + ".bias"] = 0.1 * randn(c)

    sd["embeddings.weight"] = randn(cfg.class_embed_dim, cfg.num_classes) * (1.0 / cfg.class_embed_dim) ** 0.5 * 4.0
    c0 = ch * cfg.layers[0][1]
    sd["generator.gen_z.weight"] = randn(16 * c0, cdim) * (1.0 / cdim) ** 0.5
    sd["generator.gen_z.bias"] = randn(16 * c0) * 0.05
    idx = 0
    for i, (up, cin, cout) in enumerate(cfg.layers):
        if i == cfg.attention_layer_position:
            c = ch * cin
            p = "generator.layers.%d." % idx
            conv(p + "snconv1x1_theta", c // 8, c, 1, 1.0, bias=False)
            conv(p + "snconv1x1_phi", c // 8, c, 1, 1.0, bias=False)
            conv(p + "snconv1x1_g", c // 2, c, 1, 1.0, bias=False)
            conv(p + "snconv1x1_o_conv", c, c // 2, 1, 0.5, bias=False)
            sd[p + "gamma"] = torch.full((1,), 0.5)
            idx += 1
        ci, co = ch * cin, ch * cout
        mid = ci // 4
        p = "generator.layers.%d." % idx
        bn(p + "bn_0", ci); conv(p + "conv_0", mid, ci, 1, 2.0)
        bn(p + "bn_1", mid); conv(p + "conv_1", mid, mid, 3, 2.0)
        bn(p + "bn_2", mid); conv(p + "conv_2", mid, mid, 3, 2.0)
        bn(p + "bn_3", mid); conv(p + "conv_3", co, mid, 1, 0.2)
        idx += 1
    c_last = ch * cfg.layers[-1][2]
    bn("generator.bn", c_last, conditional=False)
    conv("generator.conv_to_rgb", c_last, c_last, 3, 0.3)
    return sd


ALEX_CFG = [(3, 64, 11, 0), (64, 192, 5, 3), (192, 384, 3, 6), (384, 256, 3, 8), (256, 256, 3, 10)]
VGG_SLICES = [[(3, 64), (64, 64)], [(64, 128), (128, 128)], [(128, 256), (256, 256), (256, 256)],
              [(256, 512), (512, 512), (512, 512)], [(512, 512), (512, 512), (512, 512)]]


def lpips_state_dict(net="alex", seed=0):
    """Random-init LPIPS weights under the native loader's keys: 'net.slice{k}.{idx}.weight|bias'
    (torchvision feature indices, as in lpips/pretrained_networks.py) and 'lin{k}.weight' [C]
    (non-negative, like the trained linear layers)."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd = {}

    def conv(name, co, ci, k):
        sd[name + ".weight"] = torch.randn(co, ci, k, k, generator=g) * (2.0 / (ci * k * k)) ** 0.5
        sd[name + ".bias"] = torch.randn(co, generator=g) * 0.05

    chns = []
    if net in ("alex", "alexnet"):
        for k, (ci, co, ks, idx) in enumerate(ALEX_CFG):
            conv("net.slice%d.%d" % (k + 1, idx), co, ci, ks)
            chns.append(co)
    elif net in ("vgg", "vgg16"):
        i = 0
        for s, convs in enumerate(VGG_SLICES):
            if s > 0:
                i += 1
            for ci, co in convs:
                conv("net.slice%d.%d" % (s + 1, i), co, ci, 3)
                i += 2
            chns.append(convs[-1][1])
    else:
        raise ValueError("unsupported lpips net %r" % (net,))
    for k, c in enumerate(chns):
        sd["lin%d.weight" % k] = torch.randn(c, generator=g).abs() * (2.0 / c)
    return sd


def lpips_state_from_package(sd):
    """Convert a real ``lpips.LPIPS`` state dict ('lin0.model.1.weight' [1,C,1,1], 'net.slice1.0.weight')
    to the native loader's keys."""
    out = {}
    for k, v in sd.items():
        if k.startswith("lin") and k.endswith(".model.1.weight"):
            out[k.split(".")[0] + ".weight"] = v.reshape(-1)
        elif k.startswith("net.slice"):
            out[k] = v
    return out


def stylegan2_state_dict(size=512, channels=None, seed=0):
    """Random-init rosinality ``g_ema`` state dict (keys 'style.k.*', 'input.input', 'conv1.*',
    'to_rgb1.*', 'convs.i.*', 'to_rgbs.i.*'); N(0,1) weights under equalised-lr scaling, non-zero
    noise strengths / biases, ToRGB weights scaled down so the summed skip image stays in range."""
    ch = channels or {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}
    g = torch.Generator().manual_seed(2000 + seed)
    sd = {}

    def randn(*s):
        return torch.randn(*s, generator=g)

    for k in range(1, 9):
        sd["style.%d.weight" % k] = randn(512, 512) / 0.01
        sd["style.%d.bias" % k] = randn(512)
    sd["input.input"] = randn(1, ch[4], 4, 4)

    def styled(pre, cin, cout):
        sd[pre + ".conv.weight"] = randn(1, cout, cin, 3, 3)
        sd[pre + ".conv.modulation.weight"] = randn(cin, 512)
        sd[pre + ".conv.modulation.bias"] = torch.ones(cin)
        sd[pre + ".noise.weight"] = 0.1 * randn(1)
        sd[pre + ".activate.bias"] = 0.1 * randn(cout)

    def torgb(pre, cin):
        sd[pre + ".conv.weight"] = 0.25 * randn(1, 3, cin, 1, 1)
        sd[pre + ".conv.modulation.weight"] = randn(cin, 512)
        sd[pre + ".conv.modulation.bias"] = torch.ones(cin)
        sd[pre + ".bias"] = 0.1 * randn(1, 3, 1, 1)

    styled("conv1", ch[4], ch[4])
    torgb("to_rgb1", ch[4])
    log_size = size.bit_length() - 1
    cin = ch[4]
    for i in range(3, log_size + 1):
        cout = ch[2 ** i]
        styled("convs.%d" % (2 * (i - 3)), cin, cout)
        styled("convs.%d" % (2 * (i - 3) + 1), cout, cout)
        torgb("to_rgbs.%d" % (i - 3), cout)
        cin = cout
    return sd
