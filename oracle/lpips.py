"""ORACLE (test infrastructure) — restatement of ``lpips.LPIPS(net, spatial=True)`` as called by
pix2latent/loss_functions.py:131,142 and of the reference's loss algebra
(pix2latent/loss_functions.py:20-27, 86-148).

``lpips>=0.1`` (requirements.txt:15) is a third-party package absent from /root/reference and this
image; its published algorithm (lpips/lpips.py: LPIPS, ScalingLayer, NetLinLayer,
normalize_tensor, upsample; lpips/pretrained_networks.py: alexnet / vgg16 slices) is restated
here — SURVEY.md Appendix A.2.
PINNED (partially): the backbones — where the FLOPs are — are checked against the INSTALLED
torchvision ``alexnet().features`` / ``vgg16().features`` on the same weights at the tap indices the real
package uses (tests/test_oracle_backbones_cpu.py); the loss classes below are checked against the real
reference code (tests/golden/make_golden.py imports /root/reference with this module standing in for
``lpips``). PARITY UNPINNED for the small algebra in between, restated from the package's published
source: ScalingLayer constants, unit-normalisation eps, the 1x1 ``lin`` layers, the bilinear up-sampling.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

ALEX_CFG = [  # (Cin, Cout, k, stride, pad, maxpool_before)
    (3, 64, 11, 4, 2, False),
    (64, 192, 5, 1, 2, True),
    (192, 384, 3, 1, 1, True),
    (384, 256, 3, 1, 1, False),
    (256, 256, 3, 1, 1, False),
]
# vgg16: relu1_2, relu2_2, relu3_3, relu4_3, relu5_3
VGG_CFG = [
    [(3, 64), (64, 64)],
    [(64, 128), (128, 128)],
    [(128, 256), (256, 256), (256, 256)],
    [(256, 512), (512, 512), (512, 512)],
    [(512, 512), (512, 512), (512, 512)],
]


class ScalingLayer(nn.Module):  # lpips/lpips.py ScalingLayer
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.Tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, inp):
        return (inp - self.shift) / self.scale


class AlexFeatures(nn.Module):
    """lpips/pretrained_networks.py alexnet: slices of torchvision alexnet.features
    [0:2], [2:5], [5:8], [8:10], [10:12]; module names slice1..slice5 with the torchvision
    indices as child names so official state dicts load."""
    chns = [64, 192, 384, 256, 256]

    def __init__(self):
        super().__init__()
        idx = [0, 3, 6, 8, 10]
        self.slices = nn.ModuleList()
        for k, (ci, co, ks, st, pd, pool) in enumerate(ALEX_CFG):
            seq = nn.Sequential()
            if pool:
                seq.add_module(str(idx[k] - 1), nn.MaxPool2d(kernel_size=3, stride=2))
            seq.add_module(str(idx[k]), nn.Conv2d(ci, co, ks, st, pd))
            seq.add_module(str(idx[k] + 1), nn.ReLU(inplace=False))
            self.slices.append(seq)

    def forward(self, x):
        outs = []
        for s in self.slices:
            x = s(x)
            outs.append(x)
        return outs


class VGGFeatures(nn.Module):
    """lpips/pretrained_networks.py vgg16: torchvision vgg16.features [0:4],[4:9],[9:16],[16:23],[23:30]."""
    chns = [64, 128, 256, 512, 512]

    def __init__(self):
        super().__init__()
        self.slices = nn.ModuleList()
        i = 0
        for k, convs in enumerate(VGG_CFG):
            seq = nn.Sequential()
            if k > 0:
                seq.add_module(str(i), nn.MaxPool2d(kernel_size=2, stride=2))
                i += 1
            for ci, co in convs:
                seq.add_module(str(i), nn.Conv2d(ci, co, 3, padding=1))
                seq.add_module(str(i + 1), nn.ReLU(inplace=False))
                i += 2
            self.slices.append(seq)

    def forward(self, x):
        outs = []
        for s in self.slices:
            x = s(x)
            outs.append(x)
        return outs


def normalize_tensor(in_feat, eps=1e-10):  # lpips/__init__.py normalize_tensor
    norm_factor = torch.sqrt(torch.sum(in_feat ** 2, dim=1, keepdim=True))
    return in_feat / (norm_factor + eps)


def upsample(in_tens, out_HW):  # lpips/lpips.py upsample
    return F.interpolate(in_tens, size=out_HW, mode="bilinear", align_corners=False)


class LPIPSOracle(nn.Module):
    """lpips.LPIPS(net=..., spatial=...) forward, eval mode (dropout inactive)."""

    def __init__(self, net="alex", spatial=True):
        super().__init__()
        self.pnet_type, self.spatial = net, spatial
        self.scaling_layer = ScalingLayer()
        if net in ("alex", "alexnet"):
            self.net = AlexFeatures()
        elif net in ("vgg", "vgg16"):
            self.net = VGGFeatures()
        else:
            raise ValueError("unsupported lpips net %r" % net)
        self.chns = self.net.chns
        # linK.model.1.weight in the real package; flat [C] here
        self.lins = nn.ParameterList([nn.Parameter(torch.ones(c)) for c in self.chns])
        for p in self.parameters():
            p.requires_grad_(False)  # lpips freezes everything (pnet_tune=False, lins eval)
        self.eval()

    def forward(self, in0, in1):
        in0_input, in1_input = self.scaling_layer(in0), self.scaling_layer(in1)
        outs0, outs1 = self.net(in0_input), self.net(in1_input)
        res = []
        for kk in range(len(self.chns)):
            f0, f1 = normalize_tensor(outs0[kk]), normalize_tensor(outs1[kk])
            diff = (f0 - f1) ** 2
            lin = (diff * self.lins[kk][None, :, None, None]).sum(1, keepdim=True)  # 1x1 conv C->1, no bias
            if self.spatial:
                res.append(upsample(lin, in0.shape[2:]))
            else:
                res.append(lin.mean([2, 3], keepdim=True))
        val = res[0]
        for r in res[1:]:
            val = val + r
        return val


@torch.no_grad()
def init_random_(m: LPIPSOracle, seed=0):
    """Seeded synthetic weights: He-normal convs, non-negative ``lin`` weights (as in the real
    package, so the loss is a valid distance)."""
    g = torch.Generator().manual_seed(1000 + seed)
    for mod in m.net.modules():
        if isinstance(mod, nn.Conv2d):
            fan_in = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1]
            mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
            mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.05)
    for k, c in enumerate(m.chns):
        m.lins[k].copy_(torch.randn(c, generator=g).abs() * (2.0 / c))
    return m


def make_lpips(net="alex", seed=0, dtype=torch.float32, spatial=True):
    st = torch.random.get_rng_state()
    m = LPIPSOracle(net, spatial)
    torch.random.set_rng_state(st)
    init_random_(m, seed)
    return m.to(dtype).eval()


# ----------------------------------------------------------------------------- reference losses
def l1_loss(out, target):  # loss_functions.py:20-22
    return torch.abs(target - out)


def l2_loss(out, target):  # loss_functions.py:25-27
    return (target - out) ** 2


class ReconstructionLoss(nn.Module):  # loss_functions.py:104-124
    def __init__(self, loss_type="l1"):
        super().__init__()
        if loss_type in ["l1", 1]:
            self.loss_fn = l1_loss
        elif loss_type in ["l2", 2]:
            self.loss_fn = l2_loss
        else:
            raise ValueError("Unknown loss_type {}".format(loss_type))

    def __call__(self, output, target, weight=None, loss_mask=None):
        loss = self.loss_fn(output, target)
        if weight is not None:
            _weight = weight if loss_mask is None else (loss_mask * weight)
            n = torch.sum(loss * _weight, [1, 2, 3])
            d = torch.sum(_weight, [1, 2, 3])
            loss = n / d
        return loss


class PerceptualLoss(nn.Module):  # loss_functions.py:127-148
    def __init__(self, net="vgg", lpips_module=None, seed=0, dtype=torch.float32):
        super().__init__()
        self.loss_fn = lpips_module if lpips_module is not None else make_lpips(net, seed, dtype)

    def __call__(self, output, target, weight=None, loss_mask=None):
        loss = self.loss_fn(output, target)
        if weight is not None:
            _weight = weight if loss_mask is None else (loss_mask * weight)
            n = torch.sum(loss * _weight, [1, 2, 3])
            d = torch.sum(_weight, [1, 2, 3])
            loss = n / d
        return loss


class ProjectionLoss(nn.Module):  # loss_functions.py:86-100
    def __init__(self, lpips_net="alex", beta=10, lpips_module=None, seed=0, dtype=torch.float32):
        super().__init__()
        self.beta = beta
        self.rloss_fn = ReconstructionLoss()
        self.ploss_fn = PerceptualLoss(net=lpips_net, lpips_module=lpips_module, seed=seed, dtype=dtype)

    def __call__(self, output, target, weight=None, loss_mask=None):
        rec_loss = self.rloss_fn(output, target, weight, loss_mask)
        per_loss = self.ploss_fn(output, target, weight, loss_mask)
        return rec_loss + (self.beta * per_loss)
