"""Partial pin of the third-party LPIPS arithmetic: ``lpips/pretrained_networks.py`` wraps
``torchvision.models.alexnet().features`` / ``vgg16().features`` and taps them after the ReLUs at indices
(1, 4, 7, 9, 11) / (3, 8, 15, 22, 29). The oracle's hand-written backbones (oracle/lpips.py AlexFeatures /
VGGFeatures: kernel sizes, strides, paddings, where the max-pools sit) are checked here against the INSTALLED
torchvision modules on the same weights — an implementation this repo did not write. What stays unpinned is only
the small algebra around them (ScalingLayer constants, unit-normalisation eps, the 1x1 `lin` layers, the
bilinear up-sampling), restated from the package's published source."""
import pytest
import torch


def _into_torchvision(oracle_net, tv_features):
    """oracle keys 'slices.{k}.{idx}.weight' -> torchvision 'features.{idx}.weight' (same indices by construction)"""
    sd = {}
    for name, p in oracle_net.state_dict().items():
        _, _, idx, kind = name.split(".")
        sd["%s.%s" % (idx, kind)] = p
    missing, unexpected = tv_features.load_state_dict(sd, strict=True), None
    return tv_features.eval()


@pytest.mark.parametrize("net,taps,size", [("alex", (1, 4, 7, 9, 11), 96), ("vgg", (3, 8, 15, 22, 29), 48)])
def test_backbone_matches_torchvision(net, taps, size):
    import torchvision
    from oracle import lpips as olp
    m = olp.make_lpips(net, seed=0)
    tv = torchvision.models.alexnet(weights=None) if net == "alex" else torchvision.models.vgg16(weights=None)
    feats = _into_torchvision(m.net, tv.features)
    torch.manual_seed(0)
    x = torch.randn(2, 3, size, size)
    ours = m.net(x)
    ref, h = [], x
    with torch.no_grad():
        for i, layer in enumerate(feats):
            h = layer(h)
            if i in taps:
                ref.append(h)
    assert len(ours) == len(ref) == 5
    for a, b in zip(ours, ref):
        assert a.shape == b.shape
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    # channel counts the `lin` layers are sized by (lpips/lpips.py: chns)
    assert [o.shape[1] for o in ours] == m.chns


def test_scaling_layer_constants_and_distance_properties():
    """ScalingLayer constants as published (lpips/lpips.py), and the metric properties any LPIPS must have:
    d(x, x) = 0, d >= 0 with non-negative lin weights, symmetry."""
    from oracle import lpips as olp
    m = olp.make_lpips("alex", seed=0)
    assert torch.allclose(m.scaling_layer.shift.flatten(), torch.tensor([-.030, -.088, -.188]))
    assert torch.allclose(m.scaling_layer.scale.flatten(), torch.tensor([.458, .448, .450]))
    torch.manual_seed(1)
    a, b = torch.tanh(torch.randn(1, 3, 64, 64)), torch.tanh(torch.randn(1, 3, 64, 64))
    assert m(a, a).abs().max().item() == 0.0
    dab, dba = m(a, b), m(b, a)
    assert dab.shape == (1, 1, 64, 64) and dab.min().item() >= 0.0
    assert torch.allclose(dab, dba, rtol=1e-5, atol=1e-7)
