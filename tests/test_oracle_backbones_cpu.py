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


def test_biggan_self_attention_matches_torch_sdpa():
    """The oracle's SelfAttn (HF model.py: theta / 2x2-max-pooled phi, g; softmax over the pooled keys, no 1/sqrt(d) scale)
    against torch's own fused scaled_dot_product_attention called with scale=1 — an implementation this repo did not write."""
    import torch.nn.functional as F
    from oracle import biggan as obg
    torch.manual_seed(0)
    att = obg.SelfAttn(64).double()
    for p in att.parameters():
        torch.nn.init.normal_(p, std=0.2)
    x = torch.randn(2, 64, 8, 8, dtype=torch.float64)
    out = att(x)
    b, ch, h, w = x.shape
    q = att.snconv1x1_theta(x).view(b, ch // 8, h * w).transpose(1, 2)                       # [b, HW, C/8]
    k = F.max_pool2d(att.snconv1x1_phi(x), 2).view(b, ch // 8, h * w // 4).transpose(1, 2)    # [b, HW/4, C/8]
    v = F.max_pool2d(att.snconv1x1_g(x), 2).view(b, ch // 2, h * w // 4).transpose(1, 2)      # [b, HW/4, C/2]
    o = F.scaled_dot_product_attention(q, k, v, scale=1.0).transpose(1, 2).reshape(b, ch // 2, h, w)
    ref = x + att.gamma * att.snconv1x1_o_conv(o)
    assert torch.allclose(out, ref, rtol=1e-10, atol=1e-12)


def test_biggan_conditional_bn_matches_torch_batch_norm():
    """BigGANBatchNorm (eval statistics row + per-sample gain / offset from the condition vector) against F.batch_norm
    followed by the per-sample affine."""
    import torch.nn.functional as F
    from oracle import biggan as obg
    torch.manual_seed(1)
    bn = obg.BigGANBatchNorm(16, condition_vector_dim=8, n_stats=51, eps=1e-4).double()
    bn.running_means.normal_()
    bn.running_vars.uniform_(0.5, 2.0)
    x = torch.randn(3, 16, 5, 5, dtype=torch.float64)
    cond = torch.randn(3, 8, dtype=torch.float64)
    for trunc in (1.0, 0.5, 0.43):                      # row 50, row 25, interpolation between rows 21 and 22
        out = bn(x, trunc, cond)
        mean, var = bn.stats(trunc)
        ref = F.batch_norm(x, mean, var, None, None, False, 0.0, 1e-4)
        ref = ref * (1 + bn.scale(cond))[:, :, None, None] + bn.offset(cond)[:, :, None, None]
        assert torch.allclose(out, ref, rtol=1e-10, atol=1e-12)
    coef = 0.43 / 0.02 - 21
    m, _ = bn.stats(0.43)
    assert torch.allclose(m, bn.running_means[21] * coef + bn.running_means[22] * (1 - coef), rtol=1e-9)
