"""StyleGAN2 w / w+ / noise search on the GPU through the C-ABI, against the oracle's
rosinality Generator called as the reference does (pix2latent/model/stylegan2.py:122-125:
``model([w], input_is_latent=True, noise=noises)[0].clamp_(-1, 1)``) on the same weights:
image, gradient w.r.t. every latent row, gradient w.r.t. every noise image; the mapping network alone
(style(z)); the product model class with search='w+' (flat noise tensor, latent statistics)."""
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def tiny():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from oracle import stylegan2 as osg
    from pix2latent_b200.native import NativeStyleGAN2
    orc = osg.make_stylegan2(32, osg.TINY_CHANNELS, seed=0).cuda()
    nat = NativeStyleGAN2(32, dict(osg.TINY_CHANNELS), orc.model.state_dict())
    return orc, nat


def _ref_forward_w(orc, w, noise):
    return orc.model([w], input_is_latent=True, noise=noise)[0].clamp(-1.0, 1.0)


def test_mapping_network(tiny):
    from pix2latent_b200 import native
    orc, nat = tiny
    torch.manual_seed(0)
    z = torch.randn(37, 512, device="cuda")
    w_ref = orc.model.style(z)
    w = native.sg2_style(nat, z)
    assert rel(w, w_ref) < 1e-4
    assert nat.n_latent == orc.model.n_latent == 8


@pytest.mark.parametrize("plus", [False, True])
def test_forward_backward_w(tiny, plus):
    from pix2latent_b200 import native
    orc, nat = tiny
    torch.manual_seed(1)
    b = 3
    with torch.no_grad():
        w0 = orc.model.style(torch.randn(b, 512, device="cuda"))
    if plus:
        w0 = w0.unsqueeze(1).repeat(1, nat.n_latent, 1) + 0.3 * torch.randn(b, nat.n_latent, 512, device="cuda")
    w = w0.clone().requires_grad_(True)
    noise = [torch.randn(s, device="cuda").requires_grad_(True) for s in orc.model.noise_shapes(b)]
    ref = _ref_forward_w(orc, w, noise)
    img = native.sg2_forward_w(nat, w.detach(), [n.detach() for n in noise])
    torch.cuda.synchronize()
    print("w%s img rel %.4f" % ("+" if plus else "", rel(img, ref)))
    assert rel(img, ref) < 3e-2
    dimg = torch.randn_like(ref) * 1e-2
    ref.backward(dimg)
    dlat, dn = native.sg2_backward_w(nat, b, dimg)
    torch.cuda.synchronize()
    dw = dlat if plus else dlat.sum(1)
    print("   dw rel %.3f cos %.4f" % (rel(dw, w.grad), cos(dw, w.grad)))
    assert cos(dw, w.grad) > 0.98 and rel(dw, w.grad) < 0.2
    if plus:  # every latent row on its own (rows feed different layers)
        for r in range(nat.n_latent):
            assert cos(dlat[:, r], w.grad[:, r]) > 0.95, r
    for l, (g, n) in enumerate(zip(dn, noise)):
        assert g.shape == n.shape
        assert cos(g, n.grad) > 0.97, (l, cos(g, n.grad))
        assert rel(g, n.grad) < 0.25, (l, rel(g, n.grad))


def test_z_and_w_paths_agree_and_do_not_mix(tiny):
    """forward(z) == forward_w(style(z)); a z-search backward after forward_w is refused."""
    from pix2latent_b200 import _lib, native
    orc, nat = tiny
    torch.manual_seed(2)
    b = 2
    z = torch.randn(b, 512, device="cuda")
    noise = [torch.randn(s, device="cuda") for s in orc.model.noise_shapes(b)]
    img_z = nat.forward(z, noise)
    img_w = native.sg2_forward_w(nat, native.sg2_style(nat, z), noise)
    assert (img_z - img_w).abs().max().item() < 1e-5
    with pytest.raises(_lib.P2LError):
        nat.backward(b, torch.zeros_like(img_z))


def test_product_model_wplus_and_fused_step(tiny):
    from oracle import lpips as olp
    from pix2latent_b200 import native
    from pix2latent_b200.model.stylegan2 import StyleGAN2
    from test_biggan_gpu import lpips_native_state
    orc, _ = tiny
    from oracle import stylegan2 as osg
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = StyleGAN2(state_dict=orc.model.state_dict(), size=32, channels=dict(osg.TINY_CHANNELS), search="w+")
    assert model.latent_mean.shape == (512,) and model.latent_std.ndim == 0 and float(model.latent_std) > 0
    torch.manual_seed(3)
    b = 2
    n_noise = sum(s[-1] * s[-2] for s in model.noise_shape)
    w = (model.latent_mean[None, None] + 0.1 * torch.randn(b, model.n_latent, 512, device="cuda")).requires_grad_(True)
    nz = torch.randn(b, n_noise, device="cuda", requires_grad=True)
    out = model(w, nz)
    w2 = w.detach().clone().requires_grad_(True)
    nz2 = nz.detach().clone().requires_grad_(True)
    ref = _ref_forward_w(orc, w2, model.reshape_noise(nz2))
    assert rel(out, ref) < 3e-2
    g = torch.randn_like(ref) * 1e-2
    out.backward(g)
    ref.backward(g)
    assert cos(w.grad, w2.grad) > 0.98 and cos(nz.grad, nz2.grad) > 0.97
    # fused step with the pixel loss (32x32 is below the alex backbone's minimum size)
    lp = olp.make_lpips("alex", seed=0).cuda()
    nl = native.NativeLPIPS("alex", lpips_native_state(lp))
    target = torch.tanh(torch.randn(3, 32, 32, device="cuda"))
    tgt = nl.make_target(target, None, None, 1, 1.0, 0.0)
    noises = [n.contiguous() for n in model.reshape_noise(nz.detach())]
    loss, dlat, dn, img = native.sg2_step_w(model.native, nl, tgt, w.detach(), noises, True, 1.0 / b)
    w3 = w.detach().clone().requires_grad_(True)
    nz3 = nz.detach().clone().requires_grad_(True)
    ref_img = _ref_forward_w(orc, w3, model.reshape_noise(nz3))
    ref_loss = (target[None] - ref_img).abs().mean((1, 2, 3))
    ref_loss.mean().backward()
    assert torch.allclose(loss, ref_loss, rtol=3e-2, atol=3e-3)
    assert cos(dlat, w3.grad) > 0.95
    assert cos(torch.cat([t.reshape(b, -1) for t in dn], 1), nz3.grad) > 0.95
