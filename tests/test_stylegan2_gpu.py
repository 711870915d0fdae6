"""Parity of the native StyleGAN2 generator (z search) against oracle/stylegan2.py on the same
weights, latents and per-layer noise; then the fused step with the projection loss.
Reduced config (32x32, 128/64 channels) exercises every layer type: const input, plain and
up-sampling styled convs (zero-inserted grid + FIR), noise, ToRGB with up-sampled skips, clamp."""
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
    ch = dict(osg.TINY_CHANNELS)
    nat = NativeStyleGAN2(32, ch, orc.model.state_dict())
    return orc, nat


def test_forward_backward(tiny):
    orc, nat = tiny
    torch.manual_seed(1)
    b = 3
    z = torch.randn(b, 512, device="cuda", requires_grad=True)
    noise = [torch.randn(s, device="cuda") for s in orc.model.noise_shapes(b)]
    assert nat.num_layers == len(noise)
    ref = orc(z, noise)
    img = nat.forward(z.detach(), noise)
    torch.cuda.synchronize()
    print("img rel %.4f max %.4f" % (rel(img, ref), (img - ref).abs().max().item()))
    assert rel(img, ref) < 3e-2
    dimg = torch.randn_like(ref) * 1e-2
    ref.backward(dimg)
    dz = nat.backward(b, dimg)
    torch.cuda.synchronize()
    print("dz rel %.3f cos %.4f" % (rel(dz, z.grad), cos(dz, z.grad)))
    assert cos(dz, z.grad) > 0.98 and rel(dz, z.grad) < 0.2


def test_no_noise_and_fused_step(tiny):
    orc, nat = tiny
    from oracle import lpips as olp
    from pix2latent_b200.native import NativeLPIPS, sg2_step
    from test_biggan_gpu import lpips_native_state
    torch.manual_seed(2)
    b = 2
    z = torch.randn(b, 512, device="cuda", requires_grad=True)
    noise = [torch.randn(s, device="cuda") for s in orc.model.noise_shapes(b)]
    lp = olp.make_lpips("alex", seed=0).cuda()
    # 32x32 is too small for the alex backbone: compare the pixel term + image only via ReconstructionLoss-like target
    target = torch.tanh(torch.randn(3, 32, 32, device="cuda"))
    nl = NativeLPIPS("alex", lpips_native_state(lp))
    tgt = nl.make_target(target, None, None, 1, 1.0, 0.0)   # rec_weight 1, per_weight 0
    loss, dz, img = sg2_step(nat, nl, tgt, z.detach(), noise, True, 1.0 / b)
    ref_img = orc(z, noise)
    ref_loss = (target[None] - ref_img).abs().mean((1, 2, 3))
    ref_loss.mean().backward()
    torch.cuda.synchronize()
    print("loss", loss.tolist(), ref_loss.tolist(), "cos", cos(dz, z.grad))
    assert torch.allclose(loss, ref_loss, rtol=3e-2, atol=3e-3)
    assert cos(dz, z.grad) > 0.95


def test_narrow_level_is_zero_padded_exactly():
    """ffhq-1024's top level has 32 channels; the wrapper pads such levels to 64 (zero weights, fan-in
    scale compensated). Reduced config with a 32-channel level, through the product model class."""
    import warnings
    from oracle import stylegan2 as osg
    from pix2latent_b200.model.stylegan2 import StyleGAN2
    ch = {4: 128, 8: 64, 16: 32}
    orc = osg.make_stylegan2(16, ch, seed=1).cuda()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = StyleGAN2(state_dict=orc.model.state_dict(), size=16, channels=ch)
    assert model.channels[16] == 64
    torch.manual_seed(4)
    z = torch.randn(2, 512, device="cuda", requires_grad=True)
    noise = [torch.randn(s, device="cuda") for s in orc.model.noise_shapes(2)]
    ref = orc(z, noise)
    z2 = z.detach().clone().requires_grad_(True)
    out = model(z2, noise)
    assert rel(out, ref) < 3e-2
    g = torch.randn_like(ref) * 1e-2
    ref.backward(g)
    out.backward(g)
    assert cos(z2.grad, z.grad) > 0.98
