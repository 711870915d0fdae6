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
