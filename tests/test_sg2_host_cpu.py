"""Host logic of the StyleGAN2 step (closure._step_native_sg2) on CPU. The product runs the reference's chunks
(/root/reference pix2latent/optimizer/closure.py:32-66: split_vars -> per chunk: zero_grad, hooks, forward (the generator
draws fresh per-layer noise), loss.mean().backward(), opt.step()) as ONE physical batch with per-candidate 1/b_chunk
scales, drawing the RNG chunk by chunk in the reference's order. Here the native call is replaced by a torch stand-in and
the result is compared with the reference's own procedure (the autograd path, chunk by chunk) on the same toy generator:
same latents, same losses, same RNG consumption."""
import numpy as np
import pytest
import torch


class ToySG2(torch.nn.Module):
    """img[b,3,4,4] from z[b,8] and two per-layer noise images; noise drawn from the global RNG when not given (as
    rosinality's NoiseInjection does)."""
    noise_shape = [(1, 1, 2, 2), (1, 1, 4, 4)]

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.A = torch.nn.Parameter(torch.randn(48, 8, generator=g) * 0.5, requires_grad=False)
        self.native = object()
        self.search = "z"

    def draw_noise(self, b, device):
        return [torch.randn(b, 1, s[2], s[3], device=device) for s in self.noise_shape]

    def forward(self, z, noises=None):
        if noises is None:
            noises = self.draw_noise(z.shape[0], z.device)
        x = torch.tanh(z @ self.A.T).view(-1, 3, 4, 4)
        x = x + 0.3 * noises[1] + 0.2 * torch.nn.functional.interpolate(noises[0], scale_factor=2.0, mode="nearest")
        return x * (1 + 0.1 * z.mean(1).view(-1, 1, 1, 1))


class ToyLoss:
    def __call__(self, out, target, weight=None, loss_mask=None):
        return ((out - target).abs() * weight).flatten(1).sum(1) / weight.flatten(1).sum(1)

    def prepared_target(self, target, weight, mask):
        return (target, weight)

    def native_lpips(self):
        return None


def _vm(hook):
    from pix2latent_b200 import VariableManager
    torch.manual_seed(3)
    vm = VariableManager(device="cpu")
    vm.register("z", (8,), "input", learning_rate=0.05, hook_fn=hook)
    vm.register("target", (3, 4, 4), "output", requires_grad=False, default=torch.tanh(torch.randn(3, 4, 4)))
    vm.register("weight", (3, 4, 4), "output", requires_grad=False, default=torch.rand(3, 4, 4) + 0.2)
    return vm


@pytest.mark.parametrize("n,chunk,phys", [(5, 2, 24), (7, 3, 4), (4, 9, 24)])
def test_one_physical_batch_equals_the_reference_chunks(monkeypatch, n, chunk, phys):
    from pix2latent_b200 import native
    from pix2latent_b200.optimizer import closure
    from pix2latent_b200.utils import function_hooks as hk
    model, loss_fn = ToySG2(), ToyLoss()
    calls = []

    def fake_sg2_step(gen, lp, tgt, z, noises, want_grad, grad_scale, want_img=True, dloss=None):
        target, weight = tgt
        calls.append(z.shape[0])
        zz = z.detach().clone().requires_grad_(want_grad)
        img = model(zz, noises)
        b = z.shape[0]
        loss = loss_fn(img, target[None].expand(b, -1, -1, -1), weight[None].expand(b, -1, -1, -1))
        dz = None
        if want_grad:
            (loss * (dloss if dloss is not None else 1.0) * grad_scale).sum().backward()
            dz = zz.grad
        return loss.detach(), dz, img.detach()

    monkeypatch.setattr(native, "sg2_step", fake_sg2_step)
    monkeypatch.setattr(closure, "SG2_PHYS_BATCH", phys)
    res = {}
    for mode in ("reference", "product"):
        vm = _vm(hk.NormalPerturb(0.05))
        torch.manual_seed(11)
        variables = vm.initialize(n)
        torch.manual_seed(12)
        losses = []
        for _ in range(3):
            if mode == "product":
                out, l, _ = closure._step_native_sg2(model, variables, loss_fn, True, chunk)
            else:
                out, l, _ = closure._step_autograd(model, variables, loss_fn, True, chunk)
            losses.append(np.array(l, dtype=np.float64))
        # an evaluation-only pass and the state of the global RNG afterwards
        if mode == "product":
            out, l, _ = closure._step_native_sg2(model, variables, loss_fn, False, chunk)
        else:
            out, l, _ = closure._step_autograd(model, variables, loss_fn, False, chunk)
        losses.append(np.array(l, dtype=np.float64))
        res[mode] = (torch.stack(variables.input.z.data).detach().clone(), np.stack(losses), out.detach().clone(),
                     torch.rand(1).item())
    (zr, lr, outr, rr), (zp, lp_, outp, rp) = res["reference"], res["product"]
    assert rr == rp, "the product consumed the global RNG differently from the reference's chunk loop"
    assert np.allclose(lr, lp_, atol=1e-6), np.abs(lr - lp_).max()
    assert torch.allclose(zr, zp, atol=1e-6), (zr - zp).abs().max()
    assert torch.allclose(outr, outp, atol=1e-6)
    per_pass = -(-n // phys)
    assert calls == ([min(phys, n - i * phys) for i in range(per_pass)] * 4), calls
