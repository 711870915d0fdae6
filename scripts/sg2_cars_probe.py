"""Timing probe: StyleGAN2-cars 512^2 inner step (BASELINE.json configs[2] shapes: population 22 in
chunks 9/9/4, loss mask rows 64:-64), native path."""
import sys, warnings
sys.path.insert(0, ".")
warnings.filterwarnings("ignore")
import torch
from pix2latent_b200 import native
from pix2latent_b200.loss_functions import ProjectionLoss
from pix2latent_b200.model.stylegan2 import StyleGAN2
model = StyleGAN2("cars", allow_synthetic=True)
loss_fn = ProjectionLoss(allow_synthetic=True)
g = torch.Generator().manual_seed(1)
target = torch.tanh(0.5 * torch.randn(3, 512, 512, generator=g)).cuda()
weight = torch.zeros(3, 512, 512).cuda(); weight[:, 64:-64, :] = 1
tgt = loss_fn.prepared_target(target, weight, weight)
for b in (9, 4):
    z = torch.randn(b, 512).cuda()
    f = lambda: native.sg2_step(model.native, loss_fn.native_lpips(), tgt, z, model.draw_noise(b, "cuda"), True, 1.0 / b, want_img=False)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("StyleGAN2-cars 512: b=%d %.2f ms/step -> %.1f candidates/s ; loss %s ; mem %.1f GB" % (b, ms, b / ms * 1e3, out[0][:2].tolist(), torch.cuda.max_memory_allocated() / 1e9))
