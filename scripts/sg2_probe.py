"""Timing probe of the StyleGAN2 fused step (generator fwd + L1+10*alex-LPIPS + bwd to z) at the BASELINE configs:
    python scripts/sg2_probe.py cars 9     # LSUN-cars 512x512, chunk of 9, loss on rows 64:-64
    python scripts/sg2_probe.py ffhq 8     # FFHQ 1024x1024, 8 candidates per GPU
Seeded synthetic weights (no network). One JSON line: ms/step, candidates/s, eval-only candidates/s."""
import json
import sys
import warnings

sys.path.insert(0, ".")
warnings.filterwarnings("ignore")
import torch  # noqa: E402
from pix2latent_b200 import native  # noqa: E402
from pix2latent_b200.loss_functions import ProjectionLoss  # noqa: E402
from pix2latent_b200.model.stylegan2 import StyleGAN2  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cars"
b = int(sys.argv[2]) if len(sys.argv) > 2 else (9 if which == "cars" else 8)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
model = StyleGAN2(which, allow_synthetic=True)
R = model.im_res
loss_fn = ProjectionLoss(allow_synthetic=True)
g = torch.Generator().manual_seed(1)
target = torch.tanh(0.5 * torch.randn(3, R, R, generator=g)).cuda()
weight = torch.zeros(3, R, R, device="cuda")
if which == "cars":
    weight[:, R // 8:-(R // 8), :] = 1.0
else:
    weight[:] = 1.0
tgt = loss_fn.prepared_target(target, weight, weight if which == "cars" else None)
z = torch.fmod(torch.randn(b, 512, generator=g), 2.0).cuda()
noise = model.draw_noise(b, z.device)
out = {"workload": which, "b": b, "res": R}
for name, grad in (("step", True), ("eval_only", False)):
    f = lambda: native.sg2_step(model.native, loss_fn.native_lpips(), tgt, z, noise, grad, 1.0 / b, want_img=False)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = native.launch_count()
    e0.record()
    for _ in range(steps):
        loss, dz, _ = f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out[name] = {"ms": ms, "cand_per_s": b / ms * 1e3, "launches": (native.launch_count() - n0) // steps}
out["loss_mean"] = float(loss.mean())
print(json.dumps(out))
