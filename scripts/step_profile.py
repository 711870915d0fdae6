"""Per-launch timing of the tensor-core kernels inside one real bench step (CUDA events around every launch), as JSON
lines: one header {"ms_per_step", "conv_ms", "launches"} and one record per launch. Options are taken from P2L_OPTS;
P2L_LIB selects another build of the library (A/B of two builds on the same box).
    python scripts/step_profile.py > gpurun_out/profile.jsonl"""
import json
import sys
import warnings

sys.path.insert(0, ".")
warnings.filterwarnings("ignore")
import torch  # noqa: E402
from bench import synthetic_target  # noqa: E402
from pix2latent_b200 import _lib, native  # noqa: E402
from pix2latent_b200.loss_functions import ProjectionLoss  # noqa: E402
from pix2latent_b200.model import BigGAN  # noqa: E402

n = 18
target, weight = synthetic_target(256, "cuda")
g = torch.Generator().manual_seed(2)
z = torch.fmod(torch.randn(n, 128, generator=g), 2.0).cuda()
model = BigGAN(seed=0, allow_synthetic=True).cuda()
loss_fn = ProjectionLoss(allow_synthetic=True)
tgt = loss_fn.prepared_target(target, weight)
c = model.get_class_embedding(153).repeat(n, 1).contiguous()


def f():
    return native.biggan_step(model.native, loss_fn.native_lpips(), tgt, z, c, True, 1 / 9, want_img=False)


for _ in range(3):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    f()
e1.record()
torch.cuda.synchronize()
ms_step = e0.elapsed_time(e1) / 20
# three instrumented passes, per-launch minimum (the events serialise the launches: PDL overlap is lost, tails are exposed)
best = None
for _ in range(3):
    native.profile_enable(1)
    f()
    torch.cuda.synchronize()
    recs = _lib.profile_records()
    native.profile_read()
    native.profile_enable(0)
    best = recs if best is None else [r if r[0] < b[0] else b for r, b in zip(recs, best)]
print(json.dumps({"ms_per_step": ms_step, "conv_ms": sum(r[0] for r in best), "launches": len(best)}))
for i, (ms, fl, BN, mode, hl, grid, M, N, K) in enumerate(best):
    print(json.dumps({"i": i, "mode": "bwd" if mode else "fwd", "BN": BN, "halo": hl, "grid": grid, "M": M, "N": N, "K": K,
                      "us": round(ms * 1e3, 2), "tflops": round(fl / ms / 1e9, 1)}))
