"""Per-launch timing of the tensor-core kernels inside one real step (CUDA events), for a few
kernel-option settings. Usage: python scripts/step_profile.py [halo values...]"""
import sys, warnings
sys.path.insert(0, ".")
warnings.filterwarnings("ignore")
import torch
from pix2latent_b200 import native, _lib
from pix2latent_b200.loss_functions import ProjectionLoss
from pix2latent_b200.model import BigGAN
from bench import synthetic_target

halos = [int(x) for x in sys.argv[1:]] or [0, 1]  # halo_mode values (0 off, 1 resident-weight layers, 2 all 3x3); values > 100 set the TMA-I/O K limit instead
n = 18
target, weight = synthetic_target(256, "cuda")
z = torch.fmod(torch.randn(n, 128), 2.0).cuda()
for halo in halos:
    if halo > 100:
        _lib.set_option("tma_kmax", halo); halo = 0
    _lib.set_option("halo_rgb", 1 if halo == 3 else 0)
    if halo == 3: halo = 0
    _lib.set_option("halo_mode", halo)
    model = BigGAN(seed=0, allow_synthetic=True).cuda()          # plans are built lazily with the current options
    loss_fn = ProjectionLoss(allow_synthetic=True)
    tgt = loss_fn.prepared_target(target, weight)
    c = model.get_class_embedding(153).repeat(n, 1).contiguous()
    f = lambda: native.biggan_step(model.native, loss_fn.native_lpips(), tgt, z, c, True, 1 / 9, want_img=False)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print("halo=%d: %.3f ms/step" % (halo, e0.elapsed_time(e1) / 10), flush=True)
    native.profile_enable(1)
    f(); torch.cuda.synchronize()
    recs = _lib.profile_records()
    native.profile_read(); native.profile_enable(0)
    tot = sum(r[0] for r in recs)
    print("  conv launches %d, sum %.3f ms" % (len(recs), tot))
    for i, (ms, fl, BN, mode, hl, grid, M, N, K) in enumerate(recs):
        if ms > 0.04:
            print("  #%3d %s BN%3d halo%2d grid%3d M%8d N%5d K%5d  %7.1f us  %6.0f TF/s" % (i, "bwd" if mode else "fwd", BN, hl, grid, M, N, K, ms * 1e3, fl / ms / 1e9))
    del model, loss_fn, tgt
