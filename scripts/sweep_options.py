"""Time the bench step (BigGAN-deep-256, 18 candidates) under different kernel-selection options.
    python scripts/sweep_options.py "deep=1" "deep=1,deep_kmin=4" "tma_kmax=576" ...
Each config builds a fresh generator / LPIPS (plans are built under the options in force), checks the losses
against the default config (same inputs) and reports ms/step from CUDA events. Results: one JSON line per config."""
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from bench import synthetic_target, POP_PER_GPU, CHUNK  # noqa: E402
from pix2latent_b200 import _lib, native  # noqa: E402
from pix2latent_b200.loss_functions import ProjectionLoss  # noqa: E402
from pix2latent_b200.model import BigGAN, synth  # noqa: E402

# the library's defaults at import time are the baseline every config starts from
OPTION_KEYS = ["attn_fused", "attn_emit_t", "serpentine", "pdl", "deep", "deep_kmin", "tma_out", "tma_kmax", "halo_mode", "halo"]
DEFAULTS = {}


def run(cfg, sd, lp_sd, steps=20):
    if not DEFAULTS:
        DEFAULTS.update({k: _lib.get_option(k) for k in OPTION_KEYS})
    for k, v in DEFAULTS.items():
        _lib.set_option(k, v)
    for k, v in cfg.items():
        _lib.set_option(k, v)
    dev = torch.device("cuda", 0)
    model = BigGAN(state_dict=sd).cuda()
    loss_fn = ProjectionLoss(lpips_state_dict=dict(lp_sd))  # a new dict object: a new NativeLPIPS with fresh plans
    target, weight = synthetic_target(256, dev)
    tgt = loss_fn.prepared_target(target, weight)
    gen, lp = model.native, loss_fn.native_lpips()
    n = POP_PER_GPU
    g = torch.Generator().manual_seed(2)
    z = torch.fmod(torch.randn(n, 128, generator=g), 2.0).to(dev)
    c = model.get_class_embedding(153).repeat(n, 1).contiguous()
    for _ in range(3):
        loss, dz, dc, _ = native.biggan_step(gen, lp, tgt, z, c, True, 1.0 / CHUNK, want_img=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, dz, dc, _ = native.biggan_step(gen, lp, tgt, z, c, True, 1.0 / CHUNK, want_img=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"cfg": cfg, "ms_per_step": ms, "cand_per_s": n / ms * 1e3, "loss_mean": float(loss.mean()),
           "dz_norm": float(dz.norm()), "loss": loss.cpu(), "dz": dz.cpu()}
    del model, loss_fn, gen, lp, tgt
    torch.cuda.empty_cache()
    return out


def main():
    cfgs = [{}]
    for a in sys.argv[1:]:
        cfgs.append({kv.split("=")[0]: int(kv.split("=")[1]) for kv in a.split(",") if kv})
    sd = synth.biggan_state_dict(synth.BigGANConfig(), 0)
    lp_sd = synth.lpips_state_dict("alex", 0)
    base = None
    for cfg in cfgs:
        try:
            r = run(cfg, sd, lp_sd)
        except Exception as e:
            print(json.dumps({"cfg": cfg, "error": "%s: %s" % (type(e).__name__, e)}), flush=True)
            continue
        if base is None:
            base = dict(r)
        r["loss_maxdiff_vs_default"] = float((r["loss"] - base["loss"]).abs().max())
        r["dz_cos_vs_default"] = float(torch.nn.functional.cosine_similarity(r["dz"].flatten(), base["dz"].flatten(), dim=0))
        r["speedup_vs_default"] = base["ms_per_step"] / r["ms_per_step"]
        r.pop("loss"); r.pop("dz")
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
