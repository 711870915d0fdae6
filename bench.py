#!/usr/bin/env python
"""bench.py — candidate latents/sec of the inversion inner step (BASELINE.json metric).

One "step" = one pass of the hot path over the population shard of this rank: generator forward ->
ProjectionLoss (L1 + 10*LPIPS-alex) -> backward to the latents, on synthetic targets and seeded random-init weights of
the named architecture (no network for checkpoints).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|sg2_cars|sg2_ffhq] [--impl reference]

Workloads (BASELINE.json configs):
  c2        BigGAN-deep-256 BasinCMA inner step, population 18 per GPU in chunks of 9 (configs[1]; configs[3] at N = 8:
            144 candidates, 18 per rank, candidate-sharded — weak scaling). DEFAULT, the driver's line.
  sg2_cars  StyleGAN2 LSUN-cars 512x512, CMA population 22 in chunks 9/9/4, loss on rows 64:-64 (configs[2])
  sg2_ffhq  StyleGAN2 FFHQ 1024x1024, 8 candidates per GPU (configs[4]: population 64 over 8 GPUs)

N > 1 is launched by torchrun (one rank per GPU, NCCL); timing = CUDA events, barrier + synchronize on both sides, max
over ranks. At N > 1 the timed region of `c2` ends with the search loop's one data-path collective, the NCCL all_gather
of the per-candidate losses (configs[3] "NCCL loss allgather"). Rank 0 prints ONE JSON line.

Keys beyond the base contract:
  e2e          same metric through the C-ABI step with HOST (pinned) latents in and loss / gradients out, H2D + D2H
               inside the timed region, one host sync per step
  roofline     tensor-core kernel (conv_gemm, tcgen05): algorithmic FLOPs / summed launch durations
               measured with CUDA events in an instrumented pass inside this script; peak = MEASURED_PEAKS.json (burst
               figure when the clock record of the run shows no power cap, else sustained); `traffic` = DRAM bytes per
               tensor-core launch measured by an ncu pass this script spawns on itself (N = 1)
  cpu_baseline the oracle port of the reference path (torch fp32 on the host cores) on a bounded sample, rank 0, N = 1
  api          (c2) the same steps through the package's public API: closure.step per step, the device-resident loop
  search_loop  (c2) BasinCMA meta-iterations at population 18*N through BasinCMAOptimizer under the process group:
               rank-0 ask -> broadcast -> 30 fused gradient steps -> eval-only pass -> all_gather of losses -> rank-0
               tell; candidate-steps/s and where the time goes
--impl reference times the CPU path alone (the reference itself is pure Python over third-party packages that are not
installable offline — SURVEY.md F3 — so the arm runs the oracle port, kind "port").
"""
import argparse
import csv
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 9                 # max_batch_size of the reference examples -> 1/9 gradient scale
POP_PER_GPU = 18          # PyCMA default popsize for dim 128 (reference README.md:74)
FLOP_PER_UNIT = 121.1e9   # 2*(58.80 + 1.74) GFLOP per candidate-step (BigGAN-deep-256, alex)
METRIC = "candidate latents/sec (generator+LPIPS fwd+bwd)"
WORKLOADS = {
    "c2": dict(kind="biggan", res=256, pop=POP_PER_GPU, chunks=[9, 9], flop=FLOP_PER_UNIT,
               name="BigGAN-deep-256 BasinCMA inner step (generator fwd + L1+10*LPIPS-alex + bwd to z,c), population 18 per GPU, "
                    "256x256, grad scale 1/9 (BASELINE.json configs[1]; configs[3] at N=8)"),
    "sg2_cars": dict(kind="sg2", model="cars", res=512, pop=22, chunks=[9, 9, 4], flop=255.4e9, band=True,
                     name="StyleGAN2 LSUN-cars 512x512 CMA population 22 (chunks 9/9/4), loss on rows 64:-64, generator fwd + "
                          "L1+10*LPIPS-alex + bwd to z (BASELINE.json configs[2])"),
    "sg2_ffhq": dict(kind="sg2", model="ffhq", res=1024, pop=8, chunks=[8], flop=361.3e9, band=False,
                     name="StyleGAN2 FFHQ 1024x1024, 8 candidates per GPU (BASELINE.json configs[4]: population 64 over 8 GPUs), "
                          "generator fwd + L1+10*LPIPS-alex + bwd to z"),
}


def synthetic_target(res, device, band=False):
    """low-passed tanh(0.5*randn) target in (-1,1); BigGAN: 0.3 weight with a centred box of 1;
    StyleGAN2-cars: weight = loss_mask = 1 on rows res/8 : -res/8 (examples/invert_stylegan2_cars_basincma.py:39-42)."""
    g = torch.Generator().manual_seed(1)
    t = torch.tanh(0.5 * torch.randn(1, 3, res, res, generator=g))
    t = torch.nn.functional.avg_pool2d(t, 8)
    t = torch.nn.functional.interpolate(t, size=(res, res), mode="bilinear", align_corners=False)[0]
    if band:
        w = torch.zeros(3, res, res)
        w[:, res // 8:res - res // 8, :] = 1.0
    else:
        w = torch.full((3, res, res), 0.3)
        q = res // 4
        w[:, q:res - q, q:res - q] = 1.0
    return t.to(device), w.to(device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in o.strip().split(",")]
                if len(p) >= 6:
                    self.rows.append(p)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peaks(power_capped):
    """Tensor peak (TFLOP/s) for the roofline: the burst figure when the run's clock record shows no power cap (the SM
    clock stayed at its maximum), the sustained one otherwise."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if power_capped:
            return d.get("bf16_tflops_sustained", 1381.3), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured; sw_power_cap seen)"
        return d.get("bf16_tflops", 1632.3), "MEASURED_PEAKS.json bf16_tflops burst (of measured; no power cap in the clock record)"
    return (1400.0, "B200_PROFILING.md fallback 1.4 PFLOP/s sustained (of fallback)") if power_capped else \
        (1590.0, "B200_PROFILING.md fallback 1.59 PFLOP/s burst (of fallback)")


# ------------------------------------------------------------------------------- CPU reference arm
def _oracle_problem(workload):
    """(model, loss_fn, make_vars(n), oracle closure module) of the oracle port for a workload, on the CPU."""
    from oracle import closure as oc, lpips as olp
    import pix2latent_b200.distribution as dist
    import pix2latent_b200.utils.function_hooks as hook
    import torch.optim as optim
    W = WORKLOADS[workload]
    loss_fn = olp.ProjectionLoss(lpips_module=olp.make_lpips("alex", seed=0))
    target, weight = synthetic_target(W["res"], "cpu", band=W.get("band", False))
    base = dict(distribution=None, optimizer=optim.Adam, learning_rate=0.05, hook_fn=None, grad_free=False)
    spec = {}
    if W["kind"] == "biggan":
        from oracle import biggan as obg
        model = obg.make_biggan(obg.BigGANConfig.deep256(), seed=0, calibrate=False)
        spec["z"] = dict(base, shape=(128,), var_type="input", requires_grad=True, default=None,
                         distribution=dist.TruncatedNormalModulo(), hook_fn=hook.Clamp(2.0))
        spec["c"] = dict(base, shape=(128,), var_type="input", requires_grad=True, default=model.get_class_embedding(153)[0],
                         learning_rate=0.01)
    else:
        from oracle import stylegan2 as osg
        model = osg.make_stylegan2(W["res"], None, seed=0)
        for p in model.parameters():
            p.requires_grad_(True)   # the reference leaves the generator trainable (weight gradients are computed)
        spec["z"] = dict(base, shape=(512,), var_type="input", requires_grad=True, default=None,
                         distribution=dist.TruncatedNormalModulo(), hook_fn=hook.Clamp(2.0))
    R = W["res"]
    spec["target"] = dict(base, shape=(3, R, R), var_type="output", requires_grad=False, default=target)
    spec["weight"] = dict(base, shape=(3, R, R), var_type="output", requires_grad=False, default=weight)
    if W.get("band"):
        spec["loss_mask"] = dict(base, shape=(3, R, R), var_type="output", requires_grad=False, default=weight)

    def make_vars(n):
        torch.manual_seed(2)
        return oc.initialize(spec, n, "cpu")

    return model, loss_fn, make_vars, oc


def cpu_reference(workload, steps, warmup, budget_s=150.0, cand=None):
    """Oracle port of the reference path on the host cores: fp32 generator as the reference executes it (BigGAN: 128-channel
    rgb conv; weight gradients of the unfrozen generator; LPIPS on the target recomputed every step; chunk of <= 9), one
    Adam step per call. Bounded: one candidate-step costs seconds on the host, so the sample is sized to a wall-clock
    budget — candidates per step first, then the step count. Thread count: "all host cores" is often NOT the fastest
    setting for torch's CPU convolutions on a many-core box (oversubscription), so a 1-candidate step is timed at a few
    settings and the best one is used and reported as `cores`."""
    model, loss_fn, make_vars, oc = _oracle_problem(workload)
    ncpu = os.cpu_count()
    per_cand, threads = None, ncpu
    for t in sorted({min(ncpu, 16), min(ncpu, 32), min(ncpu, 64), ncpu}):   # ascending: the small settings are the cheap ones
        torch.set_num_threads(t)
        v = make_vars(1)
        if per_cand is None and WORKLOADS[workload]["kind"] == "biggan":
            oc.step(model, v, loss_fn, optimize=True, max_batch_size=CHUNK)   # warms the thread pool / allocator
        t0 = time.time()
        oc.step(model, v, loss_fn, optimize=True, max_batch_size=CHUNK)
        dt1 = time.time() - t0
        if per_cand is None or dt1 < per_cand:
            per_cand, threads = dt1, t
        if dt1 > 1.2 * per_cand or dt1 > 0.15 * budget_s:
            break   # getting worse (oversubscription), or too slow to keep probing
    torch.set_num_threads(threads)
    if cand is None:
        cand = int(max(1, min(CHUNK, budget_s / max(1e-6, (steps + warmup) * per_cand))))
    warmup_run = max(0, min(warmup, int(0.2 * budget_s / (cand * per_cand))))
    steps_run = max(1, min(steps, int(0.8 * budget_s / (cand * per_cand))))
    v = make_vars(cand)
    for _ in range(warmup_run):
        oc.step(model, v, loss_fn, optimize=True, max_batch_size=CHUNK)
    t0 = time.time()
    for _ in range(steps_run):
        oc.step(model, v, loss_fn, optimize=True, max_batch_size=CHUNK)
    dt = time.time() - t0
    return {"value": cand * steps_run / dt, "unit": "candidates/s", "cores": threads, "kind": "port",
            "sample": "%d candidates x %d optimise-steps (+%d warm-up after a 1-candidate sizing probe) of the %s step, oracle port, torch "
                      "fp32, %d threads (best of a probe over thread counts; %d cores on the box)"
                      % (cand, steps_run, warmup_run, workload, threads, ncpu),
            "ms_per_step": 1e3 * dt / steps_run, "candidates": cand, "steps_run": steps_run, "warmup_run": warmup_run}


def run_reference(args, rank, world):
    if rank != 0:
        return
    W = WORKLOADS[args.workload]
    r = cpu_reference(args.workload, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "candidates/s", "n_gpus": args.gpus,
        "steps": r["steps_run"], "warmup": r["warmup_run"], "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s — CPU sample of %d candidates (chunk<=9)" % (W["name"], r["candidates"]),
                   "resolution": W["res"], "lpips_net": "alex"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- native workloads
class BigGANWork:
    """configs[1] / [3]: 18 candidates per GPU, one fused C-ABI call per step (per-sample 1/9 gradient scales)."""

    def __init__(self, W, dev, rank):
        from pix2latent_b200 import native
        from pix2latent_b200.loss_functions import ProjectionLoss
        from pix2latent_b200.model import BigGAN
        self.native, self.W, self.dev, self.n = native, W, dev, W["pop"]
        self.model = BigGAN(seed=0, allow_synthetic=True).cuda()   # no network here: seeded random-init weights of the named architecture
        self.loss_fn = ProjectionLoss(allow_synthetic=True)
        self.target, self.weight = synthetic_target(256, dev)
        self.tgt = self.loss_fn.prepared_target(self.target, self.weight)
        self.gen, self.lp = self.model.native, self.loss_fn.native_lpips()
        g = torch.Generator().manual_seed(2 + rank)
        self.z = torch.fmod(torch.randn(self.n, 128, generator=g), 2.0).to(dev)
        self.c = self.model.get_class_embedding(153).repeat(self.n, 1).contiguous()
        self.scale = 1.0 / CHUNK
        self.hz, self.hc = self.z.cpu().pin_memory(), self.c.cpu().pin_memory()
        self.hl = torch.empty(self.n).pin_memory()
        self.hdz, self.hdc = torch.empty(self.n, 128).pin_memory(), torch.empty(self.n, 128).pin_memory()
        self.dz_d, self.dc_d = torch.empty(self.n, 128, device=dev), torch.empty(self.n, 128, device=dev)
        self.h2d = 2 * self.n * 128 * 4
        self.d2h = self.n * 4 + 2 * self.n * 128 * 4

    def step_dev(self, grad=True):
        return self.native.biggan_step(self.gen, self.lp, self.tgt, self.z, self.c, grad, self.scale, want_img=False)[0]

    def step_e2e(self):
        self.dz_d.copy_(self.hz, non_blocking=True)  # staging reused as device z/c
        self.dc_d.copy_(self.hc, non_blocking=True)
        l, gz, gc, _ = self.native.biggan_step(self.gen, self.lp, self.tgt, self.dz_d, self.dc_d, True, self.scale, want_img=False)
        self.hl.copy_(l, non_blocking=True)
        self.hdz.copy_(gz, non_blocking=True)
        self.hdc.copy_(gc, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the host consumes loss / gradients every step


class SG2Work:
    """configs[2] / [4]: the population with the reference's chunking (closure.py:32: split_vars by max_batch_size = 9) as the
    product runs it (closure._step_native_sg2): fresh per-layer noise drawn chunk by chunk (rosinality draws N(0,1) noise per
    layer per forward), per-candidate gradient scales 1/b_chunk, and ONE fused C-ABI call per physical batch of up to 24
    candidates — a candidate's result does not depend on the batch it is evaluated in (tests/test_determinism_gpu.py)."""

    def __init__(self, W, dev, rank):
        from pix2latent_b200 import native
        from pix2latent_b200.loss_functions import ProjectionLoss
        from pix2latent_b200.model.stylegan2 import StyleGAN2
        from pix2latent_b200.optimizer.closure import SG2_PHYS_BATCH
        self.native, self.W, self.dev, self.n = native, W, dev, W["pop"]
        self.model = StyleGAN2(W["model"], allow_synthetic=True)
        self.loss_fn = ProjectionLoss(allow_synthetic=True)
        self.target, self.weight = synthetic_target(W["res"], dev, band=W["band"])
        self.tgt = self.loss_fn.prepared_target(self.target, self.weight, self.weight if W["band"] else None)
        self.gen, self.lp = self.model.native, self.loss_fn.native_lpips()
        g = torch.Generator().manual_seed(2 + rank)
        self.z = torch.fmod(torch.randn(self.n, 512, generator=g), 2.0).to(dev)
        self.dloss = torch.cat([torch.full((c,), 1.0 / c) for c in W["chunks"]]).to(dev)
        self.phys = [(lo, min(self.n, lo + SG2_PHYS_BATCH)) for lo in range(0, self.n, SG2_PHYS_BATCH)]
        self.hz = self.z.cpu().pin_memory()
        self.hl, self.hdz = torch.empty(self.n).pin_memory(), torch.empty(self.n, 512).pin_memory()
        self.z_d = torch.empty(self.n, 512, device=dev)
        self.h2d = self.n * 512 * 4
        self.d2h = self.n * 4 + self.n * 512 * 4

    def _batches(self, z, grad):
        parts = [self.model.draw_noise(c, self.dev) for c in self.W["chunks"]]
        noise = [torch.cat([p[l] for p in parts]) if len(parts) > 1 else parts[0][l] for l in range(len(parts[0]))]
        return [self.native.sg2_step(self.gen, self.lp, self.tgt, z[lo:hi], [t[lo:hi] for t in noise], grad, 1.0,
                                     want_img=False, dloss=self.dloss[lo:hi]) for lo, hi in self.phys]

    def step_dev(self, grad=True):
        out = self._batches(self.z, grad)
        return out[0][0] if len(out) == 1 else torch.cat([o[0] for o in out])

    def step_e2e(self):
        self.z_d.copy_(self.hz, non_blocking=True)
        out = self._batches(self.z_d, True)
        self.hl.copy_(out[0][0] if len(out) == 1 else torch.cat([o[0] for o in out]), non_blocking=True)
        self.hdz.copy_(out[0][1] if len(out) == 1 else torch.cat([o[1] for o in out]), non_blocking=True)
        torch.cuda.current_stream().synchronize()


def measure_traffic(args, launches_per_step):
    """DRAM bytes per tensor-core launch, measured now: an ncu pass over one step of this script (`--ncu`), metrics
    dram__bytes_read.sum + dram__bytes_write.sum of the conv_gemm / halo launches. None when ncu is not usable."""
    ncu = "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu) or launches_per_step <= 0:
        return None, "ncu not found"
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    log = os.path.join(out, "bench_traffic_%s.csv" % args.workload)
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:conv_gemm|conv3x3",
           "--csv", "--log-file", log, sys.executable, os.path.abspath(__file__), "--ncu", "--workload", args.workload, "--steps", "1",
           "--warmup", "1"]
    rc = "?"
    try:
        rc = subprocess.run(cmd, capture_output=True, text=True, timeout=420).returncode
        rows = list(csv.DictReader(l for l in open(log) if not l.startswith("==")))
    except Exception as e:
        return None, "ncu pass failed: %s" % e
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r_ in rows:
        if r_.get("Metric Name", "").startswith("dram__bytes"):
            per.setdefault(r_["ID"], 0.0)
            per[r_["ID"]] += float(r_["Metric Value"]) * unit.get(r_["Metric Unit"], 1)
    vals = [per[k] for k in sorted(per, key=lambda s: int(s))]
    if len(vals) < launches_per_step:
        return None, "ncu pass saw %d tensor-core launches, expected >= %d (rc %s)" % (len(vals), launches_per_step, rc)
    last = vals[-int(launches_per_step):]   # the last (timed) step's launches
    return sum(last) / len(last), "ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d tensor-core launches of one step, this run" % len(last)


def api_legs(work, args, world, sync_all):
    """(c2) the same work through the package's public API (what examples/invert_*.py call)."""
    from pix2latent_b200 import VariableManager, native
    from pix2latent_b200.optimizer.closure import step as api_step
    import pix2latent_b200.utils.function_hooks as hook
    n, dev = work.n, work.dev
    out = {"unit": "candidates/s"}
    cfg_adam = native.adam_config(0.05, 0.01, clamp_z=2.0)
    dl = torch.full((n,), work.scale, device=dev)
    k_inner = max(4, min(args.steps, 50))
    out["steps_per_call"] = k_inner
    out["what"] = ("fused_loop: Clamp hook + generator fwd + loss + bwd + Adam(z lr 0.05, c lr 0.01) per step inside ONE C-ABI call "
                   "(p2l_biggan_optimize, CUDA-graph replay), losses read back once per K steps; per_step: closure.step per step "
                   "(Python hooks, one C-ABI step, device-resident Adam, losses read lazily)")
    for name, use_graph in (("fused_loop_eager", False), ("fused_loop", True)):
        zz, cc = work.z.clone(), work.c.clone()
        native.biggan_optimize(work.gen, work.lp, work.tgt, zz, cc, 4, cfg_adam, dloss=dl, want_img=False, use_graph=use_graph)
        sync_all()
        t0 = time.perf_counter()
        r = native.biggan_optimize(work.gen, work.lp, work.tgt, zz, cc, k_inner, cfg_adam, dloss=dl, want_img=False, use_graph=use_graph)
        hist = r["loss"].cpu()
        sync_all()
        out[name] = world * n * k_inner / (time.perf_counter() - t0)
        if use_graph:
            out["graph_used"] = bool(r["graph"])
            out["first_loss_mean"], out["final_loss_mean"] = float(hist[0].mean()), float(hist[-1].mean())
    vm = VariableManager(device=dev)
    vm.register("z", (128,), "input", learning_rate=0.05, hook_fn=hook.Clamp(2.0))
    vm.register("c", (128,), "input", default=work.model.get_class_embedding(153)[0], learning_rate=0.01)
    vm.register("target", (3, 256, 256), "output", requires_grad=False, default=work.target)
    vm.register("weight", (3, 256, 256), "output", requires_grad=False, default=work.weight)
    variables = vm.initialize(n)
    for _ in range(3):
        api_step(work.model, variables, work.loss_fn, optimize=True, max_batch_size=CHUNK)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(k_inner):
        _, losses, _ = api_step(work.model, variables, work.loss_fn, optimize=True, max_batch_size=CHUNK)
    float(losses[0])   # the caller reads the last step's losses
    sync_all()
    out["per_step"] = world * n * k_inner / (time.perf_counter() - t0)
    return out


def search_loop_leg(work, world, rank, sync_all, meta_steps=2, grad_steps=30):
    """BASELINE configs[3] as the reference runs it (/root/reference pix2latent/optimizer/basincma_optimizer.py:24-83,
    base_cma_optimizer.py:71-141), population 18 per GPU, through the package's BasinCMAOptimizer under the process group:
    per meta-iteration rank 0 asks, ONE broadcast of the asked z, `grad_steps` fused gradient steps on every rank's shard,
    an eval-only pass, ONE all_gather of the per-candidate losses, rank 0 tells."""
    from pix2latent_b200 import VariableManager, parallel
    from pix2latent_b200.optimizer import BasinCMAOptimizer
    import pix2latent_b200.distribution as pdist
    import pix2latent_b200.utils.function_hooks as hook
    dev, pop = work.dev, work.n * world
    vm = VariableManager(device=dev)
    vm.register("z", (128,), "input", grad_free=True, distribution=pdist.TruncatedNormalModulo(sigma=1.0, trunc=2.0),
                learning_rate=0.05, hook_fn=hook.Clamp(2.0))
    vm.register("c", (128,), "input", default=work.model.get_class_embedding(153)[0], learning_rate=0.01)
    vm.register("target", (3, 256, 256), "output", requires_grad=False, default=work.target)
    vm.register("weight", (3, 256, 256), "output", requires_grad=False, default=work.weight)
    opt = BasinCMAOptimizer(work.model, vm, work.loss_fn, max_batch_size=CHUNK, track_variables=False)
    opt.cma_seed, opt.cma_popsize, opt.show_iter = 7, pop, 10 ** 9
    phases = {"ask_broadcast": 0.0, "grad_steps": 0.0, "eval_gather_tell": 0.0, "collectives": 0.0}

    def timed(key, fn):
        def wrapper(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize()
            phases[key] += time.perf_counter() - t0
            return r
        return wrapper

    opt.cma_init = timed("ask_broadcast", opt.cma_init)
    opt.grad_steps = timed("grad_steps", opt.grad_steps)
    opt.cma_update = timed("eval_gather_tell", opt.cma_update)
    old = parallel.allgather_losses, parallel.broadcast_array
    parallel.allgather_losses = timed("collectives", old[0])
    parallel.broadcast_array = timed("collectives", old[1])
    try:
        from pix2latent_b200.utils.misc import HiddenPrints
        with HiddenPrints():
            opt.optimize(meta_steps=1, grad_steps=4, last_grad_steps=2)   # warm-up: plans, graph capture, NCCL channels
            for k in phases:
                phases[k] = 0.0
            sync_all()
            t0 = time.perf_counter()
            _, _, losses = opt.optimize(meta_steps=meta_steps, grad_steps=grad_steps, last_grad_steps=grad_steps)
            sync_all()
            dt = time.perf_counter() - t0
    finally:
        parallel.allgather_losses, parallel.broadcast_array = old
    t = torch.tensor([dt], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    n_grad = (meta_steps + 1) * grad_steps
    return {"value": pop * n_grad / dt, "unit": "candidate optimise-steps/s", "population": pop, "meta_iterations": meta_steps + 1,
            "grad_steps_per_meta_iteration": grad_steps, "seconds": dt,
            "rank0_seconds_by_phase": {k: round(v, 4) for k, v in phases.items()},
            "fused_calls": opt.fused_calls, "backend": "nccl" if world > 1 else "single process",
            "best_final_loss": float(min(losses[0][1]["loss"])),
            "what": "BasinCMAOptimizer.optimize(meta_steps=%d, grad_steps=%d, last_grad_steps=%d): ask -> broadcast -> fused Adam steps -> "
                    "eval-only pass -> all_gather(losses) -> tell; value counts the optimise-steps only" % (meta_steps, grad_steps, grad_steps)}


def run_native(args, rank, world, local_rank):
    import torch.distributed as dist
    from pix2latent_b200 import native
    import warnings
    warnings.filterwarnings("ignore")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = WORKLOADS[args.workload]
    work = BigGANWork(W, dev, rank) if W["kind"] == "biggan" else SG2Work(W, dev, rank)
    n = work.n

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    warm = args.warmup if args.ncu else max(args.warmup, 3)
    for _ in range(warm):
        work.step_dev()
    gathered = torch.empty(world * n, device=dev) if world > 1 else None
    if world > 1:
        dist.all_gather_into_tensor(gathered, work.step_dev().contiguous())   # NCCL channels up before the timed region
    sync_all()
    # ---- timed region: inputs resident in HBM. The step's working set (GBs of saved activations) exceeds the 126 MB
    # L2 many times over, so no explicit flush. N > 1: ends with the search loop's one data-path collective.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = native.launch_count()
    with ClockSampler(local_rank) as clocks:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            loss = work.step_dev()
        if world > 1:
            dist.all_gather_into_tensor(gathered, loss.contiguous())
        e1.record()
        sync_all()
    ms = e0.elapsed_time(e1)
    launches = native.launch_count() - launches0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * args.steps / (ms / 1e3)

    if args.ncu:
        if rank == 0:
            print(json.dumps({"ncu_mode": True, "ms_per_step": ms / args.steps, "gpu_launches": launches}))
        return
    # ---- end to end through the C-ABI step with host buffers
    for _ in range(3):
        work.step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        work.step_e2e()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / (float(t.item()) / 1e3)

    # ---- roofline leg: per-launch CUDA-event timing of the tensor-core kernel (instrumented pass)
    native.profile_enable(1)
    for _ in range(args.steps):
        work.step_dev()
    torch.cuda.synchronize()
    conv_ms, conv_n, lib_flops = native.profile_read()
    native.profile_enable(0)
    clk = clocks.summary()
    peak, peak_src = measured_peaks("sw_power_cap" in clk["reasons"])
    alg_flops = W["flop"] * n * args.steps
    achieved = alg_flops / (conv_ms / 1e3) / 1e12
    roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": None, "kernel": "conv_gemm_kernel / conv3x3_halo_kernel (tcgen05.mma kind::f16, 16-bit operands, fp32 accumulate)",
            "launches_per_step": conv_n / args.steps, "kernel_ms_per_step": conv_ms / args.steps,
            "kernel_share_of_step": (conv_ms / args.steps) / (ms / args.steps),
            "flops_per_launch_algorithmic": alg_flops / conv_n, "avg_launch_ms": conv_ms / conv_n,
            "executed_flops_per_step": lib_flops / args.steps, "peak_source": peak_src,
            "whole_step_tflops": W["flop"] * n / (ms / args.steps / 1e3) / 1e12}
    if world == 1 and not args.no_traffic:
        tr, src = measure_traffic(args, conv_n // args.steps)
        if tr is None:
            p = os.path.join(ROOT, "profiles", "traffic.json")
            if args.workload == "c2" and os.path.exists(p):
                tr = json.load(open(p)).get("dram_bytes_per_launch")
                src = "profiles/traffic.json (earlier ncu capture; this run: %s)" % src
        roof["traffic"], roof["traffic_source"] = tr, src

    # ---- eval-only units (generator fwd + loss, no backward): what CMAOptimizer's meta-iterations run
    # (cma_optimizer.py:46-72; SURVEY.md section 8d asks for them separately)
    for _ in range(3):
        work.step_dev(False)
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        work.step_dev(False)
    f1.record()
    sync_all()
    t = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    eval_only = {"value": world * n * args.steps / (float(t.item()) / 1e3), "unit": "candidates/s",
                 "what": "generator forward + L1+10*LPIPS loss only (no backward), inputs resident in HBM"}

    api, search = None, None
    if W["kind"] == "biggan":
        try:
            api = api_legs(work, args, world, sync_all)
        except Exception as e:  # the headline numbers above do not depend on these legs
            api = {"error": "%s: %s" % (type(e).__name__, e)}
        try:
            search = search_loop_leg(work, world, rank, sync_all)
        except Exception as e:
            search = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16" if native.act_dtype() == torch.float16 else "bf16", "data": "synthetic",
        "config": {"workload": W["name"], "population_per_gpu": n, "global_population": n * world, "chunks": W["chunks"],
                   "physical_batches": ([n] if W["kind"] == "biggan" else [hi - lo for lo, hi in work.phys]),
                   "chunking": "the reference's chunks set the per-candidate 1/b_chunk gradient scales (and the RNG draw order); "
                               "the candidates run in the physical batches listed — results are bitwise independent of the batching",
                   "resolution": W["res"], "lpips_net": "alex",
                   "parallelism": ("candidate-sharded x%d; the timed region ends with one NCCL all_gather of the per-candidate losses" % world)
                   if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (GBs of saved activations per step vs 126 MB)"},
        "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": work.h2d, "d2h_bytes_per_step": work.d2h},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roof,
        "final_loss_mean": float(loss.mean().item()),
        "eval_only": eval_only,
    }
    if api is not None:
        line["api"] = api
    if search is not None:
        line["search_loop"] = search
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(args.workload, 1, 0, budget_s=30.0, cand=3 if W["kind"] == "biggan" else 1)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu pass that measures DRAM bytes per launch")
    ap.add_argument("--ncu", action="store_true", help="launch-list mode for ncu: W warm-ups + K steps only, no e2e / profile / CPU legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
