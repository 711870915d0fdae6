#!/usr/bin/env python
"""bench.py — candidate latents/sec of the inversion inner step (BASELINE.json metric).

One "step" = one pass of the hot path over the population shard of this rank:
generator forward (BigGAN-deep-256) -> ProjectionLoss (L1 + 10*LPIPS-alex) -> backward to
dL/dz and dL/dc, for 18 candidates per GPU (BASELINE.json configs[1]; configs[3] at N=8:
144 candidates, 18 per rank, embarrassingly sharded — weak scaling). Synthetic target and
seeded random-init weights of the named architecture (no network for checkpoints).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL); timing = CUDA events, barrier +
synchronize on both sides, max over ranks. Rank 0 prints ONE JSON line.

Keys beyond the base contract:
  e2e          same metric through the C-ABI step with HOST (pinned) z/c in and loss/dz/dc out,
               H2D + D2H inside the timed region, one host sync per step
  roofline     tensor-core kernel (conv_gemm, tcgen05): algorithmic FLOPs (SURVEY.md §8d:
               121.1 GFLOP per candidate-step) / summed launch durations measured with CUDA
               events in an instrumented pass inside this script; peak = MEASURED_PEAKS.json
  cpu_baseline the oracle port of the reference path (torch fp32 on the host cores) on a bounded
               sample, rank 0 at N=1 only
--impl reference times that CPU path alone (the reference itself is pure Python over
third-party packages that are not installable offline — SURVEY.md F3 — so the arm runs the
oracle port, kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POP_PER_GPU = 18          # PyCMA default popsize for dim 128 (reference README.md:74)
CHUNK = 9                 # max_batch_size of the reference examples -> 1/9 gradient scale
FLOP_PER_UNIT = 121.1e9   # 2*(58.80 + 1.74) GFLOP per candidate-step (alex)
METRIC = "candidate latents/sec (generator+LPIPS fwd+bwd)"


def synthetic_target(res, device):
    """low-passed tanh(0.5*randn) target in (-1,1); 0.3 weight with a centred box of 1."""
    g = torch.Generator().manual_seed(1)
    t = torch.tanh(0.5 * torch.randn(1, 3, res, res, generator=g))
    t = torch.nn.functional.avg_pool2d(t, 8)
    t = torch.nn.functional.interpolate(t, size=(res, res), mode="bilinear", align_corners=False)[0]
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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1472.0), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
    return 1400.0, "B200_PROFILING.md fallback 1.4 PFLOP/s sustained (of fallback)"


# ------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(steps, warmup, budget_s=150.0, cand=None):
    """Oracle port of the reference path on the host cores: BigGAN-deep-256 fp32 generator as the
    reference executes it (128-channel rgb conv, weight gradients of the unfrozen generator, LPIPS
    on the target recomputed every step, chunk of <= 9), one Adam step per call."""
    from oracle import biggan as obg, closure as oc, lpips as olp
    import pix2latent_b200.distribution as dist
    import pix2latent_b200.utils.function_hooks as hook
    import torch.optim as optim
    model = obg.make_biggan(obg.BigGANConfig.deep256(), seed=0, calibrate=False)
    loss_fn = olp.ProjectionLoss(lpips_module=olp.make_lpips("alex", seed=0))
    target, weight = synthetic_target(256, "cpu")

    def make_vars(n):
        torch.manual_seed(2)
        spec = {
            "z": dict(shape=(128,), var_type="input", requires_grad=True, default=None, distribution=dist.TruncatedNormalModulo(),
                      optimizer=optim.Adam, learning_rate=0.05, hook_fn=hook.Clamp(2.0), grad_free=False),
            "c": dict(shape=(128,), var_type="input", requires_grad=True, default=model.get_class_embedding(153)[0],
                      distribution=None, optimizer=optim.Adam, learning_rate=0.01, hook_fn=None, grad_free=False),
            "target": dict(shape=(3, 256, 256), var_type="output", requires_grad=False, default=target, distribution=None,
                           optimizer=optim.Adam, learning_rate=0.05, hook_fn=None, grad_free=False),
            "weight": dict(shape=(3, 256, 256), var_type="output", requires_grad=False, default=weight, distribution=None,
                           optimizer=optim.Adam, learning_rate=0.05, hook_fn=None, grad_free=False),
        }
        return oc.initialize(spec, n, "cpu")

    # Bounded: one candidate-step costs seconds on the host, so the sample is sized to a wall-clock
    # budget — candidates per step first, then (if even 1 candidate x K steps is too long) the step count.
    # thread count: "all host cores" is often NOT the fastest setting for torch's CPU convolutions on
    # a many-core box (oversubscription), so a 1-candidate step is timed at a few settings and the best
    # one is used and reported as `cores`
    ncpu = os.cpu_count()
    per_cand, threads = None, ncpu
    for t in sorted({min(ncpu, 16), min(ncpu, 32), min(ncpu, 64), ncpu}):   # ascending: the small settings are the cheap ones
        torch.set_num_threads(t)
        v = make_vars(1)
        if per_cand is None:
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
            "sample": "%d candidates x %d optimise-steps (+%d warm-up after one 1-candidate sizing step) of the "
                      "BigGAN-deep-256 / alex-LPIPS step, torch fp32, %d threads (best of a probe over thread counts; %d cores on the box)"
                      % (cand, steps_run, warmup_run, threads, ncpu),
            "ms_per_step": 1e3 * dt / steps_run, "candidates": cand, "steps_run": steps_run, "warmup_run": warmup_run}


def run_reference(args, rank, world):
    if rank != 0:
        return
    r = cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "candidates/s", "n_gpus": args.gpus,
        "steps": r["steps_run"], "warmup": r["warmup_run"], "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BigGAN-deep-256 BasinCMA inner step, CPU sample of %d candidates (chunk<=9)" % r["candidates"],
                   "resolution": 256, "lpips_net": "alex"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- native arm
def run_native(args, rank, world, local_rank):
    import torch.distributed as dist
    from pix2latent_b200 import native
    from pix2latent_b200.loss_functions import ProjectionLoss
    from pix2latent_b200.model import BigGAN
    import warnings
    warnings.filterwarnings("ignore")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = POP_PER_GPU
    model = BigGAN(seed=0, allow_synthetic=True).cuda()   # no network here: seeded random-init weights of the named architecture
    loss_fn = ProjectionLoss(allow_synthetic=True)
    target, weight = synthetic_target(256, dev)
    tgt = loss_fn.prepared_target(target, weight)
    gen, lp = model.native, loss_fn.native_lpips()
    g = torch.Generator().manual_seed(2 + rank)
    z = torch.fmod(torch.randn(n, 128, generator=g), 2.0).to(dev)
    c = model.get_class_embedding(153).repeat(n, 1).contiguous()
    scale = 1.0 / CHUNK

    def step_dev():
        return native.biggan_step(gen, lp, tgt, z, c, True, scale, want_img=False)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup if args.ncu else max(args.warmup, 3)):
        step_dev()
    sync_all()
    # ---- timed region: inputs resident in HBM. The step's working set (~2.9 GB of saved
    # activations at 18 candidates) exceeds the 126 MB L2 many times over, so no explicit flush.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = native.launch_count()
    with ClockSampler(local_rank) as clocks:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            loss, dz, dc, _ = step_dev()
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
    hz, hc = z.cpu().pin_memory(), c.cpu().pin_memory()
    hl = torch.empty(n).pin_memory()
    hdz, hdc = torch.empty(n, 128).pin_memory(), torch.empty(n, 128).pin_memory()
    dz_d, dc_d = torch.empty(n, 128, device=dev), torch.empty(n, 128, device=dev)

    def step_e2e():
        dz_d.copy_(hz, non_blocking=True)  # staging reused as device z/c
        dc_d.copy_(hc, non_blocking=True)
        l, gz, gc, _ = native.biggan_step(gen, lp, tgt, dz_d, dc_d, True, scale, want_img=False)
        hl.copy_(l, non_blocking=True)
        hdz.copy_(gz, non_blocking=True)
        hdc.copy_(gc, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the host consumes loss / gradients every step

    for _ in range(3):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / (float(t.item()) / 1e3)
    h2d = 2 * n * 128 * 4
    d2h = n * 4 + 2 * n * 128 * 4

    # ---- roofline leg: per-launch CUDA-event timing of the tensor-core kernel (instrumented pass)
    native.profile_enable(1)
    for _ in range(args.steps):
        step_dev()
    torch.cuda.synchronize()
    conv_ms, conv_n, lib_flops = native.profile_read()
    native.profile_enable(0)
    peak, peak_src = measured_peaks()
    alg_flops = FLOP_PER_UNIT * n * args.steps
    achieved = alg_flops / (conv_ms / 1e3) / 1e12
    roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": None, "kernel": "conv_gemm_kernel (tcgen05.mma kind::f16, 16-bit operands, fp32 accumulate)",
            "launches_per_step": conv_n / args.steps, "kernel_ms_per_step": conv_ms / args.steps,
            "kernel_share_of_step": (conv_ms / args.steps) / (ms / args.steps),
            "flops_per_launch_algorithmic": alg_flops / conv_n, "avg_launch_ms": conv_ms / conv_n,
            "executed_flops_per_step": lib_flops / args.steps, "peak_source": peak_src,
            "whole_step_tflops": FLOP_PER_UNIT * n / (ms / args.steps / 1e3) / 1e12}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- device-resident inner loop (informational): the same step INCLUDING the Clamp hook and the per-candidate
    # Adam update, K steps per C-ABI call (p2l_biggan_optimize), losses read back once per call
    inner = None
    try:
        cfg_adam = native.adam_config(0.05, 0.01, clamp_z=2.0)
        dl = torch.full((n,), scale, device=dev)
        k_inner = max(4, min(args.steps, 50))
        inner = {"unit": "candidates/s", "steps_per_call": k_inner,
                 "what": "Clamp hook + generator fwd + loss + bwd + Adam(z lr 0.05, c lr 0.01) per step, one C-ABI call and "
                         "one D2H of the [K, n] losses per K steps"}
        for name, use_graph in (("eager", False), ("graph", True)):
            zz, cc = z.clone(), c.clone()
            native.biggan_optimize(gen, lp, tgt, zz, cc, 4, cfg_adam, dloss=dl, want_img=False, use_graph=use_graph)
            sync_all()
            t0 = time.perf_counter()
            r = native.biggan_optimize(gen, lp, tgt, zz, cc, k_inner, cfg_adam, dloss=dl, want_img=False, use_graph=use_graph)
            hist = r["loss"].cpu()
            sync_all()
            dt = time.perf_counter() - t0
            inner[name] = world * n * k_inner / dt
            if use_graph:
                inner["graph_used"] = bool(r["graph"])
                inner["final_loss_mean"] = float(hist[-1].mean())
                inner["first_loss_mean"] = float(hist[0].mean())
        # the same K steps through the product's per-step public API (closure.step: hooks in Python, one C-ABI
        # step, torch.optim.Adam over 2n per-candidate param groups, one D2H of the losses per step)
        from pix2latent_b200 import VariableManager
        from pix2latent_b200.optimizer.closure import step as api_step
        import pix2latent_b200.utils.function_hooks as hook
        vm = VariableManager(device=dev)
        vm.register("z", (128,), "input", learning_rate=0.05, hook_fn=hook.Clamp(2.0))
        vm.register("c", (128,), "input", default=model.get_class_embedding(153)[0], learning_rate=0.01)
        vm.register("target", (3, 256, 256), "output", requires_grad=False, default=target)
        vm.register("weight", (3, 256, 256), "output", requires_grad=False, default=weight)
        variables = vm.initialize(n)
        for _ in range(3):
            api_step(model, variables, loss_fn, optimize=True, max_batch_size=CHUNK)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(k_inner):
            api_step(model, variables, loss_fn, optimize=True, max_batch_size=CHUNK)
        sync_all()
        inner["api_per_step"] = world * n * k_inner / (time.perf_counter() - t0)
    except Exception as e:  # the headline numbers above do not depend on this leg
        inner = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- eval-only units (generator fwd + loss, no backward): what CMAOptimizer's meta-iterations run
    # (cma_optimizer.py:46-72; SURVEY.md section 8d asks for them separately). Informational.
    eval_only = None
    try:
        for _ in range(3):
            native.biggan_step(gen, lp, tgt, z, c, False, scale, want_img=False)
        sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            native.biggan_step(gen, lp, tgt, z, c, False, scale, want_img=False)
        f1.record()
        sync_all()
        t = torch.tensor([f0.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        eval_only = {"value": world * n * args.steps / (float(t.item()) / 1e3), "unit": "candidates/s",
                     "what": "generator forward + L1+10*LPIPS loss only (no backward), inputs resident in HBM"}
    except Exception as e:
        eval_only = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16" if native.act_dtype() == torch.float16 else "bf16", "data": "synthetic",
        "config": {"workload": "BigGAN-deep-256 BasinCMA inner step (generator fwd + L1+10*LPIPS-alex + bwd to z,c), "
                               "population 18 per GPU, 256x256, grad scale 1/9 (BASELINE.json configs[1]; configs[3] at N=8)",
                   "population_per_gpu": n, "global_population": n * world, "resolution": 256, "lpips_net": "alex",
                   "parallelism": "candidate-sharded x%d, no data-path collective" % world,
                   "l2": "inputs larger than L2 (2.9 GB of activations per step vs 126 MB)"},
        "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roof,
        "final_loss_mean": float(loss.mean().item()),
        "inner_loop": inner,
        "eval_only": eval_only,
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(1, 0, budget_s=30.0, cand=3)
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
