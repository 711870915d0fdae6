#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest inner loop" | tee gpurun_out/r1f_stage.txt
timeout -k 10 240 python -m pytest tests/test_inner_loop_gpu.py -m gpu -q -s --timeout 150 -p no:cacheprovider > gpurun_out/r1f_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r1f_stage.txt
grep -E "passed|failed|fused vs|graph vs|3\+3|fused vs per-step|AssertionError" gpurun_out/r1f_pytest.log | tail -12
echo "== sweep" | tee -a gpurun_out/r1f_stage.txt
timeout -k 10 400 python scripts/sweep_options.py "deep=1" "deep=1,deep_kmin=2" "tma_kmax=576" "tma_kmax=1152" "tma_out=0" "halo_mode=1" "halo_mode=2" "deep=1,tma_kmax=576" > gpurun_out/r1f_sweep.jsonl 2> gpurun_out/r1f_sweep.err
echo "sweep rc=$?" | tee -a gpurun_out/r1f_stage.txt
cat gpurun_out/r1f_sweep.jsonl; tail -3 gpurun_out/r1f_sweep.err
