#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== conv parity + biggan + step tests"
timeout -k 10 300 python -m pytest tests/test_conv_gemm_gpu.py tests/test_biggan_gpu.py tests/test_step_gpu.py -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/r1j_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r1j_pytest.log
echo "== sweep"
timeout -k 10 300 python scripts/sweep_options.py "halo_mode=1" "halo_mode=2" "tma_kmax=576" > gpurun_out/r1j_sweep.jsonl 2> gpurun_out/r1j_sweep.err
cat gpurun_out/r1j_sweep.jsonl; tail -3 gpurun_out/r1j_sweep.err
echo "== per-launch profile"
timeout -k 10 120 python scripts/step_profile.py 0 > gpurun_out/r1j_step_profile.txt 2>&1
head -70 gpurun_out/r1j_step_profile.txt
