#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== conv parity + biggan + step tests"
timeout -k 10 300 python -m pytest tests/test_conv_gemm_gpu.py tests/test_biggan_gpu.py tests/test_step_gpu.py -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/r1k_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r1k_pytest.log
echo "== conv parity under halo_mode=2"
P2L_OPTS="halo_mode=2" timeout -k 10 200 python -m pytest tests/test_conv_gemm_gpu.py tests/test_biggan_gpu.py -m gpu -q --timeout 150 -p no:cacheprovider 2>&1 | tail -3
echo "== sweep"
timeout -k 10 300 python scripts/sweep_options.py "halo_mode=1" "halo_mode=2" "halo_mode=1,halo=16" > gpurun_out/r1k_sweep.jsonl 2> gpurun_out/r1k_sweep.err
cat gpurun_out/r1k_sweep.jsonl; tail -3 gpurun_out/r1k_sweep.err
echo "== per-launch profile (halo_mode 0 and 1)"
timeout -k 10 120 python scripts/step_profile.py 0 1 > gpurun_out/r1k_step_profile.txt 2>&1
grep -E "ms/step|sum|K  576|N  128 K   64|N  256 K   64|N   64 K  128" gpurun_out/r1k_step_profile.txt | head -50
