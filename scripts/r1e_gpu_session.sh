#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest inner loop + conv parity" | tee gpurun_out/r1e_stage.txt
timeout -k 10 240 python -m pytest tests/test_inner_loop_gpu.py -m gpu -q -s --timeout 150 -p no:cacheprovider > gpurun_out/r1e_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r1e_stage.txt
grep -E "passed|failed|fused vs|graph vs|3\+3|GradientOptimizer fused|BasinCMA fused" gpurun_out/r1e_pytest.log | tail -12
echo "== sweep" | tee -a gpurun_out/r1e_stage.txt
timeout -k 10 400 python scripts/sweep_options.py "deep=1" "deep=1,deep_kmin=2" "tma_kmax=576" "tma_kmax=1152" "tma_out=0" "halo_mode=1" "halo_mode=2" "deep=1,tma_kmax=576" > gpurun_out/r1e_sweep.jsonl 2> gpurun_out/r1e_sweep.err
echo "sweep rc=$?" | tee -a gpurun_out/r1e_stage.txt
cat gpurun_out/r1e_sweep.jsonl; tail -3 gpurun_out/r1e_sweep.err
echo "== bench" | tee -a gpurun_out/r1e_stage.txt
timeout -k 10 240 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err
echo "bench rc=$?" | tee -a gpurun_out/r1e_stage.txt
python -c "
import json; d=json.load(open('gpurun_out/r1e_bench.json')); print(d['value'], d['e2e']['value'], d['inner_loop'])"
tail -3 gpurun_out/r1e_bench.err
