#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== sweep"
timeout -k 10 300 python scripts/sweep_options.py "pdl=1" "pdl=1,deep=1" "pdl=1" > gpurun_out/r1h_sweep.jsonl 2> gpurun_out/r1h_sweep.err
echo "sweep rc=$?"; cat gpurun_out/r1h_sweep.jsonl; tail -3 gpurun_out/r1h_sweep.err
echo "== pytest under pdl"
P2L_OPTS="pdl=1" timeout -k 10 300 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/r1h_pytest_pdl.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r1h_pytest_pdl.log
echo "== bench under pdl"
P2L_OPTS="pdl=1" timeout -k 10 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_bench_pdl.json 2> gpurun_out/r1h_bench_pdl.err
python -c "
import json; d=json.load(open('gpurun_out/r1h_bench_pdl.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['inner_loop'])"
