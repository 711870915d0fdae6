#!/bin/bash
# Final validation of the round: full -m gpu suite, smoke, bench (both arms), ncu launch list of one bench step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout -k 10 300 python -m pytest tests -m gpu -x -q --timeout 150 -p no:cacheprovider > gpurun_out/val_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/val_pytest.log
echo "== smoke"
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"
timeout -k 10 300 python bench.py --steps 100 --warmup 3 > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err
echo "bench rc=$?"; cat gpurun_out/val_bench.json; tail -3 gpurun_out/val_bench.err
echo "== ncu launch list"
timeout -k 10 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/val_launches.csv python bench.py --ncu --steps 2 --warmup 1 > gpurun_out/val_ncu_bench.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/val_ncu_bench.log; wc -l gpurun_out/val_launches.csv
