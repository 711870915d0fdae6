#!/bin/bash
# One GPU session: new -m gpu tests first, a short bench run, smoke(), then the rest of the -m gpu suite.
# Every stage has its own timeout and writes into gpurun_out/ as it goes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/r1d_smi.txt 2>&1
NEW="tests/test_inner_loop_gpu.py tests/test_transform_gpu.py tests/test_stylegan2_wplus_gpu.py"
echo "== pytest new" | tee gpurun_out/r1d_stage.txt
timeout -k 10 300 python -m pytest $NEW -m gpu -q -s --durations=15 --timeout 150 -p no:cacheprovider > gpurun_out/r1d_pytest_new.log 2>&1
echo "pytest new rc=$?" | tee -a gpurun_out/r1d_stage.txt
tail -30 gpurun_out/r1d_pytest_new.log
echo "== bench" | tee -a gpurun_out/r1d_stage.txt
timeout -k 10 240 python bench.py --steps 40 --warmup 3 > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err
echo "bench rc=$?" | tee -a gpurun_out/r1d_stage.txt
cat gpurun_out/r1d_bench.json
tail -5 gpurun_out/r1d_bench.err
echo "== smoke" | tee -a gpurun_out/r1d_stage.txt
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1d_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/r1d_stage.txt
tail -3 gpurun_out/r1d_smoke.log
echo "== pytest old" | tee -a gpurun_out/r1d_stage.txt
timeout -k 10 420 python -m pytest tests -m gpu -q --durations=15 --timeout 200 -p no:cacheprovider \
    --deselect tests/test_inner_loop_gpu.py --deselect tests/test_transform_gpu.py --deselect tests/test_stylegan2_wplus_gpu.py \
    > gpurun_out/r1d_pytest_old.log 2>&1
echo "pytest old rc=$?" | tee -a gpurun_out/r1d_stage.txt
tail -25 gpurun_out/r1d_pytest_old.log
