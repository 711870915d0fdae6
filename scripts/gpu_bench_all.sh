#!/bin/bash
# full GPU test suite + bench lines of the three workloads on 1 GPU: gpu_bench_all.sh <tag>
cd "$(dirname "$0")/.." || exit 1
tag=$1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for w in c2 sg2_cars sg2_ffhq; do
  timeout 900 python bench.py --steps 20 --warmup 3 --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  echo "bench $w rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_$w.json; tail -2 gpurun_out/${tag}_bench_$w.err
done
timeout 300 python __graft_entry__.py > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
