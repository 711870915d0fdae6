#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in lo3 lo1; do
  timeout -k 10 150 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -k regex:conv_gemm \
      -o gpurun_out/r1g_$m -f python scripts/one_conv.py $m 4 > gpurun_out/r1g_ncu_$m.log 2>&1
  echo "ncu $m rc=$?"
done
timeout -k 10 120 python -m pytest tests/test_inner_loop_gpu.py -m gpu -q --timeout 150 -p no:cacheprovider 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
