#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 85 ncu --set full --clock-control none --import-source on --launch-skip 0 --launch-count 6 -k regex:"conv_gemm|conv3x3" \
    -o gpurun_out/prof_kernels -f python scripts/one_conv.py c3,d0,c33 2 > gpurun_out/prof_ncu.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/prof_ncu.log
timeout -k 5 75 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_gemm|conv3x3" -c 236 --csv \
    --log-file gpurun_out/prof_traffic.csv python bench.py --ncu --steps 1 --warmup 1 > gpurun_out/prof_traffic.log 2>&1
echo "ncu traffic rc=$?"; wc -l gpurun_out/prof_traffic.csv
