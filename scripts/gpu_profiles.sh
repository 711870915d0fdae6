#!/bin/bash
# ncu --set full (+ source) of representative launches in isolation: gpu_profiles.sh <tag> <modes> [kernel regex]
cd "$(dirname "$0")/.."
tag=$1; modes=$2; rx=${3:-"conv_gemm|conv3x3"}
mkdir -p gpurun_out
n=$(echo "$modes" | tr ',' '\n' | wc -l)
# every mode runs 2 launches: profile the second of each (warm caches / attributes set)
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-count $((2 * n)) \
    -o gpurun_out/${tag}_kernels -f python scripts/one_conv.py "$modes" 2 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/${tag}_kernels.ncu-rep
