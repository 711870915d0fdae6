#!/bin/bash
# ncu launch list (per-launch device time) of a command: gpu_launchlist.sh <tag> <skip> <count> <cmd...>
cd "$(dirname "$0")/.." || exit 1
tag=$1; skip=$2; count=$3; shift 3
mkdir -p gpurun_out
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $skip -c $count --csv --log-file gpurun_out/${tag}_launches.csv "$@" > gpurun_out/${tag}_launches.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/${tag}_launches.csv; tail -2 gpurun_out/${tag}_launches.log
