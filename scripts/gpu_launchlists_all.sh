#!/bin/bash
# ncu launch lists (per-launch device time) of the three bench workloads: gpu_launchlists_all.sh <tag>
cd "$(dirname "$0")/.." || exit 1
tag=$1
mkdir -p gpurun_out
for wl in c2 sg2_cars sg2_ffhq; do
  timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches_${wl}.csv \
      python bench.py --workload $wl --steps 2 --warmup 1 --ncu > gpurun_out/${tag}_launches_${wl}.log 2>&1
  echo "$wl ncu rc=$?"; wc -l gpurun_out/${tag}_launches_${wl}.csv
done
