#!/bin/bash
# Single-GPU evidence of a round in one box session: full -m gpu suite, smoke, the bench lines of the three workloads, the
# reference arm, ncu launch lists, ncu --set full of representative launches (with source) and of every tensor-core launch of
# one step (raw metrics only). gpu_evidence.sh <tag>
cd "$(dirname "$0")/.." || exit 1
tag=$1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
for w in c2 sg2_cars sg2_ffhq; do
  timeout 900 python bench.py --steps 20 --warmup 3 --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  echo "bench $w rc=$?"; cut -c1-260 gpurun_out/${tag}_bench_$w.json; tail -2 gpurun_out/${tag}_bench_$w.err
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
echo "reference arm rc=$?"; cut -c1-260 gpurun_out/${tag}_bench_reference_arm.json
for wl in c2 sg2_cars sg2_ffhq; do
  timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches_${wl}.csv \
      python bench.py --workload $wl --steps 2 --warmup 1 --ncu > gpurun_out/${tag}_launches_${wl}.log 2>&1
  echo "$wl launch list rc=$?"; wc -l gpurun_out/${tag}_launches_${wl}.csv
done
modes=c3,c3a,d0,d33s,d33,c33,lo1,lo3
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|conv3x3" --launch-count 16 \
    -o gpurun_out/${tag}_kernels -f python scripts/one_conv.py "$modes" 2 > gpurun_out/${tag}_ncu_kernels.log 2>&1
echo "ncu one_conv rc=$?"; ls -la gpurun_out/${tag}_kernels.ncu-rep
# every tensor-core launch of one BigGAN step (119) and the kernels of one StyleGAN2 chunk: metrics only, exported here
timeout -k 5 1500 ncu --set full --clock-control none -k regex:"conv_gemm|conv3x3" --launch-skip 130 --launch-count 119 \
    -o /tmp/${tag}_c2_step -f python bench.py --workload c2 --steps 2 --warmup 1 --ncu > gpurun_out/${tag}_ncu_c2_step.log 2>&1
echo "ncu c2 step rc=$?"
ncu -i /tmp/${tag}_c2_step.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_c2_step_raw.csv 2>/dev/null; wc -c gpurun_out/${tag}_ncu_c2_step_raw.csv
timeout -k 5 900 ncu --set full --clock-control none -k regex:"conv_gemm|conv3x3|torgb|sg_mapping|sg_gemm" --launch-skip 100 --launch-count 100 \
    -o /tmp/${tag}_sg2_step -f python scripts/sg2_probe.py cars 9 1 > gpurun_out/${tag}_ncu_sg2_step.log 2>&1
echo "ncu sg2 step rc=$?"
ncu -i /tmp/${tag}_sg2_step.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_sg2_step_raw.csv 2>/dev/null; wc -c gpurun_out/${tag}_ncu_sg2_step_raw.csv
