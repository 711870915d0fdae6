#!/bin/bash
# GPU session r2a: full GPU test suite (determinism + parity at the benchmarked configs), option sweeps, bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
echo "stage pytest" > gpurun_out/r2a_stage.txt
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_stage.txt
tail -5 gpurun_out/r2a_pytest.log
echo "stage sweep" >> gpurun_out/r2a_stage.txt
timeout 900 python scripts/sweep_options.py "attn_fused=1" "attn_fused=1,attn_emit_t=1" "splitk=1" "prefetch_saved=1" \
    "sub_mb=32" "sub_mb=64" "sub_mb=96" "sub_mb=128" "sub_mb=192" "sub_mb=64,sub_min_tiles=296" "sub_mb=96,sub_min_tiles=1184" \
    "attn_fused=1,attn_emit_t=1,sub_mb=96" > gpurun_out/r2a_sweep.jsonl 2> gpurun_out/r2a_sweep.err
echo "sweep rc=$?" >> gpurun_out/r2a_stage.txt
P2L_LIB=$PWD/pix2latent_b200/libp2l_nopre.so timeout 300 python scripts/sweep_options.py "prefetch_saved=0" > gpurun_out/r2a_sweep_nopre.jsonl 2> gpurun_out/r2a_sweep_nopre.err
echo "stage bench" >> gpurun_out/r2a_stage.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_stage.txt
timeout 300 python __graft_entry__.py > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2a_stage.txt
cat gpurun_out/r2a_stage.txt
grep -h "cand_per_s\|error" gpurun_out/r2a_sweep.jsonl gpurun_out/r2a_sweep_nopre.jsonl | cut -c1-220
