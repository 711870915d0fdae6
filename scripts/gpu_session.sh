#!/bin/bash
# One GPU session: (optional) GPU test suite, per-launch profile of the bench step (this build and, when present, the
# round-1 build for A/B), option sweep, bench. Usage: gpu_session.sh <tag> [tests] [sweep "cfg" ...]
cd "$(dirname "$0")/.." || exit 1
tag=$1; shift
mkdir -p gpurun_out
if [ "$1" = "tests" ]; then
  shift
  timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
fi
[ -f pix2latent_b200/libp2l_r1.so ] && P2L_LIB=$PWD/pix2latent_b200/libp2l_r1.so timeout 300 python scripts/step_profile.py > gpurun_out/${tag}_profile_r1.jsonl 2> gpurun_out/${tag}_profile_r1.err
timeout 300 python scripts/step_profile.py > gpurun_out/${tag}_profile.jsonl 2> gpurun_out/${tag}_profile.err
head -1 gpurun_out/${tag}_profile*.jsonl
if [ "$1" = "sweep" ]; then
  shift
  timeout 900 python scripts/sweep_options.py "$@" > gpurun_out/${tag}_sweep.jsonl 2> gpurun_out/${tag}_sweep.err
  cut -c1-200 gpurun_out/${tag}_sweep.jsonl
fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/${tag}_bench.json
