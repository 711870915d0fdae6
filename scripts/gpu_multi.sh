#!/bin/bash
# multi-GPU session (gpurun --gpus N): NCCL bitwise test + bench at N GPUs. gpu_multi.sh <tag> <N>
cd "$(dirname "$0")/.." || exit 1
tag=$1; n=$2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_n$n.json; tail -3 gpurun_out/${tag}_bench_n$n.err
