#!/bin/bash
# StyleGAN2 session: parity tests at the real configs + timing probes. gpu_sg2.sh <tag>
cd "$(dirname "$0")/.." || exit 1
tag=$1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_sg2_gpu.py tests/test_stylegan2_gpu.py tests/test_stylegan2_wplus_gpu.py tests/test_stylegan2_api_gpu.py tests/test_determinism_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "^$" gpurun_out/${tag}_pytest.log | tail -25
timeout 300 python scripts/sg2_probe.py cars 9 > gpurun_out/${tag}_cars.json 2> gpurun_out/${tag}_cars.err; cat gpurun_out/${tag}_cars.json; tail -2 gpurun_out/${tag}_cars.err
timeout 300 python scripts/sg2_probe.py ffhq 8 > gpurun_out/${tag}_ffhq.json 2> gpurun_out/${tag}_ffhq.err; cat gpurun_out/${tag}_ffhq.json; tail -2 gpurun_out/${tag}_ffhq.err
