#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest (defaults: pdl=1 deep=1)"
timeout -k 10 300 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/r1i_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r1i_pytest.log
echo "== ncu halo c33"
P2L_OPTS="halo_mode=2" timeout -k 10 150 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -k regex:conv3x3_halo \
      -o gpurun_out/r1i_halo_c33 -f python scripts/one_conv.py c33 4 > gpurun_out/r1i_ncu_halo.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r1i_ncu_halo.log
echo "== ncu plain c33 (reference)"
timeout -k 10 150 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -k regex:conv_gemm \
      -o gpurun_out/r1i_plain_c33 -f python scripts/one_conv.py c33 4 > gpurun_out/r1i_ncu_plain.log 2>&1
echo "ncu rc=$?"
