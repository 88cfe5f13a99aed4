#!/bin/bash
set -u
mkdir -p gpurun_out
{
for i in 1 2; do
python scripts/ab_lib.py qcc_b200/lib_r01/libqcc_b200.so qft 30
python scripts/ab_lib.py qcc_b200/lib_a/libqcc_b200.so qft 30
python scripts/ab_lib.py qcc_b200/lib/libqcc_b200.so qft 30
done
python scripts/ab_lib.py qcc_b200/lib_a/libqcc_b200.so larose 28
python scripts/ab_lib.py qcc_b200/lib/libqcc_b200.so larose 28
} 2>&1 | tee gpurun_out/r02_ab3.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -3
