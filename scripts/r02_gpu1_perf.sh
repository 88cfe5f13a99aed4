#!/bin/bash
# r02 (1 GPU): A/B of kernel variants on ONE box, then the ncu launch list and one full capture.
set -u
mkdir -p gpurun_out
{
python scripts/ab_lib.py qcc_b200/lib_r01/libqcc_b200.so qft 30
python scripts/ab_lib.py qcc_b200/lib/libqcc_b200.so qft 30
for st in 3000 8000 16000; do
  echo "stagger $st"; QCC_B200_STAGGER=$st python scripts/ab_lib.py qcc_b200/lib/libqcc_b200.so qft 30
done
echo "persist 3"; QCC_B200_FUSED_PERSIST=3 python scripts/ab_lib.py qcc_b200/lib/libqcc_b200.so qft 30
echo "persist 3 stagger 8000"; QCC_B200_FUSED_PERSIST=3 QCC_B200_STAGGER=8000 python scripts/ab_lib.py qcc_b200/lib/libqcc_b200.so qft 30
python scripts/ab_lib.py qcc_b200/lib_r01/libqcc_b200.so larose 28
python scripts/ab_lib.py qcc_b200/lib/libqcc_b200.so larose 28
QCC_B200_STAGGER=8000 python scripts/ab_lib.py qcc_b200/lib/libqcc_b200.so larose 28
} 2>&1 | tee gpurun_out/r02_ab2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_qft30.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r02_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 9 -c 3 -o gpurun_out/r02_prof_fused \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/r02_prof_fused.ncu-rep
echo done
