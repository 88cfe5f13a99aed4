#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
  for lib in qcc_b200/lib_r01/libqcc_b200.so qcc_b200/lib/libqcc_b200.so; do
    python scripts/ab_lib.py $lib qft 30
    python scripts/ab_lib.py $lib larose 28
  done
done 2>&1 | tee gpurun_out/r02_ab.log
timeout 900 python bench.py --workload grover --qubits 28 2> gpurun_out/r02_grover28_1gpu_ccu.err | tail -1 > gpurun_out/r02_grover28_1gpu_ccu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_grover28_1gpu_ccu.json'))
print('grover28', d['wall_s'], d['device_ms'], d['passes'], d['check'], d['roofline']['frac'], d['host_overhead_frac'])"
timeout 600 python -m pytest tests -m gpu -x -q -k "grover or toffoli or two_controls or composites or order or multi" 2>&1 | tail -4
