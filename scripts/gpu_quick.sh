#!/bin/bash
# quick GPU check: parity tests + bench lines (+ phase-skipping timing experiments with DEBUGS="0 1 2")
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for dbg in ${DEBUGS:-0}; do
  for w in ${WORKLOADS:-qft30 larose28}; do
    QCC_B200_FUSED_DEBUG=$dbg timeout 600 python bench.py --workload $w --steps 5 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${w}_d$dbg.json
    python - <<PY
import json
try:
  d=json.load(open("gpurun_out/bench_${w}_d$dbg.json"))
  print("$w debug=$dbg", "gates/s=%.0f"%d["value"], "ms/step=%.2f"%d["ms_per_step"], "passes=%.0f"%d["passes_per_step"], "roof=%.3f"%d["roofline"]["frac"], "avg_launch_ms=%.2f"%d["roofline"]["avg_launch_ms"], d["clocks"])
except Exception as e:
  print("$w debug=$dbg FAILED", e, open("gpurun_out/bench_${w}_d$dbg.json").read()[-600:])
PY
  done
done
