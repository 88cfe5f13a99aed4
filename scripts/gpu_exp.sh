#!/bin/bash
set -u
mkdir -p gpurun_out
for st in 0 4000 8000 12000 20000; do for dbg in 0 1; do
  QCC_B200_FUSED_STAGGER_NS=$st QCC_B200_FUSED_DEBUG=$dbg timeout 600 python bench.py --workload qft30 --steps 5 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp.json
  python - <<PY
import json
d=json.load(open("gpurun_out/exp.json"))
print("stagger=$st debug=$dbg ms/step=%.2f avg_launch_ms=%.2f roof=%.3f"%(d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"]), d["clocks"]["power_w_max"])
PY
done; done
