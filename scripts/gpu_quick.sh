#!/bin/bash
# quick GPU check: parity tests + bench lines for both fused variants
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for g in 1 2; do
  for w in qft30 larose28; do
    QCC_B200_FUSED_GROUPS=$g timeout 600 python bench.py --workload $w --steps 5 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${w}_g$g.json
    python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_g$g.json"))
print("$w groups=$g", "gates/s=%.0f"%d["value"], "ms/step=%.2f"%d["ms_per_step"], "passes=%.0f"%d["passes_per_step"], "roof=%.3f"%d["roofline"]["frac"], "avg_launch_ms=%.2f"%d["roofline"]["avg_launch_ms"], d["clocks"])
PY
  done
done
