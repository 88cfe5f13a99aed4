#!/bin/bash
# Final single-GPU visit of a round: parity suite, smoke, both bench arms, ncu launch lists and one
# --set full capture of each hot kernel; summaries land in gpurun_out/ (copy what matters to profiles/).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_qft30.json
timeout 900 python bench.py --workload larose28 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_larose28.json
timeout 900 python bench.py --workload hsweep30 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_hsweep30.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_qft30.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 9 -c 3 -f -o gpurun_out/prof_fused_final \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
for f in qft30 larose28 hsweep30 reference; do python - <<PY
import json
try:
  d=json.load(open("gpurun_out/bench_$f.json"))
  r=d.get("roofline") or {}
  print("$f", "value=%.4g"%d["value"], d["unit"], "ms/step=%.2f"%d["ms_per_step"], "roof=", r.get("kernel"), r.get("frac"), "e2e=", (d.get("e2e") or {}).get("value"), "cpu=", (d.get("cpu_baseline") or {}).get("value"), "single=", (d.get("roofline_single_gate") or {}).get("frac"))
except Exception as e:
  print("$f FAILED", e, open("gpurun_out/bench_$f.json").read()[-500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 20 -c 3 -f -o gpurun_out/prof_larose_final \
  python bench.py --workload larose28 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_full_larose.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_larose28.csv \
  python bench.py --workload larose28 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_launch_larose.log 2>&1
