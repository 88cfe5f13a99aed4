#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/r02_final_bench_qft30.err | tail -1 > gpurun_out/r02_final_bench_qft30.json
timeout 900 python bench.py --workload larose28 --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | tail -1 > gpurun_out/r02_final_bench_larose28.json
timeout 900 python bench.py --workload grover --qubits 28 2> /dev/null | tail -1 > gpurun_out/r02_final_grover28_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_final_launches_qft30.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r02_final_ncu_launch.log 2>&1
python - <<'PY'
import json
for f in ("bench_qft30","bench_larose28","grover28_1gpu"):
  try:
    d=json.load(open(f"gpurun_out/r02_final_{f}.json"))
    print(f, "value %.0f ms/step %.2f"%(d["value"], d["ms_per_step"]), "roofline", (d.get("roofline") or {}).get("frac"), "e2e", d.get("e2e"), "res", (d.get("e2e_resident") or {}).get("value"), d.get("check"), d.get("wall_s"), d.get("device_ms"), d.get("clocks"))
  except Exception as e:
    print(f, "FAILED", e)
PY
echo done
