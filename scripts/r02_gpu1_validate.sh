#!/bin/bash
# r02 (1 GPU): full -m gpu suite, default bench + larose28 line, reference CPU sweep, smoke.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/r02_bench_qft30.err | tail -1 > gpurun_out/r02_bench_qft30.json
timeout 900 python bench.py --workload larose28 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_bench_larose28.err | tail -1 > gpurun_out/r02_bench_larose28.json
timeout 900 python bench.py --workload grover --qubits 28 2> gpurun_out/r02_grover28_1gpu.err | tail -1 > gpurun_out/r02_grover28_1gpu.json
timeout 900 python bench.py --cpu-sweep 2> gpurun_out/r02_cpu_sweep.err | tail -1 > gpurun_out/r02_cpu_sweep.json
python - <<'PY'
import json
for f in ("bench_qft30","bench_larose28","grover28_1gpu"):
  try:
    d=json.load(open(f"gpurun_out/r02_{f}.json"))
    print(f, "value %.0f ms/step %.2f"%(d["value"], d["ms_per_step"]), "roofline", d.get("roofline",{}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), d.get("e2e_resident",{}) and d["e2e_resident"].get("value"), d.get("check"), d.get("host_overhead_frac"), d.get("clocks"))
  except Exception as e:
    print(f, "FAILED", e); print(open(f"gpurun_out/r02_{f}.err").read()[-1500:])
try:
  d=json.load(open("gpurun_out/r02_cpu_sweep.json"))
  for x in d["xgates"]: print(x["qubits"], x.get("h_gates_per_s"), x.get("h_gbs_algorithmic"), x.get("cx_seconds"))
  for x in d["libq"]: print(x["qubits"], x["value"])
except Exception as e:
  print("sweep FAILED", e)
PY
echo done
