#!/bin/bash
# experiment visit: parity suite, then order finding N=35 a=9 (26 qubits) and Grover-28 under env variants
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-base:QCC_B200_FUSED_DEBUG=0}; do
  name=${v%%:*}; envs=${v#*:}
  env ${envs//,/ } timeout 300 python - <<PY 2>&1 | tail -2
import time, sys
sys.path.insert(0, ".")
from qcc_b200 import workloads
t0 = time.perf_counter()
qc, aux, up, down = workloads.order_finding(35, 9)
qc.sync()
t1 = time.perf_counter()
found = workloads.order_readout(qc, 35, 9)
c = qc.dev.counters()
import math
rs = sorted({r for _, _, _, r, _ in found})
print("order35 $name wall_s=%.2f gates=%d passes=%d launches=%d peaks=%d rs=%s lcm=%d" % (t1 - t0, c["gates_applied"], c["passes"], c["kernel_launches"], len(found), rs, math.lcm(*rs)))
qc.close()
PY
  env ${envs//,/ } timeout 300 python bench.py --workload grover --qubits 28 --gpus 1 2>&1 | tail -1 > gpurun_out/alg_grover_$name.json
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/alg_grover_$name.json"))
  print("grover28 $name wall_ms=%.0f gates=%d passes=%d ok=%s"%(d["ms_per_step"], d["gates"], d["passes"], d["check"].get("ok")))
except Exception as e:
  print("grover $name FAILED", e)
PY
done
