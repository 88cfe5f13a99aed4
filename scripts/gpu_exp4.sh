#!/bin/bash
# experiment visit: parity suite, then whole-algorithm runs (grover / supremacy, one GPU) under env variants
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-base:QCC_B200_FUSED_DEBUG=0}; do
  name=${v%%:*}; envs=${v#*:}
  for spec in "grover:${GROVER_Q:-28}" "supremacy:${SUP_Q:-30}"; do
    wl=${spec%%:*}; q=${spec#*:}
    env ${envs//,/ } timeout 300 python bench.py --workload $wl --qubits $q --gpus 1 2>&1 | tail -1 > gpurun_out/alg_${wl}_$name.json
    python - <<PY
import json
try:
  d=json.load(open("gpurun_out/alg_${wl}_$name.json"))
  print("$wl $name q=$q wall_ms=%.0f gates=%d passes=%d ok=%s"%(d["ms_per_step"], d["gates"], d["passes"], d["check"].get("ok")))
except Exception as e:
  print("$wl $name FAILED", e, open("gpurun_out/alg_${wl}_$name.json").read()[-400:])
PY
  done
done
