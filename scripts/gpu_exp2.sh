#!/bin/bash
# experiment visit: parity, QFT-30 under debug / persist variants, one ncu --set full capture
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --workload ${WL:-qft30} --steps 5 --no-e2e --no-cpu-baseline --no-secondary 2>&1 | tail -1 > gpurun_out/exp_$name.json
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/exp_$name.json"))
  print("$name", "ms/step=%.2f"%d["ms_per_step"], "passes=%.0f"%d["passes_per_step"], "roof=%.3f"%d["roofline"]["frac"], "avg_launch_ms=%.2f"%d["roofline"]["avg_launch_ms"], "W=%.0f"%d["clocks"]["power_w_max"])
except Exception as e:
  print("$name FAILED", e, open("gpurun_out/exp_$name.json").read()[-400:])
PY
}
for v in ${VARIANTS:-d0:QCC_B200_FUSED_DEBUG=0}; do
  name=${v%%:*}; envs=${v#*:}
  run $name ${envs//,/ }
done
if [ "${NCU:-0}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s ${NCU_SKIP:-9} -c ${NCU_COUNT:-2} -f -o gpurun_out/${NCU_OUT:-prof_fused} \
    python bench.py --workload ${WL:-qft30} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
fi
