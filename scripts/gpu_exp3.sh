#!/bin/bash
# experiment visit: parity suite, then each workload in $WLS (default larose28 qft30) under the env
# variants in $VARIANTS (name:K=V,K=V ...), then one ncu --set full capture of $NCU_WL's fused passes.
set -u
mkdir -p gpurun_out
timeout ${PYTEST_TIMEOUT:-400} python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
run() {  # workload, name, env...
  local wl=$1 name=$2; shift 2
  env "$@" timeout ${BENCH_TIMEOUT:-300} python bench.py --workload $wl --steps ${STEPS:-5} --no-e2e --no-cpu-baseline --no-secondary 2>&1 | tail -1 > gpurun_out/exp_${wl}_$name.json
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/exp_${wl}_$name.json"))
  print("$wl $name", "ms/step=%.2f"%d["ms_per_step"], "passes=%.0f"%d["passes_per_step"], "roof=%.3f"%d["roofline"]["frac"], "avg_launch_ms=%.3f"%d["roofline"]["avg_launch_ms"], "W=%.0f"%d["clocks"]["power_w_max"], "mhz=%s"%d["clocks"]["sm_mhz"])
except Exception as e:
  print("$wl $name FAILED", e, open("gpurun_out/exp_${wl}_$name.json").read()[-400:])
PY
}
for wl in ${WLS:-larose28 qft30}; do
  for v in ${VARIANTS:-base:QCC_B200_FUSED_DEBUG=0}; do
    name=${v%%:*}; envs=${v#*:}
    run $wl $name ${envs//,/ }
  done
done
if [ "${NCU_WL:-}" != "" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s ${NCU_SKIP:-9} -c ${NCU_COUNT:-3} -f -o gpurun_out/${NCU_OUT:-prof_fused} \
    python bench.py --workload $NCU_WL --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
fi
