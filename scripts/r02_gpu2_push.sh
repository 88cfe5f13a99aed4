#!/bin/bash
# r02 (gpurun --gpus 2): sharded parity under the three exchanges, then QFT-30 / larose-28 over 2 GPUs.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r02_pytest_multi2.log
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  PORT=$((29700 + RANDOM % 200))
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus 2 --no-cpu-baseline --no-e2e "$@" 2> gpurun_out/r02_$name.err | tail -1 > gpurun_out/r02_$name.json
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/r02_$name.json"))
  print("$name ms/step %.2f passes %.1f norm %.12f"%(d["ms_per_step"], d["passes_per_step"], d["norm2_after"]), d["kernel_ms"], d["exchange"])
except Exception as e:
  print("$name FAILED", e); print(open("gpurun_out/r02_$name.err").read()[-1500:])
PY
}
run qft30_2gpu_push X=1 -- --steps 6 --warmup 3
run qft30_2gpu_push_perstep X=1 -- --steps 6 --warmup 3 --flush-per-step
run qft30_2gpu_push_nofuse QCC_B200_NO_PUSH_FUSE=1 -- --steps 6 --warmup 3
run qft30_2gpu_swap QCC_B200_EXCHANGE=swap -- --steps 6 --warmup 3
run larose28_2gpu_push X=1 -- --workload larose28 --steps 3 --warmup 3
run qft31_2gpu_weak X=1 -- --steps 4 --warmup 3 --weak
echo done
