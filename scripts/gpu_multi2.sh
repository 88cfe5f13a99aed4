#!/bin/bash
# 2-GPU visit (gpurun --gpus 2): sharded parity tests at world 2, then the default bench at N=2
# (strong scaling: the same 30-qubit QFT sharded over both GPUs) and the larose stream.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_multi2.log
for wl in qft30 larose28; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus 2 --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/scale2_$wl.json
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/scale2_$wl.json"))
  print("$wl N=2 qubits=%d gates/s=%.0f ms/step=%.1f passes=%.1f"%(d["config"]["qubits"], d["value"], d["ms_per_step"], d["passes_per_step"]), d.get("exchange"))
except Exception as e:
  print("$wl N=2 FAILED", e, open("gpurun_out/scale2_$wl.json").read()[-800:])
PY
done
