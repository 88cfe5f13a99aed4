#!/bin/bash
# Multi-GPU visit (run with gpurun --gpus 8): sharded parity tests at world 2/4/8, the 1/2/4/8
# scaling line of the default bench, and the two multi-GPU configs of BASELINE.json.
set -u
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
nvidia-smi topo -m >> gpurun_out/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_multi.log
P=29600
for n in 1 2 4 8; do
  P=$((P+1))
  if [ $n = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/scale_qft_$n.json
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/scale_qft_$n.json
  fi
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/scale_qft_$n.json"))
  print("N=$n qubits=%d gates/s=%.0f ms/step=%.1f passes=%.1f"%(d["config"]["qubits"], d["value"], d["ms_per_step"], d["passes_per_step"]), d.get("exchange"))
except Exception as e:
  print("N=$n FAILED", e, open("gpurun_out/scale_qft_$n.json").read()[-800:])
PY
done
# BASELINE.json north_star: "8-GPU 34-qubit QFT wall time and NVLink GB/s" (32 GiB shards)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29619 \
  bench.py --gpus 8 --qubits 34 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee gpurun_out/qft34_8gpu.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 \
  bench.py --gpus 8 --workload supremacy --qubits 34 --depth 20 2>&1 | tail -1 | tee gpurun_out/supremacy34_8gpu.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 \
  bench.py --gpus 4 --workload grover --qubits 32 2>&1 | tail -1 | tee gpurun_out/grover32_4gpu.json
echo done
