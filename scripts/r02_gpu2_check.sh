#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "push-2 or swap-2" 2>&1 | tail -4
PORT=$((29300 + RANDOM % 500))
for env in "X=1" "QCC_B200_PEER_BARRIER=1" "X=2"; do
timeout 600 env $env python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r02_qft30_2gpu_c.json
python -c "
import json; d=json.load(open('gpurun_out/r02_qft30_2gpu_c.json')); print('$env', d['ms_per_step'], d['passes_per_step'], d['kernel_ms'], d['exchange']['nvlink_gbs_per_direction_rank0'])"
PORT=$((PORT+1))
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((PORT+2)) bench.py --gpus 2 --workload grover --qubits 28 2>/dev/null | tail -1 > gpurun_out/r02_grover28_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_grover28_2gpu.json')); print('grover28x2', d['wall_s'], d['device_ms'], d['passes'], d['check']['ok'], d['kernel_ms'], d['kernel_launches'])"
