#!/bin/bash
set -u
mkdir -p gpurun_out
PORT=$((29300 + RANDOM % 500))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r02_final_qft30_2gpu.err | grep '^{' > gpurun_out/r02_final_qft30_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_final_qft30_2gpu.json')); print(d['ms_per_step'], d['passes_per_step'], d['kernel_ms'], d['exchange']['nvlink_gbs_per_direction_rank0'], d['e2e'], d['e2e_resident'])" || tail -20 gpurun_out/r02_final_qft30_2gpu.err
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "push-2" 2>&1 | tail -2
