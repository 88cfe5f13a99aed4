#!/bin/bash
# r02 probe (gpurun --gpus 2): does CUDA IPC peer mapping work on the box; NCCL vs peer-swap exchange.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus.txt
nvidia-smi topo -m >> gpurun_out/r02_gpus.txt 2>&1
nvidia-smi nvlink -s -i 0 >> gpurun_out/r02_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_multi2_probe.log
for mode in 0 1; do
  QCC_B200_PEER_SWAP=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2960$mode \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/r02_probe_qft30_peer$mode.err | tail -1 > gpurun_out/r02_probe_qft30_peer$mode.json
  tail -3 gpurun_out/r02_probe_qft30_peer$mode.err
  python -c "
import json
d=json.load(open('gpurun_out/r02_probe_qft30_peer$mode.json'))
print('peer=$mode ms/step', d['ms_per_step'], 'passes', d.get('passes_per_step'), d.get('exchange'))
"
done
echo done
