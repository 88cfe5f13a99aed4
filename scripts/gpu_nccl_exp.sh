#!/bin/bash
# 2-GPU visit: NCCL point-to-point settings for the pair exchange (QFT-30 sharded over 2 GPUs).
set -u
mkdir -p gpurun_out
run() {
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 \
    bench.py --gpus 2 --workload qft30 --steps 6 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/nccl_$name.log 2>&1
  tail -1 gpurun_out/nccl_$name.log > gpurun_out/nccl_$name.json
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/nccl_$name.json"))
  x=d.get("exchange") or {}
  print("$name ms/step=%.1f exch/step=%.1f exch_ms=%.2f nvlink_gbs=%.0f"%(d["ms_per_step"], x.get("per_step",0), x.get("ms_per_step_rank0",0), x.get("nvlink_gbs_per_direction_rank0",0)))
except Exception as e:
  print("$name FAILED", e, open("gpurun_out/nccl_$name.log").read()[-600:])
PY
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P
grep -i -E "channel|p2p|NVLS|nvlink" gpurun_out/nccl_default.log | head -12
run ch32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run ch16 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=16
run ce NCCL_P2P_USE_CUDA_MEMCPY=1
