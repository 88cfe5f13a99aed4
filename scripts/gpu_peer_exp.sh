#!/bin/bash
# Multi-GPU visit for the peer-memory exchange (gpurun --gpus 2 | 4 | 8): sharded parity under both
# exchanges, then the default bench with NCCL send/recv vs QCC_B200_PEER_SWAP=1 (peer-swap kernel + wide
# victim window + exchange hoisting).  Written in round 1 without GPU budget left: run it first thing.
set -u
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_multi_peer.log
P=29650
for v in nccl:QCC_B200_PEER_SWAP=0 peer:QCC_B200_PEER_SWAP=1; do
  name=${v%%:*}; envs=${v#*:}
  for wl in qft30 larose28; do
    P=$((P+1))
    env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --workload $wl --steps 6 --warmup 4 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/peer_${wl}_${name}_$N.json
    python - <<PY
import json
try:
  d=json.load(open("gpurun_out/peer_${wl}_${name}_$N.json"))
  x=d.get("exchange") or {}
  print("$wl $name N=$N ms/step=%.2f gates/s=%.0f passes=%.1f exch/step=%.1f exch_ms=%.2f nvlink_gbs=%.0f norm2=%.12f"%(d["ms_per_step"], d["value"], d["passes_per_step"], x.get("per_step",0), x.get("ms_per_step_rank0",0), x.get("nvlink_gbs_per_direction_rank0",0), d.get("norm2_after", 0)))
except Exception as e:
  print("$wl $name FAILED", e, open("gpurun_out/peer_${wl}_${name}_$N.json").read()[-600:])
PY
  done
done
