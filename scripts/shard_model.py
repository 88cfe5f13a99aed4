#!/usr/bin/env python
"""Host-only model of the sharded lowering: exchange events, exchanged pairs and fused passes per step
for a workload repeated `reps` times in ONE flush (python scripts/shard_model.py qft 30 8 5).
Uses the library's own lowering (qb_shard_lower_json) and planner (qb_plan_json); no GPU."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qcc_b200 import _cabi, workloads  # noqa: E402


def to_bits(stream, n):
  out = []
  for kind, ctl, tgt, m in stream:
    m = np.asarray(m).reshape(4)
    out.append((0 if kind == 1 else 1 << (n - 1 - ctl), n - 1 - tgt, m))
  return out


def model(name, n, world, reps, mode="push", rank=0):
  if name == "qft":
    stream = workloads.qft(n)
  elif name == "larose":
    stream = workloads.larose(n, n)
  elif name == "supremacy":
    stream = workloads.supremacy(n, 20, seed=0)
  else:
    raise SystemExit(name)
  gates = to_bits(stream, n) * reps
  p = int(np.log2(world))
  nl = n - p
  if mode == "push":
    os.environ["QCC_B200_VICTIM_WINDOW"] = str(min(max(6, nl - 5), nl - 3))
    os.environ["QCC_B200_HOIST"] = "1"
    os.environ["QCC_B200_PREFETCH"] = "1"
    os.environ.setdefault("QCC_B200_LAND", "0")
  elif mode == "swap":
    os.environ["QCC_B200_VICTIM_WINDOW"] = str(max(6, nl - 5))
    os.environ["QCC_B200_HOIST"] = "1"
    os.environ["QCC_B200_PREFETCH"] = "0"
  else:
    for k in ("QCC_B200_VICTIM_WINDOW", "QCC_B200_HOIST", "QCC_B200_PREFETCH"):
      os.environ.pop(k, None)
  plan = json.loads(_cabi.shard_lower_json(n, world, rank, gates, canonicalize=False))
  events = pairs = passes = fusedev = 0
  prev_fused_tail = False
  for st in plan["steps"]:
    if st["kind"] == 1:
      events += 1
      pairs += len(st["pairs"])
      fusedev += prev_fused_tail
      prev_fused_tail = False
      continue
    g = [(x["ctl_mask"], x["target"], np.array([complex(x["m"][2 * i], x["m"][2 * i + 1]) for i in range(4)]))
         for x in st["gates"]]
    if not g:
      prev_fused_tail = False
      continue
    pl = json.loads(_cabi.plan_json(nl, g, 12))["passes"]
    passes += len(pl)
    prev_fused_tail = pl[-1]["single_gate"] < 0
  return dict(workload=name, n=n, world=world, reps=reps, mode=mode, gates=len(gates), events=events, pairs=pairs,
              passes=passes, events_on_a_pass=fusedev, per_rep=dict(events=events / reps, pairs=pairs / reps,
                                                                   passes=passes / reps))


if __name__ == "__main__":
  name, n, world, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
  mode = sys.argv[5] if len(sys.argv) > 5 else "push"
  print(json.dumps(model(name, n, world, reps, mode)))
