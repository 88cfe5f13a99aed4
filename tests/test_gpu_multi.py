"""GPU, >= 2 devices: the sharded state end to end (NCCL pair exchange, rank-bit predicates,
collective readouts) against the oracle.  One process per GPU, rendezvous over gloo on
127.0.0.1; skipped on single-GPU boxes (run with `gpurun --gpus 2`)."""
import math
import os
import socket
import sys

import numpy as np
import pytest

from helpers import ROOT, oracle, random_state
from qcc_b200 import _cabi

pytestmark = pytest.mark.gpu


def _ngpus():
  import ctypes
  n = ctypes.c_int(0)
  try:
    _cabi.lib().qb_device_count(ctypes.byref(n))
  except Exception:  # pylint: disable=broad-except
    return 0
  return n.value


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _streams(n):
  H, X, V = oracle.GATES["h"], oracle.GATES["x"], oracle.GATES["v"]
  qft = []
  for i in reversed(range(n)):
    qft.append((1, 0, i, H))
    for j in reversed(range(i)):
      qft.append((2, i, j, oracle.u1(math.pi / 2 ** (i - j))))
  rng = np.random.default_rng(3)
  names = list(oracle.GATES)
  rnd = []
  for _ in range(150):
    r = rng.random()
    m = oracle.u1(float(rng.uniform(-3, 3))) if r < 0.2 else (
        oracle.rotation([0, 0, 1.0], float(rng.uniform(-3, 3))) if r < 0.3 else oracle.GATES[names[rng.integers(len(names))]])
    t = int(rng.integers(n))
    if rng.random() < 0.45:
      rnd.append((1, 0, t, m))
    else:
      c = int(rng.integers(n))
      if c != t:
        rnd.append((2, c, t, m))
  larose = []
  for _ in range(2):
    for bit in range(n):
      larose += [(1, 0, bit, H), (1, 0, bit, V)]
      if bit:
        larose.append((2, bit, 0, X))
  return {"qft": qft, "random": rnd, "larose": larose}


def _worker(rank, world, port, n, out_dir):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "tests"))
  import torch
  import torch.distributed as dist
  dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
  try:
    nl = n - int(math.log2(world))
    for fusion in (True, False):
      ids = [_cabi.comm_unique_id() if rank == 0 else None]   # one id per communicator
      dist.broadcast_object_list(ids, src=0)
      s = _cabi.DeviceState(n, 0, rank, rank=rank, nranks=world, comm_id=ids[0])
      s.set_fusion(fusion)
      mode = s.exchange_mode()
      for name, stream in _streams(n).items():
        psi0 = random_state(n, 21)
        ex_before = s.counters()["exchanges"]
        s.set_basis(0)                       # resets the bit permutation
        s.copy_in(psi0[rank << nl:(rank + 1) << nl])
        s.xg_apply_gates(_cabi.pack_xg_gates(stream))
        want = oracle.c_run(psi0.copy(), n, stream)
        # collective readouts under whatever permutation the exchanges left behind
        n2 = s.norm2()
        idx, p = s.argmax()
        amp = s.amplitude(12345 % (1 << n))
        pb = [s.prob_bit(b) for b in (0, nl - 1, n - 1)]
        lay = s.layout()
        ex = s.counters()["exchanges"]
        shard = s.copy_out()                 # undoes the bit remap (collective) before copying
        assert s.layout()["perm"] == list(range(n))
        err = float(np.abs(shard - want[rank << nl:(rank + 1) << nl]).max())
        wi = int(np.argmax(np.abs(want) ** 2))
        ii = np.arange(1 << n)
        ok = (abs(n2 - 1.0) < 1e-12 and idx == wi and abs(p - abs(want[wi]) ** 2) < 1e-15 and
              abs(amp - want[12345 % (1 << n)]) < 1e-12 and
              all(abs(v - float(np.sum(np.abs(want[(ii >> b) & 1 == 1]) ** 2))) < 1e-12
                  for v, b in zip(pb, (0, nl - 1, n - 1))))
        ex0 = s.counters()["exchanges"]
        if name == "qft":
          # three more QFTs queued into ONE flush: events hoisted to pass boundaries across the repetitions
          s.copy_in(psi0[rank << nl:(rank + 1) << nl])
          for _ in range(3):
            s.xg_apply_gates(_cabi.pack_xg_gates(stream))
          w3 = psi0.copy()
          for _ in range(3):
            w3 = oracle.c_run(w3, n, stream)
          err = max(err, float(np.abs(s.copy_out() - w3[rank << nl:(rank + 1) << nl]).max()))
        with open(os.path.join(out_dir, f"{name}_{int(fusion)}_{rank}.txt"), "w") as f:
          f.write(f"{err} {int(ok)} {ex - ex_before} {mode} {lay['perm']}")
      s.close()
    dist.barrier()
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("exchange", ["push", "swap", "nccl"])
def test_sharded_state_matches_oracle(world, exchange, tmp_path, monkeypatch):
  """QCC_B200_EXCHANGE: "push" (default) -- double-buffered state, an exchange event of any number of
  (sharded bit, local bit) pairs is ONE all-to-all written through CUDA IPC peer mappings by the store
  stage of the fused pass before it; "swap" -- one in-place kernel per pair over the same mappings;
  "nccl" -- ncclSend/ncclRecv per pair (also what the other two fall back to, on every rank alike, where
  the mappings are unavailable)."""
  if _ngpus() < world:
    pytest.skip(f"needs {world} GPUs")
  import torch.multiprocessing as mp
  monkeypatch.setenv("QCC_B200_EXCHANGE", exchange)   # inherited by the workers
  n = 18
  mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
  for name in ("qft", "random", "larose"):
    for fusion in (1, 0):
      for r in range(world):
        err, ok, ex, mode, perm = open(tmp_path / f"{name}_{fusion}_{r}.txt").read().split(" ", 4)
        assert float(err) <= 1e-12 and ok == "1", (name, fusion, r, err, ok, perm)
        assert mode == exchange, f"asked for the {exchange} exchange, the ranks agreed on {mode}"
      if name == "qft":   # every sharded qubit comes in exactly once: ONE event when the push exchange sees the
        # whole stream, else one exchange per sharded qubit
        assert int(ex) == (1 if exchange == "push" and fusion else int(math.log2(world)))
