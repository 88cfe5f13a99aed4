"""CPU, world_size 2 and 4 over gloo: the multi-GPU path's host logic.  Each process asks the
library for ITS rank's lowering of a circuit (qb_shard_lower_json: local gate batches, exchange
steps, bit permutation), executes it on a numpy shard -- local gates with the index-bit
reference, exchanges with torch.distributed send/recv exactly as engine.cu's do_exchange lays
them out -- and the gathered shards must equal the oracle's full-state result."""
import json
import math
import os
import socket
import sys

import numpy as np
import pytest

from helpers import ROOT, apply_masked, oracle, random_state, run_bits


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _circuits(n):
  H, X, V = oracle.GATES["h"], oracle.GATES["x"], oracle.GATES["v"]
  qft = []
  for i in reversed(range(n)):
    qft.append((0, n - 1 - i, H))
    for j in reversed(range(i)):
      qft.append((1 << (n - 1 - i), n - 1 - j, oracle.u1(math.pi / 2 ** (i - j))))
  rng = np.random.default_rng(11)
  names = list(oracle.GATES)
  rnd = []
  for _ in range(120):
    r = rng.random()
    m = oracle.u1(float(rng.uniform(-3, 3))) if r < 0.2 else (
        oracle.rotation([0, 0, 1.0], float(rng.uniform(-3, 3))) if r < 0.3 else oracle.GATES[names[rng.integers(len(names))]])
    b = [int(x) for x in rng.permutation(n)[:3]]
    nctl = int(rng.choice([0, 0, 1, 1, 2]))
    mask = 0
    for c in b[1:1 + nctl]:
      mask |= 1 << c
    rnd.append((mask, b[0], m))
  larose = []
  for bit in range(n):
    larose += [(0, n - 1 - bit, H), (0, n - 1 - bit, V)]
    if bit:
      larose.append((1 << (n - 1 - bit), n - 1, X))
  # grover.py-style x layers on qubits that serve as controls in between (they stay sharded: the x is a
  # rank relabel, no exchange), plus h on them now and then (exchange; the flip is undone by a local x)
  xh = []
  rng = np.random.default_rng(23)
  top = [n - 1, n - 2, n - 3]
  for rep in range(6):
    for q in top:
      if rng.random() < 0.7:
        xh.append((0, q, X))
    for _ in range(6):
      t = int(rng.integers(0, n - 3))
      ctl = 0
      for q in top:
        if rng.random() < 0.6:
          ctl |= 1 << q
      xh.append((ctl, t, oracle.GATES[names[rng.integers(len(names))]]))
      xh.append((1 << top[int(rng.integers(3))], t, oracle.u1(float(rng.uniform(-3, 3)))))
      xh.append((0, top[int(rng.integers(3))], oracle.u1(float(rng.uniform(-3, 3)))))   # diagonal on a sharded qubit
    if rep % 2:
      xh.append((0, top[rep % 3], H))
    xh.append((1 << 0, top[(rep + 1) % 3], X))   # cx ONTO a sharded qubit: a real exchange
  for q in range(n):                             # whatever ends up sharded is left relabelled ...
    xh.append((0, q, X))
  for q in range(n):                             # ... and read through the flip by these
    xh.append((1 << q, (q + 1) % n, oracle.u1(0.1 * (q + 1))))
    xh.append((0, q, oracle.GATES["t"]))
  return {"qft": qft, "random": rnd, "larose": larose, "xlayers": xh}


def _worker(rank, world, port, n, out_dir):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "tests"))
  import torch
  import torch.distributed as dist
  from qcc_b200 import _cabi
  dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
  try:
    p = int(math.log2(world))
    nl = n - p
    for name, gates in _circuits(n).items():
      for canon in (True, False):
        plan = json.loads(_cabi.shard_lower_json(n, world, rank, gates, canonicalize=canon))
        assert plan["nl"] == nl and plan["rank"] == rank
        psi0 = random_state(n, 5)
        shard = psi0[rank << nl:(rank + 1) << nl].copy()
        retired = 0
        for st in plan["steps"]:
          if st["kind"] == 0:
            retired += st["retired"]
            for g in st["gates"]:
              m = np.array([complex(g["m"][2 * i], g["m"][2 * i + 1]) for i in range(4)])
              assert g["ctl_mask"] < (1 << nl) and g["target"] < nl
              apply_masked(shard, nl, g["ctl_mask"], g["target"], m)
            continue
          assert len({k for k, _, _ in st["pairs"]}) == len(st["pairs"])
          assert len({v for _, v, _ in st["pairs"]} | {h for _, _, h in st["pairs"]}) == \
              len(st["pairs"]) + sum(1 for _, v, h in st["pairs"] if h != v)
          for k, v, h in st["pairs"]:   # the pairs of an event are disjoint: one after the other == all at once
            b = (rank >> k) & 1
            partner = rank ^ (1 << k)
            sel = 0 if b else 1
            run, nruns = 1 << v, 1 << (nl - 1 - v)
            view = shard.reshape(nruns, 2, run)
            send = torch.from_numpy(np.ascontiguousarray(view[:, sel, :]).view(np.float64))
            recv = torch.empty_like(send)
            reqs = [dist.isend(send, partner), dist.irecv(recv, partner)]
            for r in reqs:
              r.wait()
            view[:, sel, :] = recv.numpy().view(np.complex128).reshape(nruns, run)
            if h != v:                  # landing bit: the arrived qubit moves on to bit h, bit h's occupant down to v
              idx = np.arange(1 << nl)
              bv, bh = (idx >> v) & 1, (idx >> h) & 1
              shard = shard[idx ^ ((bv ^ bh) << v) ^ ((bv ^ bh) << h)]
        if not canon:
          assert retired == len(gates)
        # gather shards on rank 0 and undo the bit permutation
        t = torch.from_numpy(shard.view(np.float64).copy())
        parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, parts, dst=0)
        if rank == 0:
          phys = np.concatenate([x.numpy().view(np.complex128) for x in parts])
          perm = plan["perm"]
          flip = plan["flip"]            # relabelled rank bits carry the negated qubit
          if canon:
            assert perm == list(range(n)) and flip == 0
          idx = np.arange(1 << n)
          pidx = np.zeros_like(idx)
          for bl in range(n):
            f = (flip >> (perm[bl] - nl)) & 1 if perm[bl] >= nl else 0
            pidx |= (((idx >> bl) & 1) ^ f) << perm[bl]
          got = phys[pidx]
          want = run_bits(psi0.copy(), n, gates)
          err = float(np.abs(got - want).max())
          with open(os.path.join(out_dir, f"{name}_{int(canon)}.txt"), "w") as f:
            f.write(f"{err} {sum(len(s['pairs']) for s in plan['steps'] if s['kind'] == 1)} {flip} "
                    f"{sum(1 for s in plan['steps'] if s['kind'] == 1)}")
    dist.barrier()
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,window,hoist,prefetch,land",
                         [(2, 9, None, 0, 0, 0), (4, 10, None, 0, 0, 0), (4, 10, 30, 0, 0, 0), (4, 10, 30, 1, 0, 0),
                          (2, 12, 30, 1, 0, 0), (4, 10, 7, 1, 1, 0), (8, 12, 9, 1, 1, 1), (4, 13, 10, 0, 1, 1),
                          (2, 11, 8, 1, 1, 1)])
def test_sharded_lowering_over_gloo(world, n, window, hoist, prefetch, land, tmp_path, monkeypatch):
  """window: the victim window of the peer-memory exchanges (any local bit above the lowest few),
  QCC_B200_VICTIM_WINDOW; hoist = 1: exchanges moved back to pass boundaries (QCC_B200_HOIST), gates in
  between lowered again; prefetch = 1: multi-bit events (QCC_B200_PREFETCH), the push exchange's setting."""
  import torch.multiprocessing as mp
  if window:
    monkeypatch.setenv("QCC_B200_VICTIM_WINDOW", str(window))
  if hoist:
    monkeypatch.setenv("QCC_B200_HOIST", "1")
  if prefetch:
    monkeypatch.setenv("QCC_B200_PREFETCH", "1")
  if land:
    monkeypatch.setenv("QCC_B200_LAND", "1")   # arriving qubits land on the highest local bits (3-cycles)
  port = _free_port()
  mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
  flips_seen = 0
  for name in ("qft", "random", "larose", "xlayers"):
    for canon in (1, 0):
      err, nex, flip, nev = open(tmp_path / f"{name}_{canon}.txt").read().split()
      assert float(err) <= 1e-12, (name, canon, err)
      flips_seen += int(flip) != 0
  assert flips_seen > 0   # the x layers leave relabelled rank bits behind (uncanonicalised run)
  # QFT touches each sharded bit as a non-diagonal target exactly once: one exchanged pair per global bit,
  # and with prefetch all of them travel in ONE event
  assert int(open(tmp_path / "qft_0.txt").read().split()[1]) == int(math.log2(world))
  if prefetch:
    assert int(open(tmp_path / "qft_0.txt").read().split()[3]) == 1


def test_lowering_is_identical_in_structure_on_every_rank():
  from qcc_b200 import _cabi
  n, world = 10, 4
  gates = _circuits(n)["random"]
  plans = [json.loads(_cabi.shard_lower_json(n, world, r, gates, canonicalize=True)) for r in range(world)]
  shape = lambda p: [(s["kind"], s.get("pairs")) for s in p["steps"]]
  assert all(shape(p) == shape(plans[0]) for p in plans)
  assert all(p["perm"] == plans[0]["perm"] and p["flip"] == plans[0]["flip"] for p in plans)


def test_hoisted_exchange_keeps_sharded_qft_at_three_passes(monkeypatch):
  """QFT-30 over 2 ranks (29 local bits), victim window of the peer-swap exchange + hoisting: the one
  exchange sits on a pass boundary, so both segments plan into whole passes -- 3 in total, the
  single-GPU count -- instead of 3 + a fragment."""
  from qcc_b200 import _cabi
  n, world = 30, 2
  gates = _circuits(n)["qft"]

  def passes(env):
    for k, v in env.items():
      monkeypatch.setenv(k, v)
    plan = json.loads(_cabi.shard_lower_json(n, world, 1, gates, canonicalize=False))
    for k in env:
      monkeypatch.delenv(k)
    total, nex = 0, 0
    for st in plan["steps"]:
      if st["kind"] == 1:
        nex += 1
        continue
      g = [(x["ctl_mask"], x["target"], np.array([complex(x["m"][2 * i], x["m"][2 * i + 1]) for i in range(4)]))
           for x in st["gates"]]
      total += len(json.loads(_cabi.plan_json(n - 1, g, 12))["passes"])
    return total, nex

  assert passes({}) == (4, 1)
  assert passes({"QCC_B200_VICTIM_WINDOW": "30", "QCC_B200_HOIST": "1"}) == (3, 1)


def test_push_event_destinations_equal_pairwise_exchanges():
  """The push exchange (engine.cu event_map / kernels.h push_apply, through qb_shard_event_dest) writes every
  amplitude of every rank straight to its place after the event.  That must be the permutation the pairs
  of the event produce when exchanged one after the other the send/recv way, and a bijection."""
  from qcc_b200 import _cabi
  rng = np.random.default_rng(7)
  for world, nl, pairs in [(2, 6, [(0, 4)]), (4, 7, [(1, 5)]), (4, 7, [(0, 3), (1, 6)]), (8, 8, [(2, 7), (0, 4)]),
                           (8, 9, [(1, 3), (2, 8), (0, 5)]), (2, 7, [(0, 3, 6)]), (8, 10, [(1, 3, 9), (2, 4, 8), (0, 7, 7)]),
                           (4, 9, [(0, 5, 8), (1, 3, 7)])]:
    full = rng.normal(size=world << nl) + 1j * rng.normal(size=world << nl)
    shards = [full[r << nl:(r + 1) << nl].copy() for r in range(world)]
    want = [x.copy() for x in shards]
    for pr in pairs:
      k, v = pr[0], pr[1]
      h = pr[2] if len(pr) > 2 else v
      nxt = [x.copy() for x in want]
      run, nruns = 1 << v, 1 << (nl - 1 - v)
      for r in range(world):
        sel = 0 if (r >> k) & 1 else 1
        nxt[r].reshape(nruns, 2, run)[:, sel, :] = want[r ^ (1 << k)].reshape(nruns, 2, run)[:, 1 - sel, :]
      if h != v:    # landing bit: then local bits v and h trade places on every rank
        idx = np.arange(1 << nl)
        bv, bh = (idx >> v) & 1, (idx >> h) & 1
        sw = idx ^ ((bv ^ bh) << v) ^ ((bv ^ bh) << h)
        nxt = [x[sw] for x in nxt]
      want = nxt
    got = np.full(world << nl, np.nan + 0j)
    for r in range(world):
      dest = _cabi.shard_event_dest(nl, world, r, pairs, np.arange(1 << nl))
      assert np.isnan(got[dest]).all()
      got[dest] = shards[r]
    assert np.array_equal(got, np.concatenate(want)), (world, nl, pairs)


def test_pair_swap_index_math_equals_send_recv_layout():
  """k_pair_swap (kernels.cu) swaps local[(h, sel, w)] with peer[(h, 1 - sel, w)] for the flattened element
  numbers k = h * 2^victim + w of ONE half of [0, 2^(nl-1)) per rank.  Restated in numpy, both ranks'
  swaps together must leave exactly what do_exchange's send/recv + copy-back leaves (the layout the gloo
  test above executes)."""
  rng = np.random.default_rng(3)
  for nl in (2, 3, 5, 7):
    for victim in range(nl):
      sh = [rng.normal(size=1 << nl) + 1j * rng.normal(size=1 << nl) for _ in range(2)]
      run, nruns = 1 << victim, 1 << (nl - 1 - victim)
      want = [x.copy() for x in sh]
      for r in (0, 1):                       # send/recv: rank r gives away the half with victim bit != its rank bit
        sel = 0 if r else 1
        want[r].reshape(nruns, 2, run)[:, sel, :] = sh[1 - r].reshape(nruns, 2, run)[:, 1 - sel, :]
      got = [x.copy() for x in sh]
      half, low = 1 << (nl - 1), (1 << victim) - 1
      for r in (0, 1):                       # the kernel: rank bit 0 takes k < half / 2, rank bit 1 the rest
        sel = 0 if r else 1
        k = np.arange(half // 2, half) if r else np.arange(0, half // 2)
        i0 = ((k >> victim) << (victim + 1)) | (k & low)
        l, p = i0 | (sel << victim), i0 | ((sel ^ 1) << victim)
        a, b = got[r][l].copy(), got[1 - r][p].copy()
        got[r][l], got[1 - r][p] = b, a
      assert all(np.array_equal(got[r], want[r]) for r in (0, 1)), (nl, victim)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_lowering_fuzz_all_ranks_in_one_process(seed, monkeypatch):
  """Random circuits (0-2 controls, diagonal / permuting / general gates) x random lowering parameters (victim
  window, hoisting, prefetch = multi-bit events, landing bits) x world 2 / 4 / 8, with and without the final
  canonicalisation: the plans of ALL ranks are executed side by side on numpy shards (events as the pairwise
  half-shard trades they stand for, then the landing-bit swap) and the un-permuted result must be the oracle's."""
  rng = np.random.default_rng(seed)
  names = list(oracle.GATES)
  from qcc_b200 import _cabi
  for trial in range(14):
    world = int(rng.choice([2, 4, 8]))
    p = int(math.log2(world))
    n = int(rng.integers(p + 6, 13))
    nl = n - p
    monkeypatch.setenv("QCC_B200_VICTIM_WINDOW", str(int(rng.integers(1, nl - 2))))
    monkeypatch.setenv("QCC_B200_HOIST", str(int(rng.integers(0, 2))))
    monkeypatch.setenv("QCC_B200_PREFETCH", str(int(rng.integers(0, 2))))
    monkeypatch.setenv("QCC_B200_LAND", str(int(rng.integers(0, 2))))
    gates = []
    for _ in range(int(rng.integers(20, 140))):
      r = rng.random()
      m = oracle.u1(float(rng.uniform(-3, 3))) if r < 0.25 else (
          oracle.rotation([0, 0, 1.0], float(rng.uniform(-3, 3))) if r < 0.3 else oracle.GATES[names[rng.integers(len(names))]])
      b = [int(x) for x in rng.permutation(n)[:3]]
      mask = 0
      for c in b[1:1 + int(rng.choice([0, 0, 1, 1, 2]))]:
        mask |= 1 << c
      gates.append((mask, b[0], m))
    psi0 = random_state(n, 100 * seed + trial)
    plans = [json.loads(_cabi.shard_lower_json(n, world, r, gates, canonicalize=bool(trial % 2))) for r in range(world)]
    shards = [psi0[r << nl:(r + 1) << nl].copy() for r in range(world)]
    assert len({len(pl["steps"]) for pl in plans}) == 1
    for si, st0 in enumerate(plans[0]["steps"]):
      if st0["kind"] == 0:
        for r in range(world):
          for g in plans[r]["steps"][si]["gates"]:
            m = np.array([complex(g["m"][2 * i], g["m"][2 * i + 1]) for i in range(4)])
            apply_masked(shards[r], nl, g["ctl_mask"], g["target"], m)
        continue
      for k, v, h in st0["pairs"]:
        run, nruns = 1 << v, 1 << (nl - 1 - v)
        new = [x.copy() for x in shards]
        for r in range(world):
          sel = 0 if (r >> k) & 1 else 1
          new[r].reshape(nruns, 2, run)[:, sel, :] = shards[r ^ (1 << k)].reshape(nruns, 2, run)[:, 1 - sel, :]
        if h != v:
          idx = np.arange(1 << nl)
          bv, bh = (idx >> v) & 1, (idx >> h) & 1
          sw = idx ^ ((bv ^ bh) << v) ^ ((bv ^ bh) << h)
          new = [x[sw] for x in new]
        shards = new
    phys = np.concatenate(shards)
    perm, flip = plans[0]["perm"], plans[0]["flip"]
    idx = np.arange(1 << n)
    pidx = np.zeros_like(idx)
    for bl in range(n):
      f = (flip >> (perm[bl] - nl)) & 1 if perm[bl] >= nl else 0
      pidx |= (((idx >> bl) & 1) ^ f) << perm[bl]
    want = run_bits(psi0.copy(), n, gates)
    assert np.abs(phys[pidx] - want).max() <= 1e-12, (seed, trial, world, n)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_qft30_lowering_quality(world):
  """Regression guard for what the multi-GPU numbers rest on: ten QFT-30s queued into one flush
  lower, on every rank alike, into 3 passes per QFT, at most 1.5 exchange events per QFT, every event on the store
  stage of a fused pass, all log2(world) sharded qubits travelling together; the NCCL-mode lowering (narrow victim
  window, one pair per exchange, no hoisting) cuts more passes."""
  from qcc_b200 import _cabi
  n, reps = 30, 10
  gates = _circuits(n)["qft"] * reps
  arr = _cabi.pack_gates(gates)
  push = [_cabi.shard_plan_stats(n, world, r, arr, 12, "push") for r in (0, world - 1)]
  assert push[0] == push[1]
  st = push[0]
  assert st["passes"] == 3 * reps and st["fused_passes"] == st["passes"]
  assert st["events"] <= 1.5 * reps and st["events_on_a_pass"] == st["events"]
  assert st["pairs"] == st["events"] * int(math.log2(world))
  nccl = _cabi.shard_plan_stats(n, world, 0, arr, 12, "nccl")
  assert nccl["passes"] > st["passes"]
