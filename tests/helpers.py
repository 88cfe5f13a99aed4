"""Shared test helpers: golden loaders, index-bit reference apply, and a numpy interpreter of
the fusion planner's pass format (so the planner can be verified on CPU, where no kernel
can run).  Test infrastructure only -- may import oracle/, never imported by qcc_b200."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import oracle  # noqa: E402

K_U, K_PHASE, K_DIAG, K_PERM, K_LADDER, K_NOP, K_SWAP, K_ULADDER, K_PARSWAP = range(9)


def load_golden(name):
  return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def stream_of(z):
  """(kind, ctl, tgt, 2x2) tuples in python numbering from a golden npz."""
  return [(int(k), int(c), int(t), m.reshape(2, 2)) for k, c, t, m in
          zip(z["kind"], z["ctl"], z["tgt"], z["mats"])]


def xg_to_bits(n, stream):
  """python-numbered stream -> [(ctl_mask, target_bit, 2x2)] following xgates.cc:45-67
  (negative controls included).  Gates that act on nothing are dropped."""
  out = []
  for kind, ctl, tgt, m in stream:
    t = n - 1 - tgt
    if kind == 1:
      out.append((0, t, m))
      continue
    c = n - 1 - ctl
    if c < n:
      if c == t:
        continue
      out.append((1 << c, t, m))
    else:
      cc = c - n
      if cc <= t or cc >= n:
        continue
      out.append((1 << cc, t, m))
  return out


def apply_masked(psi, n, ctl_mask, target, m):
  """Index-bit reference: 2x2 on bit `target` where all ctl_mask bits are 1 (in place)."""
  m = np.asarray(m, dtype=np.complex128).reshape(4)
  idx = np.arange(1 << n, dtype=np.int64)
  sel = ((idx >> target) & 1 == 0) & ((idx & ctl_mask) == ctl_mask)
  i0 = idx[sel]
  i1 = i0 | (1 << target)
  a = psi[i0].copy()
  b = psi[i1].copy()
  psi[i0] = m[0] * a + m[1] * b
  psi[i1] = m[2] * a + m[3] * b


def run_bits(psi, n, gates):
  for mask, t, m in gates:
    apply_masked(psi, n, mask, t, m)
  return psi


def random_state(n, seed):
  rng = np.random.default_rng(seed)
  v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
  return (v / np.linalg.norm(v)).astype(np.complex128)


# ---------------------------------------------------------------------------------------
# numpy interpreter of qb_plan_json output: mirrors fused.cu step by step (tile gather,
# rounds, three-level predicates, ladder tables), vectorised over tiles and groups.
# ---------------------------------------------------------------------------------------
def interpret_plan(plan_json: str, n: int, gates, psi: np.ndarray) -> np.ndarray:
  plan = json.loads(plan_json)
  retired = 0
  for p in plan["passes"]:
    retired += p["ngates"]
    if p["single_gate"] >= 0:
      # the plan carries the (possibly pre-multiplied) gate itself
      m = np.array([complex(p["m"][2 * i], p["m"][2 * i + 1]) for i in range(4)])
      if not (m[0] == 1 and m[3] == 1 and m[1] == 0 and m[2] == 0):
        apply_masked(psi, n, p["ctl_mask"], p["target"], m)
      continue
    K = p["K"]
    tb = p["tile_bits"]
    assert tb[:3] == [0, 1, 2] and sorted(tb) == tb and len(tb) == K
    non_tile = [b for b in range(n) if b not in tb]
    ntiles = 1 << (n - K)
    tids = np.arange(ntiles, dtype=np.int64)
    base = np.zeros(ntiles, dtype=np.int64)
    for k, b in enumerate(non_tile):
      base |= ((tids >> k) & 1) << b
    j = np.arange(1 << K, dtype=np.int64)
    off = np.zeros(1 << K, dtype=np.int64)
    for k, b in enumerate(tb):
      off |= ((j >> k) & 1) << b
    gidx = base[:, None] | off[None, :]
    T = psi[gidx]                               # [tile, local]
    tables = np.array([complex(x, y) for x, y in p["tables"]], dtype=np.complex128)
    outph = np.array([complex(x, y) for x, y in p["outph"]], dtype=np.complex128)
    jbtab = np.array(p["jbtab"], dtype=np.int64).reshape(len(p["rounds"]), 1 << (K - 3))
    outbits = p["outbits"]
    for ri, R in enumerate(p["rounds"]):
      assert R["nbits"] == 3
      rbit = R["rbit"]
      qmap = R["qmap"]
      assert sorted(rbit + qmap) == list(range(K)), "round bits + qmap must partition the tile"
      ng = 1 << (K - 3)
      q = np.arange(ng, dtype=np.int64)
      jb = np.zeros(ng, dtype=np.int64)
      for k, lp in enumerate(qmap):
        jb |= ((q >> k) & 1) << lp
      # the kernel reads jb and its swizzled slot from the host-built table
      assert np.array_equal(jbtab[ri] & 0xFFFF, jb)
      fold = (jb >> 3) ^ (jb >> 6) ^ (jb >> 9) ^ (jb >> 12)
      assert np.array_equal(jbtab[ri] >> 16, jb ^ (fold & 7))
      spread = np.array([sum(((e >> k) & 1) << rbit[k] for k in range(3)) for e in range(8)])
      je = jb[:, None] | spread[None, :]        # [group, e]
      # rounds of one barrier-free run: warp w (groups with (q >> 5) & 7 == w) must own the same
      # amplitudes in this round and in the next one -- the kernel only syncs the warp between them
      if R.get("nobar"):
        assert ng >= 256 and ri + 1 < len(p["rounds"])
        Rn = p["rounds"][ri + 1]
        jbn = np.zeros(ng, dtype=np.int64)
        for k, lp in enumerate(Rn["qmap"]):
          jbn |= ((q >> k) & 1) << lp
        spn = np.array([sum(((e >> k) & 1) << Rn["rbit"][k] for k in range(3)) for e in range(8)])
        jen = jbn[:, None] | spn[None, :]
        for w in range(8):
          mine = ((q >> 5) & 7) == w
          assert np.array_equal(np.sort(je[mine].ravel()), np.sort(jen[mine].ravel())), "warp sub-cube changes inside a run"
      else:
        assert ri + 1 == len(p["rounds"]) or not R.get("nobar")
      # warp-private tile I/O: the amplitudes warp w copies in (ld_map) are exactly the ones it works on
      # in the first round, the ones it copies out (st_map) exactly those of the last round -- the kernel
      # has no CTA barrier between the copy and those rounds
      if p.get("warp_io") and ri in (0, len(p["rounds"]) - 1):
        assert ng >= 512
        for name, at in (("ld_map", 0), ("st_map", len(p["rounds"]) - 1)):
          if ri != at:
            continue
          cmap = p[name]
          assert sorted(cmap) == list(range(K)) and cmap[:3] == [0, 1, 2]
          c = np.arange(1 << K, dtype=np.int64)
          jc = np.zeros(1 << K, dtype=np.int64)
          for k, lp in enumerate(cmap):
            jc |= ((c >> k) & 1) << lp
          for w in range(8):
            mine = ((q >> 5) & 7) == w
            copied = np.sort(jc[((c >> 5) & 7) == w])
            assert np.array_equal(copied, np.sort(je[mine].ravel())), f"{name}: warp {w} copies what it does not own"
      elif not p.get("warp_io"):
        assert p.get("ld_map", list(range(K))) == list(range(K)) and p.get("st_map", list(range(K))) == list(range(K))
      # direct-store round: groups straight to HBM -- 8 lanes must still cover one 128-byte run
      for flag, at in (("st_direct", len(p["rounds"]) - 1),):
        if p.get(flag) and ri == at:
          assert p["warp_io"] and R["prog"] != 0
          assert min(rbit) >= 3 and sorted(qmap[:3]) == [0, 1, 2]
      A = T[:, je]                              # [tile, group, e]
      rops = p["ops"][R["op_begin"]:R["op_end"]]
      if R.get("prog") == 3:
        # UX program: uncontrolled U's and parity swaps only (the kernel's predicate-free interpreter)
        for o in rops:
          k8 = o["kind"] & 0xFF
          assert k8 in (K_U, K_PARSWAP, K_SWAP, K_PHASE, K_LADDER, K_ULADDER) and len(rops) >= 1
          masked_u = k8 == K_U and 15 <= (o["kind"] >> 24) < 18
          if k8 in (K_SWAP, K_PHASE) or masked_u:
            assert o["flags"] == sum(1 << e for e in range(8) if (e & o["rmask"]) == o["rwant"])
          if k8 == K_U and not masked_u:
            assert o["lmask"] == 0 and o["rmask"] == 0
            real = all(o["m"][2 * i + 1] == 0.0 for i in range(4))
            colimag = not real and o["m"][1] == 0 and o["m"][5] == 0 and o["m"][2] == 0 and o["m"][6] == 0
            assert (o["kind"] >> 24) == (29 + o["tpos"] if colimag else 9 + o["tpos"] + (3 if real else 0))
      for op in rops:
        if op["kind"] & 0xFF == K_PARSWAP:
          # swap the pair on tpos where parity(control bits) ^ flip is odd
          tp = op["tpos"]
          assert not (op["rmask"] >> tp) & 1 and (op["kind"] >> 24) == 26 + tp
          par_t = np.array([bin(int(b) & op["gmask"]).count("1") & 1 for b in base])
          par_g = np.array([bin(int(x) & op["lmask"]).count("1") & 1 for x in jb])
          par = par_t[:, None] ^ par_g[None, :] ^ (op["rwant"] & 1)
          for e in range(8):
            if e & (1 << tp):
              continue
            pe = bin(e & op["rmask"]).count("1") & 1
            assert (op["lwant"] >> e) & 1 == pe
            sw = (par ^ pe) == 1
            e1 = e | (1 << tp)
            x = A[:, :, e].copy()
            y = A[:, :, e1].copy()
            A[:, :, e] = np.where(sw, y, x)
            A[:, :, e1] = np.where(sw, x, y)
          continue
        tile_ok = (base & op["gmask"]) == op["gwant"]
        grp_ok = (jb & op["lmask"]) == op["lwant"]
        ok = tile_ok[:, None] & grp_ok[None, :]                        # [tile, group]
        m = np.array(op["m"]).view(np.complex128) if False else np.array(
            [complex(op["m"][2 * i], op["m"][2 * i + 1]) for i in range(4)])
        kind = op["kind"] & 0xFF
        assert (op["kind"] >> 8) & 0xFF == op["tpos"] and (op["kind"] >> 16) & 0xFF == op["mflags"]
        if kind == K_ULADDER:
          assert op["F"][2 * (1 << op["tpos"])] == 1.0 and op["F"][2 * (1 << op["tpos"]) + 1] == 0.0
        if kind in (K_U, K_PERM, K_SWAP):
          tp = op["tpos"]
          for e in range(8):
            if e & (1 << tp) or (e & op["rmask"]) != op["rwant"]:
              continue
            e1 = e | (1 << tp)
            x = A[:, :, e].copy()
            y = A[:, :, e1].copy()
            if kind == K_U:
              nx, ny = m[0] * x + m[1] * y, m[2] * x + m[3] * y
            elif kind == K_PERM:
              nx, ny = m[1] * y, m[2] * x
            else:
              nx, ny = y, x
            A[:, :, e] = np.where(ok, nx, x)
            A[:, :, e1] = np.where(ok, ny, y)
        elif kind == K_PHASE:
          for e in range(8):
            if (e & op["rmask"]) == op["rwant"]:
              A[:, :, e] = np.where(ok, m[0] * A[:, :, e], A[:, :, e])
        elif kind in (K_LADDER, K_ULADDER):
          if kind == K_ULADDER:
            # uncontrolled butterfly on the pivot first, then the ladder on the pivot-set half
            tp = op["tpos"]
            assert op["rmask"] == 0 and op["lmask"] == 0 and op["gmask"] == 0
            for e in range(8):
              if e & (1 << tp):
                continue
              e1 = e | (1 << tp)
              x = A[:, :, e].copy()
              y = A[:, :, e1].copy()
              A[:, :, e] = m[0] * x + m[1] * y
              A[:, :, e1] = m[2] * x + m[3] * y
            lad_rmask, lad_rwant = 1 << tp, 1 << tp
          else:
            lad_rmask, lad_rwant = op["rmask"], op["rwant"]
          t0 = op["table_off"]
          # per-tile constant: three tables over the fields of the tile number (planner.cc build_ladder_tables)
          cbase = op["outph_off"]
          w0, w1, w2 = p["lad_w"]
          assert w0 + w1 + w2 == n - K
          pout = (outph[cbase + (tids & ((1 << w0) - 1))] *
                  outph[cbase + (1 << w0) + ((tids >> w0) & ((1 << w1) - 1))] *
                  outph[cbase + (1 << w0) + (1 << w1) + (tids >> (w0 + w1))])
          assert all(b not in tb for b in outbits[op["out_off"]:op["out_off"] + op["nout"]])
          # tables are indexed by the group number: T_a[q & 31] (the lane), T_b[q >> 5]
          c = pout[:, None] * tables[t0 + (q & 31)][None, :] * tables[t0 + 32 + (q >> 5)][None, :]
          F = np.array([complex(op["F"][2 * e], op["F"][2 * e + 1]) for e in range(8)])
          for e in range(8):
            if (e & lad_rmask) == lad_rwant:
              A[:, :, e] = np.where(ok, c * F[e] * A[:, :, e], A[:, :, e])
        else:
          raise AssertionError(f"unexpected op kind {kind}")
      T[:, je] = A
    psi[gidx] = T
  assert retired == len(gates), f"plan retires {retired} of {len(gates)} gates"
  return psi


def plan_summary(plan_json: str):
  plan = json.loads(plan_json)
  fused = [p for p in plan["passes"] if p["single_gate"] < 0]
  return {
      "passes": len(plan["passes"]),
      "fused": len(fused),
      "singles": len(plan["passes"]) - len(fused),
      "rounds": sum(len(p["rounds"]) for p in fused),
      "ops": sum(len(p["ops"]) for p in fused),
      "ladders": sum(1 for p in fused for o in p["ops"] if o["kind"] & 0xFF in (K_LADDER, K_ULADDER)),
      "progs": [R["prog"] for p in fused for R in p["rounds"]],
  }
