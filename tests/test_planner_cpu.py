"""CPU: the C-ABI library loads and exports everything include/qcc_b200.h declares, fails
loudly without a GPU, and its fusion planner (pure host code) produces plans that compute
the same state as gate-by-gate application -- checked by interpreting the plan with numpy
(tests/helpers.py), since no kernel can run here."""
import ctypes
import json
import math
import os
import re

import numpy as np
import pytest

from helpers import (GOLDEN, ROOT, interpret_plan, load_golden, oracle, plan_summary, random_state,
                     run_bits, stream_of, xg_to_bits)
from qcc_b200 import _cabi


def test_library_exports_every_declared_symbol():
  hdr = open(os.path.join(ROOT, "include", "qcc_b200.h")).read()
  declared = set(re.findall(r"^(?:int|const char \*)\s*\*?(qb_\w+)\(", hdr, re.M))
  assert len(declared) >= 30
  L = _cabi.lib()
  for name in declared:
    assert hasattr(L, name), f"{name} declared in qcc_b200.h but not exported"
  assert declared == set(_cabi.PROTOTYPES), "ctypes prototypes out of sync with the header"
  assert L.qb_abi_version() == 2


def test_no_gpu_means_loud_failure(has_gpu):
  if has_gpu:
    pytest.skip("GPU present")
  with pytest.raises(_cabi.QbError) as e:
    _cabi.DeviceState(4)
  assert "CUDA" in str(e.value) or "cuda" in str(e.value) or "device" in str(e.value)
  psi = np.zeros(4, dtype=np.complex128)
  g = np.eye(2, dtype=np.complex128).reshape(4) * 1j
  rc = _cabi.lib().qb_host_apply1(psi.ctypes.data, g.ctypes.data, 2, 0, 128, -1)
  assert rc < 0


def test_argument_errors_do_not_need_a_gpu():
  L = _cabi.lib()
  need = ctypes.c_size_t()
  bad = _cabi.pack_gates([(1 << 3, 3, np.eye(2))])  # control == target
  assert L.qb_plan_json(6, bad, 1, 12, None, 0, ctypes.byref(need)) == -1
  assert b"bad bits" in L.qb_last_error()
  assert L.qb_state_destroy(None) == 0


CIRCS = sorted(f for f in os.listdir(GOLDEN) if f.startswith("circ_") and f.endswith(".npz"))


@pytest.mark.parametrize("name", CIRCS + ["dense_n8.npz", "dense_n10.npz", "dense_n12.npz", "acceleration_2.npz"])
@pytest.mark.parametrize("tile_bits", [4, 7, 12])
def test_plan_reproduces_golden(name, tile_bits):
  z = load_golden(name)
  n = int(z["nbits"])
  if n < 4:
    pytest.skip("fusion needs >= 4 qubits")
  gates = xg_to_bits(n, stream_of(z))
  pj = _cabi.plan_json(n, gates, tile_bits)
  got = interpret_plan(pj, n, gates, z["psi0"].astype(np.complex128).copy())
  ref = z["final_xgates"] if "final_xgates" in z.files else z["final_spec"]
  assert np.abs(got - ref).max() <= 1e-12


def _random_bit_gates(n, count, seed):
  rng = np.random.default_rng(seed)
  names = list(oracle.GATES)
  gates = []
  for _ in range(count):
    r = rng.random()
    if r < 0.15:
      m = oracle.u1(float(rng.uniform(-3, 3)))
    elif r < 0.25:
      m = oracle.rotation([0, 0, 1.0], float(rng.uniform(-3, 3)))     # DIAG (rz)
    elif r < 0.35:
      m = oracle.rotation([1.0, 0, 0], float(rng.uniform(-3, 3)))
    else:
      m = oracle.GATES[names[rng.integers(len(names))]]
    bits = [int(b) for b in rng.permutation(n)[:3]]
    nctl = int(rng.choice([0, 0, 1, 1, 2]))
    mask = 0
    for b in bits[1:1 + nctl]:
      mask |= 1 << b
    gates.append((mask, bits[0], m))
  return gates


@pytest.mark.parametrize("n,tile_bits,seed", [(4, 4, 1), (5, 4, 2), (9, 6, 3), (13, 12, 4), (14, 10, 5), (15, 13, 6)])
def test_plan_random_circuits(n, tile_bits, seed):
  gates = _random_bit_gates(n, 160, seed)
  psi0 = random_state(n, seed)
  want = run_bits(psi0.copy(), n, gates)
  pj = _cabi.plan_json(n, gates, tile_bits)
  got = interpret_plan(pj, n, gates, psi0.copy())
  assert np.abs(got - want).max() <= 1e-12


def _qft_bits(n):
  """circuit.py:320-326 in index bits: python qubit q == bit n-1-q."""
  g = []
  for i in reversed(range(n)):
    g.append((0, n - 1 - i, oracle.GATES["h"]))
    for j in reversed(range(i)):
      g.append((1 << (n - 1 - i), n - 1 - j, oracle.u1(math.pi / 2 ** (i - j))))
  return g


def test_qft30_plans_into_three_passes():
  """configs[2]: 30 h + 435 cu1 -> 3 HBM sweeps with every ladder fused."""
  gates = _qft_bits(30)
  assert len(gates) == 465
  s = plan_summary(_cabi.plan_json(30, gates, 12))
  assert s["fused"] == 3 and s["singles"] == 0
  # 28 h+ladder; the h + single cu1 and the bare h of the tail get a one-partner / empty ladder so
  # that every round is the unrolled Hadamard+ladder program (QB_PROG_HL3U == 2)
  assert s["ladders"] == 30 and s["ops"] == 30
  plan = json.loads(_cabi.plan_json(30, gates, 12))
  assert sum(p["ngates"] for p in plan["passes"]) == 465
  assert [len(p["rounds"]) for p in plan["passes"]] == [4, 3, 3]
  assert all(R["prog"] == 2 for p in plan["passes"] for R in p["rounds"])
  # rounds that share a per-warp sub-cube need no CTA barrier between them.  Every sub-cube also holds
  # tile bits 0..2 (whole 128-byte runs, so the warp itself copies it in and out): 9 round bits in the
  # first pass (whose first round is on bits 0..2), 6 in the others -- one CTA barrier per tile.
  assert [[R["nobar"] for R in p["rounds"]] for p in plan["passes"]] == [[1, 1, 0, 0], [1, 0, 0], [1, 0, 0]]
  assert all(p["warp_io"] == 1 for p in plan["passes"])


def test_larose_plan_is_much_shorter_than_the_gate_list():
  """configs[1] gate stream (larose_benchmark.py:47-54), 28 qubits depth 2."""
  n = 28
  gates = []
  for _ in range(2):
    for bit in range(n):
      b = n - 1 - bit
      gates.append((0, b, oracle.GATES["h"]))
      gates.append((0, b, oracle.GATES["v"]))
      if bit > 0:
        gates.append((1 << b, n - 1, oracle.GATES["x"]))
  s = plan_summary(_cabi.plan_json(n, gates, 12))
  assert s["passes"] <= 10 and s["passes"] * 10 < len(gates)


def _larose_bits(n, depth):
  gates = []
  for _ in range(depth):
    for bit in range(n):
      b = n - 1 - bit
      gates.append((0, b, oracle.GATES["h"]))
      gates.append((0, b, oracle.GATES["v"]))
      if bit > 0:
        gates.append((1 << b, n - 1, oracle.GATES["x"]))
  return gates


def test_larose_rounds_are_scheduled_into_ux_programs():
  """configs[1]: the scheduler collects the h.v butterflies three to a round and turns the cx fan-in
  onto qubit 0 (larose_benchmark.py:52-53) into one parity swap per pass, so every round runs the
  predicate-free UX program: 10 rounds per depth instead of the 14 of the in-order cut."""
  from helpers import K_PARSWAP
  n = 28
  plan = json.loads(_cabi.plan_json(n, _larose_bits(n, 2), 12))
  fused = [p for p in plan["passes"] if p["single_gate"] < 0]
  assert len(fused) == 6 and len(plan["passes"]) == 6
  assert sum(len(p["rounds"]) for p in fused) <= 20
  assert all(R["prog"] == 3 for p in fused for R in p["rounds"])
  for p in fused:
    assert sum(1 for o in p["ops"] if o["kind"] & 0xFF == K_PARSWAP) == 1


@pytest.mark.parametrize("n,tile_bits,seed", [(6, 5, 11), (10, 7, 12), (13, 12, 13), (14, 11, 14)])
def test_plan_cx_fan_in_and_scheduling(n, tile_bits, seed):
  """x / cx chains onto shared targets mixed with 1-qubit gates and controlled phases: exercises the
  commutation DAG, PARSWAP merging (including x's that cancel) and the UX round program."""
  from helpers import K_PARSWAP
  rng = np.random.default_rng(seed)
  names = ["h", "v", "yroot", "t", "x", "z", "s"]
  gates = []
  for _ in range(220):
    r = rng.random()
    t = int(rng.integers(n))
    if r < 0.45:
      hot = int(rng.integers(2))                      # two popular targets
      tgt = n - 1 - hot
      c = int(rng.integers(n))
      if c == tgt or rng.random() < 0.1:
        gates.append((0, tgt, oracle.GATES["x"]))
      else:
        gates.append((1 << c, tgt, oracle.GATES["x"]))
    elif r < 0.55:
      c = int((t + 1 + rng.integers(n - 1)) % n)
      gates.append((1 << c, t, oracle.u1(float(rng.uniform(-3, 3)))))
    else:
      gates.append((0, t, oracle.GATES[names[rng.integers(len(names))]]))
  psi0 = random_state(n, seed)
  want = run_bits(psi0.copy(), n, gates)
  pj = _cabi.plan_json(n, gates, tile_bits)
  got = interpret_plan(pj, n, gates, psi0.copy())
  assert np.abs(got - want).max() <= 1e-12
  plan = json.loads(pj)
  assert any(o["kind"] & 0xFF == K_PARSWAP for p in plan["passes"] if p["single_gate"] < 0 for o in p["ops"])


def test_round_bank_classes():
  """QbRound::qmap gives group-index bits 0..2 one tile-local bit of each class (b % 3),
  which is what makes the swizzled shared-memory accesses of fused.cu conflict free."""
  gates = _random_bit_gates(14, 200, 9)
  plan = json.loads(_cabi.plan_json(14, gates, 12))
  seen = 0
  for p in plan["passes"]:
    if p["single_gate"] >= 0:
      continue
    for R in p["rounds"]:
      assert sorted(q % 3 for q in R["qmap"][:3]) == [0, 1, 2]
      seen += 1
  assert seen > 5


def test_sleator_weinfurter_runs_become_one_doubly_controlled_gate():
  """qb_flush's peephole (planner.cc fuse_ccu_runs, through qb_fuse_gates): the five gates circuit.py:227-246
  emits for ccu(a, b, t, U) -- cu(a,t,V) cx(a,b) cu(b,t,V^dagger) cx(a,b) cu(b,t,V), V = sqrt(U) -- are replaced
  by ONE gate on t controlled by a AND b (+ four identities); anything that merely looks similar is left alone.
  Checked by running both gate lists on a random state with the index-bit reference."""
  from scipy.linalg import sqrtm
  from helpers import oracle, random_state, run_bits
  n = 7
  X, H = oracle.GATES["x"], oracle.GATES["h"]

  def ccu(a, b, t, u):
    v = np.asarray(sqrtm(np.asarray(u, dtype=np.complex128)))
    return [(1 << a, t, v), (1 << a, b, X), (1 << b, t, v.conj().T), (1 << a, b, X), (1 << b, t, v)]

  gates = [(0, q, H) for q in range(n)]
  gates += ccu(0, 1, 2, X)                                   # Toffoli
  gates += ccu(5, 3, 6, oracle.GATES["z"])                   # ccz
  gates += ccu(4, 6, 0, oracle.u1(0.3))                      # ccu1
  gates += [(0, 3, oracle.GATES["v"])]
  gates += ccu(2, 4, 1, oracle.GATES["y"])
  broken = ccu(1, 2, 3, X)
  broken[2] = (broken[2][0], broken[2][1], broken[0][2])     # V instead of V^dagger: NOT a ccx
  gates += broken
  wrong_ctl = ccu(0, 5, 4, X)
  wrong_ctl[3] = (1 << 6, 5, X)                              # second cx from another control
  gates += wrong_ctl
  fused, nruns = _cabi.fuse_gates(gates)
  assert nruns == 4 and len(fused) == len(gates)
  two_ctl = [g for g in fused if bin(g[0]).count("1") == 2]
  assert len(two_ctl) == 4
  # the Toffoli came out as the exact permutation matrix (so it runs as a pure swap)
  assert np.array_equal(two_ctl[0][2], np.array([[0, 1], [1, 0]], dtype=np.complex128))
  psi0 = random_state(n, 9)
  want = run_bits(psi0.copy(), n, gates)
  got = run_bits(psi0.copy(), n, fused)
  assert np.abs(got - want).max() < 1e-13
  # and the planner's output for the fused list still matches the oracle
  from helpers import interpret_plan
  out = interpret_plan(_cabi.plan_json(n, fused, 6), n, fused, psi0.copy())
  assert np.abs(out - want).max() < 1e-12


def test_every_plan_fits_the_kernel_parameter_block():
  """qb_plan_check = the planner + the HOST half of launch_fused_pass (capacity of the 32 KiB parameter block:
  ops, rounds, ladder ops; shared memory for the tile + ladder tables).  Round 2 shipped a parameter array with
  room for 12 ladder ops per pass for a few hours -- order finding plans 14 -- and only a GPU run noticed; this
  runs the check where no GPU exists, over the streams the GPU suite and the bench execute."""
  import glob
  from helpers import GOLDEN, load_golden, stream_of, xg_to_bits
  from qcc_b200 import circuit, workloads
  streams = []
  for path in sorted(glob.glob(os.path.join(GOLDEN, "circ_*.npz"))) + [os.path.join(GOLDEN, "order_N15_a4.npz")]:
    z = load_golden(os.path.basename(path))
    n = int(z["nbits"])
    if n >= 4:
      streams.append((os.path.basename(path), n, xg_to_bits(n, stream_of(z))))
  for n in (21, 30, 34):
    streams.append((f"qft{n}", n, xg_to_bits(n, workloads.qft(n) * 3)))
  streams.append(("larose28", 28, xg_to_bits(28, workloads.larose(28, 3))))
  streams.append(("supremacy34", 34, xg_to_bits(34, workloads.supremacy(34, 20, seed=0))))
  np.random.seed(0)
  qc, _ = workloads.grover_circuit(7, qc_factory=lambda name: circuit.qc(name, eager=False), iterations=3)
  g = []
  for node in qc.ir.gates:
    if node.is_single():
      g.append((1, 0, node.idx0, np.asarray(node.gate)))
    elif node.is_ctl():
      g.append((2, node.ctl, node.idx1, np.asarray(node.gate)))
  streams.append(("grover14", 14, xg_to_bits(14, g)))
  for name, n, gates in streams:
    for K in (12, 13, 9, 6):
      if K <= n:
        assert _cabi.plan_check(n, gates, K) >= 1, (name, K)
