"""CPU: pin the oracle (oracle/qcc_oracle.c + oracle/oracle.py) against the golden vectors
generated from the reference (tests/golden/make_golden.py) and, when oracle/_ref/ exists,
against the reference builds themselves."""
import math
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN, load_golden, oracle, random_state, run_bits, stream_of, xg_to_bits

DENSE = ["dense_n1.npz", "dense_n2.npz", "dense_n5.npz", "dense_n8.npz", "dense_n10.npz", "dense_n12.npz"]
CIRCS = sorted(f for f in os.listdir(GOLDEN) if f.startswith("circ_") and f.endswith(".npz"))


@pytest.mark.parametrize("name", DENSE + ["acceleration_1.npz", "acceleration_2.npz"] + CIRCS)
def test_oracle_matches_golden(name):
  z = load_golden(name)
  n = int(z["nbits"])
  stream = stream_of(z)
  refs = [z[k] for k in ("final_spec", "final_xgates") if k in z.files]
  assert refs
  # numpy restatement
  a = oracle.run(z["psi0"].astype(np.complex128).copy(), n, stream)
  # C restatement
  b = oracle.c_run(z["psi0"].astype(np.complex128).copy(), n, stream)
  # index-bit restatement used by the plan interpreter (covers xg_to_bits' negative controls)
  c = run_bits(z["psi0"].astype(np.complex128).copy(), n, xg_to_bits(n, stream))
  for ref in refs:
    assert np.abs(a - ref).max() <= 1e-13
    assert np.abs(b - ref).max() <= 1e-13
    assert np.abs(c - ref).max() <= 1e-13


@pytest.mark.parametrize("name", DENSE)
def test_oracle_complex64_matches_xgates_float(name):
  z = load_golden(name)
  n = int(z["nbits"])
  b = oracle.c_run(z["psi0"].astype(np.complex64), n, stream_of(z))
  # the reference is built with -ffast-math; float results agree to rounding only
  assert np.abs(b - z["final_xgates_f"]).max() <= 2e-5


def test_spec_equals_xgates_in_golden():
  """circuit_test.py:69-107's claim, re-checked on our recorded runs."""
  for name in ("dense_n5.npz", "dense_n8.npz", "acceleration_2.npz"):
    z = load_golden(name)
    assert np.abs(z["final_spec"] - z["final_xgates"]).max() <= 1e-13


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (make -C oracle ref)")
def test_oracle_matches_reference_xgates_live():
  xg = oracle.RefXgates()
  rng = np.random.default_rng(3)
  n = 11
  names = list(oracle.GATES)
  stream = []
  for _ in range(150):
    m = oracle.GATES[names[rng.integers(len(names))]]
    t = int(rng.integers(n))
    if rng.random() < 0.5:
      stream.append((1, 0, t, m))
    else:
      c = int(rng.integers(-n, n))
      if c != t:
        stream.append((2, c, t, m))
  psi0 = random_state(n, 5)
  want = xg.run(psi0.copy(), n, stream)
  assert np.abs(oracle.c_run(psi0.copy(), n, stream) - want).max() <= 1e-13
  assert np.abs(oracle.run(psi0.copy(), n, stream) - want).max() <= 1e-13


def _libq_ops(z):
  ops = []
  for name, a, g in zip(z["names"], z["args"], z["gamma"]):
    name = str(name)
    if name in ("u1",):
      ops.append((name, int(a[0]), float(g)))
    elif name == "cu1":
      ops.append((name, int(a[0]), int(a[1]), float(g)))
    elif name == "ccx":
      ops.append((name, int(a[0]), int(a[1]), int(a[2])))
    elif name in ("cx", "cz"):
      ops.append((name, int(a[0]), int(a[1])))
    else:
      ops.append((name, int(a[0])))
  return ops


@pytest.mark.parametrize("name", ["libq_w4.npz", "libq_w8.npz", "libq_w12.npz"])
def test_libq_dense_model_matches_reference_libq(name):
  z = load_golden(name)
  w = int(z["width"])
  got = oracle.libq_dense(w, int(z["init"]), _libq_ops(z))
  # all-double libq build: only its pruning (apply.cc:107,150-171) separates it from dense math
  assert np.abs(got - z["final_double"]).max() <= 1e-9
  # stock float libq
  assert np.abs(got - z["final_float"]).max() <= 2e-5


def parse_print_qureg(text):
  """qureg.cc:64-78 lines -> {label: complex}."""
  out = {}
  for m in re.finditer(r"^\s*(-?\d+\.\d+) ([+-]\d+\.\d+)i\|(\d+)>", text, re.M):
    out[int(m.group(3))] = complex(float(m.group(1)), float(m.group(2)))
  return out


def test_libq_test_main_golden():
  """libq_test.cc:6-17: h(0) cx(0,1) u1(1, pi/8) on |00>."""
  text = open(os.path.join(GOLDEN, "libq_test.out")).read()
  got = parse_print_qureg(text)
  want = oracle.libq_dense(2, 0, [("h", 0), ("cx", 0, 1), ("u1", 1, math.pi / 8)])
  assert set(got) == {0, 3}
  for k, v in got.items():
    assert abs(want[k] - v) < 2e-6
  assert "# States: 2" in text


def test_qft6_golden_bit_reversal():
  """configs[0]: the transpiled libq program's labels are the bit reversal of the dense
  python index (SURVEY.md trap 2)."""
  z = load_golden("circ_qft6.npz")
  n = 6
  perm = oracle.bitrev_perm(n)
  # replay the generated C++ text's gate list through the dense libq model
  src = open(os.path.join(GOLDEN, "qft6_libq.cc")).read()
  ops = []
  for m in re.finditer(r"libq::(\w+)\(([^;]*), q\);", src):
    name, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
    if name in ("x", "h"):
      ops.append((name, int(args[0])))
    elif name == "cu1":
      ops.append((name, int(args[0]), int(args[1]), eval(args[2].replace("M_PI", "math.pi"))))  # pylint: disable=eval-used
  assert len(ops) == 4 + 21
  got = oracle.libq_dense(n, 0, ops)
  assert np.abs(got[perm] - z["final_xgates"]).max() <= 1e-12
