"""GPU: BASELINE.json's full sizes (28-30 qubits), checked through size-independent
properties because no oracle finishes at these sizes in seconds (and the reference itself
cannot run above 30 qubits, SURVEY.md trap 3):
  * unitarity: the norm stays 1;
  * QFT then its inverse circuit (circuit.py:423-456: reversed order, adjoint gates) returns
    the initial basis state;
  * QFT of a basis state has a flat spectrum |amp|^2 = 2^-n, and its amplitudes follow the
    closed form, spot-checked;
  * the fused path and the gate-by-gate path agree on sampled slices of a random state.
"""
import math

import numpy as np
import pytest

from helpers import oracle
from qcc_b200 import _cabi

pytestmark = pytest.mark.gpu


def qft_stream(n, inverse=False):
  s = []
  for i in reversed(range(n)):
    s.append((1, 0, i, oracle.GATES["h"]))
    for j in reversed(range(i)):
      s.append((2, i, j, oracle.u1(math.pi / 2 ** (i - j))))
  if inverse:
    s = [(k, c, t, np.asarray(m).conj().T) for k, c, t, m in reversed(s)]
  return s


def qft_amplitude(n, x, k):
  """Amplitude <k| QFT_noswap |x> for circuit.py:320-326 (validated against the oracle at
  n = 10 in test_qft_closed_form_small)."""
  xrev = int(format(x, f"0{n}b")[::-1], 2)
  return np.exp(2j * np.pi * ((xrev * k) % (1 << n)) / (1 << n)) / math.sqrt(1 << n)


def test_qft_closed_form_small():
  n, x = 10, 0b1011001110
  psi = np.zeros(1 << n, dtype=np.complex128)
  psi[x] = 1
  oracle.c_run(psi, n, qft_stream(n))
  want = np.array([qft_amplitude(n, x, k) for k in range(1 << n)])
  assert np.abs(psi - want).max() < 1e-12


@pytest.mark.parametrize("n", [28, 30])
def test_qft_roundtrip_and_flat_spectrum(n):
  x = 0x2F0F3A5 & ((1 << n) - 1)
  fwd = _cabi.pack_xg_gates(qft_stream(n))
  inv = _cabi.pack_xg_gates(qft_stream(n, inverse=True))
  with _cabi.DeviceState(n, x) as s:
    s.xg_apply_gates(fwd)
    assert abs(s.norm2() - 1.0) < 1e-9
    idx, p = s.argmax()
    assert abs(p * (1 << n) - 1.0) < 1e-9
    for k in (0, 1, 12345, (1 << n) - 1, 0x1234567 & ((1 << n) - 1)):
      assert abs(s.amplitude(k) - qft_amplitude(n, x, k)) < 1e-12
    cnt = s.counters()
    assert cnt["passes"] <= 4, cnt
    s.xg_apply_gates(inv)
    a = s.amplitude(x)
    assert abs(a - 1.0) < 1e-9
    assert abs(s.norm2() - 1.0) < 1e-9
    labels, amps, count = s.list_above(1e-12)
    assert count == 1 and labels[0] == x


def test_walsh_30():
  n = 30
  with _cabi.DeviceState(n, 0) as s:
    for t in range(n):
      s.xg_apply1(t, oracle.GATES["h"])
    assert abs(s.norm2() - 1.0) < 1e-9
    _, p = s.argmax()
    assert abs(p * (1 << n) - 1.0) < 1e-9
    for bit in (0, 13, 29):
      assert abs(s.prob_bit(bit) - 0.5) < 1e-9


def test_fused_equals_single_gate_path_at_28():
  n = 28
  rng = np.random.default_rng(28)
  names = ["h", "v", "yroot", "t", "x", "y", "z", "s"]
  stream = []
  for _ in range(60):
    m = oracle.GATES[names[rng.integers(len(names))]] if rng.random() < 0.8 else oracle.u1(float(rng.uniform(-3, 3)))
    t = int(rng.integers(n))
    if rng.random() < 0.5:
      stream.append((1, 0, t, m))
    else:
      c = int(rng.integers(n))
      if c != t:
        stream.append((2, c, t, m))
  packed = _cabi.pack_xg_gates(stream)
  with _cabi.DeviceState(n) as a, _cabi.DeviceState(n) as b:
    a.fill_random(5)
    b.fill_random(5)
    b.set_fusion(False)
    a.xg_apply_gates(packed)
    b.xg_apply_gates(packed)
    assert abs(a.norm2() - 1.0) < 1e-9 and abs(b.norm2() - 1.0) < 1e-9
    for first in (0, 1 << 20, (1 << 27) + 4093, (1 << 28) - (1 << 16)):
      x = a.copy_out(first, 1 << 16)
      y = b.copy_out(first, 1 << 16)
      assert np.abs(x - y).max() < 1e-15 * 1e3  # amplitudes are ~6e-5; agreement to ~1e-16 relative
    assert a.counters()["passes"] * 3 < b.counters()["passes"]


def test_larose_28_fused_equals_single_gate_path():
  """configs[1] at full size: two depths of larose_benchmark.py:47-54 (h, v, cx(bit, 0) per qubit) --
  scheduled UX rounds, the cx fan-in as parity swaps, direct-store last rounds -- against the
  gate-by-gate sweeps on sampled slices of a random state."""
  n = 28
  stream = []
  for _ in range(2):
    for bit in range(n):
      stream.append((1, 0, bit, oracle.GATES["h"]))
      stream.append((1, 0, bit, oracle.GATES["v"]))
      if bit > 0:
        stream.append((2, bit, 0, oracle.GATES["x"]))
  packed = _cabi.pack_xg_gates(stream)
  with _cabi.DeviceState(n) as a, _cabi.DeviceState(n) as b:
    a.fill_random(7)
    b.fill_random(7)
    b.set_fusion(False)
    a.xg_apply_gates(packed)
    b.xg_apply_gates(packed)
    assert abs(a.norm2() - 1.0) < 1e-9 and abs(b.norm2() - 1.0) < 1e-9
    for first in (0, (1 << 19) + 8, (1 << 27) - (1 << 15), (1 << 27) + 4093 * 16, (1 << 28) - (1 << 16)):
      x = a.copy_out(first, 1 << 16)
      y = b.copy_out(first, 1 << 16)
      assert np.abs(x - y).max() < 1e-12
    assert a.counters()["passes"] == 6 and b.counters()["passes"] == len(stream)


# ---------------------------------------------------------------------------------------------------
# Whole-vector parity against the REFERENCE at 26 and 28 qubits: the unmodified src/lib/xgates.cc build
# (oracle/_ref/libxgates.so, shipped to the GPU box) handles these sizes in a fraction of a second per gate;
# the C restatement of the same loops (oracle.c_run) stands in where that build is absent.  At >= 2^14
# tiles these are the sizes where 64-bit index arithmetic, the tile-number -> base scatter over several
# runs of non-tile bits and the per-tile constants of ladders with many partners outside the tile can go
# wrong without any small test noticing.
# ---------------------------------------------------------------------------------------------------
def _reference_run(psi, n, stream):
  if oracle.have_ref("libxgates.so"):
    return oracle.RefXgates().run(psi, n, stream), "reference xgates"
  return oracle.c_run(psi, n, stream), "oracle restatement"


def _interesting_qubits(n):
  """python qubits whose index bits are 0-2 (inside every 128-byte run), 12-14 (around the tile edge at
  K = 12), the middle and n-1 (the longest stride)."""
  bits = [0, 1, 2, 12, 13, 14, n // 2 + 3, n - 2, n - 1]
  return [n - 1 - b for b in bits]


def _big_streams(n):
  H, V, X = oracle.GATES["h"], oracle.GATES["v"], oracle.GATES["x"]
  out = {}
  # QFT blocks (circuit.py:320-326) on chosen pivots: h + the cu1 ladder to EVERY lower qubit, so the pivot on
  # index bit 0 drags a ladder with n - 1 partners, most of them outside any tile
  qft = []
  for k, p in enumerate(sorted({n - 1, n - 3, n - 14, n // 2, 1}, reverse=True)):
    qft.append((1, 0, p, H))
    lower = list(reversed(range(p)))
    if k:                                  # later pivots: nearest two partners, every third one, and the far end
      lower = [j for i, j in enumerate(lower) if i < 2 or i % 3 == 0 or j < 2][:12]
    for j in lower:
      qft.append((2, p, j, oracle.u1(math.pi / 2 ** (p - j))))
  out["qft_blocks"] = qft
  # larose_benchmark.py:47-54 on a subset of the qubits (cx onto python qubit 0 = index bit n-1, controls anywhere)
  lar = []
  for bit in _interesting_qubits(n) + [5, 9]:
    lar += [(1, 0, bit, H), (1, 0, bit, V)]
    if bit:
      lar.append((2, bit, 0, X))
  out["larose_subset"] = lar
  # random gates: targets from the interesting set, controls anywhere (mostly outside the tile)
  rng = np.random.default_rng(n)
  names = ["h", "v", "yroot", "t", "x", "y", "z", "s"]
  rnd = []
  tq = _interesting_qubits(n)
  while len(rnd) < 36:
    r = rng.random()
    m = oracle.u1(float(rng.uniform(-3, 3))) if r < 0.25 else (
        oracle.rotation([1.0, 0, 0], float(rng.uniform(-3, 3))) if r < 0.35 else oracle.GATES[names[rng.integers(len(names))]])
    t = tq[int(rng.integers(len(tq)))]
    if rng.random() < 0.4:
      rnd.append((1, 0, t, m))
    else:
      c = int(rng.integers(n))
      if c != t:
        rnd.append((2, c, t, m))
  out["random"] = rnd
  return out


@pytest.mark.parametrize("n,name,tile_bits", [(26, "qft_blocks", 12), (26, "qft_blocks", 7), (26, "larose_subset", 12),
                                              (26, "random", 12), (26, "random", 10), (28, "qft_blocks", 12),
                                              (28, "random", 12)])
def test_whole_vector_matches_the_reference_at_26_and_28_qubits(n, name, tile_bits):
  """tile_bits = 7 at 26 qubits gives the ladders 19 partner bits outside the tile (more than the 18 a
  30-qubit QFT pass has at K = 12)."""
  stream = _big_streams(n)[name]
  rng = np.random.default_rng(100 + n)
  psi0 = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
  psi0 /= np.linalg.norm(psi0)
  with _cabi.DeviceState(n) as s:
    s.set_tile_bits(tile_bits)
    s.copy_in(psi0)
    s.xg_apply_gates(_cabi.pack_xg_gates(stream))
    got = s.copy_out()
    cnt = s.counters()
  want, who = _reference_run(psi0, n, stream)      # in place: psi0 is the reference result now
  err = float(np.abs(got - want).max())
  assert err <= 1e-10, (n, name, tile_bits, who, err)
  assert cnt["passes"] < len(stream)               # it did go through fused passes
