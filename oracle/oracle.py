"""oracle.py -- Python side of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker.  The
product package ``qcc_b200`` never imports anything from ``oracle/``.

Contents (citations relative to /root/reference):
  * ``apply1`` / ``applyc``: numpy-vectorised restatement of the dense butterfly
    src/lib/state.py:80-125 == src/lib/xgates.cc:23-67 (MSB-first qubit numbers,
    negative-control quirk included).
  * ``c_apply1`` / ``c_applyc`` / ``c_run``: the same through oracle/qcc_oracle.c
    (liboracle.so) -- scalar loops in xgates' arithmetic order.
  * ``GATES``: the gate matrices of src/lib/ops.py:110-207 restated.
  * ``RefXgates`` / ``RefLibq``: loaders for the REFERENCE builds under
    oracle/_ref/ (present when oracle/Makefile's ``ref`` target has been run).
  * ``libq_dense``: dense complex128 model of the libq face (libq.h:44-64) on
    the oracle-safe gate subset, LSB-first labels (libq.h:35-40).

Parity status: pinned.  tests/test_oracle.py checks every function here against
the golden vectors under tests/golden/ (generated from the reference's Python
spec and its xgates build by tests/golden/make_golden.py) and, when
oracle/_ref/ is present, against the reference binaries directly.
"""
from __future__ import annotations

import cmath
import ctypes
import importlib.machinery
import importlib.util
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


# --------------------------------------------------------------------------
# Gate matrices (src/lib/ops.py:110-207).
# --------------------------------------------------------------------------
def _m(rows):
  return np.array(rows, dtype=np.complex128)


def u1(lam):  # ops.py:165-166
  return _m([[1.0, 0.0], [0.0, cmath.exp(1j * lam)]])


def rotation(v, theta):  # ops.py:187-195
  x = _m([[0, 1], [1, 0]])
  y = _m([[0, -1j], [1j, 0]])
  z = _m([[1, 0], [0, -1]])
  return (np.cos(theta / 2) * np.eye(2) - 1j * np.sin(theta / 2) *
          (v[0] * x + v[1] * y + v[2] * z)).astype(np.complex128)


GATES = {
    "id": _m([[1, 0], [0, 1]]),                                   # ops.py:110
    "x": _m([[0, 1], [1, 0]]),                                    # ops.py:114
    "y": _m([[0, -1j], [1j, 0]]),                                 # ops.py:118
    "z": _m([[1, 0], [0, -1]]),                                   # ops.py:122
    "h": (1 / np.sqrt(2) * np.array([[1.0, 1.0], [1.0, -1.0]])).astype(np.complex128),  # :130
    "s": _m([[1, 0], [0, 1j]]),                                   # ops.py:136
    "t": _m([[1, 0], [0, cmath.exp(cmath.pi * 1j / 4)]]),         # ops.py:146
    "v": 0.5 * _m([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]),          # ops.py:152
    "yroot": 0.5 * _m([[1 + 1j, -1 - 1j], [1 + 1j, 1 + 1j]]),     # ops.py:158
}


# --------------------------------------------------------------------------
# numpy-vectorised dense butterfly (state.py:80-125).
# --------------------------------------------------------------------------
def apply1(psi: np.ndarray, gate: np.ndarray, nbits: int, tgt: int) -> None:
  """In place.  tgt is MSB-first (state.py:85)."""
  t = nbits - tgt - 1
  if t < 0 or t >= nbits:
    raise ValueError("qubit index out of range")
  g = np.asarray(gate).reshape(4).astype(psi.dtype)
  v = psi.reshape(-1, 2, 1 << t)
  a = v[:, 0, :].copy()
  b = v[:, 1, :].copy()
  v[:, 0, :] = g[0] * a + g[1] * b
  v[:, 1, :] = g[2] * a + g[3] * b


def applyc(psi: np.ndarray, gate: np.ndarray, nbits: int, ctl: int, tgt: int) -> None:
  """In place.  ctl/tgt MSB-first; negative ctl follows state.py:118-120."""
  t = nbits - tgt - 1
  c = nbits - ctl - 1
  if t < 0 or t >= nbits:
    raise ValueError("qubit index out of range")
  if c < 0:
    raise ValueError("negative shift count")
  g = np.asarray(gate).reshape(4).astype(psi.dtype)
  n = 1 << nbits
  i = np.arange(n, dtype=np.uint64)
  lo = i[(i >> np.uint64(t)) & np.uint64(1) == 0]          # the `i` of state.py:119
  if c < nbits:
    sel = (lo >> np.uint64(c)) & np.uint64(1) == 1
  else:
    # bit c of g * 2^nbits + i  ==  bit (c - nbits) of the group base g
    gbase = lo & ~np.uint64((1 << (t + 1)) - 1)
    cc = c - nbits
    sel = ((gbase >> np.uint64(cc)) & np.uint64(1) == 1) if cc < 64 else np.zeros(lo.shape, bool)
  i0 = lo[sel].astype(np.int64)
  i1 = i0 + (1 << t)
  a = psi[i0].copy()
  b = psi[i1].copy()
  psi[i0] = g[0] * a + g[1] * b
  psi[i1] = g[2] * a + g[3] * b


def run(psi: np.ndarray, nbits: int, gates) -> np.ndarray:
  """gates: iterable of (kind, ctl, tgt, 2x2) with kind in {1, 2} (circuit.py:180-215)."""
  for kind, ctl, tgt, m in gates:
    if kind == 1:
      apply1(psi, m, nbits, tgt)
    else:
      applyc(psi, m, nbits, ctl, tgt)
  return psi


# --------------------------------------------------------------------------
# C restatement (oracle/qcc_oracle.c).
# --------------------------------------------------------------------------
class _OrcGate(ctypes.Structure):
  _fields_ = [("kind", ctypes.c_int32), ("ctl", ctypes.c_int32), ("tgt", ctypes.c_int32),
              ("pad", ctypes.c_int32), ("m", ctypes.c_double * 8)]


_lib = None


def c_lib():
  global _lib
  if _lib is None:
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
      raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
    _lib = ctypes.CDLL(path)
    for suf in ("d", "f"):
      getattr(_lib, f"orc_apply1_{suf}").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
      getattr(_lib, f"orc_applyc_{suf}").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
      getattr(_lib, f"orc_run_{suf}").argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
  return _lib


def _suf(psi):
  if psi.dtype == np.complex128:
    return "d"
  if psi.dtype == np.complex64:
    return "f"
  raise TypeError(psi.dtype)


def c_apply1(psi, gate, nbits, tgt):
  assert psi.flags.c_contiguous
  g = np.ascontiguousarray(np.asarray(gate).reshape(4), dtype=psi.dtype)
  rc = getattr(c_lib(), f"orc_apply1_{_suf(psi)}")(psi.ctypes.data, g.ctypes.data, nbits, tgt)
  if rc:
    raise ValueError(f"orc_apply1 rc={rc}")


def c_applyc(psi, gate, nbits, ctl, tgt):
  assert psi.flags.c_contiguous
  g = np.ascontiguousarray(np.asarray(gate).reshape(4), dtype=psi.dtype)
  rc = getattr(c_lib(), f"orc_applyc_{_suf(psi)}")(psi.ctypes.data, g.ctypes.data, nbits, ctl, tgt)
  if rc:
    raise ValueError(f"orc_applyc rc={rc}")


def pack_gates(gates):
  arr = (_OrcGate * len(gates))()
  for k, (kind, ctl, tgt, m) in enumerate(gates):
    arr[k].kind = int(kind)
    arr[k].ctl = int(ctl if ctl is not None else 0)
    arr[k].tgt = int(tgt)
    flat = np.asarray(m, dtype=np.complex128).reshape(4)
    for j in range(4):
      arr[k].m[2 * j] = flat[j].real
      arr[k].m[2 * j + 1] = flat[j].imag
  return arr


def c_run(psi, nbits, gates):
  assert psi.flags.c_contiguous
  arr = gates if isinstance(gates, ctypes.Array) else pack_gates(list(gates))
  rc = getattr(c_lib(), f"orc_run_{_suf(psi)}")(psi.ctypes.data, nbits, arr, len(arr))
  if rc:
    raise ValueError(f"orc_run rc={rc}")
  return psi


# --------------------------------------------------------------------------
# Reference builds (oracle/_ref/, made by `make -C oracle ref`).
# --------------------------------------------------------------------------
def have_ref(name="libxgates.so"):
  return os.path.exists(os.path.join(REF_DIR, name))


class RefXgates:
  """The reference's CPython extension, loaded by path so that no `libxgates`
  shim on sys.path can shadow it."""

  def __init__(self):
    path = os.path.join(REF_DIR, "libxgates.so")
    loader = importlib.machinery.ExtensionFileLoader("libxgates", path)
    spec = importlib.util.spec_from_loader("libxgates", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    self.mod = mod

  def apply1(self, psi, gate, nbits, tgt):
    bw = 128 if psi.dtype == np.complex128 else 64
    g = np.ascontiguousarray(np.asarray(gate).reshape(4), dtype=psi.dtype)
    self.mod.apply1(psi, g, nbits, tgt, bw)

  def applyc(self, psi, gate, nbits, ctl, tgt):
    bw = 128 if psi.dtype == np.complex128 else 64
    g = np.ascontiguousarray(np.asarray(gate).reshape(4), dtype=psi.dtype)
    self.mod.applyc(psi, g, nbits, ctl, tgt, bw)

  def run(self, psi, nbits, gates):
    for kind, ctl, tgt, m in gates:
      if kind == 1:
        self.apply1(psi, m, nbits, tgt)
      else:
        self.applyc(psi, m, nbits, ctl, tgt)
    return psi


LIBQ_OPS = {name: i for i, name in enumerate(
    ["x", "y", "z", "h", "t", "v", "yroot", "walsh", "cx", "cz", "ccx", "u1", "cu1",
     "cv", "cv_adj", "gate1"])}


class _RefqOp(ctypes.Structure):
  _fields_ = [("code", ctypes.c_int32), ("a", ctypes.c_int32), ("b", ctypes.c_int32),
              ("c", ctypes.c_int32), ("gamma", ctypes.c_double), ("m", ctypes.c_double * 8)]


def pack_libq_ops(ops):
  """ops: list of (name, args...) e.g. ('h', 3), ('cu1', 0, 2, 0.3), ('ccx', 0, 1, 2),
  ('gate1', tgt, 2x2)."""
  arr = (_RefqOp * len(ops))()
  for k, op in enumerate(ops):
    name = op[0]
    arr[k].code = LIBQ_OPS[name]
    if name in ("u1",):
      arr[k].a, arr[k].gamma = op[1], op[2]
    elif name in ("cu1",):
      arr[k].a, arr[k].b, arr[k].gamma = op[1], op[2], op[3]
    elif name == "ccx":
      arr[k].a, arr[k].b, arr[k].c = op[1], op[2], op[3]
    elif name in ("cx", "cz", "cv", "cv_adj"):
      arr[k].a, arr[k].b = op[1], op[2]
    elif name == "gate1":
      arr[k].a = op[1]
      flat = np.asarray(op[2], dtype=np.complex128).reshape(4)
      for j in range(4):
        arr[k].m[2 * j] = flat[j].real
        arr[k].m[2 * j + 1] = flat[j].imag
    else:
      arr[k].a = op[1]
  return arr


class RefLibq:
  """Stock libq (double=False) or the all-double scratch build (double=True)."""

  def __init__(self, double=False):
    path = os.path.join(REF_DIR, "libq_ref_d.so" if double else "libq_ref.so")
    self.lib = ctypes.CDLL(path)
    self.lib.refq_new.restype = ctypes.c_void_p
    self.lib.refq_new.argtypes = [ctypes.c_ulonglong, ctypes.c_int]
    self.lib.refq_delete.argtypes = [ctypes.c_void_p]
    self.lib.refq_apply.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    self.lib.refq_size.argtypes = [ctypes.c_void_p]
    self.lib.refq_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]

  def run(self, width, initval, ops):
    """Returns (labels uint64[size], amps complex128[size])."""
    q = self.lib.refq_new(initval, width)
    try:
      arr = pack_libq_ops(ops)
      rc = self.lib.refq_apply(q, arr, len(arr))
      if rc:
        raise ValueError("refq_apply failed")
      size = self.lib.refq_size(q)
      labels = np.zeros(size, dtype=np.uint64)
      amps = np.zeros(size, dtype=np.complex128)
      self.lib.refq_read(q, labels.ctypes.data, amps.ctypes.data, size)
      return labels, amps
    finally:
      self.lib.refq_delete(q)

  def run_dense(self, width, initval, ops):
    labels, amps = self.run(width, initval, ops)
    out = np.zeros(1 << width, dtype=np.complex128)
    out[labels.astype(np.int64)] = amps
    return out


# --------------------------------------------------------------------------
# Dense model of the libq face (labels LSB-first: qubit k = bit k, libq.h:35-40).
# Gate definitions: gates.cc:17-146 for the oracle-safe subset; v / yroot / cv /
# cv_adj follow the intended matrices of ops.py:152-162 (stock libq's loops are
# wrong for these -- SURVEY.md trap 4 -- so they are never compared to _ref).
# --------------------------------------------------------------------------
def libq_dense(width, initval, ops, dtype=np.complex128):
  psi = np.zeros(1 << width, dtype=dtype)
  psi[initval] = 1.0
  n = width

  def one(m, tq):       # libq qubit tq == dense bit tq == MSB-first index n-1-tq
    apply1(psi, m, n, n - 1 - tq)

  def ctl(m, cq, tq):
    applyc(psi, m, n, n - 1 - cq, n - 1 - tq)

  for op in ops:
    name = op[0]
    if name in ("x", "y", "z", "h", "t", "v", "yroot"):
      one(GATES[name], op[1])
    elif name == "walsh":
      for i in range(op[1]):
        one(GATES["h"], i)
    elif name == "u1":
      one(u1(op[2]), op[1])
    elif name == "cx":
      ctl(GATES["x"], op[1], op[2])
    elif name == "cz":
      ctl(GATES["z"], op[1], op[2])
    elif name == "cu1":
      ctl(u1(op[3]), op[1], op[2])
    elif name == "cv":
      ctl(GATES["v"], op[1], op[2])
    elif name == "cv_adj":
      ctl(GATES["v"].conj().T, op[1], op[2])
    elif name in ("s", "sdag", "tdag", "vdag", "yrootdag", "hdag", "xdag", "ydag", "zdag"):
      base = GATES[name[:-3]] if name.endswith("dag") else GATES[name]
      one(base.conj().T if name.endswith("dag") else base, op[1])
    elif name in ("rx", "ry", "rz"):
      axis = {"rx": [1.0, 0, 0], "ry": [0, 1.0, 0], "rz": [0, 0, 1.0]}[name]
      one(rotation(axis, op[2]), op[1])
    elif name in ("crx", "cry", "crz"):
      axis = {"crx": [1.0, 0, 0], "cry": [0, 1.0, 0], "crz": [0, 0, 1.0]}[name]
      ctl(rotation(axis, op[3]), op[1], op[2])
    elif name in ("ch", "cs", "ct", "cy", "cyroot"):
      ctl(GATES[name[1:]], op[1], op[2])
    elif name == "ccx":
      c0, c1, tq = op[1], op[2], op[3]
      idx = np.arange(1 << n)
      sel = ((idx >> c0) & 1 == 1) & ((idx >> c1) & 1 == 1) & ((idx >> tq) & 1 == 0)
      i0 = idx[sel]
      i1 = i0 | (1 << tq)
      a = psi[i0].copy()
      psi[i0] = psi[i1]
      psi[i1] = a
    elif name == "gate1":
      one(np.asarray(op[2]).reshape(2, 2), op[1])
    else:
      raise ValueError(name)
  return psi


def bitrev_perm(n):
  """perm[i] = bit-reversal of i over n bits: libq label <-> Python dense index
  (SURVEY.md appendix C)."""
  idx = np.arange(1 << n, dtype=np.int64)
  out = np.zeros_like(idx)
  for b in range(n):
    out |= ((idx >> b) & 1) << (n - 1 - b)
  return out
