#!/usr/bin/env python
"""A/B timing of two builds of the library on the same box: python scripts/ab_lib.py <lib.so> [workload] [n].
Only ABI-stable entry points are used (create, fill_random, xg_apply_gates, timer), so older builds load."""
import ctypes
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qcc_b200 import _cabi, workloads  # noqa: E402

path = sys.argv[1]
wl = sys.argv[2] if len(sys.argv) > 2 else "qft"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
L = ctypes.CDLL(path)
P = ctypes.c_void_p
L.qb_state_create.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(P)]
L.qb_fill_random.argtypes = [P, ctypes.c_uint64]
L.qb_xg_apply_gates.argtypes = [P, ctypes.POINTER(_cabi.qb_xg_gate), ctypes.c_int64]
L.qb_timer_start.argtypes = [P]
L.qb_timer_stop.argtypes = [P, ctypes.POINTER(ctypes.c_double)]
L.qb_sync.argtypes = [P]
L.qb_state_destroy.argtypes = [P]
L.qb_last_error.restype = ctypes.c_char_p
stream = workloads.qft(n) if wl == "qft" else workloads.larose(n, n)
packed = _cabi.pack_xg_gates(stream)
h = P()
assert L.qb_state_create(n, 0, 0, ctypes.byref(h)) == 0, L.qb_last_error()
L.qb_fill_random(h, 1234)
res = []
for rep in range(3):
  for _ in range(3):
    L.qb_xg_apply_gates(h, packed, len(packed))
  L.qb_sync(h)
  L.qb_timer_start(h)
  K = 10 if wl == "qft" else 3
  for _ in range(K):
    L.qb_xg_apply_gates(h, packed, len(packed))
  ms = ctypes.c_double()
  L.qb_timer_stop(h, ctypes.byref(ms))
  res.append(ms.value / K)
print(os.path.relpath(path, ROOT), wl, n, "ms/step", " ".join(f"{x:.2f}" for x in res))
L.qb_state_destroy(h)
