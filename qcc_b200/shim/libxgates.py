"""libxgates -- drop-in for the reference's CPython extension of the same name.

The reference's circuit.py does `import libxgates as xgates` and calls
`xgates.apply1(psi, gate, nbits, tgt, bit_width)` / `xgates.applyc(psi, gate, nbits, ctl, tgt,
bit_width)` once per gate on the numpy state it owns (src/lib/circuit.py:36-41, 195-214;
the extension is src/lib/xgates.cc:89-145).  Put this directory on PYTHONPATH *instead of*
the directory holding the reference's libxgates.so and the unchanged reference code runs its
gates on the B200:

    PYTHONPATH=/path/to/qcc:/path/to/repo/qcc_b200/shim python src/supremacy.py

Semantics kept: positional arguments, in-place update of `psi`, complex64 unless
bit_width == 128, python qubit numbering including negative control indices, None return.
Differences: a dtype / contiguity mismatch raises TypeError (the reference silently updates
a temporary copy and leaves `psi` unchanged); an out-of-range qubit
raises ValueError instead of exit(1) (xgates.cc:28-32).

This path is correct but PCIe-bound by construction -- the state crosses the bus twice per
gate because the caller owns it as a host array.  The fast path is qcc_b200.circuit.qc, which
keeps the state in HBM.  There is no CPU fallback: without the CUDA library the import fails.
"""
import os
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(os.path.dirname(_here))
if _root not in sys.path:
  sys.path.insert(0, _root)

from qcc_b200 import _cabi  # noqa: E402

_lib = _cabi.lib()  # raises if the library has not been built
_DEVICE = int(os.environ.get("QCC_B200_DEVICE", "-1"))
_calls = {"apply1": 0, "applyc": 0}
if os.environ.get("QCC_B200_SHIM_LOG"):   # tests: proof that the reference's gates came through here
  import atexit
  import json

  def _write_log():
    with open(os.environ["QCC_B200_SHIM_LOG"], "w") as f:
      json.dump(_calls, f)

  atexit.register(_write_log)


def _check(psi, gate, nbits, bit_width):
  want = np.complex128 if bit_width == 128 else np.complex64
  if not isinstance(psi, np.ndarray) or psi.dtype != want or not psi.flags.c_contiguous:
    raise TypeError(f"psi must be a C-contiguous numpy array of {np.dtype(want)} for bit_width={bit_width}")
  if psi.size != 1 << nbits:
    raise ValueError(f"psi has {psi.size} amplitudes, expected 2^{nbits}")
  g = np.ascontiguousarray(np.asarray(gate).reshape(4), dtype=want)
  return g


def apply1(psi, gate, nbits, tgt, bit_width):
  """Apply a single-qubit gate in place (xgates.cc:89-107)."""
  g = _check(psi, gate, nbits, bit_width)
  _calls["apply1"] += 1
  rc = _lib.qb_host_apply1(psi.ctypes.data, g.ctypes.data, int(nbits), int(tgt), int(bit_width), _DEVICE)
  if rc == -1:
    raise ValueError(_lib.qb_last_error().decode())
  _cabi.check(rc)
  return None


def applyc(psi, gate, nbits, ctl, tgt, bit_width):
  """Apply a controlled single-qubit gate in place (xgates.cc:126-145)."""
  g = _check(psi, gate, nbits, bit_width)
  _calls["applyc"] += 1
  rc = _lib.qb_host_applyc(psi.ctypes.data, g.ctypes.data, int(nbits), int(ctl), int(tgt), int(bit_width),
                           _DEVICE)
  if rc == -1:
    raise ValueError(_lib.qb_last_error().decode())
  _cabi.check(rc)
  return None
