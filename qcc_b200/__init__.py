"""qcc_b200 -- B200-native state-vector gate application behind qcc's interfaces.

Only what the hot path needs lives here:
  csrc/      hand-written sm_100a CUDA + the C ABI (include/qcc_b200.h)
  _cabi.py   ctypes binding of that ABI (no torch, no CPU fallback)
  ops.py     the gate matrices of the reference's src/lib/ops.py
  circuit.py device-resident mirror of the reference's circuit.qc operator surface
  libq/      C++ header + library with the reference's libq API over the same ABI
  shim/      `libxgates.py`: drop-in for the reference's CPython extension
"""
__version__ = "0.1.0"
