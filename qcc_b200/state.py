"""Registers and the device-resident state proxy of the python face.

`Reg` mirrors src/lib/state.py:249-267 (a named run of global qubit indices with initial
values).  `DevicePsi` stands where the reference has `qc.psi`, a numpy `State`
(src/lib/state.py:15-154): it offers the same readouts (nbits, ampl, prob, phase, maxprob,
indexing, dump) but every one of them is a device reduction or a small device->host copy
through the C ABI; the 2^n vector itself never leaves HBM unless the caller asks for it
with np.asarray(qc.psi) / qc.psi.numpy()."""
from __future__ import annotations

import cmath
import math
from typing import List, Tuple

import numpy as np

from qcc_b200 import helper


class Reg:
  """A register: `size` consecutive global qubits starting at `global_reg`."""

  def __init__(self, size: int, it=0, global_reg: int = None):
    self.size = size
    self.global_idx = list(range(global_reg, global_reg + size))
    self.val = [0] * size
    self.global_reg = global_reg
    if it:
      if isinstance(it, int):
        it = format(it, f"0{size}b")                      # MSB first (state.py:256-258)
      if isinstance(it, (str, tuple, list)):
        for k, ch in enumerate(it):
          if ch in ("1", 1):
            self.val[k] = 1

  def __getitem__(self, idx):
    return self.global_idx[idx]

  def __setitem__(self, idx: int, val: int) -> None:
    self.val[idx] = val

  def __len__(self) -> int:
    return self.size

  @property
  def nbits(self) -> int:
    return self.size

  def __str__(self) -> str:
    return "|" + "".join(str(v) for v in self.val) + ">"


class DevicePsi:
  """Readout surface over the engine's state (see module docstring)."""

  MATERIALIZE_LIMIT = 28  # np.asarray(psi) beyond 4 GiB must be asked for explicitly

  def __init__(self, dev):
    self._dev = dev

  @property
  def nbits(self) -> int:
    return self._dev.nqubits

  def __len__(self) -> int:
    return 1 << self._dev.nqubits

  @property
  def shape(self):
    return (len(self),)

  # -- single amplitudes -------------------------------------------------------------
  def __getitem__(self, idx):
    if isinstance(idx, slice):
      start, stop, step = idx.indices(len(self))
      if step != 1:
        return self.numpy()[idx]
      self._whole_vector_only("slicing")
      return self._dev.copy_out(start, max(0, stop - start))
    idx = int(idx)
    if idx < 0:
      idx += len(self)
    return self._dev.amplitude(idx)

  def ampl(self, *bits) -> complex:                       # state.py:30-33
    return self._dev.amplitude(helper.bits2val(bits))

  def prob(self, *bits) -> float:                         # state.py:36-40
    a = self.ampl(*bits)
    return (a.conjugate() * a).real

  def phase(self, *bits) -> float:                        # state.py:42-46
    return math.degrees(cmath.phase(self.ampl(*bits)))

  # -- reductions ----------------------------------------------------------------------
  def maxprob(self) -> Tuple[List[int], float]:           # state.py:60-78
    idx, p = self._dev.argmax()
    return helper.val2bits(idx, self.nbits), p

  def norm2(self) -> float:
    return self._dev.norm2()

  def prob_of_qubit(self, qubit: int) -> float:
    """P(python qubit `qubit` == 1)."""
    return self._dev.prob_bit(self.nbits - 1 - qubit)

  def weight_of_qubit(self, qubit: int, value: int) -> float:
    """Sum of |amp|^2 over the basis states whose python qubit `qubit` equals `value` (not normalised)."""
    return self._dev.prob_bit_value(self.nbits - 1 - qubit, value)

  def nonzero(self, threshold: float = 1e-12, cap: int = 1 << 16):
    """(index, amplitude) of every basis state with |amp|^2 >= threshold, ascending index."""
    labels, amps, total = self._dev.list_above(threshold, cap)
    return labels, amps, total

  # -- whole vector ----------------------------------------------------------------------
  def _whole_vector_only(self, what: str) -> None:
    if getattr(self._dev, "nranks", 1) > 1:
      raise NotImplementedError(f"{what} a sharded state: this process holds 1/{self._dev.nranks} of the vector; "
                                "use psi.local_slice() (this rank's slice of the canonical vector) or the "
                                "collective readouts (ampl, prob, maxprob, nonzero, ...)")

  def local_slice(self) -> np.ndarray:
    """This rank's contiguous slice [rank * 2^n / nranks, (rank + 1) * 2^n / nranks) of the canonical vector
    (the whole vector when the state is not sharded).  Collective on a sharded state: the engine first undoes
    whatever bit layout the exchange events left behind."""
    return self._dev.copy_out()

  def numpy(self, force: bool = False) -> np.ndarray:
    self._whole_vector_only("copying out")
    if self.nbits > self.MATERIALIZE_LIMIT and not force:
      raise MemoryError(f"refusing to copy a {self.nbits}-qubit state to the host implicitly; "
                        "call psi.numpy(force=True)")
    return self._dev.copy_out()

  def __array__(self, dtype=None, copy=None):
    a = self.numpy()
    return a.astype(dtype) if dtype is not None else a

  def is_close(self, other, tol: float = 1e-6) -> bool:
    return bool(np.allclose(self.numpy(), np.asarray(other), atol=tol))

  def dump(self, desc: str = None, prob_only: bool = True) -> None:   # state.py:127-154
    if desc:
      print("|", end="")
      for i in range(self.nbits - 1, -1, -1):
        print(i % 10, end="")
      print(f"> '{desc}'")
    labels, amps, _ = self.nonzero(1e-12 if prob_only else -1.0, 1 << 16)
    for lab, a in zip(labels, amps):
      bits = helper.val2bits(int(lab), self.nbits)
      p = (a.conjugate() * a).real
      print("|{}> ({}): ampl: {:+.2f} prob: {:.2f} Phase: {:5.1f}".format(
          "".join(str(b) for b in bits), int(lab), complex(a), p, math.degrees(cmath.phase(a))))
