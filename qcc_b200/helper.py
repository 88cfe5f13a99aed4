"""Index/bit conventions of the reference's python face (src/lib/helper.py:18-31, 84-106):
qubit 0 is the MOST significant bit of the dense state index."""
from __future__ import annotations

import math
from typing import Iterable, List


def bits2val(bits: Iterable[int]) -> int:
  """(1, 0, 1) -> 5: first bit is the most significant (helper.py:18-23)."""
  v = 0
  for b in bits:
    v = (v << 1) | (1 if b else 0)
  return v


def val2bits(val: int, nbits: int) -> List[int]:
  """5, 3 -> [1, 0, 1] (helper.py:26-31)."""
  return [(val >> (nbits - 1 - k)) & 1 for k in range(nbits)]


def bits2frac(bits: Iterable[int]) -> float:
  """(1, 0, 1) -> 0.625: the bits as a binary fraction, first bit = 1/2 (helper.py:34-37)."""
  return sum((1 if b else 0) * 2.0 ** (-(k + 1)) for k, b in enumerate(bits))


def pi_fractions(val, pi: str = "pi") -> str:
  """Pretty-print an angle as a fraction of pi when it is one (helper.py:84-106): the
  transpiler relies on this to emit M_PI/2, -M_PI/4, ... instead of decimals."""
  if val is None:
    return ""
  if val == 0:
    return "0"
  for mult in (1, 2, 3):
    for den in range(-128, 128):
      if den and math.isclose(val, mult * math.pi / den):
        head = "" if mult == 1 else f"{mult}*"
        sign = "-" if den < 0 else ""
        tail = "" if abs(den) == 1 else f"/{abs(den)}"
        return f"{sign}{head}{pi}{tail}"
  return f"{val}"
