"""Gate matrices of the reference's operator library (src/lib/ops.py:110-207), as plain
complex128 numpy 2x2 arrays with a `.name`.  Only definitions the gate-application path
needs; the reference's full-matrix operator algebra (O(4^n)) is out of scope."""
from __future__ import annotations

import cmath
import math

import numpy as np


class Gate(np.ndarray):
  """2x2 complex128 matrix with a name (mirrors ops.Operator's use in circuit.py)."""

  def __new__(cls, rows, name=None):
    obj = np.asarray(rows, dtype=np.complex128).reshape(2, 2).view(cls)
    obj.name = name
    return obj

  def __array_finalize__(self, obj):
    self.name = getattr(obj, "name", None)

  def adjoint(self) -> "Gate":                       # ops.py:27
    return Gate(np.conj(np.asarray(self)).T, self.name)


def Identity():                                        # ops.py:110
  return Gate([[1.0, 0.0], [0.0, 1.0]], "Id")


def PauliX():                                          # ops.py:114
  return Gate([[0.0, 1.0], [1.0, 0.0]], "X")


def PauliY():                                          # ops.py:118
  return Gate([[0.0, -1.0j], [1.0j, 0.0]], "Y")


def PauliZ():                                          # ops.py:122
  return Gate([[1.0, 0.0], [0.0, -1.0]], "Z")


def Hadamard():                                        # ops.py:130
  return Gate(1 / np.sqrt(2) * np.array([[1.0, 1.0], [1.0, -1.0]]), "H")


def Phase():                                           # ops.py:136
  return Gate([[1.0, 0.0], [0.0, 1.0j]], "S")


def Sgate():                                           # ops.py:141
  return Phase()


def Tgate():                                           # ops.py:146
  return Gate([[1.0, 0.0], [0.0, cmath.exp(cmath.pi * 1j / 4)]], "T")


def Vgate():                                           # ops.py:152
  return Gate(0.5 * np.array([(1 + 1j, 1 - 1j), (1 - 1j, 1 + 1j)]), "V")


def Yroot():                                           # ops.py:158
  return Gate(0.5 * np.array([(1 + 1j, -1 - 1j), (1 + 1j, 1 + 1j)]), "YRoot")


def U1(lam: float):                                    # ops.py:165
  return Gate([(1.0, 0.0), (0.0, cmath.exp(1j * lam))], "U1")


def U3(theta: float, phi: float, lam: float):          # ops.py:170
  return Gate([(np.cos(theta / 2), -cmath.exp(1j * lam) * np.sin(theta / 2)),
               (cmath.exp(1j * phi) * np.sin(theta / 2),
                cmath.exp(1j * (phi + lam)) * np.cos(theta / 2))], "U3")


def Rk(k: int):                                        # ops.py:178
  return U1(2 * math.pi / (2 ** k))


def Rotation(vparm, theta: float, name: str):          # ops.py:187-195
  v = np.asarray(vparm)
  if v.shape != (3,) or not math.isclose(v @ v, 1) or not np.all(np.isreal(v)):
    raise ValueError("Rotation vector v must be a 3D real unit vector.")
  m = np.cos(theta / 2) * np.asarray(Identity()) - 1j * np.sin(theta / 2) * (
      v[0] * np.asarray(PauliX()) + v[1] * np.asarray(PauliY()) + v[2] * np.asarray(PauliZ()))
  return Gate(m, name + f"({theta:.3f})")


def RotationX(theta: float):                           # ops.py:198
  return Rotation([1.0, 0.0, 0.0], theta, "Rx")


def RotationY(theta: float):                           # ops.py:202
  return Rotation([0.0, 1.0, 0.0], theta, "Ry")


def RotationZ(theta: float):                           # ops.py:206
  return Rotation([0.0, 0.0, 1.0], theta, "Rz")
