"""class qc -- the reference's circuit surface (src/lib/circuit.py:68-534) over a state that
lives in B200 HBM.

Same constructor, same state builders, same gate methods with the same argument meaning,
same IR recording / sub-circuit / inverse / control_by machinery, so the reference's
algorithms (`src/*.py`) run on it by swapping the import.  What differs:

  * `qc.psi` is a `state.DevicePsi` proxy, not a numpy array: its readouts (maxprob, prob,
    ampl, indexing, dump) are device reductions / tiny copies; `np.asarray(qc.psi)` copies
    the whole vector out (refused above 28 qubits unless forced).
  * every gate is handed to the engine (C ABI, python qubit numbering: qb_xg_apply1 /
    qb_xg_applyc == xgates.cc:23-67 incl. the negative-control predicate), which queues and
    fuses them; nothing is computed on the CPU and there is no fallback.
  * states are built on the device from basis labels (`reg`, `zeros`, `ones`, `bitstring`,
    `rand_bits`); dense factors (`qubit(alpha, beta)`, `arange`, `random`, `state`) are
    combined on the host while the circuit is still small and uploaded once.
  * transpilation is selected per call (`dump_to_file(libq=...)`, `libq()`), not by absl flags.
"""
from __future__ import annotations

import math
import random as _random
from typing import List, Optional, Tuple

import numpy as np

from qcc_b200 import _cabi, dumpers, ir, ops, state

HOST_COMBINE_LIMIT = 28  # largest state (qubits) we will kron / expand through host memory


_SQRTM_CACHE = {}


def _sqrtm2(m: np.ndarray) -> np.ndarray:
  """Principal square root of a 2x2 (circuit.py:238 uses scipy.linalg.sqrtm).  Memoised on the matrix bytes: a
  Toffoli ladder asks for the root of the same X thousands of times."""
  a = np.ascontiguousarray(np.asarray(m, dtype=np.complex128))
  key = a.tobytes()
  r = _SQRTM_CACHE.get(key)
  if r is None:
    from scipy.linalg import sqrtm
    r = np.asarray(sqrtm(a), dtype=np.complex128)
    r.setflags(write=False)
    if len(_SQRTM_CACHE) < 4096:
      _SQRTM_CACHE[key] = r
  return r


class qc:
  """Quantum circuit: device-resident state + gate surface + IR."""

  def __init__(self, name=None, eager: bool = True, *, device: int = 0, fusion: bool = True,
               tile_bits: int = 12, rank: int = 0, nranks: int = 1, comm_id: bytes = None):
    """rank / nranks / comm_id: this process holds one shard of a state split over nranks GPUs
    (one process per GPU; see _cabi.DeviceState).  Every gate and readout is then collective."""
    self.name = name
    self.ir = ir.Ir()
    self.eager = eager
    self.build_ir = not eager
    self.global_reg = 0
    self.sub_circuits = 0
    self._device = device
    self._fusion = fusion
    self._tile_bits = tile_bits
    self._shard = dict(rank=rank, nranks=nranks, comm_id=comm_id)
    self._dev: Optional[_cabi.DeviceState] = None
    self._gbuf = _cabi.GateBuffer()      # gate records not yet handed to the engine (see _queue)
    self._pending: List[Tuple[str, object, int]] = []   # ('basis', bits, n) | ('dense', vec, n)

    self.simple_gates = [
        ["h", ops.Hadamard()], ["s", ops.Sgate()], ["t", ops.Tgate()], ["v", ops.Vgate()],
        ["x", ops.PauliX()], ["y", ops.PauliY()], ["z", ops.PauliZ()], ["yroot", ops.Yroot()],
    ]
    for gname, gate in self.simple_gates:           # circuit.py:97-101
      self.add_single(gname, gate)
      self.add_single(gname + "dag", gate.adjoint())
      self.add_ctl("c" + gname, gate)
      self.add_ctl("c" + gname + "dag", gate.adjoint())

  # --- state ---------------------------------------------------------------------
  @property
  def nbits(self) -> int:
    n = self._dev.nqubits if self._dev is not None else 0
    return n + sum(f[2] for f in self._pending)

  def _new_device_state(self, n: int, label: int = 0) -> _cabi.DeviceState:
    dev = _cabi.DeviceState(n, label, self._device, **self._shard)
    dev.set_fusion(self._fusion)
    if n - int(math.log2(self._shard["nranks"])) >= 4:
      dev.set_tile_bits(self._tile_bits)
    return dev

  def _materialize(self) -> None:
    """Bring every pending factor onto the device (tensor product order = call order)."""
    if not self._pending:
      if self._dev is None:
        raise AssertionError("circuit has no qubits yet")
      return
    self._flush_gates()      # queued gates act on the state as it is now, before it grows
    total = self.nbits
    if self._shard["nranks"] > 1 and not (self._dev is None and all(f[0] == "basis" for f in self._pending)):
      raise NotImplementedError("sharded circuits must be built from basis registers before the first gate")
    if self._dev is None and all(f[0] == "basis" for f in self._pending):
      label = 0
      for _, bits, _ in self._pending:
        for b in bits:
          label = (label << 1) | int(b)
      self._dev = self._new_device_state(total, label)
      self._pending = []
      return
    if total > HOST_COMBINE_LIMIT:
      raise MemoryError(f"combining dense factors into a {total}-qubit state would go through host "
                        f"memory; build large states from basis registers only")
    vec = self._dev.copy_out() if self._dev is not None else np.ones(1, dtype=np.complex128)
    for kind, payload, n in self._pending:
      if kind == "basis":
        label = 0
        for b in payload:
          label = (label << 1) | int(b)
        out = np.zeros(vec.size << n, dtype=np.complex128)
        out[label::1 << n] = vec
        vec = out
      else:
        vec = np.kron(vec, np.asarray(payload, dtype=np.complex128))
    if self._dev is not None:
      self._dev.close()
    self._dev = self._new_device_state(total, 0)
    self._dev.copy_in(vec)
    self._pending = []

  @property
  def dev(self) -> _cabi.DeviceState:
    """The engine handle, with every gate issued so far handed over (readouts, copies and direct
    engine calls all come through here, so they observe all prior gates -- gates_jit.cc's protocol)."""
    self._materialize()
    self._flush_gates()
    return self._dev

  def _flush_gates(self) -> None:
    if self._gbuf.n and self._dev is not None:
      self._dev.xg_apply_buffer(self._gbuf)

  def _queue(self, kind: int, ctl: int, tgt: int, gate) -> None:
    """One gate record for the engine.  Records are batched on the host and cross the C ABI thousands at
    a time (qb_xg_apply_gates); the engine queues and fuses them as before."""
    if self._pending or self._dev is None:
      self._materialize()
    if self._gbuf.append(kind, ctl, tgt, gate):
      self._flush_gates()

  @property
  def psi(self) -> state.DevicePsi:
    return state.DevicePsi(self.dev)

  @psi.setter
  def psi(self, vec) -> None:
    vec = np.ascontiguousarray(np.asarray(vec), dtype=np.complex128).reshape(-1)
    n = int(round(math.log2(vec.size)))
    assert 1 << n == vec.size, "state length must be a power of two"
    if self._dev is not None:
      self._dev.close()
    self._pending = []
    self._gbuf.n = 0          # gates queued for the state that is being replaced
    self._dev = self._new_device_state(n, 0)
    self._dev.copy_in(vec)

  def _tprod(self, kind: str, payload, nqubits: int) -> None:
    self._pending.append((kind, payload, nqubits))
    self.global_reg += nqubits

  def reg(self, size: int, it=0, *, name: str = None) -> state.Reg:      # circuit.py:125-129
    ret = state.Reg(size, it, self.global_reg)
    self._tprod("basis", list(ret.val), size)
    self.ir.reg(size, name, ret)
    return ret

  def qubit(self, alpha=None, beta=None) -> None:                        # circuit.py:131-133
    if alpha is None and beta is None:
      raise ValueError("Both alpha and beta need to be specified")
    if beta is None:
      beta = math.sqrt(1.0 - np.conj(alpha) * alpha)
    if alpha is None:
      alpha = math.sqrt(1.0 - np.conj(beta) * beta)
    if not math.isclose(np.conj(alpha) * alpha + np.conj(beta) * beta, 1.0):
      raise ValueError("Qubit probabilities do not add to 1.")
    self._tprod("dense", np.array([alpha, beta], dtype=np.complex128), 1)

  def zeros(self, n: int) -> None:
    self._tprod("basis", [0] * n, n)

  def ones(self, n: int) -> None:
    self._tprod("basis", [1] * n, n)

  def bitstring(self, *bits) -> None:
    assert len(bits), "Need to specify at least 1 qubit"
    assert all(b in (0, 1) for b in bits), "Bits must be 0 or 1"
    self._tprod("basis", list(bits), len(bits))

  def rand_bits(self, n: int) -> None:
    self._tprod("basis", [_random.randint(0, 1) for _ in range(n)], n)

  def arange(self, n: int) -> None:                                      # circuit.py:149-151
    self.psi = np.arange(0, 2 ** n, dtype=np.float64)
    self.global_reg += n

  def random(self, n: int = 1) -> None:                                  # circuit.py:153-155
    from scipy.stats import unitary_group
    u = np.asarray(unitary_group.rvs(1 << n), dtype=np.complex128)
    self.psi = u[:, 0]

  def state(self, t) -> state.Reg:                                       # circuit.py:161-166
    vec = np.asarray(t, dtype=np.complex128).reshape(-1)
    n = int(round(math.log2(vec.size)))
    ret = state.Reg(n, 0, self.global_reg)
    self._tprod("dense", vec, n)
    self.ir.reg(n, "state", ret)
    return ret

  # --- gates -----------------------------------------------------------------------
  @staticmethod
  def _ctl_by_0(ctl):
    if isinstance(ctl, (int, np.integer)):
      return int(ctl), False
    return ctl[0], True

  def add_single(self, name: str, gate) -> None:
    setattr(self, name, lambda idx, cond=True: self.apply1(gate, idx, name) if cond else None)

  def add_ctl(self, name: str, gate) -> None:
    setattr(self, name, lambda idx0, idx1, cond=True: self.applyc(gate, idx0, idx1, name) if cond else None)

  def apply1(self, gate, idx_set, name: str = None, *, val: float = None) -> None:   # circuit.py:180-197
    indices = []
    if isinstance(idx_set, (int, np.integer)):
      indices.append(int(idx_set))
    if isinstance(idx_set, (state.Reg, list)):
      indices += idx_set[:]
    for idx in indices:
      if self.build_ir:
        self.ir.single(name, idx, gate, val)
      if self.eager:
        assert idx < self.nbits, "Invalid qubit index"
        self._queue(1, 0, idx, gate)

  def applyc(self, gate, ctl, idx, name: str = None, *, val: float = None) -> None:  # circuit.py:199-215
    if isinstance(idx, state.Reg):
      assert len(idx) == 1, "Controlled n-qbit register not supported"
      idx = idx[0]
    ctl_qubit, by_0 = self._ctl_by_0(ctl)
    self.x(ctl_qubit, by_0)
    if self.build_ir:
      self.ir.controlled(name, ctl_qubit, idx, gate, val)
    if self.eager:
      assert idx < self.nbits, "Invalid qubit index"
      self._queue(2, ctl_qubit, idx, gate)
    self.x(ctl_qubit, by_0)

  def cx0(self, idx0: int, idx1: int) -> None:
    self.apply1(ops.PauliX(), idx0, "x")
    self.applyc(ops.PauliX(), idx0, idx1, "cx")
    self.apply1(ops.PauliX(), idx0, "x")

  def cu(self, idx0: int, idx1: int, op, desc: str = None) -> None:
    assert np.asarray(op).shape[0] == 2, "cu only supports 2x2 operators"
    self.applyc(op, idx0, idx1, desc)

  def ccu(self, idx0, idx1, idx2: int, op, desc: str = "") -> None:     # circuit.py:227-246
    """Sleator-Weinfurter: cu(sqrt U), cx, cu(sqrt U^dagger), cx, cu(sqrt U)."""
    i0, c0_by_0 = self._ctl_by_0(idx0)
    i1, c1_by_0 = self._ctl_by_0(idx1)
    opname = getattr(op, "name", None) or "U"
    with self.scope(self.ir, f"CC{opname}\\{desc}({idx0},{idx1},{idx2})"):
      self.x(i0, c0_by_0)
      self.x(i1, c1_by_0)
      root = ops.Gate(_sqrtm2(op), opname)
      self.cu(i0, idx2, root, opname + "^{1/2}")
      self.cx(i0, i1)
      self.cu(i1, idx2, root.adjoint(), opname + "^t")
      self.cx(i0, i1)
      self.cu(i1, idx2, root, opname + "^{1/2}")
      self.x(i1, c1_by_0)
      self.x(i0, c0_by_0)

  def ccx(self, idx0, idx1, idx2: int) -> None:
    self.ccu(idx0, idx1, idx2, ops.PauliX(), "ccx")

  def toffoli(self, idx0, idx1, idx2: int) -> None:
    self.ccu(idx0, idx1, idx2, ops.PauliX(), "ccx")

  def u1(self, idx: int, val) -> None:
    self.apply1(ops.U1(val), idx, "u1", val=val)

  def cu1(self, idx0: int, idx1: int, value) -> None:
    self.applyc(ops.U1(value), idx0, idx1, "cu1", val=value)

  def ccu1(self, idx0: int, idx1: int, tgt: int, value) -> None:
    self.ccu(idx0, idx1, tgt, ops.U1(value))

  def rx(self, idx: int, theta: float) -> None:
    self.apply1(ops.RotationX(theta), idx, "rx", val=theta)

  def ry(self, idx: int, theta: float) -> None:
    self.apply1(ops.RotationY(theta), idx, "ry", val=theta)

  def rz(self, idx: int, theta: float) -> None:
    self.apply1(ops.RotationZ(theta), idx, "rz", val=theta)

  def crx(self, ctl: int, idx: int, theta: float) -> None:
    self.applyc(ops.RotationX(theta), ctl, idx, "crx", val=theta)

  def cry(self, ctl: int, idx: int, theta: float) -> None:
    self.applyc(ops.RotationY(theta), ctl, idx, "cry", val=theta)

  def crz(self, ctl: int, idx: int, theta: float) -> None:
    self.applyc(ops.RotationZ(theta), ctl, idx, "crz", val=theta)

  def unitary(self, op, idx) -> None:
    """General k-qubit operator (circuit.py:283-284): O(4^k 2^n) full-matrix math in the
    reference; not part of the gate-application hot path.  Supported for 2x2 only."""
    m = np.asarray(op)
    if m.shape != (2, 2):
      raise NotImplementedError("qc.unitary beyond 2x2 is outside the gate-application path")
    self.apply1(ops.Gate(m, "unitary"), idx, "unitary")

  # --- measurement -----------------------------------------------------------------
  def measure_bit(self, idx: int, tostate: int = 0, collapse: bool = True):   # circuit.py:287-292
    """Probability of qubit `idx` being `tostate`; optionally project + renormalise.  The
    reference goes through a 4^n density matrix (ops.py:426-460); here it is one device
    reduction and, for the collapse, one diagonal gate diag(1/sqrt p, 0) on that qubit."""
    prob = self.psi.weight_of_qubit(idx, 1 if tostate == 1 else 0)   # trace(P rho), computed directly
    if collapse:
      assert math.sqrt(max(prob, 0.0)) > 1e-10, "Measurement collapses to 0.0."
      s = 1.0 / math.sqrt(prob)
      proj = [[s, 0.0], [0.0, 0.0]] if tostate == 0 else [[0.0, 0.0], [0.0, s]]
      self.dev.xg_apply1(idx, np.array(proj, dtype=np.complex128))
    return prob, self.psi

  def pauli_expectation(self, idx: int) -> float:
    p0, _ = self.measure_bit(idx, 0, False)
    return p0 - (1 - p0)

  # --- composites ---------------------------------------------------------------------
  def swap(self, idx0: int, idx1: int) -> None:                           # circuit.py:303-310
    with self.scope(self.ir, f"swap({idx0}, {idx1})"):
      self.cx(idx1, idx0)
      self.cx(idx0, idx1)
      self.cx(idx1, idx0)

  def cswap(self, ctl, idx0, idx1) -> None:                               # circuit.py:312-318
    with self.scope(self.ir, f"cswap({ctl}, {idx0}, {idx1})"):
      self.cx(idx1, idx0)
      self.ccx(ctl, idx0, idx1)
      self.cx(idx1, idx0)

  def qft(self, reg, with_swaps: bool = False) -> None:                   # circuit.py:320-328
    for i in reversed(range(len(reg))):
      self.h(reg[i])
      for j in reversed(range(i)):
        self.cu1(reg[i], reg[j], np.pi / 2 ** (i - j))
    if with_swaps:
      self.flip(reg)

  def inverse_qft(self, reg, with_swaps: bool = False) -> None:           # circuit.py:330-339
    if with_swaps:
      self.flip(reg)
    for idx, r in enumerate(reg):
      self.h(r)
      if idx != len(reg) - 1:
        for y in range(idx, -1, -1):
          self.cu1(reg[idx + 1], reg[y], -np.pi / 2 ** (idx + 1 - y))

  def multi_control(self, ctl, idx1, aux, gate, desc: str = "") -> None:  # circuit.py:341-392
    """Multi-controlled gate with n-1 ancillae: ccx ladder into aux, one controlled gate,
    uncompute.  Controls given as [q] are control-by-0."""
    if aux:
      assert len(aux) >= len(ctl) - 1, "Incorrect number of ancilla qubits."
    gname = getattr(gate, "name", None) or "U"
    with self.scope(self.ir, f"multi-{gname}({ctl}, {idx1}) # {desc})"):
      if not ctl:
        self.apply1(gate, idx1, desc)
        return
      if isinstance(ctl, state.Reg):
        ctl = ctl[:]
      if len(ctl) == 1:
        self.applyc(gate, ctl[0], idx1, desc)
        return
      if len(ctl) == 2:
        self.ccu(ctl[0], ctl[1], idx1, gate, desc)
        return
      self.ccx(ctl[0], ctl[1], aux[0])
      top = 0
      for i in range(2, len(ctl)):
        self.ccx(ctl[i], aux[top], aux[top + 1])
        top += 1
      self.applyc(gate, aux[top], idx1, desc)
      top -= 1
      for i in range(len(ctl) - 1, 1, -1):
        self.ccx(ctl[i], aux[top], aux[top + 1])
        top -= 1
      self.ccx(ctl[0], ctl[1], aux[0])

  def flip(self, reg) -> None:
    for i in range(len(reg) // 2):
      self.swap(reg[i], reg[len(reg) - 1 - i])

  # --- circuits of circuits ---------------------------------------------------------------
  class scope:
    def __init__(self, ir_param, desc: str):
      self.ir = ir_param
      self.desc = desc

    def __enter__(self):
      self.ir.section(self.desc)

    def __exit__(self, t, value, traceback):
      self.ir.end_section()

  def qc(self, qc_parm: "qc", offset: int = 0) -> None:                   # circuit.py:401-412
    for gate in list(qc_parm.ir.gates):
      if gate.is_single():
        self.apply1(gate.gate, gate.idx0 + offset, gate.name, val=gate.val)
      if gate.is_ctl():
        self.applyc(gate.gate, gate.ctl + offset, gate.idx1 + offset, gate.name, val=gate.val)

  def run(self) -> None:                                                  # circuit.py:414-423
    """Apply the recorded gates to this circuit's state without re-recording them.  The
    whole IR is handed to the engine as one batch (one plan, fused passes)."""
    dev = self.dev
    stream = []
    for gate in self.ir.gates:
      if gate.is_single():
        stream.append((1, 0, gate.idx0, gate.gate))
      elif gate.is_ctl():
        stream.append((2, gate.ctl, gate.idx1, gate.gate))
    if stream:
      dev.xg_apply_gates(_cabi.pack_xg_gates(stream))

  def inverse(self) -> "qc":                                              # circuit.py:425-456
    newqc = qc(self.name, eager=False, device=self._device, fusion=self._fusion,
               tile_bits=self._tile_bits)
    for gate in self.ir.gates[::-1]:
      val = -gate.val if gate.val else None
      if gate.is_single():
        newqc.apply1(gate.gate.adjoint(), gate.idx0, gate.name + "*", val=val)
      if gate.is_ctl():
        newqc.applyc(gate.gate.adjoint(), gate.ctl, gate.idx1, gate.name + "*", val=val)
    return newqc

  def control_by(self, ctl: int) -> None:                                 # circuit.py:470-489
    assert not self.eager, "control_by() used in non-eager circuit."
    res = ir.Ir()
    for gate in self.ir.gates:
      if gate.is_single():
        gate.to_ctl(ctl)
        res.add_node(gate)
        continue
      if gate.is_ctl():
        sub = qc("multi", eager=False)
        sub.multi_control([ctl, gate.ctl], gate.idx1, None, gate.gate, gate.desc)
        for g in sub.ir.gates:
          res.add_node(g)
    self.ir = res

  def sub(self, name: str = "") -> "qc":
    sub = qc(f"inner_{self.sub_circuits}{name}", eager=False)
    self.sub_circuits += 1
    return sub

  # --- output ----------------------------------------------------------------------------
  def stats(self) -> str:
    return ("Circuit Statistics\n" + "  Qubits: {}\n".format(self.nbits) +
            "  Gates : {}\n".format(self.ir.ngates))

  def libq(self) -> str:
    """C++ program for this circuit's IR against libq.h (dumpers.libq)."""
    return dumpers.libq(self.ir)

  def qasm(self) -> str:
    """OPENQASM 2.0 text for this circuit's IR (dumpers.qasm)."""
    return dumpers.qasm(self.ir)

  def cirq(self) -> str:
    """Cirq script for this circuit's IR (dumpers.cirq)."""
    return dumpers.cirq(self.ir)

  def dump_to_file(self, libq: str = None, qasm: str = None, cirq: str = None) -> None:
    """circuit.py:505-520: the reference takes the file names from absl flags (--libq, --qasm, --cirq);
    here they are arguments."""
    for path, text in ((libq, self.libq), (qasm, self.qasm), (cirq, self.cirq)):
      if path:
        with open(path, "w") as f:
          print(text(), file=f)

  def dump(self, *, desc=None, draw=False, pstate=True) -> None:
    if desc:
      print(desc)
    if self.name:
      print(f"Circuit: {self.name}, Gates: {len(self.ir.gates)}, QBits: {self.nbits}")
    print(self.ir, end="")
    if pstate and (self._dev is not None or self._pending):
      self.psi.dump("Current state")

  def sync(self) -> None:
    if self._dev is not None or self._pending:
      self.dev.sync()

  def close(self) -> None:
    if self._dev is not None:
      self._dev.close()
      self._dev = None
