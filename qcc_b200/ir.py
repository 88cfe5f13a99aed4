"""Gate IR recorded by circuit.qc when it is not (only) executing eagerly.

Mirrors what the reference's src/lib/ir.py:11-154 exposes to its users (dumpers, qc.qc(),
qc.inverse(), qc.control_by()): an ordered list of nodes, each a single-qubit gate, a
controlled gate, or a section marker, plus the register table the transpiler needs."""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Any, List, Optional, Tuple


class Op(enum.Enum):
  UNK = 0
  SINGLE = 1
  CTL = 2
  SECTION = 3
  END_SECTION = 4


@dataclass
class Node:
  opcode: Op
  name: Optional[str]
  _idx0: Any
  _idx1: Any
  gate: Any = None
  val: Optional[float] = None

  def is_single(self) -> bool:
    return self.opcode == Op.SINGLE

  def is_ctl(self) -> bool:
    return self.opcode == Op.CTL

  def is_gate(self) -> bool:
    return self.opcode in (Op.SINGLE, Op.CTL)

  def is_section(self) -> bool:
    return self.opcode == Op.SECTION

  def is_end_section(self) -> bool:
    return self.opcode == Op.END_SECTION

  @property
  def desc(self):
    return self.name

  @property
  def idx0(self) -> int:
    if not self.is_single():
      raise AssertionError("idx0 is only defined for single-qubit gates")
    return self._idx0

  @property
  def ctl(self) -> int:
    if not self.is_ctl():
      raise AssertionError("ctl is only defined for controlled gates")
    return self._idx0

  @property
  def idx1(self) -> int:
    if not self.is_ctl():
      raise AssertionError("idx1 is only defined for controlled gates")
    return self._idx1

  def to_ctl(self, ctl: int) -> None:
    """Turn a single-qubit node into the same gate controlled by `ctl` (ir.py:44-48)."""
    self.opcode = Op.CTL
    self._idx1 = self._idx0
    self._idx0 = ctl
    self.name = "c" + (self.name or "*unk*")

  def __str__(self) -> str:
    from qcc_b200 import helper
    nm = self.name or "*unk*"
    if self.is_single():
      s = f"{nm}({self._idx0})"
    elif self.is_ctl():
      s = f"{nm}({self._idx0}, {self._idx1})"
    elif self.is_section():
      return f"|-- {nm} ---"
    else:
      return ""
    if self.val:
      s += f"({helper.pi_fractions(self.val)})"
    return s


@dataclass
class Ir:
  gates: List[Node] = field(default_factory=list)
  regs: List[Tuple[int, Optional[str], int]] = field(default_factory=list)   # (global idx, reg name, idx in reg)
  regset: List[Tuple[Optional[str], int, Any]] = field(default_factory=list)  # (name, size, Reg)
  nregs: int = 0
  _ngates: int = 0

  @property
  def ngates(self) -> int:
    return self._ngates

  def reg(self, size: int, name, register) -> None:
    self.regset.append((name, size, register))
    for i in range(size):
      self.regs.append((self.nregs + i, name, i))
    self.nregs += size

  def add_node(self, node: Node) -> None:
    self.gates.append(node)
    self._ngates += 1

  def single(self, name, idx0, gate, val=None) -> None:
    self.add_node(Node(Op.SINGLE, name, idx0, None, gate, val))

  def controlled(self, name, idx0, idx1, gate, val=None) -> None:
    self.add_node(Node(Op.CTL, name, idx0, idx1, gate, val))

  def section(self, desc) -> None:
    self.gates.append(Node(Op.SECTION, desc, 0, 0))

  def end_section(self) -> None:
    self.gates.append(Node(Op.END_SECTION, None, 0, 0))

  def __str__(self) -> str:
    out, depth = [], 0
    for node in self.gates:
      if node.is_end_section():
        depth -= 1
        continue
      out.append("  " * depth + str(node))
      if node.is_section():
        depth += 1
    return "\n".join(out) + ("\n" if out else "")
