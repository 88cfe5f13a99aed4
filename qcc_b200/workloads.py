"""Gate streams of the reference workloads named in BASELINE.json, as flat lists of
(kind, ctl, tgt, 2x2) in the reference's python qubit numbering (kind 1 = apply1,
2 = applyc -- what circuit.py:180-215 hands to xgates per call)."""
from __future__ import annotations

import math
import random

import numpy as np

from qcc_b200 import ops


def qft(n: int, reg=None):
  """circuit.py:320-326 (qft without swaps): per i from n-1 down: h(i), then
  cu1(i, j, pi / 2^(i-j)) for j < i.  30 qubits: 30 h + 435 cu1 = 465 gates."""
  reg = list(range(n)) if reg is None else list(reg)
  h = np.asarray(ops.Hadamard())
  out = []
  for i in reversed(range(len(reg))):
    out.append((1, 0, reg[i], h))
    for j in reversed(range(i)):
      out.append((2, reg[i], reg[j], np.asarray(ops.U1(math.pi / 2 ** (i - j)))))
  return out


def inverse_of(stream):
  """circuit.py:423-456: reversed order, adjoint gates."""
  return [(k, c, t, np.asarray(m).conj().T) for k, c, t, m in reversed(stream)]


def larose(n: int, depth: int):
  """larose_benchmark.py:47-54: per depth, per bit: h, v, and cx(bit, 0) for bit > 0.
  28 qubits, depth 28: 784 h + 784 v + 756 cx = 2324 gates."""
  h, v, x = (np.asarray(g) for g in (ops.Hadamard(), ops.Vgate(), ops.PauliX()))
  out = []
  for _ in range(depth):
    for bit in range(n):
      out.append((1, 0, bit, h))
      out.append((1, 0, bit, v))
      if bit > 0:
        out.append((2, bit, 0, x))
  return out


def hsweep(n: int):
  """The headline micro-benchmark: one h on every qubit."""
  h = np.asarray(ops.Hadamard())
  return [(1, 0, q, h) for q in range(n)]


# The eight CZ layouts of supremacy.py:53-97 on its 6 x 6 grid, one digit per qubit: 0 = no gate starts
# here, d = a cz from this qubit to qubit + d (1: right neighbour, 6: the qubit below).  Constant data of the
# reference's workload (configs[4]); the stream is only the reference's stream with this exact table.
_ROW = "000000"
SUPREMACY_PATTERNS = tuple([int(ch) for ch in txt] for txt in (
    ("001000" "100010") * 3,
    ("100010" "001000") * 3,
    _ROW + "060606" + _ROW + "060606" + _ROW + _ROW,
    _ROW + "606060" + _ROW + "606060" + _ROW + _ROW,
    ("000100" "010000") * 3,
    ("010000" "000100") * 3,
    "606060" + _ROW + "060606" + _ROW + "606060" + _ROW,
    "060606" + _ROW + "060606" + _ROW + "060606" + _ROW,
))


def supremacy_layers(n: int, depth: int, rng):
  """The gate grid of supremacy.py:123-158 (build_circuit): depth + 1 layers of n marks ('h', 't', 'u',
  'cz' or None).  One rng.randint(0, 7) per middle layer picks the pattern, exactly as the reference draws."""
  layer = ["h"] * n
  layers = [layer]
  for _ in range(depth - 1):
    nxt = [None] * n
    pat = SUPREMACY_PATTERNS[rng.randint(0, 7)]
    for i in range(min(n, len(pat))):
      if pat[i] and i + pat[i] < n:
        nxt[i] = nxt[i + pat[i]] = "cz"
    for i in range(n):
      if nxt[i] == "cz":
        continue
      if layer[i] == "cz":
        nxt[i] = "u"
      elif layer[i] in ("u", "h"):
        nxt[i] = "t"
    layer = nxt
    layers.append(layer)
  layers.append(["h"] * n)
  return layers


def supremacy(n: int, depth: int, seed: int = 0):
  """The gate stream supremacy.py runs for --nbits n --depth depth after random.seed(seed): build_circuit
  (supremacy.py:123-158) then the gate calls of sim_circuit (:208-240) -- layers 0 .. depth-1 (the closing h
  layer is built but not simulated, :218), 'u' drawn as v or yroot by one more randint each, and cz marks
  paired with the right neighbour and / or the qubit six further on, whichever also carries a mark (:234-239).
  random.Random(seed) is the generator random.seed(seed) installs, so the draws are the reference's."""
  rng = random.Random(seed)
  layers = supremacy_layers(n, depth, rng)
  g = {k: np.asarray(v) for k, v in (("h", ops.Hadamard()), ("t", ops.Tgate()), ("v", ops.Vgate()),
                                     ("yroot", ops.Yroot()), ("z", ops.PauliZ()))}
  out = []
  for d in range(depth):
    s = list(layers[d])
    for i in range(n):
      if s[i] is None:
        continue
      if s[i] in ("h", "t"):
        out.append((1, 0, i, g[s[i]]))
      elif s[i] == "u":
        out.append((1, 0, i, g["v"] if rng.randint(0, 1) == 0 else g["yroot"]))
      else:
        if i < n - 1 and s[i + 1] == "cz":
          out.append((2, i, i + 1, g["z"]))
          s[i + 1] = None
        if i < n - 6 and s[i + 6] == "cz":
          out.append((2, i, i + 6, g["z"]))
          s[i + 6] = None
  return out


def grover_circuit(nbits: int, marked: int = None, *, device: int = 0, qc_factory=None, iterations: int = None):
  """Circuit Grover of grover.py:124-168 on the device-resident qc: `nbits` search qubits,
  one ancilla in |1>, nbits-1 helper qubits for multi_control (2*nbits qubits in total;
  nbits=16 is configs[3], 32 qubits).  `marked` defaults to np.random.randint(0, 2^nbits),
  the draw make_f1 makes (grover.py:20-27).  Returns (qc, marked_bits)."""
  from qcc_b200 import circuit, helper
  if marked is None:
    marked = int(np.random.randint(0, 1 << nbits))
  bits = helper.val2bits(marked, nbits)
  qc = qc_factory("Grover") if qc_factory else circuit.qc("Grover", device=device)
  reg = qc.reg(nbits, 0)
  qc.reg(1, 1)
  aux = qc.reg(nbits - 1, 0)
  if iterations is None:
    iterations = int(math.pi / 4 * math.sqrt(2 ** nbits))
  idx = list(range(nbits))
  x, z = ops.PauliX(), ops.PauliZ()

  def multi_masked(gate, allow):
    for i in idx:
      if bits[i] == allow:
        qc.apply1(gate, i, gate.name)

  qc.h(list(range(nbits + 1)))
  for _ in range(iterations):
    multi_masked(x, 0)                                             # phase inversion
    qc.multi_control(reg, nbits, aux, x, "Phase Inversion")
    multi_masked(x, 0)
    qc.h(idx)                                                      # inversion about the mean
    qc.x(idx)
    qc.multi_control(reg, nbits, aux, z, "Mean Inversion")
    qc.x(idx)
    qc.h(idx)
  return qc, bits


def qft_adder(n: int, init_a: int, init_b: int, factor: float = 1.0, *, eager: bool = True, qc_factory=None):
  """a + factor * b with the QFT adder of arith_quantum.py:43-75 on two (n+1)-qubit registers
  (n = 12, a = 2, b = 3 is the 26-qubit circuit behind src/libq/libq_arith_test.cc).
  Returns (qc, a_register, b_register); the sum is read from the a register, least
  significant bit first (arith_quantum.py:36)."""
  from qcc_b200 import circuit, helper
  qc = qc_factory("qadd") if qc_factory else circuit.qc("qadd", eager=eager)
  a = qc.reg(n + 1, helper.val2bits(init_a, n)[::-1], name="a")
  b = qc.reg(n + 1, helper.val2bits(init_b, n)[::-1], name="b")

  def qft_step(reg, k):                       # arith_quantum.py:43-46
    qc.h(reg[k])
    for i in range(k):
      qc.cu1(reg[k - (i + 1)], reg[k], math.pi / float(2 ** (i + 1)))

  def evolve(k):                              # arith_quantum.py:49-52
    for i in range(k + 1):
      qc.cu1(b[k - i], a[k], factor * math.pi / float(2 ** i))

  def inverse_qft_step(reg, k):               # arith_quantum.py:55-58
    for i in range(k):
      qc.cu1(reg[i], reg[k], -1 * math.pi / float(2 ** (k - i)))
    qc.h(reg[k])

  for i in range(n + 1):
    qft_step(a, n - i)
  for i in range(n + 1):
    evolve(n - i)
  for i in range(n + 1):
    inverse_qft_step(a, i)
  return qc, a, b


# ---------------------------------------------------------------------------------------------
# Order finding: the phase-estimation circuit of order_finding.py:152-183 --
# Beauregard-style modular multiplication out of Fourier-space adders -- on the device-resident qc,
# with the readout of order_finding.py:185-202 done by a device-side listing instead of a Python
# scan over all 2^n bit strings.
# ---------------------------------------------------------------------------------------------
def _modinv(a: int, m: int) -> int:
  """a^-1 mod m (order_finding.py:33-50)."""
  g, x, _ = _egcd(a % m, m)
  assert g == 1, f"modular inverse ({a}, {m}) does not exist"
  return x % m


def _egcd(a: int, b: int):
  if a == 0:
    return b, 0, 1
  g, y, x = _egcd(b % a, a)
  return g, x - (b // a) * y, y


def _fourier_angles(a: int, n: int):
  """Rotation angle of each qubit of an n-qubit Fourier-space register when the classical number
  `a` is added to it (order_finding.py:53-62): qubit n-1-i collects a's bits i.. with halving weights."""
  out = [0.0] * n
  for i in range(n):
    acc = 0.0
    for j in range(i, n):
      if a & (1 << (n - j - 1)):
        acc += 2.0 ** (-(j - i))
    out[n - i - 1] = acc * math.pi
  return out


def order_finding(number: int, a: int, *, eager: bool = True, qc_factory=None):
  """Quantum order finding for `a` modulo `number` (order_finding.py:152-183): 4 * nbits + 2
  qubits with nbits = number.bit_length() -- N=15, a=4: 18 qubits; N=21, a=11: 22 qubits (the circuit
  behind src/libq/libq_order22_test.cc).  Returns (qc, aux, up, down); `up` holds the phase."""
  from qcc_b200 import circuit
  nbits = number.bit_length()
  qc = qc_factory("order_finding") if qc_factory else circuit.qc("order_finding", eager=eager)
  aux = qc.reg(nbits + 2)          # adder / multiplier work register (order_finding.py:165)
  up = qc.reg(nbits * 2)           # phase register (:168)
  down = qc.reg(nbits)             # the number being multiplied (:171)
  _modinv(int(a), number)

  def add(q, val, n, factor):                       # order_finding.py:65-69
    for k, angle in enumerate(_fourier_angles(val, n)):
      qc.u1(q[k], factor * angle)

  def cadd(q, ctl, val, n, factor):                 # :72-76
    for k, angle in enumerate(_fourier_angles(val, n)):
      qc.cu1(ctl, q[k], factor * angle)

  def ccadd(q, c1, c2, val, n, factor):             # :79-84
    for k, angle in enumerate(_fourier_angles(val, n)):
      qc.ccu1(c1, c2, q[k], factor * angle)

  def cc_add_mod(q, c1, c2, anc, val, n):           # :95-109
    ccadd(q, c1, c2, val, n, 1.0)
    add(q, number, n, -1.0)
    qc.inverse_qft(q[:n])
    qc.cx(q[n - 1], anc)
    qc.qft(q[:n])
    cadd(q, anc, number, n, 1.0)
    ccadd(q, c1, c2, val, n, -1.0)
    qc.inverse_qft(q[:n])
    qc.cx0(q[n - 1], anc)
    qc.qft(q[:n])
    ccadd(q, c1, c2, val, n, 1.0)

  def cc_add_mod_inverse(q, c1, c2, anc, val, n):   # :112-126
    ccadd(q, c1, c2, val, n, -1.0)
    qc.inverse_qft(q[:n])
    qc.cx0(q[n - 1], anc)
    qc.qft(q[:n])
    ccadd(q, c1, c2, val, n, 1.0)
    cadd(q, anc, number, n, -1.0)
    qc.inverse_qft(q[:n])
    qc.cx(q[n - 1], anc)
    qc.qft(q[:n])
    add(q, number, n, 1.0)
    ccadd(q, c1, c2, val, n, -1.0)

  def cmult_mod(ctl, mult):                         # :129-149
    n = nbits
    qc.qft(aux[:n + 1])
    for i in range(n):
      cc_add_mod(aux, down[i], ctl, aux[n + 1], ((2 ** i) * mult) % number, n + 1)
    qc.inverse_qft(aux[:n + 1])
    for i in range(n):
      qc.cswap(ctl, down[i], aux[i])
    inv = _modinv(mult, number)
    qc.qft(aux[:n + 1])
    for i in reversed(range(n)):
      cc_add_mod_inverse(aux, down[i], ctl, aux[n + 1], ((2 ** i) * inv) % number, n + 1)
    qc.inverse_qft(aux[:n + 1])

  qc.h(up)                                          # :177-181
  qc.x(down[0])
  for i in range(nbits * 2):
    cmult_mod(up[i], int(a ** (2 ** i)))
  qc.inverse_qft(up[:2 * nbits], with_swaps=True)
  return qc, aux, up, down


def order_readout(qc, number: int, a: int, threshold: float = 0.01):
  """order_finding.py:185-202 without the scan over 2^n bit strings: every basis state with
  probability > threshold, as (x, phase, prob, order guess r, factor guesses)."""
  import fractions
  from qcc_b200 import helper
  nbits = number.bit_length()
  n = 4 * nbits + 2
  labels, amps, _ = qc.psi.nonzero(threshold)
  out = []
  for idx, amp in zip(labels, amps):
    bits = helper.val2bits(int(idx), n)
    bitslice = bits[nbits + 2: nbits + 2 + nbits * 2][::-1]
    x = helper.bits2val(bitslice)
    phase = helper.bits2frac(bitslice)
    r = fractions.Fraction(phase).limit_denominator(8).denominator
    guesses = [math.gcd(a ** (r // 2) - 1, number), math.gcd(a ** (r // 2) + 1, number)]
    out.append((x, phase, float(abs(amp) ** 2), r, guesses))
  return out
