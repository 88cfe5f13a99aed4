"""make_golden.py -- generate the golden fixtures under tests/golden/ FROM THE REFERENCE.

Run in the build container only (it imports /root/reference and uses the reference
builds in oracle/_ref/):

    make -C oracle ref && python tests/golden/make_golden.py

Everything written here is small (n <= 12 qubits) and committed; tests never need
/root/reference at run time.  What is recorded and where it comes from:

  dense_*.npz        seeded random apply1/applyc streams.  `final_spec` is the
                     reference's pure-Python definition State.apply1/applyc
                     (src/lib/state.py:80-125), `final_xgates` is the reference's
                     xgates build with bit_width=128 (src/lib/xgates.cc:23-67),
                     `final_xgates_f` the complex64 path.
  acceleration.npz   the two gate sequences of circuit_test.py:69-107 (including
                     the negative control indices) and their final states.
  circ_*.npz         gate streams recorded as IR from the reference's circuit.qc
                     (qft, larose, supremacy.py, grover.py, multi_control, swap,
                     cswap, inverse_qft...) + the final state of running them.
  libq_*.npz         stock libq (float, sparse) and the all-double libq build on
                     the oracle-safe gate subset, read from the qureg struct.
  libq_*_test.out    stdout of the reference's three libq test mains.
  qft6_libq.cc       dumpers.libq() text for the 6-qubit QFT (configs[0]).
  *.qasm, *_cirq.py.txt  SURVEY 8(f)4: dumpers.qasm() / dumpers.cirq() text for the 6-qubit QFT and a
                     two-register circuit with every gate name the cirq emitter knows.
  order_N15_a4.npz   SURVEY 8(f)1: the gate stream the reference's order_finding.py builds
                     for N=15, a=4 (18 qubits, 12 353 IR nodes), and what its xgates build
                     computes from it: the basis states with p > 0.01 that the script's
                     measurement loop prints (order_finding.py:185-202), 64 sampled
                     amplitudes and the norm (the 4 MiB state itself is not stored).
"""
import math
import os
import random
import subprocess
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from absl import flags  # noqa: E402
from src.lib import circuit, ops, state  # noqa: E402  (prints the no-libxgates banner: fine)

flags.FLAGS(["make_golden", "--tensor_width=128"])

from oracle import oracle  # noqa: E402

XG = oracle.RefXgates()


def gate_pool(rng):
  th = rng.uniform(-math.pi, math.pi)
  return [
      ops.Hadamard(), ops.PauliX(), ops.PauliY(), ops.PauliZ(), ops.Sgate(), ops.Tgate(),
      ops.Vgate(), ops.Yroot(), ops.U1(th), ops.RotationX(th), ops.RotationY(th),
      ops.RotationZ(th), ops.Vgate().adjoint(), ops.U3(th, 0.3, -1.2),
  ]


def save_stream(path, n, psi0, stream, **finals):
  kind = np.array([g[0] for g in stream], dtype=np.int32)
  ctl = np.array([g[1] if g[1] is not None else 0 for g in stream], dtype=np.int32)
  tgt = np.array([g[2] for g in stream], dtype=np.int32)
  mats = np.array([np.asarray(g[3], dtype=np.complex128).reshape(4) for g in stream])
  names = np.array([g[4] if len(g) > 4 and g[4] else "" for g in stream])
  np.savez_compressed(path, nbits=n, psi0=psi0, kind=kind, ctl=ctl, tgt=tgt, mats=mats,
                      names=names, **finals)
  print("wrote", os.path.relpath(path, ROOT), len(stream), "gates")


def random_stream(n, ngates, seed, allow_neg_ctl=False):
  rng = np.random.default_rng(seed)
  stream = []
  for _ in range(ngates):
    g = gate_pool(rng)[rng.integers(0, 14)]
    t = int(rng.integers(0, n))
    if rng.random() < 0.5 or n == 1:
      stream.append((1, None, t, np.array(g), g.name))
    else:
      lo = -n if allow_neg_ctl else 0
      c = t
      while c == t:
        c = int(rng.integers(lo, n))
      stream.append((2, c, t, np.array(g), g.name))
  return stream


def dense_cases():
  for n, ngates, seed, neg in [(1, 6, 1, False), (2, 20, 2, False), (5, 60, 3, True),
                               (8, 120, 4, True), (10, 200, 5, False), (12, 150, 6, True)]:
    rng = np.random.default_rng(100 + seed)
    psi0 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi0 /= np.linalg.norm(psi0)
    stream = random_stream(n, ngates, seed, neg)
    finals = {}
    if n <= 8:  # the pure-Python definition is slow
      s = state.State(psi0.copy())
      for kind, c, t, m, _ in stream:
        if kind == 1:
          s.apply1(m.reshape(2, 2), t)
        else:
          s.applyc(m.reshape(2, 2), c, t)
      finals["final_spec"] = np.asarray(s)
    x = psi0.copy()
    XG.run(x, n, [g[:4] for g in stream])
    finals["final_xgates"] = x
    xf = psi0.astype(np.complex64)
    XG.run(xf, n, [g[:4] for g in stream])
    finals["final_xgates_f"] = xf
    save_stream(os.path.join(HERE, f"dense_n{n}.npz"), n, psi0, stream, **finals)


def acceleration_case():
  """circuit_test.py:69-107, recorded gate by gate."""
  out = {}
  # part 1
  psi = state.bitstring(1, 0, 1, 0)
  stream = []
  for i in range(4):
    for g in (ops.PauliX(), ops.PauliY(), ops.PauliZ(), ops.Hadamard()):
      psi.apply1(g, i)
      stream.append((1, None, i, np.array(g), g.name))
    if i:
      psi.applyc(ops.U1(1.1), 0, i)
      stream.append((2, 0, i, np.array(ops.U1(1.1)), "cu1"))
  save_stream(os.path.join(HERE, "acceleration_1.npz"), 4, np.asarray(state.bitstring(1, 0, 1, 0)),
              stream, final_spec=np.asarray(psi))
  # part 2 (negative controls)
  psi = state.bitstring(1, 0, 1, 0, 1)
  stream = []
  for n in range(5):
    psi.apply1(ops.Hadamard(), n)
    stream.append((1, None, n, np.array(ops.Hadamard()), "h"))
    for sign in (1.0, -1.0):
      for i in range(0, 5):
        u = ops.U1(sign * math.pi / float(2 ** (i + 1)))
        psi.applyc(u, n - (i + 1), n)
        stream.append((2, n - (i + 1), n, np.array(u), "cu1"))
    psi.apply1(ops.Hadamard(), n)
    stream.append((1, None, n, np.array(ops.Hadamard()), "h"))
  x = np.asarray(state.bitstring(1, 0, 1, 0, 1)).copy()
  XG.run(x, 5, [g[:4] for g in stream])
  save_stream(os.path.join(HERE, "acceleration_2.npz"), 5,
              np.asarray(state.bitstring(1, 0, 1, 0, 1)), stream,
              final_spec=np.asarray(psi), final_xgates=x)
  return out


def ir_stream(qc):
  stream = []
  for g in qc.ir.gates:
    if g.is_single():
      stream.append((1, None, g.idx0, np.array(g.gate), g.name, g.val))
    elif g.is_ctl():
      stream.append((2, g.ctl, g.idx1, np.array(g.gate), g.name, g.val))
  return stream


def save_circuit(name, qc, psi0, extra=None):
  stream = ir_stream(qc)
  n = int(math.log2(len(psi0)))
  vals = np.array([np.nan if g[5] is None else g[5] for g in stream], dtype=np.float64)
  x = np.asarray(psi0).astype(np.complex128).copy()
  XG.run(x, n, [g[:4] for g in stream])
  extra = dict(extra or {})
  extra["vals"] = vals
  save_stream(os.path.join(HERE, f"circ_{name}.npz"), n, np.asarray(psi0), stream,
              final_xgates=x, **extra)
  return x


class Capture(circuit.qc):
  """circuit.qc that records IR *and* executes (for scripts that build eager circuits)."""
  made = []

  def __init__(self, *a, **k):
    super().__init__(*a, **k)
    self.build_ir = True
    Capture.made.append(self)


def circuit_cases():
  # configs[0]: 6-qubit QFT of |101101> (SURVEY.md 8d input 1)
  qc = circuit.qc("qft6", eager=False)
  r = qc.reg(6, 0b101101)
  psi0 = np.asarray(qc.psi).copy()
  qc.qft(r)
  final = save_circuit("qft6", qc, psi0)
  qc.run()
  assert np.allclose(np.asarray(qc.psi), final, atol=1e-12)
  from src.lib import dumpers
  with open(os.path.join(HERE, "qft6_libq.cc"), "w") as f:
    f.write(dumpers.libq(qc.ir))

  # QFT / inverse QFT with swaps on a random state, 9 qubits
  rng = np.random.default_rng(7)
  qc = circuit.qc("qft9", eager=False)
  r = qc.reg(9, 0)
  psi0 = rng.normal(size=512) + 1j * rng.normal(size=512)
  psi0 /= np.linalg.norm(psi0)
  qc.qft(r, with_swaps=True)
  qc.inverse_qft(r, with_swaps=False)
  save_circuit("qft9_swaps_iqft", qc, psi0)

  # larose_benchmark.py:45-54 gate stream, small
  for n, depth in [(8, 3), (11, 2)]:
    qc = circuit.qc("larose", eager=False)
    qc.reg(n, 5, name="q")
    psi0 = np.asarray(qc.psi).copy()
    for _ in range(depth):
      for bit in range(n):
        qc.h(bit)
        qc.v(bit)
        if bit > 0:
          qc.cx(bit, 0)
    save_circuit(f"larose_n{n}_d{depth}", qc, psi0)

  # supremacy.py:123-158 + 208-240, seeded
  import src.supremacy as supremacy
  supremacy.circuit.qc = Capture
  try:
    for n, depth in [(10, 8), (12, 10)]:
      random.seed(0)
      Capture.made.clear()
      states = supremacy.build_circuit(n, depth)
      supremacy.sim_circuit(states, n, depth, 53, 20)
      qc = Capture.made[-1]
      psi0 = np.zeros(1 << n, dtype=np.complex128)
      psi0[0] = 1
      final = save_circuit(f"supremacy_n{n}_d{depth}", qc, psi0)
      assert np.allclose(np.asarray(qc.psi), final, atol=1e-10)
  finally:
    supremacy.circuit.qc = circuit.qc

  # grover.py:124-168, seeded: nbits=4 -> 8 qubits total
  import src.grover as grover
  grover.circuit.qc = Capture
  try:
    for nb in (3, 4):
      np.random.seed(0)
      Capture.made.clear()
      grover.run_experiment_circuit(nb)
      qc = Capture.made[-1]
      n = qc.psi.nbits
      # initial state: reg(nb,0) (x) |1> (x) reg(nb-1,0)
      psi0 = np.zeros(1 << n, dtype=np.complex128)
      psi0[1 << (nb - 1)] = 1
      final = save_circuit(f"grover_{nb}", qc, psi0)
      assert np.allclose(np.asarray(qc.psi), final, atol=1e-10)
      maxbits, maxprob = qc.psi.maxprob()
      np.savez_compressed(os.path.join(HERE, f"grover_{nb}_readout.npz"),
                          maxbits=np.array(maxbits), maxprob=maxprob)
  finally:
    grover.circuit.qc = circuit.qc

  # composites: toffoli, swap, cswap, multi_control incl. control-by-0, ccu1, rotations
  qc = circuit.qc("composites", eager=False)
  r = qc.reg(9, 0)
  psi0 = rng.normal(size=512) + 1j * rng.normal(size=512)
  psi0 /= np.linalg.norm(psi0)
  qc.toffoli(0, 3, 5)
  qc.swap(1, 7)
  qc.cswap(2, 4, 8)
  qc.multi_control([0, [1], 2, [3]], 8, [4, 5, 6, 7][:3], ops.PauliX(), "mc")
  qc.ccu1(0, 1, 2, 0.77)
  qc.rx(3, 0.3)
  qc.cry(3, 4, -1.3)
  qc.crz(8, 0, 2.1)
  qc.cx0(6, 2)
  qc.sdag(5)
  qc.cvdag(1, 6)
  qc.cyroot(7, 0)
  save_circuit("composites9", qc, psi0)
  inv = qc.inverse()
  save_circuit("composites9_inverse", inv, psi0)


SAFE = ["x", "y", "z", "h", "t", "u1", "cu1", "cx", "cz", "ccx"]


def libq_cases():
  for width, nops, seed in [(4, 40, 11), (8, 120, 12), (12, 200, 13)]:
    rng = np.random.default_rng(seed)
    ops_list = [("walsh", width)] if seed != 11 else []
    for _ in range(nops):
      name = SAFE[rng.integers(0, len(SAFE))]
      qs = [int(q) for q in rng.permutation(width)[:3]]
      if name in ("x", "y", "z", "h", "t"):
        ops_list.append((name, qs[0]))
      elif name == "u1":
        ops_list.append((name, qs[0], float(rng.uniform(-3, 3))))
      elif name == "cu1":
        ops_list.append((name, qs[0], qs[1], float(rng.uniform(-3, 3))))
      elif name in ("cx", "cz"):
        ops_list.append((name, qs[0], qs[1]))
      else:
        ops_list.append((name, qs[0], qs[1], qs[2]))
    init = int(rng.integers(0, 1 << width))
    f = oracle.RefLibq(False).run_dense(width, init, ops_list)
    d = oracle.RefLibq(True).run_dense(width, init, ops_list)
    names = np.array([o[0] for o in ops_list])
    args = np.zeros((len(ops_list), 3), dtype=np.int32)
    gam = np.zeros(len(ops_list))
    for k, o in enumerate(ops_list):
      ints = [a for a in o[1:] if isinstance(a, int)]
      args[k, :len(ints)] = ints
      fl = [a for a in o[1:] if isinstance(a, float)]
      if fl:
        gam[k] = fl[0]
    path = os.path.join(HERE, f"libq_w{width}.npz")
    np.savez_compressed(path, width=width, init=init, names=names, args=args, gamma=gam,
                        final_float=f, final_double=d)
    print("wrote", os.path.relpath(path, ROOT))
  for t in ("libq_test", "libq_arith_test", "libq_order22_test"):
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", t)], capture_output=True, text=True,
                         check=True).stdout
    with open(os.path.join(HERE, t + ".out"), "w") as f:
      f.write(out)
    print("wrote", t + ".out")
  # fingerprint of the libq:: call sequence of the reference's generated adder test (the file
  # itself is reference source and is not copied): our transpiler must reproduce it exactly
  import hashlib
  lines = [l.strip() for l in open(os.path.join(REF, "src/libq/libq_arith_test.cc")) if l.strip().startswith("libq::")]
  with open(os.path.join(HERE, "libq_arith_test.calls.sha256"), "w") as f:
    f.write(f"{hashlib.sha256(chr(10).join(lines).encode()).hexdigest()} {len(lines)}\n")


def emitter_cases():
  """the reference's qasm / cirq text (dumpers.py:20-37, 89-161) for the 6-qubit QFT and
  for a two-register circuit using every gate name its cirq emitter knows."""
  from src.lib import dumpers
  qc = circuit.qc("qft6", eager=False)
  r = qc.reg(6, 0b101101)
  qc.qft(r)
  with open(os.path.join(HERE, "qft6.qasm"), "w") as f:
    f.write(dumpers.qasm(qc.ir))
  # (dumpers.cirq cannot print cu1 / cv: it reads op.idx0 of a controlled node, which ir.py asserts
  # against -- so the cirq golden is limited to the names that work: h x y z cx cz u1)
  qc = circuit.qc("mixed", eager=False)
  a = qc.reg(3, 0b101, name="a")
  b = qc.reg(2, 0, name="b")
  qc.h(a[0]); qc.x(a[1]); qc.y(a[2]); qc.z(b[0])
  qc.cx(a[0], b[1]); qc.cz(a[2], b[0]); qc.u1(b[1], math.pi / 8)
  with open(os.path.join(HERE, "mixed_cirq.py.txt"), "w") as f:
    f.write(dumpers.cirq(qc.ir))
  qc.cu1(a[1], b[0], -math.pi / 4); qc.cv(a[0], a[2]); qc.cu1(b[1], a[0], 0.3)
  with open(os.path.join(HERE, "mixed.qasm"), "w") as f:
    f.write(dumpers.qasm(qc.ir))
  print("wrote qft6.qasm mixed.qasm mixed_cirq.py.txt")


def order_finding_case(number=15, a=4):
  """order_finding.py:152-183 recorded as IR through the reference's own functions, then run with
  its xgates build.  Only the readout is stored (see the module docstring)."""
  from src import order_finding as of    # defines its --N / --a flags on import: harmless here
  nbits = number.bit_length()
  n = 4 * nbits + 2
  qc = circuit.qc("order_finding", eager=False)
  aux = qc.reg(nbits + 2)
  up = qc.reg(nbits * 2)
  down = qc.reg(nbits)
  qc.h(up)
  qc.x(down[0])
  for i in range(nbits * 2):
    of.cmultmodn(qc, up[i], down, aux, int(a ** (2 ** i)), number, nbits)
  of.inverse_qft(qc, up, 2 * nbits, with_swaps=True)
  stream = ir_stream(qc)
  psi = np.zeros(1 << n, dtype=np.complex128)
  psi[0] = 1.0
  XG.run(psi, n, [g[:4] for g in stream])
  p = np.abs(psi) ** 2
  keep = np.nonzero(p > 0.01)[0]
  rng = np.random.default_rng(15)
  sample = np.sort(rng.choice(1 << n, size=64, replace=False))
  psi0 = np.zeros(2, dtype=np.complex128)   # placeholder: the initial state is |0...0>
  save_stream(os.path.join(HERE, f"order_N{number}_a{a}.npz"), n, psi0, stream,
              labels=keep.astype(np.int64), probs=p[keep], sample_idx=sample.astype(np.int64),
              sample_amp=psi[sample], norm2=float(p.sum()))
  for k in keep:
    bits = [int(b) for b in format(int(k), f"0{n}b")]
    bitslice = bits[nbits + 2: nbits + 2 + nbits * 2][::-1]
    print("  x =", int("".join(map(str, bitslice)), 2), "p =", round(float(p[k]), 4))


if __name__ == "__main__":
  if len(sys.argv) > 1 and sys.argv[1] == "order":
    order_finding_case()
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == "emitters":
    emitter_cases()
    sys.exit(0)
  dense_cases()
  acceleration_case()
  circuit_cases()
  libq_cases()
  order_finding_case()
  emitter_cases()
