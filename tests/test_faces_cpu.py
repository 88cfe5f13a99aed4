"""CPU: host-side logic of the python face (IR recording, composites, transpiler text) and
that programs written against the reference's libq header compile and link against ours."""
import os
import subprocess

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, load_golden, stream_of
from qcc_b200 import circuit, helper, ops


def ir_stream(qc):
  out = []
  for g in qc.ir.gates:
    if g.is_single():
      out.append((1, 0, g.idx0, np.asarray(g.gate), g.name, g.val))
    elif g.is_ctl():
      out.append((2, g.ctl, g.idx1, np.asarray(g.gate), g.name, g.val))
  return out


def assert_same_stream(ours, z):
  ref = stream_of(z)
  assert len(ours) == len(ref)
  for (k, c, t, m, name, _), (rk, rc, rt, rm), rname in zip(ours, ref, z["names"]):
    assert (k, t) == (rk, rt)
    if k == 2:
      assert c == rc
    assert np.abs(np.asarray(m) - rm).max() < 1e-15, (name, rname)
    assert (name or "") == str(rname)


def test_qft6_ir_and_transpiler_text_match_the_reference():
  """configs[0]: same IR as the reference's circuit.qc and byte-identical dumpers.libq text."""
  qc = circuit.qc("qft6", eager=False)
  r = qc.reg(6, 0b101101)
  qc.qft(r)
  assert_same_stream(ir_stream(qc), load_golden("circ_qft6.npz"))
  want = open(os.path.join(GOLDEN, "qft6_libq.cc")).read()
  assert qc.libq() == want


def test_larose_and_composite_streams_match_the_reference():
  qc = circuit.qc("larose", eager=False)
  qc.reg(8, 5, name="q")
  for _ in range(3):
    for bit in range(8):
      qc.h(bit)
      qc.v(bit)
      if bit > 0:
        qc.cx(bit, 0)
  assert_same_stream(ir_stream(qc), load_golden("circ_larose_n8_d3.npz"))

  qc = circuit.qc("composites", eager=False)
  qc.reg(9, 0)
  qc.toffoli(0, 3, 5)
  qc.swap(1, 7)
  qc.cswap(2, 4, 8)
  qc.multi_control([0, [1], 2, [3]], 8, [4, 5, 6], ops.PauliX(), "mc")
  qc.ccu1(0, 1, 2, 0.77)
  qc.rx(3, 0.3)
  qc.cry(3, 4, -1.3)
  qc.crz(8, 0, 2.1)
  qc.cx0(6, 2)
  qc.sdag(5)
  qc.cvdag(1, 6)
  qc.cyroot(7, 0)
  z = load_golden("circ_composites9.npz")
  ours = ir_stream(qc)
  ref = stream_of(z)
  assert len(ours) == len(ref) == 68
  for (k, c, t, m, _, _), (rk, rc, rt, rm) in zip(ours, ref):
    assert (k, t) == (rk, rt) and (k == 1 or c == rc)
    assert np.abs(np.asarray(m) - rm).max() < 1e-12      # sqrtm of X agrees to rounding
  inv = qc.inverse()
  zi = load_golden("circ_composites9_inverse.npz")
  for (k, c, t, m, _, _), (rk, rc, rt, rm) in zip(ir_stream(inv), stream_of(zi)):
    assert (k, t) == (rk, rt) and (k == 1 or c == rc)
    assert np.abs(np.asarray(m) - rm).max() < 1e-12


@pytest.mark.parametrize("n,depth", [(10, 8), (12, 10)])
def test_supremacy_stream_is_the_reference_stream(n, depth):
  """workloads.supremacy(n, depth, seed) must emit, gate for gate, what supremacy.py's build_circuit +
  sim_circuit (supremacy.py:123-158, 208-240) issue after random.seed(seed): the golden files hold the IR
  the reference recorded for seed 0 (tests/golden/make_golden.py)."""
  from qcc_b200 import workloads
  z = load_golden(f"circ_supremacy_n{n}_d{depth}.npz")
  got = workloads.supremacy(n, depth, seed=0)
  assert len(got) == len(z["kind"])
  for (kind, ctl, tgt, m), k, c, t, gm, name in zip(got, z["kind"], z["ctl"], z["tgt"], z["mats"], z["names"]):
    assert kind == int(k) and tgt == int(t) and (kind == 1 or ctl == int(c)), name
    assert np.array_equal(np.asarray(m).reshape(4), gm), name
  # the pattern table: 8 layouts of the 6 x 6 grid, offsets 1 (right) and 6 (down) only
  assert len(workloads.SUPREMACY_PATTERNS) == 8
  assert all(len(p) == 36 and set(p) <= {0, 1, 6} for p in workloads.SUPREMACY_PATTERNS)


def test_qft_swaps_inverse_qft_stream():
  qc = circuit.qc("qft9", eager=False)
  r = qc.reg(9, 0)
  qc.qft(r, with_swaps=True)
  qc.inverse_qft(r, with_swaps=False)
  assert_same_stream(ir_stream(qc), load_golden("circ_qft9_swaps_iqft.npz"))


def test_multi_control_gate_count():
  """circuit_test.py:152-158: 5 controls -> 41 gates."""
  qc = circuit.qc("multi", eager=False)
  ctl = qc.reg(5, 0)
  aux = qc.reg(4, 0)
  tgt = qc.reg(1, 0)
  qc.multi_control(ctl, tgt[0], aux, ops.PauliX(), "x")
  assert qc.ir.ngates == 41
  assert "Gates : 41" in qc.stats()


def test_helpers_and_reg():
  assert helper.bits2val((1, 0, 1)) == 5 and helper.val2bits(5, 4) == [0, 1, 0, 1]
  assert helper.pi_fractions(np.pi / 8, "M_PI") == "M_PI/8"
  assert helper.pi_fractions(-np.pi, "M_PI") == "-M_PI"
  assert helper.pi_fractions(3 * np.pi / 4) == "3*pi/4"
  assert helper.pi_fractions(0.1234) == "0.1234"
  qc = circuit.qc(eager=False)
  a = qc.reg(3, 0b110)
  b = qc.reg(2, [0, 1])
  assert a.val == [1, 1, 0] and b.val == [0, 1] and b[0] == 3 and qc.nbits == 5
  assert str(a) == "|110>"


def test_non_eager_circuits_never_touch_the_gpu(has_gpu):
  """A transpile-only circuit (larose_benchmark.py:45-55) must work without a device and
  without allocating 2^28 amplitudes anywhere."""
  qc = circuit.qc(eager=False)
  qc.reg(28, 3, name="q")
  for bit in range(28):
    qc.h(bit)
    qc.v(bit)
    if bit:
      qc.cx(bit, 0)
  text = qc.libq()
  assert "libq::new_qureg(0, 28);" in text and text.count("libq::cx(") == 27
  assert qc._dev is None


@pytest.mark.parametrize("src", ["golden/qft6_libq.cc", "libq_progs/bell_u1.cc", "libq_progs/all_gates.cc"])
def test_libq_programs_compile_against_our_header(src, tmp_path):
  exe = tmp_path / "prog"
  cmd = ["g++", "-O1", "-I" + os.path.join(ROOT, "qcc_b200", "libq"), os.path.join(ROOT, "tests", src),
         "-L" + os.path.join(ROOT, "qcc_b200", "lib"), "-lqcc_libq", "-lqcc_b200",
         "-Wl,-rpath," + os.path.join(ROOT, "qcc_b200", "lib"), "-o", str(exe)]
  subprocess.run(cmd, check=True, capture_output=True)
  assert exe.exists()


def test_transpiled_adder_reproduces_the_reference_test_program():
  """arith_quantum.py's 12-bit adder on our surface, transpiled, gives exactly the
  libq:: call sequence of the reference's generated src/libq/libq_arith_test.cc (fingerprint
  recorded by make_golden.py; 280 calls)."""
  import hashlib
  from qcc_b200 import workloads
  qc, _, _ = workloads.qft_adder(12, 2, 3, eager=False)
  lines = [l.strip() for l in qc.libq().splitlines() if l.strip().startswith("libq::")]
  want, count = open(os.path.join(GOLDEN, "libq_arith_test.calls.sha256")).read().split()
  assert len(lines) == int(count)
  assert hashlib.sha256("\n".join(lines).encode()).hexdigest() == want


def test_order_finding_stream_matches_the_reference():
  """order_finding.py:152-183 (N=15, a=4, 18 qubits) written against our surface records
  gate for gate the stream the reference's own functions record (ccu1 -> sqrt expansions, cswap, cx0,
  inverse_qft with swaps), and the oracle reproduces what the reference's xgates build computes from it."""
  from helpers import oracle
  from qcc_b200 import workloads
  z = load_golden("order_N15_a4.npz")
  qc, aux, up, down = workloads.order_finding(15, 4, eager=False)
  assert qc.nbits == 18 == int(z["nbits"]) and (len(aux), len(up), len(down)) == (6, 8, 4)
  ours = ir_stream(qc)
  assert len(ours) == 10297
  assert_same_stream(ours, z)
  psi = np.zeros(1 << 18, dtype=np.complex128)
  psi[0] = 1.0
  oracle.c_run(psi, 18, [g[:4] for g in ours])
  p = np.abs(psi) ** 2
  assert abs(p.sum() - float(z["norm2"])) < 1e-10
  assert np.array_equal(np.nonzero(p > 0.01)[0], z["labels"])
  assert np.abs(p[z["labels"]] - z["probs"]).max() < 1e-10
  assert np.abs(psi[z["sample_idx"]] - z["sample_amp"]).max() < 1e-10


def test_bench_reference_arm_prints_the_contract_line():
  """bench.py --impl reference (the reference's own xgates build timed on the host, no GPU): one JSON
  line with the keys the driver reads.  Small register here; the default is the 30-qubit workload."""
  import json
  import sys
  from helpers import oracle
  if not oracle.have_ref("libxgates.so"):
    pytest.skip("oracle/_ref/libxgates.so not built (make -C oracle ref needs /root/reference)")
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--qubits", "14",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, check=True).stdout
  d = json.loads(out.strip().splitlines()[-1])
  assert d["impl"] == "reference" and d["metric"] == "gate-applies/sec" and d["unit"] == "gates/s"
  assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
  assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
  assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
  assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
  assert d["config"]["workload"] == "qft30" and d["dtype"] == "f64"


def test_qasm_and_cirq_text_match_the_reference():
  """byte-identical OPENQASM text (dumpers.py:20-37) for the QFT-6 circuit and a
  two-register circuit, byte-identical Cirq text (dumpers.py:89-161) for the gate names the reference's
  cirq emitter can actually print (h x y z cx cz u1 -- on cu1 / cv it trips over its own ir.py assert)."""
  import math
  qc = circuit.qc("qft6", eager=False)
  r = qc.reg(6, 0b101101)
  qc.qft(r)
  assert qc.qasm() == open(os.path.join(GOLDEN, "qft6.qasm")).read()
  qc = circuit.qc("mixed", eager=False)
  a = qc.reg(3, 0b101, name="a")
  b = qc.reg(2, 0, name="b")
  qc.h(a[0]); qc.x(a[1]); qc.y(a[2]); qc.z(b[0])
  qc.cx(a[0], b[1]); qc.cz(a[2], b[0]); qc.u1(b[1], math.pi / 8)
  assert qc.cirq() == open(os.path.join(GOLDEN, "mixed_cirq.py.txt")).read()
  qc.cu1(a[1], b[0], -math.pi / 4); qc.cv(a[0], a[2]); qc.cu1(b[1], a[0], 0.3)
  assert qc.qasm() == open(os.path.join(GOLDEN, "mixed.qasm")).read()
  # the controlled-matrix forms, which only our emitter can print: control index first
  text = qc.cirq()
  assert "qc.append(cirq.MatrixGate(m).controlled()(r[1], r[3]))" in text
  assert "m = np.array([(1+1j, 1-1j), (1-1j, 1+1j)]) * 0.5\nqc.append(cirq.MatrixGate(m).controlled()(r[0], r[2]))" in text
  assert "cmath.exp(1j * -pi/4)" in text and text.endswith("print(res_str.encode('utf-8'))\n")


def test_gate_batching_keeps_program_order(monkeypatch):
  """circuit.qc batches gate records on the host (GateBuffer) and hands them to the engine in bulk.  With a
  recording stand-in for the engine handle: every readout, direct engine call, register growth and psi
  replacement sees exactly the gates issued before it, in order, and nothing is handed over twice."""
  from qcc_b200 import _cabi

  log = []

  class FakeDev:
    def __init__(self, n, label=0, device=0, **kw):
      self.nqubits, self.nranks = n, kw.get("nranks", 1)
      log.append(("create", n, label))

    def set_fusion(self, on): pass
    def set_tile_bits(self, k): pass
    def close(self): log.append(("close",))
    def sync(self): log.append(("sync",))

    def xg_apply_buffer(self, buf):
      for k in range(buf.n):
        log.append(("gate", int(buf.kind[k]), int(buf.ctl[k]), int(buf.tgt[k]), complex(buf.m[k][3])))
      buf.n = 0

    def xg_apply1(self, tgt, m): log.append(("direct1", tgt))
    def argmax(self): log.append(("argmax",)); return 0, 1.0
    def amplitude(self, i): log.append(("ampl", i)); return 1.0 + 0j
    def prob_bit_value(self, bit, value): log.append(("weight", bit, value)); return 0.5
    def copy_out(self, first=0, count=None): log.append(("copy_out",)); return np.ones(1 << self.nqubits, dtype=np.complex128)
    def copy_in(self, arr, first=0): log.append(("copy_in", len(arr)))

  monkeypatch.setattr(_cabi, "DeviceState", FakeDev)
  qc = circuit.qc("batch")
  qc.reg(3, 0b101)
  qc.h(0)
  qc.cx(0, 1)
  qc.u1(2, 0.5)
  assert [e[0] for e in log] == ["create"]                     # nothing crossed the ABI yet
  qc.psi.maxprob()
  assert [e[0] for e in log] == ["create", "gate", "gate", "gate", "argmax"]
  assert [(e[1], e[2], e[3]) for e in log if e[0] == "gate"] == [(1, 0, 0), (2, 0, 1), (1, 0, 2)]
  qc.x(1)
  qc.measure_bit(1, 1, collapse=True)                          # readout, then a direct engine call
  assert [e[0] for e in log][5:] == ["gate", "weight", "direct1"]
  qc.t(0)
  qc.reg(1, 1)                                                 # the state grows: queued gates first, then the copy
  qc.h(3)
  qc.psi[0]
  tail = [e[0] for e in log][8:]
  assert tail == ["gate", "copy_out", "close", "create", "copy_in", "gate", "ampl"], tail
  qc.z(2)
  qc.psi = np.ones(16) / 4.0                                   # replaced wholesale: the queued z is dropped with it
  qc.s(1)
  qc.sync()
  assert [e[0] for e in log][-5:] == ["close", "create", "copy_in", "gate", "sync"]
  n_gates = sum(1 for e in log if e[0] == "gate")
  assert n_gates == 7                                          # h cx u1 x t h s -- each exactly once
  # a full buffer is handed over on its own
  big = circuit.qc("big")
  big.reg(2, 0)
  for _ in range(big._gbuf.cap + 5):
    big.h(0)
  assert sum(1 for e in log if e[0] == "gate") == n_gates + big._gbuf.cap
