"""GPU: the three reference-facing faces over the C ABI.
  * circuit.qc (python operator surface, device-resident state)
  * libq C++ face (programs compiled against qcc_b200/libq/libq.h)
  * libxgates shim (the two callables the reference's circuit.py binds)
Everything is compared to golden vectors recorded from the reference or to the oracle."""
import json
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, load_golden, oracle, random_state, stream_of
from qcc_b200 import circuit, ops, workloads

pytestmark = pytest.mark.gpu
TOL = 1e-12


# ---------------------------------------------------------------------------------------
# circuit.qc
# ---------------------------------------------------------------------------------------
def test_qc_qft6_eager_matches_reference():
  qc = circuit.qc("qft6")
  r = qc.reg(6, 0b101101)
  qc.qft(r)
  z = load_golden("circ_qft6.npz")
  assert np.abs(np.asarray(qc.psi) - z["final_xgates"]).max() <= TOL
  assert qc.nbits == 6 and qc.psi.nbits == 6


def test_qc_acceleration_sequences():
  """circuit_test.py:69-107 on our surface, incl. negative control indices."""
  qc = circuit.qc()
  qc.bitstring(1, 0, 1, 0)
  for i in range(4):
    qc.x(i)
    qc.y(i)
    qc.z(i)
    qc.h(i)
    if i:
      qc.cu1(0, i, 1.1)
  assert np.abs(np.asarray(qc.psi) - load_golden("acceleration_1.npz")["final_spec"]).max() <= TOL
  qc = circuit.qc()
  qc.bitstring(1, 0, 1, 0, 1)
  for n in range(5):
    qc.h(n)
    for i in range(0, 5):
      qc.cu1(n - (i + 1), n, math.pi / float(2 ** (i + 1)))
    for i in range(0, 5):
      qc.cu1(n - (i + 1), n, -math.pi / float(2 ** (i + 1)))
    qc.h(n)
  assert np.abs(np.asarray(qc.psi) - load_golden("acceleration_2.npz")["final_spec"]).max() <= TOL


def test_qc_composites_and_inverse():
  z = load_golden("circ_composites9.npz")
  qc = circuit.qc("composites")
  qc.psi = z["psi0"]
  qc.global_reg = 9
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
  assert np.abs(np.asarray(qc.psi) - z["final_xgates"]).max() <= 1e-11


def test_qc_toffoli_swap_truth_tables():
  """circuit_test.py:19-58 style checks."""
  for bits in [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1)]:
    qc = circuit.qc()
    qc.bitstring(*bits)
    qc.toffoli(0, 1, 2)
    want = (bits[0], bits[1], bits[2] ^ (bits[0] & bits[1]))
    assert abs(qc.psi.prob(*want) - 1.0) < 1e-12
    qc.swap(0, 2)
    assert abs(qc.psi.prob(want[2], want[1], want[0]) - 1.0) < 1e-12
    qc.cswap(1, 0, 2)
    w2 = (want[2], want[1], want[0])
    w3 = (w2[2], w2[1], w2[0]) if w2[1] else w2
    assert abs(qc.psi.prob(*w3) - 1.0) < 1e-12


def test_qc_rotation_probabilities():
  """circuit_test.py:62-67."""
  qc = circuit.qc()
  qc.bitstring(0)
  qc.rx(0, 2 * np.arcsin(0.5))
  assert abs(qc.psi.prob(0) - 0.75) < 1e-12 and abs(qc.psi.prob(1) - 0.25) < 1e-12


def test_qc_state_builders_and_readouts():
  qc = circuit.qc()
  qc.reg(2, 0b10)
  qc.qubit(alpha=0.6)                      # dense factor: host kron, then upload
  qc.zeros(1)
  qc.ones(1)
  v = np.asarray(qc.psi)
  want = np.kron(np.kron(np.kron(np.array([0, 0, 1, 0]), np.array([0.6, 0.8])), [1, 0]), [0, 1])
  assert np.abs(v - want).max() < 1e-15
  assert qc.nbits == 5
  bits, p = qc.psi.maxprob()
  assert bits == [1, 0, 1, 0, 1] and abs(p - 0.64) < 1e-12
  assert abs(qc.psi.ampl(1, 0, 0, 0, 1) - 0.6) < 1e-15
  assert abs(qc.psi[0b10101] - 0.8) < 1e-15
  assert len(qc.psi) == 32
  # adding a register after gates ran expands the device state
  qc.h(0)
  qc.reg(1, 1)
  assert qc.nbits == 6 and abs(np.linalg.norm(np.asarray(qc.psi)) - 1.0) < 1e-12
  assert abs(qc.psi.prob_of_qubit(5) - 1.0) < 1e-12
  t = random_state(4, 1)
  qc2 = circuit.qc()
  r = qc2.state(t)
  assert len(r) == 4 and np.abs(np.asarray(qc2.psi) - t).max() == 0
  qc3 = circuit.qc()
  qc3.arange(3)
  assert np.array_equal(np.asarray(qc3.psi), np.arange(8).astype(np.complex128))


def test_qc_measure_bit():
  psi0 = random_state(6, 9)
  qc = circuit.qc()
  qc.psi = psi0
  idx = np.arange(64)
  p1 = float(np.sum(np.abs(psi0[(idx >> (5 - 2)) & 1 == 1]) ** 2))
  prob, _ = qc.measure_bit(2, 1, collapse=False)
  assert abs(prob - p1) < 1e-12
  assert abs(qc.pauli_expectation(2) - (1 - 2 * p1)) < 1e-12
  prob, _ = qc.measure_bit(2, 0, collapse=True)
  want = psi0.copy()
  want[(idx >> 3) & 1 == 1] = 0
  want /= np.linalg.norm(want)
  assert abs(prob - (1 - p1)) < 1e-12
  assert np.abs(np.asarray(qc.psi) - want).max() < 1e-12


def test_qc_run_subcircuit_inverse():
  """circuit_test.py:109-150 style: record, run, splice, invert."""
  main = circuit.qc("main")
  main.reg(5, 0b10110)
  sub = main.sub("s")
  sub.h(0)
  sub.cx(0, 1)
  sub.cu1(1, 2, 0.4)
  sub.ry(2, 1.2)
  main.qc(sub, offset=1)
  before = np.asarray(main.psi).copy()
  main.qc(sub.inverse(), offset=1)
  after = np.asarray(main.psi)
  want = np.zeros(32, dtype=np.complex128)
  want[0b10110] = 1
  assert np.abs(after - want).max() < 1e-12 and np.abs(before - want).max() > 0.1
  rec = circuit.qc("rec", eager=False)
  r = rec.reg(7, 0b1011001)
  rec.qft(r)
  rec.run()
  eager = circuit.qc("eager")
  r = eager.reg(7, 0b1011001)
  eager.qft(r)
  assert np.abs(np.asarray(rec.psi) - np.asarray(eager.psi)).max() < 1e-14


@pytest.mark.parametrize("nb", [3, 4])
def test_qc_grover_matches_reference_run(nb):
  """grover.py:124-168 rebuilt on our surface == the reference's own run (golden)."""
  z = load_golden(f"circ_grover_{nb}.npz")
  rd = load_golden(f"grover_{nb}_readout.npz")
  np.random.seed(0)
  qc, bits = workloads.grover_circuit(nb)
  assert np.abs(np.asarray(qc.psi) - z["final_xgates"]).max() <= 1e-10
  maxbits, maxprob = qc.psi.maxprob()
  assert maxbits[:nb] == bits
  assert maxbits == [int(b) for b in rd["maxbits"]] and abs(maxprob - float(rd["maxprob"])) < 1e-10


def test_qc_supremacy_style_stream_is_unitary_and_deterministic():
  s1 = workloads.supremacy(16, 12, seed=3)
  s2 = workloads.supremacy(16, 12, seed=3)
  assert len(s1) == len(s2) and all(a[:3] == b[:3] for a, b in zip(s1, s2))
  from qcc_b200 import _cabi
  want = np.zeros(1 << 16, dtype=np.complex128)
  want[0] = 1
  oracle.c_run(want, 16, s1)
  with _cabi.DeviceState(16, 0) as s:
    s.xg_apply_gates(_cabi.pack_xg_gates(s1))
    assert np.abs(s.copy_out() - want).max() <= TOL


# ---------------------------------------------------------------------------------------
# libq C++ face
# ---------------------------------------------------------------------------------------
def build_and_run(src, tmp_path):
  exe = tmp_path / "prog"
  lib = os.path.join(ROOT, "qcc_b200", "lib")
  subprocess.run(["g++", "-O1", "-I" + os.path.join(ROOT, "qcc_b200", "libq"), os.path.join(ROOT, "tests", src),
                  "-L" + lib, "-lqcc_libq", "-lqcc_b200", "-Wl,-rpath," + lib, "-o", str(exe)],
                 check=True, capture_output=True)
  return subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=120).stdout


def parse_print_qureg(text):
  out = {}
  for m in re.finditer(r"^\s*(-?\d+\.\d+) ([+-]\d+\.\d+)i\|(\d+)> \(([\d.e+-]+)\) \(\|([01 ]+)>\)", text, re.M):
    out[int(m.group(3))] = (complex(float(m.group(1)), float(m.group(2))), float(m.group(4)), m.group(5))
  return out


def test_libq_bell_program_matches_reference_output(tmp_path):
  """Same program as src/libq/libq_test.cc; the reference's stdout is the golden."""
  out = build_and_run("libq_progs/bell_u1.cc", tmp_path)
  ref = open(os.path.join(GOLDEN, "libq_test.out")).read()
  ours, theirs = parse_print_qureg(out), parse_print_qureg(ref)
  assert set(ours) == set(theirs) == {0, 3}
  for k in ours:
    assert abs(ours[k][0] - theirs[k][0]) < 2e-6 and abs(ours[k][1] - theirs[k][1]) < 1e-6
    assert ours[k][2] == theirs[k][2]            # bit-string rendering
  assert " # States: 2" in out and "States with non-zero probability:" in out


def test_libq_transpiled_qft6_matches_python_face(tmp_path):
  """configs[0]: circuit.qc -> dumpers.libq text (the reference's own output, committed as
  golden) -> g++ against our libq.h -> run -> same state as the python path, up to the
  bit reversal between the two faces (SURVEY.md trap 2)."""
  out = build_and_run("golden/qft6_libq.cc", tmp_path)
  got = parse_print_qureg(out)
  z = load_golden("circ_qft6.npz")
  perm = oracle.bitrev_perm(6)
  assert len(got) == 64
  for label, (amp, prob, _) in got.items():
    want = z["final_xgates"][perm[label]]
    assert abs(amp - want) < 2e-6 and abs(prob - abs(want) ** 2) < 1e-6
  assert "# of qubits        : 6" in out and "Maximum # of states: 64, theoretical: 128" in out


def test_libq_every_gate_name(tmp_path):
  out = build_and_run("libq_progs/all_gates.cc", tmp_path)
  amps = np.zeros(32, dtype=np.complex128)
  for m in re.finditer(r"^amp (\d+) (\S+) (\S+)$", out, re.M):
    amps[int(m.group(1))] = complex(float(m.group(2)), float(m.group(3)))
  g601 = np.array([[0.6, 0.8j], [0.8j, 0.6]])
  g601 = g601.astype(np.complex64).astype(np.complex128)   # libq_gate1 takes complex<float>
  prog = [("walsh", 5), ("x", 0), ("y", 1), ("z", 2), ("h", 3), ("t", 4), ("v", 0), ("yroot", 1), ("s", 2),
          ("cx", 0, 1), ("cz", 1, 2), ("ccx", 0, 1, 3), ("u1", 2, 0.3), ("cu1", 3, 4, math.pi / 16),
          ("cv", 4, 0), ("cv_adj", 2, 3), ("rx", 0, 0.7), ("ry", 1, -0.4), ("rz", 2, 1.1),
          ("crx", 0, 4, 0.2), ("cry", 1, 3, 0.9), ("crz", 2, 0, -0.6), ("sdag", 1), ("tdag", 2), ("vdag", 3),
          ("yrootdag", 4), ("ch", 0, 2), ("cs", 1, 4), ("ct", 3, 0), ("cy", 4, 1), ("cyroot", 2, 3),
          ("gate1", 4, g601)]
  want = oracle.libq_dense(5, 5, prog)
  assert np.abs(amps - want).max() < 1e-12
  m = re.search(r"size (\d+) width 5 norm2 (\S+)", out)
  assert int(m.group(1)) == int(np.sum(np.abs(want) ** 2 >= 1e-6 / 32))
  assert abs(float(m.group(2)) - np.sum(np.abs(want) ** 2)) < 1e-12   # gate1 matrix is not exactly unitary in float


# ---------------------------------------------------------------------------------------
# libxgates shim
# ---------------------------------------------------------------------------------------
def test_libxgates_shim_is_a_drop_in():
  sys.path.insert(0, os.path.join(ROOT, "qcc_b200", "shim"))
  try:
    import libxgates
  finally:
    sys.path.pop(0)
  assert sorted(n for n in dir(libxgates) if not n.startswith("_"))[:2] == ["apply1", "applyc"]
  z = load_golden("dense_n8.npz")
  for dtype, bw, key, tol in ((np.complex128, 128, "final_xgates", TOL), (np.complex64, 64, "final_xgates_f", 2e-5)):
    psi = np.ascontiguousarray(z["psi0"].astype(dtype))
    for kind, c, t, m in stream_of(z):
      g = m.reshape(4).astype(dtype)
      ret = libxgates.apply1(psi, g, 8, t, bw) if kind == 1 else libxgates.applyc(psi, g, 8, c, t, bw)
      assert ret is None
    assert np.abs(psi - z[key]).max() <= tol
  with pytest.raises(TypeError):
    libxgates.apply1(np.zeros(4, dtype=np.complex64), np.eye(2).reshape(4), 2, 0, 128)
  with pytest.raises(ValueError):
    libxgates.apply1(np.zeros(4, dtype=np.complex128), np.eye(2, dtype=np.complex128).reshape(4) * 1j, 2, 2, 128)


# ---------------------------------------------------------------------------------------
# the QFT adder behind src/libq/libq_arith_test.cc, end to end on both faces
# ---------------------------------------------------------------------------------------
def _arith_golden_label():
  text = open(os.path.join(GOLDEN, "libq_arith_test.out")).read()
  got = parse_print_qureg(text)
  assert len(got) == 1
  return next(iter(got))                      # 24581


def test_qft_adder_26_qubits_python_face():
  from qcc_b200 import helper
  n = 12
  qc, a, _ = workloads.qft_adder(n, 2, 3)
  assert qc.nbits == 26
  maxbits, p = qc.psi.maxprob()
  assert abs(p - 1.0) < 1e-9
  assert helper.bits2val(maxbits[0:n + 1][::-1]) == 5          # arith_quantum.py:34-39
  # same basis state as the reference's libq run of this circuit, up to the bit reversal between faces
  label = sum(bit << q for q, bit in enumerate(maxbits))       # python qubit q == libq bit q
  assert label == _arith_golden_label()
  labels, amps, total = qc.psi.nonzero(1e-9)
  assert total == 1 and abs(amps[0] - 1.0) < 1e-9


def test_qft_adder_transpiled_to_libq(tmp_path):
  """circuit.qc (non-eager) -> qc.libq() -> g++ against our libq.h -> run: prints the same single
  basis state the reference's own libq build prints for src/libq/libq_arith_test.cc."""
  qc, _, _ = workloads.qft_adder(12, 2, 3, eager=False)
  src = tmp_path / "adder.cc"
  src.write_text(qc.libq())
  exe = tmp_path / "adder"
  lib = os.path.join(ROOT, "qcc_b200", "lib")
  subprocess.run(["g++", "-O1", "-I" + os.path.join(ROOT, "qcc_b200", "libq"), str(src), "-L" + lib,
                  "-lqcc_libq", "-lqcc_b200", "-Wl,-rpath," + lib, "-o", str(exe)], check=True, capture_output=True)
  out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=300).stdout
  got = parse_print_qureg(out)
  assert list(got) == [_arith_golden_label()]
  amp, prob, bits = got[_arith_golden_label()]
  assert abs(amp - 1.0) < 1e-5 and abs(prob - 1.0) < 1e-5
  ref = parse_print_qureg(open(os.path.join(GOLDEN, "libq_arith_test.out")).read())
  assert bits == ref[_arith_golden_label()][2]
  assert "# of qubits        : 26" in out


# ---------------------------------------------------------------------------------------
# order finding (order_finding.py:152-204) end to end
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("fusion", [True, False])
def test_order_finding_golden_stream_on_device(fusion):
  """The 10 297-gate stream the reference records for N=15, a=4 (18 qubits): the device state has the
  basis states with p > 0.01, the sampled amplitudes and the norm of the reference's xgates run."""
  from qcc_b200 import _cabi
  z = load_golden("order_N15_a4.npz")
  n = int(z["nbits"])
  with _cabi.DeviceState(n, 0) as s:
    s.set_fusion(fusion)
    s.xg_apply_gates(_cabi.pack_xg_gates(stream_of(z)))
    assert abs(s.norm2() - float(z["norm2"])) < 1e-10
    labels, amps, total = s.list_above(0.01)
    assert total == len(z["labels"]) and np.array_equal(np.sort(labels), z["labels"])
    order = np.argsort(labels)
    assert np.abs(np.abs(np.asarray(amps)[order]) ** 2 - z["probs"]).max() < 1e-10
    for idx, want in zip(z["sample_idx"], z["sample_amp"]):
      assert abs(s.amplitude(int(idx)) - want) < 1e-10
    cnt = s.counters()
    assert cnt["gates_applied"] == 10297
    if fusion:
      assert cnt["passes"] * 50 < cnt["gates_applied"]


@pytest.mark.parametrize("number,a,order", [(15, 4, 2), (21, 11, 6)])
def test_order_finding_python_face(number, a, order):
  """order_finding.py main on the device-resident surface (18 and 22 qubits -- the second is the
  circuit behind src/libq/libq_order22_test.cc): the phase register peaks at multiples of 1/order and
  the continued-fraction readout of order_finding.py:185-202 recovers the order."""
  qc, aux, up, down = workloads.order_finding(number, a)
  nbits = number.bit_length()
  assert qc.nbits == 4 * nbits + 2
  assert abs(qc.psi.norm2() - 1.0) < 1e-9
  found = workloads.order_readout(qc, number, a)
  assert found and sum(p for _, _, p, _, _ in found) > 0.5
  rs = set()
  for x, phase, p, r, guesses in found:
    # every peak sits (to register resolution) on a multiple of 1/order
    k = round(phase * order)
    assert abs(phase - k / order) < 2.0 ** (-2 * nbits) * 1.01 + 1e-12, (x, phase)
    rs.add(r)
  assert all(order % r == 0 for r in rs) and math.lcm(*rs) == order
  assert pow(a, order, number) == 1
  qc.close()


# ---------------------------------------------------------------------------------------
# The UNMODIFIED reference over the libxgates shim: its own test file
# and its supremacy.py, with PYTHONPATH = <reference tree>:<qcc_b200/shim>, so that circuit.py:36-41
# binds OUR apply1 / applyc.  The tree is staged by __graft_entry__.build() under the git-ignored
# baseline/_ref/qcc (it travels to the GPU box); /root/reference is used where it exists.
# ---------------------------------------------------------------------------------------
def _reference_tree():
  for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref", "qcc")):
    if os.path.exists(os.path.join(cand, "src", "lib", "circuit_test.py")):
      return cand
  return None


def _run_reference(script, args, tmp_path):
  ref = _reference_tree()
  log = str(tmp_path / "shim_calls.json")
  env = dict(os.environ, PYTHONPATH=ref + os.pathsep + os.path.join(ROOT, "qcc_b200", "shim"),
             QCC_B200_SHIM_LOG=log)
  r = subprocess.run([sys.executable, os.path.join(ref, "src", script)] + args, cwd=ref, env=env,
                     capture_output=True, text=True, timeout=900)
  calls = json.load(open(log)) if os.path.exists(log) else None
  return r, calls


@pytest.mark.parametrize("width", [64, 128])
def test_reference_circuit_test_passes_over_the_shim(width, tmp_path):
  """src/lib/circuit_test.py (21 tests; :69-107 is the one that pins xgates against the Python definition,
  negative controls included) run unchanged, every gate executed by the B200 through libxgates.py."""
  if _reference_tree() is None:
    pytest.skip("no reference tree staged (baseline/_ref/qcc)")
  r, calls = _run_reference(os.path.join("lib", "circuit_test.py"), [f"--tensor_width={width}"], tmp_path)
  out = r.stdout + r.stderr
  assert r.returncode == 0, out[-3000:]
  assert "Could not find 'libxgates.so'" not in out      # the reference did not fall back to Python
  m = re.search(r"Ran (\d+) tests", out)
  assert m and int(m.group(1)) >= 20 and "OK" in out, out[-2000:]
  assert calls and calls["apply1"] > 100 and calls["applyc"] > 100, calls


@pytest.mark.parametrize("width", [64, 128])
def test_reference_supremacy_runs_over_the_shim(width, tmp_path):
  """src/supremacy.py --nbits 20 (its default size) unchanged: build_circuit + sim_circuit, one shim call per gate."""
  if _reference_tree() is None:
    pytest.skip("no reference tree staged (baseline/_ref/qcc)")
  r, calls = _run_reference("supremacy.py", ["--nbits=20", "--depth=12", f"--tensor_width={width}"], tmp_path)
  out = r.stdout + r.stderr
  assert r.returncode == 0, out[-3000:]
  assert "Could not find 'libxgates.so'" not in out
  assert "Estimated sim for FULL experiment" in out
  assert calls and calls["apply1"] + calls["applyc"] > 100, calls
