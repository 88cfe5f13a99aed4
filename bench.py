#!/usr/bin/env python
"""bench.py -- gate-applies/sec on the reference's headline workloads, on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload qft30|larose28|hsweep30]
                    [--impl ours|reference]

Metric (BASELINE.json): gate-applies/sec at 30 qubits + achieved HBM GB/s.  A "step" is one
full pass of the workload's gate stream over the resident 2^n complex128 state:
  qft30     circuit.qc.qft on 30 qubits: 30 h + 435 cu1 = 465 gates          (configs[2], default)
  larose28  larose_benchmark.py at 28 qubits, depth 28: 2324 gates            (configs[1])
  hsweep30  one h on each of 30 qubits, gate-by-gate (no fusion): the "30-qubit single-qubit
            gate application" roofline target of BASELINE.json
The state (16 GiB at 30 qubits) is >> the 126 MB L2, so no explicit L2 flush is needed.

Timing: CUDA events on the engine's own stream (qb_timer_*), synchronised on both sides,
max over ranks.  `value` counts inputs resident in HBM; `e2e` goes through the host-buffer
C-ABI path (pinned host state -> qb_copy_in -> gates -> qb_copy_out) every step.
`roofline` is the dominant kernel's algorithmic bytes / its summed CUDA-event time, measured
in the same timed region (qb_profile_*), against MEASURED_PEAKS.json's hbm_gbs.
`cpu_baseline` times the reference's own xgates build (oracle/_ref/libxgates.so, 1 thread --
the reference has no threading) on a bounded sample of the same gate stream; `cpu_baseline.libq`
adds the reference's other implementation, stock libq, on a dense 24-qubit register (it stops at 28).

N > 1 (torchrun): the SAME workload (same qubit count, same gate stream) with its state sharded
over the N GPUs by the top log2(N) index bits -- strong scaling, so metric and config do not
change with N.  Gates on sharded qubits cost a pairwise half-shard exchange over NVLink
(ncclSend/ncclRecv) and a bit remap; diagonal gates and controls on sharded qubits cost nothing
extra.  `value` is gates / max-over-ranks time; exchange counts, bytes and GB/s are reported.
--weak instead grows the state to n + log2(N) qubits (16 GiB per GPU at every N).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# One-off workloads of the multi-GPU configs (not part of the default line): run the whole
# algorithm once through the python circuit.qc surface on a sharded state.
#   grover:     grover.py:124-168 with --qubits/2 search bits (configs[3]: --qubits 32 --gpus 4)
#   supremacy:  supremacy.py-style random circuit, depth 20 (configs[4]: --qubits 34 --gpus 8)
ALGOS = ("grover", "supremacy")

WORKLOADS = {
    "qft30": dict(n=30, desc="30-qubit QFT (circuit.qc.qft): 30 h + 435 cu1", fusion=True),
    "larose28": dict(n=28, desc="28-qubit larose_benchmark depth 28: 784 h + 784 v + 756 cx", fusion=True),
    "hsweep30": dict(n=30, desc="h on each of 30 qubits, one kernel per gate (no fusion)", fusion=False),
}


def build_stream(name, n):
  from qcc_b200 import workloads
  if name.startswith("supremacy"):
    return workloads.supremacy(n, 20, seed=0)
  if name.startswith("qft"):
    return workloads.qft(n)
  if name.startswith("larose"):
    return workloads.larose(n, n)
  if name.startswith("hsweep"):
    return workloads.hsweep(n)
  raise ValueError(name)


def hbm_peak():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    try:
      return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # pylint: disable=broad-except
      pass
  return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
  """SM clock, power and throttle reasons while the timed region runs: NVML every 10 ms (the timed
  region of the default run is a quarter of a second), nvidia-smi every 200 ms when NVML is missing."""
  Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")
  NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
  NVML_BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

  def __init__(self, index, period=0.01):
    super().__init__(daemon=True)
    self.index = index
    self.period = period
    self.samples = []     # (sm_mhz, power_w, set of reasons)
    self.sm_max = None
    self.stop_flag = False
    self.source = "nvidia-smi"
    self.nvml = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
      self.nvml = pynvml
      self.source = "nvml"
    except Exception:  # pylint: disable=broad-except
      self.nvml = None

  def _sample_nvml(self):
    nv = self.nvml
    sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
    try:
      pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
    except Exception:  # pylint: disable=broad-except
      pw = None
    try:
      bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
    except Exception:  # pylint: disable=broad-except
      bits = 0
    self.samples.append((sm, pw, {nm for nm, b in self.NVML_BITS.items() if bits & b}))

  def _sample_smi(self):
    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                          "--format=csv,noheader,nounits"], capture_output=True, text=True,
                         timeout=5).stdout.strip()
    if not out:
      return
    f = [x.strip() for x in out.split(",")]
    num = lambda x: float(x) if x.replace(".", "").isdigit() else None
    if self.sm_max is None:
      self.sm_max = num(f[1])
    if num(f[0]) is not None:
      self.samples.append((num(f[0]), num(f[2]), {nm for k, nm in enumerate(self.NAMES) if f[3 + k] == "Active"}))

  def run(self):
    while not self.stop_flag:
      try:
        if self.nvml is not None:
          self._sample_nvml()
        else:
          self._sample_smi()
      except Exception:  # pylint: disable=broad-except
        pass
      time.sleep(self.period if self.nvml is not None else 0.2)

  def summary(self):
    self.stop_flag = True
    self.join(timeout=6)
    if not self.samples:
      return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no clock samples (nvml / nvidia-smi unavailable)"]}
    sm = sorted(s[0] for s in self.samples)
    pw = [s[1] for s in self.samples if s[1] is not None]
    reasons = [nm for nm in self.NAMES if any(nm in s[2] for s in self.samples)]
    return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max,
            "power_w_max": max(pw) if pw else None, "samples": len(self.samples), "source": self.source,
            "reasons": reasons}


def cpu_reference_sample(n, stream, budget_s=20.0, max_gates=8):
  """Time the REFERENCE xgates build on the first gates of `stream` at full size n.
  Returns (gates_per_sec, description)."""
  from oracle import oracle
  if not oracle.have_ref("libxgates.so"):
    return None, "oracle/_ref/libxgates.so missing"
  xg = oracle.RefXgates()
  scale = 1.0
  nn = n
  while True:
    try:
      psi = np.zeros(1 << nn, dtype=np.complex128)
      break
    except MemoryError:
      nn -= 2
      scale /= 4.0
  psi[1] = 1.0
  psi[:] += 1.0 / (1 << (nn // 2 + 1))  # dense, so every pair does real arithmetic; also first-touches the pages
  t_total, done = 0.0, 0
  picks = [stream[(k * len(stream)) // max_gates] for k in range(max_gates)]   # spread over the whole stream
  for kind, c, t, m in picks:
    if nn != n:  # remap qubit numbers into the smaller register
      t = min(t, nn - 1)
      c = min(c, nn - 1)
      if kind == 2 and c == t:
        c = (t + 1) % nn
    t0 = time.perf_counter()
    if kind == 1:
      xg.apply1(psi, m, nn, t)
    else:
      xg.applyc(psi, m, nn, c, t)
    t_total += time.perf_counter() - t0
    done += 1
    if t_total > budget_s:
      break
  gps = done / t_total * scale
  desc = (f"{done} gates spread evenly over the {len(stream)}-gate stream with reference xgates (bit_width=128) on a "
          f"dense {nn}-qubit numpy state, {t_total:.1f} s of CPU work, 1 thread")
  if nn != n:
    desc += f"; host RAM could not hold {n} qubits: scaled by 4^-{(n - nn) // 2} (time ~ 2^n)"
  return gps, desc


def cpu_reference_libq_sample(width=24, ngates=26):
  """The reference's OTHER implementation of the path, stock libq (sparse, complex64; SURVEY 8d asks for
  it beside xgates): a dense `width`-qubit register (walsh first, since libq's cost follows the number of
  non-zero states) and the first gates of that width's QFT -- h is libq_gate1 (apply.cc:78-176, a hash
  rebuild per gate), cu1 a linear scan (gates.cc:85-94).  Timed as (walsh + gates) - (walsh), 1 thread."""
  import math
  from oracle import oracle
  if not oracle.have_ref("libq_ref.so"):
    return None
  lq = oracle.RefLibq(double=False)
  ops = []
  for i in reversed(range(width)):          # circuit.py:320-326 in libq qubit numbers
    ops.append(("h", i))
    for j in reversed(range(i)):
      ops.append(("cu1", i, j, math.pi / 2 ** (i - j)))
  ops = ops[:ngates]
  t0 = time.perf_counter()
  lq.run(width, 0, [("walsh", width)])
  t1 = time.perf_counter()
  lq.run(width, 0, [("walsh", width)] + ops)
  t2 = time.perf_counter()
  dt = max((t2 - t1) - (t1 - t0), 1e-9)
  nh = sum(1 for o in ops if o[0] == "h")
  return {"value": len(ops) / dt, "unit": "gates/s", "cores": 1, "kind": "reference", "qubits": width,
          "sample": (f"stock libq (src/libq, float, -O3 -ffast-math) on a dense {width}-qubit register: first "
                     f"{len(ops)} gates of QFT-{width} ({nh} h + {len(ops) - nh} cu1) in {dt:.2f} s after a walsh "
                     f"of {t1 - t0:.2f} s; libq stops at 28 qubits and its time grows with 2^n")}


def cpu_sweep(args):
  """BASELINE.md section 4, steps 2-3, on this box's host cores (1 thread: the reference has no threading):
  reference xgates (bit_width=128) per gate at n = 26 / 28 / 30 -- h on EVERY target position, cx with a near and
  a far control -- and stock libq on the QFT IR at n <= 28 after a walsh.  One JSON line."""
  from oracle import oracle
  out = {"impl": "reference", "kind": "cpu_sweep", "host_cores": os.cpu_count(), "threads": 1, "xgates": [], "libq": []}
  if not oracle.have_ref("libxgates.so"):
    print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libxgates.so not built"}))
    return
  xg = oracle.RefXgates()
  H = oracle.GATES["h"]
  X = oracle.GATES["x"]
  for n in args.sweep_qubits:
    try:
      psi = np.full(1 << n, 1.0 / np.sqrt(float(1 << n)), dtype=np.complex128)
    except MemoryError:
      out["xgates"].append({"qubits": n, "error": "host RAM"})
      continue
    xg.apply1(psi, H, n, 0)    # first touch
    per_target = []
    for t in range(n):
      t0 = time.perf_counter()
      xg.apply1(psi, H, n, t)
      per_target.append(time.perf_counter() - t0)
    cx = {}
    for name, c, t in (("near control (ctl 1, tgt 0)", 1, 0), ("far control (ctl n-1, tgt 0)", n - 1, 0),
                       ("far target (ctl 0, tgt n-1)", 0, n - 1)):
      t0 = time.perf_counter()
      xg.applyc(psi, X, n, c, t)
      cx[name] = time.perf_counter() - t0
    byts = 32.0 * (1 << n)
    out["xgates"].append({
        "qubits": n, "h_seconds_per_python_target": per_target,
        "h_gates_per_s": n / sum(per_target), "h_gbs_algorithmic": byts * n / sum(per_target) / 1e9,
        "h_slowest_s": max(per_target), "h_fastest_s": min(per_target),
        "cx_seconds": cx, "cx_gbs_algorithmic": {k: byts / 2 / v / 1e9 for k, v in cx.items()}})
    del psi
  for width in args.sweep_libq:
    r = cpu_reference_libq_sample(width=width, ngates=2 * width)
    if r:
      out["libq"].append(r)
  # n = 32 / 34: no reference runs (xgates is int-indexed, libq stops at 28); the reference's own rule
  # (supremacy.py:283-297): time ~ gates x bytes
  base = next((x for x in out["xgates"] if x.get("qubits") == max(args.sweep_qubits) and "h_gates_per_s" in x), None)
  if base:
    out["extrapolated"] = {f"{m} qubits": {"h_gates_per_s": base["h_gates_per_s"] / 2 ** (m - base["qubits"]),
                                          "rule": "time proportional to gates x bytes (supremacy.py:283-297), from "
                                                  f"the {base['qubits']}-qubit measurement; NOT measured"}
                           for m in (32, 34)}
  print(json.dumps(out))


def run_reference_arm(args, wl, stream):
  """--impl reference: the reference's own CPU implementation of the path, bounded sample per step."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  n = wl["n"]
  total = args.steps + args.warmup
  from oracle import oracle
  if not oracle.have_ref("libxgates.so"):
    print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libxgates.so not built"}))
    return
  xg = oracle.RefXgates()
  psi = np.zeros(1 << n, dtype=np.complex128)
  psi[:] = 1.0 / np.sqrt(float(1 << n))
  # one probe gate decides how many gates a step can afford (whole run <= ~150 s)
  t0 = time.perf_counter()
  xg.apply1(psi, stream[0][3], n, stream[0][2])
  probe = time.perf_counter() - t0
  per_step = max(1, min(len(stream), int(150.0 / max(probe, 1e-3) / max(total, 1))))
  # The sample is STRIDED over the whole stream (stride coprime with its length), so that over the run the
  # sampled gates have the stream's own mix of gate kinds and target positions -- xgates' cost per gate does not
  # depend on the state, so the sampled rate is an unbiased estimate of the whole-stream rate.
  stride = max(1, len(stream) // max(1, per_step * total))
  while stride > 1 and np.gcd(stride, len(stream)) != 1:
    stride -= 1
  pos = 0

  def step():
    nonlocal pos
    for _ in range(per_step):
      kind, c, t, m = stream[pos % len(stream)]
      pos += stride
      if kind == 1:
        xg.apply1(psi, m, n, t)
      else:
        xg.applyc(psi, m, n, c, t)

  for _ in range(args.warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    step()
  dt = time.perf_counter() - t0
  value = per_step * args.steps / dt
  sample = (f"{per_step} gates of the {args.workload} stream per step, taken every {stride}-th gate over the whole "
            f"{len(stream)}-gate stream ({per_step * total} gates in the run, same kind / target mix as the stream), on "
            f"a dense {n}-qubit complex128 numpy state, reference xgates (src/lib/xgates.cc built -O3 -ffast-math), "
            f"1 thread; a full step of {len(stream)} gates would take ~{len(stream) / max(value, 1e-9) / 60:.0f} min here")
  print(json.dumps({
      "impl": "reference", "metric": "gate-applies/sec", "value": value, "unit": "gates/s",
      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
      "config": {"workload": args.workload, "desc": wl["desc"], "qubits": n, "gates_per_step": per_step},
      "cpu_baseline": {"value": value, "unit": "gates/s", "cores": 1, "kind": "reference", "sample": sample},
      "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "host_cores": os.cpu_count(),
  }))


def _roofline_of(prof, shard_len, peak, peak_src):
  """roofline block of the dominant compute kernel class of a profile (qb_profile_read)."""
  names = {"apply1": "k_apply_u", "phase": "k_apply_phase", "fused": "k_fused_pass", "fused_push": "k_fused_pass (push store)"}
  cand = [k for k in names if prof[k]["launches"]]
  if not cand:
    return None
  dom = max(cand, key=lambda k: prof[k]["ms"])
  d = prof[dom]
  ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
  return {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
          "traffic": None, "peak_source": peak_src, "launches": d["launches"], "avg_launch_ms": d["ms"] / d["launches"],
          "algorithmic_bytes_per_launch": d["bytes"] / d["launches"],
          "note": "per GPU: algorithmic bytes (32 B x shard amplitudes per sweep) / summed CUDA-event time of that "
                  "kernel class on rank 0"}


def run_algorithm(args):
  """Whole-algorithm runs of the multi-GPU configs (configs[3] grover-32 x 4, configs[4] supremacy-34 x 8):
  ONE execution through the python circuit.qc surface (grover) / the packed stream (supremacy), timed by
  wall clock (what the user waits for: gate dispatch, lowering, planning, kernels, readout) and by the
  engine's CUDA events per kernel class (device time, roofline of the dominant kernel)."""
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  dist = None
  from qcc_b200 import _cabi, circuit, workloads

  def new_comm():
    if world == 1:
      return None
    ids = [_cabi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return ids[0]

  if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  n = args.qubits or (32 if args.workload == "grover" else 34)
  peak, peak_src = hbm_peak()
  factory = lambda name: circuit.qc(name, device=local_rank, rank=rank, nranks=world, comm_id=new_comm())
  extra = {}
  sampler = ClockSampler(local_rank, period=0.05)
  sampler.start()
  t0 = time.perf_counter()
  if args.workload == "grover":
    nb = n // 2
    np.random.seed(0)
    qc = factory("Grover")
    prof_on = []

    def fac(_name):
      return qc

    # the state exists after the first register + gate; profiling is switched on by a hook on materialisation
    orig = qc._new_device_state

    def hooked(*a, **k):
      nonlocal t0
      dev = orig(*a, **k)
      dev.profile_enable(True)
      prof_on.append(dev)
      t0 = time.perf_counter()   # the state exists (CUDA context, allocation, |0..0>): the run starts here
      return dev

    qc._new_device_state = hooked
    qc, bits = workloads.grover_circuit(nb, qc_factory=fac, iterations=args.iterations or None)
    maxbits, maxprob = qc.psi.maxprob()
    check = {"marked": bits, "found": maxbits[:nb], "ok": maxbits[:nb] == bits, "maxprob": maxprob}
    desc = (f"grover.py:124-168 (circuit version), {nb} search bits + 1 ancilla + {nb - 1} multi_control helpers = "
            f"{n} qubits, np.random.seed(0)" + (f", {args.iterations} iterations" if args.iterations else ""))
  else:
    stream = workloads.supremacy(n, args.depth, seed=0)
    packed = _cabi.pack_xg_gates(stream)
    qc = factory("supremacy")
    qc.reg(n, 0)
    qc.dev.profile_enable(True)
    t0 = time.perf_counter()
    qc.dev.timer_start()
    qc.dev.xg_apply_gates(packed)
    extra["device_ms_circuit"] = qc.dev.timer_stop()
    wall_circuit = time.perf_counter() - t0
    norm = qc.psi.norm2()
    amp0 = qc.psi[0]
    # full-size property: the circuit followed by its inverse (reversed stream, adjoint matrices) is the
    # identity -- every pass, exchange event and relabel of the forward run has to be undone exactly
    inv = [(k, c, t, np.conj(np.asarray(m).reshape(2, 2)).T) for k, c, t, m in reversed(stream)]
    qc.dev.xg_apply_gates(_cabi.pack_xg_gates(inv))
    back = qc.psi[0]
    norm_back = qc.psi.norm2()
    check = {"norm2": norm, "amp0_after_circuit": [amp0.real, amp0.imag],
             "amp0_after_circuit_then_inverse": [back.real, back.imag], "norm2_after_inverse": norm_back,
             "ok": abs(norm - 1.0) < 1e-9 and abs(back - 1.0) < 1e-9 and abs(norm_back - 1.0) < 1e-9}
    extra["wall_s_circuit"] = wall_circuit
    desc = (f"supremacy.py build_circuit + sim_circuit gate stream (supremacy.py:123-158, 208-240), {n} qubits, "
            f"depth {args.depth}, random.seed(0): {len(stream)} gates; then the inverse stream as a check")
  qc.sync()
  wall = time.perf_counter() - t0
  clocks = sampler.summary()
  c = qc.dev.counters()
  prof = qc.dev.profile_read(reset=True)
  mode = qc.dev.exchange_mode()
  nl = n - int(np.log2(world))
  device_ms = sum(v["ms"] for v in prof.values())
  # parity of the sharded engine against a single GPU on the same seeded generator at a size one GPU holds
  parity = None
  if args.workload == "supremacy" and args.check_qubits and world > 1:
    m = min(args.check_qubits, n)
    st = workloads.supremacy(m, args.depth, seed=0)
    pk = _cabi.pack_xg_gates(st)
    qs = circuit.qc("check", device=local_rank, rank=rank, nranks=world, comm_id=new_comm())
    qs.reg(m, 0)
    qs.dev.xg_apply_gates(pk)
    rng = np.random.default_rng(5)
    idx = [int(x) for x in rng.integers(0, 1 << m, size=48)] + [0, (1 << m) - 1]
    got = np.array([qs.psi[i] for i in idx])
    ex_events = qs.dev.counters()["exchanges"]
    qs.close()
    if rank == 0:
      with _cabi.DeviceState(m, 0, local_rank) as one:
        one.xg_apply_gates(pk)
        want = np.array([one.amplitude(i) for i in idx])
      err = float(np.abs(got - want).max())
      parity = {"qubits": m, "sampled_amplitudes": len(idx), "max_abs_err_vs_single_gpu": err, "ok": err <= 1e-10,
                "exchange_events": ex_events, "gates": len(st)}
  if dist is not None:
    dist.barrier()
  if rank == 0:
    gates = c["gates_applied"]
    print(json.dumps({
        "metric": "gate-applies/sec", "value": gates / wall, "unit": "gates/s", "n_gpus": world,
        "steps": 1, "warmup": 0, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": desc, "qubits": n, "shard_qubits": nl,
                   "parallelism": f"state sharded over {world} GPUs, exchange mode {mode}" if world > 1 else "1 GPU",
                   "timing": "value = gates / wall clock of the one full run, from the moment the state exists to the "
                             "last readout: python gate dispatch, lowering, planning, kernels, readouts; device_ms = "
                             "summed CUDA-event time of every kernel class on rank 0"},
        "gates": gates, "passes": c["passes"], "gpu_launches": c["kernel_launches"],
        "wall_s": wall, "device_ms": device_ms, "host_overhead_frac": 1.0 - device_ms * 1e-3 / wall if wall else None,
        "kernel_ms": {k: v["ms"] for k, v in prof.items() if v["launches"]},
        "kernel_launches": {k: v["launches"] for k, v in prof.items() if v["launches"]},
        "roofline": _roofline_of(prof, 1 << nl, peak, peak_src),
        "exchange": {"mode": mode, "events": c["exchanges"], "bytes_sent_per_rank": c["bytes_exchanged"],
                     "fused_push_passes": prof["fused_push"]["launches"],
                     "nvlink_gbs_per_direction_rank0":
                         c["bytes_exchanged"] / ((prof["fused_push"]["ms"] + prof["exchange"]["ms"]) * 1e-3) / 1e9
                         if (prof["fused_push"]["ms"] + prof["exchange"]["ms"]) > 0 else None} if world > 1 else None,
        "check": check, "parity_vs_single_gpu": parity, "clocks": clocks, **extra}))
  qc.close()
  if dist is not None:
    dist.destroy_process_group()


def run_matrix(args):
  """--matrix qft:28,qft:30,supremacy:34,...: the north-star table (QFT and the supremacy.py random circuit at
  28-34 qubits on N GPUs) in ONE process per rank -- one line of JSON per entry, same timing rules as the default
  line (CUDA events on the engine's stream, max over ranks, K steps queued and flushed inside the timed region)."""
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  dist = None
  if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  from qcc_b200 import _cabi
  peak, peak_src = hbm_peak()
  for entry in args.matrix.split(","):
    name, n = entry.split(":")
    n = int(n)
    line = {"workload": name, "qubits": n, "n_gpus": world}
    s = None
    try:
      stream = build_stream(name, n)
      packed = _cabi.pack_xg_gates(stream)
      comm_id = None
      if world > 1:
        ids = [_cabi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm_id = ids[0]
      s = _cabi.DeviceState(n, 0, local_rank, rank=rank, nranks=world, comm_id=comm_id)
      s.set_tile_bits(args.tile_bits)
      s.fill_random(1234)
      for _ in range(args.warmup):
        s.xg_apply_gates(packed)
      s.sync()
      if dist is not None:
        dist.barrier()
      c0 = s.counters()
      s.profile_enable(True)
      s.profile_read(reset=True)
      s.timer_start()
      for _ in range(args.steps):
        s.xg_apply_gates(packed)
      ms = s.timer_stop()
      prof = s.profile_read(reset=True)
      s.profile_enable(False)
      c1 = s.counters()
      if dist is not None:
        import torch
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
      norm = s.norm2()
      sent = c1["bytes_exchanged"] - c0["bytes_exchanged"]
      carrier = prof["fused_push"]["ms"] + prof["exchange"]["ms"]
      nl = n - int(np.log2(world))
      line.update({
          "gates_per_step": len(stream), "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
          "value": len(stream) * args.steps / (ms * 1e-3), "unit": "gates/s", "shard_qubits": nl,
          "passes_per_step": (c1["passes"] - c0["passes"]) / args.steps,
          "achieved_gbs_swept_per_gpu": (c1["bytes_swept"] - c0["bytes_swept"]) / (ms * 1e-3) / 1e9,
          "frac_of_hbm_roofline_per_gpu": (c1["bytes_swept"] - c0["bytes_swept"]) / (ms * 1e-3) / 1e9 / peak,
          "roofline": _roofline_of(prof, 1 << nl, peak, peak_src),
          "kernel_ms": {k: v["ms"] for k, v in prof.items() if v["launches"]},
          "exchange": None if world == 1 else {
              "mode": s.exchange_mode(), "events_per_step": (c1["exchanges"] - c0["exchanges"]) / args.steps,
              "bytes_sent_per_rank_per_step": sent / args.steps,
              "nvlink_gbs_per_direction_rank0": sent / (carrier * 1e-3) / 1e9 if carrier else None},
          "norm2_after": norm})
    except Exception as ex:  # pylint: disable=broad-except
      line["error"] = str(ex)[:300]
    if s is not None:
      s.close()
    if rank == 0:
      print(json.dumps(line), flush=True)
  if dist is not None:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--workload", default="qft30", choices=sorted(WORKLOADS) + list(ALGOS))
  ap.add_argument("--depth", type=int, default=20)
  ap.add_argument("--iterations", type=int, default=0, help="grover: override the iteration count (debug)")
  ap.add_argument("--check-qubits", type=int, default=28,
                  help="supremacy on N > 1 GPUs: also run the generator at this size sharded AND on one GPU and "
                       "compare sampled amplitudes (0 = skip)")
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--qubits", type=int, default=0, help="override the workload's qubit count (debug)")
  ap.add_argument("--tile-bits", type=int, default=12)
  ap.add_argument("--e2e-steps", type=int, default=2)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-secondary", action="store_true")
  ap.add_argument("--matrix", default="", help="comma list of workload:qubits entries, e.g. qft:28,supremacy:30 "
                                               "(one JSON line each; see run_matrix)")
  ap.add_argument("--cpu-sweep", action="store_true",
                  help="reference CPU baselines of BASELINE.md section 4 (xgates per target at 26/28/30 qubits, libq "
                       "on the QFT IR): host only, one JSON line")
  ap.add_argument("--sweep-qubits", type=int, nargs="*", default=[26, 28, 30])
  ap.add_argument("--sweep-libq", type=int, nargs="*", default=[22, 24, 26])
  ap.add_argument("--weak", action="store_true", help="N > 1: grow the state to n + log2(N) qubits")
  ap.add_argument("--flush-per-step", action="store_true",
                  help="flush the gate queue after every step instead of once at the end of the timed region "
                       "(sharded states: the lowering then cannot place exchange events across step boundaries)")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
  if args.cpu_sweep:
    return cpu_sweep(args)
  if args.matrix:
    return run_matrix(args)
  if args.workload in ALGOS:
    return run_algorithm(args)
  wl = dict(WORKLOADS[args.workload])
  if args.qubits:
    wl["n"] = args.qubits
  n = wl["n"]
  stream = build_stream(args.workload, n)

  if args.impl == "reference":
    run_reference_arm(args, wl, stream)
    return

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  dist = None
  if world > 1:
    import torch
    import torch.distributed as dist  # plumbing only: barrier + max-reduce of the timings
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

  from qcc_b200 import _cabi
  comm_id = None
  n_shard = n
  if world > 1:
    if world & (world - 1):
      raise SystemExit("--gpus must be a power of two")
    ids = [_cabi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm_id = ids[0]
    if args.weak:
      n = n_shard + int(np.log2(world))     # one bigger state, same shard per GPU
      wl["n"] = n
      stream = build_stream(args.workload, n)
    else:
      n_shard = n - int(np.log2(world))
  packed = _cabi.pack_xg_gates(stream)
  ngates = len(stream)
  s = _cabi.DeviceState(n, 0, local_rank, rank=rank, nranks=world, comm_id=comm_id)
  s.set_fusion(wl["fusion"])
  if wl["fusion"]:
    s.set_tile_bits(args.tile_bits)
  s.fill_random(1234)

  def barrier():
    s.sync()
    if dist is not None:
      dist.barrier()

  for _ in range(args.warmup):
    s.xg_apply_gates(packed)
    if args.flush_per_step:
      s.flush()
  barrier()
  sampler = ClockSampler(local_rank)
  sampler.start()
  c0 = s.counters()
  s.profile_enable(True)
  s.profile_read(reset=True)
  t_wall0 = time.perf_counter()
  s.timer_start()
  for _ in range(args.steps):
    s.xg_apply_gates(packed)     # queued (the C ABI's deferred-gate protocol, gates_jit.cc:53-132) ...
    if args.flush_per_step:
      s.flush()
  ms = s.timer_stop()            # ... and flushed here at the latest: plans, launches, waits for the device
  barrier()
  wall = time.perf_counter() - t_wall0
  prof = s.profile_read(reset=True)
  s.profile_enable(False)
  c1 = s.counters()
  clocks = sampler.summary()
  if dist is not None:
    import torch
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
  norm = s.norm2()

  # ---- roofline of the dominant kernel class --------------------------------------------
  peak, peak_src = hbm_peak()
  dom = max((k for k in prof if k not in ("exchange", "fused_push")), key=lambda k: prof[k]["ms"])
  d = prof[dom]
  roof = None
  if d["launches"]:
    ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": {"apply1": "k_apply_u", "phase": "k_apply_phase", "fused": "k_fused_pass",
                                      "aux": "aux"}[dom],
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "peak_source": peak_src, "launches": d["launches"],
            "avg_launch_ms": d["ms"] / d["launches"],
            "algorithmic_bytes_per_launch": d["bytes"] / d["launches"],
            "share_of_step": d["ms"] / ms if ms else None,
            "note": "achieved = algorithmic bytes (32 B x 2^n per full sweep; SURVEY 8d) / summed "
                    "CUDA-event time of that kernel inside the timed region; traffic (ncu dram bytes) "
                    "is recorded under profiles/"}
  by_gate_alg = (c1["bytes_algorithmic"] - c0["bytes_algorithmic"]) / (ms * 1e-3) / 1e9
  # ncu dram bytes per launch of the same kernels: a CONSTANT from profiles/r02_traffic.json (one --set full
  # capture, provenance recorded there and copied into the line), not something this run measured
  try:
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    if roof and n == traffic.get("qubits") and world == 1:
      roof["traffic"] = traffic.get(roof["kernel"])
      roof["traffic_source"] = "profiles/r02_traffic.json: " + traffic.get("provenance", {}).get(roof["kernel"], "?")
  except Exception:  # pylint: disable=broad-except
    traffic = {}

  # ---- secondary: the single-gate kernel's roofline at the same size (one h per qubit, no fusion)
  single = None
  if wl["fusion"] and not args.no_secondary and world == 1:
    from qcc_b200 import workloads
    hs = _cabi.pack_xg_gates(workloads.hsweep(n))
    s.set_fusion(False)
    for _ in range(2):
      s.xg_apply_gates(hs)
    s.sync()
    s.profile_enable(True)
    s.profile_read(reset=True)
    s.timer_start()
    for _ in range(3):
      s.xg_apply_gates(hs)
    hms = s.timer_stop()
    hp = s.profile_read(reset=True)["apply1"]
    s.profile_enable(False)
    s.set_fusion(True)
    s.set_tile_bits(args.tile_bits)
    ach = hp["bytes"] / (hp["ms"] * 1e-3) / 1e9
    single = {"workload": f"h on each of {n} qubits, one k_apply_u launch per gate", "kernel": "k_apply_u",
              "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
              "launches": hp["launches"], "avg_launch_ms": hp["ms"] / hp["launches"],
              "gates_per_s": 3 * n / (hms * 1e-3),
              "traffic": traffic.get("k_apply_u") if n == traffic.get("qubits") else None}

  # ---- e2e: host buffers through the C ABI ------------------------------------------------
  e2e = None
  e2e_res = None
  if rank == 0 or world > 1:
    if not args.no_e2e:
      try:
        # every rank owns a pinned host buffer for ITS slice of the canonical vector (the whole vector on one GPU)
        cnt = (1 << n) // world
        host = _cabi.PinnedBuffer(cnt)
        host.array[:] = 0
        if rank == 0:
          host.array[5] = 1.0

        def e2e_step():
          # qb_copy_in of a whole shard resets the bit layout; qb_copy_out brings it back to the canonical one
          # first (exchange events on a sharded state), so the host always sees its slice of the logical vector
          _cabi.check(_cabi.lib().qb_copy_in(s._h, 0, cnt, host.array.ctypes.data))
          s.xg_apply_gates(packed)
          _cabi.check(_cabi.lib().qb_copy_out(s._h, 0, cnt, host.array.ctypes.data))

        e2e_step()                       # warm-up of the path
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
          e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        # the two PCIe copies on their own (nothing queued, canonical layout), so that the step can be read against them
        t1 = time.perf_counter()
        _cabi.check(_cabi.lib().qb_copy_in(s._h, 0, cnt, host.array.ctypes.data))
        t2 = time.perf_counter()
        _cabi.check(_cabi.lib().qb_copy_out(s._h, 0, cnt, host.array.ctypes.data))
        t3 = time.perf_counter()
        e2e = {"value": ngates * args.e2e_steps / dt, "unit": "gates/s",
               "h2d_bytes_per_step": (1 << n) * 16 + len(packed) * 80 * world, "d2h_bytes_per_step": (1 << n) * 16,
               "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
               "copy_in_ms": (t2 - t1) * 1e3, "copy_out_ms": (t3 - t2) * 1e3,
               "step_over_the_two_copies": dt / args.e2e_steps / (t3 - t1),
               "pcie_gbs_per_gpu": {"h2d": cnt * 16 / (t2 - t1) / 1e9, "d2h": cnt * 16 / (t3 - t2) / 1e9},
               "path": "pinned host complex128 state -> qb_copy_in -> qb_xg_apply_gates -> qb_copy_out "
                       "(what a host-buffer caller such as the libxgates/libq faces pays per circuit)" +
                       ("" if world == 1 else f"; each of the {world} ranks moves its 1/{world} slice of the canonical "
                        "vector, bytes are totals over the ranks, time is rank 0's wall clock between two barriers")}
        host.close()
      except Exception as ex:  # pylint: disable=broad-except
        e2e = {"value": None, "unit": "gates/s", "error": str(ex)[:200]}
    # resident flavour: what a circuit.qc()/libq user pays -- init label in, readout out
    t0 = time.perf_counter()
    reps = max(2, args.e2e_steps)
    for _ in range(reps):
      s.set_basis(5)
      s.xg_apply_gates(packed)
      s.argmax()
    dt = time.perf_counter() - t0
    e2e_res = {"value": ngates * reps / dt, "unit": "gates/s", "h2d_bytes_per_step": len(packed) * 80 + 16,
               "d2h_bytes_per_step": 16 * 1184, "ms_per_step": dt / reps * 1e3,
               "path": "qb_set_basis -> qb_xg_apply_gates -> qb_argmax (state stays in HBM, as with the "
                       "reference where it stays in one numpy array)"}

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    try:
      gps, desc = cpu_reference_sample(n, stream)
      if gps:
        cpu = {"value": gps, "unit": "gates/s", "cores": 1, "kind": "reference", "sample": desc,
               "host_cores": os.cpu_count()}
      else:
        cpu = {"value": None, "unit": "gates/s", "cores": 1, "kind": "reference", "sample": desc}
    except Exception as ex:  # pylint: disable=broad-except
      cpu = {"value": None, "unit": "gates/s", "cores": 1, "kind": "reference", "sample": f"failed: {ex}"[:200]}
    try:
      # in a child process: the reference's C code must not be able to take the bench line down with it
      out = subprocess.run([sys.executable, "-c",
                            "import json, bench; print(json.dumps(bench.cpu_reference_libq_sample()))"],
                           capture_output=True, text=True, timeout=120, cwd=ROOT)
      cpu["libq"] = json.loads(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else {
          "value": None, "sample": f"child exited {out.returncode}"}
    except Exception as ex:  # pylint: disable=broad-except
      cpu["libq"] = {"value": None, "sample": f"failed: {ex}"[:200]}

  if rank == 0:
    value = ngates * args.steps / (ms * 1e-3)
    xch = prof.get("exchange", {"launches": 0, "ms": 0.0, "bytes": 0.0})
    fpush = prof.get("fused_push", {"launches": 0, "ms": 0.0, "bytes": 0.0})
    exchange = None
    if world > 1:
      mode = s.exchange_mode()
      sent = c1["bytes_exchanged"] - c0["bytes_exchanged"]
      carrier_ms = xch["ms"] + fpush["ms"]     # kernels that carried exchange traffic (+ their barriers)
      exchange = {"mode": mode, "events_per_step": (c1["exchanges"] - c0["exchanges"]) / args.steps,
                  "bytes_sent_per_rank_per_step": sent / args.steps,
                  "fused_push_passes_per_step": fpush["launches"] / args.steps,
                  "fused_push_pass_avg_ms_rank0": fpush["ms"] / fpush["launches"] if fpush["launches"] else None,
                  "standalone_ms_per_step_rank0": xch["ms"] / args.steps,
                  "nvlink_gbs_per_direction_rank0": (sent / (carrier_ms * 1e-3) / 1e9) if carrier_ms else None,
                  "note": {"push": "an exchange event swaps k sharded bits with k local ones in ONE all-to-all: the "
                                   "store stage of the fused pass before it writes every amplitude to its "
                                   "destination rank's alternate buffer through CUDA IPC peer mappings (NVLink "
                                   "posted writes); bytes sent = (1 - 2^-k) of a shard per event; GB/s = bytes "
                                   "sent / time of the kernels that carried them (a pass + its exchange)",
                           "swap": "one in-place pair-swap kernel per (sharded bit, local bit) pair over peer mappings",
                           "nccl": "ncclSend/ncclRecv of half a shard per pair + copy-back"}[mode]}
    line = {
        "metric": "gate-applies/sec", "value": value, "unit": "gates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak" if (args.weak and world > 1) else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "qubits": n, "gates_per_step": ngates,
                   "state_bytes": (1 << n) * 16, "l2": "state (>= 4 GiB) >> 126 MB L2; no flush needed",
                   "fusion": wl["fusion"], "tile_bits": args.tile_bits if wl["fusion"] else None,
                   "shard_qubits": n_shard,
                   "parallelism": "1 GPU" if world == 1 else
                   f"one {n}-qubit state sharded over {world} GPUs by its top {int(np.log2(world))} index bits; "
                   "gates on sharded qubits cost an exchange event over NVLink (see exchange.mode)",
                   "queue": "flush per step" if args.flush_per_step else
                   "the K steps are queued and flushed once inside the timed region"},
        "passes_per_step": (c1["passes"] - c0["passes"]) / args.steps,
        "gpu_launches": c1["kernel_launches"] - c0["kernel_launches"],
        "achieved_gbs_algorithmic_by_gate": by_gate_alg,
        "achieved_gbs_swept": (c1["bytes_swept"] - c0["bytes_swept"]) / (ms * 1e-3) / 1e9,
        "roofline": roof, "roofline_single_gate": single, "kernel_ms": {k: v["ms"] for k, v in prof.items() if v["launches"]},
        "exchange": exchange, "cpu_baseline": cpu, "e2e": e2e, "e2e_resident": e2e_res, "clocks": clocks,
        "wall_s_timed_region": wall, "norm2_after": norm,
    }
    print(json.dumps(line))
  s.close()
  if dist is not None:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
