"""ctypes binding of the C ABI (include/qcc_b200.h).  No torch, no CPU fallback.

The shared library is built in-tree by ``make -C qcc_b200/csrc`` (or
``__graft_entry__.build()``) into ``qcc_b200/lib/libqcc_b200.so``.  If it is missing the
import fails loudly; if it loads but no B200 is visible, every state operation raises
``QbError`` with the library's own message.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libqcc_b200.so")

QB_OK = 0
QB_KCLASS = {"apply1": 0, "phase": 1, "fused": 2, "aux": 3, "exchange": 4, "fused_push": 5}


class QbError(RuntimeError):
  pass


class qb_gate(ctypes.Structure):
  _fields_ = [("ctl_mask", ctypes.c_uint64), ("target", ctypes.c_int32), ("flags", ctypes.c_int32),
              ("m", ctypes.c_double * 8)]


class qb_xg_gate(ctypes.Structure):
  _fields_ = [("kind", ctypes.c_int32), ("ctl", ctypes.c_int32), ("tgt", ctypes.c_int32),
              ("pad", ctypes.c_int32), ("m", ctypes.c_double * 8)]


class qb_counters(ctypes.Structure):
  _fields_ = [(n, ctypes.c_uint64) for n in
              ("gates_applied", "kernel_launches", "passes", "bytes_algorithmic", "bytes_swept",
               "exchanges", "bytes_exchanged")]


class qb_profile(ctypes.Structure):
  _fields_ = [("launches", ctypes.c_uint64 * 6), ("ms", ctypes.c_double * 6),
              ("bytes", ctypes.c_double * 6)]


_P = ctypes.c_void_p
_I = ctypes.c_int
_U64 = ctypes.c_uint64
_DP = ctypes.POINTER(ctypes.c_double)

# name -> argtypes (restype is int unless listed in _RESTYPES)
PROTOTYPES = {
    "qb_abi_version": [],
    "qb_last_error": [],
    "qb_device_count": [ctypes.POINTER(_I)],
    "qb_device_info": [_I, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(_I),
                       ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(_I), ctypes.POINTER(_I)],
    "qb_state_create": [_I, _U64, _I, ctypes.POINTER(_P)],
    "qb_comm_get_unique_id": [_P],
    "qb_state_create_sharded": [_I, _U64, _I, _I, _I, _P, ctypes.POINTER(_P)],
    "qb_state_layout": [_P, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I)],
    "qb_canonicalize": [_P],
    "qb_state_exchange_mode": [_P, ctypes.POINTER(_I)],
    "qb_shard_lower_json": [_I, _I, _I, ctypes.POINTER(qb_gate), ctypes.c_int64, _I, ctypes.c_char_p,
                            ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)],
    "qb_plan_check": [_I, ctypes.POINTER(qb_gate), ctypes.c_int64, _I, ctypes.POINTER(ctypes.c_int64)],
    "qb_fuse_gates": [ctypes.POINTER(qb_gate), ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)],
    "qb_shard_plan_stats": [_I, _I, _I, ctypes.POINTER(qb_gate), ctypes.c_int64, _I, _I, _I, _I,
                            ctypes.POINTER(ctypes.c_int64)],
    "qb_shard_event_dest": [_I, _I, _I, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I), _I, _P, _P,
                            ctypes.c_int64],
    "qb_state_destroy": [_P],
    "qb_state_nqubits": [_P, ctypes.POINTER(_I)],
    "qb_set_basis": [_P, _U64],
    "qb_fill_random": [_P, _U64],
    "qb_copy_in": [_P, _U64, _U64, _P],
    "qb_copy_out": [_P, _U64, _U64, _P],
    "qb_apply1": [_P, _I, _DP],
    "qb_applyc": [_P, _I, _I, _DP],
    "qb_applycc": [_P, _I, _I, _I, _DP],
    "qb_apply_gates": [_P, ctypes.POINTER(qb_gate), ctypes.c_int64],
    "qb_xg_apply1": [_P, _I, _DP],
    "qb_xg_applyc": [_P, _I, _I, _DP],
    "qb_xg_apply_gates": [_P, ctypes.POINTER(qb_xg_gate), ctypes.c_int64],
    "qb_set_fusion": [_P, _I],
    "qb_flush": [_P],
    "qb_sync": [_P],
    "qb_get_amplitude": [_P, _U64, _DP],
    "qb_norm2": [_P, _DP],
    "qb_argmax": [_P, ctypes.POINTER(_U64), _DP],
    "qb_prob_bit": [_P, _I, _DP],
    "qb_prob_bit_value": [_P, _I, _I, _DP],
    "qb_list_above": [_P, ctypes.c_double, _U64, _P, _P, ctypes.POINTER(_U64)],
    "qb_host_apply1": [_P, _P, _I, _I, _I, _I],
    "qb_host_applyc": [_P, _P, _I, _I, _I, _I, _I],
    "qb_host_run": [_P, _I, ctypes.POINTER(qb_xg_gate), ctypes.c_int64, _I],
    "qb_host_alloc": [ctypes.c_size_t, ctypes.POINTER(_P)],
    "qb_host_free": [_P],
    "qb_get_counters": [_P, ctypes.POINTER(qb_counters)],
    "qb_profile_enable": [_P, _I],
    "qb_profile_read": [_P, ctypes.POINTER(qb_profile), _I],
    "qb_timer_start": [_P],
    "qb_timer_stop": [_P, _DP],
    "qb_plan_json": [_I, ctypes.POINTER(qb_gate), ctypes.c_int64, _I, ctypes.c_char_p, ctypes.c_size_t,
                     ctypes.POINTER(ctypes.c_size_t)],
    "qb_set_tile_bits": [_P, _I],
}
_RESTYPES = {"qb_last_error": ctypes.c_char_p}

_lib = None


def lib():
  """The loaded C-ABI library.  Raises if it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise QbError(f"{LIB_PATH} is missing: build it with `make -C qcc_b200/csrc` "
                    "(or __graft_entry__.build()); there is no Python/CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    for name, args in PROTOTYPES.items():
      fn = getattr(L, name)
      fn.argtypes = args
      fn.restype = _RESTYPES.get(name, ctypes.c_int)
    _lib = L
  return _lib


def check(rc: int):
  if rc != QB_OK:
    raise QbError(f"qcc_b200 error {rc}: {lib().qb_last_error().decode()}")


def mat8(m) -> ctypes.Array:
  """2x2 complex (any array-like) -> the ABI's 8 doubles."""
  flat = np.ascontiguousarray(np.asarray(m, dtype=np.complex128).reshape(4))
  return (ctypes.c_double * 8)(*flat.view(np.float64))


def pack_xg_gates(gates) -> ctypes.Array:
  """gates: iterable of (kind, ctl, tgt, 2x2) in python qubit numbering."""
  gates = list(gates)
  arr = (qb_xg_gate * len(gates))()
  for k, g in enumerate(gates):
    kind, ctl, tgt, m = g[0], g[1], g[2], g[3]
    arr[k].kind = int(kind)
    arr[k].ctl = int(ctl) if ctl is not None else 0
    arr[k].tgt = int(tgt)
    flat = np.ascontiguousarray(np.asarray(m, dtype=np.complex128).reshape(4)).view(np.float64)
    for j in range(8):
      arr[k].m[j] = flat[j]
  return arr


# qb_xg_gate as a numpy record: 4 x int32 + 4 x complex128 = 80 bytes, no padding
XG_DTYPE = np.dtype([("kind", np.int32), ("ctl", np.int32), ("tgt", np.int32), ("pad", np.int32),
                     ("m", np.complex128, (4,))])
assert XG_DTYPE.itemsize == ctypes.sizeof(qb_xg_gate) == 80


class GateBuffer:
  """Host-side batch of gate records in the ABI's layout: the python face appends here (a few numpy
  stores per gate) and hands the whole batch to qb_xg_apply_gates in ONE foreign call, instead of one
  ctypes call + matrix conversion per gate."""

  def __init__(self, cap: int = 8192):
    self.cap = cap
    self.arr = np.zeros(cap, dtype=XG_DTYPE)
    self.kind, self.ctl, self.tgt, self.m = self.arr["kind"], self.arr["ctl"], self.arr["tgt"], self.arr["m"]
    self.n = 0

  def append(self, kind: int, ctl: int, tgt: int, gate) -> bool:
    """Returns True when the buffer is full and must be handed over."""
    k = self.n
    self.kind[k] = kind
    self.ctl[k] = ctl
    self.tgt[k] = tgt
    self.m[k] = np.asarray(gate, dtype=np.complex128).reshape(4)
    self.n = k + 1
    return self.n == self.cap


def pack_gates(gates) -> ctypes.Array:
  """gates: iterable of (ctl_mask, target_bit, 2x2) in index-bit numbering."""
  gates = list(gates)
  arr = (qb_gate * len(gates))()
  for k, (mask, tgt, m) in enumerate(gates):
    arr[k].ctl_mask = int(mask)
    arr[k].target = int(tgt)
    flat = np.ascontiguousarray(np.asarray(m, dtype=np.complex128).reshape(4)).view(np.float64)
    for j in range(8):
      arr[k].m[j] = flat[j]
  return arr


def plan_json(nqubits: int, gates, tile_bits: int = 12) -> str:
  """Fusion plan for index-bit gates (host only, no GPU needed)."""
  arr = gates if isinstance(gates, ctypes.Array) else pack_gates(gates)
  need = ctypes.c_size_t(0)
  check(lib().qb_plan_json(nqubits, arr, len(arr), tile_bits, None, 0, ctypes.byref(need)))
  buf = ctypes.create_string_buffer(need.value)
  check(lib().qb_plan_json(nqubits, arr, len(arr), tile_bits, buf, need.value, ctypes.byref(need)))
  return buf.value.decode()


def shard_lower_json(nqubits: int, nranks: int, rank: int, gates, canonicalize: bool = True) -> str:
  """Per-rank lowering of index-bit gates for a sharded state (host only)."""
  arr = gates if isinstance(gates, ctypes.Array) else pack_gates(gates)
  need = ctypes.c_size_t(0)
  check(lib().qb_shard_lower_json(nqubits, nranks, rank, arr, len(arr), int(canonicalize), None, 0,
                                  ctypes.byref(need)))
  buf = ctypes.create_string_buffer(need.value)
  check(lib().qb_shard_lower_json(nqubits, nranks, rank, arr, len(arr), int(canonicalize), buf, need.value,
                                  ctypes.byref(need)))
  return buf.value.decode()


def plan_check(nqubits: int, gates, tile_bits: int = 12) -> int:
  """Plan + host half of every pass's kernel launch (host only); returns the number of passes, raises QbError
  when a pass does not fit the kernel."""
  arr = gates if isinstance(gates, ctypes.Array) else pack_gates(gates)
  n = ctypes.c_int64(0)
  check(lib().qb_plan_check(nqubits, arr, len(arr), tile_bits, ctypes.byref(n)))
  return n.value


def fuse_gates(gates):
  """The flush peephole on a list of (ctl_mask, target_bit, 2x2): returns (fused list, runs replaced)."""
  arr = pack_gates(gates)
  n = ctypes.c_int64(0)
  check(lib().qb_fuse_gates(arr, len(arr), ctypes.byref(n)))
  out = [(int(g.ctl_mask), int(g.target), np.array(list(g.m)).view(np.complex128).reshape(2, 2)) for g in arr]
  return out, n.value


def shard_plan_stats(nqubits: int, nranks: int, rank: int, gates, tile_bits: int = 12, mode: str = "push") -> dict:
  """Events / pairs / passes a flush of `gates` (index-bit numbering) costs on one rank (host only)."""
  arr = gates if isinstance(gates, ctypes.Array) else pack_gates(gates)
  nl = nqubits - int(np.log2(nranks))
  window, hoist, prefetch = {"push": (min(max(6, nl - 5), nl - 3), 1, 1), "swap": (max(6, nl - 5), 1, 0),
                             "nccl": (6, 0, 0)}[mode]
  st = (ctypes.c_int64 * 8)()
  check(lib().qb_shard_plan_stats(nqubits, nranks, rank, arr, len(arr), tile_bits, window, hoist, prefetch, st))
  return dict(zip(("events", "pairs", "passes", "fused_passes", "events_on_a_pass", "rounds", "ops", "ccu_fused"), st))


def shard_event_dest(nlocal: int, nranks: int, rank: int, pairs, local: np.ndarray) -> np.ndarray:
  """Distributed destination index of each local index under one exchange event (host only).
  pairs: (rank bit, victim bit[, landing bit])."""
  rb = (ctypes.c_int * len(pairs))(*[int(p[0]) for p in pairs])
  vb = (ctypes.c_int * len(pairs))(*[int(p[1]) for p in pairs])
  lb = (ctypes.c_int * len(pairs))(*[int(p[2]) if len(p) > 2 else int(p[1]) for p in pairs])
  local = np.ascontiguousarray(local, dtype=np.uint64)
  dest = np.empty_like(local)
  check(lib().qb_shard_event_dest(nlocal, nranks, rank, rb, vb, lb, len(pairs), local.ctypes.data, dest.ctypes.data,
                                  local.size))
  return dest


def comm_unique_id() -> bytes:
  buf = ctypes.create_string_buffer(128)
  check(lib().qb_comm_get_unique_id(buf))
  return buf.raw


class PinnedBuffer:
  """Page-locked complex128 host vector (numpy view over cudaHostAlloc memory)."""

  def __init__(self, count: int):
    p = _P()
    check(lib().qb_host_alloc(count * 16, ctypes.byref(p)))
    self._p = p
    buf = (ctypes.c_double * (2 * count)).from_address(p.value)
    self.array = np.frombuffer(buf, dtype=np.complex128, count=count)

  def close(self):
    if getattr(self, "_p", None):
      self.array = None
      lib().qb_host_free(self._p)
      self._p = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


class DeviceState:
  """Owning handle of a device-resident 2^n complex128 amplitude vector."""

  def __init__(self, nqubits: int, init_label: int = 0, device: int = 0, *, rank: int = 0,
               nranks: int = 1, comm_id: bytes = None):
    """nranks > 1: this process holds shard `rank` of a state split over nranks GPUs (all
    calls on it are then collective); comm_id is comm_unique_id() of rank 0, shared out of band."""
    h = _P()
    if nranks > 1:
      idbuf = ctypes.create_string_buffer(comm_id, 128)
      check(lib().qb_state_create_sharded(nqubits, init_label, device, rank, nranks, idbuf, ctypes.byref(h)))
    else:
      check(lib().qb_state_create(nqubits, init_label, device, ctypes.byref(h)))
    self._h = h
    self.nqubits = nqubits
    self.rank, self.nranks = rank, nranks

  def close(self):
    if getattr(self, "_h", None):
      lib().qb_state_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self.close()

  # -- init / copies
  def set_basis(self, label: int):
    check(lib().qb_set_basis(self._h, label))

  def fill_random(self, seed: int):
    check(lib().qb_fill_random(self._h, seed))

  def copy_in(self, arr: np.ndarray, first: int = 0):
    a = np.ascontiguousarray(arr, dtype=np.complex128)
    check(lib().qb_copy_in(self._h, first, a.size, a.ctypes.data))

  def copy_out(self, first: int = 0, count: int | None = None) -> np.ndarray:
    if count is None:
      count = ((1 << self.nqubits) // self.nranks) - first
    out = np.empty(count, dtype=np.complex128)
    check(lib().qb_copy_out(self._h, first, count, out.ctypes.data))
    return out

  # -- gates (index-bit numbering)
  def apply1(self, target: int, m):
    check(lib().qb_apply1(self._h, target, mat8(m)))

  def applyc(self, control: int, target: int, m):
    check(lib().qb_applyc(self._h, control, target, mat8(m)))

  def applycc(self, c0: int, c1: int, target: int, m):
    check(lib().qb_applycc(self._h, c0, c1, target, mat8(m)))

  def apply_gates(self, packed):
    check(lib().qb_apply_gates(self._h, packed, len(packed)))

  # -- gates (python numbering)
  def xg_apply1(self, tgt: int, m):
    check(lib().qb_xg_apply1(self._h, tgt, mat8(m)))

  def xg_applyc(self, ctl: int, tgt: int, m):
    check(lib().qb_xg_applyc(self._h, ctl, tgt, mat8(m)))

  def xg_apply_gates(self, packed):
    check(lib().qb_xg_apply_gates(self._h, packed, len(packed)))

  def xg_apply_buffer(self, buf: "GateBuffer"):
    """Hand over (and empty) a GateBuffer."""
    if buf.n:
      n, buf.n = buf.n, 0
      check(lib().qb_xg_apply_gates(self._h, ctypes.cast(buf.arr.ctypes.data, ctypes.POINTER(qb_xg_gate)), n))

  # -- sharding
  def layout(self) -> dict:
    nl, r, nr = _I(), _I(), _I()
    perm = (_I * self.nqubits)()
    check(lib().qb_state_layout(self._h, ctypes.byref(nl), ctypes.byref(r), ctypes.byref(nr), perm))
    return {"nlocal": nl.value, "rank": r.value, "nranks": nr.value, "perm": list(perm)}

  def exchange_mode(self) -> str:
    m = _I()
    check(lib().qb_state_exchange_mode(self._h, ctypes.byref(m)))
    return {-1: "none", 0: "nccl", 1: "swap", 2: "push"}[m.value]

  def canonicalize(self):
    check(lib().qb_canonicalize(self._h))

  # -- queue
  def set_fusion(self, on: bool):
    check(lib().qb_set_fusion(self._h, 1 if on else 0))

  def set_tile_bits(self, k: int):
    check(lib().qb_set_tile_bits(self._h, k))

  def flush(self):
    check(lib().qb_flush(self._h))

  def sync(self):
    check(lib().qb_sync(self._h))

  # -- readouts
  def amplitude(self, index: int) -> complex:
    out = (ctypes.c_double * 2)()
    check(lib().qb_get_amplitude(self._h, index, out))
    return complex(out[0], out[1])

  def norm2(self) -> float:
    out = ctypes.c_double()
    check(lib().qb_norm2(self._h, ctypes.byref(out)))
    return out.value

  def argmax(self):
    idx = _U64()
    p = ctypes.c_double()
    check(lib().qb_argmax(self._h, ctypes.byref(idx), ctypes.byref(p)))
    return idx.value, p.value

  def prob_bit(self, bit: int) -> float:
    out = ctypes.c_double()
    check(lib().qb_prob_bit(self._h, bit, ctypes.byref(out)))
    return out.value

  def prob_bit_value(self, bit: int, value: int) -> float:
    out = ctypes.c_double()
    check(lib().qb_prob_bit_value(self._h, bit, value, ctypes.byref(out)))
    return out.value

  def list_above(self, threshold: float, cap: int = 1 << 16):
    labels = np.zeros(cap, dtype=np.uint64)
    amps = np.zeros(cap, dtype=np.complex128)
    cnt = _U64()
    check(lib().qb_list_above(self._h, threshold, cap, labels.ctypes.data, amps.ctypes.data,
                              ctypes.byref(cnt)))
    n = min(cnt.value, cap)
    return labels[:n], amps[:n], cnt.value

  # -- measurement
  def counters(self) -> dict:
    c = qb_counters()
    check(lib().qb_get_counters(self._h, ctypes.byref(c)))
    return {n: getattr(c, n) for n, _ in qb_counters._fields_}

  def profile_enable(self, on: bool):
    check(lib().qb_profile_enable(self._h, 1 if on else 0))

  def profile_read(self, reset: bool = True) -> dict:
    p = qb_profile()
    check(lib().qb_profile_read(self._h, ctypes.byref(p), 1 if reset else 0))
    return {name: {"launches": p.launches[k], "ms": p.ms[k], "bytes": p.bytes[k]}
            for name, k in QB_KCLASS.items()}

  def timer_start(self):
    check(lib().qb_timer_start(self._h))

  def timer_stop(self) -> float:
    ms = ctypes.c_double()
    check(lib().qb_timer_stop(self._h, ctypes.byref(ms)))
    return ms.value
