// engine.cu -- host side of the C ABI (include/qcc_b200.h): state lifetime, the gate
// queue, dispatch to single-gate kernels or fused passes, readouts, the host-buffer
// entry points the `libxgates` shim binds, and the measurement helpers.
//
// Reference behaviour mirrored here (paths relative to the reference tree):
//   * gate semantics: src/lib/xgates.cc:23-67 (dense butterfly, python numbering,
//     negative-control predicate) and src/libq/gates.cc:9-146 (named gates);
//   * queue / flush protocol: src/libq/gates_jit.cc:53-132 -- gates may be deferred
//     and are guaranteed applied at flush and before anything observes the state;
//   * readouts: src/lib/state.py:30-78 (ampl / prob / maxprob), src/libq/qureg.cc:64-86.
// There is no CPU execution path in this file: without a CUDA device every state
// operation fails with QB_ERR_CUDA.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/qcc_b200.h"
#include "comm.h"
#include "kernels.h"
#include "planner.h"
#include "qb_types.h"
#include "shard.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return fail(QB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                 \
  } while (0)

#define QB(call)            \
  do {                      \
    int rc__ = (call);      \
    if (rc__ != QB_OK) return rc__; \
  } while (0)

#define NC(api, call)                                                                   \
  do {                                                                                  \
    ncclResult_t r__ = (call);                                                          \
    if (r__ != ncclSuccess)                                                             \
      return fail(QB_ERR_COMM, "%s failed: %s (%s:%d)", #call, (api)->GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

struct ProfRec {
  int kclass;
  double bytes;
  cudaEvent_t e0, e1;
  std::string tag;   // QCC_B200_TRACE_FLUSH: what this launch was (printed with its time)
};

}  // namespace

struct qb_state {
  int n = 0;            // LOCAL index bits of this rank's shard (== nq when not sharded)
  uint64_t len = 0;     // 2^n amplitudes held here
  int nq = 0;           // logical qubits of the whole state
  int device = 0;
  // sharding (shard.h): rank bits are the top physical index bits
  int rank = 0, nranks = 1, p = 0;
  std::vector<int> perm;          // logical index bit -> physical index bit
  uint32_t flip = 0;              // relabelled rank bits (shard.h): rank bit k carries the negated qubit
  int victim_window = qb::kVictimWindow;  // wider when exchanges go through peer memory
  // How exchange events run (decided collectively at creation, QCC_B200_EXCHANGE=nccl|swap|push overrides):
  //   QB_X_NCCL  ncclSend/ncclRecv per pair into a half-shard bounce buffer + copy-back
  //   QB_X_SWAP  one in-place kernel per pair over CUDA IPC peer mappings (k_pair_swap)
  //   QB_X_PUSH  the state is double-buffered; an event (any number of pairs) is ONE out-of-place all-to-all
  //              written by the store stage of the last fused pass before it (or k_push_remap)
  int xmode = 0;
  double2 *base = nullptr;        // the allocation: one shard, or two (push)
  int bufsel = 0;                 // push: which half of `base` psi is (all ranks flip together)
  ncclComm_t comm = nullptr;
  double2 *xbuf = nullptr;        // half-shard receive buffer for exchanges
  size_t xbuf_bytes = 0;
  cudaStream_t xstream = nullptr; // copy-back of received pieces overlaps the next piece's transfer
  std::vector<cudaEvent_t> xevents;
  // CUDA IPC mappings of the other ranks' allocations (swap / push)
  std::vector<double2 *> peer_base; // [rank] -> mapping of its `base` (our own pointer for ourselves)
  int nbuf = 1;                     // shards in `base` (2: push); the barrier counters sit behind them
  unsigned long long xbar_epoch = 0;
  bool peer_barrier = false;        // barriers through the counters in peer memory instead of an NCCL all-reduce
  double *d_sync = nullptr;        // scratch of the stream-ordered cross-rank barrier
  cudaStream_t stream = nullptr;
  double2 *psi = nullptr;
  bool fusion = true;
  int tile_bits = 12;
  std::vector<QbGate> queue;
  // scratch
  double *d_scalar = nullptr;            // 1 double
  unsigned long long *d_counter = nullptr;
  double *d_blk_prob = nullptr;
  uint64_t *d_blk_idx = nullptr;
  uint64_t *d_list_labels = nullptr;     // qb_list_above output buffers, kept between calls
  double2 *d_list_amps = nullptr;
  uint64_t list_cap = 0;
  // fused-pass staging (device copies of the current plan)
  void *d_plan = nullptr;
  size_t d_plan_cap = 0;
  void *h_plan = nullptr;  // pinned
  size_t h_plan_cap = 0;
  cudaEvent_t plan_free = nullptr;  // signalled when the staged plan has been consumed
  // measurement
  qb_counters cnt{};
  bool profiling = false;
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> event_pool;
  qb_profile prof{};
  cudaEvent_t t0 = nullptr, t1 = nullptr;
};

namespace {

int classify(const double m[8]) { return qb::classify_matrix(m); }

double gate_bytes(const qb_state *s, const QbGate &g) {
  // general 32N, controlled-general / 1-bit diagonal 16N, ...
  int nb = __builtin_popcountll(g.ctl_mask);
  double n = double(s->len) * 32.0;
  if (g.kind == QB_K_NOP) return 0.0;
  if (g.kind == QB_K_PHASE) nb += 1;
  return n / double(uint64_t(1) << nb);
}

cudaEvent_t get_event(qb_state *s) {
  if (!s->event_pool.empty()) {
    cudaEvent_t e = s->event_pool.back();
    s->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

struct ProfScope {
  qb_state *s;
  ProfRec rec;
  bool on;
  ProfScope(qb_state *st, int kclass, double bytes, const std::string &tag = std::string()) : s(st), on(st->profiling) {
    if (on) {
      rec.kclass = kclass;
      rec.bytes = bytes;
      rec.tag = tag;
      rec.e0 = get_event(s);
      rec.e1 = get_event(s);
      cudaEventRecord(rec.e0, s->stream);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(rec.e1, s->stream);
      s->prof_pending.push_back(rec);
    }
  }
};

int resolve_profile(qb_state *s) {
  for (auto &r : s->prof_pending) {
    float ms = 0.f;
    CU(cudaEventSynchronize(r.e1));
    CU(cudaEventElapsedTime(&ms, r.e0, r.e1));
    s->prof.launches[r.kclass] += 1;
    s->prof.ms[r.kclass] += ms;
    s->prof.bytes[r.kclass] += r.bytes;
    if (!r.tag.empty()) fprintf(stderr, "qcc_b200 launch: %.3f ms  %s\n", ms, r.tag.c_str());
    s->event_pool.push_back(r.e0);
    s->event_pool.push_back(r.e1);
  }
  s->prof_pending.clear();
  return QB_OK;
}

int run_single(qb_state *s, const QbGate &g) {
  if (g.kind == QB_K_NOP) {
    s->cnt.gates_applied += 1;
    return QB_OK;
  }
  double bytes = gate_bytes(s, g);
  {
    ProfScope ps(s, g.kind == QB_K_PHASE ? QB_KCLASS_PHASE : QB_KCLASS_APPLY1, bytes);
    CU(qb::launch_gate(s->psi, s->n, g, true, s->stream));
  }
  s->cnt.gates_applied += 1;
  s->cnt.kernel_launches += 1;
  s->cnt.passes += 1;
  s->cnt.bytes_algorithmic += uint64_t(bytes);
  s->cnt.bytes_swept += uint64_t(bytes);
  return QB_OK;
}

int ensure_plan_buffers(qb_state *s, size_t bytes) {
  if (bytes > s->h_plan_cap) {
    if (s->h_plan) cudaFreeHost(s->h_plan);
    if (s->d_plan) cudaFree(s->d_plan);
    size_t cap = std::max<size_t>(bytes * 2, 1 << 20);
    CU(cudaMallocHost(&s->h_plan, cap));
    CU(cudaMalloc(&s->d_plan, cap));
    s->h_plan_cap = s->d_plan_cap = cap;
  }
  return QB_OK;
}

enum { QB_X_NCCL = 0, QB_X_SWAP = 1, QB_X_PUSH = 2 };
constexpr size_t kBarrierBytes = 4096;   // arrival counters of the peer-memory barrier, behind the shard(s)

// Fused execution: plan the queue into tile-resident passes and launch one kernel per pass.  With
// `push` (an exchange event follows these gates, push mode) the LAST pass, when it is a fused pass,
// stores through the event's bit permutation into the alternate buffers; *pushed says whether it did.
double host_now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return double(ts.tv_sec) * 1e3 + double(ts.tv_nsec) * 1e-6;
}

// One run of local gates of a flush.  All segments of a flush are planned first and staged as ONE blob with one
// upload (stage_segments), so that the host never waits for the device between the segments of a sharded flush.
struct Segment {
  const std::vector<QbGate> *gates = nullptr;
  qb::Plan plan;
  bool planned = false;      // false: gate by gate (fusion off, or a shard too small for a tile)
  size_t blob_off = 0;
};

int stage_segments(qb_state *s, std::vector<Segment> *segs, const std::vector<char> &event_follows) {
  static const bool trace = getenv("QCC_B200_TRACE_FLUSH") != nullptr;
  const double t_a = trace ? host_now_ms() : 0.0;
  size_t total = 0, npass = 0, ngates = 0;
  for (size_t k = 0; k < segs->size(); ++k) {
    Segment &sg = (*segs)[k];
    if (!sg.gates || sg.gates->empty() || !(s->fusion && s->n > QB_TILE_LOW)) continue;
    // an exchange event follows: the segment's last pass is made a fused pass whenever it has anything to do,
    // so that the event can ride on its store stage
    qb::plan_gates(s->n, sg.gates->data(), int64_t(sg.gates->size()), s->tile_bits, &sg.plan, event_follows[k] != 0);
    sg.planned = true;
    sg.blob_off = total;
    total += (sg.plan.blob_bytes() + 255) & ~size_t(255);
    npass += sg.plan.passes.size();
    ngates += sg.gates->size();
  }
  if (!total) return QB_OK;
  const double t_b = trace ? host_now_ms() : 0.0;
  // The staging buffers are reused flush after flush: wait until the previous flush's kernels have consumed them.
  CU(cudaEventSynchronize(s->plan_free));
  QB(ensure_plan_buffers(s, total));
  const double t_c = trace ? host_now_ms() : 0.0;
  for (Segment &sg : *segs)
    if (sg.planned) sg.plan.serialize(static_cast<char *>(s->h_plan) + sg.blob_off);
  CU(cudaMemcpyAsync(s->d_plan, s->h_plan, total, cudaMemcpyHostToDevice, s->stream));
  if (trace)
    fprintf(stderr, "qcc_b200 stage: %zu gates in %zu segment(s), %zu passes: plan %.2f ms, wait for the previous flush "
            "%.2f ms, stage %zu KiB %.2f ms\n", ngates, segs->size(), npass, t_b - t_a, t_c - t_b, total >> 10,
            host_now_ms() - t_c);
  return QB_OK;
}

// Launch the passes of one staged segment.  With `push` (an exchange event follows these gates, push mode) the
// LAST pass, when it is a fused pass, stores through the event's bit permutation into the alternate buffers;
// *pushed says whether it did.
int launch_segment(qb_state *s, const Segment &sg, const qb::PushMap *push, bool *pushed) {
  if (!sg.gates || sg.gates->empty()) return QB_OK;
  if (!sg.planned) {
    for (const QbGate &g : *sg.gates) QB(run_single(s, g));
    return QB_OK;
  }
  const qb::Plan &plan = sg.plan;
  const char *dbase = static_cast<const char *>(s->d_plan) + sg.blob_off;
  for (size_t k = 0; k < plan.passes.size(); ++k) {
    const qb::PlannedPass &pp = plan.passes[k];
    if (pp.single_gate >= 0) {
      // planner decided this gate is better off as a plain sweep (e.g. a lone PHASE)
      QB(run_single(s, pp.single));
      s->cnt.gates_applied += uint64_t(pp.ngates - 1);  // no-op gates retired alongside
      continue;
    }
    qb::DevicePass dp;
    dp.desc = pp.desc;
    dp.ops = pp.ops.data();
    dp.rounds = pp.rounds.data();
    dp.tables = pp.desc.ntable ? reinterpret_cast<const double2 *>(dbase + pp.tables_off) : nullptr;
    dp.outph = pp.outph.empty() ? nullptr : reinterpret_cast<const double2 *>(dbase + pp.outph_off);
    dp.jbtab = reinterpret_cast<const uint32_t *>(dbase + pp.jbtab_off);
    const bool carry = push && k + 1 == plan.passes.size();
    if (carry) dp.push = push;
    double sweep = double(s->len) * 32.0;
    std::string tag;
    static const bool trace = getenv("QCC_B200_TRACE_FLUSH") != nullptr;
    if (trace && s->profiling) {
      tag = carry ? "push pass, tile bits" : "pass, tile bits";
      for (int b = 0; b < pp.desc.K; ++b) tag += " " + std::to_string(pp.desc.tile_bits[b]);
      if (carry) {
        tag += "; moved";
        for (int b = 0; b < push->nmoved; ++b) tag += " " + std::to_string(push->src[b]) + "->" + std::to_string(push->dst[b]);
      }
      tag += "; rounds " + std::to_string(pp.desc.nrounds) + " ops " + std::to_string(pp.desc.nops);
    }
    {
      ProfScope ps(s, carry ? QB_KCLASS_FUSED_PUSH : QB_KCLASS_FUSED, sweep, tag);
      CU(qb::launch_fused_pass(s->psi, s->n, dp, s->stream));
    }
    if (carry && pushed) *pushed = true;
    s->cnt.kernel_launches += 1;
    s->cnt.passes += 1;
    s->cnt.bytes_swept += uint64_t(sweep);
    s->cnt.gates_applied += uint64_t(pp.ngates);
    s->cnt.bytes_algorithmic += uint64_t(pp.bytes_algorithmic_per_amp * double(s->len));
  }
  return QB_OK;
}

int run_local(qb_state *s, const std::vector<QbGate> &q) {
  if (q.empty()) return QB_OK;
  std::vector<Segment> segs(1);
  segs[0].gates = &q;
  QB(stage_segments(s, &segs, std::vector<char>(1, 0)));
  QB(launch_segment(s, segs[0], nullptr, nullptr));
  if (segs[0].planned) CU(cudaEventRecord(s->plan_free, s->stream));
  return QB_OK;
}

// min over ranks of *val (collective; synchronises the stream)
int agree_min(qb_state *s, double *val) {
  const qb::NcclApi *nc = qb::nccl_api(nullptr);
  if (!nc || !s->comm) return fail(QB_ERR_COMM, "no communicator");
  if (!s->d_sync) CU(cudaMalloc(&s->d_sync, sizeof(double)));
  CU(cudaMemcpyAsync(s->d_sync, val, sizeof *val, cudaMemcpyHostToDevice, s->stream));
  NC(nc, nc->AllReduce(s->d_sync, s->d_sync, 1, ncclDouble, ncclMin, s->comm, s->stream));
  CU(cudaMemcpyAsync(val, s->d_sync, sizeof *val, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return QB_OK;
}

// Stream-ordered barrier over all ranks: everything the ranks enqueued before it (their kernels' writes
// into peer memory included -- a kernel's stores are performed when it completes) is done before anything
// enqueued after it starts.
int stream_barrier(qb_state *s, cudaStream_t st = nullptr) {
  if (!st) st = s->stream;
  if (s->peer_barrier) {
    unsigned long long *rows[qb::kPushMaxRanks];
    for (int q = 0; q < s->nranks; ++q)
      rows[q] = reinterpret_cast<unsigned long long *>(s->peer_base[size_t(q)] + uint64_t(s->nbuf) * s->len);
    CU(qb::launch_peer_barrier(rows, s->rank, s->nranks, ++s->xbar_epoch, st));
    return QB_OK;
  }
  const qb::NcclApi *nc = qb::nccl_api(nullptr);
  if (!nc || !s->comm) return fail(QB_ERR_COMM, "no communicator");
  NC(nc, nc->AllReduce(s->d_sync, s->d_sync, 1, ncclDouble, ncclMin, s->comm, st));
  return QB_OK;
}

// Map every other rank's allocation into this process (CUDA IPC over NVLink peer access).  Collective:
// the outcome is agreed on (min over ranks), so either every rank uses peer memory or none does.
int map_peers(qb_state *s, bool *mapped) {
  *mapped = false;
  const qb::NcclApi *nc = qb::nccl_api(nullptr);
  if (!nc || !s->comm) return fail(QB_ERR_COMM, "no communicator");
  s->peer_base.assign(size_t(s->nranks), nullptr);
  std::vector<cudaIpcMemHandle_t> handles(static_cast<size_t>(s->nranks));
  unsigned char *dh = nullptr;
  CU(cudaMalloc(&dh, sizeof(cudaIpcMemHandle_t) * size_t(s->nranks)));
  double ok = 1.0;
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, s->base) != cudaSuccess) {
    cudaGetLastError();
    memset(&mine, 0, sizeof mine);
    ok = 0.0;
  }
  CU(cudaMemcpyAsync(dh + sizeof mine * size_t(s->rank), &mine, sizeof mine, cudaMemcpyHostToDevice, s->stream));
  NC(nc, nc->AllGather(dh + sizeof mine * size_t(s->rank), dh, sizeof mine, ncclUint8, s->comm, s->stream));
  CU(cudaMemcpyAsync(handles.data(), dh, sizeof mine * size_t(s->nranks), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  cudaFree(dh);
  QB(agree_min(s, &ok));   // nobody opens a handle that some rank could not make
  if (ok != 0.0) {
    ok = 1.0;
    for (int r = 0; r < s->nranks; ++r) {
      if (r == s->rank) {
        s->peer_base[size_t(r)] = s->base;
        continue;
      }
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, handles[size_t(r)], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0.0;
        break;
      }
      s->peer_base[size_t(r)] = static_cast<double2 *>(ptr);
    }
    QB(agree_min(s, &ok));
  }
  if (ok != 0.0) {
    *mapped = true;
    return QB_OK;
  }
  for (int r = 0; r < s->nranks; ++r)
    if (r != s->rank && s->peer_base[size_t(r)]) cudaIpcCloseMemHandle(s->peer_base[size_t(r)]);
  s->peer_base.clear();
  if (s->rank == 0) fprintf(stderr, "qcc_b200: CUDA IPC peer mapping unavailable, exchanges use NCCL send/recv\n");
  return QB_OK;
}

// One pair of an event, in place: swap physical global bit (n + rank_bit) with local bit `victim` -- every
// rank trades the half of its shard whose victim bit differs from its own rank bit with rank ^ (1 << rank_bit).
int do_exchange(qb_state *s, int rank_bit, int victim) {
  std::string why;
  const qb::NcclApi *nc = qb::nccl_api(&why);
  if (!nc || !s->comm) return fail(QB_ERR_COMM, "exchange without a communicator: %s", why.c_str());
  const int b = (s->rank >> rank_bit) & 1;
  const int partner = s->rank ^ (1 << rank_bit);
  const size_t half_bytes = size_t(s->len / 2) * sizeof(double2);
  if (s->xmode != QB_X_NCCL) {
    // ONE kernel swaps our outgoing half with the partner's in place, through the peer mapping: no
    // receive buffer, no copy-back, both NVLink directions busy (we read and write the partner's shard
    // for one half of the elements, it reads and writes ours for the other half).  The two stream-ordered
    // barriers fence it: nobody touches a shard its owner is still computing on, and nobody computes
    // on a shard its partner is still swapping into.
    ProfScope ps(s, QB_KCLASS_EXCHANGE, double(half_bytes));
    double2 *peer = s->peer_base[size_t(partner)] + (s->bufsel ? s->len : 0);
    QB(stream_barrier(s));
    CU(qb::launch_pair_swap(s->psi, peer, s->n, victim, b ? 0 : 1, b, s->stream));
    QB(stream_barrier(s));
    s->cnt.exchanges += 1;
    s->cnt.bytes_exchanged += half_bytes;
    s->cnt.kernel_launches += 1;
    return QB_OK;
  }
  // NCCL: the half is 2^(n-1-victim) contiguous runs of 2^victim amplitudes; they are received into
  // xbuf (NCCL must not write into memory it is still sending from) and copied back in place.
  const uint64_t run = uint64_t(1) << victim;
  const uint64_t nruns = uint64_t(1) << (s->n - 1 - victim);
  const uint64_t sel = b ? 0 : 1;  // we give away the half whose victim bit is NOT our rank bit
  if (s->xbuf_bytes < half_bytes) {
    if (s->xbuf) cudaFree(s->xbuf);
    s->xbuf = nullptr;
    s->xbuf_bytes = 0;
    cudaError_t e = cudaMalloc(&s->xbuf, half_bytes);
    if (e != cudaSuccess) return fail(QB_ERR_NOMEM, "exchange buffer of %zu MiB: %s", half_bytes >> 20, cudaGetErrorString(e));
    s->xbuf_bytes = half_bytes;
  }
  {
    // Pipelined in pieces of <= 1 GiB: while piece k+1 crosses NVLink, piece k is copied from
    // the receive buffer back into the shard on a second stream.
    ProfScope ps(s, QB_KCLASS_EXCHANGE, double(half_bytes));
    if (!s->xstream) CU(cudaStreamCreateWithFlags(&s->xstream, cudaStreamNonBlocking));
    const uint64_t piece = std::min<uint64_t>(run, uint64_t(1) << 26);
    const uint64_t per_run = run / piece;
    const uint64_t npieces = nruns * per_run;
    while (s->xevents.size() < 2) {
      cudaEvent_t e = nullptr;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      s->xevents.push_back(e);
    }
    for (uint64_t k = 0; k < npieces; ++k) {
      const uint64_t h = k / per_run, w = k % per_run;
      const uint64_t off = (h << (victim + 1)) | (sel << victim) | (w * piece);
      double2 *rx = s->xbuf + k * piece;
      NC(nc, nc->GroupStart());
      NC(nc, nc->Send(s->psi + off, size_t(piece) * 2, ncclDouble, partner, s->comm, s->stream));
      NC(nc, nc->Recv(rx, size_t(piece) * 2, ncclDouble, partner, s->comm, s->stream));
      NC(nc, nc->GroupEnd());
      CU(cudaEventRecord(s->xevents[0], s->stream));
      CU(cudaStreamWaitEvent(s->xstream, s->xevents[0], 0));
      CU(cudaMemcpyAsync(s->psi + off, rx, size_t(piece) * sizeof(double2), cudaMemcpyDeviceToDevice, s->xstream));
    }
    CU(cudaEventRecord(s->xevents[1], s->xstream));
    CU(cudaStreamWaitEvent(s->stream, s->xevents[1], 0));
  }
  s->cnt.exchanges += 1;
  s->cnt.bytes_exchanged += half_bytes;
  s->cnt.kernel_launches += 1;
  return QB_OK;
}

// The bit permutation of an exchange event as the store map of one rank (push mode); out[] is left empty.
// Pair k: local victim bit -> rank bit, rank bit -> landing bit, and (landing != victim) landing bit -> victim bit.
int event_map(int nl, int p, int rank, const int *rank_bits, const int *victims, const int *lands, size_t np,
              qb::PushMap *m) {
  *m = qb::PushMap();
  if (np == 0 || 2 * np > size_t(qb::kPushMaxMoved)) return fail(QB_ERR_ARG, "exchange event with %zu pairs", np);
  m->nl = nl;
  uint64_t rt = 0;
  for (int b = 0; b < p; ++b) {
    int dest = nl + b;   // a rank bit outside the event stays a rank bit
    for (size_t k = 0; k < np; ++k)
      if (rank_bits[k] == b) dest = lands ? lands[k] : victims[k];
    rt |= uint64_t((rank >> b) & 1) << dest;
  }
  m->rank_term = rt;
  uint64_t seen = 0;
  for (size_t k = 0; k < np; ++k) {
    const int land = lands ? lands[k] : victims[k];
    if (victims[k] < QB_TILE_LOW || victims[k] >= nl || land < QB_TILE_LOW || land >= nl || rank_bits[k] < 0 ||
        rank_bits[k] >= p)
      return fail(QB_ERR_ARG, "exchange pair (rank bit %d, victim bit %d, landing bit %d)", rank_bits[k], victims[k], land);
    for (size_t j = 0; j < k; ++j)
      if (rank_bits[j] == rank_bits[k]) return fail(QB_ERR_ARG, "exchange pairs overlap");
    if ((seen >> victims[k] & 1) || (land != victims[k] && (seen >> land & 1))) return fail(QB_ERR_ARG, "exchange pairs overlap");
    seen |= (uint64_t(1) << victims[k]) | (uint64_t(1) << land);
    m->src[m->nmoved] = victims[k];
    m->dst[m->nmoved] = nl + rank_bits[k];
    m->moved_mask |= uint64_t(1) << victims[k];
    ++m->nmoved;
    if (land != victims[k]) {
      m->src[m->nmoved] = land;
      m->dst[m->nmoved] = victims[k];
      m->moved_mask |= uint64_t(1) << land;
      ++m->nmoved;
    }
  }
  return QB_OK;
}

int make_push_map(const qb_state *s, const qb::ShardStep &ev, qb::PushMap *m) {
  if (ev.rank_bits.size() != ev.victims.size() || ev.lands.size() != ev.victims.size())
    return fail(QB_ERR_ARG, "bad exchange event");
  QB(event_map(s->n, s->p, s->rank, ev.rank_bits.data(), ev.victims.data(), ev.lands.data(), ev.rank_bits.size(), m));
  for (int r = 0; r < s->nranks; ++r) m->out[r] = s->peer_base[size_t(r)] + (s->bufsel ? 0 : s->len);
  return QB_OK;
}

// After every rank's push of an event has been enqueued: fence, then the alternate buffers are the state.
int finish_push(qb_state *s, const qb::ShardStep &ev) {
  {
    ProfScope ps(s, QB_KCLASS_EXCHANGE, 0.0);
    QB(stream_barrier(s));
  }
  s->bufsel ^= 1;
  s->psi = s->base + (s->bufsel ? s->len : 0);
  s->cnt.exchanges += 1;
  const uint64_t shard_bytes = uint64_t(s->len) * sizeof(double2);
  s->cnt.bytes_exchanged += shard_bytes - (shard_bytes >> ev.rank_bits.size());
  return QB_OK;
}

int do_event(qb_state *s, const qb::ShardStep &ev) {
  if (s->xmode != QB_X_PUSH) {
    for (size_t k = 0; k < ev.rank_bits.size(); ++k) QB(do_exchange(s, ev.rank_bits[k], ev.victims[k]));
    return QB_OK;
  }
  qb::PushMap pm;
  QB(make_push_map(s, ev, &pm));
  {
    const double shard_bytes = double(s->len) * sizeof(double2);
    ProfScope ps(s, QB_KCLASS_EXCHANGE, shard_bytes - shard_bytes / double(uint64_t(1) << ev.rank_bits.size()));
    CU(qb::launch_push_remap(s->psi, pm, s->stream));
  }
  s->cnt.kernel_launches += 1;
  return finish_push(s, ev);
}

int run_steps(qb_state *s, const std::vector<qb::ShardStep> &steps) {
  static const bool no_fuse = getenv("QCC_B200_NO_PUSH_FUSE") != nullptr;
  std::vector<Segment> segs(steps.size());
  std::vector<char> event_follows(steps.size(), 0);
  for (size_t k = 0; k < steps.size(); ++k) {
    if (steps[k].kind != 0) continue;
    segs[k].gates = &steps[k].gates;
    event_follows[k] = s->xmode == QB_X_PUSH && !no_fuse && k + 1 < steps.size() && steps[k + 1].kind == 1;
  }
  QB(stage_segments(s, &segs, event_follows));
  bool any_planned = false;
  for (size_t k = 0; k < steps.size(); ++k) {
    const qb::ShardStep &st = steps[k];
    if (st.kind == 1) {
      QB(do_event(s, st));
      continue;
    }
    any_planned = any_planned || segs[k].planned;
    // push mode: the event that follows these gates rides on the store stage of their last pass
    const qb::ShardStep *ev = event_follows[k] ? &steps[k + 1] : nullptr;
    qb::PushMap pm;
    if (ev) QB(make_push_map(s, *ev, &pm));
    bool pushed = false;
    const uint64_t before = s->cnt.gates_applied;
    QB(launch_segment(s, segs[k], ev ? &pm : nullptr, &pushed));
    s->cnt.gates_applied = before + uint64_t(st.retired);
    if (pushed) {
      QB(finish_push(s, *ev));
      ++k;
    }
  }
  if (any_planned) CU(cudaEventRecord(s->plan_free, s->stream));
  return QB_OK;
}

void make_layout(const qb_state *s, qb::ShardLayout *L) {
  L->n = s->nq;
  L->nl = s->n;
  L->p = s->p;
  L->rank = s->rank;
  L->perm = s->perm;
  L->flip = s->flip;
  L->window = s->victim_window;
  L->hoist = s->xmode != QB_X_NCCL ? 1 : 0;
  L->prefetch = s->xmode == QB_X_PUSH ? 1 : 0;
  if (getenv("QCC_B200_NO_PREFETCH")) L->prefetch = 0;
  // landing bits (ShardStep::lands) are opt-in: measured on 8 GPUs at 34 qubits they did not help
  L->land = s->xmode == QB_X_PUSH && getenv("QCC_B200_LAND") && atoi(getenv("QCC_B200_LAND")) ? 1 : 0;
  L->pass_targets = std::max(1, s->tile_bits - QB_TILE_LOW);
}


int flush_impl(qb_state *s);

// QCC_B200_TRACE_FLUSH=1: one line per flush on stderr (gates, host milliseconds spent lowering + planning +
// launching) -- where the host time of a long gate stream goes.
int flush(qb_state *s) {
  static const bool trace = getenv("QCC_B200_TRACE_FLUSH") != nullptr;
  if (!trace || s->queue.empty()) return flush_impl(s);
  const size_t ng = s->queue.size();
  const uint64_t p0 = s->cnt.passes;
  const double t0 = host_now_ms();
  const int rc = flush_impl(s);
  fprintf(stderr, "qcc_b200 flush: %zu gates -> %llu passes, host %.2f ms\n", ng,
          (unsigned long long)(s->cnt.passes - p0), host_now_ms() - t0);
  return rc;
}

int flush_impl(qb_state *s) {
  if (s->queue.empty()) return QB_OK;
  CU(cudaSetDevice(s->device));
  std::vector<QbGate> q;
  q.swap(s->queue);
  static const bool no_ccu = getenv("QCC_B200_NO_CCU_FUSE") != nullptr;
  if (!no_ccu && s->fusion) qb::fuse_ccu_runs(q.data(), int64_t(q.size()));
  if (s->nranks == 1) return run_local(s, q);
  qb::ShardLayout L;
  make_layout(s, &L);
  std::vector<qb::ShardStep> steps;
  qb::lower_for_rank(&L, q.data(), int64_t(q.size()), &steps);
  s->perm = L.perm;
  s->flip = L.flip;
  return run_steps(s, steps);
}

// logical index -> (owner rank, local index) under the current bit permutation
void locate(const qb_state *s, uint64_t logical, int *rank, uint64_t *local) {
  uint64_t phys = 0;
  for (int b = 0; b < s->nq; ++b)
    if (logical >> b & 1) phys |= uint64_t(1) << s->perm[size_t(b)];
  *rank = int((phys >> s->n) ^ s->flip);   // a relabelled rank bit holds the negated qubit
  *local = phys & (s->len - 1);
}

uint64_t logical_of(const qb_state *s, int rank, uint64_t local) {
  const uint64_t phys = (uint64_t(uint32_t(rank) ^ s->flip) << s->n) | local;
  uint64_t logical = 0;
  for (int b = 0; b < s->nq; ++b)
    if (phys >> s->perm[size_t(b)] & 1) logical |= uint64_t(1) << b;
  return logical;
}

// Sum a device double over all ranks (no-op when not sharded).
int allreduce_scalar(qb_state *s, double *dptr) {
  if (s->nranks == 1) return QB_OK;
  std::string why;
  const qb::NcclApi *nc = qb::nccl_api(&why);
  if (!nc) return fail(QB_ERR_COMM, "%s", why.c_str());
  NC(nc, nc->AllReduce(dptr, dptr, 1, ncclDouble, ncclSum, s->comm, s->stream));
  return QB_OK;
}

int enqueue(qb_state *s, uint64_t ctl_mask, int target, const double m[8]) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  if (!m) return fail(QB_ERR_ARG, "null gate matrix");
  if (target < 0 || target >= s->nq) return fail(QB_ERR_ARG, "target bit %d out of range [0,%d)", target, s->nq);
  if (s->nq < 64 && (ctl_mask >> s->nq)) return fail(QB_ERR_ARG, "control mask has bits >= %d", s->nq);
  if (ctl_mask >> target & 1) return fail(QB_ERR_ARG, "control and target coincide (bit %d)", target);
  if (__builtin_popcountll(ctl_mask) > 3) return fail(QB_ERR_UNSUPPORTED, "more than 3 controls");
  QbGate g;
  g.ctl_mask = ctl_mask;
  g.target = target;
  g.kind = classify(m);
  memcpy(g.m, m, sizeof g.m);
  if (!s->fusion && s->nranks == 1) {
    CU(cudaSetDevice(s->device));
    return run_single(s, g);
  }
  s->queue.push_back(g);
  if (!s->fusion) return flush(s);
  if (s->queue.size() >= 16384) return flush(s);
  return QB_OK;
}

// xgates.cc:45-67 predicate with python numbering -> (ctl_mask | no-op) in index bits.
// Returns 1 if the gate acts on nothing, 0 if *mask is valid, <0 on error.
int xg_control_mask(int nbits, int ctl, int tgt, uint64_t *mask) {
  int t = nbits - tgt - 1;
  long long c = (long long)nbits - ctl - 1;
  if (c < 0) return fail(QB_ERR_ARG, "control index %d >= nbits %d (negative shift in xgates)", ctl, nbits);
  if (c < nbits) {
    if (c == t) return 1;  // pair base has the target bit clear: predicate never true
    *mask = uint64_t(1) << c;
    return 0;
  }
  // bit c of (g << nbits) + i == bit (c - nbits) of the pair-group base g, whose low
  // t+1 bits are zero and which is < 2^nbits.
  long long cc = c - nbits;
  if (cc <= t || cc >= nbits) return 1;
  *mask = uint64_t(1) << cc;
  return 0;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int qb_abi_version(void) { return QB_ABI_VERSION; }

const char *qb_last_error(void) { return g_err.c_str(); }

int qb_device_count(int *count) {
  if (!count) return fail(QB_ERR_ARG, "null pointer");
  *count = 0;
  CU(cudaGetDeviceCount(count));
  return QB_OK;
}

int qb_device_info(int device, char *name, size_t name_len, int *sm_count, size_t *total_mem,
                   int *cc_major, int *cc_minor) {
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, device));
  if (name && name_len) snprintf(name, name_len, "%s", p.name);
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (total_mem) *total_mem = p.totalGlobalMem;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return QB_OK;
}

int qb_comm_get_unique_id(void *id128) {
  if (!id128) return fail(QB_ERR_ARG, "null pointer");
  std::string why;
  const qb::NcclApi *nc = qb::nccl_api(&why);
  if (!nc) return fail(QB_ERR_COMM, "%s", why.c_str());
  static_assert(sizeof(ncclUniqueId) == 128, "unique id size");
  ncclUniqueId id;
  NC(nc, nc->GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return QB_OK;
}

static int create_impl(int nqubits, uint64_t init_label, int device, int rank, int nranks, const void *id128,
                       qb_state **out);

int qb_state_create(int nqubits, uint64_t init_label, int device, qb_state **out) {
  return create_impl(nqubits, init_label, device, 0, 1, nullptr, out);
}

int qb_state_create_sharded(int nqubits, uint64_t init_label, int device, int rank, int nranks,
                            const void *id128, qb_state **out) {
  if (nranks < 1 || (nranks & (nranks - 1))) return fail(QB_ERR_ARG, "nranks %d is not a power of two", nranks);
  if (rank < 0 || rank >= nranks) return fail(QB_ERR_ARG, "rank %d out of range", rank);
  if (nranks > 1 && !id128) return fail(QB_ERR_ARG, "sharded state needs the communicator id");
  return create_impl(nqubits, init_label, device, rank, nranks, id128, out);
}

static int create_impl(int nqubits, uint64_t init_label, int device, int rank, int nranks, const void *id128,
                       qb_state **out) {
  if (!out) return fail(QB_ERR_ARG, "null out pointer");
  *out = nullptr;
  if (nqubits < 1 || nqubits > 40) return fail(QB_ERR_ARG, "nqubits %d out of range [1,40]", nqubits);
  int pbits = 0;
  while ((1 << pbits) < nranks) ++pbits;
  if (nqubits - pbits < 4 && nranks > 1)
    return fail(QB_ERR_ARG, "%d qubits over %d ranks leaves fewer than 4 local bits", nqubits, nranks);
  if (nqubits < 64 && (init_label >> nqubits)) return fail(QB_ERR_ARG, "init label does not fit %d qubits", nqubits);
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(QB_ERR_CUDA, "no CUDA device visible (this engine has no CPU path)");
  if (device < 0) device = 0;
  if (device >= ndev) return fail(QB_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(QB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  qb_state *s = new qb_state;
  s->nq = nqubits;
  s->n = nqubits - pbits;
  s->len = uint64_t(1) << s->n;
  s->device = device;
  s->rank = rank;
  s->nranks = nranks;
  s->p = pbits;
  s->perm.resize(size_t(nqubits));
  for (int b = 0; b < nqubits; ++b) s->perm[size_t(b)] = b;
  size_t freeb = 0, totalb = 0;
  cudaMemGetInfo(&freeb, &totalb);
  const size_t need = size_t(s->len) * sizeof(double2);
  const size_t margin = size_t(64) << 20;
  if (nranks == 1 && need + margin > freeb) {
    delete s;
    return fail(QB_ERR_NOMEM, "state needs %zu MiB, device has %zu MiB free", need >> 20, freeb >> 20);
  }
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_scalar, sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&s->d_sync, sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&s->d_counter, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&s->d_blk_prob, sizeof(double) * qb::argmax_blocks());
  if (e == cudaSuccess) e = cudaMalloc(&s->d_blk_idx, sizeof(uint64_t) * qb::argmax_blocks());
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->plan_free, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(s->plan_free, s->stream);
  if (e == cudaSuccess) e = cudaEventCreate(&s->t0);
  if (e == cudaSuccess) e = cudaEventCreate(&s->t1);
  if (e == cudaSuccess) e = qb::fused_configure(device);
  if (e == cudaSuccess && nranks == 1) e = cudaMalloc(&s->base, need);
  if (e != cudaSuccess) {
    int rc = fail(e == cudaErrorMemoryAllocation ? QB_ERR_NOMEM : QB_ERR_CUDA, "state allocation failed: %s",
                  cudaGetErrorString(e));
    qb_state_destroy(s);
    return rc;
  }
  if (nranks > 1) {
    std::string why;
    const qb::NcclApi *nc = qb::nccl_api(&why);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = nc ? nc->CommInitRank(&s->comm, nranks, id, rank) : ncclSystemError;
    if (r != ncclSuccess) {
      int rc = fail(QB_ERR_COMM, "ncclCommInitRank: %s", nc ? nc->GetErrorString(r) : why.c_str());
      qb_state_destroy(s);
      return rc;
    }
    // Exchange mode.  Every decision below is collective (min over ranks), so all ranks end up in the same
    // mode: push needs room for a second shard on every GPU and the peer mappings; swap only the mappings.
    int mode = nranks <= qb::kPushMaxRanks ? QB_X_PUSH : QB_X_SWAP;
    if (const char *env = getenv("QCC_B200_EXCHANGE")) {
      if (!strcmp(env, "nccl")) mode = QB_X_NCCL;
      else if (!strcmp(env, "swap")) mode = QB_X_SWAP;
      else if (!strcmp(env, "push") && nranks <= qb::kPushMaxRanks) mode = QB_X_PUSH;
    }
    int rc = QB_OK;
    if (mode == QB_X_PUSH) {
      double ok = 2 * need + margin <= freeb ? 1.0 : 0.0;
      if (ok != 0.0 && cudaMalloc(&s->base, 2 * need + kBarrierBytes) != cudaSuccess) {
        cudaGetLastError();
        s->base = nullptr;
        ok = 0.0;
      }
      rc = agree_min(s, &ok);
      if (rc == QB_OK && ok == 0.0) {
        if (s->base) cudaFree(s->base);
        s->base = nullptr;
        mode = QB_X_SWAP;
      } else {
        s->nbuf = 2;
      }
    }
    if (rc == QB_OK && !s->base) {
      double ok = need + margin <= freeb ? 1.0 : 0.0;
      if (ok != 0.0 && cudaMalloc(&s->base, need + kBarrierBytes) != cudaSuccess) {
        cudaGetLastError();
        s->base = nullptr;
        ok = 0.0;
      }
      rc = agree_min(s, &ok);
      if (rc == QB_OK && ok == 0.0)
        rc = fail(QB_ERR_NOMEM, "shard needs %zu MiB on every rank (this device has %zu MiB free)", need >> 20, freeb >> 20);
    }
    if (rc == QB_OK && mode != QB_X_NCCL) {
      // the barrier counters behind the shard(s) start at zero on every rank before anybody can touch them
      if (cudaMemsetAsync(s->base + uint64_t(s->nbuf) * s->len, 0, kBarrierBytes, s->stream) != cudaSuccess)
        rc = fail(QB_ERR_CUDA, "barrier counters");
      bool mapped = false;
      if (rc == QB_OK) rc = map_peers(s, &mapped);   // its agreement rounds order the memset before any peer access
      if (!mapped) mode = QB_X_NCCL;
      // The counters-in-peer-memory barrier (k_peer_barrier) is opt-in: 23 us instead of 85 us per event on 2
      // GPUs, but runs with it were not reproducibly as fast as runs with the NCCL all-reduce.
      else s->peer_barrier = nranks <= qb::kPushMaxRanks && getenv("QCC_B200_PEER_BARRIER") != nullptr;
    }
    if (rc != QB_OK) {
      qb_state_destroy(s);
      return rc;
    }
    s->xmode = mode;
    // peer-memory exchanges do not care how many contiguous runs the exchanged part is made of
    // Shards of 8 GiB and more: victims stay above bit 8.  The per-launch trace of QFT-34 on 8 GPUs
    // (profiles/r02_qft34_8gpu_trace_run2.err) has the push passes at 43 ms with victims 27-29,
    // 53-80 ms with 9-11 and 174-199 ms with 5-7; the lowering model gives the same number of events and passes
    // either way.
    const int min_victim = s->n >= 29 ? qb::kPeerSwapMinVictimLarge : qb::kPeerSwapMinVictim;
    if (mode != QB_X_NCCL) s->victim_window = std::max(qb::kVictimWindow, s->n - min_victim);
    if (mode == QB_X_PUSH) s->victim_window = std::min(s->victim_window, s->n - QB_TILE_LOW);  // bits 0..2 never move
  }
  s->psi = s->base;
  int rc = qb_set_basis(s, init_label);
  if (rc != QB_OK) {
    qb_state_destroy(s);
    return rc;
  }
  *out = s;
  return QB_OK;
}

int qb_state_destroy(qb_state *s) {
  if (!s) return QB_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (auto &r : s->prof_pending) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto e : s->event_pool) cudaEventDestroy(e);
  // peer mappings: every rank unmaps the others' vectors BEFORE anybody frees its own (collective)
  bool had_peers = false;
  for (int r = 0; r < int(s->peer_base.size()); ++r)
    if (r != s->rank && s->peer_base[size_t(r)]) {
      cudaIpcCloseMemHandle(s->peer_base[size_t(r)]);
      had_peers = true;
    }
  s->peer_base.clear();
  if (had_peers && s->comm && s->d_sync) {
    const qb::NcclApi *nc = qb::nccl_api(nullptr);
    if (nc && nc->AllReduce(s->d_sync, s->d_sync, 1, ncclDouble, ncclMin, s->comm, s->stream) == ncclSuccess)
      cudaStreamSynchronize(s->stream);
  }
  if (s->comm) {
    const qb::NcclApi *nc = qb::nccl_api(nullptr);
    if (nc) nc->CommDestroy(s->comm);
  }
  if (s->xbuf) cudaFree(s->xbuf);
  if (s->d_sync) cudaFree(s->d_sync);
  for (auto e : s->xevents) cudaEventDestroy(e);
  if (s->xstream) cudaStreamDestroy(s->xstream);
  if (s->base) cudaFree(s->base);
  if (s->d_scalar) cudaFree(s->d_scalar);
  if (s->d_counter) cudaFree(s->d_counter);
  if (s->d_blk_prob) cudaFree(s->d_blk_prob);
  if (s->d_blk_idx) cudaFree(s->d_blk_idx);
  if (s->d_list_labels) cudaFree(s->d_list_labels);
  if (s->d_list_amps) cudaFree(s->d_list_amps);
  if (s->d_plan) cudaFree(s->d_plan);
  if (s->h_plan) cudaFreeHost(s->h_plan);
  if (s->plan_free) cudaEventDestroy(s->plan_free);
  if (s->t0) cudaEventDestroy(s->t0);
  if (s->t1) cudaEventDestroy(s->t1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return QB_OK;
}

int qb_state_nqubits(qb_state *s, int *nqubits) {
  if (!s || !nqubits) return fail(QB_ERR_ARG, "null pointer");
  *nqubits = s->nq;
  return QB_OK;
}

int qb_set_basis(qb_state *s, uint64_t label) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  if (s->nq < 64 && (label >> s->nq)) return fail(QB_ERR_ARG, "label out of range");
  CU(cudaSetDevice(s->device));
  s->queue.clear();
  for (int b = 0; b < s->nq; ++b) s->perm[size_t(b)] = b;
  s->flip = 0;
  CU(cudaMemsetAsync(s->psi, 0, size_t(s->len) * sizeof(double2), s->stream));
  const double2 one = make_double2(1.0, 0.0);
  if (int(label >> s->n) == s->rank)
    CU(cudaMemcpyAsync(s->psi + (label & (s->len - 1)), &one, sizeof one, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));  // `one` is a stack temporary
  s->cnt.kernel_launches += 1;
  return QB_OK;
}

int qb_fill_random(qb_state *s, uint64_t seed) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  CU(cudaSetDevice(s->device));
  s->queue.clear();
  for (int b = 0; b < s->nq; ++b) s->perm[size_t(b)] = b;
  s->flip = 0;
  CU(qb::launch_fill_random(s->psi, s->len, uint64_t(s->rank) << s->n, seed, s->stream));
  double n2 = 0.0;
  CU(qb::launch_norm2(s->psi, s->len, s->d_scalar, s->stream));
  QB(allreduce_scalar(s, s->d_scalar));
  CU(cudaMemcpyAsync(&n2, s->d_scalar, sizeof n2, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  CU(qb::launch_scale(s->psi, s->len, 1.0 / std::sqrt(n2), s->stream));
  s->cnt.kernel_launches += 3;
  return QB_OK;
}

// Sharded states: the copies address this rank's slice of the CANONICAL vector (identity bit layout), so a
// layout left behind by exchange events or rank relabels is undone first (collective, like every call on
// a sharded state).  Overwriting the whole shard needs no data movement: the layout is simply reset.
static bool layout_is_canonical(const qb_state *s) {
  if (s->flip) return false;
  for (int b = 0; b < s->nq; ++b)
    if (s->perm[size_t(b)] != b) return false;
  return true;
}

int qb_copy_in(qb_state *s, uint64_t first, uint64_t count, const double *host) {
  if (!s || !host) return fail(QB_ERR_ARG, "null pointer");
  if (first > s->len || count > s->len - first) return fail(QB_ERR_ARG, "range out of bounds");
  QB(flush(s));
  if (s->nranks > 1 && !layout_is_canonical(s)) {
    if (first == 0 && count == s->len) {
      for (int b = 0; b < s->nq; ++b) s->perm[size_t(b)] = b;
      s->flip = 0;
    } else {
      QB(qb_canonicalize(s));
    }
  }
  CU(cudaMemcpyAsync(s->psi + first, host, size_t(count) * sizeof(double2), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return QB_OK;
}

int qb_copy_out(qb_state *s, uint64_t first, uint64_t count, double *host) {
  if (!s || !host) return fail(QB_ERR_ARG, "null pointer");
  if (first > s->len || count > s->len - first) return fail(QB_ERR_ARG, "range out of bounds");
  QB(flush(s));
  if (s->nranks > 1 && !layout_is_canonical(s)) QB(qb_canonicalize(s));
  CU(cudaMemcpyAsync(host, s->psi + first, size_t(count) * sizeof(double2), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return QB_OK;
}

// ---- gates, index-bit numbering ---------------------------------------------------
int qb_apply1(qb_state *s, int target, const double m[8]) { return enqueue(s, 0, target, m); }

int qb_applyc(qb_state *s, int control, int target, const double m[8]) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  if (control < 0 || control >= s->nq) return fail(QB_ERR_ARG, "control bit %d out of range", control);
  return enqueue(s, uint64_t(1) << control, target, m);
}

int qb_applycc(qb_state *s, int c0, int c1, int target, const double m[8]) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  if (c0 < 0 || c0 >= s->nq || c1 < 0 || c1 >= s->nq) return fail(QB_ERR_ARG, "control bit out of range");
  if (c0 == c1) return fail(QB_ERR_ARG, "identical controls");
  return enqueue(s, (uint64_t(1) << c0) | (uint64_t(1) << c1), target, m);
}

int qb_apply_gates(qb_state *s, const qb_gate *gates, int64_t ngates) {
  if (!s || (!gates && ngates)) return fail(QB_ERR_ARG, "null pointer");
  for (int64_t k = 0; k < ngates; ++k) QB(enqueue(s, gates[k].ctl_mask, gates[k].target, gates[k].m));
  return QB_OK;
}

// ---- gates, python numbering --------------------------------------------------------
int qb_xg_apply1(qb_state *s, int tgt, const double m[8]) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  return enqueue(s, 0, s->nq - tgt - 1, m);  // out-of-range tgt is rejected by enqueue
}

int qb_xg_applyc(qb_state *s, int ctl, int tgt, const double m[8]) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  int t = s->nq - tgt - 1;
  if (t < 0 || t >= s->nq) return fail(QB_ERR_ARG, "target qubit %d out of range", tgt);
  uint64_t mask = 0;
  int r = xg_control_mask(s->nq, ctl, tgt, &mask);
  if (r < 0) return r;
  if (r == 1) {  // acts on nothing; still counts as an applied gate
    s->cnt.gates_applied += 1;
    return QB_OK;
  }
  return enqueue(s, mask, t, m);
}

int qb_xg_apply_gates(qb_state *s, const qb_xg_gate *gates, int64_t ngates) {
  if (!s || (!gates && ngates)) return fail(QB_ERR_ARG, "null pointer");
  for (int64_t k = 0; k < ngates; ++k) {
    const qb_xg_gate &g = gates[k];
    if (g.kind == 1) QB(qb_xg_apply1(s, g.tgt, g.m));
    else if (g.kind == 2) QB(qb_xg_applyc(s, g.ctl, g.tgt, g.m));
    else return fail(QB_ERR_ARG, "gate %lld: kind %d", (long long)k, g.kind);
  }
  return QB_OK;
}

// ---- queue control --------------------------------------------------------------------
int qb_set_fusion(qb_state *s, int fusion) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  QB(flush(s));
  s->fusion = fusion != 0;
  return QB_OK;
}

int qb_flush(qb_state *s) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  return flush(s);
}

int qb_sync(qb_state *s) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  QB(flush(s));
  CU(cudaStreamSynchronize(s->stream));
  return QB_OK;
}

// ---- readouts ---------------------------------------------------------------------------
int qb_get_amplitude(qb_state *s, uint64_t index, double out[2]) {
  if (!s || !out) return fail(QB_ERR_ARG, "null pointer");
  if (s->nq < 64 && (index >> s->nq)) return fail(QB_ERR_ARG, "index out of range");
  if (s->nranks == 1) return qb_copy_out(s, index, 1, out);
  // sharded: collective -- the owner reads, everybody receives
  QB(flush(s));
  int owner = 0;
  uint64_t local = 0;
  locate(s, index, &owner, &local);
  std::string why;
  const qb::NcclApi *nc = qb::nccl_api(&why);
  if (!nc) return fail(QB_ERR_COMM, "%s", why.c_str());
  double *two = s->d_blk_prob;  // scratch, >= 2 doubles
  if (owner == s->rank) CU(cudaMemcpyAsync(two, s->psi + local, sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
  NC(nc, nc->Broadcast(two, two, 2, ncclDouble, owner, s->comm, s->stream));
  CU(cudaMemcpyAsync(out, two, sizeof(double2), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return QB_OK;
}

int qb_norm2(qb_state *s, double *out) {
  if (!s || !out) return fail(QB_ERR_ARG, "null pointer");
  QB(flush(s));
  {
    ProfScope ps(s, QB_KCLASS_AUX, double(s->len) * 16.0);
    CU(qb::launch_norm2(s->psi, s->len, s->d_scalar, s->stream));
  }
  s->cnt.kernel_launches += 1;
  QB(allreduce_scalar(s, s->d_scalar));
  CU(cudaMemcpyAsync(out, s->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return QB_OK;
}

int qb_prob_bit_value(qb_state *s, int bit, int value, double *p) {
  if (!s || !p) return fail(QB_ERR_ARG, "null pointer");
  if (bit < 0 || bit >= s->nq) return fail(QB_ERR_ARG, "bit out of range");
  if (value != 0 && value != 1) return fail(QB_ERR_ARG, "bit value must be 0 or 1");
  QB(flush(s));
  {
    ProfScope ps(s, QB_KCLASS_AUX, double(s->len) * 8.0);
    const int pb = s->perm[size_t(bit)];
    if (pb < s->n) {
      CU(qb::launch_prob_mask(s->psi, s->len, uint64_t(1) << pb, uint64_t(value) << pb, s->d_scalar, s->stream));
    } else if (int(((uint32_t(s->rank) ^ s->flip) >> (pb - s->n)) & 1u) == value) {
      CU(qb::launch_norm2(s->psi, s->len, s->d_scalar, s->stream));  // the bit has this value on the whole shard
    } else {
      CU(cudaMemsetAsync(s->d_scalar, 0, sizeof(double), s->stream));
    }
  }
  s->cnt.kernel_launches += 1;
  QB(allreduce_scalar(s, s->d_scalar));
  CU(cudaMemcpyAsync(p, s->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return QB_OK;
}

int qb_prob_bit(qb_state *s, int bit, double *p_one) { return qb_prob_bit_value(s, bit, 1, p_one); }

int qb_argmax(qb_state *s, uint64_t *index, double *prob) {
  if (!s || !index || !prob) return fail(QB_ERR_ARG, "null pointer");
  QB(flush(s));
  int nb = qb::argmax_blocks();
  {
    ProfScope ps(s, QB_KCLASS_AUX, double(s->len) * 16.0);
    qb::LogicalMap lm;
    if (s->nranks > 1) {
      lm.identity = 0;
      lm.n = s->n;
      lm.hi = logical_of(s, s->rank, 0);
      for (int b = 0; b < s->nq; ++b)
        if (s->perm[size_t(b)] < s->n) lm.lpos[s->perm[size_t(b)]] = b;
    }
    CU(qb::launch_argmax(s->psi, s->len, s->d_blk_prob, s->d_blk_idx, lm, s->stream));
  }
  s->cnt.kernel_launches += 1;
  std::vector<double> hp(nb);
  std::vector<uint64_t> hi(nb);
  CU(cudaMemcpyAsync(hp.data(), s->d_blk_prob, sizeof(double) * nb, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(hi.data(), s->d_blk_idx, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  double best = -1.0;
  uint64_t bi = 0;
  for (int b = 0; b < nb; ++b)
    if (hp[b] > best || (hp[b] == best && hi[b] < bi)) {
      best = hp[b];
      bi = hi[b];
    }
  if (s->nranks > 1) {
    // the kernel already reports logical indices: pick the global winner (ties: lowest logical index)
    std::string why;
    const qb::NcclApi *nc = qb::nccl_api(&why);
    if (!nc) return fail(QB_ERR_COMM, "%s", why.c_str());
    double mine[2];
    mine[0] = best;
    memcpy(&mine[1], &bi, sizeof bi);
    double *dsend = s->d_blk_prob;                 // 2 doubles
    double *drecv = s->d_blk_prob + 2;             // 2 * nranks doubles (argmax_blocks() >> 2 * 8 + 2)
    CU(cudaMemcpyAsync(dsend, mine, sizeof mine, cudaMemcpyHostToDevice, s->stream));
    NC(nc, nc->AllGather(dsend, drecv, 2, ncclDouble, s->comm, s->stream));
    std::vector<double> all(size_t(2 * s->nranks));
    CU(cudaMemcpyAsync(all.data(), drecv, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    best = -1.0;
    for (int r = 0; r < s->nranks; ++r) {
      uint64_t idx;
      memcpy(&idx, &all[size_t(2 * r + 1)], sizeof idx);
      if (all[size_t(2 * r)] > best || (all[size_t(2 * r)] == best && idx < bi)) {
        best = all[size_t(2 * r)];
        bi = idx;
      }
    }
  }
  *index = bi;
  *prob = best;
  return QB_OK;
}

int qb_list_above(qb_state *s, double threshold, uint64_t cap, uint64_t *labels, double *amps,
                  uint64_t *count) {
  if (!s || !count) return fail(QB_ERR_ARG, "null pointer");
  if (cap && (!labels || !amps)) return fail(QB_ERR_ARG, "null output buffers with cap > 0");
  QB(flush(s));
  // sharded: every rank lists its shard, then the lists are all-gathered -- the buffers hold one slot of
  // `cap` entries per rank
  const uint64_t slots = s->nranks > 1 ? uint64_t(s->nranks) + 1 : 1;
  if (cap * slots > s->list_cap) {
    if (s->d_list_labels) cudaFree(s->d_list_labels);
    if (s->d_list_amps) cudaFree(s->d_list_amps);
    s->d_list_labels = nullptr;
    s->d_list_amps = nullptr;
    s->list_cap = 0;
    cudaError_t e = cudaMalloc(&s->d_list_labels, cap * slots * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_list_amps, cap * slots * sizeof(double2));
    if (e != cudaSuccess) return fail(QB_ERR_NOMEM, "list buffers: %s", cudaGetErrorString(e));
    s->list_cap = cap * slots;
  }
  uint64_t *d_labels = s->d_list_labels;
  double2 *d_amps = s->d_list_amps;
  {
    ProfScope ps(s, QB_KCLASS_AUX, double(s->len) * 16.0);
    CU(qb::launch_list_above(s->psi, s->len, threshold, cap, s->d_counter, d_labels, d_amps, s->stream));
  }
  s->cnt.kernel_launches += 1;
  unsigned long long found = 0;
  CU(cudaMemcpyAsync(&found, s->d_counter, sizeof found, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  uint64_t got = std::min<uint64_t>(found, cap);
  std::vector<uint64_t> hl(got);
  std::vector<double2> ha(got);
  if (got) {
    CU(cudaMemcpyAsync(hl.data(), d_labels, got * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(ha.data(), d_amps, got * sizeof(double2), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
  }
  uint64_t total = found;
  if (s->nranks > 1) {
    // labels under their LOGICAL names, then everybody's list to everybody (collective, like every readout)
    for (uint64_t i = 0; i < got; ++i) hl[i] = logical_of(s, s->rank, hl[i]);
    std::string why;
    const qb::NcclApi *nc = qb::nccl_api(&why);
    if (!nc) return fail(QB_ERR_COMM, "%s", why.c_str());
    const int R = s->nranks;
    uint64_t *d_cnt = s->d_blk_idx;   // scratch: 1 + R counters
    CU(cudaMemcpyAsync(d_cnt, &found, sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
    NC(nc, nc->AllGather(d_cnt, d_cnt + 1, 1, ncclUint64, s->comm, s->stream));
    std::vector<uint64_t> cnts(static_cast<size_t>(R));
    CU(cudaMemcpyAsync(cnts.data(), d_cnt + 1, sizeof(uint64_t) * size_t(R), cudaMemcpyDeviceToHost, s->stream));
    if (cap) {
      if (got) CU(cudaMemcpyAsync(d_labels, hl.data(), got * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
      NC(nc, nc->AllGather(d_labels, d_labels + cap, size_t(cap), ncclUint64, s->comm, s->stream));
      NC(nc, nc->AllGather(d_amps, d_amps + cap, size_t(cap) * 2, ncclDouble, s->comm, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    total = 0;
    for (int r = 0; r < R; ++r) total += cnts[size_t(r)];
    hl.clear();
    ha.clear();
    for (int r = 0; r < R && cap; ++r) {
      const uint64_t g = std::min<uint64_t>(cnts[size_t(r)], cap);
      if (!g) continue;
      const size_t at = hl.size();
      hl.resize(at + g);
      ha.resize(at + g);
      CU(cudaMemcpy(hl.data() + at, d_labels + cap * uint64_t(r + 1), g * sizeof(uint64_t), cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(ha.data() + at, d_amps + cap * uint64_t(r + 1), g * sizeof(double2), cudaMemcpyDeviceToHost));
    }
    got = hl.size();
  }
  // the kernel's slots are in arrival order; present them by ascending label
  std::vector<uint64_t> order(got);
  for (uint64_t i = 0; i < got; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return hl[a] < hl[b]; });
  const uint64_t nout = std::min<uint64_t>(got, cap);
  for (uint64_t i = 0; i < nout; ++i) {
    labels[i] = hl[order[i]];
    amps[2 * i] = ha[order[i]].x;
    amps[2 * i + 1] = ha[order[i]].y;
  }
  *count = total;
  return QB_OK;
}

// ---- host-buffer entry points (the libxgates binding) -----------------------------------
namespace {
struct HostScratch {
  int device = -1;
  void *d = nullptr;
  size_t cap = 0;
  cudaStream_t st = nullptr;
};
thread_local HostScratch g_hs;

int host_scratch(int device, size_t bytes, HostScratch **out) {
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(QB_ERR_CUDA, "no CUDA device visible (this engine has no CPU path)");
  if (device < 0) device = 0;
  if (device >= ndev) return fail(QB_ERR_ARG, "device %d out of range", device);
  CU(cudaSetDevice(device));
  HostScratch &h = g_hs;
  if (h.device != device) {
    if (h.d) cudaFree(h.d);
    h.d = nullptr;
    h.cap = 0;
    if (h.st) cudaStreamDestroy(h.st);
    h.st = nullptr;
    h.device = device;
  }
  if (!h.st) CU(cudaStreamCreateWithFlags(&h.st, cudaStreamNonBlocking));
  if (bytes > h.cap) {
    if (h.d) cudaFree(h.d);
    h.d = nullptr;
    h.cap = 0;
    cudaError_t e = cudaMalloc(&h.d, bytes);
    if (e != cudaSuccess) return fail(QB_ERR_NOMEM, "host-path scratch of %zu MiB: %s", bytes >> 20, cudaGetErrorString(e));
    h.cap = bytes;
  }
  *out = &h;
  return QB_OK;
}

int host_gate(void *psi, const void *gate, int nbits, uint64_t mask, int t, int bit_width, int device) {
  bool dbl = bit_width == 128;
  size_t esz = dbl ? sizeof(double2) : sizeof(float2);
  size_t bytes = (size_t(1) << nbits) * esz;
  HostScratch *h = nullptr;
  QB(host_scratch(device, bytes, &h));
  QbGate g;
  g.ctl_mask = mask;
  g.target = t;
  if (dbl) memcpy(g.m, gate, sizeof g.m);
  else
    for (int k = 0; k < 8; ++k) g.m[k] = static_cast<const float *>(gate)[k];
  g.kind = classify(g.m);
  if (g.kind == QB_K_NOP) return QB_OK;
  CU(cudaMemcpyAsync(h->d, psi, bytes, cudaMemcpyHostToDevice, h->st));
  CU(qb::launch_gate(h->d, nbits, g, dbl, h->st));
  CU(cudaMemcpyAsync(psi, h->d, bytes, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return QB_OK;
}
}  // namespace

int qb_host_apply1(void *psi, const void *gate, int nbits, int tgt, int bit_width, int device) {
  if (!psi || !gate) return fail(QB_ERR_ARG, "null pointer");
  if (nbits < 1 || nbits > 40) return fail(QB_ERR_ARG, "nbits out of range");
  int t = nbits - tgt - 1;
  if (t < 0 || t >= nbits) return fail(QB_ERR_ARG, "Negative qubit index in apply1(): tgt %d, nbits %d", tgt, nbits);
  return host_gate(psi, gate, nbits, 0, t, bit_width, device);
}

int qb_host_applyc(void *psi, const void *gate, int nbits, int ctl, int tgt, int bit_width,
                   int device) {
  if (!psi || !gate) return fail(QB_ERR_ARG, "null pointer");
  if (nbits < 1 || nbits > 40) return fail(QB_ERR_ARG, "nbits out of range");
  int t = nbits - tgt - 1;
  if (t < 0 || t >= nbits) return fail(QB_ERR_ARG, "Negative qubit index in applyc(): tgt %d, nbits %d", tgt, nbits);
  uint64_t mask = 0;
  int r = xg_control_mask(nbits, ctl, tgt, &mask);
  if (r < 0) return r;
  if (r == 1) return QB_OK;
  return host_gate(psi, gate, nbits, mask, t, bit_width, device);
}

int qb_host_run(void *psi, int nbits, const qb_xg_gate *gates, int64_t ngates, int device) {
  if (!psi || (!gates && ngates)) return fail(QB_ERR_ARG, "null pointer");
  qb_state *s = nullptr;
  QB(qb_state_create(nbits, 0, device, &s));
  int rc = qb_copy_in(s, 0, s->len, static_cast<const double *>(psi));
  if (rc == QB_OK) rc = qb_xg_apply_gates(s, gates, ngates);
  if (rc == QB_OK) rc = qb_copy_out(s, 0, s->len, static_cast<double *>(psi));
  qb_state_destroy(s);
  return rc;
}

int qb_host_alloc(size_t bytes, void **out) {
  if (!out) return fail(QB_ERR_ARG, "null pointer");
  *out = nullptr;
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(QB_ERR_CUDA, "no CUDA device visible");
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) return fail(QB_ERR_NOMEM, "cudaHostAlloc(%zu MiB): %s", bytes >> 20, cudaGetErrorString(e));
  return QB_OK;
}

int qb_host_free(void *p) {
  if (p) CU(cudaFreeHost(p));
  return QB_OK;
}

// ---- measurement helpers ------------------------------------------------------------------
int qb_get_counters(qb_state *s, qb_counters *out) {
  if (!s || !out) return fail(QB_ERR_ARG, "null pointer");
  *out = s->cnt;
  return QB_OK;
}

int qb_profile_enable(qb_state *s, int enable) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  QB(qb_sync(s));
  QB(resolve_profile(s));
  s->profiling = enable != 0;
  return QB_OK;
}

int qb_profile_read(qb_state *s, qb_profile *out, int reset) {
  if (!s || !out) return fail(QB_ERR_ARG, "null pointer");
  QB(qb_sync(s));
  QB(resolve_profile(s));
  *out = s->prof;
  if (reset) s->prof = qb_profile{};
  return QB_OK;
}

int qb_timer_start(qb_state *s) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  QB(qb_sync(s));
  CU(cudaEventRecord(s->t0, s->stream));
  return QB_OK;
}

int qb_timer_stop(qb_state *s, double *ms) {
  if (!s || !ms) return fail(QB_ERR_ARG, "null pointer");
  QB(flush(s));
  CU(cudaEventRecord(s->t1, s->stream));
  CU(cudaEventSynchronize(s->t1));
  float f = 0.f;
  CU(cudaEventElapsedTime(&f, s->t0, s->t1));
  *ms = f;
  return QB_OK;
}

int qb_state_layout(qb_state *s, int *nlocal, int *rank, int *nranks, int *perm) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  QB(flush(s));
  if (nlocal) *nlocal = s->n;
  if (rank) *rank = s->rank;
  if (nranks) *nranks = s->nranks;
  if (perm)
    for (int b = 0; b < s->nq; ++b) perm[b] = s->perm[size_t(b)];
  return QB_OK;
}

int qb_state_exchange_mode(qb_state *s, int *mode) {
  if (!s || !mode) return fail(QB_ERR_ARG, "null pointer");
  *mode = s->nranks > 1 ? s->xmode : -1;
  return QB_OK;
}

int qb_canonicalize(qb_state *s) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  QB(flush(s));
  if (s->nranks == 1) return QB_OK;
  qb::ShardLayout L;
  make_layout(s, &L);
  std::vector<qb::ShardStep> steps;
  qb::canonicalize_steps(&L, &steps);
  s->perm = L.perm;
  s->flip = L.flip;
  const uint64_t before = s->cnt.gates_applied;
  int rc = run_steps(s, steps);
  s->cnt.gates_applied = before;  // layout moves are not gates of the caller's circuit
  return rc;
}

int qb_shard_lower_json(int nqubits, int nranks, int rank, const qb_gate *gates, int64_t ngates, int canonicalize,
                        char *buf, size_t cap, size_t *needed) {
  if ((!gates && ngates) || !needed) return fail(QB_ERR_ARG, "null pointer");
  if (nranks < 1 || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks) return fail(QB_ERR_ARG, "bad rank/nranks");
  int pbits = 0;
  while ((1 << pbits) < nranks) ++pbits;
  if (nqubits - pbits < 1 || nqubits > 40) return fail(QB_ERR_ARG, "nqubits out of range");
  std::vector<QbGate> q;
  for (int64_t k = 0; k < ngates; ++k) {
    const qb_gate &g = gates[k];
    if (g.target < 0 || g.target >= nqubits || (g.ctl_mask >> g.target & 1) || (g.ctl_mask >> nqubits))
      return fail(QB_ERR_ARG, "gate %lld: bad bits", (long long)k);
    QbGate x;
    x.ctl_mask = g.ctl_mask;
    x.target = g.target;
    x.kind = classify(g.m);
    memcpy(x.m, g.m, sizeof x.m);
    q.push_back(x);
  }
  qb::ShardLayout L;
  L.n = nqubits;
  L.nl = nqubits - pbits;
  L.p = pbits;
  L.rank = rank;
  L.perm.resize(size_t(nqubits));
  for (int b = 0; b < nqubits; ++b) L.perm[size_t(b)] = b;
  if (const char *w = getenv("QCC_B200_VICTIM_WINDOW")) L.window = std::max(1, atoi(w));  // tests: the peer-swap window
  if (const char *h = getenv("QCC_B200_HOIST")) L.hoist = atoi(h);                        // ... and its exchange hoisting
  if (const char *h = getenv("QCC_B200_PREFETCH")) L.prefetch = atoi(h);                  // ... and multi-bit events
  if (const char *h = getenv("QCC_B200_LAND")) L.land = atoi(h);                          // ... landing on the high bits
  std::vector<qb::ShardStep> steps;
  qb::lower_for_rank(&L, q.data(), ngates, &steps);
  if (canonicalize) qb::canonicalize_steps(&L, &steps);
  std::string js = qb::steps_to_json(L, steps);
  *needed = js.size() + 1;
  if (buf && cap >= js.size() + 1) memcpy(buf, js.c_str(), js.size() + 1);
  return QB_OK;
}

int qb_shard_plan_stats(int nqubits, int nranks, int rank, const qb_gate *gates, int64_t ngates, int tile_bits,
                        int window, int hoist, int prefetch, int64_t stats[8]) {
  if ((!gates && ngates) || !stats) return fail(QB_ERR_ARG, "null pointer");
  if (nranks < 1 || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks) return fail(QB_ERR_ARG, "bad rank/nranks");
  int pbits = 0;
  while ((1 << pbits) < nranks) ++pbits;
  if (nqubits - pbits < 4 || nqubits > 40) return fail(QB_ERR_ARG, "nqubits out of range");
  std::vector<QbGate> q;
  q.reserve(size_t(ngates));
  for (int64_t k = 0; k < ngates; ++k) {
    const qb_gate &g = gates[k];
    if (g.target < 0 || g.target >= nqubits || (g.ctl_mask >> g.target & 1) || (g.ctl_mask >> nqubits))
      return fail(QB_ERR_ARG, "gate %lld: bad bits", (long long)k);
    QbGate x;
    x.ctl_mask = g.ctl_mask;
    x.target = g.target;
    x.kind = classify(g.m);
    memcpy(x.m, g.m, sizeof x.m);
    q.push_back(x);
  }
  for (int k = 0; k < 8; ++k) stats[k] = 0;
  if (!getenv("QCC_B200_NO_CCU_FUSE")) stats[7] = qb::fuse_ccu_runs(q.data(), ngates);
  std::vector<qb::ShardStep> steps;
  if (nranks > 1) {
    qb::ShardLayout L;
    L.n = nqubits;
    L.nl = nqubits - pbits;
    L.p = pbits;
    L.rank = rank;
    L.perm.resize(size_t(nqubits));
    for (int b = 0; b < nqubits; ++b) L.perm[size_t(b)] = b;
    L.window = std::max(1, std::min(window, L.nl));
    L.hoist = hoist;
    L.prefetch = prefetch;
    L.land = prefetch && getenv("QCC_B200_LAND") && atoi(getenv("QCC_B200_LAND")) ? 1 : 0;
    L.pass_targets = std::max(1, tile_bits - QB_TILE_LOW);
    qb::lower_for_rank(&L, q.data(), ngates, &steps);
  } else {
    qb::ShardStep st;
    st.gates = q;
    steps.push_back(st);
  }
  bool tail_fused = false;
  for (size_t si = 0; si < steps.size(); ++si) {
    const qb::ShardStep &st = steps[si];
    if (st.kind == 1) {
      stats[0] += 1;
      stats[1] += int64_t(st.rank_bits.size());
      stats[4] += tail_fused ? 1 : 0;
      tail_fused = false;
      continue;
    }
    tail_fused = false;
    if (st.gates.empty()) continue;
    qb::Plan plan;
    qb::plan_gates(nqubits - pbits, st.gates.data(), int64_t(st.gates.size()), tile_bits, &plan,
                   prefetch && si + 1 < steps.size() && steps[si + 1].kind == 1);
    for (const qb::PlannedPass &pp : plan.passes) {
      stats[2] += 1;
      if (pp.single_gate < 0) {
        stats[3] += 1;
        stats[5] += int64_t(pp.rounds.size());
        stats[6] += int64_t(pp.ops.size());
      }
    }
    tail_fused = !plan.passes.empty() && plan.passes.back().single_gate < 0;
  }
  return QB_OK;
}

int qb_plan_check(int nqubits, const qb_gate *gates, int64_t ngates, int tile_bits, int64_t *passes) {
  if ((!gates && ngates) || !passes) return fail(QB_ERR_ARG, "null pointer");
  if (nqubits < 4 || nqubits > 40) return fail(QB_ERR_ARG, "nqubits out of range [4,40]");
  std::vector<QbGate> q(static_cast<size_t>(ngates));
  for (int64_t k = 0; k < ngates; ++k) {
    const qb_gate &g = gates[k];
    if (g.target < 0 || g.target >= nqubits || (g.ctl_mask >> g.target & 1) || (g.ctl_mask >> nqubits))
      return fail(QB_ERR_ARG, "gate %lld: bad bits", (long long)k);
    q[size_t(k)].ctl_mask = g.ctl_mask;
    q[size_t(k)].target = g.target;
    q[size_t(k)].kind = classify(g.m);
    memcpy(q[size_t(k)].m, g.m, sizeof g.m);
  }
  qb::fuse_ccu_runs(q.data(), ngates);
  qb::Plan plan;
  qb::plan_gates(nqubits, q.data(), ngates, tile_bits, &plan);
  plan.blob_bytes();
  *passes = int64_t(plan.passes.size());
  for (size_t k = 0; k < plan.passes.size(); ++k) {
    const qb::PlannedPass &pp = plan.passes[k];
    if (pp.single_gate >= 0) continue;
    qb::DevicePass dp;
    dp.desc = pp.desc;
    dp.ops = pp.ops.data();
    dp.rounds = pp.rounds.data();
    cudaError_t e = qb::check_fused_pass(nqubits, dp);
    if (e != cudaSuccess)
      return fail(QB_ERR_UNSUPPORTED, "pass %zu of %zu (%d ops, %d rounds, %d table entries) does not fit the kernel's "
                  "parameter block / shared memory", k, plan.passes.size(), pp.desc.nops, pp.desc.nrounds, pp.desc.ntable);
  }
  return QB_OK;
}

int qb_fuse_gates(qb_gate *gates, int64_t ngates, int64_t *fused) {
  if ((!gates && ngates) || !fused) return fail(QB_ERR_ARG, "null pointer");
  std::vector<QbGate> q(static_cast<size_t>(ngates));
  for (int64_t k = 0; k < ngates; ++k) {
    q[size_t(k)].ctl_mask = gates[k].ctl_mask;
    q[size_t(k)].target = gates[k].target;
    q[size_t(k)].kind = classify(gates[k].m);
    memcpy(q[size_t(k)].m, gates[k].m, sizeof gates[k].m);
  }
  *fused = qb::fuse_ccu_runs(q.data(), ngates);
  for (int64_t k = 0; k < ngates; ++k) {
    gates[k].ctl_mask = q[size_t(k)].ctl_mask;
    gates[k].target = q[size_t(k)].target;
    memcpy(gates[k].m, q[size_t(k)].m, sizeof gates[k].m);
  }
  return QB_OK;
}

int qb_shard_event_dest(int nlocal, int nranks, int rank, const int *rank_bits, const int *victims, const int *lands,
                        int npairs, const uint64_t *local, uint64_t *dest, int64_t count) {
  if (!rank_bits || !victims || (!local && count) || (!dest && count)) return fail(QB_ERR_ARG, "null pointer");
  if (nranks < 1 || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks) return fail(QB_ERR_ARG, "bad rank/nranks");
  int pbits = 0;
  while ((1 << pbits) < nranks) ++pbits;
  qb::PushMap m;
  QB(event_map(nlocal, pbits, rank, rank_bits, victims, lands, size_t(npairs), &m));
  for (int64_t k = 0; k < count; ++k) dest[k] = qb::push_apply(m, local[k]);
  return QB_OK;
}

int qb_set_tile_bits(qb_state *s, int tile_bits) {
  if (!s) return fail(QB_ERR_ARG, "null state");
  if (tile_bits < 4 || tile_bits > QB_MAX_TILE_BITS) return fail(QB_ERR_ARG, "tile_bits out of range [4,%d]", QB_MAX_TILE_BITS);
  QB(flush(s));
  s->tile_bits = tile_bits;
  return QB_OK;
}

int qb_plan_json(int nqubits, const qb_gate *gates, int64_t ngates, int tile_bits, char *buf,
                 size_t cap, size_t *needed) {
  if ((!gates && ngates) || !needed) return fail(QB_ERR_ARG, "null pointer");
  if (nqubits < 4 || nqubits > 40) return fail(QB_ERR_ARG, "nqubits out of range [4,40]");
  std::vector<QbGate> q;
  q.reserve(size_t(ngates));
  for (int64_t k = 0; k < ngates; ++k) {
    const qb_gate &g = gates[k];
    if (g.target < 0 || g.target >= nqubits || (g.ctl_mask >> g.target & 1) || (g.ctl_mask >> nqubits))
      return fail(QB_ERR_ARG, "gate %lld: bad bits", (long long)k);
    QbGate x;
    x.ctl_mask = g.ctl_mask;
    x.target = g.target;
    x.kind = classify(g.m);
    memcpy(x.m, g.m, sizeof x.m);
    q.push_back(x);
  }
  qb::Plan plan;
  qb::plan_gates(nqubits, q.data(), ngates, tile_bits, &plan);
  std::string js = plan.to_json();
  *needed = js.size() + 1;
  if (buf && cap >= js.size() + 1) memcpy(buf, js.c_str(), js.size() + 1);
  return QB_OK;
}

}  // extern "C"
