// kernels.cu -- single-gate sweeps and init/readout kernels for sm_100a.
//
// Every gate of the reference's hot path (xgates.cc:23-67, apply.cc:110-144,
// gates.cc:17-146) is a strided butterfly over the dense amplitude vector.  The
// kernels here are the UNFUSED form: one launch = one gate = one sweep over exactly
// the amplitudes the gate touches.  They are bandwidth-bound (0.25-0.5 flop/B in
// fp64), so the whole design is about the memory system:
//   * 128-bit accesses: one amplitude (complex128) per LDG.128/STG.128;
//   * consecutive lanes take consecutive "free" indices p, and the touched index is
//     p with a 0 (target) or 1 (control) bit inserted -- for any target >= 5 a warp
//     reads two fully coalesced 512 B runs; for targets 0..4 the two loads of a warp
//     together cover one contiguous 1 KiB window, so every DRAM sector that is
//     fetched is fully used;
//   * controlled and diagonal gates enumerate only the indices they change (half,
//     quarter, eighth of the vector) instead of predicating a full sweep, which is
//     what makes their algorithmic bytes (16N, 8N) the real traffic;
//   * 4 independent pairs per thread are loaded before any is used (8 LDG.128 in
//     flight per thread) and streaming cache hints (ld.global.cs / st.global.cs)
//     keep the 126 MB L2 from thrashing on data that is never re-read.
// Grids are sized from the pair count; at 30 qubits that is 2^19 CTAs of 256
// threads, i.e. thousands of waves over the 148 SMs, so no tail effect.
#include <cuda_runtime.h>

#include <algorithm>
#include <stdint.h>

#include "kernels.h"

namespace qb {

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// Bits to insert into a dense "free" counter to obtain a touched index.
struct InsertSpec {
  int n;
  int pos[4];        // ascending final positions
  uint64_t setmask;  // bits forced to 1 after insertion (controls / phase bits)
};

__device__ __forceinline__ uint64_t expand_index(uint64_t p, const InsertSpec &s) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < s.n) {
      uint64_t low = (uint64_t(1) << s.pos[k]) - 1;
      p = ((p & ~low) << 1) | (p & low);
    }
  }
  return p | s.setmask;
}

template <typename V>
struct Mat2 {
  V a, b, c, d;
};

template <typename V>
__device__ __forceinline__ V cmul(V x, V y) {
  V r;
  r.x = x.x * y.x - x.y * y.y;
  r.y = x.x * y.y + x.y * y.x;
  return r;
}
template <typename V>
__device__ __forceinline__ V cadd(V x, V y) {
  V r;
  r.x = x.x + y.x;
  r.y = x.y + y.y;
  return r;
}

// General 2x2 on pairs (i0, i0 | tmask); same arithmetic as xgates.cc:34-37.
template <typename V>
__global__ void __launch_bounds__(kThreads)
k_apply_u(V *__restrict__ psi, uint64_t npairs, InsertSpec ins, uint64_t tmask, Mat2<V> m) {
  uint64_t base = (uint64_t(blockIdx.x) * kUnroll) * kThreads + threadIdx.x;
  uint64_t idx[kUnroll];
  V x[kUnroll], y[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    uint64_t p = base + uint64_t(u) * kThreads;
    if (p < npairs) {
      idx[u] = expand_index(p, ins);
      x[u] = __ldcs(psi + idx[u]);
      y[u] = __ldcs(psi + (idx[u] | tmask));
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    uint64_t p = base + uint64_t(u) * kThreads;
    if (p < npairs) {
      V t1 = cadd(cmul(m.a, x[u]), cmul(m.b, y[u]));
      V t2 = cadd(cmul(m.c, x[u]), cmul(m.d, y[u]));
      __stcs(psi + idx[u], t1);
      __stcs(psi + (idx[u] | tmask), t2);
    }
  }
}

// diag(1, p): multiply every amplitude whose index has all `setmask` bits set.
template <typename V>
__global__ void __launch_bounds__(kThreads)
k_apply_phase(V *__restrict__ psi, uint64_t count, InsertSpec ins, V phase) {
  uint64_t base = (uint64_t(blockIdx.x) * kUnroll) * kThreads + threadIdx.x;
  uint64_t idx[kUnroll];
  V x[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    uint64_t p = base + uint64_t(u) * kThreads;
    if (p < count) {
      idx[u] = expand_index(p, ins);
      x[u] = __ldcs(psi + idx[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    uint64_t p = base + uint64_t(u) * kThreads;
    if (p < count) __stcs(psi + idx[u], cmul(phase, x[u]));
  }
}

template <typename V, typename T>
cudaError_t launch_gate_t(V *psi, int nbits, const QbGate &g, cudaStream_t st) {
  // positions to insert: target (as 0 for U/DIAG/PERM, as 1 for PHASE) and controls (as 1)
  InsertSpec ins{};
  uint64_t ones = g.ctl_mask;
  uint64_t all = g.ctl_mask | (uint64_t(1) << g.target);
  if (__builtin_popcountll(all) > 4) return cudaErrorInvalidValue;
  int n = 0;
  for (int b = 0; b < nbits; ++b)
    if (all >> b & 1) ins.pos[n++] = b;
  ins.n = n;
  uint64_t count = uint64_t(1) << (nbits - n);
  unsigned blocks = unsigned((count + uint64_t(kThreads) * kUnroll - 1) / (uint64_t(kThreads) * kUnroll));
  if (g.kind == QB_K_PHASE) {
    ins.setmask = all;
    V ph;
    ph.x = T(g.m[6]);
    ph.y = T(g.m[7]);
    k_apply_phase<V><<<blocks, kThreads, 0, st>>>(psi, count, ins, ph);
  } else {
    ins.setmask = ones;
    Mat2<V> m;
    m.a.x = T(g.m[0]); m.a.y = T(g.m[1]);
    m.b.x = T(g.m[2]); m.b.y = T(g.m[3]);
    m.c.x = T(g.m[4]); m.c.y = T(g.m[5]);
    m.d.x = T(g.m[6]); m.d.y = T(g.m[7]);
    k_apply_u<V><<<blocks, kThreads, 0, st>>>(psi, count, ins, uint64_t(1) << g.target, m);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gate(void *psi, int nbits, const QbGate &g, bool is_double, cudaStream_t st) {
  if (g.kind == QB_K_NOP) return cudaSuccess;
  if (is_double) return launch_gate_t<double2, double>(static_cast<double2 *>(psi), nbits, g, st);
  return launch_gate_t<float2, float>(static_cast<float2 *>(psi), nbits, g, st);
}

// ---------------------------------------------------------------------------
// init / readout
// ---------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(kThreads) k_fill_random(double2 *psi, uint64_t n, uint64_t off, uint64_t seed) {
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t a = splitmix64(seed ^ (2 * (off + i)));
    uint64_t b = splitmix64(seed ^ (2 * (off + i) + 1));
    double2 v;
    v.x = double(int64_t(a >> 11)) * (1.0 / 4503599627370496.0) - 1.0;  // [-1, 1)
    v.y = double(int64_t(b >> 11)) * (1.0 / 4503599627370496.0) - 1.0;
    psi[i] = v;
  }
}

__global__ void __launch_bounds__(kThreads) k_scale(double2 *psi, uint64_t n, double f) {
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double2 v = psi[i];
    v.x *= f;
    v.y *= f;
    psi[i] = v;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kThreads)
k_prob_mask(const double2 *__restrict__ psi, uint64_t n, uint64_t mask, uint64_t want, double *out) {
  __shared__ double part[kThreads / 32];
  double acc = 0.0;
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    if ((i & mask) == want) {
      double2 v = psi[i];
      acc += v.x * v.x + v.y * v.y;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// Index space in which "lowest index wins a tie" is decided (state.py:75 np.abs(psi).argmax() on the LOGICAL
// vector): identity for an unsharded state; for a shard, local bit b is logical bit lpos[b] and the rank bits
// contribute `hi`.
__device__ __forceinline__ uint64_t logical_index(const LogicalMap &m, uint64_t i) {
  if (m.identity) return i;
  uint64_t l = m.hi;
#pragma unroll 1
  for (int b = 0; b < m.n; ++b) l |= ((i >> b) & 1) << m.lpos[b];
  return l;
}

__global__ void __launch_bounds__(kThreads)
k_argmax(const double2 *__restrict__ psi, uint64_t n, double *blk_prob, uint64_t *blk_idx, const LogicalMap lm) {
  __shared__ double sp[kThreads / 32];
  __shared__ uint64_t si[kThreads / 32];
  double best = -1.0;
  uint64_t bidx = ~uint64_t(0);   // LOGICAL index of the best so far
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double2 v = psi[i];
    double p = v.x * v.x + v.y * v.y;
    if (p > best) {
      best = p;
      bidx = logical_index(lm, i);
    } else if (p == best && !lm.identity) {   // identity: ascending i per thread, strict > keeps the lowest index
      const uint64_t l = logical_index(lm, i);
      if (l < bidx) bidx = l;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double op = __shfl_xor_sync(0xffffffffu, best, o);
    uint64_t oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (op > best || (op == best && oi < bidx)) {
      best = op;
      bidx = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    sp[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = bidx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w)
      if (sp[w] > best || (sp[w] == best && si[w] < bidx)) {
        best = sp[w];
        bidx = si[w];
      }
    blk_prob[blockIdx.x] = best;
    blk_idx[blockIdx.x] = bidx;
  }
}

__global__ void __launch_bounds__(kThreads)
k_list_above(const double2 *__restrict__ psi, uint64_t n, double thr, uint64_t cap,
             unsigned long long *counter, uint64_t *labels, double2 *amps) {
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double2 v = psi[i];
    if (v.x * v.x + v.y * v.y >= thr) {
      unsigned long long slot = atomicAdd(counter, 1ull);
      if (slot < cap) {
        labels[slot] = i;
        amps[slot] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_cvt_f2d(const float2 *in, double2 *out, uint64_t n) {
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float2 v = in[i];
    out[i] = make_double2(v.x, v.y);
  }
}

// 148 SMs x 8 resident CTAs of 256 threads: one full wave of grid-stride workers.
constexpr int kReduceBlocks = 148 * 8;

unsigned stride_blocks(uint64_t n) {
  uint64_t want = (n + kThreads - 1) / kThreads;
  if (want < 1) want = 1;
  return unsigned(want < uint64_t(kReduceBlocks) ? want : kReduceBlocks);
}

}  // namespace

cudaError_t launch_fill_random(double2 *psi, uint64_t n, uint64_t index_offset, uint64_t seed, cudaStream_t st) {
  k_fill_random<<<stride_blocks(n), kThreads, 0, st>>>(psi, n, index_offset, seed);
  return cudaGetLastError();
}

cudaError_t launch_scale(double2 *psi, uint64_t n, double f, cudaStream_t st) {
  k_scale<<<stride_blocks(n), kThreads, 0, st>>>(psi, n, f);
  return cudaGetLastError();
}

cudaError_t launch_norm2(const double2 *psi, uint64_t n, double *out, cudaStream_t st) {
  return launch_prob_mask(psi, n, 0, 0, out, st);
}

cudaError_t launch_prob_mask(const double2 *psi, uint64_t n, uint64_t mask, uint64_t want, double *out,
                             cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), st);
  if (e != cudaSuccess) return e;
  k_prob_mask<<<stride_blocks(n), kThreads, 0, st>>>(psi, n, mask, want, out);
  return cudaGetLastError();
}

int argmax_blocks() { return kReduceBlocks; }

cudaError_t launch_argmax(const double2 *psi, uint64_t n, double *blk_prob, uint64_t *blk_idx, const LogicalMap &lm,
                          cudaStream_t st) {
  // always kReduceBlocks entries: blocks beyond the data report prob = -1
  k_argmax<<<kReduceBlocks, kThreads, 0, st>>>(psi, n, blk_prob, blk_idx, lm);
  return cudaGetLastError();
}

cudaError_t launch_list_above(const double2 *psi, uint64_t n, double thr, uint64_t cap,
                              unsigned long long *counter, uint64_t *labels, double2 *amps,
                              cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  k_list_above<<<stride_blocks(n), kThreads, 0, st>>>(psi, n, thr, cap, counter, labels, amps);
  return cudaGetLastError();
}

// ---- pair exchange over peer memory --------------------------------------------------------------
// Flattened element number k = h * 2^victim + w addresses the k-th amplitude of a half shard: local index
// (h << (victim + 1)) | (sel << victim) | w.  The kernel swaps local[(h, sel, w)] with peer[(h, 1 - sel, w)]
// for k in this rank's half of [0, 2^(nbits-1)).  Two elements per thread per iteration, both loads of
// both sides issued before any store: the remote side is an NVLink round trip away.
namespace {
__global__ void __launch_bounds__(256) k_pair_swap(double2 *__restrict__ local, double2 *__restrict__ peer, int victim,
                                                   uint64_t sel_local, uint64_t k_begin, uint64_t k_end) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  const uint64_t lowmask = (uint64_t(1) << victim) - 1;
  uint64_t k = k_begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; k + stride < k_end; k += 2 * stride) {
    const uint64_t k2 = k + stride;
    const uint64_t i0 = ((k >> victim) << (victim + 1)) | (k & lowmask);
    const uint64_t i1 = ((k2 >> victim) << (victim + 1)) | (k2 & lowmask);
    const uint64_t l0 = i0 | (sel_local << victim), r0 = i0 | ((sel_local ^ 1) << victim);
    const uint64_t l1 = i1 | (sel_local << victim), r1 = i1 | ((sel_local ^ 1) << victim);
    const double2 a0 = local[l0], b0 = peer[r0], a1 = local[l1], b1 = peer[r1];
    local[l0] = b0;
    peer[r0] = a0;
    local[l1] = b1;
    peer[r1] = a1;
  }
  if (k < k_end) {
    const uint64_t i0 = ((k >> victim) << (victim + 1)) | (k & lowmask);
    const uint64_t l0 = i0 | (sel_local << victim), r0 = i0 | ((sel_local ^ 1) << victim);
    const double2 a0 = local[l0], b0 = peer[r0];
    local[l0] = b0;
    peer[r0] = a0;
  }
}
}  // namespace

cudaError_t launch_pair_swap(double2 *local, double2 *peer, int nbits, int victim, int sel_local, int upper,
                             cudaStream_t st) {
  if (nbits < 1 || victim < 0 || victim >= nbits) return cudaErrorInvalidValue;
  const uint64_t half = uint64_t(1) << (nbits - 1);     // elements in a half shard
  const uint64_t k_begin = upper ? half / 2 : 0, k_end = upper ? half : half / 2;
  if (half == 1) {                                       // a 1-amplitude half: the lower rank swaps it
    if (upper) return cudaSuccess;
    k_pair_swap<<<1, 256, 0, st>>>(local, peer, victim, uint64_t(sel_local), 0, half);
    return cudaGetLastError();
  }
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t want = (k_end - k_begin + 511) / 512;
  const unsigned blocks = unsigned(std::min<uint64_t>(want, uint64_t(sms) * 8));
  k_pair_swap<<<blocks, 256, 0, st>>>(local, peer, victim, uint64_t(sel_local), k_begin, k_end);
  return cudaGetLastError();
}

// ---- push exchange without a pass to ride on ------------------------------------------------------
// Plain out-of-place remap: amplitude e of this shard goes to out[a >> nl][a & (2^nl - 1)], a = sigma(rank, e).
// Moved bits are >= 3 (engine.cu), so 8 consecutive lanes still write one 128-byte run; 4 independent
// 16-byte copies per thread, the remote ones posted writes over NVLink.
namespace {
__global__ void __launch_bounds__(256) k_push_remap(const double2 *__restrict__ psi, const __grid_constant__ PushMap m) {
  __shared__ double2 *s_out[kPushMaxRanks];
  if (threadIdx.x < kPushMaxRanks) s_out[threadIdx.x] = m.out[threadIdx.x];
  __syncthreads();
  const uint64_t n = uint64_t(1) << m.nl;
  const uint64_t lmask = n - 1;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t e0 = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e0 < n; e0 += 4 * stride) {
    double2 v[4];
    uint64_t a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t e = e0 + uint64_t(u) * stride;
      if (e < n) {
        v[u] = __ldcs(psi + e);
        uint64_t x = (e & ~m.moved_mask) | m.rank_term;
#pragma unroll
        for (int k = 0; k < kPushMaxMoved; ++k)
          if (k < m.nmoved) x |= ((e >> m.src[k]) & 1) << m.dst[k];
        a[u] = x;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t e = e0 + uint64_t(u) * stride;
      if (e < n) __stcs(s_out[a[u] >> m.nl] + (a[u] & lmask), v[u]);
    }
  }
}
}  // namespace

// ---- cross-rank barrier over peer memory -----------------------------------------------------------
// Every rank owns a row of arrival counters, flags[q] = how many barriers rank q has entered, living in its
// state allocation and mapped into every other rank (CUDA IPC).  Thread q of the one-warp kernel announces this
// rank's arrival in rank q's row (release store at system scope, after whatever this stream did before: the
// kernel boundary has already performed those writes) and then waits until rank q has announced itself here.
// A few microseconds over NVLink instead of an NCCL all-reduce launch; usable on any stream.
namespace {
struct BarrierPeers {
  unsigned long long *row[kPushMaxRanks];   // row[q]: rank q's counters (row[rank]: our own)
};
__global__ void __launch_bounds__(32) k_peer_barrier(const BarrierPeers peers, int rank, int nranks, unsigned long long epoch) {
  const int q = int(threadIdx.x);
  if (q >= nranks) return;
  unsigned long long *theirs = peers.row[q] + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
  const unsigned long long *mine = peers.row[rank] + q;
  unsigned long long seen = 0;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
  } while (seen < epoch);
}
}  // namespace

cudaError_t launch_peer_barrier(unsigned long long *const *rows, int rank, int nranks, unsigned long long epoch,
                                cudaStream_t st) {
  if (nranks > kPushMaxRanks) return cudaErrorInvalidValue;
  BarrierPeers p{};
  for (int q = 0; q < nranks; ++q) p.row[q] = rows[q];
  k_peer_barrier<<<1, 32, 0, st>>>(p, rank, nranks, epoch);
  return cudaGetLastError();
}

cudaError_t launch_push_remap(const double2 *psi, const PushMap &m, cudaStream_t st) {
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t n = uint64_t(1) << m.nl;
  const uint64_t want = (n + 1023) / 1024;
  const unsigned blocks = unsigned(std::min<uint64_t>(std::max<uint64_t>(want, 1), uint64_t(sms) * 8));
  k_push_remap<<<blocks, 256, 0, st>>>(psi, m);
  return cudaGetLastError();
}

cudaError_t launch_cvt_f2d(const float2 *in, double2 *out, uint64_t n, cudaStream_t st) {
  k_cvt_f2d<<<stride_blocks(n), kThreads, 0, st>>>(in, out, n);
  return cudaGetLastError();
}

}  // namespace qb
