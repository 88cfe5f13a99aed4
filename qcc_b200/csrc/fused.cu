// fused.cu -- tile-resident fused pass for sm_100a: many gates per HBM sweep.
//
// One CTA = one TILE of 2^K amplitudes (K <= 13, default 12 = 64 KiB), see qb_types.h.
//
//   1. LOAD   the tile is gathered from HBM into shared memory: thread t of 256 takes
//             tile-local indices t, t+256, ...; 8 consecutive lanes read one 128-byte
//             run, and the planner pads the tile with the lowest free index bits so the
//             runs of one CTA are mostly adjacent (a tile whose bits are 0..K-1 is one
//             contiguous 64 KiB block).  8 LDG.128 are in flight per thread before the
//             first STS; with 3 CTAs resident per SM (3 x 67 KiB shared) one CTA's
//             arithmetic overlaps the other two's loads/stores.
//   2. ROUNDS each thread owns groups of 8 amplitudes that differ only in the round's 3
//             tile-local bits, pulls them into 16 fp64 registers, runs every op of the
//             round on registers, writes them back: ONE shared-memory round trip for
//             any number of gates on those 3 qubits, plus every diagonal gate that
//             happens to be queued in between.
//   3. STORE  the mirror image of LOAD with streaming stores.
//
// Shared-memory layout: the tile is stored XOR-swizzled, slot(j) = j ^ (fold(j >> 3) & 7)
// with fold(x) = x ^ x>>3 ^ x>>6 ^ x>>9, in 16-byte units.  The 16-byte bank group of j is
// then the XOR of its index bits taken mod 3, so (a) the linear LOAD/STORE pattern is
// conflict free, and (b) for ANY choice of round bits the planner can hand group-index
// bits 0..2 to one free local bit of each class (QbRound::qmap), which makes every
// quarter-warp of an LDS.128/STS.128 hit 8 distinct bank groups.  The swizzle is linear
// over XOR, so slot(base | spread(e)) = slot(base) ^ slot(spread(e)): one XOR per register.
//
// Phase ladders (QB_K_LADDER): the phase of an amplitude is
//     C_tile * T_lo[j & 63] * T_hi[j >> 6] * F[e]
// C_tile = product over partner bits outside the tile (computed once per CTA into shared
// memory), T_* = host-built 64-entry tables over the tile-local partner bits, F = the 8
// combinations of the round's own bits.  A whole QFT ladder (up to n-1 cu1 gates) costs
// two table loads and two complex multiplies per group plus two per touched amplitude.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace qb {

namespace {

constexpr int kFThreads = 256;
constexpr int kMaxLadders = 64;

struct FusedParams {
  double2 *psi;
  int nbits;
  QbPassDesc desc;
  const QbOp *ops;
  const QbRound *rounds;
  const double2 *tables;
  const int32_t *outbits;
};

__device__ __forceinline__ uint32_t swz(uint32_t j) {
  uint32_t x = j >> 3;
  x ^= x >> 3;
  x ^= x >> 6;
  return j ^ (x & 7u);
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ double2 mad2(double2 m0, double2 x, double2 m1, double2 y) {
  // m0 * x + m1 * y, same association as xgates.cc:34-35
  double2 p = cmul(m0, x), q = cmul(m1, y);
  return make_double2(p.x + q.x, p.y + q.y);
}

template <int TP>
__device__ __forceinline__ void op_u(double2 (&a)[8], double2 ma, double2 mb, double2 mc, double2 md,
                                     uint32_t rmask, uint32_t rwant) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    if ((uint32_t(e) & rmask) == rwant) {
      double2 x = a[e], y = a[e | (1 << TP)];
      a[e] = mad2(ma, x, mb, y);
      a[e | (1 << TP)] = mad2(mc, x, md, y);
    }
  }
}

template <int TP>
__device__ __forceinline__ void op_perm(double2 (&a)[8], double2 mb, double2 mc, uint32_t rmask,
                                        uint32_t rwant) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    if ((uint32_t(e) & rmask) == rwant) {
      double2 x = a[e], y = a[e | (1 << TP)];
      a[e] = cmul(mb, y);
      a[e | (1 << TP)] = cmul(mc, x);
    }
  }
}

template <int TP>
__device__ __forceinline__ void op_swap(double2 (&a)[8], uint32_t rmask, uint32_t rwant) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    if ((uint32_t(e) & rmask) == rwant) {
      double2 x = a[e];
      a[e] = a[e | (1 << TP)];
      a[e | (1 << TP)] = x;
    }
  }
}

__global__ void __launch_bounds__(kFThreads, 3) k_fused_pass(const __grid_constant__ FusedParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = P.desc.K;
  const uint32_t tileN = 1u << K;
  double2 *tile = reinterpret_cast<double2 *>(smem_raw);
  double2 *s_pout = tile + tileN;                                      // kMaxLadders
  uint32_t *hi_off = reinterpret_cast<uint32_t *>(s_pout + kMaxLadders);  // 2^(K-3)
  const uint32_t tid = threadIdx.x;

  // ---- tile base: scatter blockIdx.x over the non-tile index bits -----------------
  uint64_t base = 0;
  {
    uint64_t t = blockIdx.x;
    const uint64_t tmask = P.desc.tile_mask;
    for (int b = 0; b < P.nbits; ++b) {
      if (!((tmask >> b) & 1)) {
        base |= (t & 1) << b;
        t >>= 1;
      }
    }
  }
  // ---- offset of every 8-amplitude run of the tile (in units of 8 amplitudes) ------
  for (uint32_t h = tid; h < (tileN >> 3); h += kFThreads) {
    uint64_t off = 0;
    for (int k = 3; k < K; ++k) off |= uint64_t((h >> (k - 3)) & 1u) << P.desc.tile_bits[k];
    hi_off[h] = uint32_t(off >> 3);
  }
  // ---- per-tile constants of the phase ladders ---------------------------------------
  const int hi_bits = K > QB_LADDER_CHUNK ? K - QB_LADDER_CHUNK : 0;
  for (int oi = tid; oi < P.desc.nops; oi += kFThreads) {
    const QbOp *op = P.ops + oi;
    if (op->kind == QB_K_LADDER) {
      const double2 *tb = P.tables + op->table_off + 64 + (1 << hi_bits) + 8;
      double2 c = tb[0];
      const int32_t *ob = P.outbits + op->out_off;
      for (int k = 0; k < op->nout; ++k)
        if ((base >> ob[k]) & 1) c = cmul(c, tb[1 + k]);
      s_pout[op->flags] = c;
    }
  }
  __syncthreads();

  // ---- LOAD ----------------------------------------------------------------------------
  double2 *__restrict__ psi = P.psi;
  for (uint32_t s0 = 0; s0 < tileN; s0 += kFThreads * 8) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      uint32_t j = s0 + u * kFThreads + tid;
      if (j < tileN) v[u] = __ldcs(psi + (base | (uint64_t(hi_off[j >> 3]) << 3) | (j & 7u)));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      uint32_t j = s0 + u * kFThreads + tid;
      if (j < tileN) tile[swz(j)] = v[u];
    }
  }
  __syncthreads();

  // ---- ROUNDS --------------------------------------------------------------------------
  const uint32_t ngroups = tileN >> 3;
  for (int r = 0; r < P.desc.nrounds; ++r) {
    const QbRound *R = P.rounds + r;
    const uint32_t d0 = swz(1u << R->rbit[0]);
    const uint32_t d1 = swz(1u << R->rbit[1]);
    const uint32_t d2 = swz(1u << R->rbit[2]);
    const int ob = R->op_begin, oe = R->op_end;
    for (uint32_t q = tid; q < ngroups; q += kFThreads) {
      uint32_t jb = 0;
      for (int k = 0; k < K - 3; ++k) jb |= ((q >> k) & 1u) << R->qmap[k];
      const uint32_t pb = swz(jb);
      double2 a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        a[e] = tile[pb ^ ((e & 1) ? d0 : 0u) ^ ((e & 2) ? d1 : 0u) ^ ((e & 4) ? d2 : 0u)];

      for (int oi = ob; oi < oe; ++oi) {
        const QbOp *op = P.ops + oi;
        const int4 h0 = __ldg(reinterpret_cast<const int4 *>(op));      // kind tpos lmask lwant
        const int4 h1 = __ldg(reinterpret_cast<const int4 *>(op) + 1);  // rmask rwant table_off flags
        const ulonglong2 h2 = __ldg(reinterpret_cast<const ulonglong2 *>(op) + 2);  // gmask gwant
        if ((base & h2.x) != h2.y) continue;                            // uniform per CTA
        if ((jb & uint32_t(h0.z)) != uint32_t(h0.w)) continue;          // per group
        const uint32_t rmask = uint32_t(h1.x), rwant = uint32_t(h1.y);
        const double2 *mp = reinterpret_cast<const double2 *>(op->m);
        switch (h0.x) {
          case QB_K_U: {
            const double2 ma = __ldg(mp), mb = __ldg(mp + 1), mc = __ldg(mp + 2), md = __ldg(mp + 3);
            if (h0.y == 0) op_u<0>(a, ma, mb, mc, md, rmask, rwant);
            else if (h0.y == 1) op_u<1>(a, ma, mb, mc, md, rmask, rwant);
            else op_u<2>(a, ma, mb, mc, md, rmask, rwant);
            break;
          }
          case QB_K_PERM: {
            const double2 mb = __ldg(mp + 1), mc = __ldg(mp + 2);
            if (h0.y == 0) op_perm<0>(a, mb, mc, rmask, rwant);
            else if (h0.y == 1) op_perm<1>(a, mb, mc, rmask, rwant);
            else op_perm<2>(a, mb, mc, rmask, rwant);
            break;
          }
          case QB_K_SWAP: {
            if (h0.y == 0) op_swap<0>(a, rmask, rwant);
            else if (h0.y == 1) op_swap<1>(a, rmask, rwant);
            else op_swap<2>(a, rmask, rwant);
            break;
          }
          case QB_K_PHASE: {
            const double2 ph = __ldg(mp);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if ((uint32_t(e) & rmask) == rwant) a[e] = cmul(ph, a[e]);
            break;
          }
          case QB_K_LADDER: {
            const double2 *tb = P.tables + h1.z;
            double2 c = cmul(s_pout[h1.w], __ldg(tb + (jb & 63u)));
            if (hi_bits) c = cmul(c, __ldg(tb + 64 + (jb >> QB_LADDER_CHUNK)));
            const double2 *F = tb + 64 + (1 << hi_bits);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if ((uint32_t(e) & rmask) == rwant) a[e] = cmul(cmul(c, __ldg(F + e)), a[e]);
            break;
          }
          default:
            break;
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e)
        tile[pb ^ ((e & 1) ? d0 : 0u) ^ ((e & 2) ? d1 : 0u) ^ ((e & 4) ? d2 : 0u)] = a[e];
    }
    __syncthreads();
  }

  // ---- STORE ---------------------------------------------------------------------------
  for (uint32_t s0 = 0; s0 < tileN; s0 += kFThreads * 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      uint32_t j = s0 + u * kFThreads + tid;
      if (j < tileN) __stcs(psi + (base | (uint64_t(hi_off[j >> 3]) << 3) | (j & 7u)), tile[swz(j)]);
    }
  }
}

size_t fused_smem_bytes(int K) {
  return (size_t(1) << K) * sizeof(double2) + kMaxLadders * sizeof(double2) +
         (size_t(1) << (K - 3)) * sizeof(uint32_t);
}

}  // namespace

cudaError_t fused_configure(int device) {
  (void)device;
  return cudaFuncSetAttribute(k_fused_pass, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              int(fused_smem_bytes(QB_MAX_TILE_BITS)));
}

cudaError_t launch_fused_pass(double2 *psi, int nbits, const DevicePass &p, cudaStream_t st) {
  FusedParams P;
  P.psi = psi;
  P.nbits = nbits;
  P.desc = p.desc;
  P.ops = p.ops;
  P.rounds = p.rounds;
  P.tables = p.tables;
  P.outbits = p.outbits;
  const int K = p.desc.K;
  if (K < 4 || K > QB_MAX_TILE_BITS || K > nbits) return cudaErrorInvalidValue;
  unsigned blocks = 1u << (nbits - K);
  k_fused_pass<<<blocks, kFThreads, fused_smem_bytes(K), st>>>(P);
  return cudaGetLastError();
}

}  // namespace qb
