// fused.cu -- tile-resident fused pass for sm_100a: many gates per HBM sweep.
//
// One CTA = one TILE of 2^K amplitudes (K <= 13, default 12 = 64 KiB), see qb_types.h.
//
//   0. STAGE  the pass's op and round descriptors (<= 48 x 128 B) are copied into shared
//             memory once, so the per-op decode in the hot loop is LDS broadcasts, not
//             dependent global loads.
//   1. LOAD   the tile is gathered from HBM into shared memory: thread t of 256 takes
//             tile-local indices t, t+256, ...; 8 consecutive lanes read one 128-byte
//             run, and the planner pads the tile with the lowest free index bits so the
//             runs of one CTA are mostly adjacent (a tile whose bits are 0..K-1 is one
//             contiguous 64 KiB block).  8 LDG.128 are in flight per thread before the
//             first STS; the CTAs resident on an SM are in different phases, so one
//             CTA's arithmetic overlaps the others' loads and stores.
//   2. ROUNDS each thread owns groups of 8 amplitudes that differ only in the round's 3
//             tile-local bits, pulls NG groups (16 or 32 fp64 registers) out of shared
//             memory, runs every op of the round on registers -- each op is decoded once
//             and applied to all NG groups -- and writes them back: ONE shared-memory
//             round trip for any number of gates on those 3 qubits, plus every diagonal
//             gate queued in between.
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
// Arithmetic is fp64 on the CUDA cores (no tensor cores: 0.5-3 flop/B).  At 12 fused
// h+ladder stages per sweep the fp64 pipe, not HBM, is the limiter, so the op forms are
// chosen to minimise DFMA/DMUL count: real matrices (h, ry) cost 8 instead of 20 per pair;
// h followed by its cu1 ladder is ONE op (ULADDER): y' = (c x + d y) * phase.
//
// Phase ladders: the phase of an amplitude is C_tile * T_lo[j & 63] * T_hi[j >> 6] * F[e];
// C_tile = product over partner bits outside the tile (computed once per CTA into shared
// memory), T_* = host-built 64-entry tables over the tile-local partner bits, F = the 8
// combinations of the round's own bits.
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

// m0 * x + m1 * y, same association as xgates.cc:34-35
__device__ __forceinline__ double2 mad2(double2 m0, double2 x, double2 m1, double2 y) {
  double2 p = cmul(m0, x), q = cmul(m1, y);
  return make_double2(p.x + q.x, p.y + q.y);
}

// real m0, m1
__device__ __forceinline__ double2 mad2r(double m0, double2 x, double m1, double2 y) {
  return make_double2(m0 * x.x + m1 * y.x, m0 * x.y + m1 * y.y);
}

struct Mat {
  double2 a, b, c, d;
};

// ---- butterflies on the 8 registers of one group ------------------------------------
template <int TP, bool REAL>
__device__ __forceinline__ void bfly_all(double2 (&a)[8], const Mat &m) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    const double2 x = a[e], y = a[e | (1 << TP)];
    if (REAL) {
      a[e] = mad2r(m.a.x, x, m.b.x, y);
      a[e | (1 << TP)] = mad2r(m.c.x, x, m.d.x, y);
    } else {
      a[e] = mad2(m.a, x, m.b, y);
      a[e | (1 << TP)] = mad2(m.c, x, m.d, y);
    }
  }
}

template <int TP>
__device__ __forceinline__ void bfly_masked(double2 (&a)[8], const Mat &m, uint32_t rmask, uint32_t rwant) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    if ((uint32_t(e) & rmask) == rwant) {
      const double2 x = a[e], y = a[e | (1 << TP)];
      a[e] = mad2(m.a, x, m.b, y);
      a[e | (1 << TP)] = mad2(m.c, x, m.d, y);
    }
  }
}

template <int TP>
__device__ __forceinline__ void perm_masked(double2 (&a)[8], double2 mb, double2 mc, uint32_t rmask,
                                            uint32_t rwant) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    if ((uint32_t(e) & rmask) == rwant) {
      const double2 x = a[e], y = a[e | (1 << TP)];
      a[e] = cmul(mb, y);
      a[e | (1 << TP)] = cmul(mc, x);
    }
  }
}

template <int TP>
__device__ __forceinline__ void swap_masked(double2 (&a)[8], uint32_t rmask, uint32_t rwant) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    if ((uint32_t(e) & rmask) == rwant) {
      const double2 x = a[e];
      a[e] = a[e | (1 << TP)];
      a[e | (1 << TP)] = x;
    }
  }
}

// U on the pivot, then the pivot's ladder on the pivot-set output: y' = (c x + d y) * (cf * F[e1])
template <int TP, bool REAL>
__device__ __forceinline__ void uladder(double2 (&a)[8], const Mat &m, double2 cf, const double2 *F) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    const double2 x = a[e], y = a[e | (1 << TP)];
    const double2 ph = cmul(cf, F[e | (1 << TP)]);
    double2 t;
    if (REAL) {
      a[e] = mad2r(m.a.x, x, m.b.x, y);
      t = mad2r(m.c.x, x, m.d.x, y);
    } else {
      a[e] = mad2(m.a, x, m.b, y);
      t = mad2(m.c, x, m.d, y);
    }
    a[e | (1 << TP)] = cmul(t, ph);
  }
}

#define QB_DISPATCH_TP(tp, CALL0, CALL1, CALL2) \
  do {                                          \
    if ((tp) == 0) { CALL0; }                   \
    else if ((tp) == 1) { CALL1; }              \
    else { CALL2; }                             \
  } while (0)

template <int NG>
__global__ void __launch_bounds__(kFThreads, NG == 1 ? 3 : 2)
k_fused_pass(const __grid_constant__ FusedParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = P.desc.K;
  const uint32_t tileN = 1u << K;
  double2 *tile = reinterpret_cast<double2 *>(smem_raw);
  double2 *s_pout = tile + tileN;                                       // kMaxLadders
  QbOp *s_ops = reinterpret_cast<QbOp *>(s_pout + kMaxLadders);         // QB_MAX_PASS_OPS
  QbRound *s_rounds = reinterpret_cast<QbRound *>(s_ops + QB_MAX_PASS_OPS);  // QB_MAX_PASS_ROUNDS
  uint32_t *hi_off = reinterpret_cast<uint32_t *>(s_rounds + QB_MAX_PASS_ROUNDS);  // 2^(K-3)
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
  // ---- STAGE: descriptors -> shared memory ------------------------------------------
  {
    const int4 *src = reinterpret_cast<const int4 *>(P.ops);
    int4 *dst = reinterpret_cast<int4 *>(s_ops);
    const int n16 = P.desc.nops * int(sizeof(QbOp) / 16);
    for (int i = tid; i < n16; i += kFThreads) dst[i] = __ldg(src + i);
    const int32_t *rs = reinterpret_cast<const int32_t *>(P.rounds);
    int32_t *rd = reinterpret_cast<int32_t *>(s_rounds);
    const int n4 = P.desc.nrounds * int(sizeof(QbRound) / 4);
    for (int i = tid; i < n4; i += kFThreads) rd[i] = __ldg(rs + i);
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
    if (op->kind == QB_K_LADDER || op->kind == QB_K_ULADDER) {
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
    const QbRound *R = s_rounds + r;
    const uint32_t d0 = swz(1u << R->rbit[0]);
    const uint32_t d1 = swz(1u << R->rbit[1]);
    const uint32_t d2 = swz(1u << R->rbit[2]);
    const int ob = R->op_begin, oe = R->op_end;
    for (uint32_t q0 = tid; q0 < ngroups; q0 += kFThreads * NG) {
      uint32_t jb[NG], pb[NG];
      bool valid[NG];
      double2 a[NG][8];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const uint32_t q = q0 + g * kFThreads;
        valid[g] = q < ngroups;
        uint32_t j = 0;
        for (int k = 0; k < K - 3; ++k) j |= ((q >> k) & 1u) << R->qmap[k];
        jb[g] = j;
        pb[g] = swz(j);
        if (valid[g]) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            a[g][e] = tile[pb[g] ^ ((e & 1) ? d0 : 0u) ^ ((e & 2) ? d1 : 0u) ^ ((e & 4) ? d2 : 0u)];
        }
      }

      for (int oi = ob; oi < oe; ++oi) {
        const QbOp *op = s_ops + oi;
        const int4 h0 = *reinterpret_cast<const int4 *>(op);            // kind tpos lmask lwant
        const int4 h1 = *(reinterpret_cast<const int4 *>(op) + 1);      // rmask rwant table_off flags
        const ulonglong2 h2 = *(reinterpret_cast<const ulonglong2 *>(op) + 2);  // gmask gwant
        if ((base & h2.x) != h2.y) continue;                            // uniform per CTA
        const int kind = h0.x, tp = h0.y;
        const uint32_t lmask = uint32_t(h0.z), lwant = uint32_t(h0.w);
        const uint32_t rmask = uint32_t(h1.x), rwant = uint32_t(h1.y);
        bool ok[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) ok[g] = valid[g] && ((jb[g] & lmask) == lwant);
        const double2 *mp = reinterpret_cast<const double2 *>(op->m);
        switch (kind) {
          case QB_K_U: {
            Mat m{mp[0], mp[1], mp[2], mp[3]};
            if (rmask == 0) {
              if (op->mflags & QB_MF_REAL) {
#pragma unroll
                for (int g = 0; g < NG; ++g)
                  if (ok[g])
                    QB_DISPATCH_TP(tp, (bfly_all<0, true>(a[g], m)), (bfly_all<1, true>(a[g], m)),
                                   (bfly_all<2, true>(a[g], m)));
              } else {
#pragma unroll
                for (int g = 0; g < NG; ++g)
                  if (ok[g])
                    QB_DISPATCH_TP(tp, (bfly_all<0, false>(a[g], m)), (bfly_all<1, false>(a[g], m)),
                                   (bfly_all<2, false>(a[g], m)));
              }
            } else {
#pragma unroll
              for (int g = 0; g < NG; ++g)
                if (ok[g])
                  QB_DISPATCH_TP(tp, (bfly_masked<0>(a[g], m, rmask, rwant)), (bfly_masked<1>(a[g], m, rmask, rwant)),
                                 (bfly_masked<2>(a[g], m, rmask, rwant)));
            }
            break;
          }
          case QB_K_ULADDER: {
            Mat m{mp[0], mp[1], mp[2], mp[3]};
            const double2 *tb = P.tables + h1.z;
            const double2 *F = tb + 64 + (1 << hi_bits);
            const double2 cp = s_pout[h1.w];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
              if (!ok[g]) continue;
              double2 c = cmul(cp, __ldg(tb + (jb[g] & 63u)));
              if (hi_bits) c = cmul(c, __ldg(tb + 64 + (jb[g] >> QB_LADDER_CHUNK)));
              if (op->mflags & QB_MF_REAL)
                QB_DISPATCH_TP(tp, (uladder<0, true>(a[g], m, c, F)), (uladder<1, true>(a[g], m, c, F)),
                               (uladder<2, true>(a[g], m, c, F)));
              else
                QB_DISPATCH_TP(tp, (uladder<0, false>(a[g], m, c, F)), (uladder<1, false>(a[g], m, c, F)),
                               (uladder<2, false>(a[g], m, c, F)));
            }
            break;
          }
          case QB_K_PERM: {
            const double2 mb = mp[1], mc = mp[2];
#pragma unroll
            for (int g = 0; g < NG; ++g)
              if (ok[g])
                QB_DISPATCH_TP(tp, (perm_masked<0>(a[g], mb, mc, rmask, rwant)),
                               (perm_masked<1>(a[g], mb, mc, rmask, rwant)),
                               (perm_masked<2>(a[g], mb, mc, rmask, rwant)));
            break;
          }
          case QB_K_SWAP: {
#pragma unroll
            for (int g = 0; g < NG; ++g)
              if (ok[g])
                QB_DISPATCH_TP(tp, (swap_masked<0>(a[g], rmask, rwant)), (swap_masked<1>(a[g], rmask, rwant)),
                               (swap_masked<2>(a[g], rmask, rwant)));
            break;
          }
          case QB_K_PHASE: {
            const double2 ph = mp[0];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
              if (!ok[g]) continue;
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if ((uint32_t(e) & rmask) == rwant) a[g][e] = cmul(ph, a[g][e]);
            }
            break;
          }
          case QB_K_LADDER: {
            const double2 *tb = P.tables + h1.z;
            const double2 *F = tb + 64 + (1 << hi_bits);
            const double2 cp = s_pout[h1.w];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
              if (!ok[g]) continue;
              double2 c = cmul(cp, __ldg(tb + (jb[g] & 63u)));
              if (hi_bits) c = cmul(c, __ldg(tb + 64 + (jb[g] >> QB_LADDER_CHUNK)));
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if ((uint32_t(e) & rmask) == rwant) a[g][e] = cmul(cmul(c, __ldg(F + e)), a[g][e]);
            }
            break;
          }
          default:
            break;
        }
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (valid[g]) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            tile[pb[g] ^ ((e & 1) ? d0 : 0u) ^ ((e & 2) ? d1 : 0u) ^ ((e & 4) ? d2 : 0u)] = a[g][e];
        }
      }
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
         QB_MAX_PASS_OPS * sizeof(QbOp) + QB_MAX_PASS_ROUNDS * sizeof(QbRound) +
         (size_t(1) << (K - 3)) * sizeof(uint32_t);
}

int g_groups = 0;  // 0 = not configured

}  // namespace

cudaError_t fused_configure(int device) {
  (void)device;
  static_assert(sizeof(QbOp) == 128, "QbOp is read as 16-byte pieces");
  static_assert(sizeof(QbRound) % 4 == 0 && (QB_MAX_PASS_OPS * sizeof(QbOp)) % 16 == 0, "smem layout");
  if (g_groups == 0) {
    const char *e = getenv("QCC_B200_FUSED_GROUPS");
    g_groups = (e && atoi(e) == 1) ? 1 : 2;
  }
  cudaError_t err = cudaFuncSetAttribute(k_fused_pass<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(fused_smem_bytes(QB_MAX_TILE_BITS)));
  if (err != cudaSuccess) return err;
  return cudaFuncSetAttribute(k_fused_pass<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
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
  if (p.desc.nops > QB_MAX_PASS_OPS || p.desc.nrounds > QB_MAX_PASS_ROUNDS) return cudaErrorInvalidValue;
  unsigned blocks = 1u << (nbits - K);
  if (g_groups == 1)
    k_fused_pass<1><<<blocks, kFThreads, fused_smem_bytes(K), st>>>(P);
  else
    k_fused_pass<2><<<blocks, kFThreads, fused_smem_bytes(K), st>>>(P);
  return cudaGetLastError();
}

}  // namespace qb
