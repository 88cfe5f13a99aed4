// fused.cu -- tile-resident fused pass for sm_100a: many gates per HBM sweep.
//
// PERSISTENT kernel: one CTA of 512 threads per SM walks over TILES of 2^K amplitudes
// (K <= 13, default 12 = 64 KiB; see qb_types.h) with a two-deep shared-memory ring: while
// tile i is being computed on, tile i+1 is already streaming in through cp.async (LDGSTS,
// 16 B per request, L2-only), so HBM latency is hidden behind the fp64 work instead of being
// paid twice per tile.  (With one-shot CTAs the resident CTAs of an SM ran in lockstep --
// all loading, then all computing -- and the phases simply added up: 9.4 + 3.7 + 6 ms.)
//
//   0. STAGE  the pass's op and round descriptors (<= 48 x 128 B) are copied into shared
//             memory once, so the per-op decode in the hot loop is LDS broadcasts, not
//             dependent global loads.
//   1. LOAD   the tile is gathered from HBM straight into (swizzled) shared memory with
//             cp.async: thread t of 512 takes tile-local indices t, t+512, ...; 8
//             consecutive lanes fetch one 128-byte run, and the planner pads the tile with
//             the lowest free index bits so the runs of one tile are mostly adjacent (a
//             tile whose bits are 0..K-1 is one contiguous 64 KiB block).  No registers are
//             staged and nobody waits: the copy of the NEXT tile is issued before the
//             rounds of the current one start.
//   2. ROUNDS each thread owns one group of 8 amplitudes that differ only in the round's 3
//             tile-local bits, pulls it into 16 fp64 registers, runs every op of the round
//             on registers and writes it back: ONE shared-memory round trip for any number
//             of gates on those 3 qubits, plus every diagonal gate queued in between.
//   3. STORE  shared -> HBM with streaming 128-bit stores (fire and forget).
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
// C_tile = product over partner bits outside the tile (one warp per ladder, once per tile,
// folded into that tile's copy of T_lo), T_* = host-built 64-entry tables over the tile-local
// partner bits, F = the 8 combinations of the round's own bits (constant bank).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace qb {

namespace {

constexpr int kFThreads = 256;
constexpr int kMaxLadders = 64;

struct FusedParams {
  double2 *psi;
  int nbits;
  QbPassDesc desc;
  const double2 *tables;
  const double2 *outph;
  const int32_t *outbits;
  const uint32_t *jbtab;
  int debug;  // timing experiments only -- 1: skip the op loop, 2: skip the rounds, 4: skip the store, 8: skip the load
  int nbuf;   // tile buffers in the shared-memory ring (1 or 2)
  int stagger_ns;  // first-wave start offset between the CTA slots of an SM (see launch_fused_pass)
  int sms;
  int nbits_out;  // entries in outbits
  // Pass descriptors travel as kernel parameters (7 KiB of the 32 KiB parameter space): the
  // per-op decode in the hot loop is then LDC from the constant bank (warp-uniform index), which
  // costs neither shared-memory wavefronts nor LSU issue slots.
  QbRound rounds[QB_MAX_PASS_ROUNDS];
  QbOp ops[QB_MAX_PASS_OPS];
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

// The QFT case: U = r * [[1, 1], [1, -1]] (Hadamard) on the pivot, then the ladder.  The scale r
// is folded into the phase, so a pair costs 14 fp64 instructions instead of 16:
//   x' = r (x + y),   y' = (x - y) * (r * cf * F[e1]);  F[pivot only] is exactly 1 by construction
// (the planner moves the pivot's own phase into the per-tile constant).
template <int TP>
__device__ __forceinline__ void hladder(double2 (&a)[8], double r, double2 cf, const double2 *F) {
  const double2 cr = make_double2(r * cf.x, r * cf.y);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    const double2 x = a[e], y = a[e | (1 << TP)];
    const double2 ph = (e == 0) ? cr : cmul(cr, F[e | (1 << TP)]);
    const double2 d = make_double2(x.x - y.x, x.y - y.y);
    a[e] = make_double2(r * (x.x + y.x), r * (x.y + y.y));
    a[e | (1 << TP)] = cmul(d, ph);
  }
}

#define QB_DISPATCH_TP(tp, CALL0, CALL1, CALL2) \
  do {                                          \
    if ((tp) == 0) { CALL0; }                   \
    else if ((tp) == 1) { CALL1; }              \
    else { CALL2; }                             \
  } while (0)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const uint32_t sa = uint32_t(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// FULL: the tile has a multiple of 256 groups (K >= 11), so the group loop has a trip count that
// is uniform across the CTA and needs no bounds test.  That matters beyond the saved compare: with a
// thread-dependent loop condition the compiler must treat the op loop inside as divergent and keeps
// the op index -- and with it every descriptor load -- in vector registers (vector-indexed LDC,
// vector compares and branches per op); with a uniform trip count the decode runs on the uniform
// datapath.
template <bool FULL>
__global__ void __launch_bounds__(kFThreads, 2) k_fused_pass(const __grid_constant__ FusedParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = P.desc.K;
  const uint32_t tileN = 1u << K;
  const int nbuf = P.nbuf;
  double2 *tiles = reinterpret_cast<double2 *>(smem_raw);               // nbuf x 2^K
  double2 *s_pout = tiles + size_t(nbuf) * tileN;                       // kMaxLadders
  uint32_t *s_active = reinterpret_cast<uint32_t *>(s_pout + kMaxLadders);  // 4 words: ops whose
                                                                        // outside-tile predicate holds for this tile
  double2 *s_tab = s_pout + kMaxLadders + 1;                            // ntable
  double2 *s_outph = s_tab + P.desc.ntable;                             // nout_total (+1 pad)
  uint32_t *hi_off = reinterpret_cast<uint32_t *>(s_outph + P.desc.nout_total + 1);  // 2^(K-3)
  int32_t *s_outbits = reinterpret_cast<int32_t *>(hi_off + (tileN >> 3));  // nout_total
  const QbOp *s_ops = P.ops;        // constant bank
  const QbRound *s_rounds = P.rounds;
  const uint32_t tid = threadIdx.x;
  double2 *__restrict__ psi = P.psi;

  // De-phase the CTAs that share an SM.  All CTAs do identical work, so the ones launched
  // together stay in lockstep for the whole kernel -- every resident CTA loading, then every
  // one computing -- and HBM, shared memory and the fp64 pipe are used one after the other
  // instead of concurrently.  Delaying the 2nd / 3rd CTA slot of each SM once, in the first
  // wave only, keeps the slots a third of a tile apart from then on.
  if (P.stagger_ns && blockIdx.x < 3u * uint32_t(P.sms)) {
    const uint32_t slot = blockIdx.x / uint32_t(P.sms);
    for (uint32_t k = 0; k < slot; ++k) __nanosleep(uint32_t(P.stagger_ns));
  }

  // ---- STAGE (once per CTA): ladder tables, run offsets -> shared memory --------------------
  {
    for (int i = tid; i < P.desc.ntable; i += kFThreads) s_tab[i] = __ldg(P.tables + i);
    // ladder constants: per-tile factors are rebuilt from these for every tile
    for (int i = tid; i < P.desc.nout_total; i += kFThreads) s_outph[i] = __ldg(P.outph + i);
    for (int i = tid; i < P.nbits_out; i += kFThreads) s_outbits[i] = __ldg(P.outbits + i);
    // offset of every 8-amplitude run of a tile, in units of 8 amplitudes
    for (uint32_t h = tid; h < (tileN >> 3); h += kFThreads) {
      uint64_t off = 0;
      for (int k = 3; k < K; ++k) off |= uint64_t((h >> (k - 3)) & 1u) << P.desc.tile_bits[k];
      hi_off[h] = uint32_t(off >> 3);
    }
  }
  __syncthreads();

  const uint32_t ntiles = 1u << (P.nbits - K);
  const uint64_t tmask = P.desc.tile_mask;
  // tile number -> index bits outside the tile
  auto tile_base = [&](uint32_t t) {
    uint64_t b = 0, tt = t;
    if (P.desc.nseg >= 0) {
#pragma unroll
      for (int r = 0; r < QB_MAX_SEGS; ++r) {
        if (r < P.desc.nseg) {
          const int len = P.desc.seg_len[r];
          b |= (tt & ((uint64_t(1) << len) - 1)) << P.desc.seg_pos[r];
          tt >>= len;
        }
      }
      return b;
    }
    for (int bit = 0; bit < P.nbits; ++bit) {
      if (!((tmask >> bit) & 1)) {
        b |= (tt & 1) << bit;
        tt >>= 1;
      }
    }
    return b;
  };
  auto issue_load = [&](uint32_t t, double2 *buf) {
    const uint64_t b = tile_base(t);
    if (!(P.debug & 8))
      for (uint32_t j = tid; j < tileN; j += kFThreads)
        cp_async16(buf + swz(j), psi + (b | (uint64_t(hi_off[j >> 3]) << 3) | (j & 7u)));
    cp_async_commit();
  };

  const int hi_bits = K > QB_LADDER_CHUNK ? K - QB_LADDER_CHUNK : 0;
  const uint32_t ngroups = tileN >> 3;
  int cur = 0;
  if (blockIdx.x < ntiles) issue_load(blockIdx.x, tiles);
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint32_t tn = t + gridDim.x;
    const bool more = tn < ntiles;
    double2 *tile = tiles + size_t(cur) * tileN;
    if (nbuf == 2 && more) issue_load(tn, tiles + size_t(cur ^ 1) * tileN);  // prefetch
    const uint64_t base = tile_base(t);
    // which ops apply to this tile at all (their controls outside the tile): one bit per op
    if (tid < 64) {
      const bool on = int(tid) < P.desc.nops && (base & s_ops[tid].gmask) == s_ops[tid].gwant;
      const uint32_t bal = __ballot_sync(0xffffffffu, on);
      if ((tid & 31u) == 0) s_active[tid >> 5] = bal;
    }
    // per-tile constants of the phase ladders (overlaps the wait for the tile's data): one
    // warp per ladder, lane k owns outside bit k, product by butterfly shuffles
    for (int oi = int(tid >> 5); oi < P.desc.nops; oi += kFThreads / 32) {
      const QbOp *op = s_ops + oi;
      const int k8 = op->kind & 0xff;
      if (k8 == QB_K_LADDER || k8 == QB_K_ULADDER) {
        const int lane = int(tid & 31u);
        const double2 *ph = s_outph + op->outph_off;
        double2 c = make_double2(1.0, 0.0);
        if (lane < op->nout && ((base >> s_outbits[op->out_off + lane]) & 1)) c = ph[1 + lane];
        if (lane == 0) c = cmul(c, ph[0]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          double2 d;
          d.x = __shfl_xor_sync(0xffffffffu, c.x, o);
          d.y = __shfl_xor_sync(0xffffffffu, c.y, o);
          c = cmul(c, d);
        }
        // Fold the per-tile constant into this tile's copy of T_lo (64 entries, 2 per lane), so
        // the hot loop needs one multiply less and one dependent shared-memory read less per op.
        const double2 *t0 = P.tables + op->table_off;
        double2 *t1 = s_tab + op->table_off;
        t1[lane] = cmul(__ldg(t0 + lane), c);
        t1[lane + 32] = cmul(__ldg(t0 + lane + 32), c);
      }
    }
    if (nbuf == 2 && more) cp_async_wait<1>();
    else cp_async_wait<0>();
    __syncthreads();

    // ---- ROUNDS ------------------------------------------------------------------------
    for (int r = 0; r < ((P.debug & 2) ? 0 : P.desc.nrounds); ++r) {
      const QbRound *R = s_rounds + r;
      const uint32_t d0 = swz(1u << R->rbit[0]);
      const uint32_t d1 = swz(1u << R->rbit[1]);
      const uint32_t d2 = swz(1u << R->rbit[2]);
      const int ob = R->op_begin, oe = (P.debug & 1) ? R->op_begin : R->op_end;
      const uint32_t *jbt = P.jbtab + (size_t(r) << P.desc.ngroups_log2);
      const uint32_t giters = FULL ? (ngroups / kFThreads) : ((ngroups + kFThreads - 1) / kFThreads);
      for (uint32_t git = 0; git < giters; ++git) {
        const uint32_t q = git * kFThreads + tid;
        if (!FULL && q >= ngroups) break;
        const uint32_t w = __ldg(jbt + q);
        const uint32_t jb = w & 0xffffu, pb = w >> 16;
        double2 a[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          a[e] = tile[pb ^ ((e & 1) ? d0 : 0u) ^ ((e & 2) ? d1 : 0u) ^ ((e & 4) ? d2 : 0u)];

        // One dense opcode per (kind, target position, matrix class): a single jump-table switch
        // replaces the chain of compare-and-branch steps a generic (kind, tpos, flags) decode needs.
        const uint64_t active = (uint64_t(s_active[1]) << 32) | s_active[0];
#pragma unroll 1
        for (int oi = ob; oi < oe; ++oi) {
          if (!((active >> oi) & 1)) continue;                            // uniform per tile
          const QbOp *op = s_ops + oi;
          const double2 *mp = reinterpret_cast<const double2 *>(op->m);
          const int opc = int(uint32_t(op->kind) >> 24);
          if (opc < QB_OPC_U_ALL) {
            // ULADDER family: uncontrolled butterfly on the pivot + the pivot's phase ladder
            const double2 *tb = s_tab + op->table_off;
            const double2 *F = reinterpret_cast<const double2 *>(op->F);  // constant bank
            double2 c = tb[(jb & 63u) ^ ((jb >> 3) & 7u)];   // already carries the per-tile constant
            if (hi_bits) c = cmul(c, tb[64 + (jb >> QB_LADDER_CHUNK)]);
            switch (opc) {
              case 0: { const Mat m{mp[0], mp[1], mp[2], mp[3]}; uladder<0, false>(a, m, c, F); break; }
              case 1: { const Mat m{mp[0], mp[1], mp[2], mp[3]}; uladder<1, false>(a, m, c, F); break; }
              case 2: { const Mat m{mp[0], mp[1], mp[2], mp[3]}; uladder<2, false>(a, m, c, F); break; }
              case 3: case 4: case 5: {
                Mat m;
                m.a.x = mp[0].x; m.b.x = mp[1].x; m.c.x = mp[2].x; m.d.x = mp[3].x;
                if (opc == 3) uladder<0, true>(a, m, c, F);
                else if (opc == 4) uladder<1, true>(a, m, c, F);
                else uladder<2, true>(a, m, c, F);
                break;
              }
              case 6: hladder<0>(a, mp[0].x, c, F); break;
              case 7: hladder<1>(a, mp[0].x, c, F); break;
              default: hladder<2>(a, mp[0].x, c, F); break;
            }
            continue;
          }
          const uint32_t rmask = op->rmask, rwant = op->rwant;
          if ((jb & op->lmask) != op->lwant) continue;                   // per group
          switch (opc) {
            case QB_OPC_U_ALL + 0: { const Mat m{mp[0], mp[1], mp[2], mp[3]}; bfly_all<0, false>(a, m); break; }
            case QB_OPC_U_ALL + 1: { const Mat m{mp[0], mp[1], mp[2], mp[3]}; bfly_all<1, false>(a, m); break; }
            case QB_OPC_U_ALL + 2: { const Mat m{mp[0], mp[1], mp[2], mp[3]}; bfly_all<2, false>(a, m); break; }
            case QB_OPC_U_ALL + 3: case QB_OPC_U_ALL + 4: case QB_OPC_U_ALL + 5: {
              Mat m;
              m.a.x = mp[0].x; m.b.x = mp[1].x; m.c.x = mp[2].x; m.d.x = mp[3].x;
              if (opc == QB_OPC_U_ALL + 3) bfly_all<0, true>(a, m);
              else if (opc == QB_OPC_U_ALL + 4) bfly_all<1, true>(a, m);
              else bfly_all<2, true>(a, m);
              break;
            }
            case QB_OPC_U_MASKED + 0: { const Mat m{mp[0], mp[1], mp[2], mp[3]}; bfly_masked<0>(a, m, rmask, rwant); break; }
            case QB_OPC_U_MASKED + 1: { const Mat m{mp[0], mp[1], mp[2], mp[3]}; bfly_masked<1>(a, m, rmask, rwant); break; }
            case QB_OPC_U_MASKED + 2: { const Mat m{mp[0], mp[1], mp[2], mp[3]}; bfly_masked<2>(a, m, rmask, rwant); break; }
            case QB_OPC_PERM + 0: perm_masked<0>(a, mp[1], mp[2], rmask, rwant); break;
            case QB_OPC_PERM + 1: perm_masked<1>(a, mp[1], mp[2], rmask, rwant); break;
            case QB_OPC_PERM + 2: perm_masked<2>(a, mp[1], mp[2], rmask, rwant); break;
            case QB_OPC_SWAP + 0: swap_masked<0>(a, rmask, rwant); break;
            case QB_OPC_SWAP + 1: swap_masked<1>(a, rmask, rwant); break;
            case QB_OPC_SWAP + 2: swap_masked<2>(a, rmask, rwant); break;
            case QB_OPC_PHASE: {
              const double2 ph = mp[0];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if ((uint32_t(e) & rmask) == rwant) a[e] = cmul(ph, a[e]);
              break;
            }
            case QB_OPC_LADDER: {
              const double2 *tb = s_tab + op->table_off;
              const double2 *F = reinterpret_cast<const double2 *>(op->F);
              double2 c = tb[(jb & 63u) ^ ((jb >> 3) & 7u)];
              if (hi_bits) c = cmul(c, tb[64 + (jb >> QB_LADDER_CHUNK)]);
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if ((uint32_t(e) & rmask) == rwant) a[e] = cmul(cmul(c, F[e]), a[e]);
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
    if (!(P.debug & 4))
    for (uint32_t j = tid; j < tileN; j += kFThreads)
      __stcs(psi + (base | (uint64_t(hi_off[j >> 3]) << 3) | (j & 7u)), tile[swz(j)]);
    __syncthreads();  // every read of this buffer is done before a later copy lands in it
    if (nbuf == 1 && more) issue_load(tn, tiles);
    if (nbuf == 2) cur ^= 1;
  }
}

size_t fused_smem_bytes(int K, int ntable, int nbuf, int nout_total) {
  return size_t(nbuf) * (size_t(1) << K) * sizeof(double2) + (kMaxLadders + 1) * sizeof(double2) +
         size_t(ntable) * sizeof(double2) + (size_t(1) << (K - 3)) * sizeof(uint32_t) +
         size_t(nout_total + 1) * sizeof(double2) + size_t(nout_total + 4) * sizeof(int32_t);
}

constexpr size_t kSmemLimit = 227 * 1024;
int g_sms = 0;

}  // namespace

cudaError_t fused_configure(int device) {
  static_assert(sizeof(QbOp) == 256, "QbOp layout");
  static_assert(sizeof(QbRound) % 4 == 0 && (QB_MAX_PASS_OPS * sizeof(QbOp)) % 16 == 0, "smem layout");
  cudaError_t err = cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, device);
  if (err != cudaSuccess) return err;
  err = cudaFuncSetAttribute(k_fused_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemLimit));
  if (err != cudaSuccess) return err;
  return cudaFuncSetAttribute(k_fused_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemLimit));
}

cudaError_t launch_fused_pass(double2 *psi, int nbits, const DevicePass &p, cudaStream_t st) {
  FusedParams P;
  P.psi = psi;
  P.nbits = nbits;
  P.desc = p.desc;
  if (p.desc.nops > QB_MAX_PASS_OPS || p.desc.nrounds > QB_MAX_PASS_ROUNDS) return cudaErrorInvalidValue;
  memcpy(P.ops, p.ops, size_t(p.desc.nops) * sizeof(QbOp));
  memcpy(P.rounds, p.rounds, size_t(p.desc.nrounds) * sizeof(QbRound));
  P.tables = p.tables;
  P.outph = p.outph;
  P.outbits = p.outbits;
  P.jbtab = p.jbtab;
  P.nbits_out = p.noutbits;
  static const int dbg = getenv("QCC_B200_FUSED_DEBUG") ? atoi(getenv("QCC_B200_FUSED_DEBUG")) : 0;
  static const int force_nbuf = getenv("QCC_B200_FUSED_NBUF") ? atoi(getenv("QCC_B200_FUSED_NBUF")) : 0;
  P.debug = dbg;
  static const int stagger = getenv("QCC_B200_FUSED_STAGGER_NS") ? atoi(getenv("QCC_B200_FUSED_STAGGER_NS")) : 0;
  P.stagger_ns = stagger;
  P.sms = g_sms;
  const int K = p.desc.K;
  if (K < 4 || K > QB_MAX_TILE_BITS || K > nbits) return cudaErrorInvalidValue;
  if (p.desc.nops > QB_MAX_PASS_OPS || p.desc.nrounds > QB_MAX_PASS_ROUNDS) return cudaErrorInvalidValue;
  const unsigned ntiles = 1u << (nbits - K);
  // Default: one CTA per tile, single buffer, two CTAs of 256 threads resident per SM (register
  // budget 128/thread: the 16 fp64 amplitude registers plus a complex 2x2 and ladder phases fit
  // without spilling; at 3 CTAs / 80 registers the spills and ladder-table misses cost more
  // than the extra CTA gains).  QCC_B200_FUSED_NBUF=2 selects the persistent two-deep cp.async
  // ring instead (one CTA per SM); measured slower (profiles/r01_fused_experiments.md).
  int nbuf = force_nbuf == 2 && fused_smem_bytes(K, p.desc.ntable, 2, p.desc.nout_total) <= kSmemLimit ? 2 : 1;
  const size_t smem = fused_smem_bytes(K, p.desc.ntable, nbuf, p.desc.nout_total);
  if (smem > kSmemLimit) return cudaErrorInvalidValue;
  P.nbuf = nbuf;
  // Persistent CTAs: 2 per SM, each walking tiles blockIdx.x, +grid, ... so that the per-CTA
  // staging (26 KiB of ladder tables, run offsets) is paid once per SM slot, not once per tile
  // (it was 3.5 of 16 ms per pass when every tile had its own CTA).
  static const int persist = getenv("QCC_B200_FUSED_PERSIST") ? atoi(getenv("QCC_B200_FUSED_PERSIST")) : 2;
  unsigned blocks = ntiles;
  if (persist > 0 && ntiles > unsigned(persist * g_sms)) blocks = unsigned(persist * g_sms);
  if (nbuf == 2) blocks = ntiles < unsigned(g_sms) ? ntiles : unsigned(g_sms);
  if (((1u << (K - 3)) % kFThreads) == 0)
    k_fused_pass<true><<<blocks, kFThreads, smem, st>>>(P);
  else
    k_fused_pass<false><<<blocks, kFThreads, smem, st>>>(P);
  return cudaGetLastError();
}

}  // namespace qb
