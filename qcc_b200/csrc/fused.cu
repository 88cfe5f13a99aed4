// fused.cu -- tile-resident fused pass for sm_100a: many gates per HBM sweep.
//
// One CTA of 256 threads = one TILE of 2^K amplitudes (K <= 13, default 12 = 64 KiB; see qb_types.h).
// Up to three CTAs are resident per SM (80 registers, 64 KiB tile + <= 9 KiB of ladder tables each), and
// because they start and finish at different times one streams its tile while the others compute.
// (A persistent CTA per SM with a ring of tile buffers was measured in round 1 and lost by 29 %: warps
// that move through the rounds in lockstep cover the fp64 and shared-memory latencies worse than warps
// of different CTAs in different phases -- DESIGN.md, experiment log.)
//
//   1. LOAD   the tile is gathered from HBM straight into (swizzled) shared memory with cp.async
//             (LDGSTS, 16 B per request, L2-only): 8 consecutive lanes fetch one 128-byte run, and the
//             planner pads the tile with the lowest free index bits so the runs of one tile are mostly
//             adjacent (a tile whose bits are 0..K-1 is one contiguous 64 KiB block).  With warp_io
//             (K >= 12) every warp copies exactly the sub-cube it will work on and waits for its own
//             copies only.  Meanwhile the ladder tables are staged and the per-tile ladder constants
//             computed.
//   2. ROUNDS each thread owns groups of 8 amplitudes that differ only in the round's 3 tile-local
//             bits, pulls one into 16 fp64 registers, runs every op of the round on registers and
//             writes it back: ONE shared-memory round trip for any number of gates on those 3 qubits,
//             plus every diagonal gate queued in between.  Rounds of one run keep each warp inside its
//             own sub-cube (warp sync only); a CTA barrier separates runs.  Code paths: the unrolled
//             Hadamard+ladder program (HL3, every round of a QFT), the UX program (straight-line
//             butterflies + a lean predicate-light interpreter: everything Grover, order finding, larose
//             and the supremacy circuits need), and the generic interpreter.
//   3. STORE  the last round writes its groups straight to HBM with streaming stores when it can
//             (st_direct); otherwise each warp stores the sub-cube of the last run from shared memory.
//             On a sharded state the store stage can carry an EXCHANGE EVENT (PushMap, kernels.h): every
//             amplitude is written to where the event's bit permutation of the distributed index puts
//             it -- this rank's alternate buffer or, through CUDA IPC peer mappings, another rank's, as
//             posted writes over NVLink -- so the all-to-all costs no sweep of its own.
//
// Shared-memory layout: the tile is stored XOR-swizzled, slot(j) = j ^ (fold(j >> 3) & 7)
// with fold(x) = x ^ x>>3 ^ x>>6 ^ x>>9, in 16-byte units.  The 16-byte bank group of j is
// then the XOR of its index bits taken mod 3, so (a) the linear LOAD/STORE pattern is
// conflict free, and (b) for ANY choice of round bits the planner can hand group-index
// bits 0..2 to one free local bit of each class (QbRound::qmap), which makes every
// quarter-warp of an LDS.128/STS.128 hit 8 distinct bank groups.  The swizzle is linear
// over XOR, so slot(base | spread(e)) = slot(base) ^ slot(spread(e)): one XOR per register, and
// every thread / iteration / round-bit term of an address is a host-computed constant.
//
// Arithmetic is fp64 on the CUDA cores (no tensor cores: 0.5-3 flop/B).  The op forms are chosen to
// minimise DFMA/DMUL count: real matrices (h, ry) cost 8 instead of 16 per pair, a matrix with a
// real and an imaginary column (h.v = H.S) 8 too; h followed by its cu1 ladder is ONE op (ULADDER):
// y' = (c x + d y) * phase.
//
// Phase ladders: the phase of an amplitude is T_a[lane] * T_b[q >> 5] * F[e]; T_* = host-built
// tables over the group number q, F = the 8 combinations of the round's own bits (constant bank);
// the product over partner bits outside the tile (one warp per ladder, once per tile) is folded into
// that tile's copy of T_b.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "kernels.h"

namespace qb {

namespace {

constexpr int kFThreads = 256;
// Host-precomputed per-round constants of the round programs (all derivable from QbRound / QbOp;
// kept out of the hot loop's uniform-datapath instruction stream).
struct RoundAux {
  uint32_t b[3];     // byte XOR that sets round bit k in a swizzled slot
  uint32_t pbi[(1 << (QB_MAX_TILE_BITS - 3)) / kFThreads];  // byte slot of group 256 * git
  uint32_t jbi[(1 << (QB_MAX_TILE_BITS - 3)) / kFThreads];  // tile-local base index of group 256 * git
  uint32_t pbt[8];   // byte XOR that group-number bit k (a thread-index bit) contributes to the swizzled slot
  uint32_t jbk[8];   // ... and the tile-local index bit it drives (1 << qmap[k])
  uint32_t gbk[8];   // ... and its index offset in units of 8 amplitudes (direct rounds; 0 for tile positions 0..2)
  uint32_t gb[3];    // direct rounds: index offset of round bit k, in units of 8 amplitudes
  uint32_t gji[(1 << (QB_MAX_TILE_BITS - 3)) / kFThreads];  // direct rounds: index offset of group 256 * git, same units
  uint32_t ta[3];    // HL3: byte offset of the three ladders' tables
  uint32_t ux;       // UX: leading uncontrolled U's on ascending round positions, run as straight-line code:
                     //     bits 0..1 = their number, bits 4+2k..5+2k = class at position k (0 none, 1 complex, 2 real,
                     //     3 column-imaginary)
  double s;          // HL3: product of the three Hadamard scales
};

struct FusedParams {
  double2 *psi;
  int nbits;
  QbPassDesc desc;
  const double2 *tables;
  const double2 *outph;
  const uint32_t *jbtab;
  int nlad;                                   // ladder ops of the pass ...
  // (any op of the pass can be one: the planner's QB_MAX_PASS_LADDERS counts LADDER items, and close_round
  // synthesises more for the short tails of Hadamard+ladder rounds -- order finding reaches 14 per pass)
  uint32_t lad_tab[QB_MAX_PASS_OPS];          // ... first entry of their lookup tables in `tables`
  uint32_t lad_ph[QB_MAX_PASS_OPS];           // ... first entry of their per-tile-constant tables in `outph`
  int push_on;   // 1: the store stage writes through `push` (exchange event fused into this pass)
  PushMap push;
  int stagger;   // experiment (QCC_B200_STAGGER=cycles): first-wave CTAs of resident slot j start j * stagger clocks late
  int nsm;
  int debug;  // timing experiments only -- 1: skip the op loop, 2: skip the rounds, 4: skip the store, 8: skip the load, 16: no round programs, 32: (unused), 64: CTA barrier after every round
  // per copy iteration i (thread t moves copy index t + 256 i, see QbPassDesc::ld_map / st_map): global
  // offset of copy index 256 i in units of 8 amplitudes, and the XOR that takes the byte slot of copy
  // index t to the byte slot of copy index t + 256 i
  // copy index bits 3..7 (thread bits): index offset in units of 8 amplitudes / byte XOR of the swizzled slot
  uint32_t ld_gbit[8], ld_sbit[8], st_gbit[8], st_sbit[8];
  uint32_t ld_goff[(1 << QB_MAX_TILE_BITS) / kFThreads];
  uint32_t ld_sxor[(1 << QB_MAX_TILE_BITS) / kFThreads];
  uint32_t st_goff[(1 << QB_MAX_TILE_BITS) / kFThreads];
  uint32_t st_sxor[(1 << QB_MAX_TILE_BITS) / kFThreads];
  RoundAux aux[QB_MAX_PASS_ROUNDS];
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

// m0 * x + m1 * y as one multiply and three fused multiply-adds per component (8 fp64 instructions
// per output instead of 10): the form the uncontrolled butterflies use.
__device__ __forceinline__ double2 mad2f(double2 m0, double2 x, double2 m1, double2 y) {
  double re = m0.x * x.x;
  re = fma(-m0.y, x.y, re);
  re = fma(m1.x, y.x, re);
  re = fma(-m1.y, y.y, re);
  double im = m0.x * x.y;
  im = fma(m0.y, x.x, im);
  im = fma(m1.x, y.y, im);
  im = fma(m1.y, y.x, im);
  return make_double2(re, im);
}

// real m0, m1
__device__ __forceinline__ double2 mad2r(double m0, double2 x, double m1, double2 y) {
  return make_double2(m0 * x.x + m1 * y.x, m0 * x.y + m1 * y.y);
}

struct Mat {
  double2 a, b, c, d;
};

// shared-memory accesses by 32-bit shared address (base in a uniform register + per-thread offset)
__device__ __forceinline__ double2 lds128(uint32_t sa) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(sa));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t sa, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(sa), "d"(v.x), "d"(v.y) : "memory");
}

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
      a[e] = mad2f(m.a, x, m.b, y);
      a[e | (1 << TP)] = mad2f(m.c, x, m.d, y);
    }
  }
}

// first column real (a, c), second column imaginary (i b, i d):  x' = a x + i b y,  y' = c x + i d y
template <int TP>
__device__ __forceinline__ void bfly_colimag(double2 (&a)[8], double ma, double mb, double mc, double md) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    const double2 x = a[e], y = a[e | (1 << TP)];
    a[e] = make_double2(fma(-mb, y.y, ma * x.x), fma(mb, y.x, ma * x.y));
    a[e | (1 << TP)] = make_double2(fma(-md, y.y, mc * x.x), fma(md, y.x, mc * x.y));
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

// butterfly on the pairs selected by bit e of sel (sel is uniform: a precomputed round-level predicate)
template <int TP>
__device__ __forceinline__ void bfly_sel(double2 (&a)[8], const Mat &m, uint32_t sel) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    if ((sel >> e) & 1u) {
      const double2 x = a[e], y = a[e | (1 << TP)];
      a[e] = mad2f(m.a, x, m.b, y);
      a[e | (1 << TP)] = mad2f(m.c, x, m.d, y);
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

// Parity-controlled swap (QB_K_PARSWAP): pair e of position TP is swapped when bit e of sel is set
// (sel = parity table of the round's own control bits, complemented when the parity of the control
// bits outside the round -- tile-local and outside-tile -- xor the constant flip is odd).
template <int TP>
__device__ __forceinline__ void parswap(double2 (&a)[8], uint32_t sel) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e & (1 << TP)) continue;
    const bool sw = (sel >> e) & 1u;
    const double2 x = a[e], y = a[e | (1 << TP)];
    a[e] = make_double2(sw ? y.x : x.x, sw ? y.y : x.y);
    a[e | (1 << TP)] = make_double2(sw ? x.x : y.x, sw ? x.y : y.y);
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

// Direct store (QbPassDesc::st_direct): the last round of a pass writes its groups straight back to HBM, so
// the tile skips one shared-memory write + read and the barrier before the store.  gp = psi + tile base; the
// thread's part of the index (its group-number bits scattered to the tile bits) is built once per round, the
// iteration part and the round bits come from RoundAux.  (Direct LOADS of the first round were measured in
// round 1 and lost 7-17 %: only 8 loads per thread in flight instead of the cp.async copy's 16.)
// IO (template parameter of the round programs): 2 = store directly, 0 = through shared memory.

// ---- round program HL3: three Hadamard+ladder stages on round positions 0, 1, 2 ---------------
// The shape of every round of a QFT (circuit.py:320-328): h(b0) + ladder(b0), h(b1) + ladder(b1),
// h(b2) + ladder(b2).  Fully unrolled, no op decode: per group 8 LDS.128, three table lookups (two
// reads each: one per lane, one broadcast), ~130 fp64 instructions, 8 STS.128.  The three
// 1/sqrt(2) factors are applied once (s = r0 r1 r2: the y' path of stage 0 takes it inside its
// phase, the x' path as one multiply).
// UPPER: the ladder partners inside the round are all above their pivot, so stage 1 has one
// non-trivial in-round phase and stage 2 none (true for the QFT; otherwise all of F is used).
template <bool UPPER, bool FULL, bool SCALED, int IO = 0, int THREADS = kFThreads>
__device__ __forceinline__ void round_hl3(const uint32_t tile_sa, const uint32_t tab_sa,
                                          const uint32_t *__restrict__ jbt, const double2 *__restrict__ F0,
                                          const double2 *__restrict__ F1, const double2 *__restrict__ F2,
                                          const QbRound *__restrict__ R, const RoundAux *__restrict__ X,
                                          const uint32_t ngroups, const uint32_t tid, double2 *const gp,
                                          const uint64_t g_t) {
  const double s = X->s;
  const uint32_t giters = FULL ? (ngroups / THREADS) : ((ngroups + THREADS - 1) / THREADS);
  const uint32_t b0 = X->b[0], b1 = X->b[1], b2 = X->b[2];
  // ladder tables: T_a[lane] (fixed per thread), T_b[q >> 5] (uniform per warp)
  const uint32_t lane16 = (tid & 31u) << 4;
  const uint32_t ta0 = tab_sa + X->ta[0] + lane16;
  const uint32_t ta1 = tab_sa + X->ta[1] + lane16;
  const uint32_t ta2 = tab_sa + X->ta[2] + lane16;
  // Group q = tid + 256 git -> swizzled byte slot of its base index (the bits of q scattered to
  // qmap[]).  The slot is linear over XOR in q, so the per-thread part is built once per round
  // from tid and the per-iteration part is uniform: no table load at the head of the dependency
  // chain of every group.
  uint32_t pb_t = 0;
  if (FULL) {
#pragma unroll
    for (int k = 0; k < 8; ++k) pb_t ^= ((tid >> k) & 1u) ? X->pbt[k] : 0u;
  }
#pragma unroll 1
  for (uint32_t git = 0; git < giters; ++git) {
    const uint32_t q = git * THREADS + tid;
    const uint32_t sub = THREADS == kFThreads ? git : (q >> 8);  // which block of 256 groups (uniform at 256 threads)
    uint32_t pb;  // base slot in bytes
    if (FULL) {
      pb = pb_t ^ X->pbi[sub];
    } else {
      if (q >= ngroups) break;
      pb = (__ldg(jbt + q) >> 12) & 0xffff0u;
    }
    double2 a[8];
    double2 *gq = nullptr;
    if (IO != 0) gq = gp + (g_t + (uint64_t(X->gji[sub]) << 3));
#pragma unroll
    for (int e = 0; e < 8; ++e)
      a[e] = lds128(tile_sa + (pb ^ ((e & 1) ? b0 : 0u) ^ ((e & 2) ? b1 : 0u) ^ ((e & 4) ? b2 : 0u)));
    const uint32_t tboff = (32u + (q >> 5)) << 4;
    // stage 0: pairs (e, e|1)
    {
      double2 c0 = cmul(lds128(ta0), lds128(ta0 - lane16 + tboff));
      if (SCALED) {
        c0.x *= s;
        c0.y *= s;
      }
      const double2 p3 = cmul(c0, F0[3]), p5 = cmul(c0, F0[5]), p7 = cmul(c0, F0[7]);
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const double2 x = a[e], y = a[e | 1];
        const double2 ph = e == 0 ? c0 : (e == 2 ? p3 : (e == 4 ? p5 : p7));
        const double2 d = make_double2(x.x - y.x, x.y - y.y);
        a[e] = SCALED ? make_double2(s * (x.x + y.x), s * (x.y + y.y)) : make_double2(x.x + y.x, x.y + y.y);
        a[e | 1] = cmul(d, ph);
      }
    }
    // stage 1: pairs (e, e|2)
    {
      const double2 c1 = cmul(lds128(ta1), lds128(ta1 - lane16 + tboff));
      double2 p2 = c1, p3, p6, p7;
      if (UPPER) {
        p3 = c1;
        p6 = cmul(c1, F1[6]);
        p7 = p6;
      } else {
        p3 = cmul(c1, F1[3]);
        p6 = cmul(c1, F1[6]);
        p7 = cmul(c1, F1[7]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = (k & 1) | ((k & 2) << 1);
        const double2 x = a[e], y = a[e | 2];
        const double2 ph = k == 0 ? p2 : (k == 1 ? p3 : (k == 2 ? p6 : p7));
        const double2 d = make_double2(x.x - y.x, x.y - y.y);
        a[e] = make_double2(x.x + y.x, x.y + y.y);
        a[e | 2] = cmul(d, ph);
      }
    }
    // stage 2: pairs (e, e|4)
    {
      const double2 c2 = cmul(lds128(ta2), lds128(ta2 - lane16 + tboff));
      double2 p4 = c2, p5 = c2, p6 = c2, p7 = c2;
      if (!UPPER) {
        p5 = cmul(c2, F2[5]);
        p6 = cmul(c2, F2[6]);
        p7 = cmul(c2, F2[7]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double2 x = a[e], y = a[e | 4];
        const double2 ph = e == 0 ? p4 : (e == 1 ? p5 : (e == 2 ? p6 : p7));
        const double2 d = make_double2(x.x - y.x, x.y - y.y);
        a[e] = make_double2(x.x + y.x, x.y + y.y);
        a[e | 4] = cmul(d, ph);
      }
    }
    if (IO & 2) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        __stcs(gq + (uint64_t(((e & 1) ? X->gb[0] : 0u) + ((e & 2) ? X->gb[1] : 0u) + ((e & 4) ? X->gb[2] : 0u)) << 3), a[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        sts128(tile_sa + (pb ^ ((e & 1) ? b0 : 0u) ^ ((e & 2) ? b1 : 0u) ^ ((e & 4) ? b2 : 0u)), a[e]);
    }
  }
}

// ---- round program UX: uncontrolled butterflies and parity swaps only ------------------------
// Every round of larose_benchmark.py:47-54 after scheduling (h.v on three qubits; the cx fan-in onto
// qubit 0 as ONE parity swap), and the single-qubit layers of supremacy-style circuits.  The leading
// U's (one per round position, ascending -- the planner's order) are straight-line code selected by
// uniform branches: no loop-carried register assignment, so no register moves and no decode.  What
// follows them (the parity swap, a second U on a position) goes through a lean interpreter without
// predicates, matrices straight from the constant bank through uniform registers.
template <int TP>
__device__ __forceinline__ void ux_stage(double2 (&a)[8], const QbOp *__restrict__ op, const uint32_t cls) {
  const double2 *mp = reinterpret_cast<const double2 *>(op->m);
  if (cls == 1) {
    const Mat m{mp[0], mp[1], mp[2], mp[3]};
    bfly_all<TP, false>(a, m);
  } else if (cls == 2) {
    Mat m;
    m.a.x = mp[0].x; m.b.x = mp[1].x; m.c.x = mp[2].x; m.d.x = mp[3].x;
    bfly_all<TP, true>(a, m);
  } else {
    bfly_colimag<TP>(a, mp[0].x, mp[1].y, mp[2].x, mp[3].y);
  }
}

template <bool FULL, int IO = 0, int THREADS = kFThreads>
__device__ __forceinline__ void round_ux(const uint32_t tile_sa, const uint32_t *__restrict__ jbt,
                                         const QbOp *__restrict__ o, const int nops,
                                         const QbRound *__restrict__ R, const RoundAux *__restrict__ X,
                                         const uint32_t ngroups, const uint32_t tid, const uint64_t base,
                                         double2 *const gp, const uint64_t g_t, const uint32_t tab_sa) {
  const uint32_t giters = FULL ? (ngroups / THREADS) : ((ngroups + THREADS - 1) / THREADS);
  const uint32_t b0 = X->b[0], b1 = X->b[1], b2 = X->b[2];
  const uint32_t ux = X->ux;
  const int nu = int(ux & 3u);
  const uint32_t c0 = (ux >> 4) & 3u, c1 = (ux >> 6) & 3u, c2 = (ux >> 8) & 3u;
  const QbOp *o0 = o, *o1 = o + (c0 ? 1 : 0), *o2 = o1 + (c1 ? 1 : 0);
  const uint32_t uni = (c0 == c1 && c1 == c2) ? c0 : 0u;  // all three stages present and of one matrix class
  uint32_t jb_t = 0, pb_t = 0;
  if (FULL) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool bit = (tid >> k) & 1u;
      jb_t |= bit ? X->jbk[k] : 0u;
      pb_t ^= bit ? X->pbt[k] : 0u;
    }
  }
#pragma unroll 1
  for (uint32_t git = 0; git < giters; ++git) {
    const uint32_t q = git * THREADS + tid;
    const uint32_t sub = THREADS == kFThreads ? git : (q >> 8);
    uint32_t pb, jb;
    if (FULL) {
      pb = pb_t ^ X->pbi[sub];
      jb = jb_t | X->jbi[sub];
    } else {
      if (q >= ngroups) break;
      const uint32_t w = __ldg(jbt + q);
      jb = w & 0xffffu;
      pb = (w >> 12) & 0xffff0u;
    }
    double2 a[8];
    double2 *gq = nullptr;
    if (IO != 0) gq = gp + (g_t + (uint64_t(X->gji[sub]) << 3));
#pragma unroll
    for (int e = 0; e < 8; ++e)
      a[e] = lds128(tile_sa + (pb ^ ((e & 1) ? b0 : 0u) ^ ((e & 2) ? b1 : 0u) ^ ((e & 4) ? b2 : 0u)));
    if (uni == 3) {
      // all three positions carry a column-imaginary U (every round of larose): one straight-line block,
      // no control-flow joins between the stages, so no register moves to line results up
      const double2 *m0 = reinterpret_cast<const double2 *>(o0->m);
      const double2 *m1 = reinterpret_cast<const double2 *>(o1->m);
      const double2 *m2 = reinterpret_cast<const double2 *>(o2->m);
      bfly_colimag<0>(a, m0[0].x, m0[1].y, m0[2].x, m0[3].y);
      bfly_colimag<1>(a, m1[0].x, m1[1].y, m1[2].x, m1[3].y);
      bfly_colimag<2>(a, m2[0].x, m2[1].y, m2[2].x, m2[3].y);
    } else if (uni == 1) {
      ux_stage<0>(a, o0, 1);
      ux_stage<1>(a, o1, 1);
      ux_stage<2>(a, o2, 1);
    } else if (uni == 2) {
      ux_stage<0>(a, o0, 2);
      ux_stage<1>(a, o1, 2);
      ux_stage<2>(a, o2, 2);
    } else {
      if (c0) ux_stage<0>(a, o0, c0);
      if (c1) ux_stage<1>(a, o1, c1);
      if (c2) ux_stage<2>(a, o2, c2);
    }
#pragma unroll 1
    for (int oi = nu; oi < nops; ++oi) {
      const QbOp *op = o + oi;
      const int opc = int(uint32_t(op->kind) >> 24);
      // controls outside the tile: uniform per CTA (a PARSWAP's gmask is a parity, handled below)
      if (!(opc >= QB_OPC_PARSWAP && opc < QB_OPC_PARSWAP + 3) && (base & op->gmask) != op->gwant) continue;
      if (opc < QB_OPC_U_ALL || opc == QB_OPC_LADDER) {
        // phase ladder (optionally fused with the butterfly on its pivot): phase = T_a[lane] * T_b[q >> 5]
        // (the tile's copy of T_b carries the per-tile constant) * F[e]
        const uint32_t tb = tab_sa + (uint32_t(op->table_off) << 4);
        const double2 c = cmul(lds128(tb + ((q & 31u) << 4)), lds128(tb + ((32u + (q >> QB_LADDER_LANE_BITS)) << 4)));
        const double2 *F = reinterpret_cast<const double2 *>(op->F);
        const double2 *mp = reinterpret_cast<const double2 *>(op->m);
        if (opc == QB_OPC_LADDER) {
          if ((jb & op->lmask) != op->lwant) continue;
          const uint32_t rmask = op->rmask, rwant = op->rwant;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if ((uint32_t(e) & rmask) == rwant) a[e] = cmul(cmul(c, F[e]), a[e]);
        } else if (opc < 3) {
          const Mat m{mp[0], mp[1], mp[2], mp[3]};
          if (opc == 0) uladder<0, false>(a, m, c, F);
          else if (opc == 1) uladder<1, false>(a, m, c, F);
          else uladder<2, false>(a, m, c, F);
        } else if (opc < 6) {
          Mat m;
          m.a.x = mp[0].x; m.b.x = mp[1].x; m.c.x = mp[2].x; m.d.x = mp[3].x;
          if (opc == 3) uladder<0, true>(a, m, c, F);
          else if (opc == 4) uladder<1, true>(a, m, c, F);
          else uladder<2, true>(a, m, c, F);
        } else {
          if (opc == 6) hladder<0>(a, mp[0].x, c, F);
          else if (opc == 7) hladder<1>(a, mp[0].x, c, F);
          else hladder<2>(a, mp[0].x, c, F);
        }
      } else if (opc >= QB_OPC_U_CI) {
        const uint32_t tp = uint32_t(opc - QB_OPC_U_CI);
        if (tp == 0) ux_stage<0>(a, op, 3);
        else if (tp == 1) ux_stage<1>(a, op, 3);
        else ux_stage<2>(a, op, 3);
      } else if ((opc >= QB_OPC_SWAP && opc <= QB_OPC_PHASE) || (opc >= QB_OPC_U_MASKED && opc < QB_OPC_U_MASKED + 3)) {
        // controlled swap (cx, ccx, ...) / controlled phase / controlled 2x2: controls outside the tile
        // are uniform per CTA, tile bits outside the round one test per group, round bits precomputed
        // (op->flags)
        if ((jb & op->lmask) != op->lwant) continue;
        const uint32_t sel = uint32_t(op->flags);
        if (opc < QB_OPC_SWAP) {
          const double2 *mp = reinterpret_cast<const double2 *>(op->m);
          const Mat m{mp[0], mp[1], mp[2], mp[3]};
          if (opc == QB_OPC_U_MASKED + 0) bfly_sel<0>(a, m, sel);
          else if (opc == QB_OPC_U_MASKED + 1) bfly_sel<1>(a, m, sel);
          else bfly_sel<2>(a, m, sel);
        } else if (opc == QB_OPC_PHASE) {
          const double2 ph = *reinterpret_cast<const double2 *>(op->m);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if ((sel >> e) & 1u) a[e] = cmul(ph, a[e]);
        } else if (opc == QB_OPC_SWAP + 0) {
          parswap<0>(a, sel);
        } else if (opc == QB_OPC_SWAP + 1) {
          parswap<1>(a, sel);
        } else {
          parswap<2>(a, sel);
        }
      } else if (opc >= QB_OPC_PARSWAP) {
        const uint32_t odd = (uint32_t(__popcll(base & op->gmask)) + uint32_t(__popc(jb & op->lmask)) + op->rwant) & 1u;
        const uint32_t sel = op->lwant ^ (0u - odd);
        if (opc == QB_OPC_PARSWAP + 0) parswap<0>(a, sel);
        else if (opc == QB_OPC_PARSWAP + 1) parswap<1>(a, sel);
        else parswap<2>(a, sel);
      } else {
        const uint32_t k = uint32_t(opc - QB_OPC_U_ALL);  // 0..2 complex, 3..5 real
        const uint32_t cls = k < 3 ? 1u : 2u;
        const uint32_t tp = k < 3 ? k : k - 3;
        if (tp == 0) ux_stage<0>(a, op, cls);
        else if (tp == 1) ux_stage<1>(a, op, cls);
        else ux_stage<2>(a, op, cls);
      }
    }
    if (IO & 2) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        __stcs(gq + (uint64_t(((e & 1) ? X->gb[0] : 0u) + ((e & 2) ? X->gb[1] : 0u) + ((e & 4) ? X->gb[2] : 0u)) << 3), a[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        sts128(tile_sa + (pb ^ ((e & 1) ? b0 : 0u) ^ ((e & 2) ? b1 : 0u) ^ ((e & 4) ? b2 : 0u)), a[e]);
    }
  }
}

#define QB_DISPATCH_TP(tp, CALL0, CALL1, CALL2) \
  do {                                          \
    if ((tp) == 0) { CALL0; }                   \
    else if ((tp) == 1) { CALL1; }              \
    else { CALL2; }                             \
  } while (0)

__device__ __forceinline__ void cp_async16(uint32_t sa, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Between two rounds.  Rounds of one run (QbRound::nobar, planner.cc assign_group_maps) keep every
// warp inside its own sub-cube of the tile, so only the lanes of a warp have to see each other's
// stores; otherwise the whole CTA meets.
__device__ __forceinline__ void round_sync(bool warp_only) {
  if (warp_only) __syncwarp();
  else __syncthreads();
}

// One round that has a round program (QbRound::prog != GENERIC).  (Specialising this code on the round number, so
// that every per-round constant sits at a compile-time address of the parameter block, was measured in round 2:
// no gain on QFT-30 -- fp64 instructions take their constants through uniform registers either way -- and 8 %
// slower on larose-28 from the code growth.)
template <bool FULL, int THREADS>
__device__ __forceinline__ void program_round(const FusedParams &P, const int r, const int ob, const int oe,
                                              const uint32_t tile_sa, const uint32_t tab_sa,
                                              const uint32_t ngroups, const uint32_t tid, const uint64_t base,
                                              const bool direct) {
  const QbRound *R = P.rounds + r;
  const RoundAux *X = P.aux + r;
  const QbOp *o = P.ops + ob;
  const double2 *F0 = reinterpret_cast<const double2 *>(o[0].F);
  const double2 *F1 = reinterpret_cast<const double2 *>(o[1].F);
  const double2 *F2 = reinterpret_cast<const double2 *>(o[2].F);
  const uint32_t *jbt = P.jbtab + (size_t(r) << P.desc.ngroups_log2);
  const bool upper = R->prog == QB_PROG_HL3U;
  const bool st = FULL && direct && r + 1 == P.desc.nrounds && P.desc.st_direct;
  double2 *const gp = P.psi + base;
  uint64_t g_t = 0;
  if (st) {
    // the thread's group-number bits (0..7) scattered to the index bits they drive (lanes 0..7 sit on
    // tile positions 0..2 = index bits 0..2; the other terms are multiples of 8 amplitudes)
    uint32_t g8 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool bit = (tid >> k) & 1u;
      if (k < 3) g_t |= bit ? uint64_t(X->jbk[k]) : 0u;
      else g8 += bit ? X->gbk[k] : 0u;
    }
    g_t |= uint64_t(g8) << 3;
  }
#define QB_ROUND_IO(CALL)                                    \
  do {                                                       \
    if (!FULL || THREADS != kFThreads || !st) { constexpr int IO = 0; CALL; } \
    else { constexpr int IO = FULL && THREADS == kFThreads ? 2 : 0; CALL; }   \
  } while (0)
  if (R->prog == QB_PROG_UX) {
    QB_ROUND_IO((round_ux<FULL, IO, THREADS>(tile_sa, jbt, o, oe - ob, R, X, ngroups, tid, base, gp, g_t, tab_sa)));
  } else if (X->s != 1.0) {
    if (upper) QB_ROUND_IO((round_hl3<true, FULL, true, IO, THREADS>(tile_sa, tab_sa, jbt, F0, F1, F2, R, X, ngroups, tid, gp, g_t)));
    else QB_ROUND_IO((round_hl3<false, FULL, true, IO, THREADS>(tile_sa, tab_sa, jbt, F0, F1, F2, R, X, ngroups, tid, gp, g_t)));
  } else {
    if (upper) QB_ROUND_IO((round_hl3<true, FULL, false, IO, THREADS>(tile_sa, tab_sa, jbt, F0, F1, F2, R, X, ngroups, tid, gp, g_t)));
    else QB_ROUND_IO((round_hl3<false, FULL, false, IO, THREADS>(tile_sa, tab_sa, jbt, F0, F1, F2, R, X, ngroups, tid, gp, g_t)));
  }
#undef QB_ROUND_IO
}

// FULL: the tile has a multiple of 256 groups (K >= 11), so the group loop has a trip count that
// is uniform across the CTA and needs no bounds test.  That matters beyond the saved compare: with a
// thread-dependent loop condition the compiler must treat the op loop inside as divergent and keeps
// the op index -- and with it every descriptor load -- in vector registers (vector-indexed LDC,
// vector compares and branches per op); with a uniform trip count the decode runs on the uniform
// datapath.
// FAST: every round of the pass has a round program, so the op interpreter is not compiled in; the
// kernel then fits 80 registers and a third CTA per SM (24 instead of 16 warps to hide the
// shared-memory and fp64 latencies of the rounds).
// PUSH: the store stage carries an exchange event (its own instantiation, so the plain pass pays nothing for it).
template <bool FULL, bool FAST, bool PUSH>
__global__ void __launch_bounds__(kFThreads, FAST ? 3 : 2) k_fused_pass(const __grid_constant__ FusedParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = P.desc.K;
  const uint32_t tileN = 1u << K;
  double2 *tile = reinterpret_cast<double2 *>(smem_raw);                // 2^K
  double2 *s_tab = tile + tileN;                                        // ntable
  uint32_t *s_active = reinterpret_cast<uint32_t *>(s_tab + P.desc.ntable);  // 4 words: ops whose
                                                                        // outside-tile predicate holds for this tile
  double2 **s_out = reinterpret_cast<double2 **>(s_active + 4);         // push: destination buffer of every rank
  if (PUSH && threadIdx.x < kPushMaxRanks) s_out[threadIdx.x] = P.push.out[threadIdx.x];
  const QbOp *s_ops = P.ops;        // constant bank
  const QbRound *s_rounds = P.rounds;
  const uint32_t tid = threadIdx.x;
  double2 *__restrict__ psi = P.psi;


  // Tile <-> HBM addressing.  Thread t moves tile-local indices j = t + 256 i.  Both the global
  // offset of j (its bits scattered to tile_bits[]) and its swizzled slot are linear over XOR and
  // t, 256 i have no bit in common, so address(j) = address(t) (+ or ^) address(256 i): one
  // per-thread term computed here, one per-iteration term from the kernel parameters -- an add, an
  // XOR and the copy itself per 16 bytes.
  // With warp_io the copy index is not the tile-local index: its warp bits (5..7) select the per-warp
  // sub-cube of the first (load) / last (store) run of rounds, so a warp moves what it computes on.
  // (per-bit terms precomputed on the host: positions 0..2 keep their slot bits under the swizzle XOR)
  uint32_t g8_ld = 0, g8_st = 0;
  uint32_t s_ld = swz(tid & 7u) << 4, s_st = s_ld;
#pragma unroll
  for (int k = 3; k < 8; ++k) {
    const bool bit = (tid >> k) & 1u;
    g8_ld += bit ? P.ld_gbit[k] : 0u;
    g8_st += bit ? P.st_gbit[k] : 0u;
    s_ld ^= bit ? P.ld_sbit[k] : 0u;
    s_st ^= bit ? P.st_sbit[k] : 0u;
  }
  const uint64_t g_ld = (uint64_t(g8_ld) << 3) | (tid & 7u), g_st = (uint64_t(g8_st) << 3) | (tid & 7u);
  const bool warp_io = P.desc.warp_io != 0;
  const uint32_t io_iters = tileN > kFThreads ? tileN / kFThreads : 1u;
  const bool io_on = tid < tileN;
  const uint32_t tile_sa = uint32_t(__cvta_generic_to_shared(tile));
  const uint32_t tab_sa = uint32_t(__cvta_generic_to_shared(s_tab));

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
  // direct first / last round: no copy-in / copy-out of the tile (FAST or not, the first / last round
  // then has a round program; the debug switches that bypass the programs turn it off on the host)
  const bool st_direct = !PUSH && FULL && P.desc.st_direct;
  auto issue_load = [&](uint64_t b) {
    if (!(P.debug & 8) && io_on) {
      const double2 *src = psi + (b | g_ld);
      if (io_iters == 16) {  // K = 12: constant-bank operands with immediate addresses
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i)
          cp_async16(tile_sa + (s_ld ^ P.ld_sxor[i]), src + (uint64_t(P.ld_goff[i]) << 3));
      } else {
#pragma unroll 4
        for (uint32_t i = 0; i < io_iters; ++i)
          cp_async16(tile_sa + (s_ld ^ P.ld_sxor[i]), src + (uint64_t(P.ld_goff[i]) << 3));
      }
    }
    cp_async_commit();
  };

  const uint32_t ngroups = tileN >> 3;
  const int gbits = K - QB_ROUND_BITS;
  uint64_t base = blockIdx.x < ntiles ? tile_base(blockIdx.x) : 0;
  if (P.stagger > 0 && blockIdx.x >= uint32_t(P.nsm) && blockIdx.x < 3u * uint32_t(P.nsm)) {
    // the CTAs of the first wave would all load, then all compute, then all store together, and their
    // successors inherit the phase: the second and third resident CTA of an SM start a fraction of a tile late
    const long long t0 = clock64(), wait = (long long)(P.stagger) * (long long)(blockIdx.x / uint32_t(P.nsm));
    while (clock64() - t0 < wait) {
    }
  }
  if (blockIdx.x < ntiles) issue_load(base);
  // ---- STAGE (once per CTA, behind the first tile's copy): the per-lane halves T_a of the ladder tables ->
  // shared memory.  The T_b halves are written per tile below (with the per-tile constant folded in), by
  // other threads: staging them here as well would need a barrier in between.
  for (uint32_t x = tid; x < (uint32_t(P.nlad) << QB_LADDER_LANE_BITS); x += kFThreads) {
    const uint32_t e = P.lad_tab[x >> QB_LADDER_LANE_BITS] + (x & ((1u << QB_LADDER_LANE_BITS) - 1u));
    s_tab[e] = __ldg(P.tables + e);
  }
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint32_t tn = t + gridDim.x;
    const bool more = tn < ntiles;
    // which ops apply to this tile at all (their controls outside the tile): one bit per op
    if (!FAST && tid < 128) {
      const bool on = int(tid) < P.desc.nops && ((s_ops[tid].kind & 0xff) == QB_K_PARSWAP ||  // its gmask is a parity
                                                  (base & s_ops[tid].gmask) == s_ops[tid].gwant);
      const uint32_t bal = __ballot_sync(0xffffffffu, on);
      if ((tid & 31u) == 0) s_active[tid >> 5] = bal;
    }
    // per-tile constants of the phase ladders, folded into this tile's copy of T_b (<= 32 entries per ladder)
    // so the hot loop needs one multiply and one dependent shared-memory read less per op.  One thread per
    // (ladder, T_b entry): three table entries selected by the fields of the tile number (host-built,
    // planner.cc build_ladder_tables; L2-resident: 3 x 64 entries per ladder at 30 qubits), three complex
    // multiplies, no shuffles, no serial chain -- it overlaps the wait for the tile's data.
    {
      const uint32_t f0 = t & ((1u << P.desc.lad_w[0]) - 1u);
      const uint32_t f1 = (1u << P.desc.lad_w[0]) + ((t >> P.desc.lad_w[0]) & ((1u << P.desc.lad_w[1]) - 1u));
      const uint32_t f2 = (1u << P.desc.lad_w[0]) + (1u << P.desc.lad_w[1]) + (t >> (P.desc.lad_w[0] + P.desc.lad_w[1]));
      const int nb_log2 = gbits > QB_LADDER_LANE_BITS ? gbits - QB_LADDER_LANE_BITS : 0;
      for (uint32_t x = tid; x < (uint32_t(P.nlad) << nb_log2); x += kFThreads) {
        const uint32_t l = x >> nb_log2;
        const double2 *ph = P.outph + P.lad_ph[l];
        const uint32_t e = P.lad_tab[l] + (1u << QB_LADDER_LANE_BITS) + (x & ((1u << nb_log2) - 1u));
        const double2 c = cmul(cmul(__ldg(ph + f0), __ldg(ph + f1)), __ldg(ph + f2));
        s_tab[e] = cmul(__ldg(P.tables + e), c);
      }
    }
    if (warp_io) {
      // tables, per-tile constants and the active mask are CTA-wide; the tile data is not: every warp
      // waits for its own copies only (they cover exactly the sub-cube it works on until the first
      // CTA barrier between rounds)
      __syncthreads();
      cp_async_wait<0>();
      __syncwarp();
    } else {
      cp_async_wait<0>();
      __syncthreads();
    }

    // ---- ROUNDS ------------------------------------------------------------------------
    for (int r = 0; r < ((P.debug & 2) ? 0 : P.desc.nrounds); ++r) {
      const QbRound *R = s_rounds + r;
      const int ob = R->op_begin, oe = (P.debug & 1) ? R->op_begin : R->op_end;
      const uint32_t *jbt = P.jbtab + (size_t(r) << P.desc.ngroups_log2);
      if (FAST || (R->prog != QB_PROG_GENERIC && !(P.debug & (1 | 16)))) {
        program_round<FULL, kFThreads>(P, r, ob, oe, tile_sa, tab_sa, ngroups, tid, base, true);
        round_sync((R->nobar || (warp_io && r + 1 == P.desc.nrounds)) && !(P.debug & 64));
        continue;
      }
      if (!FAST) {
      const uint32_t d0 = swz(1u << R->rbit[0]);
      const uint32_t d1 = swz(1u << R->rbit[1]);
      const uint32_t d2 = swz(1u << R->rbit[2]);
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
        const uint64_t active_lo = (uint64_t(s_active[1]) << 32) | s_active[0];
        const uint64_t active_hi = (uint64_t(s_active[3]) << 32) | s_active[2];
#pragma unroll 1
        for (int oi = ob; oi < oe; ++oi) {
          if (!(((oi < 64 ? active_lo : active_hi) >> (oi & 63)) & 1)) continue;  // uniform per tile
          const QbOp *op = s_ops + oi;
          const double2 *mp = reinterpret_cast<const double2 *>(op->m);
          const int opc = int(uint32_t(op->kind) >> 24);
          if (opc < QB_OPC_U_ALL) {
            // ULADDER family: uncontrolled butterfly on the pivot + the pivot's phase ladder
            const double2 *tb = s_tab + op->table_off;
            const double2 *F = reinterpret_cast<const double2 *>(op->F);  // constant bank
            // T_b already carries the per-tile constant
            const double2 c = cmul(tb[q & 31u], tb[32u + (q >> QB_LADDER_LANE_BITS)]);
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
          if (opc >= QB_OPC_U_CI) {
            if (opc == QB_OPC_U_CI + 0) bfly_colimag<0>(a, mp[0].x, mp[1].y, mp[2].x, mp[3].y);
            else if (opc == QB_OPC_U_CI + 1) bfly_colimag<1>(a, mp[0].x, mp[1].y, mp[2].x, mp[3].y);
            else bfly_colimag<2>(a, mp[0].x, mp[1].y, mp[2].x, mp[3].y);
            continue;
          }
          if (opc >= QB_OPC_PARSWAP) {
            const uint32_t odd = (uint32_t(__popcll(base & op->gmask)) + uint32_t(__popc(jb & op->lmask)) + op->rwant) & 1u;
            const uint32_t sel = op->lwant ^ (0u - odd);
            if (opc == QB_OPC_PARSWAP + 0) parswap<0>(a, sel);
            else if (opc == QB_OPC_PARSWAP + 1) parswap<1>(a, sel);
            else parswap<2>(a, sel);
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
              const double2 c = cmul(tb[q & 31u], tb[32u + (q >> QB_LADDER_LANE_BITS)]);
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
      round_sync((R->nobar || (warp_io && r + 1 == P.desc.nrounds)) && !(P.debug & 64));
      }
    }

    // ---- STORE ---------------------------------------------------------------------------
    if (PUSH && io_on) {
      // the tile goes to where the exchange event puts it: distributed index = sigma(rank, base) | sigma(thread
      // term) | sigma(iteration term); its top bits name the destination rank, the rest the slot in that
      // rank's alternate buffer (posted writes over NVLink for the other ranks)
      uint64_t sb = (base & ~P.push.moved_mask) | P.push.rank_term | g_st;
#pragma unroll
      for (int k = 0; k < kPushMaxMoved; ++k)
        if (k < P.push.nmoved) sb |= ((base >> P.push.src[k]) & 1) << P.push.dst[k];
      const uint64_t lmask = (uint64_t(1) << P.push.nl) - 1;
#pragma unroll 4
      for (uint32_t i = 0; i < io_iters; ++i) {
        const uint64_t a = sb + (uint64_t(P.st_goff[i]) << 3);
        __stcs(s_out[a >> P.push.nl] + (a & lmask), lds128(tile_sa + (s_st ^ P.st_sxor[i])));
      }
    } else if (!(P.debug & 4) && io_on && !st_direct) {
      double2 *dst = psi + (base | g_st);
      if (io_iters == 16) {
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i)
          __stcs(dst + (uint64_t(P.st_goff[i]) << 3), lds128(tile_sa + (s_st ^ P.st_sxor[i])));
      } else {
#pragma unroll 4
        for (uint32_t i = 0; i < io_iters; ++i)
          __stcs(dst + (uint64_t(P.st_goff[i]) << 3), lds128(tile_sa + (s_st ^ P.st_sxor[i])));
      }
    }
    if (more) base = tile_base(tn);
    __syncthreads();  // every read of the tile is done before the next copy lands in it
    if (more) issue_load(base);
  }
}

size_t fused_smem_bytes(int K, int ntable) {
  return (size_t(1) << K) * sizeof(double2) + size_t(ntable) * sizeof(double2) + 4 * sizeof(uint32_t) +
         kPushMaxRanks * sizeof(double2 *);
}

constexpr size_t kSmemLimit = 227 * 1024;
int g_sms = 0;

template <bool FULL, bool FAST, bool PUSH>
cudaError_t configure_one() {
  cudaError_t err = cudaFuncSetAttribute(k_fused_pass<FULL, FAST, PUSH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(kSmemLimit));
  if (err != cudaSuccess) return err;
  return cudaFuncSetAttribute(k_fused_pass<FULL, FAST, PUSH>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

}  // namespace

cudaError_t fused_configure(int device) {
  static_assert(sizeof(QbOp) == 256, "QbOp layout");
  static_assert(sizeof(QbRound) % 4 == 0 && (QB_MAX_PASS_OPS * sizeof(QbOp)) % 16 == 0, "smem layout");
  static_assert(QB_MAX_PASS_OPS <= 128, "the active-op mask of the interpreter has 128 bits");
  static_assert(sizeof(FusedParams) <= 32764, "kernel parameter space");
  cudaError_t err = cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, device);
  if (err != cudaSuccess) return err;
  if ((err = configure_one<true, true, false>()) != cudaSuccess) return err;
  if ((err = configure_one<true, false, false>()) != cudaSuccess) return err;
  if ((err = configure_one<false, true, false>()) != cudaSuccess) return err;
  if ((err = configure_one<false, false, false>()) != cudaSuccess) return err;
  if ((err = configure_one<true, true, true>()) != cudaSuccess) return err;
  if ((err = configure_one<true, false, true>()) != cudaSuccess) return err;
  if ((err = configure_one<false, true, true>()) != cudaSuccess) return err;
  return configure_one<false, false, true>();
}

namespace {
cudaError_t fused_pass_impl(double2 *psi, int nbits, const DevicePass &p, cudaStream_t st, bool launch);
}

cudaError_t launch_fused_pass(double2 *psi, int nbits, const DevicePass &p, cudaStream_t st) {
  return fused_pass_impl(psi, nbits, p, st, true);
}

// Everything launch_fused_pass does on the host (capacity checks of the parameter block, address terms, round
// constants) without launching: the CPU tests run every plan through it (qb_plan_check), so a planner output
// the kernel's parameter block cannot hold is caught where no GPU exists.
cudaError_t check_fused_pass(int nbits, const DevicePass &p) { return fused_pass_impl(nullptr, nbits, p, nullptr, false); }

namespace {
cudaError_t fused_pass_impl(double2 *psi, int nbits, const DevicePass &p, cudaStream_t st, bool launch) {
  FusedParams P;
  P.psi = psi;
  P.nbits = nbits;
  P.desc = p.desc;
  if (p.desc.nops > QB_MAX_PASS_OPS || p.desc.nrounds > QB_MAX_PASS_ROUNDS) return cudaErrorInvalidValue;
  memcpy(P.ops, p.ops, size_t(p.desc.nops) * sizeof(QbOp));
  memcpy(P.rounds, p.rounds, size_t(p.desc.nrounds) * sizeof(QbRound));
  P.tables = p.tables;
  P.outph = p.outph;
  P.nlad = 0;
  for (int k = 0; k < p.desc.nops; ++k) {
    const int k8 = p.ops[k].kind & 0xff;
    if (k8 != QB_K_LADDER && k8 != QB_K_ULADDER) continue;
    if (P.nlad == QB_MAX_PASS_OPS) return cudaErrorInvalidValue;
    P.lad_tab[P.nlad] = uint32_t(p.ops[k].table_off);
    P.lad_ph[P.nlad] = uint32_t(p.ops[k].outph_off);
    ++P.nlad;
  }
  P.jbtab = p.jbtab;
  static const int dbg = getenv("QCC_B200_FUSED_DEBUG") ? atoi(getenv("QCC_B200_FUSED_DEBUG")) : 0;
  P.debug = dbg;
  static const int stagger = getenv("QCC_B200_STAGGER") ? atoi(getenv("QCC_B200_STAGGER")) : 0;
  P.stagger = stagger;
  P.nsm = g_sms;
  const int K = p.desc.K;
  if (K < 4 || K > QB_MAX_TILE_BITS || K > nbits) return cudaErrorInvalidValue;
  const unsigned ntiles = 1u << (nbits - K);
  const size_t smem = fused_smem_bytes(K, p.desc.ntable);
  if (smem > kSmemLimit) return cudaErrorInvalidValue;
  // copy-loop address terms of tile-local index 256 i (see the kernel)
  auto swz_h = [](uint32_t j) {
    uint32_t x = j >> 3;
    x ^= x >> 3;
    x ^= x >> 6;
    return j ^ (x & 7u);
  };
  P.ld_goff[0] = P.ld_sxor[0] = P.st_goff[0] = P.st_sxor[0] = 0;
  for (int k = 0; k < 8; ++k) {
    P.ld_gbit[k] = P.ld_sbit[k] = P.st_gbit[k] = P.st_sbit[k] = 0;
    if (k >= 3 && k < K) {
      P.ld_gbit[k] = uint32_t((uint64_t(1) << p.desc.tile_bits[p.desc.ld_map[k]]) >> 3);
      P.st_gbit[k] = uint32_t((uint64_t(1) << p.desc.tile_bits[p.desc.st_map[k]]) >> 3);
      P.ld_sbit[k] = swz_h(1u << p.desc.ld_map[k]) << 4;
      P.st_sbit[k] = swz_h(1u << p.desc.st_map[k]) << 4;
    }
  }
  for (uint32_t i = 0; i < (1u << K) / kFThreads; ++i) {
    // copy index 256 i: its bits 8.. drive tile-local positions ld_map[8..] / st_map[8..]
    uint32_t jl = 0, js = 0;
    uint64_t gl = 0, gs = 0;
    for (int k = 8; k < K; ++k) {
      const uint32_t bit = (i >> (k - 8)) & 1u;
      jl |= bit << p.desc.ld_map[k];
      js |= bit << p.desc.st_map[k];
      gl |= uint64_t(bit) << p.desc.tile_bits[p.desc.ld_map[k]];
      gs |= uint64_t(bit) << p.desc.tile_bits[p.desc.st_map[k]];
    }
    P.ld_goff[i] = uint32_t(gl >> 3);
    P.ld_sxor[i] = swz_h(jl) << 4;
    P.st_goff[i] = uint32_t(gs >> 3);
    P.st_sxor[i] = swz_h(js) << 4;
  }
  for (int r = 0; r < p.desc.nrounds; ++r) {
    const QbRound &R = p.rounds[r];
    RoundAux &X = P.aux[r];
    memset(&X, 0, sizeof X);
    if (R.prog == QB_PROG_GENERIC) continue;
    X.s = 1.0;
    const bool hl = R.prog == QB_PROG_HL3 || R.prog == QB_PROG_HL3U;
    for (int k = 0; k < 3; ++k) {
      X.b[k] = swz_h(1u << R.rbit[k]) << 4;
      if (hl) {
        X.ta[k] = uint32_t(p.ops[R.op_begin + k].table_off) << 4;
        X.s *= p.ops[R.op_begin + k].m[0];
      }
    }
    for (uint32_t git = 0; git < (1u << (K - 3)) / kFThreads; ++git) {
      uint32_t jb = 0;
      for (int k = 8; k < K - 3; ++k) jb |= ((git >> (k - 8)) & 1u) << R.qmap[k];
      X.pbi[git] = swz_h(jb) << 4;
      X.jbi[git] = jb;
      uint64_t g = 0;
      for (int k = 8; k < K - 3; ++k) g |= uint64_t((git >> (k - 8)) & 1u) << p.desc.tile_bits[R.qmap[k]];
      X.gji[git] = uint32_t(g >> 3);
    }
    for (int k = 0; k < 3; ++k) X.gb[k] = uint32_t((uint64_t(1) << p.desc.tile_bits[R.rbit[k]]) >> 3);
    for (int k = 0; k < 8 && k < K - 3; ++k) {
      X.pbt[k] = swz_h(1u << R.qmap[k]) << 4;
      X.jbk[k] = 1u << R.qmap[k];
      X.gbk[k] = uint32_t((uint64_t(1) << p.desc.tile_bits[R.qmap[k]]) >> 3);
    }
    if (R.prog == QB_PROG_UX) {
      int prev = -1;
      uint32_t nu = 0;
      for (int k = R.op_begin; k < R.op_end && nu < 3; ++k) {
        const QbOp &o = p.ops[k];
        const int opc = int(uint32_t(o.kind) >> 24);
        uint32_t cls = 0;
        if (opc >= QB_OPC_U_ALL && opc < QB_OPC_U_ALL + 3) cls = 1;
        else if (opc >= QB_OPC_U_ALL + 3 && opc < QB_OPC_U_ALL + 6) cls = 2;
        else if (opc >= QB_OPC_U_CI && opc < QB_OPC_U_CI + 3) cls = 3;
        if (cls == 0 || o.tpos <= prev || o.gmask != 0) break;  // the straight-line prefix has no predicates
        X.ux |= cls << (4 + 2 * o.tpos);
        prev = o.tpos;
        ++nu;
      }
      X.ux |= nu;
    }
  }
  auto is_hl = [&](int r) { return p.rounds[r].prog == QB_PROG_HL3 || p.rounds[r].prog == QB_PROG_HL3U; };
  // The Hadamard scales are plain scalars: collect those of all program rounds of the pass in the
  // first one (a round with s == 1 skips the multiplies).
  {
    int first = -1;
    double prod = 1.0;
    for (int r = 0; r < p.desc.nrounds; ++r)
      if (is_hl(r)) {
        if (first < 0) first = r;
        prod *= P.aux[r].s;
        P.aux[r].s = 1.0;
      }
    static const bool no_hoist = getenv("QCC_B200_NO_SCALE_HOIST") != nullptr;
    if (first >= 0 && !no_hoist) P.aux[first].s = prod;
    else if (no_hoist)
      for (int r = 0; r < p.desc.nrounds; ++r)
        if (is_hl(r)) {
          P.aux[r].s = 1.0;
          for (int k = 0; k < 3; ++k) P.aux[r].s *= p.ops[p.rounds[r].op_begin + k].m[0];
        }
  }
  if (dbg & (1 | 2 | 4 | 8 | 16)) P.desc.st_direct = 0;  // timing experiments use the copy path
  P.push_on = 0;
  if (p.push) {
    // exchange event fused into the store stage: every store address term goes through the event's bit
    // permutation (linear over OR on disjoint bit sets, so thread / iteration / tile terms stay separate)
    P.push_on = 1;
    P.push = *p.push;
    P.desc.st_direct = 0;
    auto sig = [&](uint64_t local) { return push_apply(*p.push, local) & ~p.push->rank_term; };
    for (int k = 3; k < 8 && k < K; ++k)
      P.st_gbit[k] = uint32_t(sig(uint64_t(1) << p.desc.tile_bits[p.desc.st_map[k]]) >> 3);
    for (uint32_t i = 0; i < (1u << K) / kFThreads; ++i) {
      uint64_t gs = 0;
      for (int k = 8; k < K; ++k) gs |= uint64_t((i >> (k - 8)) & 1u) << p.desc.tile_bits[p.desc.st_map[k]];
      P.st_goff[i] = uint32_t(sig(gs) >> 3);
    }
  }
  // FAST: no round needs the op interpreter (and the debug switches that fall back to it are off)
  bool fast = !(dbg & (1 | 16));
  for (int r = 0; r < p.desc.nrounds; ++r)
    if (p.rounds[r].prog == QB_PROG_GENERIC) fast = false;
  // One CTA per tile by default: up to 3 CTAs (FAST passes at K = 12: 64 KiB tile + 9 KiB of tables
  // each; else 2) are resident per SM, and because they start and finish at different times one
  // streams its tile while the others compute.  QCC_B200_FUSED_PERSIST=n instead launches n
  // persistent CTAs per SM that walk the tiles (measured 4 % slower on QFT-30: the per-CTA table
  // staging it saves is small, and the hardware scheduler balances the SMs better).
  static const int persist = getenv("QCC_B200_FUSED_PERSIST") ? atoi(getenv("QCC_B200_FUSED_PERSIST")) : 0;
  unsigned blocks = ntiles;
  if (persist > 0 && ntiles > unsigned(persist * g_sms)) blocks = unsigned(persist * g_sms);
  const bool full = ((1u << (K - 3)) % kFThreads) == 0;
  if (!launch) return cudaSuccess;
  if (P.push_on) {
    if (full && fast) k_fused_pass<true, true, true><<<blocks, kFThreads, smem, st>>>(P);
    else if (full) k_fused_pass<true, false, true><<<blocks, kFThreads, smem, st>>>(P);
    else if (fast) k_fused_pass<false, true, true><<<blocks, kFThreads, smem, st>>>(P);
    else k_fused_pass<false, false, true><<<blocks, kFThreads, smem, st>>>(P);
  } else if (full && fast) k_fused_pass<true, true, false><<<blocks, kFThreads, smem, st>>>(P);
  else if (full) k_fused_pass<true, false, false><<<blocks, kFThreads, smem, st>>>(P);
  else if (fast) k_fused_pass<false, true, false><<<blocks, kFThreads, smem, st>>>(P);
  else k_fused_pass<false, false, false><<<blocks, kFThreads, smem, st>>>(P);
  return cudaGetLastError();
}
}  // namespace

}  // namespace qb
