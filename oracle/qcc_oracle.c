/*
 * qcc_oracle.c -- CPU restatement of the reference's dense gate-application
 * path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 * The product (qcc_b200/) never links, imports or calls anything in oracle/.
 *
 * What it restates (all citations relative to /root/reference):
 *   - orc_apply1_{d,f}: src/lib/xgates.cc:23-41 (apply1<T>) which is itself
 *     the C++ form of the Python definition src/lib/state.py:80-100.
 *   - orc_applyc_{d,f}: src/lib/xgates.cc:45-67 (applyc<T>) / state.py:102-125,
 *     including the "negative control index" behaviour that
 *     src/lib/circuit_test.py:94-104 exercises: the predicate is evaluated on
 *     idx = g * 2^nbits + i, so a control position >= nbits tests a bit of the
 *     pair-group base g.  state.py does this in arbitrary precision; here it
 *     is done in 128-bit-safe form (bit c of (g << nbits) + i).
 *   - orc_run_d: replays a flat gate list through the two functions above
 *     (the shape of circuit.py:180-215 in eager mode).
 *
 * Pinning: tests/test_oracle.py checks these functions against (i) golden
 * vectors generated from the reference's own Python spec and xgates
 * (tests/golden/make_golden.py), and (ii) the reference xgates build
 * oracle/_ref/libxgates.so when present.
 *
 * Arithmetic order matches xgates.cc exactly (t1 = g0*a + g1*b; t2 = g2*a +
 * g3*b with complex multiply expanded as (ar*br - ai*bi, ar*bi + ai*br)), no
 * -ffast-math, so double results are reproducible bit-for-bit on any host.
 */
#include <stdint.h>
#include <stddef.h>

typedef struct { double re, im; } cd;
typedef struct { float re, im; } cf;

#define CMUL_RE(a, b) ((a).re * (b).re - (a).im * (b).im)
#define CMUL_IM(a, b) ((a).re * (b).im + (a).im * (b).re)

/* Predicate of xgates.cc:57-58 / state.py:118-120 without overflow:
 * bit `ctl` of (g << nbits) + i, where i < 2^nbits so no carries cross. */
static inline int ctl_bit(uint64_t g, uint64_t i, int nbits, int ctl) {
  if (ctl < nbits) return (int)((i >> ctl) & 1u);
  int c = ctl - nbits;
  if (c >= 64) return 0;
  return (int)((g >> c) & 1u);
}

#define DEFINE_APPLY(SUF, T)                                                   \
  /* xgates.cc:23-41.  tgt is the reference's MSB-first qubit index. */        \
  int orc_apply1_##SUF(T *psi, const T *gate, int nbits, int tgt) {            \
    int t = nbits - tgt - 1;                                                   \
    if (t < 0 || t >= nbits) return -1; /* xgates.cc:28-32 exits the process */\
    uint64_t q2 = (uint64_t)1 << t, n = (uint64_t)1 << nbits;                  \
    for (uint64_t g = 0; g < n; g += q2 << 1) {                                \
      for (uint64_t i = g; i < g + q2; ++i) {                                  \
        T a = psi[i], b = psi[i + q2], t1, t2;                                 \
        t1.re = CMUL_RE(gate[0], a) + CMUL_RE(gate[1], b);                     \
        t1.im = CMUL_IM(gate[0], a) + CMUL_IM(gate[1], b);                     \
        t2.re = CMUL_RE(gate[2], a) + CMUL_RE(gate[3], b);                     \
        t2.im = CMUL_IM(gate[2], a) + CMUL_IM(gate[3], b);                     \
        psi[i] = t1;                                                           \
        psi[i + q2] = t2;                                                      \
      }                                                                        \
    }                                                                          \
    return 0;                                                                  \
  }                                                                            \
  /* xgates.cc:45-67.  ctl/tgt are MSB-first; ctl may be negative. */          \
  int orc_applyc_##SUF(T *psi, const T *gate, int nbits, int ctl, int tgt) {   \
    int t = nbits - tgt - 1;                                                   \
    int c = nbits - ctl - 1;                                                   \
    if (t < 0 || t >= nbits) return -1;                                        \
    if (c < 0) return -2; /* 1 << negative: UB in C, ValueError in Python */   \
    uint64_t q2 = (uint64_t)1 << t, n = (uint64_t)1 << nbits;                  \
    for (uint64_t g = 0; g < n; g += q2 << 1) {                                \
      for (uint64_t i = g; i < g + q2; ++i) {                                  \
        if (!ctl_bit(g, i, nbits, c)) continue;                                \
        T a = psi[i], b = psi[i + q2], t1, t2;                                 \
        t1.re = CMUL_RE(gate[0], a) + CMUL_RE(gate[1], b);                     \
        t1.im = CMUL_IM(gate[0], a) + CMUL_IM(gate[1], b);                     \
        t2.re = CMUL_RE(gate[2], a) + CMUL_RE(gate[3], b);                     \
        t2.im = CMUL_IM(gate[2], a) + CMUL_IM(gate[3], b);                     \
        psi[i] = t1;                                                           \
        psi[i + q2] = t2;                                                      \
      }                                                                        \
    }                                                                          \
    return 0;                                                                  \
  }

DEFINE_APPLY(d, cd)
DEFINE_APPLY(f, cf)

/* Flat gate record used by the test harness: kind 1 = single (circuit.py:180),
 * kind 2 = controlled (circuit.py:199).  Indices are the reference's
 * MSB-first qubit numbers; m is the 2x2 row-major (a b c d), re/im pairs. */
typedef struct {
  int32_t kind;
  int32_t ctl;
  int32_t tgt;
  int32_t pad;
  double m[8];
} orc_gate;

int orc_run_d(cd *psi, int nbits, const orc_gate *gates, int64_t ngates) {
  for (int64_t k = 0; k < ngates; ++k) {
    const orc_gate *g = &gates[k];
    cd m[4];
    for (int j = 0; j < 4; ++j) { m[j].re = g->m[2 * j]; m[j].im = g->m[2 * j + 1]; }
    int rc;
    if (g->kind == 1) rc = orc_apply1_d(psi, m, nbits, g->tgt);
    else if (g->kind == 2) rc = orc_applyc_d(psi, m, nbits, g->ctl, g->tgt);
    else rc = -3;
    if (rc) return rc;
  }
  return 0;
}

int orc_run_f(cf *psi, int nbits, const orc_gate *gates, int64_t ngates) {
  for (int64_t k = 0; k < ngates; ++k) {
    const orc_gate *g = &gates[k];
    cf m[4];
    for (int j = 0; j < 4; ++j) { m[j].re = (float)g->m[2 * j]; m[j].im = (float)g->m[2 * j + 1]; }
    int rc;
    if (g->kind == 1) rc = orc_apply1_f(psi, m, nbits, g->tgt);
    else if (g->kind == 2) rc = orc_applyc_f(psi, m, nbits, g->ctl, g->tgt);
    else rc = -3;
    if (rc) return rc;
  }
  return 0;
}
