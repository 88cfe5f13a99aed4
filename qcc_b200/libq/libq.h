// libq.h -- the reference's libq C++ API (src/libq/libq.h:44-69), served by the B200 engine.
//
// A program written against the reference header -- hand-written like
// src/libq/libq_test.cc, or emitted by the transpiler (src/lib/dumpers.py:40-86) --
// compiles unchanged against this one:
//
//   g++ -O2 -Iqcc_b200/libq prog.cc -Lqcc_b200/lib -lqcc_libq -lqcc_b200 \
//       -Wl,-rpath,$PWD/qcc_b200/lib
//
// What is the same: namespace, type names (cmplx, state_t, qureg), every function name and
// argument order, LSB-first qubit numbering (qubit k = bit k of the basis label,
// reference libq.h:35-40), print formats (qureg.cc:64-86), `q->width` / `q->size`.
//
// What is different, on purpose:
//   * The register is a DENSE 2^width complex128 vector in B200 HBM, not a sparse hash of
//     complex<float>.  Width is limited by HBM (33 qubits on one B200), not by the
//     reference's int-overflowing hash (28, qureg.cc:18,25).
//   * Gate calls are asynchronous and fused (the queue/flush protocol of
//     gates_jit.cc:53-132).  `size`, `maxsize`, `state[]`, `amplitude[]` are host mirrors:
//     refreshed after every call when width <= 16, otherwise at flush / print_qureg /
//     print_qureg_stats / sync.  "Non-zero" means |amp|^2 >= 1e-6 / 2^width, the
//     reference's own pruning limit (apply.cc:107,150-171).
//   * v, yroot, cv, cv_adj implement the intended matrices (src/lib/ops.py:152-162 and
//     their controlled forms).  The reference's loops for these are wrong (gates.cc:9-15,
//     48-54, 96-118 apply the gate `size` times / uncontrolled).
//   * Angles are double (the reference takes float and evaluates cos/sin in float).
//   * Every gate name the transpiler can emit exists (appendix A of SURVEY.md): the
//     reference's 15 plus s, the *dag forms, rx/ry/rz, and all single-controlled forms.
#ifndef QCC_B200_LIBQ_H_
#define QCC_B200_LIBQ_H_

#include <complex>

namespace libq {

typedef std::complex<float> cmplx;
typedef std::complex<double> cmplxd;
typedef unsigned long long state_t;

struct qureg_impl;  // engine handle + bookkeeping, private to libq.cc

struct qureg_t {
  cmplx *amplitude;  // host mirror of the listed amplitudes (see header comment)
  state_t *state;    // their basis labels, ascending
  int width;         // number of qubits
  int size;          // number of basis states with non-negligible probability
  int maxsize;       // largest size observed
  int hash_computes; // number of general (non-diagonal, non-permutation) gate calls
  qureg_impl *impl;

  bool bit_is_set(int index, int target) const { return (state[index] >> target) & 1ULL; }
};
typedef struct qureg_t qureg;

// --- lifetime / output (reference qureg.cc) ------------------------------------
qureg *new_qureg(state_t initval, int width);
void delete_qureg(qureg *reg);
void print_qureg(qureg *reg);
void print_qureg_stats(qureg *reg);
void flush(qureg *reg);

// --- the reference's gate set (reference gates.cc) -----------------------------
void x(int target, qureg *reg);
void y(int target, qureg *reg);
void z(int target, qureg *reg);
void h(int target, qureg *reg);
void t(int target, qureg *reg);
void v(int target, qureg *reg);
void yroot(int target, qureg *reg);
void walsh(int width, qureg *reg);
void cx(int control, int target, qureg *reg);
void cz(int control, int target, qureg *reg);
void ccx(int control0, int control1, int target, qureg *reg);
void u1(int target, double gamma, qureg *reg);
void cu1(int control, int target, double gamma, qureg *reg);
void cv(int control, int target, qureg *reg);
void cv_adj(int control, int target, qureg *reg);

float probability(cmplx ampl);
void libq_gate1(int target, cmplx m[4], qureg *reg);

// --- names the transpiler emits that the reference header lacks ---------------
void s(int target, qureg *reg);
void sdag(int target, qureg *reg);
void tdag(int target, qureg *reg);
void vdag(int target, qureg *reg);
void hdag(int target, qureg *reg);
void xdag(int target, qureg *reg);
void ydag(int target, qureg *reg);
void zdag(int target, qureg *reg);
void yrootdag(int target, qureg *reg);
void rx(int target, double theta, qureg *reg);
void ry(int target, double theta, qureg *reg);
void rz(int target, double theta, qureg *reg);
void ch(int control, int target, qureg *reg);
void cs(int control, int target, qureg *reg);
void ct(int control, int target, qureg *reg);
void cy(int control, int target, qureg *reg);
void cyroot(int control, int target, qureg *reg);
void chdag(int control, int target, qureg *reg);
void csdag(int control, int target, qureg *reg);
void ctdag(int control, int target, qureg *reg);
void cvdag(int control, int target, qureg *reg);
void cxdag(int control, int target, qureg *reg);
void cydag(int control, int target, qureg *reg);
void czdag(int control, int target, qureg *reg);
void cyrootdag(int control, int target, qureg *reg);
void crx(int control, int target, double theta, qureg *reg);
void cry(int control, int target, double theta, qureg *reg);
void crz(int control, int target, double theta, qureg *reg);
// general (controlled) 2x2 in double precision
void gate1(int target, const cmplxd m[4], qureg *reg);
void cu(int control, int target, const cmplxd m[4], qureg *reg);

// --- engine extras ----------------------------------------------------------------
void sync(qureg *reg);                       // apply everything queued, refresh the mirrors
cmplxd amplitude_of(state_t label, qureg *reg);
double norm2(qureg *reg);

}  // namespace libq

#endif  // QCC_B200_LIBQ_H_
