// libq.cc -- the libq C++ face (libq.h) over the C ABI (include/qcc_b200.h).
//
// Reference counterparts: src/libq/qureg.cc (lifetime, printing), src/libq/gates.cc (named
// gates), src/libq/apply.cc (libq_gate1).  Nothing here computes amplitudes: every gate is
// a 2x2 handed to the engine, every readout is a device reduction / compaction.
#include "libq.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "qcc_b200.h"

namespace libq {

struct qureg_impl {
  qb_state *st = nullptr;
  std::vector<cmplx> amps;
  std::vector<state_t> labels;
  unsigned long long listed_total = 0;  // states above the limit on the device
  bool dirty = false;                   // gates queued since the mirrors were refreshed
};

namespace {

const int kEagerWidth = 16;                  // refresh `size` after every call up to here
const unsigned long long kListCap = 1 << 20; // most states the host mirrors will hold

void die(const char *what) {
  // The reference has no error channel (all functions are void, apply.cc:173 prints to
  // stderr); an engine failure here is unrecoverable for the caller, so say why and stop.
  fprintf(stderr, "libq (qcc_b200): %s: %s\n", what, qb_last_error());
  exit(EXIT_FAILURE);
}

double limit_of(const qureg *reg) {  // apply.cc:107
  return (1.0 / double(1ULL << reg->width)) * 1e-6;
}

void refresh(qureg *reg) {
  qureg_impl *im = reg->impl;
  if (!im->dirty) return;
  unsigned long long cap = kListCap;
  if (reg->width < 20) cap = 1ULL << reg->width;
  std::vector<uint64_t> lab(cap);
  std::vector<double> amp(2 * cap);
  uint64_t count = 0;
  if (qb_list_above(im->st, limit_of(reg), cap, lab.data(), amp.data(), &count) != QB_OK) die("qb_list_above");
  unsigned long long got = count < cap ? count : cap;
  im->labels.resize(got);
  im->amps.resize(got);
  for (unsigned long long i = 0; i < got; ++i) {
    im->labels[i] = lab[i];
    im->amps[i] = cmplx(float(amp[2 * i]), float(amp[2 * i + 1]));
  }
  im->listed_total = count;
  reg->state = im->labels.data();
  reg->amplitude = im->amps.data();
  reg->size = count > 0x7fffffffULL ? 0x7fffffff : int(count);
  if (reg->size > reg->maxsize) reg->maxsize = reg->size;
  im->dirty = false;
}

// Callers of the reference read the public fields between gates (libq_test.cc:12 reads q->size right after its
// gates), and `maxsize` -- "Maximum # of states" of print_qureg_stats, qureg.cc:80-86 -- is a running maximum over
// the program, part of the stdout the golden tests compare.  So registers of up to kEagerWidth qubits keep their
// mirrors exact after EVERY gate (a device listing per gate: microseconds at these sizes, and the listing
// buffers live in the engine state, not allocated per call); wider registers refresh lazily at flush / print /
// sync, where the queue can fuse.
void touched(qureg *reg) {
  reg->impl->dirty = true;
  if (reg->width <= kEagerWidth) refresh(reg);
}

void pack(const cmplxd m[4], double out[8]) {
  for (int k = 0; k < 4; ++k) {
    out[2 * k] = m[k].real();
    out[2 * k + 1] = m[k].imag();
  }
}

void one(int target, const cmplxd m[4], qureg *reg) {
  double mm[8];
  pack(m, mm);
  if (qb_apply1(reg->impl->st, target, mm) != QB_OK) die("qb_apply1");
  touched(reg);
}

void ctl(int control, int target, const cmplxd m[4], qureg *reg) {
  double mm[8];
  pack(m, mm);
  if (qb_applyc(reg->impl->st, control, target, mm) != QB_OK) die("qb_applyc");
  touched(reg);
}

void adj(const cmplxd m[4], cmplxd out[4]) {
  out[0] = std::conj(m[0]);
  out[1] = std::conj(m[2]);
  out[2] = std::conj(m[1]);
  out[3] = std::conj(m[3]);
}

// Gate matrices: gates.cc for x y z h t u1; src/lib/ops.py:136-207 for the rest.
const double kR = 0.70710678118654752440;  // sqrt(1/2), gates.cc:42
const cmplxd I(0.0, 1.0);
const cmplxd MX[4] = {0, 1, 1, 0};
const cmplxd MY[4] = {0, -I, I, 0};
const cmplxd MZ[4] = {1, 0, 0, -1};
const cmplxd MH[4] = {kR, kR, kR, -kR};
const cmplxd MS[4] = {1, 0, 0, I};
const cmplxd MV[4] = {cmplxd(0.5, 0.5), cmplxd(0.5, -0.5), cmplxd(0.5, -0.5), cmplxd(0.5, 0.5)};
const cmplxd MYROOT[4] = {cmplxd(0.5, 0.5), cmplxd(-0.5, -0.5), cmplxd(0.5, 0.5), cmplxd(0.5, 0.5)};

void phase_matrix(double gamma, cmplxd out[4]) {  // gates.cc:62-65 cexp, in double
  out[0] = 1;
  out[1] = 0;
  out[2] = 0;
  out[3] = cmplxd(cos(gamma), sin(gamma));
}

void rot_matrix(int axis, double theta, cmplxd out[4]) {  // ops.py:187-207
  double c = cos(theta / 2), s = sin(theta / 2);
  if (axis == 0) {
    out[0] = c; out[1] = cmplxd(0, -s); out[2] = cmplxd(0, -s); out[3] = c;
  } else if (axis == 1) {
    out[0] = c; out[1] = -s; out[2] = s; out[3] = c;
  } else {
    out[0] = cmplxd(c, -s); out[1] = 0; out[2] = 0; out[3] = cmplxd(c, s);
  }
}

}  // namespace

// ---- lifetime / output ---------------------------------------------------------------
float probability(cmplx ampl) { return ampl.real() * ampl.real() + ampl.imag() * ampl.imag(); }

qureg *new_qureg(state_t initval, int width) {
  qureg *reg = new qureg;
  reg->impl = new qureg_impl;
  reg->width = width;
  reg->maxsize = 0;
  reg->hash_computes = 0;
  const char *dev = getenv("QCC_B200_DEVICE");
  if (qb_state_create(width, initval, dev ? atoi(dev) : 0, &reg->impl->st) != QB_OK) die("qb_state_create");
  reg->impl->labels.assign(1, initval);
  reg->impl->amps.assign(1, cmplx(1.0f, 0.0f));
  reg->impl->listed_total = 1;
  reg->state = reg->impl->labels.data();
  reg->amplitude = reg->impl->amps.data();
  reg->size = 1;
  return reg;
}

void delete_qureg(qureg *reg) {
  if (!reg) return;
  qb_state_destroy(reg->impl->st);
  delete reg->impl;
  delete reg;
}

void sync(qureg *reg) {
  if (qb_sync(reg->impl->st) != QB_OK) die("qb_sync");
  refresh(reg);
}

void print_qureg(qureg *reg) {  // format of qureg.cc:64-78; states in ascending label order
  sync(reg);
  printf("States with non-zero probability:\n");
  for (size_t i = 0; i < reg->impl->labels.size(); ++i) {
    cmplx a = reg->amplitude[i];
    printf("  % f %+fi|%llu> (%e) (|", a.real(), a.imag(), reg->state[i], probability(a));
    for (int j = reg->width - 1; j >= 0; --j) {
      if (j % 4 == 3) printf(" ");
      printf("%i", int((reg->state[i] >> j) & 1ULL));
    }
    printf(">)\n");
  }
  if (reg->impl->listed_total > reg->impl->labels.size())
    printf("  ... %llu more states not listed\n",
           reg->impl->listed_total - (unsigned long long)reg->impl->labels.size());
}

void print_qureg_stats(qureg *reg) {  // format of qureg.cc:80-86
  sync(reg);
  long long theo = 2LL << reg->width;
  printf("# of qubits        : %d\n", reg->width);
  printf("# of hash computes : %d\n", reg->hash_computes);
  printf("Maximum # of states: %d, theoretical: %lld, %.3f%%\n", reg->maxsize, theo,
         100.0 * reg->maxsize / double(theo));
}

void flush(qureg *reg) { print_qureg_stats(reg); }  // gates.cc:148-150

cmplxd amplitude_of(state_t label, qureg *reg) {
  double a[2];
  if (qb_get_amplitude(reg->impl->st, label, a) != QB_OK) die("qb_get_amplitude");
  return cmplxd(a[0], a[1]);
}

double norm2(qureg *reg) {
  double v = 0;
  if (qb_norm2(reg->impl->st, &v) != QB_OK) die("qb_norm2");
  return v;
}

// ---- gates -----------------------------------------------------------------------------
void gate1(int target, const cmplxd m[4], qureg *reg) {
  reg->hash_computes += 1;
  one(target, m, reg);
}

void cu(int control, int target, const cmplxd m[4], qureg *reg) { ctl(control, target, m, reg); }

void libq_gate1(int target, cmplx m[4], qureg *reg) {  // apply.cc:78
  cmplxd md[4] = {m[0], m[1], m[2], m[3]};
  gate1(target, md, reg);
}

void x(int target, qureg *reg) { one(target, MX, reg); }
void y(int target, qureg *reg) { one(target, MY, reg); }
void z(int target, qureg *reg) { one(target, MZ, reg); }
void h(int target, qureg *reg) { gate1(target, MH, reg); }
void s(int target, qureg *reg) { one(target, MS, reg); }
void v(int target, qureg *reg) { gate1(target, MV, reg); }
void yroot(int target, qureg *reg) { gate1(target, MYROOT, reg); }

void t(int target, qureg *reg) {  // gates.cc:76-83
  cmplxd m[4];
  phase_matrix(M_PI / 4.0, m);
  one(target, m, reg);
}

void walsh(int width, qureg *reg) {  // gates.cc:56-60
  for (int i = 0; i < width; ++i) h(i, reg);
}

void u1(int target, double gamma, qureg *reg) {
  cmplxd m[4];
  phase_matrix(gamma, m);
  one(target, m, reg);
}

void cu1(int control, int target, double gamma, qureg *reg) {
  cmplxd m[4];
  phase_matrix(gamma, m);
  ctl(control, target, m, reg);
}

void cx(int control, int target, qureg *reg) { ctl(control, target, MX, reg); }
void cy(int control, int target, qureg *reg) { ctl(control, target, MY, reg); }
void cz(int control, int target, qureg *reg) { ctl(control, target, MZ, reg); }
void ch(int control, int target, qureg *reg) { ctl(control, target, MH, reg); }
void cs(int control, int target, qureg *reg) { ctl(control, target, MS, reg); }
void cv(int control, int target, qureg *reg) { ctl(control, target, MV, reg); }
void cyroot(int control, int target, qureg *reg) { ctl(control, target, MYROOT, reg); }

void ct(int control, int target, qureg *reg) {
  cmplxd m[4];
  phase_matrix(M_PI / 4.0, m);
  ctl(control, target, m, reg);
}

void cv_adj(int control, int target, qureg *reg) {
  cmplxd m[4];
  adj(MV, m);
  ctl(control, target, m, reg);
}

void ccx(int control0, int control1, int target, qureg *reg) {  // gates.cc:138-146
  double mm[8];
  pack(MX, mm);
  if (qb_applycc(reg->impl->st, control0, control1, target, mm) != QB_OK) die("qb_applycc");
  touched(reg);
}

#define QCC_DAG1(name, M)                          \
  void name(int target, qureg *reg) {              \
    cmplxd m[4];                                   \
    adj(M, m);                                     \
    one(target, m, reg);                           \
  }
#define QCC_DAGC(name, M)                                   \
  void name(int control, int target, qureg *reg) {         \
    cmplxd m[4];                                            \
    adj(M, m);                                              \
    ctl(control, target, m, reg);                           \
  }
QCC_DAG1(sdag, MS)
QCC_DAG1(vdag, MV)
QCC_DAG1(hdag, MH)
QCC_DAG1(xdag, MX)
QCC_DAG1(ydag, MY)
QCC_DAG1(zdag, MZ)
QCC_DAG1(yrootdag, MYROOT)
QCC_DAGC(chdag, MH)
QCC_DAGC(csdag, MS)
QCC_DAGC(cvdag, MV)
QCC_DAGC(cxdag, MX)
QCC_DAGC(cydag, MY)
QCC_DAGC(czdag, MZ)
QCC_DAGC(cyrootdag, MYROOT)

void tdag(int target, qureg *reg) {
  cmplxd m[4];
  phase_matrix(-M_PI / 4.0, m);
  one(target, m, reg);
}

void ctdag(int control, int target, qureg *reg) {
  cmplxd m[4];
  phase_matrix(-M_PI / 4.0, m);
  ctl(control, target, m, reg);
}

void rx(int target, double theta, qureg *reg) {
  cmplxd m[4];
  rot_matrix(0, theta, m);
  gate1(target, m, reg);
}
void ry(int target, double theta, qureg *reg) {
  cmplxd m[4];
  rot_matrix(1, theta, m);
  gate1(target, m, reg);
}
void rz(int target, double theta, qureg *reg) {
  cmplxd m[4];
  rot_matrix(2, theta, m);
  one(target, m, reg);
}
void crx(int control, int target, double theta, qureg *reg) {
  cmplxd m[4];
  rot_matrix(0, theta, m);
  ctl(control, target, m, reg);
}
void cry(int control, int target, double theta, qureg *reg) {
  cmplxd m[4];
  rot_matrix(1, theta, m);
  ctl(control, target, m, reg);
}
void crz(int control, int target, double theta, qureg *reg) {
  cmplxd m[4];
  rot_matrix(2, theta, m);
  ctl(control, target, m, reg);
}

}  // namespace libq
