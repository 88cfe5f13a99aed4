// libq_ref_driver.cc -- thin extern "C" handle over the REFERENCE libq, compiled
// together with the reference's own sources where they lie under
// /root/reference/src/libq (see oracle/Makefile).  TEST INFRASTRUCTURE ONLY:
// the product never links this.
//
// It lets the Python tests replay a flat op list through stock libq
// (libq.h:44-64) and read the sparse result straight from the struct
// (libq.h:14-41: state[], amplitude[], size) instead of parsing the 6-digit
// text print_qureg emits (qureg.cc:64-78).
//
// Built twice: against stock libq.h (cmplx = std::complex<float>) and against a
// scratch copy whose one typedef is switched to double (tier T2d).
// REFQ_REAL tells this file which one it is being compiled with.
#include <stdint.h>
#include <stdlib.h>

#include "libq.h"

#ifndef REFQ_REAL
#define REFQ_REAL float
#endif

extern "C" {

// op codes shared with oracle/oracle.py (LIBQ_OPS)
enum {
  RQ_X = 0, RQ_Y, RQ_Z, RQ_H, RQ_T, RQ_V, RQ_YROOT, RQ_WALSH,
  RQ_CX, RQ_CZ, RQ_CCX, RQ_U1, RQ_CU1, RQ_CV, RQ_CV_ADJ, RQ_GATE1
};

struct refq_op {
  int32_t code;
  int32_t a, b, c;     // qubit arguments in libq.h order
  double gamma;        // u1 / cu1 angle
  double m[8];         // libq_gate1 matrix (re, im) x 4
};

void *refq_new(unsigned long long initval, int width) {
  return libq::new_qureg(initval, width);
}

void refq_delete(void *q) { libq::delete_qureg(static_cast<libq::qureg *>(q)); }

int refq_apply(void *qv, const refq_op *ops, int64_t nops) {
  libq::qureg *q = static_cast<libq::qureg *>(qv);
  for (int64_t k = 0; k < nops; ++k) {
    const refq_op &o = ops[k];
    switch (o.code) {
      case RQ_X: libq::x(o.a, q); break;
      case RQ_Y: libq::y(o.a, q); break;
      case RQ_Z: libq::z(o.a, q); break;
      case RQ_H: libq::h(o.a, q); break;
      case RQ_T: libq::t(o.a, q); break;
      case RQ_V: libq::v(o.a, q); break;
      case RQ_YROOT: libq::yroot(o.a, q); break;
      case RQ_WALSH: libq::walsh(o.a, q); break;
      case RQ_CX: libq::cx(o.a, o.b, q); break;
      case RQ_CZ: libq::cz(o.a, o.b, q); break;
      case RQ_CCX: libq::ccx(o.a, o.b, o.c, q); break;
      case RQ_U1: libq::u1(o.a, o.gamma, q); break;
      case RQ_CU1: libq::cu1(o.a, o.b, o.gamma, q); break;
      case RQ_CV: libq::cv(o.a, o.b, q); break;
      case RQ_CV_ADJ: libq::cv_adj(o.a, o.b, q); break;
      case RQ_GATE1: {
        libq::cmplx m[4];
        for (int j = 0; j < 4; ++j)
          m[j] = libq::cmplx(static_cast<REFQ_REAL>(o.m[2 * j]),
                             static_cast<REFQ_REAL>(o.m[2 * j + 1]));
        libq::libq_gate1(o.a, m, q);
        break;
      }
      default: return -1;
    }
  }
  return 0;
}

int refq_size(void *q) { return static_cast<libq::qureg *>(q)->size; }
int refq_width(void *q) { return static_cast<libq::qureg *>(q)->width; }

// Copy the sparse register out: labels[i], amps[2i], amps[2i+1] as double.
int refq_read(void *qv, unsigned long long *labels, double *amps, int64_t cap) {
  libq::qureg *q = static_cast<libq::qureg *>(qv);
  int64_t n = q->size < cap ? q->size : cap;
  for (int64_t i = 0; i < n; ++i) {
    labels[i] = q->state[i];
    amps[2 * i] = q->amplitude[i].real();
    amps[2 * i + 1] = q->amplitude[i].imag();
  }
  return static_cast<int>(n);
}

void refq_print(void *q) { libq::print_qureg(static_cast<libq::qureg *>(q)); }

int refq_real_bytes(void) { return static_cast<int>(sizeof(REFQ_REAL)); }

}  // extern "C"
