// Exercises every gate name of the libq face on 5 qubits and prints the final state with
// full precision (one "label re im" line per basis state above the print threshold).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "libq.h"

int main() {
  libq::qureg* q = libq::new_qureg(5, 5);
  libq::walsh(5, q);
  libq::x(0, q); libq::y(1, q); libq::z(2, q); libq::h(3, q); libq::t(4, q);
  libq::v(0, q); libq::yroot(1, q); libq::s(2, q);
  libq::cx(0, 1, q); libq::cz(1, 2, q); libq::ccx(0, 1, 3, q);
  libq::u1(2, 0.3, q); libq::cu1(3, 4, M_PI / 16, q);
  libq::cv(4, 0, q); libq::cv_adj(2, 3, q);
  libq::rx(0, 0.7, q); libq::ry(1, -0.4, q); libq::rz(2, 1.1, q);
  libq::crx(0, 4, 0.2, q); libq::cry(1, 3, 0.9, q); libq::crz(2, 0, -0.6, q);
  libq::sdag(1, q); libq::tdag(2, q); libq::vdag(3, q); libq::yrootdag(4, q);
  libq::ch(0, 2, q); libq::cs(1, 4, q); libq::ct(3, 0, q); libq::cy(4, 1, q); libq::cyroot(2, 3, q);
  libq::cmplx m[4] = {libq::cmplx(0.6f, 0.0f), libq::cmplx(0.0f, 0.8f), libq::cmplx(0.0f, 0.8f), libq::cmplx(0.6f, 0.0f)};
  libq::libq_gate1(4, m, q);
  libq::sync(q);
  printf("size %d width %d norm2 %.15f\n", q->size, q->width, libq::norm2(q));
  for (unsigned long long i = 0; i < 32; ++i) {
    libq::cmplxd a = libq::amplitude_of(i, q);
    printf("amp %llu %.17g %.17g\n", i, a.real(), a.imag());
  }
  libq::flush(q);
  libq::delete_qureg(q);
  return EXIT_SUCCESS;
}
