// Same gate sequence as the reference's hand-written smoke test (src/libq/libq_test.cc:6-17):
// Bell pair, then a phase on qubit 1; reads q->size like the reference does.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "libq.h"

int main() {
  libq::qureg* q = libq::new_qureg(0, 2);
  libq::h(0, q);
  libq::cx(0, 1, q);
  libq::u1(1, M_PI / 8.0, q);
  printf(" # States: %d\n", q->size);
  libq::print_qureg(q);
  libq::delete_qureg(q);
  return EXIT_SUCCESS;
}
