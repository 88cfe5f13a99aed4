// planner.h -- gate-stream -> fused-pass planner (pure host code, no CUDA calls).
//
// The reference's only fusion attempt is src/libq/gates_jit.cc:53-132: a queue of
// diagonal / permutation ops replayed per basis state in one sweep, flushed by any
// general gate.  On a GPU the same idea pays off because a sweep costs an HBM round
// trip: this planner cuts the queued gate stream into PASSES (one HBM sweep each) and
// each pass into ROUNDS (one shared-memory round trip each) -- see qb_types.h.
#ifndef QCC_B200_CSRC_PLANNER_H_
#define QCC_B200_CSRC_PLANNER_H_

#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "qb_types.h"

namespace qb {

struct Cplx {
  double x, y;
};

struct PlannedPass {
  QbPassDesc desc{};
  std::vector<QbOp> ops;
  std::vector<QbRound> rounds;
  std::vector<Cplx> tables;      // per ladder: T_a[32], T_b[2^(K-8)], indexed by group number (staged in smem)
  std::vector<Cplx> outph;       // per ladder: three tables of per-tile constants over the fields of the tile number
  std::vector<int32_t> outbits;  // per ladder: the outside partner bits (plan dump only)
  std::vector<uint32_t> jbtab;   // per round, per group q: jb | swizzled slot(jb) << 16
  // filled by Plan::layout(): byte offsets inside the serialized blob
  size_t ops_off = 0, rounds_off = 0, tables_off = 0, outbits_off = 0, outph_off = 0, jbtab_off = 0;
  int noutbits = 0;
  int64_t single_gate = -1;  // >= 0: run `single` with the plain sweep kernel instead
  QbGate single{};           // the (possibly pre-multiplied) gate of a single-gate pass
  int64_t ngates = 0;        // gate records this pass retires (including no-ops)
  double bytes_algorithmic_per_amp = 0.0;
};

struct Plan {
  std::vector<PlannedPass> passes;
  size_t blob_bytes();               // lays the passes out and returns the blob size
  void serialize(char *dst) const;   // after blob_bytes()
  std::string to_json() const;       // for the CPU tests (tests/ interprets it with numpy)
};

// How a 2x2 acts (QbKind): exact comparisons on purpose -- a matrix that is merely close to
// diagonal still goes through the general butterfly so results track the reference.
int classify_matrix(const double m[8]);

// SURVEY.md 8(d) algorithmic bytes per amplitude of the full vector for one gate.
double gate_bytes_per_amp(const QbGate &g);

// Peephole over a queued stream, in place: every five-gate Sleator-Weinfurter run (circuit.py:227-246) becomes
// one doubly-controlled gate + four no-ops.  Returns the number of runs replaced.  See planner.cc.
int64_t fuse_ccu_runs(QbGate *gates, int64_t ngates);

// fuse_last: the last pass is made a fused pass even where a plain single-gate sweep would be cheaper (a sharded
// state's exchange event is about to ride on its store stage).
void plan_gates(int nbits, const QbGate *gates, int64_t ngates, int tile_bits, Plan *out, bool fuse_last = false);

}  // namespace qb

#endif  // QCC_B200_CSRC_PLANNER_H_
