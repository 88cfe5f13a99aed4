// shard.h -- host-side lowering of a logical gate stream onto ONE rank's shard of a state
// that is split across 2^p GPUs by its top p physical index bits.  Pure host code.
//
// The reference has no counterpart (its only "scaling" is the arithmetic extrapolation of
// src/supremacy.py:255-297, which assumes zero communication).  Rules:
//   * logical index bit b lives at physical bit perm[b]; physical bits [0, nl) are local,
//     physical bit nl + k is bit k of the rank;
//   * a control on a global bit is a per-rank predicate: ranks whose bit is 0 skip the gate;
//   * a diagonal gate (PHASE / DIAG) whose target is global needs no communication: on a
//     given rank it is a phase on the remaining local bits (or a scalar);
//   * a non-diagonal gate whose target is global first swaps that global bit with a local
//     "victim" bit -- one pairwise half-shard exchange with rank ^ (1 << k) -- and updates
//     perm, so every later gate on that qubit is local.  The victim is the high local bit
//     whose logical qubit is needed as a non-diagonal target farthest in the future.
//   * an uncontrolled x on a global bit moves no data at all: the rank bit is RELABELLED (`flip`), i.e.
//     from then on it carries the negation of the logical qubit; controls and diagonal gates read it
//     through the flip, and the next exchange of that bit undoes it with one local x on the victim
//     (grover.py's x layers around its multi-controlled gates hit sharded control qubits every
//     iteration: 10.75 -> 6.75 exchanges per iteration on 4 ranks);
// Every rank runs the same lowering on the same stream, so all ranks agree on the exchange
// sequence without talking to each other.
#ifndef QCC_B200_CSRC_SHARD_H_
#define QCC_B200_CSRC_SHARD_H_

#include <stdint.h>

#include <string>
#include <vector>

#include "qb_types.h"

namespace qb {

struct ShardStep {
  // kind 0: run `gates` (physical LOCAL bits, ready for plan_gates / launch_gate) on the shard
  // kind 1: exchange EVENT -- for every k, swap physical global bit nl + rank_bits[k] with local bit
  //         victims[k].  The pairs of one event are disjoint, so the event is one bit permutation of the
  //         distributed index: executed as ONE all-to-all (each rank keeps 2^-k of its shard and sends
  //         2^-k to each of the 2^k - 1 ranks that differ from it in those rank bits) by the push
  //         exchange of engine.cu, or pair after pair by the in-place / NCCL exchanges.
  int kind = 0;
  std::vector<QbGate> gates;
  int64_t retired = 0;   // logical gate records this step accounts for (incl. skipped ones)
  std::vector<int> rank_bits;
  std::vector<int> victims;
  // lands[k]: the local bit where the qubit that comes in from rank bit rank_bits[k] is put.  == victims[k]: a
  // plain swap.  Otherwise a 3-cycle (push exchange only, which is an out-of-place remap anyway): victim bit ->
  // rank bit, rank bit -> lands[k], lands[k] -> victim bit.  Landing on the HIGHEST local bits keeps the low
  // address bits of every (source, destination) flow free.  Opt-in (QCC_B200_LAND=1): measured on 8 GPUs at 34
  // qubits it did not make the slow push passes (low victim bits) any faster.
  std::vector<int> lands;
};

struct ShardLayout {
  int n = 0;       // logical qubits
  int nl = 0;      // local physical bits
  int p = 0;       // global physical bits (nranks = 2^p)
  int rank = 0;
  std::vector<int> perm;  // logical bit -> physical bit
  uint32_t flip = 0;      // bit k set: rank bit k carries the NEGATION of the logical qubit mapped to it
  int window = 6;         // victims come from the top `window` local bits (kVictimWindow unless the exchange
                          // does not care about contiguity: the peer-swap kernel, engine.cu)
  // hoist = 1: an exchange is moved back from the gate that needs it to the latest point where the fusion
  // planner would start a new pass anyway (pass_targets new target bits per pass), as far as the victim
  // allows -- the stream is then planned in whole passes on both sides of the exchange instead of ending
  // one segment with a fragment.  Pays off with the wide victim window (the victim can be a qubit that
  // is long done); off with NCCL's narrow one.
  int hoist = 0;
  int pass_targets = 9;
  // prefetch = 1: an event also brings in every other sharded qubit that is needed (as a mixing target)
  // before the local qubit it would evict -- the all-to-all moves 1 - 2^-k of a shard for k bits, so the
  // extra bits are nearly free, whereas a separate event later costs another half shard.  Only with the
  // push exchange; the pairwise exchanges pay half a shard per bit either way.
  int prefetch = 0;
  int land = 0;    // 1: arriving qubits land on the highest local bits (see ShardStep::lands); push exchange only, opt-in
};

// Victims are taken from the top `kVictimWindow` local bits so that the exchanged half
// shard is made of at most 2^(kVictimWindow-1) contiguous runs (one NCCL send / recv each).  The
// peer-swap kernel only needs runs of a few hundred bytes: its window is everything above bit
// kPeerSwapMinVictim.
constexpr int kVictimWindow = 6;
constexpr int kPeerSwapMinVictim = 5;
constexpr int kPeerSwapMinVictimLarge = 9;   // shards with >= 29 local bits (engine.cu create_impl)

// Lowers `gates` (LOGICAL index bits, kinds already classified) for layout->rank, updating
// layout->perm as exchanges are scheduled.
void lower_for_rank(ShardLayout *layout, const QbGate *gates, int64_t ngates, std::vector<ShardStep> *steps);

// Exchange sequence (as steps of kind 1) + local bit transpositions (as cx triples inside
// kind-0 steps) that bring perm back to the identity.
void canonicalize_steps(ShardLayout *layout, std::vector<ShardStep> *steps);

std::string steps_to_json(const ShardLayout &layout, const std::vector<ShardStep> &steps);

}  // namespace qb

#endif  // QCC_B200_CSRC_SHARD_H_
