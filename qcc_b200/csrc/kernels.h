// kernels.h -- host-callable launchers of the sm_100a kernels in kernels.cu / fused.cu.
#ifndef QCC_B200_CSRC_KERNELS_H_
#define QCC_B200_CSRC_KERNELS_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "qb_types.h"

namespace qb {

// ---- single-gate kernels (kernels.cu) ----------------------------------------
// psi: 2^nbits complex (double2 when is_double, else float2).  One launch, one sweep
// over exactly the amplitudes the gate touches.
cudaError_t launch_gate(void *psi, int nbits, const QbGate &g, bool is_double, cudaStream_t st);

// ---- init / readout kernels (kernels.cu) -------------------------------------
// value of element i is a hash of (seed, index_offset + i): shards of one state agree with the whole
cudaError_t launch_fill_random(double2 *psi, uint64_t n, uint64_t index_offset, uint64_t seed, cudaStream_t st);
cudaError_t launch_scale(double2 *psi, uint64_t n, double f, cudaStream_t st);
// out: 1 double (sum |psi|^2), zeroed by the launcher
cudaError_t launch_norm2(const double2 *psi, uint64_t n, double *out, cudaStream_t st);
// out: 1 double = sum of |psi_i|^2 over i with (i & mask) == want
cudaError_t launch_prob_mask(const double2 *psi, uint64_t n, uint64_t mask, uint64_t want, double *out,
                             cudaStream_t st);
// Per-block partial argmax: blk_prob[b], blk_idx[b] for b < argmax_blocks(); host finishes.  blk_idx holds
// LOGICAL indices (LogicalMap: where each local bit and the rank bits sit in the logical index), and ties go to
// the lowest logical index -- np.abs(psi).argmax() of state.py:75 whatever the exchange history of a shard.
struct LogicalMap {
  int identity = 1;
  int n = 0;          // local bits
  int lpos[40] = {};  // local physical bit -> logical bit
  uint64_t hi = 0;    // contribution of this rank's rank bits
};
int argmax_blocks();
cudaError_t launch_argmax(const double2 *psi, uint64_t n, double *blk_prob, uint64_t *blk_idx, const LogicalMap &lm,
                          cudaStream_t st);
// Compaction of indices with |psi|^2 >= thr: counter (1 x u64, zeroed by launcher), and up to
// cap (label, amp) pairs in arbitrary order.
cudaError_t launch_list_above(const double2 *psi, uint64_t n, double thr, uint64_t cap,
                              unsigned long long *counter, uint64_t *labels, double2 *amps,
                              cudaStream_t st);
// Pair exchange through peer memory, in place (engine.cu do_exchange, QCC_B200_EXCHANGE=swap): swaps this rank's outgoing
// half -- the amplitudes whose bit `victim` equals `sel_local` -- element by element with the partner's
// outgoing half (bit `victim` == 1 - sel_local) in the partner's shard, reached through `peer` (a CUDA IPC
// mapping of its state vector).  Each rank of the pair handles one half of the element range (`upper`),
// so every element pair is swapped exactly once and nothing is staged.
cudaError_t launch_pair_swap(double2 *local, double2 *peer, int nbits, int victim, int sel_local, int upper,
                             cudaStream_t st);
// Push exchange (engine.cu do_push_event): one exchange event is a permutation sigma of the bits of the
// DISTRIBUTED index (rank << nl | local index).  Every rank writes each of its amplitudes to where it belongs
// afterwards -- out[destination rank] + destination local index, the other ranks' buffers reached through
// CUDA IPC peer mappings over NVLink -- out of place, into the alternate buffer of the double-buffered
// state.  The fused pass does this in its store stage (fused.cu), k_push_remap as a plain copy.
constexpr int kPushMaxRanks = 16;
constexpr int kPushMaxMoved = 8;
struct PushMap {
  int nl = 0;                       // local index bits
  int nmoved = 0;                   // local source bits that move
  int src[kPushMaxMoved] = {};      // local source bit ...
  int dst[kPushMaxMoved] = {};      // ... -> bit of the distributed index it lands on (>= nl: a rank bit)
  uint64_t moved_mask = 0;          // OR of 1 << src[k]
  uint64_t rank_term = 0;           // sigma(rank << nl): where this rank's own rank bits land
  double2 *out[kPushMaxRanks] = {}; // per destination rank: base of the buffer the event fills
};
inline uint64_t push_apply(const PushMap &m, uint64_t local) {   // distributed destination index of a local index
  uint64_t a = (local & ~m.moved_mask) | m.rank_term;
  for (int k = 0; k < m.nmoved; ++k) a |= ((local >> m.src[k]) & 1) << m.dst[k];
  return a;
}
cudaError_t launch_push_remap(const double2 *psi, const PushMap &m, cudaStream_t st);
// Stream-ordered barrier over all ranks through counters in peer memory: rows[q] = rank q's row of kPushMaxRanks
// arrival counters (mapped), epoch = 1, 2, 3, ... the same on every rank.
cudaError_t launch_peer_barrier(unsigned long long *const *rows, int rank, int nranks, unsigned long long epoch,
                                cudaStream_t st);
cudaError_t launch_cvt_f2d(const float2 *in, double2 *out, uint64_t n, cudaStream_t st);

// ---- fused tile-resident pass (fused.cu) ---------------------------------------
struct DevicePass {
  QbPassDesc desc;
  const QbOp *ops;        // HOST: copied into the kernel parameters
  const QbRound *rounds;  // HOST
  const double2 *tables;  // device (ladder lookup tables, staged into smem) or nullptr
  const double2 *outph;   // device (per ladder: three tables of per-tile constants, planner.cc) or nullptr
  const uint32_t *jbtab;  // device (per round, per group: base index | swizzled slot << 16)
  const PushMap *push = nullptr;  // HOST: the pass stores through this exchange event's bit permutation
};
cudaError_t fused_configure(int device);  // opt in to large dynamic shared memory, query SM count
cudaError_t launch_fused_pass(double2 *psi, int nbits, const DevicePass &p, cudaStream_t st);
cudaError_t check_fused_pass(int nbits, const DevicePass &p);   // the host half of the launch, no GPU needed

}  // namespace qb

#endif  // QCC_B200_CSRC_KERNELS_H_
