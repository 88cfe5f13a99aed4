// qb_types.h -- internal types shared by the host engine, the fusion planner and
// the CUDA kernels.  Nothing here is part of the C ABI (include/qcc_b200.h).
#ifndef QCC_B200_CSRC_QB_TYPES_H_
#define QCC_B200_CSRC_QB_TYPES_H_

#include <stdint.h>

// How a 2x2 acts, decided once on the host from its entries (engine.cu classify):
//   U      general 2x2 on (psi[i0], psi[i1])                          h v yroot rx ry ...
//   PHASE  diag(1, p): psi[i] *= p where target and control bits set   z s t u1 cz cu1
//   DIAG   diag(d0, d1), d0 != 1                                       rz
//   PERM   antidiagonal | 0 b ; c 0 |: swap with weights               x y cx ccx
// PHASE only touches half (or a quarter, an eighth) of the vector; PERM with b = c = 1
// is the pure label permutation of gates.cc:17-21,120-146.
enum QbKind : int32_t {
  QB_K_U = 0,
  QB_K_PHASE = 1,
  QB_K_DIAG = 2,
  QB_K_PERM = 3,
  QB_K_LADDER = 4,  // fused run of PHASE gates sharing one pivot bit (QFT ladder)
  QB_K_NOP = 5,
  QB_K_SWAP = 6,    // PERM with b = c = 1: moves amplitudes, no arithmetic
  QB_K_ULADDER = 7, // uncontrolled U on the pivot immediately followed by that pivot's LADDER
  QB_K_PARSWAP = 8, // product of x / cx gates sharing one target: swap the pair where the PARITY of the
                    // control bits (xor a constant) is odd -- cx(a,t) cx(b,t) ... = X_t^(a xor b xor ...)
};

// A gate in physical index-bit terms, as queued by the engine.
struct QbGate {
  uint64_t ctl_mask;  // bits that must be 1 (never contains `target`)
  int32_t target;
  int32_t kind;       // QbKind
  double m[8];        // a b c d as (re, im)
};

// ---------------------------------------------------------------------------
// Fused-pass format (produced by planner.cc, consumed by fused.cu).
//
// A PASS is one HBM sweep.  It fixes a set of K "tile bits" (index-bit positions,
// ascending, always containing bits 0..QB_TILE_LOW-1 so that every tile is made of
// >= 128-byte contiguous runs).  A TILE is the 2^K amplitudes obtained by fixing all
// non-tile bits; a CTA loads one tile into shared memory, applies every op of the
// pass, and stores it back.  Inside a tile an amplitude is addressed by its K-bit
// tile-local index j (bit k of j = index bit tile_bits[k]).
//
// Ops are grouped into ROUNDS.  A round names QB_ROUND_BITS tile-local positions
// rbit[]; each thread pulls the 2^3 amplitudes that differ only in those bits into
// registers, applies all ops of the round there, and writes them back -- one
// shared-memory round trip per round, not per gate.  The remaining K-3 local bits
// enumerate the groups; qmap[] says which local bit each group-index bit drives and
// is chosen by the planner so that 8 consecutive lanes always fall into 8 distinct
// 16-byte bank groups of the swizzled tile (see fused.cu).
//
// Every op carries a three-level predicate, all of the form (x & mask) == want:
//   g*: on the index bits outside the tile  -> uniform per CTA
//   l*: on tile-local bits outside the round -> one test per thread-group
//   r*: on round positions (3 bits)          -> resolved per register at unroll time
// ---------------------------------------------------------------------------
#define QB_TILE_LOW 3        // bits 0..2 are always tile bits (8 amps = 128 B runs)
#define QB_MAX_TILE_BITS 13  // 2^13 * 16 B = 128 KiB
#define QB_ROUND_BITS 3      // 8 amplitudes = 16 doubles in registers per thread
#define QB_LADDER_LANE_BITS 5 // ladder lookup tables: T_a over group-index bits 0..4 (the lane), T_b over the rest
#define QB_MAX_PASS_OPS 96   // ops of one pass travel as kernel parameters (96 * 256 B of the 32 KiB parameter space)
#define QB_MAX_PASS_ROUNDS 16
#define QB_MAX_SEGS 6
#define QB_MAX_PASS_LADDERS 12  // their lookup tables (48 x 16 B each at K = 12) are staged in shared memory too
#define QB_MF_REAL 1         // all four entries of m are real: 8 instead of 20 flops per pair
#define QB_MF_HADAMARD 2     // m = r * [[1, 1], [1, -1]] with real r
#define QB_MF_COLIMAG 4      // first column real, second column imaginary (R * diag(1, +-i), e.g. h then v, h then s): 8 flops per pair
// Dense opcodes (bits 24..31 of QbOp::kind): one jump-table switch in the kernel.
#define QB_OPC_ULADDER 0     // +tpos complex, +3+tpos real, +6+tpos Hadamard
#define QB_OPC_U_ALL 9       // uncontrolled U: +tpos complex, +3+tpos real
#define QB_OPC_U_MASKED 15   // U with controls inside the tile: +tpos
#define QB_OPC_PERM 18       // +tpos
#define QB_OPC_SWAP 21       // +tpos
#define QB_OPC_PHASE 24
#define QB_OPC_LADDER 25
#define QB_OPC_PARSWAP 26    // +tpos
#define QB_OPC_U_CI 29       // uncontrolled U, QB_MF_COLIMAG matrix: +tpos

// Round programs (QbRound::prog).  The op list stays the definition of what a round computes;
// prog only names a fully unrolled code path in fused.cu for op lists of a known shape.
#define QB_PROG_GENERIC 0    // interpret the op list
#define QB_PROG_HL3 1        // exactly three Hadamard+ladder ops on round positions 0, 1, 2 in that order
#define QB_PROG_HL3U 2       // HL3 whose in-round ladder partners are all above their pivot (the QFT shape)
#define QB_PROG_UX 3         // every op is an uncontrolled U (any round position; complex, real or column-imaginary)
                             // or a PARSWAP: leading U's on ascending positions run as straight-line code, the rest
                             // through a lean interpreter without predicates (12 opcodes)

struct alignas(16) QbOp {   // 256 bytes; lives in the kernel parameters (constant bank)
  int32_t kind;      // QbKind (never DIAG/NOP: the planner lowers those) | tpos << 8 | mflags << 16 | opcode << 24
  int32_t tpos;      // U/PERM/SWAP: target position inside rbit[]
  uint32_t lmask, lwant;   // PARSWAP: lmask = tile-local control bits (parity), lwant = 8-bit table, bit e = parity(e & rmask)
  uint32_t rmask, rwant;   // PARSWAP: rmask = round-position control bits (parity), rwant = constant flip (0 / 1)
  int32_t table_off; // LADDER: first double2 of this op's tables in the pass table buffer
  int32_t flags;     // LADDER: slot (0..63) of its per-tile constant in shared memory
  uint64_t gmask, gwant;   // PARSWAP: gmask = control bits outside the tile (parity), gwant unused
  double m[8];       // U/PERM: a b c d; PHASE: p in m[0..1]
  int32_t nout;      // LADDER: number of partner bits outside the tile (plan dump only)
  int32_t out_off;   // LADDER: first entry in the pass's outside-bit list (plan dump only)
  int32_t mflags;    // QB_MF_* properties of m
  int32_t outph_off; // LADDER: first entry of its three per-tile-constant tables in the pass's outside-phase array
  double F[16];      // LADDER: phase of the round's own bits for each of the 8 registers (re, im)
};

struct QbRound {
  int32_t nbits;                     // == min(K, QB_ROUND_BITS)
  int32_t rbit[QB_ROUND_BITS];       // tile-local positions, ascending
  int32_t qmap[QB_MAX_TILE_BITS];    // local position driven by group-index bit k
  int32_t op_begin, op_end;          // range in the pass's op array
  int32_t nobar;                     // 1: the next round stays inside each warp's sub-cube (warp sync, no CTA barrier)
  int32_t prog;                      // QB_PROG_*: the kernel's specialised code path for this round's op list
};

struct QbPassDesc {
  int32_t K;                           // tile bits
  int32_t nrounds;
  int32_t nops;
  int32_t ntable;                      // double2 entries in the (staged) ladder table buffer
  int32_t ngroups_log2;                // K - 3
  int32_t nseg;                        // runs of consecutive non-tile index bits (tile number -> base)
  int32_t seg_pos[QB_MAX_SEGS];        // first index bit of run r
  int32_t seg_len[QB_MAX_SEGS];        // its length; nseg > QB_MAX_SEGS is flagged as nseg = -1 (generic loop)
  int32_t nout_total;                  // entries in the pass's outside-phase array
  int32_t pad_;
  int32_t lad_w[3];                    // field widths of the tile number for the per-tile ladder constants (planner.cc)
  int32_t pad3_;
  int32_t tile_bits[QB_MAX_TILE_BITS + 3];  // index-bit positions, ascending
  uint64_t tile_mask;                  // OR of 1 << tile_bits[k]
  // Tile <-> HBM copy maps: bit k of the copy index c = tid + 256 * iteration drives tile-local position
  // ld_map[k] (load) / st_map[k] (store); positions 0..2 always stay with bits 0..2 (8 lanes = one 128-byte run).
  // warp_io = 1: bits 5..7 of c (the warp number) drive the three positions outside the per-warp sub-cube of the
  // first (load) / last (store) run of rounds, so a warp copies exactly the amplitudes it computes on and waits
  // only for its own copies -- no CTA barrier between load and first round or between last round and store.
  // warp_io = 0: identity maps, CTA barriers.
  int32_t warp_io;
  int32_t ld_map[QB_MAX_TILE_BITS];
  int32_t st_map[QB_MAX_TILE_BITS];
  // st_direct = 1: the last round writes its groups straight back to HBM.  Needs warp_io, a round
  // program, round bits outside tile positions 0..2 and lanes 0..7 of a warp on positions 0..2 (so 8 lanes
  // still move one 128-byte run).  Saves one shared-memory write and one read of the whole tile.
  int32_t st_direct, pad2_;
};

#endif  // QCC_B200_CSRC_QB_TYPES_H_
