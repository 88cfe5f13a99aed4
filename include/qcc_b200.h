/*
 * qcc_b200.h -- C ABI of the B200-native state-vector gate-application engine.
 *
 * This is the drop-in boundary for the one hot path of qcc4cp/qcc: applying
 * 1-qubit and controlled gates to a dense 2^n complex amplitude vector.
 * Plain C, plain pointers and sizes; no torch / numpy / C++ types cross it.
 * The library behind it (qcc_b200/lib/libqcc_b200.so) is hand-written CUDA for
 * sm_100a; there is NO CPU fallback -- every entry point that touches a state
 * returns QB_ERR_CUDA (and qb_last_error() says why) when no B200 is usable.
 *
 * Reference interfaces each group replaces (paths relative to the reference
 * tree, qcc4cp/qcc @ 4605fd8):
 *
 *   group                        replaces
 *   ---------------------------  ------------------------------------------------
 *   qb_host_apply1/applyc        src/lib/xgates.cc:89-107 (apply1_c) and :126-145
 *                                (applyc_c): the two callables of the CPython module
 *                                `libxgates` that src/lib/circuit.py:36-41 binds.
 *   qb_xg_apply1/applyc          the kernels behind them, src/lib/xgates.cc:23-41 /
 *                                :45-67 (== src/lib/state.py:80-125), on a state
 *                                that stays resident in HBM.
 *   qb_state_create/destroy      src/libq/qureg.cc:11-62 (new_qureg/delete_qureg) and
 *                                src/lib/circuit.py:125-129 (qc.reg).
 *   qb_apply1/applyc/applycc,    src/libq/apply.cc:78-176 (libq_gate1) and the named
 *   qb_apply_gates               gates of src/libq/gates.cc:9-146 (h x y z t v yroot
 *                                u1 cu1 cx cz ccx cv cv_adj walsh).
 *   qb_flush                     src/libq/gates.cc:148-150 (flush) and the queue/flush
 *                                protocol of src/libq/gates_jit.cc:53-132.
 *   qb_list_above, qb_norm2,     src/libq/qureg.cc:64-86 (print_qureg[_stats]) and the
 *   qb_argmax, qb_prob_bit,      numpy readouts of src/lib/state.py:30-78
 *   qb_get_amplitude, qb_copy_*  (ampl/prob/maxprob) on qc.psi.
 *
 * Conventions
 *   - Amplitudes are complex128, interleaved (re, im), index i in [0, 2^n).
 *   - "bit" arguments are positions in that index, LSB = bit 0.  This is libq's
 *     qubit numbering (libq.h:35-40).  The Python face numbers qubits MSB-first
 *     (state.py:85): python qubit q == bit n-1-q; the qb_xg_* entry points take
 *     python numbering and do that conversion (including the negative-control
 *     behaviour circuit_test.py:94-104 relies on).
 *   - A 2x2 gate is 8 doubles: a.re a.im b.re b.im c.re c.im d.re d.im for
 *     | a b ; c d | (xgates.cc:18-22).
 *   - Gate calls are asynchronous and may be queued and fused; they take effect
 *     no later than the next qb_flush / qb_sync / readout on the same state.
 *   - All functions return QB_OK (0) or a negative QB_ERR_* code; they never
 *     exit the process (xgates.cc:28-32 does) and never throw.
 *   - One state may be used from one host thread at a time.
 */
#ifndef QCC_B200_H_
#define QCC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB_OK 0
#define QB_ERR_ARG (-1)      /* bad argument (qubit out of range, null pointer...) */
#define QB_ERR_CUDA (-2)     /* CUDA runtime error / no device: see qb_last_error() */
#define QB_ERR_NOMEM (-3)    /* state does not fit */
#define QB_ERR_COMM (-4)     /* multi-GPU communicator error */
#define QB_ERR_UNSUPPORTED (-5)

#define QB_ABI_VERSION 2

typedef struct qb_state qb_state; /* opaque */

/* One queued gate.  ctl_mask: index bits that must all be 1 for the gate to act
 * (0 = uncontrolled; 1 bit = cx/cu1-style; 2 bits = ccx-style). */
typedef struct qb_gate {
  uint64_t ctl_mask;
  int32_t target;   /* index bit the 2x2 acts on */
  int32_t flags;    /* reserved, 0 */
  double m[8];
} qb_gate;

/* Gate record in the reference's python numbering: what circuit.py:180-215 passes
 * to xgates per call.  kind 1 = apply1 (ctl ignored), 2 = applyc. */
typedef struct qb_xg_gate {
  int32_t kind;
  int32_t ctl;
  int32_t tgt;
  int32_t pad;
  double m[8];
} qb_xg_gate;

/* Engine counters (monotonic per state). */
typedef struct qb_counters {
  uint64_t gates_applied;    /* gate records executed */
  uint64_t kernel_launches;  /* CUDA kernels launched on the state's stream */
  uint64_t passes;           /* HBM sweeps (fused pass or single-gate kernel) */
  uint64_t bytes_algorithmic;/* sum over gates of SURVEY 8(d) bytes */
  uint64_t bytes_swept;      /* bytes the launched passes read+write by construction */
  uint64_t exchanges;        /* multi-GPU shard exchanges */
  uint64_t bytes_exchanged;
} qb_counters;

/* Per-kernel-class device time, filled while profiling is enabled. */
#define QB_KCLASS_APPLY1 0   /* single general gate, strided butterfly */
#define QB_KCLASS_PHASE 1    /* single diagonal gate */
#define QB_KCLASS_FUSED 2    /* tile-resident fused pass */
#define QB_KCLASS_AUX 3      /* init / reductions / compaction */
#define QB_KCLASS_EXCHANGE 4 /* multi-GPU exchange on its own: NCCL send/recv + copy-back, in-place pair
                              * swap or out-of-place push kernel, and the barriers that fence them */
#define QB_KCLASS_FUSED_PUSH 5 /* fused pass whose store stage carries an exchange event (peer writes over NVLink) */
#define QB_KCLASS_COUNT 6
typedef struct qb_profile {
  uint64_t launches[QB_KCLASS_COUNT];
  double ms[QB_KCLASS_COUNT];          /* summed CUDA-event durations */
  double bytes[QB_KCLASS_COUNT];       /* algorithmic bytes of those launches */
} qb_profile;

/* ---- library ---------------------------------------------------------------- */
int qb_abi_version(void);
const char *qb_last_error(void);
int qb_device_count(int *count);
/* name: >= 128 bytes */
int qb_device_info(int device, char *name, size_t name_len, int *sm_count, size_t *total_mem,
                   int *cc_major, int *cc_minor);

/* ---- state lifetime (qureg.cc:11-62, circuit.py:125-164) -------------------- */
/* Dense 2^nqubits complex128 vector on `device`, initialised to basis state |init_label>. */
int qb_state_create(int nqubits, uint64_t init_label, int device, qb_state **out);
/* Multi-GPU: the state is split across `nranks` (power of two) processes, one GPU each, by its
 * top log2(nranks) index bits.  Every rank calls this with the same arguments except device /
 * rank; id128 is the 128-byte communicator id rank 0 obtained from qb_comm_get_unique_id and
 * passed to the others out of band (bench.py: torch.distributed broadcast).  On a sharded state
 * every gate call, flush and readout is COLLECTIVE: all ranks must make the same calls in the
 * same order.  Gates whose target is a sharded bit trigger an exchange event (sharded bits
 * swapped with local ones: one all-to-all over NVLink peer memory written by the store stage of the
 * preceding fused pass; ncclSend/ncclRecv pair exchanges where peer mappings are unavailable) and a
 * logical->physical bit remap; diagonal gates and controls on sharded bits need no communication.  qb_copy_in/out address this rank's
 * slice of the canonical vector (they undo the bit remap first, collectively); qb_list_above returns the
 * same merged listing of the whole state on every rank. */
int qb_comm_get_unique_id(void *id128);
int qb_state_create_sharded(int nqubits, uint64_t init_label, int device, int rank, int nranks,
                            const void *id128, qb_state **out);
/* nlocal: index bits held per rank; perm[b] (nqubits ints): physical bit of logical bit b. */
int qb_state_layout(qb_state *s, int *nlocal, int *rank, int *nranks, int *perm);
/* How exchange events run on this state: 0 ncclSend/ncclRecv per pair, 1 in-place peer-memory swap per
 * pair, 2 push (double-buffered, one all-to-all per event fused into the preceding pass); -1 not sharded.
 * QCC_B200_EXCHANGE=nccl|swap|push asks for one; the outcome is agreed on by all ranks at creation. */
int qb_state_exchange_mode(qb_state *s, int *mode);
/* Exchanges / local bit swaps that bring perm back to the identity. */
int qb_canonicalize(qb_state *s);
int qb_state_destroy(qb_state *s);
int qb_state_nqubits(qb_state *s, int *nqubits);
int qb_set_basis(qb_state *s, uint64_t label);
/* Deterministic pseudo-random normalised state (benchmarks / property tests). */
int qb_fill_random(qb_state *s, uint64_t seed);
/* Host <-> device copies of a contiguous index range; buffers are complex128. */
int qb_copy_in(qb_state *s, uint64_t first, uint64_t count, const double *host);
int qb_copy_out(qb_state *s, uint64_t first, uint64_t count, double *host);

/* ---- gate application, index-bit numbering (libq.h:50-64, apply.cc:78) ------- */
int qb_apply1(qb_state *s, int target, const double m[8]);
int qb_applyc(qb_state *s, int control, int target, const double m[8]);
int qb_applycc(qb_state *s, int control0, int control1, int target, const double m[8]);
int qb_apply_gates(qb_state *s, const qb_gate *gates, int64_t ngates);

/* ---- gate application, python numbering (xgates.cc:23-67) -------------------- */
int qb_xg_apply1(qb_state *s, int tgt, const double m[8]);
int qb_xg_applyc(qb_state *s, int ctl, int tgt, const double m[8]);
int qb_xg_apply_gates(qb_state *s, const qb_xg_gate *gates, int64_t ngates);

/* ---- queue control (gates_jit.cc:123-127, gates.cc:148-150) ------------------ */
/* fusion = 1 (default): gate calls queue up and qb_flush plans tile-resident fused
 * passes; fusion = 0: every gate call launches its own kernel immediately. */
int qb_set_fusion(qb_state *s, int fusion);
int qb_flush(qb_state *s); /* plan + launch everything queued (asynchronous) */
int qb_sync(qb_state *s);  /* qb_flush + wait for the device */

/* ---- readouts (state.py:30-78, qureg.cc:64-86); all imply qb_flush ----------- */
int qb_get_amplitude(qb_state *s, uint64_t index, double out[2]);
int qb_norm2(qb_state *s, double *out);                       /* sum |psi_i|^2 */
int qb_argmax(qb_state *s, uint64_t *index, double *prob);    /* state.py:60-78 */
int qb_prob_bit(qb_state *s, int bit, double *p_one);         /* P(bit == 1) */
/* sum of |psi_i|^2 over the indices whose `bit` equals `value`: trace(P rho) of ops.py:426-460 for the
 * projector onto that value, computed directly (no 1 - p cancellation; states need not be normalised) */
int qb_prob_bit_value(qb_state *s, int bit, int value, double *p);
/* Sparse listing for print_qureg: every index with |psi|^2 >= threshold, ascending
 * index order.  *count receives the total number found; if it exceeds cap, an unspecified
 * subset of `cap` entries is stored (retry with a larger cap).  labels/amps may be NULL
 * with cap 0 to just count. */
int qb_list_above(qb_state *s, double threshold, uint64_t cap, uint64_t *labels, double *amps,
                  uint64_t *count);

/* ---- host-buffer entry points: what the `libxgates` module binds -------------
 * psi: 2^nbits complex64 (bit_width != 128) or complex128 (== 128), C-contiguous, updated
 * in place (H2D, kernel, D2H inside the call).  gate: 4 complex of the same type.
 * device < 0 selects the current/default device. */
int qb_host_apply1(void *psi, const void *gate, int nbits, int tgt, int bit_width, int device);
int qb_host_applyc(void *psi, const void *gate, int nbits, int ctl, int tgt, int bit_width,
                   int device);
/* Whole gate stream on a host complex128 buffer: one upload, fused passes, one download. */
int qb_host_run(void *psi, int nbits, const qb_xg_gate *gates, int64_t ngates, int device);

/* Page-locked host buffers, so callers (bench.py's e2e leg, the libq face) can hand the
 * host-buffer and copy entry points memory the DMA engines can stream at full PCIe rate. */
int qb_host_alloc(size_t bytes, void **out);
int qb_host_free(void *p);

/* ---- measurement helpers ------------------------------------------------------ */
int qb_get_counters(qb_state *s, qb_counters *out);
int qb_profile_enable(qb_state *s, int enable); /* CUDA events around every launch */
int qb_profile_read(qb_state *s, qb_profile *out, int reset); /* implies qb_sync */
/* CUDA-event stopwatch on the state's own stream. */
int qb_timer_start(qb_state *s);
int qb_timer_stop(qb_state *s, double *ms); /* implies qb_sync */

/* ---- planner introspection (host only; no GPU needed) ------------------------- */
/* Plans `ngates` gates for an n-qubit single-device state exactly as qb_flush would and
 * writes the fused-pass plan (tile bits, rounds, ops, ladder tables) as JSON text.
 * *needed receives the text length including the NUL; call with cap 0 to size the buffer.
 * The CPU tests interpret this plan with numpy to check the planner without a GPU. */
int qb_plan_json(int nqubits, const qb_gate *gates, int64_t ngates, int tile_bits, char *buf,
                 size_t cap, size_t *needed);
/* The per-rank lowering of a gate stream for a sharded state (local gate batches + exchange
 * steps + final bit permutation) as JSON; canonicalize != 0 appends the steps that restore the
 * identity layout.  Host only: the CPU tests execute it with numpy shards + gloo send/recv. */
int qb_shard_lower_json(int nqubits, int nranks, int rank, const qb_gate *gates, int64_t ngates,
                        int canonicalize, char *buf, size_t cap, size_t *needed);
/* Plans `ngates` gates as qb_flush would (peephole included) and runs every fused pass through the HOST half of
 * its kernel launch (capacity of the parameter block and of shared memory, address terms): QB_ERR_UNSUPPORTED if
 * the planner produced a pass the kernel cannot take.  No GPU needed; the CPU tests run it over every golden. */
int qb_plan_check(int nqubits, const qb_gate *gates, int64_t ngates, int tile_bits, int64_t *passes);
/* The queue's peephole, applied in place to a gate list (host only): every adjacent five-gate
 * Sleator-Weinfurter run cu(a,t,V) cx(a,b) cu(b,t,V^dagger) cx(a,b) cu(b,t,V) -- how circuit.py:227-246 spells
 * ccx / ccu / ccu1 -- becomes ONE doubly-controlled V^2 on t followed by four identity gates (so the gate count is
 * kept).  *fused receives the number of runs replaced.  qb_flush does this before planning unless
 * QCC_B200_NO_CCU_FUSE is set. */
int qb_fuse_gates(qb_gate *gates, int64_t ngates, int64_t *fused);
/* Counts only, same lowering + planning as a flush of these gates on rank `rank` (identity layout to start
 * with): stats[0] exchange events, [1] exchanged (sharded bit, local bit) pairs, [2] passes, [3] of them fused
 * passes, [4] events that ride on the store stage of a fused pass, [5] rounds, [6] ops, [7] Sleator-Weinfurter runs fused.  window / hoist /
 * prefetch are the lowering parameters (push exchange: nlocal - 5 capped at nlocal - 3, 1, 1; NCCL: 6, 0, 0).
 * Host only. */
int qb_shard_plan_stats(int nqubits, int nranks, int rank, const qb_gate *gates, int64_t ngates, int tile_bits,
                        int window, int hoist, int prefetch, int64_t stats[8]);
/* Where the push exchange of engine.cu writes: for one exchange event (npairs pairs: local victim bit -> rank
 * bit, rank bit -> landing bit, landing bit -> victim bit; lands == NULL or lands[k] == victims[k]: plain swap)
 * and each local index of `rank`'s shard, the index in the DISTRIBUTED vector (destination rank << nlocal |
 * destination local index).  Host only; the CPU tests check it against the pairwise send/recv layout. */
int qb_shard_event_dest(int nlocal, int nranks, int rank, const int *rank_bits, const int *victims, const int *lands,
                        int npairs, const uint64_t *local, uint64_t *dest, int64_t count);
/* Tile size (log2 amplitudes per CTA tile, 4..13, default 12) used by qb_flush. */
int qb_set_tile_bits(qb_state *s, int tile_bits);

#ifdef __cplusplus
}
#endif
#endif /* QCC_B200_H_ */
