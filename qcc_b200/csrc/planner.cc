// planner.cc -- see planner.h / qb_types.h for the pass format.
//
// Pipeline:  gates -> merged gates -> items (plain gate | phase ladder) -> passes -> rounds
//            -> ops.
//
//  0. 2x2 pre-multiplication.  A gate is multiplied into an earlier gate with the same
//     target and controls when everything in between commutes with it trivially (touches
//     neither its target nor has its target among their bits): h(q); v(q) of
//     larose_benchmark.py:50-51 becomes one general 2x2.
//  1. Ladder merge.  Consecutive PHASE gates (they all commute) on <= 2 bits that share
//     one "pivot" bit are merged into a LADDER item: for every index with the pivot
//     set, multiply by the product of the partner phases whose bit is set.  This is the
//     controlled-phase ladder that follows each h in circuit.py:320-328 (qft): up to
//     n-1 cu1 gates become one op whose phase is looked up in two small tables.
//  2. Pass cut.  Only U / PERM / SWAP items need their target inside the tile (they mix
//     two amplitudes); PHASE / DIAG / LADDER items are diagonal and can run in any tile.
//     A pass greedily collects items until it would need more than K-3 distinct targets
//     above bit 2.  Tile bits = {0,1,2} + targets, padded with the lowest unused bits so
//     that tiles are as contiguous in memory as possible.
//  3. Round cut.  Inside a pass, ops are grouped while their targets fit in 3 tile-local bits
//     (8 amplitudes per thread in registers).  The items of a pass are list-scheduled over
//     their commutation DAG (schedule_rounds): uncontrolled U's on distinct bits are collected
//     three to a round (the predicate-free UX round program of fused.cu), diagonal items ride along
//     wherever they are ready, and a bit that many later gates target (the cx fan-in of
//     larose_benchmark.py:52-53) is taken into a round only when that work can finish there.
//     x / cx gates onto one target that end up adjacent become ONE parity-controlled swap
//     (PARSWAP).  The in-order cut is kept as the fallback whenever scheduling does not
//     need fewer rounds.
#include "planner.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace qb {

namespace {

inline Cplx cmulh(Cplx a, Cplx b) { return Cplx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

struct Item {
  int kind;            // QbKind; LADDER for merged runs
  int64_t first_gate;  // index of the first source gate
  int64_t ngates;      // source gates merged into this item (incl. skipped no-ops)
  double bytes_per_amp;
  // plain gate
  QbGate g;
  // ladder
  int pivot = -1;
  std::vector<int> pbits;     // partner index bits
  std::vector<Cplx> pphase;   // their phases
  Cplx self{1.0, 0.0};        // phase applied whenever the pivot is set
};

bool needs_target(int kind) { return kind == QB_K_U || kind == QB_K_PERM || kind == QB_K_SWAP; }

uint64_t item_bits(const Item &it) {
  if (it.kind == QB_K_LADDER) {
    uint64_t b = uint64_t(1) << it.pivot;
    for (int p : it.pbits) b |= uint64_t(1) << p;
    return b;
  }
  return it.g.ctl_mask | (uint64_t(1) << it.g.target);
}

// Sufficient test.  An item is diagonal in every bit it touches except the target of a
// U / PERM / SWAP ("mix" bit).  Two items commute when neither mixes a bit the other touches; two
// pure X-type gates on the same target (x, cx, ccx) commute as well: X^f X^g = X^(f xor g).
bool items_commute(const Item &a, const Item &b) {
  const uint64_t ba = item_bits(a), bb = item_bits(b);
  const uint64_t ma = needs_target(a.kind) ? uint64_t(1) << a.g.target : 0;
  const uint64_t mb = needs_target(b.kind) ? uint64_t(1) << b.g.target : 0;
  if (!(ma & bb) && !(mb & ba)) return true;
  if (a.kind == QB_K_SWAP && b.kind == QB_K_SWAP && a.g.target == b.g.target) return true;
  return false;
}

// x / cx: a pure swap with at most one control -- the building block of a PARSWAP op
bool is_parity_x(const Item &it) { return it.kind == QB_K_SWAP && __builtin_popcountll(it.g.ctl_mask) <= 1; }

struct Merged {
  QbGate g;
  int64_t first_gate;
  int64_t ngates;
  double bytes_per_amp;
};

// --- 0. pre-multiplication of 2x2s on the same (controls, target) -----------------
void premultiply(const QbGate *gates, int64_t ng, bool enable, std::vector<Merged> *out) {
  for (int64_t i = 0; i < ng; ++i) {
    const QbGate &g = gates[i];
    if (g.kind == QB_K_NOP) {
      if (!out->empty()) out->back().ngates += 1;
      else out->push_back(Merged{g, i, 1, 0.0});
      continue;
    }
    const uint64_t gbits = g.ctl_mask | (uint64_t(1) << g.target);
    bool merged = false;
    if (enable) {
      int looked = 0;
      for (auto it = out->rbegin(); it != out->rend() && looked < 64; ++it, ++looked) {
        QbGate &e = it->g;
        if (e.kind == QB_K_NOP) continue;
        const uint64_t ebits = e.ctl_mask | (uint64_t(1) << e.target);
        if (e.target == g.target && e.ctl_mask == g.ctl_mask) {
          // new = g * e  (e acts first)
          Cplx a[4], b[4], r[4];
          for (int k = 0; k < 4; ++k) {
            a[k] = Cplx{g.m[2 * k], g.m[2 * k + 1]};
            b[k] = Cplx{e.m[2 * k], e.m[2 * k + 1]};
          }
          for (int row = 0; row < 2; ++row)
            for (int col = 0; col < 2; ++col) {
              Cplx p = cmulh(a[row * 2 + 0], b[0 * 2 + col]);
              Cplx q = cmulh(a[row * 2 + 1], b[1 * 2 + col]);
              r[row * 2 + col] = Cplx{p.x + q.x, p.y + q.y};
            }
          for (int k = 0; k < 4; ++k) {
            e.m[2 * k] = r[k].x;
            e.m[2 * k + 1] = r[k].y;
          }
          e.kind = classify_matrix(e.m);
          it->ngates += 1;
          it->bytes_per_amp += gate_bytes_per_amp(g);
          merged = true;
          break;
        }
        // can g move in front of e?  only if neither changes a bit the other looks at
        if ((gbits >> e.target & 1) || (ebits >> g.target & 1)) break;
      }
    }
    if (!merged) out->push_back(Merged{g, i, 1, gate_bytes_per_amp(g)});
  }
}

// --- 1. ladder merge -------------------------------------------------------------
void build_items(const std::vector<Merged> &mg, bool ladders, std::vector<Item> *items) {
  struct Open {
    bool active = false;
    int64_t first = 0, count = 0;
    double bytes = 0;
    uint64_t cand = 0;  // candidate pivot bits
    std::vector<std::pair<uint64_t, Cplx>> parts;  // (bit set of the gate, phase)
    QbGate first_gate;
  } open;

  auto close = [&]() {
    if (!open.active) return;
    Item it;
    it.first_gate = open.first;
    it.ngates = open.count;
    it.bytes_per_amp = open.bytes;
    if (open.parts.size() == 1) {
      it.kind = QB_K_PHASE;
      it.g = open.first_gate;
    } else {
      it.kind = QB_K_LADDER;
      it.pivot = __builtin_ctzll(open.cand);
      for (auto &p : open.parts) {
        uint64_t rest = p.first & ~(uint64_t(1) << it.pivot);
        if (rest == 0) {
          it.self = cmulh(it.self, p.second);
        } else {
          int b = __builtin_ctzll(rest);
          size_t k = 0;
          for (; k < it.pbits.size(); ++k)
            if (it.pbits[k] == b) break;
          if (k == it.pbits.size()) {
            it.pbits.push_back(b);
            it.pphase.push_back(p.second);
          } else {
            it.pphase[k] = cmulh(it.pphase[k], p.second);
          }
        }
      }
    }
    items->push_back(std::move(it));
    open = Open();
  };

  for (size_t mi = 0; mi < mg.size(); ++mi) {
    const QbGate &g = mg[mi].g;
    const int64_t i = mg[mi].first_gate;
    const int64_t cnt = mg[mi].ngates;
    const double gbytes = mg[mi].bytes_per_amp;
    if (g.kind == QB_K_NOP) {
      // retire it with whatever item is open / came before; costs nothing
      if (open.active) {
        open.count += cnt;
      } else if (!items->empty()) {
        items->back().ngates += cnt;
      } else {
        Item it;
        it.kind = QB_K_NOP;
        it.first_gate = i;
        it.ngates = cnt;
        it.bytes_per_amp = 0;
        it.g = g;
        items->push_back(it);
      }
      continue;
    }
    uint64_t bits = g.ctl_mask | (uint64_t(1) << g.target);
    if (ladders && g.kind == QB_K_PHASE && __builtin_popcountll(bits) <= 2) {
      Cplx ph{g.m[6], g.m[7]};
      if (open.active && (open.cand & bits)) {
        open.cand &= bits;
        open.parts.push_back({bits, ph});
        open.count += cnt;
        open.bytes += gbytes;
        continue;
      }
      close();
      open.active = true;
      open.first = i;
      open.count = cnt;
      open.bytes = gbytes;
      open.cand = bits;
      open.parts.push_back({bits, ph});
      open.first_gate = g;
      continue;
    }
    close();
    Item it;
    it.kind = g.kind;
    if (g.kind == QB_K_PERM && g.m[2] == 1.0 && g.m[3] == 0.0 && g.m[4] == 1.0 && g.m[5] == 0.0)
      it.kind = QB_K_SWAP;
    it.first_gate = i;
    it.ngates = cnt;
    it.bytes_per_amp = gbytes;
    it.g = g;
    items->push_back(it);
  }
  close();
}

// --- helpers ---------------------------------------------------------------------
struct TileMap {
  int K = 0;
  int bits[QB_MAX_TILE_BITS + 3];
  int lpos[64];  // index bit -> tile-local position or -1
  uint64_t mask = 0;
};

TileMap make_tile(int nbits, int K, const std::vector<int> &targets_hi) {
  TileMap t;
  std::vector<int> b;
  int low = std::min(QB_TILE_LOW, K);
  for (int i = 0; i < low; ++i) b.push_back(i);
  for (int x : targets_hi) b.push_back(x);
  // pad with the lowest unused bits: keeps each tile as contiguous as possible
  for (int i = low; i < nbits && int(b.size()) < K; ++i)
    if (std::find(b.begin(), b.end(), i) == b.end()) b.push_back(i);
  std::sort(b.begin(), b.end());
  t.K = int(b.size());
  for (int i = 0; i < 64; ++i) t.lpos[i] = -1;
  for (int k = 0; k < t.K; ++k) {
    t.bits[k] = b[k];
    t.lpos[b[k]] = k;
    t.mask |= uint64_t(1) << b[k];
  }
  return t;
}

struct PendingOp {
  const Item *it;
  int variant;  // DIAG: 0 -> d0 where target clear, 1 -> d1 where target set; else 0
};

// Split "index bits in `mask` must equal `want`" into the three predicate levels.
void split_pred(const TileMap &tm, const int *rbit, int nr, uint64_t mask, uint64_t want, QbOp *op) {
  op->lmask = op->lwant = op->rmask = op->rwant = 0;
  op->gmask = op->gwant = 0;
  for (int b = 0; b < 64; ++b) {
    if (!(mask >> b & 1)) continue;
    uint64_t w = want >> b & 1;
    int lp = tm.lpos[b];
    if (lp < 0) {
      op->gmask |= uint64_t(1) << b;
      op->gwant |= w << b;
      continue;
    }
    int rp = -1;
    for (int k = 0; k < nr; ++k)
      if (rbit[k] == lp) rp = k;
    if (rp >= 0) {
      op->rmask |= 1u << rp;
      op->rwant |= uint32_t(w) << rp;
    } else {
      op->lmask |= 1u << lp;
      op->lwant |= uint32_t(w) << lp;
    }
  }
}

// How the tile NUMBER (bit i = i-th index bit outside the tile, ascending) is cut into three fields for
// the per-tile ladder constants: widths w0 >= w1 >= w2, w0 + w1 + w2 = nbits - K.
void ladder_field_widths(int nbits, int K, int w[3]) {
  const int nt = nbits - K;
  w[0] = (nt + 2) / 3;
  w[1] = (nt - w[0] + 1) / 2;
  w[2] = nt - w[0] - w[1];
}

void build_ladder_tables(const TileMap &tm, const QbRound &r, const Item &it, int slot, int nbits, PlannedPass *pp,
                         QbOp *op) {
  const int K = tm.K;
  // per tile-local position phase (1 where the position is not a partner / the pivot)
  Cplx per[QB_MAX_TILE_BITS];
  for (int k = 0; k < K; ++k) per[k] = Cplx{1.0, 0.0};
  std::vector<int> out_bits;
  std::vector<Cplx> out_ph;
  for (size_t k = 0; k < it.pbits.size(); ++k) {
    int lp = tm.lpos[it.pbits[k]];
    if (lp >= 0) per[lp] = cmulh(per[lp], it.pphase[k]);
    else {
      out_bits.push_back(it.pbits[k]);
      out_ph.push_back(it.pphase[k]);
    }
  }
  int plp = tm.lpos[it.pivot];
  // The pivot's own phase (a u1 merged into the ladder) applies to every touched amplitude.  It
  // goes into the lookup tables only when the pivot is a tile bit outside the round; otherwise
  // (pivot outside the tile, or a round bit) it is folded into the per-tile constant, which keeps
  // F[pivot bit only] exactly 1 -- the Hadamard+ladder op relies on that.
  Cplx self_out = it.self;
  bool pivot_in_round = false;
  for (int k = 0; k < r.nbits; ++k)
    if (plp >= 0 && r.rbit[k] == plp) pivot_in_round = true;
  if (plp >= 0 && !pivot_in_round) {
    per[plp] = cmulh(per[plp], it.self);
    self_out = Cplx{1.0, 0.0};
  }
  // Table layout (double2 units from table_off), indexed by the GROUP number q of the op's round
  // (group-index bit k drives tile-local position r.qmap[k]; the round's own bits are not group
  // bits, so no entry is wasted on them):
  //   [0, 32)                     T_a[q & 31]  -- one entry per lane: conflict-free reads
  //   [32, 32 + 2^(K-3-5))        T_b[q >> 5]  -- uniform per warp: broadcast reads; the kernel
  //                                               folds the per-tile constant into its copy
  // The 8 combinations of the round's own bits, F[e], travel inside the op descriptor (constant
  // bank).  The per-tile constant -- the product of the phases of the partner bits OUTSIDE the tile that
  // are set in the tile's base, times the self phase when the pivot is outside the tile -- is looked up
  // too: the tile number is cut into three fields (ladder_field_widths) and a second array (outph, from
  // outph_off) holds one small table per field, C_0[2^w0] (carrying the constant factor), C_1[2^w1],
  // C_2[2^w2]; the kernel multiplies three entries per tile and ladder instead of walking the bits
  // (18 outside bits at 30 qubits: 3 x 64 entries).
  op->table_off = int32_t(pp->tables.size());
  const int gbits = K - r.nbits;                       // group-index bits
  const int a_bits = std::min(gbits, QB_LADDER_LANE_BITS);
  const int b_bits = gbits - a_bits;
  for (int v = 0; v < (1 << QB_LADDER_LANE_BITS); ++v) {
    Cplx p{1.0, 0.0};
    for (int k = 0; k < a_bits; ++k)
      if (v >> k & 1) p = cmulh(p, per[r.qmap[k]]);
    pp->tables.push_back(p);
  }
  for (int v = 0; v < (1 << b_bits); ++v) {
    Cplx p{1.0, 0.0};
    for (int k = 0; k < b_bits; ++k)
      if (v >> k & 1) p = cmulh(p, per[r.qmap[a_bits + k]]);
    pp->tables.push_back(p);
  }
  for (int e = 0; e < 8; ++e) {
    Cplx p{1.0, 0.0};
    for (int k = 0; k < r.nbits; ++k)
      if (e >> k & 1) p = cmulh(p, per[r.rbit[k]]);
    op->F[2 * e] = p.x;
    op->F[2 * e + 1] = p.y;
  }
  op->outph_off = int32_t(pp->outph.size());
  {
    std::vector<Cplx> per_t;   // phase of tile-number bit i
    for (int b = 0; b < nbits; ++b) {
      if (tm.mask >> b & 1) continue;
      Cplx ph{1.0, 0.0};
      for (size_t k = 0; k < out_bits.size(); ++k)
        if (out_bits[k] == b) ph = cmulh(ph, out_ph[k]);
      per_t.push_back(ph);
    }
    int w[3];
    ladder_field_widths(nbits, K, w);
    int first = 0;
    for (int f = 0; f < 3; ++f) {
      for (int v = 0; v < (1 << w[f]); ++v) {
        Cplx p = f == 0 ? self_out : Cplx{1.0, 0.0};
        for (int k = 0; k < w[f]; ++k)
          if (v >> k & 1) p = cmulh(p, per_t[size_t(first + k)]);
        pp->outph.push_back(p);
      }
      first += w[f];
    }
  }
  op->nout = int32_t(out_bits.size());
  op->out_off = int32_t(pp->outbits.size());
  for (int b : out_bits) pp->outbits.push_back(b);   // kept for the plan dump; the kernel reads the tables
  op->flags = slot;
}

// One round as collected by emit_pass before the group maps are assigned.
struct RoundPlan {
  std::vector<int> rset;        // round bits (tile-local positions), padded to nr and sorted
  std::vector<PendingOp> pend;
  std::vector<int> order;       // group-index bit k -> tile-local position (qmap)
  bool nobar = false;           // the next round works on the same per-warp sub-cube: no CTA barrier after
};

void pad_round_bits(const TileMap &tm, std::vector<int> *rset) {
  const int nr = std::min(tm.K, QB_ROUND_BITS);
  // pad the round's bit set with the lowest unused local positions
  for (int k = 0; k < tm.K && int(rset->size()) < nr; ++k)
    if (std::find(rset->begin(), rset->end(), k) == rset->end()) rset->push_back(k);
  std::sort(rset->begin(), rset->end());
}

// `cand` in the order they should be used, except that the first three are one bit of each
// swizzle class when available.  The tile is stored swizzled (fused.cu): the 16-byte bank group of
// local index j is (j ^ j>>3 ^ j>>6 ^ j>>9 ...) & 7, so local bit b feeds bank-group bit b % 3;
// giving group-index bits 0,1,2 one bit of each class makes 8 consecutive lanes land in 8 distinct
// bank groups.
std::vector<int> class_first(std::vector<int> cand) {
  std::vector<int> order;
  for (int cls = 0; cls < 3; ++cls)
    for (size_t k = 0; k < cand.size(); ++k)
      if (cand[k] >= 0 && cand[k] % 3 == cls) {
        order.push_back(cand[k]);
        cand[k] = -1;
        break;
      }
  for (int b : cand)
    if (b >= 0) order.push_back(b);
  return order;
}

// Group maps.  Default: the free local bits in class-first order.  For tiles with >= 256 groups
// (K >= 11) consecutive rounds are collected into RUNS that share one (K-3)-bit sub-cube S of the
// tile containing all their round bits: the group-index bits that enumerate a warp's work (lane
// bits 0..4 and the iteration bits 8..) are mapped into S, the warp number (bits 5..7) to the 3
// tile bits outside S.  Every warp then reads and writes the same 2^(K-3) amplitudes in every round
// of the run, so the rounds of a run are separated by a warp-level sync only and the warps of a
// CTA drift apart instead of meeting at a barrier after every round.
// With >= 512 groups (K >= 12) every S also contains tile bits 0..2, i.e. a warp's sub-cube is made
// of whole 128-byte runs: the warp itself copies the sub-cube of the first run in from HBM and the
// sub-cube of the last run out (QbPassDesc::ld_map / st_map, warp_io), waiting only for its own
// copies -- the load, the rounds and the store of different warps of a CTA overlap.
void assign_group_maps(const TileMap &tm, std::vector<RoundPlan> *rounds, QbPassDesc *desc) {
  const int K = tm.K;
  const int nr = std::min(K, QB_ROUND_BITS);
  const int gbits = K - nr;
  desc->warp_io = 0;
  for (int k = 0; k < QB_MAX_TILE_BITS; ++k) desc->ld_map[k] = desc->st_map[k] = k;
  for (RoundPlan &rp : *rounds) {
    std::vector<int> freeb;
    for (int k = 0; k < K; ++k)
      if (std::find(rp.rset.begin(), rp.rset.end(), k) == rp.rset.end()) freeb.push_back(k);
    rp.order = class_first(freeb);
    rp.nobar = false;
  }
  if (gbits < 8) return;  // fewer than 256 groups: one barrier per round
  static const bool no_warp_io = getenv("QCC_B200_NO_WARP_IO") != nullptr;
  const bool low3 = gbits >= 9 && !no_warp_io && !rounds->empty();
  std::vector<int> first_S, last_S;
  for (size_t i = 0; i < rounds->size();) {
    std::vector<int> S;
    if (low3) S = {0, 1, 2};
    auto merged = [&](const std::vector<int> &base, const std::vector<int> &add) {
      std::vector<int> U = base;
      for (int b : add)
        if (std::find(U.begin(), U.end(), b) == U.end()) U.push_back(b);
      return U;
    };
    // pad a sub-cube to gbits bits: first whatever swizzle class (bit % 3) some round of the run would
    // otherwise miss among its free bits (its lanes 0..2 need one bit of each class for conflict-free
    // shared-memory access), then the lowest tile bits
    auto padded = [&](std::vector<int> base, size_t r0, size_t r1) {
      for (size_t r = r0; r < r1; ++r)
        for (int cls = 0; cls < 3 && int(base.size()) < gbits; ++cls) {
          bool have = false;
          for (int b : base)
            if (b % 3 == cls && std::find((*rounds)[r].rset.begin(), (*rounds)[r].rset.end(), b) == (*rounds)[r].rset.end())
              have = true;
          for (int k = 0; k < K && !have; ++k)
            if (k % 3 == cls && std::find(base.begin(), base.end(), k) == base.end()) {
              base.push_back(k);
              have = true;
            }
        }
      for (int k = 0; k < K && int(base.size()) < gbits; ++k)
        if (std::find(base.begin(), base.end(), k) == base.end()) base.push_back(k);
      return base;
    };
    auto classes_ok = [&](const std::vector<int> &full, size_t r0, size_t r1) {
      for (size_t r = r0; r < r1; ++r) {
        int seen = 0;
        for (int b : full)
          if (std::find((*rounds)[r].rset.begin(), (*rounds)[r].rset.end(), b) == (*rounds)[r].rset.end()) seen |= 1 << (b % 3);
        if (seen != 7) return false;
      }
      return true;
    };
    S = merged(S, (*rounds)[i].rset);
    size_t j = i + 1;
    for (; j < rounds->size(); ++j) {
      std::vector<int> U = merged(S, (*rounds)[j].rset);
      if (int(U.size()) > gbits) break;
      // a longer run must not cost any of its rounds its conflict-free lane bits
      if (!classes_ok(padded(U, i, j + 1), i, j + 1) && classes_ok(padded(S, i, j), i, j)) break;
      S = U;
    }
    if (j - i >= 2 || low3) {
      S = padded(S, i, j);
      std::sort(S.begin(), S.end());
      std::vector<int> W;
      for (int k = 0; k < K; ++k)
        if (std::find(S.begin(), S.end(), k) == S.end()) W.push_back(k);
      for (size_t r = i; r < j; ++r) {
        RoundPlan &rp = (*rounds)[r];
        std::vector<int> in_s;
        for (int b : S)
          if (std::find(rp.rset.begin(), rp.rset.end(), b) == rp.rset.end()) in_s.push_back(b);
        in_s = class_first(in_s);  // gbits - 3 bits: 5 lane bits, then the iteration bits
        std::vector<int> order(size_t(gbits), -1);
        size_t u = 0;
        for (int k = 0; k < 5; ++k) order[size_t(k)] = in_s[u++];
        for (int k = 8; k < gbits; ++k) order[size_t(k)] = in_s[u++];
        for (int k = 5; k < 8; ++k) order[size_t(k)] = W[size_t(k - 5)];
        rp.order = order;
        rp.nobar = r + 1 < j;
      }
      if (i == 0) first_S = S;
      if (j == rounds->size()) last_S = S;
    }
    i = j;
  }
  if (low3) {
    // copy index bits: 0..4 (lane) -> S[0..4] (S[0..2] = positions 0..2), 5..7 (warp) -> W, 8.. -> S[5..]
    auto fill = [&](const std::vector<int> &S, int32_t *map) {
      std::vector<int> W;
      for (int k = 0; k < K; ++k)
        if (std::find(S.begin(), S.end(), k) == S.end()) W.push_back(k);
      for (int k = 0; k < 5; ++k) map[k] = S[size_t(k)];
      for (int k = 5; k < 8; ++k) map[k] = W[size_t(k - 5)];
      for (int k = 8; k < K; ++k) map[k] = S[size_t(k - 3)];
    };
    fill(first_S, desc->ld_map);
    fill(last_S, desc->st_map);
    desc->warp_io = 1;
  }
}

// Round cut in program order: a new round whenever a fourth target bit shows up.
std::vector<RoundPlan> inorder_rounds(const TileMap &tm, const std::vector<const Item *> &items) {
  std::vector<RoundPlan> out;
  RoundPlan cur;
  const int nr = std::min(tm.K, QB_ROUND_BITS);
  auto flush_round = [&]() {
    if (!cur.pend.empty()) {
      pad_round_bits(tm, &cur.rset);
      out.push_back(cur);
    }
    cur = RoundPlan();
  };
  for (const Item *it : items) {
    if (needs_target(it->kind)) {
      int lp = tm.lpos[it->g.target];
      if (std::find(cur.rset.begin(), cur.rset.end(), lp) == cur.rset.end()) {
        if (int(cur.rset.size()) == nr) flush_round();
        cur.rset.push_back(lp);
      }
    }
    cur.pend.push_back(PendingOp{it, 0});
    if (it->kind == QB_K_DIAG) cur.pend.push_back(PendingOp{it, 1});
  }
  flush_round();
  return out;
}

// List scheduling over the commutation DAG of the pass's items.  A round is grown from the ready
// items: everything ready that fits the round's bits is absorbed (diagonal items always fit); a new
// round bit is taken from the ready item whose bit strands the least work -- items on that bit that
// could not finish inside this round and would force the bit into a later round again -- with
// uncontrolled U's preferred (they are what the predicate-free UX round program runs; those on tile
// positions 0..2 first) and program order as the tie break, so a stream without freedom (the QFT)
// comes out exactly as written.
std::vector<RoundPlan> schedule_rounds(const TileMap &tm, const std::vector<const Item *> &items) {
  const int n = int(items.size());
  const int nr = std::min(tm.K, QB_ROUND_BITS);
  std::vector<std::vector<int>> preds;
  preds.resize(size_t(n));
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < j; ++i)
      if (!items_commute(*items[size_t(i)], *items[size_t(j)])) preds[size_t(j)].push_back(i);
  std::vector<char> done(static_cast<size_t>(n), 0);
  std::vector<int> lp_of(static_cast<size_t>(n), -1);
  for (int j = 0; j < n; ++j)
    if (needs_target(items[size_t(j)]->kind)) lp_of[size_t(j)] = tm.lpos[items[size_t(j)]->g.target];
  int left = n;
  std::vector<RoundPlan> out;
  while (left > 0) {
    std::vector<int> rset, picked;
    std::vector<char> in_round(static_cast<size_t>(n), 0);
    auto in_rset = [&](const std::vector<int> &rs, int lp) { return std::find(rs.begin(), rs.end(), lp) != rs.end(); };
    auto ready = [&](int j) {
      for (int i : preds[size_t(j)])
        if (!done[size_t(i)]) return false;
      return true;
    };
    auto take = [&](int j) {
      done[size_t(j)] = 1;
      in_round[size_t(j)] = 1;
      picked.push_back(j);
      --left;
    };
    for (;;) {
      // absorb everything ready that fits, in program order, until nothing moves
      for (bool moved = true; moved;) {
        moved = false;
        for (int j = 0; j < n; ++j) {
          if (done[size_t(j)] || !ready(j)) continue;
          if (lp_of[size_t(j)] >= 0 && !in_rset(rset, lp_of[size_t(j)])) continue;
          take(j);
          moved = true;
        }
      }
      if (int(rset.size()) == nr) break;
      // candidates for a new round bit
      int best = -1, best_stranded = 0, best_cls = 0;
      for (int j = 0; j < n; ++j) {
        if (done[size_t(j)] || lp_of[size_t(j)] < 0 || !ready(j)) continue;
        const int lp = lp_of[size_t(j)];
        std::vector<int> rs = rset;
        rs.push_back(lp);
        // which undone items could finish in a round with bits rs (program order == topological)
        std::vector<char> can(static_cast<size_t>(n), 0);
        int stranded = 0;
        for (int k = 0; k < n; ++k) {
          if (done[size_t(k)]) continue;
          bool ok = lp_of[size_t(k)] < 0 || in_rset(rs, lp_of[size_t(k)]);
          for (int i : preds[size_t(k)])
            if (!done[size_t(i)] && !can[size_t(i)]) ok = false;
          can[size_t(k)] = ok;
          if (!ok && lp_of[size_t(k)] == lp) ++stranded;
        }
        const Item &it = *items[size_t(j)];
        // class 0: uncontrolled U on tile positions 0..2 -- taken first, so that these bits are used up in
        // the early rounds and the LAST round of the pass can store straight to HBM (st_direct needs its
        // round bits outside positions 0..2); class 1: other uncontrolled U's; class 2: the rest
        const bool free_u = it.kind == QB_K_U && it.g.ctl_mask == 0;
        const int cls = free_u ? (lp < QB_TILE_LOW ? 0 : 1) : 2;
        if (best < 0 || stranded < best_stranded || (stranded == best_stranded && cls < best_cls)) {
          best = j;
          best_stranded = stranded;
          best_cls = cls;
        }
      }
      if (best < 0) break;
      rset.push_back(lp_of[size_t(best)]);
      take(best);
    }
    if (picked.empty()) break;  // cannot happen: the first undone item is always ready
    // Uncontrolled U's with no predecessor inside the round commute with everything scheduled
    // before them here: move them to the front, ascending by bit.  Rounds holding
    // ladders keep their order (a U must stay next to its ladder to fuse into one ULADDER op).
    bool has_ladder = false;
    for (int j : picked)
      if (items[size_t(j)]->kind == QB_K_LADDER) has_ladder = true;
    std::vector<int> order;
    if (!has_ladder) {
      std::vector<int> front;
      for (int j : picked) {
        const Item &it = *items[size_t(j)];
        bool free_u = it.kind == QB_K_U && it.g.ctl_mask == 0;
        for (int i : preds[size_t(j)])
          if (in_round[size_t(i)]) free_u = false;
        if (free_u) front.push_back(j);
      }
      std::sort(front.begin(), front.end(), [&](int a, int b) { return lp_of[size_t(a)] < lp_of[size_t(b)]; });
      order = front;
      for (int j : picked)
        if (std::find(front.begin(), front.end(), j) == front.end()) order.push_back(j);
    } else {
      order = picked;
    }
    RoundPlan rp;
    rp.rset = rset;
    for (int j : order) {
      rp.pend.push_back(PendingOp{items[size_t(j)], 0});
      if (items[size_t(j)]->kind == QB_K_DIAG) rp.pend.push_back(PendingOp{items[size_t(j)], 1});
    }
    pad_round_bits(tm, &rp.rset);
    out.push_back(std::move(rp));
  }
  return out;
}

void close_round(const TileMap &tm, RoundPlan &rp, int *nladders, int nbits, PlannedPass *pp) {
  std::vector<int> &rset = rp.rset;
  std::vector<PendingOp> &pend = rp.pend;
  const std::vector<int> &order = rp.order;
  QbRound r{};
  const int K = tm.K;
  const int nr = std::min(K, QB_ROUND_BITS);
  r.nbits = nr;
  for (int k = 0; k < nr; ++k) r.rbit[k] = rset[k];
  for (size_t k = 0; k < order.size(); ++k) r.qmap[k] = order[k];
  r.nobar = rp.nobar ? 1 : 0;
  // per-group base index and its swizzled shared-memory slot, so the kernel does one load
  // instead of a K-3 step bit scatter per group per round
  for (uint32_t q = 0; q < (1u << (K - 3)); ++q) {
    uint32_t jb = 0;
    for (int k = 0; k < K - 3; ++k) jb |= ((q >> k) & 1u) << order[size_t(k)];
    uint32_t x = jb >> 3;
    x ^= x >> 3;
    x ^= x >> 6;
    pp->jbtab.push_back(jb | ((jb ^ (x & 7u)) << 16));
  }
  r.op_begin = int32_t(pp->ops.size());
  // The tail of a QFT: "h + one cu1" and a bare h are not ladders, but when the whole round is three
  // Hadamards on round positions 0, 1, 2, each followed by at most one ladder / two-bit phase on its
  // own target, giving the short ones a one-partner / empty ladder lets the round run as the
  // unrolled Hadamard+ladder program instead of through the op interpreter.
  std::vector<Item> synth;
  synth.reserve(pend.size());
  std::vector<const Item *> lad_of(pend.size(), nullptr);  // per U op: its (possibly synthesized) ladder
  std::vector<char> skip(pend.size(), 0);
  {
    bool ok = nr == QB_ROUND_BITS;
    size_t idx = 0;
    std::vector<std::pair<size_t, size_t>> found;  // (U index, follower index or npos)
    for (int k = 0; ok && k < 3; ++k) {
      if (idx >= pend.size()) { ok = false; break; }
      const Item &u = *pend[idx].it;
      const double *m = u.g.m;
      const bool had = u.kind == QB_K_U && u.g.ctl_mask == 0 && m[1] == 0.0 && m[3] == 0.0 && m[5] == 0.0 &&
                       m[7] == 0.0 && m[0] == m[2] && m[0] == m[4] && m[0] == -m[6];
      if (!had || tm.lpos[u.g.target] != r.rbit[k]) { ok = false; break; }
      size_t fol = size_t(-1);
      if (idx + 1 < pend.size()) {
        const Item &f = *pend[idx + 1].it;
        const uint64_t fb = f.g.ctl_mask | (uint64_t(1) << f.g.target);
        if (f.kind == QB_K_LADDER && f.pivot == u.g.target) fol = idx + 1;
        else if (f.kind == QB_K_PHASE && __builtin_popcountll(fb) == 2 && (fb >> u.g.target & 1)) fol = idx + 1;
      }
      found.push_back({idx, fol});
      idx += fol == size_t(-1) ? 1 : 2;
    }
    if (ok && idx == pend.size()) {
      for (auto &uf : found) {
        const Item &u = *pend[uf.first].it;
        if (uf.second != size_t(-1) && pend[uf.second].it->kind == QB_K_LADDER) {
          lad_of[uf.first] = pend[uf.second].it;
        } else {
          Item lad;
          lad.kind = QB_K_LADDER;
          lad.pivot = u.g.target;
          if (uf.second != size_t(-1)) {
            const Item &f = *pend[uf.second].it;
            const uint64_t fb = (f.g.ctl_mask | (uint64_t(1) << f.g.target)) & ~(uint64_t(1) << u.g.target);
            lad.pbits.push_back(__builtin_ctzll(fb));
            lad.pphase.push_back(Cplx{f.g.m[6], f.g.m[7]});
          }
          synth.push_back(lad);
          lad_of[uf.first] = &synth.back();
        }
        if (uf.second != size_t(-1)) skip[uf.second] = 1;
      }
    }
  }
  // x / cx gates onto one target: each is merged into an earlier PARSWAP of this round on the same
  // target when it commutes with everything in between (larose_benchmark.py:52-53 -- after
  // scheduling, all cx(bit, 0) of a pass sit next to each other).
  struct Par {
    bool leader = false;
    uint64_t mask = 0;
    int flip = 0;
  };
  std::vector<Par> par(pend.size());
  {
    std::vector<size_t> emitted;
    for (size_t pi = 0; pi < pend.size(); ++pi) {
      if (skip[pi]) continue;
      const Item &it = *pend[pi].it;
      if (is_parity_x(it)) {
        bool merged = false;
        for (size_t k = emitted.size(); k-- > 0;) {
          const size_t e = emitted[k];
          const Item &ei = *pend[e].it;
          if (par[e].leader && ei.g.target == it.g.target) {
            par[e].mask ^= it.g.ctl_mask;
            par[e].flip ^= it.g.ctl_mask ? 0 : 1;
            skip[pi] = 1;
            merged = true;
            break;
          }
          if (par[e].leader) {
            // against the merged op, not just its first gate: X_te^(parity of mask)
            if ((par[e].mask >> it.g.target & 1) || (it.g.ctl_mask >> ei.g.target & 1)) break;
          } else if (!items_commute(ei, it)) {
            break;
          }
        }
        if (merged) continue;
        par[pi].leader = true;
        par[pi].mask = it.g.ctl_mask;
        par[pi].flip = it.g.ctl_mask ? 0 : 1;
      }
      emitted.push_back(pi);
    }
  }
  for (size_t pi = 0; pi < pend.size(); ++pi) {
    if (skip[pi]) continue;
    if (par[pi].leader && par[pi].mask == 0 && par[pi].flip == 0) continue;  // the x's cancelled
    const PendingOp &po = pend[pi];
    const Item &it = *po.it;
    QbOp op{};
    op.kind = it.kind;
    op.tpos = 0;
    if (par[pi].leader) {
      op.kind = QB_K_PARSWAP;
      int lp = tm.lpos[it.g.target];
      for (int k = 0; k < nr; ++k)
        if (r.rbit[k] == lp) op.tpos = k;
      for (int b = 0; b < 64; ++b) {
        if (!(par[pi].mask >> b & 1)) continue;
        int blp = tm.lpos[b];
        if (blp < 0) {
          op.gmask |= uint64_t(1) << b;
          continue;
        }
        int rp = -1;
        for (int k = 0; k < nr; ++k)
          if (r.rbit[k] == blp) rp = k;
        if (rp >= 0) op.rmask |= 1u << rp;
        else op.lmask |= 1u << blp;
      }
      op.rwant = uint32_t(par[pi].flip);
      for (int e = 0; e < 8; ++e)
        if (__builtin_popcount(uint32_t(e) & op.rmask) & 1) op.lwant |= 1u << e;
    } else if (lad_of[pi] || (it.kind == QB_K_U && it.g.ctl_mask == 0 && pi + 1 < pend.size() &&
                       pend[pi + 1].it->kind == QB_K_LADDER && pend[pi + 1].it->pivot == it.g.target)) {
      // h(q) + the cu1 ladder hanging off q (circuit.py:323-326): one op, y' = (c x + d y) * phase
      const Item &lad = lad_of[pi] ? *lad_of[pi] : *pend[pi + 1].it;
      op.kind = QB_K_ULADDER;
      int lp = tm.lpos[it.g.target];
      for (int k = 0; k < nr; ++k)
        if (r.rbit[k] == lp) op.tpos = k;
      memcpy(op.m, it.g.m, sizeof op.m);
      build_ladder_tables(tm, r, lad, (*nladders)++, nbits, pp, &op);
      if (!lad_of[pi]) ++pi;
    } else if (it.kind == QB_K_LADDER) {
      op.kind = QB_K_LADDER;
      split_pred(tm, r.rbit, nr, uint64_t(1) << it.pivot, uint64_t(1) << it.pivot, &op);
      build_ladder_tables(tm, r, it, (*nladders)++, nbits, pp, &op);
    } else if (it.kind == QB_K_PHASE) {
      uint64_t bits = it.g.ctl_mask | (uint64_t(1) << it.g.target);
      split_pred(tm, r.rbit, nr, bits, bits, &op);
      op.m[0] = it.g.m[6];
      op.m[1] = it.g.m[7];
    } else if (it.kind == QB_K_DIAG) {
      // lowered to two phase ops: d0 where the target bit is clear, d1 where it is set
      uint64_t bits = it.g.ctl_mask | (uint64_t(1) << it.g.target);
      uint64_t want = it.g.ctl_mask | (po.variant ? (uint64_t(1) << it.g.target) : 0);
      split_pred(tm, r.rbit, nr, bits, want, &op);
      op.kind = QB_K_PHASE;
      op.m[0] = it.g.m[po.variant ? 6 : 0];
      op.m[1] = it.g.m[po.variant ? 7 : 1];
    } else {  // U / PERM / SWAP
      split_pred(tm, r.rbit, nr, it.g.ctl_mask, it.g.ctl_mask, &op);
      int lp = tm.lpos[it.g.target];
      for (int k = 0; k < nr; ++k)
        if (r.rbit[k] == lp) op.tpos = k;
      memcpy(op.m, it.g.m, sizeof op.m);
    }
    if ((op.kind == QB_K_U || op.kind == QB_K_ULADDER || op.kind == QB_K_PERM) && op.m[1] == 0.0 &&
        op.m[3] == 0.0 && op.m[5] == 0.0 && op.m[7] == 0.0)
      op.mflags |= QB_MF_REAL;
    if ((op.mflags & QB_MF_REAL) && op.m[0] == op.m[2] && op.m[0] == op.m[4] && op.m[0] == -op.m[6])
      op.mflags |= QB_MF_HADAMARD;
    if (op.kind == QB_K_U && !(op.mflags & QB_MF_REAL) && op.m[1] == 0.0 && op.m[5] == 0.0 && op.m[2] == 0.0 &&
        op.m[6] == 0.0)
      op.mflags |= QB_MF_COLIMAG;  // a, c real; b, d imaginary
    // the kernel decodes one word per op: kind | tpos << 8 | mflags << 16 | dense opcode << 24
    int opc = 0;
    const bool real = op.mflags & QB_MF_REAL;
    const bool unmasked = op.lmask == 0 && op.rmask == 0;
    switch (op.kind & 0xff) {
      case QB_K_ULADDER:
        opc = QB_OPC_ULADDER + op.tpos + ((op.mflags & QB_MF_HADAMARD) ? 6 : (real ? 3 : 0));
        break;
      case QB_K_U:
        opc = !unmasked ? QB_OPC_U_MASKED + op.tpos
                        : ((op.mflags & QB_MF_COLIMAG) ? QB_OPC_U_CI + op.tpos : QB_OPC_U_ALL + op.tpos + (real ? 3 : 0));
        break;
      case QB_K_PERM: opc = QB_OPC_PERM + op.tpos; break;
      case QB_K_SWAP: opc = QB_OPC_SWAP + op.tpos; break;
      case QB_K_PARSWAP: opc = QB_OPC_PARSWAP + op.tpos; break;
      case QB_K_PHASE: opc = QB_OPC_PHASE; break;
      default: opc = QB_OPC_LADDER; break;
    }
    if ((op.kind & 0xff) == QB_K_SWAP || (op.kind & 0xff) == QB_K_PHASE ||
        ((op.kind & 0xff) == QB_K_U && opc >= QB_OPC_U_MASKED && opc < QB_OPC_U_MASKED + 3)) {
      // which of the 8 registers the round-level predicate selects, precomputed for the lean interpreter
      op.flags = 0;
      for (int e = 0; e < 8; ++e)
        if ((uint32_t(e) & op.rmask) == op.rwant) op.flags |= 1 << e;
    }
    op.kind = (op.kind & 0xff) | ((op.tpos & 0xff) << 8) | ((op.mflags & 0xff) << 16) | (opc << 24);
    pp->ops.push_back(op);
  }
  r.op_end = int32_t(pp->ops.size());
  // Round program: three Hadamard+ladder ops on round positions 0, 1, 2 (what every round of a QFT
  // looks like) run through a fully unrolled path in the kernel.
  r.prog = QB_PROG_GENERIC;
  if (nr == QB_ROUND_BITS && r.op_end - r.op_begin == 3) {
    bool hl3 = true;
    for (int k = 0; k < 3; ++k) {
      const QbOp &o = pp->ops[size_t(r.op_begin + k)];
      if (int(uint32_t(o.kind) >> 24) != QB_OPC_ULADDER + 6 + k || o.gmask != 0) hl3 = false;
    }
    if (hl3) {
      const QbOp &o1 = pp->ops[size_t(r.op_begin + 1)], &o2 = pp->ops[size_t(r.op_begin + 2)];
      auto isone = [](const double *F, int e) { return F[2 * e] == 1.0 && F[2 * e + 1] == 0.0; };
      const bool upper = isone(o1.F, 3) && o1.F[14] == o1.F[12] && o1.F[15] == o1.F[13] && isone(o2.F, 5) &&
                         isone(o2.F, 6) && isone(o2.F, 7);
      r.prog = upper ? QB_PROG_HL3U : QB_PROG_HL3;
    }
  }
  // Round program UX: uncontrolled U's, parity swaps (every round of larose_benchmark), and -- unless
  // QCC_B200_UX_NARROW is set -- controlled swaps (cx / ccx), controlled phases (z s t u1 cz cu1 ...) and
  // controlled 2x2s (cv, ch, crx, ... and the sqrt(X) of the Toffoli expansion) and phase ladders:
  // everything the lean interpreter of fused.cu handles without the generic one's registers.
  if (r.prog == QB_PROG_GENERIC && nr == QB_ROUND_BITS && r.op_end > r.op_begin) {
    static const bool narrow = getenv("QCC_B200_UX_NARROW") != nullptr;
    bool ux = true;
    for (int k = r.op_begin; k < r.op_end; ++k) {
      const QbOp &o = pp->ops[size_t(k)];
      const int opc = int(uint32_t(o.kind) >> 24);
      const bool u_all = ((opc >= QB_OPC_U_ALL && opc < QB_OPC_U_ALL + 6) || (opc >= QB_OPC_U_CI && opc < QB_OPC_U_CI + 3)) &&
                         (o.gmask == 0 || !narrow);
      const bool psw = opc >= QB_OPC_PARSWAP && opc < QB_OPC_PARSWAP + 3;
      const bool cswap_or_phase = !narrow && ((opc >= QB_OPC_SWAP && opc < QB_OPC_SWAP + 3) || opc == QB_OPC_PHASE ||
                                              (opc >= QB_OPC_U_MASKED && opc < QB_OPC_U_MASKED + 3));
      const bool ladder = !narrow && (opc < QB_OPC_U_ALL || opc == QB_OPC_LADDER);  // ULADDER family / LADDER
      if (!u_all && !psw && !cswap_or_phase && !ladder) ux = false;
    }
    if (ux) r.prog = QB_PROG_UX;
  }
  pp->rounds.push_back(r);
}

void emit_pass(int nbits, int K, const std::vector<const Item *> &items, const std::vector<int> &targets_hi,
               Plan *out, bool must_fuse = false) {
  if (items.empty()) return;
  // Cheap cases first: one real gate, or a handful of plain PHASE gates whose own sweeps
  // move less data than a full 32 B/amplitude pass.
  int64_t real = 0;
  bool all_phase = true;
  double bytes = 0;
  for (const Item *it : items) {
    if (it->kind == QB_K_NOP) continue;
    real += 1;
    bytes += it->bytes_per_amp;
    if (it->kind != QB_K_PHASE) all_phase = false;
  }
  if (real == 0 || (!must_fuse && ((real == 1 && items.size() == 1 && items[0]->kind != QB_K_LADDER) ||
                                   (all_phase && bytes <= 24.0)))) {
    for (const Item *it : items) {
      PlannedPass pp;
      pp.single_gate = it->first_gate;
      pp.single = it->g;
      pp.ngates = it->ngates;
      pp.bytes_algorithmic_per_amp = it->bytes_per_amp;
      out->passes.push_back(std::move(pp));
    }
    return;
  }
  PlannedPass pp;
  TileMap tm = make_tile(nbits, K, targets_hi);
  pp.desc.K = tm.K;
  pp.desc.tile_mask = tm.mask;
  for (int k = 0; k < tm.K; ++k) pp.desc.tile_bits[k] = tm.bits[k];
  int nlad = 0;
  std::vector<const Item *> live;
  for (const Item *it : items) {
    pp.ngates += it->ngates;
    pp.bytes_algorithmic_per_amp += it->bytes_per_amp;
    if (it->kind != QB_K_NOP) live.push_back(it);
  }
  std::vector<RoundPlan> rplans = inorder_rounds(tm, live);
  static const bool no_sched = getenv("QCC_B200_NO_SCHED") != nullptr;
  if (!no_sched && tm.K >= QB_ROUND_BITS) {
    std::vector<RoundPlan> sched = schedule_rounds(tm, live);
    if (sched.size() <= rplans.size()) rplans.swap(sched);
  }
  assign_group_maps(tm, &rplans, &pp.desc);
  for (RoundPlan &rp : rplans) close_round(tm, rp, &nlad, nbits, &pp);
  ladder_field_widths(nbits, tm.K, pp.desc.lad_w);
  pp.desc.nrounds = int32_t(pp.rounds.size());
  pp.desc.nops = int32_t(pp.ops.size());
  {
    static const bool no_direct = getenv("QCC_B200_NO_DIRECT") != nullptr;
    auto direct_ok = [&](const QbRound &R) {
      if (!pp.desc.warp_io || no_direct || R.prog == QB_PROG_GENERIC) return false;
      int low = 0;
      for (int k = 0; k < 3; ++k) {
        if (R.rbit[k] < 3) return false;
        if (R.qmap[k] < 3) low |= 1 << R.qmap[k];
      }
      return low == 7;
    };
    pp.desc.st_direct = !pp.rounds.empty() && direct_ok(pp.rounds.back()) ? 1 : 0;
  }
  pp.desc.ntable = int32_t(pp.tables.size());
  pp.desc.ngroups_log2 = tm.K - 3;
  // runs of consecutive non-tile bits: the kernel scatters the tile number over them
  {
    int nseg = 0;
    bool overflow = false;
    for (int b = 0; b < nbits;) {
      if (tm.mask >> b & 1) {
        ++b;
        continue;
      }
      int e = b;
      while (e < nbits && !(tm.mask >> e & 1)) ++e;
      if (nseg < QB_MAX_SEGS) {
        pp.desc.seg_pos[nseg] = b;
        pp.desc.seg_len[nseg] = e - b;
        ++nseg;
      } else {
        overflow = true;
      }
      b = e;
    }
    pp.desc.nseg = overflow ? -1 : nseg;
  }
  pp.desc.nout_total = int32_t(pp.outph.size());
  pp.noutbits = int(pp.outbits.size());
  out->passes.push_back(std::move(pp));
}

size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

}  // namespace

int classify_matrix(const double m[8]) {
  auto z = [&](int k) { return m[2 * k] == 0.0 && m[2 * k + 1] == 0.0; };
  auto one = [&](int k) { return m[2 * k] == 1.0 && m[2 * k + 1] == 0.0; };
  if (z(1) && z(2)) {
    if (one(0) && one(3)) return QB_K_NOP;
    if (one(0)) return QB_K_PHASE;
    return QB_K_DIAG;
  }
  if (z(0) && z(3)) return QB_K_PERM;
  return QB_K_U;
}

double gate_bytes_per_amp(const QbGate &g) {
  if (g.kind == QB_K_NOP) return 0.0;
  int nb = __builtin_popcountll(g.ctl_mask);
  if (g.kind == QB_K_PHASE || g.kind == QB_K_DIAG) nb += 1;  // diagonal = half
  return 32.0 / double(uint64_t(1) << nb);
}

// ---- peephole: Sleator-Weinfurter recognition -------------------------------------------------------
// circuit.py:227-246 (ccu) spells every doubly-controlled gate -- each Toffoli of multi_control's ladder
// (circuit.py:341-392), ccu1, cswap -- as five gates:
//     cu(a, t, V)  cx(a, b)  cu(b, t, V^dagger)  cx(a, b)  cu(b, t, V)        with V^2 = U
// which is exactly  U on t where a AND b are set.  Found in the queued stream (the five gates adjacent, the
// matrices matching to 1e-13), the run is replaced by ONE doubly-controlled gate U = V V plus four no-ops that
// keep the gate count: a quarter of the amplitudes touched once instead of three half-vector butterflies and
// two swaps, and b stops being a mixing target (fewer tile bits per pass, fewer exchanges on a sharded state).
// Entries of U within 1e-14 of 0 / +-1 are snapped so that X^(1/2) squared is again the pure swap.
int64_t fuse_ccu_runs(QbGate *g, int64_t n) {
  auto one_ctl = [](const QbGate &x) { return x.ctl_mask != 0 && (x.ctl_mask & (x.ctl_mask - 1)) == 0; };
  auto is_cx = [&](const QbGate &x) {
    return one_ctl(x) && (x.kind == QB_K_PERM || x.kind == QB_K_SWAP) && x.m[0] == 0.0 && x.m[1] == 0.0 && x.m[2] == 1.0 &&
           x.m[3] == 0.0 && x.m[4] == 1.0 && x.m[5] == 0.0 && x.m[6] == 0.0 && x.m[7] == 0.0;
  };
  const double tol = 1e-13;
  int64_t fused = 0;
  for (int64_t i = 0; i + 4 < n; ++i) {
    const QbGate &g1 = g[i], &g2 = g[i + 1], &g3 = g[i + 2], &g4 = g[i + 3], &g5 = g[i + 4];
    if (g1.kind == QB_K_NOP || !one_ctl(g1) || !is_cx(g2) || !is_cx(g4)) continue;
    const uint64_t a = g1.ctl_mask, bmask = uint64_t(1) << g2.target;
    const int t = g1.target;
    if (g2.ctl_mask != a || g4.ctl_mask != a || g4.target != g2.target || g2.target == t || (a >> t & 1)) continue;
    if (g3.ctl_mask != bmask || g5.ctl_mask != bmask || g3.target != t || g5.target != t) continue;
    if (g3.kind == QB_K_NOP || g5.kind == QB_K_NOP) continue;
    bool ok = true;
    // g5 == g1 and g3 == g1^dagger
    const int tr[4] = {0, 2, 1, 3};
    for (int k = 0; k < 4 && ok; ++k) {
      if (fabs(g5.m[2 * k] - g1.m[2 * k]) > tol || fabs(g5.m[2 * k + 1] - g1.m[2 * k + 1]) > tol) ok = false;
      if (fabs(g3.m[2 * k] - g1.m[2 * tr[k]]) > tol || fabs(g3.m[2 * k + 1] + g1.m[2 * tr[k] + 1]) > tol) ok = false;
    }
    if (!ok) continue;
    // V must be unitary for g3 to undo it where only one control is set
    Cplx v[4], vd[4], u[4], id[4];
    for (int k = 0; k < 4; ++k) {
      v[k] = Cplx{g1.m[2 * k], g1.m[2 * k + 1]};
      vd[k] = Cplx{g3.m[2 * k], g3.m[2 * k + 1]};
    }
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 2; ++c) {
        Cplx p = cmulh(v[r * 2], v[c]), q = cmulh(v[r * 2 + 1], v[2 + c]);
        u[r * 2 + c] = Cplx{p.x + q.x, p.y + q.y};
        Cplx p2 = cmulh(vd[r * 2], v[c]), q2 = cmulh(vd[r * 2 + 1], v[2 + c]);
        id[r * 2 + c] = Cplx{p2.x + q2.x, p2.y + q2.y};
      }
    if (fabs(id[0].x - 1) > tol || fabs(id[0].y) > tol || fabs(id[3].x - 1) > tol || fabs(id[3].y) > tol ||
        fabs(id[1].x) > tol || fabs(id[1].y) > tol || fabs(id[2].x) > tol || fabs(id[2].y) > tol)
      continue;
    QbGate f{};
    f.ctl_mask = a | bmask;
    f.target = t;
    auto snap = [](double x) {
      if (fabs(x) < 1e-14) return 0.0;
      if (fabs(x - 1.0) < 1e-14) return 1.0;
      if (fabs(x + 1.0) < 1e-14) return -1.0;
      return x;
    };
    for (int k = 0; k < 4; ++k) {
      f.m[2 * k] = snap(u[k].x);
      f.m[2 * k + 1] = snap(u[k].y);
    }
    f.kind = classify_matrix(f.m);
    g[i] = f;
    for (int k = 1; k < 5; ++k) {
      QbGate nop{};
      nop.ctl_mask = 0;
      nop.target = t;
      nop.kind = QB_K_NOP;
      nop.m[0] = 1.0;
      nop.m[6] = 1.0;
      g[i + k] = nop;
    }
    fused += 1;
    i += 4;
  }
  return fused;
}

void plan_gates(int nbits, const QbGate *gates, int64_t ngates, int tile_bits, Plan *out, bool fuse_last) {
  out->passes.clear();
  int K = std::max(4, std::min(tile_bits, QB_MAX_TILE_BITS));
  K = std::min(K, nbits);
  const int cap = K - std::min(QB_TILE_LOW, K);
  static const bool no_ladder = getenv("QCC_B200_NO_LADDER") != nullptr;
  static const bool no_premul = getenv("QCC_B200_NO_PREMUL") != nullptr;
  std::vector<Merged> merged;
  premultiply(gates, ngates, !no_premul, &merged);
  std::vector<Item> items;
  build_items(merged, !no_ladder, &items);
  std::vector<const Item *> cur;
  std::vector<int> targets;
  int nlad_pass = 0;  // ladders this pass holds (their tables are staged in shared memory)
  int nops = 0;       // ops this pass will hold (staged in shared memory by the kernel)
  int nrounds_ub = 1; // upper bound on its rounds: a new round at most every 3 new targets
  std::vector<int> round_targets;
  for (const Item &it : items) {
    bool new_target = needs_target(it.kind) && it.g.target >= QB_TILE_LOW &&
                      std::find(targets.begin(), targets.end(), it.g.target) == targets.end();
    int cost = it.kind == QB_K_NOP ? 0 : (it.kind == QB_K_DIAG ? 2 : 1);
    bool new_round = false;
    if (needs_target(it.kind) &&
        std::find(round_targets.begin(), round_targets.end(), it.g.target) == round_targets.end()) {
      if (int(round_targets.size()) == QB_ROUND_BITS) new_round = true;
    }
    bool lad_full = it.kind == QB_K_LADDER && nlad_pass == QB_MAX_PASS_LADDERS;
    if ((new_target && int(targets.size()) == cap) || nops + cost > QB_MAX_PASS_OPS || lad_full ||
        (new_round && nrounds_ub == QB_MAX_PASS_ROUNDS)) {
      emit_pass(nbits, K, cur, targets, out);
      cur.clear();
      targets.clear();
      round_targets.clear();
      nops = 0;
      nlad_pass = 0;
      nrounds_ub = 1;
      new_round = false;
      new_target = needs_target(it.kind) && it.g.target >= QB_TILE_LOW;
    }
    if (new_round) {
      round_targets.clear();
      nrounds_ub += 1;
    }
    if (needs_target(it.kind) &&
        std::find(round_targets.begin(), round_targets.end(), it.g.target) == round_targets.end())
      round_targets.push_back(it.g.target);
    if (new_target) targets.push_back(it.g.target);
    nops += cost;
    if (it.kind == QB_K_LADDER) nlad_pass += 1;
    cur.push_back(&it);
  }
  emit_pass(nbits, K, cur, targets, out, fuse_last);
}

size_t Plan::blob_bytes() {
  size_t off = 0;
  for (PlannedPass &p : passes) {
    if (p.single_gate >= 0) continue;
    p.ops_off = off;
    off = align16(off + p.ops.size() * sizeof(QbOp));
    p.rounds_off = off;
    off = align16(off + p.rounds.size() * sizeof(QbRound));
    p.tables_off = off;
    off = align16(off + p.tables.size() * sizeof(Cplx));
    p.outph_off = off;
    off = align16(off + p.outph.size() * sizeof(Cplx));
    p.jbtab_off = off;
    off = align16(off + p.jbtab.size() * sizeof(uint32_t));
  }
  return std::max<size_t>(off, 16);
}

void Plan::serialize(char *dst) const {
  for (const PlannedPass &p : passes) {
    if (p.single_gate >= 0) continue;
    memcpy(dst + p.ops_off, p.ops.data(), p.ops.size() * sizeof(QbOp));
    memcpy(dst + p.rounds_off, p.rounds.data(), p.rounds.size() * sizeof(QbRound));
    if (!p.tables.empty()) memcpy(dst + p.tables_off, p.tables.data(), p.tables.size() * sizeof(Cplx));
    if (!p.outph.empty()) memcpy(dst + p.outph_off, p.outph.data(), p.outph.size() * sizeof(Cplx));
    memcpy(dst + p.jbtab_off, p.jbtab.data(), p.jbtab.size() * sizeof(uint32_t));
  }
}

std::string Plan::to_json() const {
  std::string s = "{\"passes\":[";
  char buf[512];
  bool firstp = true;
  for (const PlannedPass &p : passes) {
    if (!firstp) s += ",";
    firstp = false;
    if (p.single_gate >= 0) {
      snprintf(buf, sizeof buf, "{\"single_gate\":%lld,\"ngates\":%lld,\"ctl_mask\":%llu,\"target\":%d,\"m\":[",
               (long long)p.single_gate, (long long)p.ngates, (unsigned long long)p.single.ctl_mask,
               p.single.target);
      s += buf;
      for (int k = 0; k < 8; ++k) {
        snprintf(buf, sizeof buf, "%s%.17g", k ? "," : "", p.single.m[k]);
        s += buf;
      }
      s += "]}";
      continue;
    }
    snprintf(buf, sizeof buf, "{\"single_gate\":-1,\"ngates\":%lld,\"K\":%d,\"warp_io\":%d,\"st_direct\":%d,\"lad_w\":[%d,%d,%d],\"ld_map\":[",
             (long long)p.ngates, p.desc.K, p.desc.warp_io, p.desc.st_direct, p.desc.lad_w[0], p.desc.lad_w[1],
             p.desc.lad_w[2]);
    s += buf;
    for (int k = 0; k < p.desc.K; ++k) {
      snprintf(buf, sizeof buf, "%s%d", k ? "," : "", p.desc.ld_map[k]);
      s += buf;
    }
    s += "],\"st_map\":[";
    for (int k = 0; k < p.desc.K; ++k) {
      snprintf(buf, sizeof buf, "%s%d", k ? "," : "", p.desc.st_map[k]);
      s += buf;
    }
    s += "],\"tile_bits\":[";
    for (int k = 0; k < p.desc.K; ++k) {
      snprintf(buf, sizeof buf, "%s%d", k ? "," : "", p.desc.tile_bits[k]);
      s += buf;
    }
    s += "],\"rounds\":[";
    for (size_t r = 0; r < p.rounds.size(); ++r) {
      const QbRound &R = p.rounds[r];
      snprintf(buf, sizeof buf, "%s{\"nbits\":%d,\"rbit\":[%d,%d,%d],\"op_begin\":%d,\"op_end\":%d,\"prog\":%d,\"nobar\":%d,\"qmap\":[",
               r ? "," : "", R.nbits, R.rbit[0], R.rbit[1], R.rbit[2], R.op_begin, R.op_end, R.prog, R.nobar);
      s += buf;
      for (int k = 0; k < p.desc.K - R.nbits; ++k) {
        snprintf(buf, sizeof buf, "%s%d", k ? "," : "", R.qmap[k]);
        s += buf;
      }
      s += "]}";
    }
    s += "],\"ops\":[";
    for (size_t o = 0; o < p.ops.size(); ++o) {
      const QbOp &O = p.ops[o];
      snprintf(buf, sizeof buf,
               "%s{\"kind\":%d,\"tpos\":%d,\"lmask\":%u,\"lwant\":%u,\"rmask\":%u,\"rwant\":%u,"
               "\"gmask\":%llu,\"gwant\":%llu,\"table_off\":%d,\"nout\":%d,\"out_off\":%d,\"outph_off\":%d,"
               "\"flags\":%d,\"mflags\":%d,\"m\":[",
               o ? "," : "", O.kind, O.tpos, O.lmask, O.lwant, O.rmask, O.rwant, (unsigned long long)O.gmask,
               (unsigned long long)O.gwant, O.table_off, O.nout, O.out_off, O.outph_off, O.flags, O.mflags);
      s += buf;
      for (int k = 0; k < 8; ++k) {
        snprintf(buf, sizeof buf, "%s%.17g", k ? "," : "", O.m[k]);
        s += buf;
      }
      s += "],\"F\":[";
      for (int k = 0; k < 16; ++k) {
        snprintf(buf, sizeof buf, "%s%.17g", k ? "," : "", O.F[k]);
        s += buf;
      }
      s += "]}";
    }
    s += "],\"tables\":[";
    for (size_t t = 0; t < p.tables.size(); ++t) {
      snprintf(buf, sizeof buf, "%s[%.17g,%.17g]", t ? "," : "", p.tables[t].x, p.tables[t].y);
      s += buf;
    }
    s += "],\"outph\":[";
    for (size_t t = 0; t < p.outph.size(); ++t) {
      snprintf(buf, sizeof buf, "%s[%.17g,%.17g]", t ? "," : "", p.outph[t].x, p.outph[t].y);
      s += buf;
    }
    s += "],\"jbtab\":[";
    for (size_t t = 0; t < p.jbtab.size(); ++t) {
      snprintf(buf, sizeof buf, "%s%u", t ? "," : "", p.jbtab[t]);
      s += buf;
    }
    s += "],\"outbits\":[";
    for (size_t t = 0; t < p.outbits.size(); ++t) {
      snprintf(buf, sizeof buf, "%s%d", t ? "," : "", p.outbits[t]);
      s += buf;
    }
    s += "]}";
  }
  s += "]}";
  return s;
}

}  // namespace qb
