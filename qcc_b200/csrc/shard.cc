// shard.cc -- see shard.h.
#include "shard.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace qb {

namespace {

const double kX[8] = {0, 0, 1, 0, 1, 0, 0, 0};

std::vector<int> inverse_of(const std::vector<int> &perm) {
  std::vector<int> inv(perm.size());
  for (size_t b = 0; b < perm.size(); ++b) inv[size_t(perm[b])] = int(b);
  return inv;
}

void close_step(ShardStep *cur, std::vector<ShardStep> *steps) {
  if (!cur->gates.empty() || cur->retired) steps->push_back(std::move(*cur));
  *cur = ShardStep();
}

bool nondiagonal(int kind) { return kind == QB_K_U || kind == QB_K_PERM || kind == QB_K_SWAP; }

bool is_plain_x(const QbGate &g) {
  return g.ctl_mask == 0 && (g.kind == QB_K_PERM || g.kind == QB_K_SWAP) && g.m[2] == 1.0 && g.m[3] == 0.0 &&
         g.m[4] == 1.0 && g.m[5] == 0.0;
}

bool relabel_enabled() {
  static const bool no_relabel = getenv("QCC_B200_NO_RELABEL") != nullptr;
  return !no_relabel;
}

// does gate g need its target qubit to be LOCAL?  (a plain x on a sharded qubit is a rank relabel)
bool needs_local(const QbGate &g) { return nondiagonal(g.kind) && !(relabel_enabled() && is_plain_x(g)); }

// first gate index >= from that needs logical qubit lq local; ngates if none
int64_t next_need(const QbGate *gates, int64_t ngates, int64_t from, int lq) {
  for (int64_t j = from; j < ngates; ++j)
    if (gates[j].target == lq && needs_local(gates[j])) return j;
  return ngates;
}

// value of the LOGICAL qubit that lives at global physical position pb, on this rank
int global_bit(const ShardLayout &L, int pb) { return int(((uint32_t(L.rank) ^ L.flip) >> (pb - L.nl)) & 1u); }

// One exchange event: every (global position, victim) pair is exchanged -- the victim's qubit goes to the rank
// bit, the rank bit's qubit to the landing bit (the victim bit itself unless L->land picks a higher one, whose
// occupant then moves down into the victim bit).  A relabelled (flipped) rank bit arrives as the negation of its
// qubit: one local x on the landing bit puts it right, and the new occupant of the rank bit is plain.
void exchange_event(ShardLayout *L, const std::vector<std::pair<int, int>> &pairs, ShardStep *cur,
                    std::vector<ShardStep> *steps) {
  close_step(cur, steps);
  ShardStep ex;
  ex.kind = 1;
  // landing bits: the highest local bits that are not victims of this event, highest first, for the pairs in
  // order; a pair whose victim is already above what is left keeps the plain swap
  std::vector<int> lands(pairs.size());
  {
    std::vector<char> taken(static_cast<size_t>(L->nl), 0);
    for (const auto &pr : pairs) taken[size_t(pr.second)] = 1;
    int h = L->nl - 1;
    for (size_t k = 0; k < pairs.size(); ++k) {
      lands[k] = pairs[k].second;
      if (!L->land) continue;
      while (h >= 0 && taken[size_t(h)]) --h;
      if (h > pairs[k].second) {
        lands[k] = h;
        taken[size_t(h)] = 1;
      }
    }
  }
  for (size_t k = 0; k < pairs.size(); ++k) {
    const int global_pos = pairs[k].first, victim = pairs[k].second, land = lands[k];
    ex.rank_bits.push_back(global_pos - L->nl);
    ex.victims.push_back(victim);
    ex.lands.push_back(land);
    std::vector<int> inv = inverse_of(L->perm);
    const int q_in = inv[size_t(global_pos)], q_out = inv[size_t(victim)], q_down = inv[size_t(land)];
    L->perm[size_t(q_out)] = global_pos;
    L->perm[size_t(q_in)] = land;
    if (land != victim) L->perm[size_t(q_down)] = victim;
    const uint32_t fb = 1u << (global_pos - L->nl);
    if (L->flip & fb) {
      L->flip &= ~fb;
      QbGate x{};
      x.ctl_mask = 0;
      x.target = land;
      x.kind = QB_K_PERM;
      memcpy(x.m, kX, sizeof kX);
      cur->gates.push_back(x);
    }
  }
  steps->push_back(ex);
}

void exchange_bits(ShardLayout *L, int global_pos, int victim, ShardStep *cur, std::vector<ShardStep> *steps) {
  exchange_event(L, {{global_pos, victim}}, cur, steps);
}

// Where to put the exchange that gate i needs: the latest pass boundary of the local stream in
// [seg_start, i] from which on the victim's qubit is not needed local any more (it becomes sharded
// at that point).  Pass boundaries are estimated the way the fusion planner cuts: a new pass when one
// more distinct target bit above QB_TILE_LOW would not fit.  Depends on the stream, the permutation
// and the victim only, so every rank finds the same point.
int64_t hoist_point(const ShardLayout &L, const QbGate *gates, int64_t seg_start, int64_t i, int victim) {
  std::vector<int> inv = inverse_of(L.perm);
  const int lv = inv[size_t(victim)];
  int64_t j0 = seg_start;
  for (int64_t k = i - 1; k >= seg_start; --k)
    if (gates[k].target == lv && needs_local(gates[k])) {
      j0 = k + 1;
      break;
    }
  int64_t best = i;
  bool found = false;
  if (j0 == seg_start && L.hoist > 1) {   // the very start of the segment is a boundary too (previous event / flush)
    best = seg_start;
    found = true;
  }
  std::vector<int> targets;
  for (int64_t k = seg_start; k < i; ++k) {
    if (!nondiagonal(gates[k].kind)) continue;
    const int pt = L.perm[size_t(gates[k].target)];
    if (pt >= L.nl || pt < QB_TILE_LOW) continue;   // relabelled x on a sharded bit / always-resident low bits
    if (std::find(targets.begin(), targets.end(), pt) != targets.end()) continue;
    if (int(targets.size()) == L.pass_targets) {
      targets.clear();
      if (k >= j0) {
        best = k;
        found = true;
      }
    }
    targets.push_back(pt);
  }
  return found ? best : i;
}

// Victim for the qubit gate i needs: among the local bits of the window, prefer one that lets the
// exchange sit on a pass boundary (hoisting; its qubit is not needed again before that boundary),
// then Belady -- the qubit whose next use as a mixing target is farthest -- then the higher bit.
// A hoistable victim that is needed again almost at once is not preferred (it would come straight back).
void choose_victim(const ShardLayout &L, const QbGate *gates, int64_t ngates, int64_t i, int64_t seg_start,
                   int *victim, int64_t *point) {
  std::vector<int> inv = inverse_of(L.perm);
  int best = L.nl - 1;
  int64_t best_dist = -1, best_point = i;
  bool best_hoists = false;
  const int64_t soon = 4 * int64_t(L.pass_targets);
  for (int v = L.nl - 1; v >= std::max(0, L.nl - L.window); --v) {
    const int lv = inv[size_t(v)];
    const int64_t dist = next_need(gates, ngates, i + 1, lv) - i;
    const int64_t hp = L.hoist ? hoist_point(L, gates, seg_start, i, v) : i;
    const bool hoists = hp < i && (dist > soon || dist >= ngates - i);
    bool better;
    if (best_dist < 0) better = true;
    else if (hoists != best_hoists) better = hoists;
    else better = dist > best_dist;
    if (better) {
      best = v;
      best_dist = dist;
      best_hoists = hoists;
      best_point = hoists ? hp : i;
    }
  }
  *victim = best;
  *point = best_point;
}

// Prefetch: the event at point j (needed by gate i) also swaps in every other sharded qubit that is
// needed before the best remaining victim is, soonest first.
void add_prefetch_pairs(const ShardLayout &L, const QbGate *gates, int64_t ngates, int64_t j,
                        std::vector<std::pair<int, int>> *pairs) {
  std::vector<int> inv = inverse_of(L.perm);
  struct Cand {
    int pos;
    int64_t need;
  };
  std::vector<Cand> globals, locals;
  auto used = [&](int pos) {
    for (const auto &pr : *pairs)
      if (pr.first == pos || pr.second == pos) return true;
    return false;
  };
  for (int G = L.nl; G < L.n; ++G) {
    if (used(G)) continue;
    const int64_t u = next_need(gates, ngates, j, inv[size_t(G)]);
    if (u < ngates) globals.push_back(Cand{G, u});
  }
  if (globals.empty()) return;
  for (int v = L.nl - 1; v >= std::max(0, L.nl - L.window); --v)
    if (!used(v)) locals.push_back(Cand{v, next_need(gates, ngates, j, inv[size_t(v)])});
  std::stable_sort(globals.begin(), globals.end(), [](const Cand &a, const Cand &b) { return a.need < b.need; });
  std::stable_sort(locals.begin(), locals.end(), [](const Cand &a, const Cand &b) { return a.need > b.need; });
  size_t li = 0;
  for (const Cand &g : globals) {
    if (li >= locals.size() || locals[li].need <= g.need) break;
    pairs->push_back({g.pos, locals[li].pos});
    ++li;
  }
}

}  // namespace

void lower_for_rank(ShardLayout *L, const QbGate *gates, int64_t ngates, std::vector<ShardStep> *steps) {
  ShardStep cur;
  const int nl = L->nl;
  // state of the open step before each gate since seg_start (for rolling back to a hoisted exchange point)
  struct Mark {
    size_t ngates;
    int64_t retired;
    uint32_t flip;
  };
  std::vector<Mark> marks;
  int64_t seg_start = 0;
  for (int64_t i = 0; i < ngates; ++i) {
    const QbGate &g = gates[i];
    marks.push_back(Mark{cur.gates.size(), cur.retired, L->flip});
    if (g.kind == QB_K_NOP) {
      cur.retired += 1;
      continue;
    }
    if (nondiagonal(g.kind) && L->perm[size_t(g.target)] >= nl) {
      if (is_plain_x(g) && relabel_enabled()) {   // x on a sharded qubit: relabel the rank bit, move nothing
        L->flip ^= 1u << (L->perm[size_t(g.target)] - nl);
        cur.retired += 1;
        continue;
      }
      int victim = 0;
      int64_t j = i;
      choose_victim(*L, gates, ngates, i, seg_start, &victim, &j);
      std::vector<std::pair<int, int>> pairs{{L->perm[size_t(g.target)], victim}};
      if (L->prefetch) add_prefetch_pairs(*L, gates, ngates, j, &pairs);
      if (j < i) {   // undo what the open step holds of gates j .. i-1; they are lowered again below
        const Mark &mk = marks[size_t(j - seg_start)];
        cur.gates.resize(mk.ngates);
        cur.retired = mk.retired;
        L->flip = mk.flip;
      }
      exchange_event(L, pairs, &cur, steps);
      seg_start = j;
      marks.clear();
      i = j - 1;     // resume at gate j under the new layout (gate i itself now finds its target local)
      continue;
    }
    const int pt = L->perm[size_t(g.target)];
    uint64_t lm = 0;
    bool active = true;
    for (int b = 0; b < L->n; ++b) {
      if (!(g.ctl_mask >> b & 1)) continue;
      const int pb = L->perm[size_t(b)];
      if (pb < nl) lm |= uint64_t(1) << pb;
      else if (!global_bit(*L, pb)) active = false;
    }
    cur.retired += 1;
    if (!active) continue;  // a global control bit is 0 on this rank: the gate touches nothing here
    QbGate out{};
    if (pt < nl) {
      out = g;
      out.ctl_mask = lm;
      out.target = pt;
      cur.gates.push_back(out);
      continue;
    }
    // diagonal gate whose target bit is a rank bit: a phase on the remaining local bits
    const int tb = global_bit(*L, pt);
    double pr, pi;
    if (g.kind == QB_K_PHASE) {
      if (!tb) continue;
      pr = g.m[6];
      pi = g.m[7];
    } else {  // DIAG
      pr = tb ? g.m[6] : g.m[0];
      pi = tb ? g.m[7] : g.m[1];
      if (pr == 1.0 && pi == 0.0) continue;
    }
    if (lm == 0) {  // scalar on the whole shard
      out.ctl_mask = 0;
      out.target = 0;
      out.kind = QB_K_DIAG;
      out.m[0] = pr; out.m[1] = pi; out.m[6] = pr; out.m[7] = pi;
    } else {
      const int t2 = __builtin_ctzll(lm);
      out.ctl_mask = lm & ~(uint64_t(1) << t2);
      out.target = t2;
      out.kind = QB_K_PHASE;
      out.m[0] = 1.0; out.m[6] = pr; out.m[7] = pi;
    }
    cur.gates.push_back(out);
  }
  close_step(&cur, steps);
}

void canonicalize_steps(ShardLayout *L, std::vector<ShardStep> *steps) {
  const int nl = L->nl, n = L->n;
  ShardStep cur;
  auto local_swap = [&](int a, int b) {  // swap the contents of local physical bits a and b: 3 cx
    if (a == b) return;
    for (int k = 0; k < 3; ++k) {
      QbGate g{};
      g.kind = QB_K_PERM;
      memcpy(g.m, kX, sizeof kX);
      g.ctl_mask = uint64_t(1) << ((k & 1) ? b : a);
      g.target = (k & 1) ? a : b;
      cur.gates.push_back(g);
    }
    std::vector<int> inv = inverse_of(L->perm);
    const int la = inv[size_t(a)], lb = inv[size_t(b)];
    L->perm[size_t(la)] = b;
    L->perm[size_t(lb)] = a;
  };
  // relabelled rank bits first: bring each one local (the exchange un-flips it with a local x); the loop
  // below then puts every qubit back where it belongs
  for (int G = nl; G < n; ++G)
    if (L->flip >> (G - nl) & 1u) exchange_bits(L, G, nl - 1, &cur, steps);
  for (int G = nl; G < n; ++G) {
    if (L->perm[size_t(G)] == G) continue;
    int where = L->perm[size_t(G)];  // physical position of logical bit G
    if (where >= nl) {               // sitting in another rank bit: bring it local first
      exchange_bits(L, where, nl - 1, &cur, steps);
      where = nl - 1;
    }
    if (where < nl - L->window) {  // keep the exchanged half shard in few contiguous runs
      local_swap(where, nl - 1);
      where = nl - 1;
    }
    exchange_bits(L, G, where, &cur, steps);
  }
  for (int q = 0; q < nl; ++q)
    while (L->perm[size_t(q)] != q) local_swap(q, L->perm[size_t(q)]);
  close_step(&cur, steps);
}

std::string steps_to_json(const ShardLayout &L, const std::vector<ShardStep> &steps) {
  std::string s = "{\"n\":" + std::to_string(L.n) + ",\"nl\":" + std::to_string(L.nl) + ",\"rank\":" +
                  std::to_string(L.rank) + ",\"perm\":[";
  char buf[256];
  for (size_t b = 0; b < L.perm.size(); ++b) s += (b ? "," : "") + std::to_string(L.perm[b]);
  s += "],\"flip\":" + std::to_string(L.flip) + ",\"steps\":[";
  for (size_t k = 0; k < steps.size(); ++k) {
    const ShardStep &st = steps[k];
    if (k) s += ",";
    if (st.kind == 1) {
      s += "{\"kind\":1,\"pairs\":[";
      for (size_t j = 0; j < st.rank_bits.size(); ++j) {
        snprintf(buf, sizeof buf, "%s[%d,%d,%d]", j ? "," : "", st.rank_bits[j], st.victims[j], st.lands[j]);
        s += buf;
      }
      s += "]}";
      continue;
    }
    snprintf(buf, sizeof buf, "{\"kind\":0,\"retired\":%lld,\"gates\":[", (long long)st.retired);
    s += buf;
    for (size_t j = 0; j < st.gates.size(); ++j) {
      const QbGate &g = st.gates[j];
      snprintf(buf, sizeof buf, "%s{\"ctl_mask\":%llu,\"target\":%d,\"kind\":%d,\"m\":[", j ? "," : "",
               (unsigned long long)g.ctl_mask, g.target, g.kind);
      s += buf;
      for (int e = 0; e < 8; ++e) {
        snprintf(buf, sizeof buf, "%s%.17g", e ? "," : "", g.m[e]);
        s += buf;
      }
      s += "]}";
    }
    s += "]}";
  }
  s += "]}";
  return s;
}

}  // namespace qb
