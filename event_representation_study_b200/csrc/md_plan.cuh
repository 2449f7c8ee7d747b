// Accumulator plan of the mixed-density tile kernel, buildable at run time (any tuple) and at compile time
// (the ERGO-12 tuples, so that their kernels are fully specialised).
#pragma once
#include "md_device.cuh"

namespace evrep {

// representations/optimized_representation.py:86-115 (v2, active) and :16-66 (v1, commented out)
constexpr int8_t kErgoWin2[12] = {0, 3, 2, 6, 5, 6, 2, 5, 1, 0, 4, 1};
constexpr int8_t kErgoFunc2[12] = {EVREP_FUNC_POLARITY, EVREP_FUNC_TIMESTAMP_NEG, EVREP_FUNC_COUNT_NEG, EVREP_FUNC_POLARITY,
                                   EVREP_FUNC_COUNT_POS, EVREP_FUNC_COUNT, EVREP_FUNC_TIMESTAMP_POS, EVREP_FUNC_COUNT_NEG,
                                   EVREP_FUNC_TIMESTAMP_NEG, EVREP_FUNC_TIMESTAMP_POS, EVREP_FUNC_TIMESTAMP, EVREP_FUNC_COUNT};
constexpr int8_t kErgoAgg2[12] = {EVREP_AGG_VARIANCE, EVREP_AGG_VARIANCE, EVREP_AGG_MEAN, EVREP_AGG_SUM, EVREP_AGG_MEAN, EVREP_AGG_SUM,
                                  EVREP_AGG_MEAN, EVREP_AGG_MEAN, EVREP_AGG_MAX, EVREP_AGG_MAX, EVREP_AGG_MAX, EVREP_AGG_MEAN};
constexpr int8_t kErgoWin1[12] = {0, 2, 2, 3, 5, 0, 0, 4, 2, 6, 1, 1};
constexpr int8_t kErgoFunc1[12] = {EVREP_FUNC_TIMESTAMP, EVREP_FUNC_TIMESTAMP_POS, EVREP_FUNC_TIMESTAMP_NEG, EVREP_FUNC_COUNT_NEG,
                                   EVREP_FUNC_COUNT_POS, EVREP_FUNC_POLARITY, EVREP_FUNC_TIMESTAMP, EVREP_FUNC_COUNT,
                                   EVREP_FUNC_TIMESTAMP_POS, EVREP_FUNC_COUNT, EVREP_FUNC_TIMESTAMP_POS, EVREP_FUNC_TIMESTAMP_NEG};
constexpr int8_t kErgoAgg1[12] = {EVREP_AGG_MAX, EVREP_AGG_SUM, EVREP_AGG_MEAN, EVREP_AGG_SUM, EVREP_AGG_MEAN, EVREP_AGG_VARIANCE,
                                  EVREP_AGG_VARIANCE, EVREP_AGG_SUM, EVREP_AGG_MEAN, EVREP_AGG_SUM, EVREP_AGG_SUM, EVREP_AGG_SUM};

// limb width for windows of at most n_max events: a limb sum must fit 32 bits
constexpr int md_limb_width(int64_t n_max) {
  int nbits = 0;
  while (nbits < 31 && ((int64_t)1 << nbits) <= n_max) ++nbits;  // n_max < 2^nbits
  int lw = 32 - nbits;
  return lw > 31 ? 31 : (lw < 1 ? 1 : lw);
}

// Returns 0, or 1 when the plan needs too many accumulator words.  C in 1..EVREP_MAX_CHANNELS, stacking valid.
// packed = true: 16-bit counters (two per word) and 16-bit limbs; exact as long as a bucket holds < 65536 events.
constexpr int md_plan_build(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking, int lw, bool packed, MdPlan& P) {
  P = MdPlan{};
  P.C = C;
  P.stacking = stacking;
  P.packed = packed ? 1 : 0;
  if (packed) lw = 16;
  P.lw = lw;
  P.nl1 = (31 + lw - 1) / lw;
  P.nl2 = (62 + lw - 1) / lw;
  const int n_win = stacking == EVREP_STACK_SBN ? 7 : 8;
  for (int c = 0; c < C; ++c) {
    MdChan& ch = P.ch[c];
    int wi = win[c];
    if (wi < 0 && wi >= -n_win) wi += n_win;  // the reference indexes a Python list: negative indices wrap
    ch.func = (uint8_t)func[c];
    ch.agg = (uint8_t)agg[c];
    ch.win = (uint8_t)wi;
    ch.g_main = ch.g_pos = ch.g_neg = ch.g_oth = -1;
    // an unknown window / function / aggregation raises inside the reference's make_stack and is
    // swallowed into an all-zero channel (mixed_density_event_stack.py:120-127)
    if (wi < 0 || wi >= n_win || func[c] < 0 || func[c] > EVREP_FUNC_COUNT_NEG || agg[c] < 0 || agg[c] > EVREP_AGG_MIN) {
      ch.valid = 0;
      continue;
    }
    ch.valid = 1;
    const int f = func[c], a = agg[c];
    // requests: need[k] for class k (0 all, 1 pos, 2 neg, 3 neither); `main_cls` names the group md_value starts from
    int need[4] = {0, 0, 0, 0};
    int main_cls = -1;
    bool all_counts = false;  // the channel needs the number of events of every class (= count over "all")
    if (f == EVREP_FUNC_POLARITY) {
      need[1] = need[2] = G_CNT;
      if (a != EVREP_AGG_SUM) need[3] = G_CNT;
    } else {
      const bool is_count = (f == EVREP_FUNC_COUNT || f == EVREP_FUNC_COUNT_POS || f == EVREP_FUNC_COUNT_NEG);
      const int cls = (f == EVREP_FUNC_COUNT || f == EVREP_FUNC_TIMESTAMP) ? 0 : (f == EVREP_FUNC_COUNT_POS || f == EVREP_FUNC_TIMESTAMP_POS) ? 1 : 2;
      int main_need = 0;
      bool count_needed = false;
      if (is_count) {
        if (a == EVREP_AGG_SUM) count_needed = true;
        else if (a != EVREP_AGG_VARIANCE) main_need = G_PRES;  // mean / max / min of ones: "touched"; variance of a constant is 0
      } else {
        main_need = a == EVREP_AGG_SUM ? (G_ST | G_PRES) : a == EVREP_AGG_MEAN ? G_ST : a == EVREP_AGG_MAX ? G_MAX : a == EVREP_AGG_MIN ? G_MIN : (G_ST | G_ST2);
        count_needed = (a == EVREP_AGG_MEAN || a == EVREP_AGG_VARIANCE);
      }
      if (count_needed) {
        if (cls == 0) all_counts = true; else main_need |= G_CNT;
      }
      if (all_counts) need[1] = need[2] = need[3] = G_CNT;
      if (main_need) { need[cls] |= main_need; main_cls = cls; }
    }
    for (int k = 0; k < 4; ++k) {
      if (!need[k]) continue;
      const int bit = k * 8 + wi;
      int g = 0;
      for (; g < P.G; ++g)
        if (P.grp[g].bit == bit) break;
      if (g == P.G) { P.grp[g].bit = (uint8_t)bit; P.grp[g].flags = 0; ++P.G; }
      P.grp[g].flags |= (uint8_t)need[k];
      if (k == main_cls) ch.g_main = (int8_t)g;
      if (f == EVREP_FUNC_POLARITY || all_counts) {
        if (k == 1) ch.g_pos = (int8_t)g; else if (k == 2) ch.g_neg = (int8_t)g; else if (k == 3) ch.g_oth = (int8_t)g;
      }
    }
  }
  // Packed plans pair their 16-bit counters two per word.  An event can bump both counters of a word with ONE
  // shared-memory atomic when it belongs to both groups, so pairs are chosen greedily by how many events hit both
  // (same polarity class - or class "all" - and overlapping index windows), and presence bits become counters when that
  // does not widen the per-pixel footprint: a presence bit costs a read and an atomicOr of its own, a counter that
  // shares a word with one the event bumps anyway costs nothing.
  for (int pass = packed ? 0 : 1; pass < 2; ++pass) {
    const bool convert = pass == 0;  // pass 0: presence -> counter; kept only if the stride does not grow (else pass 1 redoes it)
    MdPlan Q = P;
    int words = 0, pres_bits = 0;
    bool any_pres = false;
    for (int g = 0; g < Q.G; ++g) {
      MdGroup& G = Q.grp[g];
      if (G.flags & G_CNT) G.flags &= (uint8_t)~G_PRES;                          // a count subsumes the presence bit
      if ((G.flags & G_PRES) && (G.flags & (G_MAX | G_MIN))) G.flags &= (uint8_t)~G_PRES;  // so does a latest / earliest-timestamp word
      if ((G.flags & G_PRES) && convert) G.flags = (uint8_t)((G.flags & ~G_PRES) | G_CNT);
      if (G.flags & G_PRES) any_pres = true;
    }
    if (any_pres) Q.w_pres = words++;
    if (packed) {
      // index windows in units of n / 24 (mixed_density_event_stack.py:55-74): W0 all, W1-W3 thirds, W4-W6 nested suffixes
      const int lo[8] = {0, 0, 8, 16, 12, 18, 21, 0}, hi[8] = {24, 8, 16, 24, 24, 24, 24, 0};
      int cg[MD_MAX_GROUPS] = {}, nc = 0;
      bool done[MD_MAX_GROUPS] = {};
      for (int g = 0; g < Q.G; ++g)
        if (Q.grp[g].flags & G_CNT) cg[nc++] = g;
      for (int left = nc; left > 0;) {
        int bi = -1, bj = -1, best = -1;
        for (int i = 0; i < nc; ++i) {
          if (done[i]) continue;
          if (bi < 0) bi = i;  // fallback: the first free counter, alone or with the next free one
          for (int j = i + 1; j < nc; ++j) {
            if (done[j]) continue;
            const int ci = Q.grp[cg[i]].bit >> 3, cj = Q.grp[cg[j]].bit >> 3, wi = Q.grp[cg[i]].bit & 7, wj = Q.grp[cg[j]].bit & 7;
            if (ci != cj && ci != 0 && cj != 0) continue;  // no event is in both
            const int a0 = lo[wi] > lo[wj] ? lo[wi] : lo[wj], a1 = hi[wi] < hi[wj] ? hi[wi] : hi[wj];
            int score = (a1 > a0 ? a1 - a0 : 0) * ((ci == 0 && cj == 0) ? 2 : 1);
            if (ci == 3 || cj == 3) score = score > 0 ? 1 : 0;  // "neither" class: p == 0 next to p == -1, rare
            if (score > best) { best = score; bi = i; bj = j; }
          }
        }
        if (best <= 0) {  // nothing left that merges: pair the remaining counters in order
          bj = -1;
          for (int j = bi + 1; j < nc; ++j)
            if (!done[j]) { bj = j; break; }
        }
        Q.grp[cg[bi]].w_cnt = (uint8_t)words;
        Q.grp[cg[bi]].cnt_shift = 0;
        done[bi] = true;
        --left;
        if (bj >= 0) {
          Q.grp[cg[bj]].w_cnt = (uint8_t)words;
          Q.grp[cg[bj]].cnt_shift = 16;
          done[bj] = true;
          --left;
        }
        ++words;
      }
    } else {
      for (int g = 0; g < Q.G; ++g)
        if (Q.grp[g].flags & G_CNT) Q.grp[g].w_cnt = (uint8_t)words++;
    }
    for (int g = 0; g < Q.G; ++g) {
      MdGroup& G = Q.grp[g];
      if (G.flags & G_PRES) G.pres_bit = (uint8_t)pres_bits++;
      if (G.flags & G_MAX) G.w_max = (uint8_t)words++;
      if (G.flags & G_MIN) G.w_min = (uint8_t)words++;
      if (G.flags & G_ST) { G.w_st = (uint8_t)words; words += Q.nl1; }
      if (G.flags & G_ST2) { G.w_st2 = (uint8_t)words; words += Q.nl2; }
      if (words > 250) return 1;
    }
    if (words == 0) words = 1;
    Q.words = words;
    Q.stride = (words > C ? words : C) | 1;  // odd: bank-conflict-free, and room for the C outputs written in place
    if (convert) {  // compare with the footprint without the conversion
      int base_words = 0, halves = 0;
      bool pres = false;
      for (int g = 0; g < P.G; ++g) {
        uint8_t f = P.grp[g].flags;
        if (f & G_CNT) f &= (uint8_t)~G_PRES;
        if ((f & G_PRES) && (f & (G_MAX | G_MIN))) f &= (uint8_t)~G_PRES;
        if (f & G_PRES) pres = true;
        if (f & G_CNT) ++halves;
        base_words += ((f & G_MAX) ? 1 : 0) + ((f & G_MIN) ? 1 : 0) + ((f & G_ST) ? P.nl1 : 0) + ((f & G_ST2) ? P.nl2 : 0);
      }
      base_words += (pres ? 1 : 0) + (halves + 1) / 2;
      if (base_words == 0) base_words = 1;
      const int base_stride = (base_words > C ? base_words : C) | 1;
      if (Q.stride > base_stride) continue;  // conversion would cost shared memory: redo without it
    }
    P = Q;
    return 0;
  }
  return 0;
}

// compile-time ERGO-12 plans for a menu of limb widths
constexpr int kErgoLimbMenu[] = {16, 14, 12, 10, 8};  // windows below 2^16, 2^18, 2^20, 2^22, 2^24 events
template <int VER, int LW, bool PACKED = false>
struct ErgoPlan {
  static constexpr MdPlan make() {
    MdPlan P{};
    md_plan_build(VER == 2 ? kErgoWin2 : kErgoWin1, VER == 2 ? kErgoFunc2 : kErgoFunc1, VER == 2 ? kErgoAgg2 : kErgoAgg1, 12,
                  EVREP_STACK_SBN, LW, PACKED, P);
    return P;
  }
  static constexpr MdPlan value = make();
};

}  // namespace evrep
