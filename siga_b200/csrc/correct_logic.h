// Per-read control logic of k-mer error correction -- KmerCorrector::process / try2Correct of the reference
// (src/correct_processor.cpp:72-232) -- written once for the CUDA kernel (correct.cu: one warp per read) and for the
// test-only host stand-in (tests/hostemu).
//
// What the reference does per read of n bases (k = kmer-size):
//   * the n-k+1 k-mers are counted in the FORWARD index only (FMIndex::Interval::occurrences, src/fmindex.h:67-86);
//     a k-mer is solid when its count reaches requiredSupport(min phred over its bases) = minSupport, or minSupport+1
//     when that phred is >= 20 (CorrectThreshold, src/correct_processor.cpp:24-46); a base is solid when a solid
//     k-mer covers it (110-135);
//   * all bases solid -> the (possibly corrected) read is written; otherwise, for at most kmer-rounds rounds, the
//     leftmost non-solid base that try2Correct can fix is fixed: first through its leftmost covering k-mer, then
//     through its rightmost one; a fix needs EXACTLY ONE of the three other bases to give that k-mer a count of at
//     least max(countVector[..] + kmer-count-offset, requiredSupport(phred of the base)) (146-177, 196-227).
//     `countVector` is never filled in the reference (115), so the first term is just kmer-count-offset;
//   * a read that is not all-solid at the end is NOT written at all (PostCorrector, 258-264); reads shorter than k
//     likewise (84-88).
//
// The reference spends (n-k+1)(k-1) backward-search steps per counting round (7200 rank lookups at n = 150, k = 31).
// Only the PREDICATE count >= threshold is observable, and it can be certified with far fewer steps, exactly:
//   * a backward search over a string LONGER than k that still has >= T occurrences proves every k-mer inside it
//     solid (a substring occurs at least as often as its superstring);
//   * a string no longer than k with < T occurrences proves every k-mer that contains it non-solid;
//   * a k-mer searched alone is decided either way.
// corr_resolve() therefore gives each of the kCorrSlots parallel searches a run of still-undecided k-mers: the search
// starts at the right end of the rightmost one (so that k-mer is always decided -- progress is guaranteed) and keeps
// extending to the left through the run while the string is frequent enough; what remains undecided is split again.
// A clean 150-base read is settled by one pass of ~45 steps (15 k-mers per search) instead of 15 passes of 30; the
// k-mers around an error are settled by the two or three searches that die on it.  The solid bits, hence every
// decision and the output, are identical to the reference's (tests/test_cpu_correct.py compares the written files).
//
// Work split: the control flow is uniform over the lanes that work on a read; the parallel primitives are
// Eng::windows(gen, out) (the certificate searches) and Eng::qualify(gen) ("does this k-mer, with one base
// substituted, reach this count?" for try2Correct), each for up to kCorrSlots queries at once.  The device engine
// gives each query to a 4-lane quad (backward search on the 128-byte rank blocks); the host engine loops.
#ifndef SGX_CORRECT_LOGIC_H_
#define SGX_CORRECT_LOGIC_H_

#include "fm_layout.h"

namespace sgx {

static const int kCorrSlots = 8;           // queries per engine call = quads per warp
static const int kCorrAttempts = 8;        // try2Correct attempts evaluated speculatively at once (3 queries each)
static const int kCorrDeadCap = 16;        // remembered rare strings per read
static const uint32_t kCorrDeadNone = 0xffffffffu;
static const uint32_t kCorrNone = 0xffffffffu;
static const int kCorrPhredCutoff = 20;    // CorrectThreshold::_cutoff (src/correct_processor.cpp:41)
static const int kCorrPhredBase = 33;      // Quality::Phred::fromchar (src/quality.h:30-34)
static const int kCorrWinWords = 4 * kCorrSlots;  // result area of Eng::windows: {a, e, drop_low, drop_high} per slot;
                                                  // between certificate passes it holds the try2Correct attempt list

struct CorrectParams {
  uint32_t k;             // kmer-size
  uint32_t rounds;        // kmer-rounds (_maxAttempts)
  uint64_t low, high;     // requiredSupport below / from the phred cut-off (int -> size_t like the reference's return)
  uint64_t count_offset;  // kmer-count-offset
};

// try2Correct query: the k-mer at kidx with base sub_pos replaced by sub_sym; answer bit = count >= min_count.
struct KmerQuery {
  uint32_t kidx;
  uint32_t sub_pos;
  uint32_t sub_sym;    // rank 1..4
  uint64_t min_count;
  bool have;
};

// Certificate search: backward search of read[a .. e] from e towards a.  Result per threshold T in {low, high}:
// drop_T = the largest position i in [a, e] with count(read[i..e]) < T, kCorrNone when there is none (the search
// stops once both are found, the interval is empty, or a is reached).
struct WinQuery {
  uint32_t a, e;
  bool have;
};

// ---- bit ranges over 32-bit words (bit i of the array = bit i&31 of word i>>5), hi inclusive, lo <= hi
SGX_HD uint32_t corr_range_mask(uint32_t w, uint32_t lo, uint32_t hi) {
  // bits of word w that fall into [lo, hi]
  const uint32_t first = w == (lo >> 5) ? (lo & 31u) : 0u, last = w == (hi >> 5) ? (hi & 31u) : 31u;
  return (0xffffffffu << first) & (0xffffffffu >> (31u - last));
}
SGX_HD bool corr_bits_any(const uint32_t* a, uint32_t lo, uint32_t hi) {
  for (uint32_t w = lo >> 5; w <= (hi >> 5); ++w)
    if (a[w] & corr_range_mask(w, lo, hi)) return true;
  return false;
}
SGX_HD bool corr_bits_all(const uint32_t* a, uint32_t lo, uint32_t hi) {
  for (uint32_t w = lo >> 5; w <= (hi >> 5); ++w) {
    const uint32_t m = corr_range_mask(w, lo, hi);
    if ((a[w] & m) != m) return false;
  }
  return true;
}
SGX_HD bool corr_bit(const uint32_t* a, uint32_t i) { return (a[i >> 5] >> (i & 31u)) & 1u; }
// Index of the lowest set bit (z != 0) / of the r-th (0-based) set bit (r < popc(z)).
SGX_HD uint32_t corr_ctz(uint32_t z) {
#ifdef __CUDA_ARCH__
  return (uint32_t)(__ffs((int)z) - 1);
#else
  return (uint32_t)__builtin_ctz(z);
#endif
}
SGX_HD uint32_t corr_nth_bit(uint32_t z, uint32_t r) {
#ifdef __CUDA_ARCH__
  return __fns(z, 0u, (int)r + 1);
#else
  while (r--) z &= z - 1;
  return (uint32_t)__builtin_ctz(z);
#endif
}
SGX_HD void corr_clear_range(uint32_t* a, uint32_t lo, uint32_t hi) {
  for (uint32_t w = lo >> 5; w <= (hi >> 5); ++w) a[w] &= ~corr_range_mask(w, lo, hi);
}
// Number of ZERO bits of a in [lo, hi].
SGX_HD uint32_t corr_count_zeros(const uint32_t* a, uint32_t lo, uint32_t hi) {
  uint32_t c = 0;
  for (uint32_t w = lo >> 5; w <= (hi >> 5); ++w) c += (uint32_t)popc32(~a[w] & corr_range_mask(w, lo, hi));
  return c;
}
// Position of the r-th (0-based) ZERO bit of a in [lo, hi]; the caller guarantees it exists.
SGX_HD uint32_t corr_select_zero(const uint32_t* a, uint32_t lo, uint32_t hi, uint32_t r) {
  for (uint32_t w = lo >> 5; w <= (hi >> 5); ++w) {
    uint32_t z = ~a[w] & corr_range_mask(w, lo, hi);
    const uint32_t c = (uint32_t)popc32(z);
    if (r >= c) {
      r -= c;
      continue;
    }
    return w * 32 + corr_nth_bit(z, r);
  }
  return hi;
}

SGX_HD uint64_t corr_max64(uint64_t a, uint64_t b) { return a > b ? a : b; }

// Symbol of a try2Correct query at read position j.
SGX_HD int corr_query_sym(const KmerQuery& q, const uint32_t* seq, uint32_t j) {
  return j == q.sub_pos ? (int)q.sub_sym : read_rank(seq, j);
}

// The a-th (0..2) base of A,C,G,T other than `cur` (ranks), in the order try2Correct visits them (205-208).
SGX_HD int corr_alt(int cur, int a) {
  const int r = a + 1;
  return r >= cur ? r + 1 : r;
}

// Working state of one read; every array is private to the lanes working on the read (shared memory on the device).
struct CorrectState {
  uint32_t* seq;         // working sequence, 2-bit packed like the read store, ceil(n/16) words
  const uint32_t* hq;    // bit j: phred of base j >= 20; nullptr when the input has no qualities (phred 15)
  uint32_t* khigh;       // bit p: every base of k-mer p has phred >= 20 (it needs `high` occurrences); unused without hq
  uint32_t* known;       // bit p: the solid bit of k-mer p is decided for the current sequence
  uint32_t* ksolid;      // bit p: k-mer p is solid
  uint32_t* unsolid;     // bit j: base j is covered by no solid k-mer
  uint32_t* tmp;         // scratch of corr_dilate                               (all five: ceil(n/32) words)
  uint32_t* win;         // kCorrWinWords words: results of Eng::windows
  uint32_t* dead;        // kCorrDeadCap words: rare strings of the current sequence, start | end << 16 (count < low)
};

// dst bit i = OR of src bits i-d (left) or i+d (right) for d = 0..k-1, over nw words; doubling, log2(k) + 1 rounds of
// word-parallel shift-and-or (every lane owns words lane, lane + lanes, ..).  dst, tmp, src are distinct arrays.
SGX_HD uint32_t corr_shifted_word(const uint32_t* a, uint32_t nw, uint32_t w, uint32_t d, bool left) {
  const uint32_t ws = d >> 5, bs = d & 31u;
  uint32_t v = 0;
  if (left) {
    if (w >= ws) v = a[w - ws] << bs;
    if (bs && w >= ws + 1) v |= a[w - ws - 1] >> (32u - bs);
  } else {
    if (w + ws < nw) v = a[w + ws] >> bs;
    if (bs && w + ws + 1 < nw) v |= a[w + ws + 1] << (32u - bs);
  }
  return v;
}
template <class Eng>
SGX_HD void corr_dilate(Eng& eng, const uint32_t* src, uint32_t* dst, uint32_t* tmp, uint32_t nw, uint32_t k, bool left) {
  const uint32_t* cur = src;
  uint32_t* out = dst;
  uint32_t* other = tmp;
  uint32_t covered = 1;
  while (covered < k) {
    const uint32_t d = covered < k - covered ? covered : k - covered;
    for (uint32_t w = eng.lane_id(); w < nw; w += eng.num_lanes()) out[w] = cur[w] | corr_shifted_word(cur, nw, w, d, left);
    eng.sync();
    covered += d;
    cur = out;
    uint32_t* t = out;
    out = other;
    other = t;
  }
  if (cur != dst) {
    for (uint32_t w = eng.lane_id(); w < nw; w += eng.num_lanes()) dst[w] = cur[w];
    eng.sync();
  }
}

// Rare strings ("dead" strings) found by the certificate searches are remembered per read: read[s..e] with fewer than
// `low` occurrences.  They prove k-mers non-solid after a later fix elsewhere, and they prune try2Correct: substituting
// base i of a k-mer that contains a dead string not covering i leaves the dead string in place, so no substitution
// can lift that k-mer to `low` or more occurrences (src/correct_processor.cpp:196-227 would return false).
SGX_HD bool corr_dead_blocks(const CorrectState& st, uint32_t n_dead, uint32_t k, uint32_t i, uint32_t kidx) {
  const uint32_t m = n_dead < (uint32_t)kCorrDeadCap ? n_dead : (uint32_t)kCorrDeadCap;
  for (uint32_t d = 0; d < m; ++d) {
    const uint32_t v = st.dead[d];
    if (v == kCorrDeadNone) continue;
    const uint32_t ds = v & 0xffffu, de = v >> 16;
    if (ds >= kidx && de <= kidx + k - 1 && (i < ds || i > de)) return true;
  }
  return false;
}

// Decides the solid bit of every k-mer in [d0, d1) (see the header comment).
template <class Eng>
SGX_HD void corr_resolve(Eng& eng, const CorrectParams& P, const CorrectState& st, uint32_t d0, uint32_t d1, uint32_t* n_dead) {
  const uint32_t k = P.k;
  const bool dead_both = P.high >= P.low;  // fewer than `low` occurrences also means fewer than `high`
  if (eng.leader()) {
    corr_clear_range(st.known, d0, d1 - 1);
    corr_clear_range(st.ksolid, d0, d1 - 1);
    // k-mers that contain a remembered rare string are not solid
    const uint32_t m = *n_dead < (uint32_t)kCorrDeadCap ? *n_dead : (uint32_t)kCorrDeadCap;
    for (uint32_t d = 0; d < m; ++d) {
      const uint32_t v = st.dead[d];
      if (v == kCorrDeadNone) continue;
      const uint32_t ds = v & 0xffffu, de = v >> 16;
      if (de - ds + 1 > k) continue;
      const uint32_t first = de >= k - 1 ? de - (k - 1) : 0;
      const uint32_t lo = first > d0 ? first : d0, hi = ds < d1 - 1 ? ds : d1 - 1;
      if (lo > hi) continue;
      for (uint32_t w = lo >> 5; w <= (hi >> 5); ++w) {
        uint32_t mk = corr_range_mask(w, lo, hi);
        if (st.hq && !dead_both) mk &= ~st.khigh[w];
        st.known[w] |= mk;
      }
    }
  }
  eng.sync();
  for (;;) {
    const uint32_t U = corr_count_zeros(st.known, d0, d1 - 1);
    if (U == 0) return;
    const uint32_t per = (U + kCorrSlots - 1) / kCorrSlots;
    eng.windows([&](int slot) {
      WinQuery q;
      const uint32_t r0 = (uint32_t)slot * per;
      q.have = r0 < U;
      q.a = q.e = 0;
      if (q.have) {
        const uint32_t r1 = (r0 + per < U ? r0 + per : U) - 1;
        q.a = corr_select_zero(st.known, d0, d1 - 1, r0);
        q.e = corr_select_zero(st.known, d0, d1 - 1, r1) + k - 1;
      }
      return q;
    }, st.win);
    eng.sync();
    // slot s is applied by lane s: the bit ranges of different slots overlap, hence the atomic or
    for (int s = (int)eng.lane_id(); s < kCorrSlots && (uint32_t)s * per < U; s += (int)eng.num_lanes()) {
      const uint32_t a = st.win[4 * s], e = st.win[4 * s + 1];
      const uint32_t p_hi = e - (k - 1);  // the rightmost k-mer of the search; read[p..e] has >= k symbols for p <= p_hi
      if (st.win[4 * s + 2] != kCorrNone) {
        uint32_t ring = *n_dead;
        for (int s2 = 0; s2 < s; ++s2)
          if (st.win[4 * s2 + 2] != kCorrNone) ++ring;
        st.dead[ring % kCorrDeadCap] = st.win[4 * s + 2] | (e << 16);
      }
      for (int t = 0; t < 2; ++t) {  // t = 0: k-mers that need `low`, t = 1: those that need `high`
        if (t == 1 && !st.hq) break;
        const uint32_t drop = st.win[4 * s + 2 + t];
        // solid: read[p..e] is frequent enough for every p > drop
        const uint32_t f = drop == kCorrNone ? a : drop + 1;
        if (f <= p_hi) {
          for (uint32_t w = f >> 5; w <= (p_hi >> 5); ++w) {
            uint32_t m = corr_range_mask(w, f, p_hi);
            if (st.hq) m &= t ? st.khigh[w] : ~st.khigh[w];
            eng.atomic_or(&st.known[w], m);
            eng.atomic_or(&st.ksolid[w], m);
          }
        }
        // not solid: read[drop..e] is too rare and lies inside every k-mer p in [p_hi, drop]
        if (drop != kCorrNone && drop >= p_hi) {
          const uint32_t last = drop < d1 - 1 ? drop : d1 - 1;
          for (uint32_t w = p_hi >> 5; w <= (last >> 5); ++w) {
            uint32_t m = corr_range_mask(w, p_hi, last);
            if (st.hq) m &= t ? st.khigh[w] : ~st.khigh[w];
            eng.atomic_or(&st.known[w], m);
          }
        }
      }
    }
    eng.sync();
    // every lane keeps the same count of remembered strings
    for (int s = 0; s < kCorrSlots && (uint32_t)s * per < U; ++s)
      if (st.win[4 * s + 2] != kCorrNone) ++*n_dead;
    eng.sync();
  }
}

// Returns true when the read ends all-solid (the caller then emits st.seq); *n_fixed = number of rounds that changed
// a base.
template <class Eng>
SGX_HD bool correct_read(Eng& eng, const CorrectParams& P, uint32_t n, const CorrectState& st, uint32_t* n_fixed) {
  *n_fixed = 0;
  const uint32_t k = P.k;
  if (n < k) return false;
  const uint32_t nk = n - k + 1;
  const uint32_t nw32 = (n + 31) / 32;
  for (uint32_t w = eng.lane_id(); w < nw32; w += eng.num_lanes()) {
    st.known[w] = 0;
    st.ksolid[w] = 0;
    if (st.hq) st.unsolid[w] = ~st.hq[w];  // bases below the phred cut-off (bits past the read end count as such)
  }
  eng.sync();
  if (st.hq) {
    // k-mer p needs `high` unless one of its bases p .. p+k-1 is below the cut-off
    corr_dilate(eng, st.unsolid, st.khigh, st.tmp, nw32, k, false);
    for (uint32_t w = eng.lane_id(); w < nw32; w += eng.num_lanes()) st.khigh[w] = ~st.khigh[w];
  }
  if (eng.leader())
    for (int d = 0; d < kCorrDeadCap; ++d) st.dead[d] = kCorrDeadNone;
  eng.sync();
  uint32_t n_dead = 0;       // rare strings remembered so far (ring of kCorrDeadCap)
  uint32_t d0 = 0, d1 = nk;  // k-mers whose solid bit is unknown or stale
  uint32_t rounds = 0;
  for (;;) {
    // ---- solid k-mers (src/correct_processor.cpp:110-135)
    corr_resolve(eng, P, st, d0, d1, &n_dead);
    // ---- solid bases
    //      base j is solid when one of the k-mers j-k+1 .. j is
    corr_dilate(eng, st.ksolid, st.unsolid, st.tmp, nw32, k, true);
    for (uint32_t w = eng.lane_id(); w < nw32; w += eng.num_lanes())
      st.unsolid[w] = ~st.unsolid[w] & corr_range_mask(w, 0, n - 1);
    eng.sync();
    bool all_solid = true;
    for (uint32_t w = 0; w < nw32; ++w)
      if (st.unsolid[w]) all_solid = false;
    // ---- src/correct_processor.cpp:137-144
    if (all_solid) return true;
    if (++rounds > P.rounds) return false;
    // ---- fix the leftmost fixable non-solid base (146-170).  An attempt = (base, covering k-mer): the leftmost
    //      covering k-mer of a base first, then its rightmost one.  Attempts a remembered rare string rules out are
    //      skipped; kCorrAttempts of the others are evaluated speculatively at once and the first success counts.
    bool corrected = false;
    uint32_t cursor = 0;  // 2 * base + side of the next attempt to look at
    while (!corrected) {
      int na = 0;
      while (na < kCorrAttempts && cursor < 2 * n) {
        const uint32_t j = cursor >> 1, side = cursor & 1u;
        if (side == 0) {
          const uint32_t w = j >> 5;
          const uint32_t rest = st.unsolid[w] & (0xffffffffu << (j & 31u));
          if (!rest) {
            cursor = (w + 1) * 64;
            continue;
          }
          const uint32_t b = corr_ctz(rest);
          if (w * 32 + b != j) {
            cursor = (w * 32 + b) * 2;
            continue;
          }
        }
        ++cursor;
        const uint32_t kidx = side == 0 ? (j + 1 >= k ? j + 1 - k : 0) : (j < n - k ? j : n - k);
        const uint64_t need = corr_max64(P.count_offset, (st.hq && corr_bit(st.hq, j)) ? P.high : P.low);
        if (P.low <= need && corr_dead_blocks(st, n_dead, k, j, kidx)) continue;
        if (eng.leader()) st.win[na] = j | (side << 31);
        ++na;
      }
      if (na == 0) break;
      eng.sync();
      uint32_t res = 0;
      for (int call = 0; call * kCorrSlots < na * 3; ++call) {
        const uint32_t bits = eng.qualify([&](int slot) {
          const int t = call * kCorrSlots + slot;
          const int b = t / 3, a = t % 3;
          KmerQuery q;
          q.have = b < na;
          const uint32_t att = q.have ? st.win[b] : 0u;
          const uint32_t i = att & 0x7fffffffu;
          q.kidx = (att >> 31) == 0 ? (i + 1 >= k ? i + 1 - k : 0) : (i < n - k ? i : n - k);
          q.sub_pos = i;
          q.sub_sym = q.have ? (uint32_t)corr_alt(read_rank(st.seq, i), a) : 1u;
          q.min_count = corr_max64(P.count_offset, (q.have && st.hq && corr_bit(st.hq, i)) ? P.high : P.low);
          return q;
        });
        res |= bits << (kCorrSlots * call);
      }
      for (int b = 0; b < na && !corrected; ++b) {
        const uint32_t three = (res >> (3 * b)) & 7u;
        if (three == 1u || three == 2u || three == 4u) {  // exactly one base reaches the count (209-224)
          const uint32_t i = st.win[b] & 0x7fffffffu;
          const int a = three == 1u ? 0 : (three == 2u ? 1 : 2);
          const uint32_t alt = (uint32_t)corr_alt(read_rank(st.seq, i), a);
          eng.sync();
          if (eng.leader()) {
            const uint32_t sh = 2 * (i & 15u);
            st.seq[i >> 4] = (st.seq[i >> 4] & ~(3u << sh)) | ((alt - 1u) << sh);
            // rare strings that covered the fixed base are gone
            for (int d = 0; d < kCorrDeadCap; ++d) {
              const uint32_t v = st.dead[d];
              if (v != kCorrDeadNone && (v & 0xffffu) <= i && i <= (v >> 16)) st.dead[d] = kCorrDeadNone;
            }
          }
          corrected = true;
          d0 = i + 1 >= k ? i + 1 - k : 0;
          d1 = (i < nk - 1 ? i : nk - 1) + 1;
          ++*n_fixed;
        }
      }
      eng.sync();
    }
    if (!corrected) return false;  // 172-176
  }
}

// ---------------------------------------------------------------------------------------------------------
// Host engine (tests/hostemu only): the queries one after the other on the scalar rank query.
// ---------------------------------------------------------------------------------------------------------
struct HostCorrEng {
  const FmView* fm;
  const uint32_t* seq;
  uint32_t k;
  uint64_t low, high;
  uint64_t lookups;
  // cost model of the device engine: a call lasts as long as its longest query (all quads of a warp step together)
  uint64_t win_calls, win_steps, qual_calls, qual_steps;
  bool leader() const { return true; }
  void sync() const {}
  uint32_t lane_id() const { return 0; }
  uint32_t num_lanes() const { return 1; }
  void atomic_or(uint32_t* p, uint32_t m) const { *p |= m; }
  // FMIndex::Interval::get (src/fmindex.h:67-86) with one sound shortcut: the interval only shrinks, so the search
  // can stop as soon as it is smaller than the count asked for.
  bool reaches(const KmerQuery& q, uint32_t* n_steps) {
    HostGroup g;
    uint32_t j = q.kidx + k - 1;
    int c = corr_query_sym(q, seq, j);
    pos_t lo = fm->pred[c], hi = fm->pred[c + 1] - 1;
    uint64_t L[5], U[5];
    uint32_t steps = k - 1;
    while (steps > 0 && ival_valid(lo, hi) && hi - lo + 1 >= q.min_count) {
      --j;
      --steps;
      c = corr_query_sym(q, seq, j);
      fm_occ(*fm, lo - 1, L, g);
      fm_occ(*fm, hi, U, g);
      lookups += 2;
      lo = fm->pred[c] + L[c];
      hi = fm->pred[c] + U[c] - 1;
    }
    *n_steps = k - 1 - steps;
    return (ival_valid(lo, hi) ? hi - lo + 1 : 0) >= q.min_count;
  }
  template <class Gen>
  uint32_t qualify(Gen gen) {
    uint32_t bits = 0, longest = 0;
    for (int s = 0; s < kCorrSlots; ++s) {
      const KmerQuery q = gen(s);
      uint32_t st = 0;
      if (q.have && reaches(q, &st)) bits |= 1u << s;
      if (st > longest) longest = st;
    }
    ++qual_calls;
    qual_steps += longest;
    return bits;
  }
  template <class Gen>
  void windows(Gen gen, uint32_t* out) {
    HostGroup g;
    uint32_t longest = 0;
    for (int s = 0; s < kCorrSlots; ++s) {
      const WinQuery q = gen(s);
      if (!q.have) continue;
      uint32_t pos = q.e, dl = kCorrNone, dh = kCorrNone;
      int c = read_rank(seq, pos);
      pos_t lo = fm->pred[c], hi = fm->pred[c + 1] - 1;
      uint64_t L[5], U[5];
      for (;;) {
        const bool valid = ival_valid(lo, hi);
        const uint64_t size = valid ? hi - lo + 1 : 0;
        if (dl == kCorrNone && size < low) dl = pos;
        if (dh == kCorrNone && size < high) dh = pos;
        if (!valid || (dl != kCorrNone && dh != kCorrNone) || pos == q.a) break;
        --pos;
        c = read_rank(seq, pos);
        fm_occ(*fm, lo - 1, L, g);
        fm_occ(*fm, hi, U, g);
        lookups += 2;
        lo = fm->pred[c] + L[c];
        hi = fm->pred[c] + U[c] - 1;
      }
      if (q.e - pos > longest) longest = q.e - pos;
      out[4 * s] = q.a;
      out[4 * s + 1] = q.e;
      out[4 * s + 2] = dl;
      out[4 * s + 3] = dh;
    }
    ++win_calls;
    win_steps += longest;
  }
};

}  // namespace sgx

#endif  // SGX_CORRECT_LOGIC_H_
