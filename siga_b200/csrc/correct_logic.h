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
// Work split here: the control flow is uniform over the lanes that work on a read and the only parallel primitive is
// Eng::qualify(gen): up to kCorrSlots k-mer queries (a k-mer of the working sequence, optionally with one base
// substituted, and the count it has to reach) answered at once as a bit mask.  The device engine gives each query to
// a 4-lane quad (backward search on the 128-byte rank blocks); the host engine loops.  Counts are a pure function of
// the k-mer, so the reference's per-read k-mer cache (103-124) has no observable effect; after a fix only the k-mers
// covering the fixed base are counted again.
#ifndef SGX_CORRECT_LOGIC_H_
#define SGX_CORRECT_LOGIC_H_

#include "fm_layout.h"

namespace sgx {

static const int kCorrSlots = 8;           // queries per qualify() call = quads per warp
static const int kCorrChunk = 4;           // non-solid bases tried per speculative step (6 queries each)
static const uint32_t kCorrNoSub = 0xffffffffu;
static const int kCorrPhredCutoff = 20;    // CorrectThreshold::_cutoff (src/correct_processor.cpp:41)
static const int kCorrPhredBase = 33;      // Quality::Phred::fromchar (src/quality.h:30-34)

struct CorrectParams {
  uint32_t k;             // kmer-size
  uint32_t rounds;        // kmer-rounds (_maxAttempts)
  uint64_t low, high;     // requiredSupport below / from the phred cut-off (int -> size_t like the reference's return)
  uint64_t count_offset;  // kmer-count-offset
};

struct KmerQuery {
  uint32_t kidx;       // first base of the k-mer
  uint32_t sub_pos;    // base replaced by sub_sym, kCorrNoSub: none
  uint32_t sub_sym;    // rank 1..4
  uint64_t min_count;  // the answer bit is count >= min_count
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

SGX_HD uint64_t corr_max64(uint64_t a, uint64_t b) { return a > b ? a : b; }

// Symbol of the queried k-mer at read position j.
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
  uint32_t* ksolid;      // bit p: k-mer p is solid, ceil(n/32) words
  uint32_t* unsolid;     // bit j: base j is covered by no solid k-mer, ceil(n/32) words
};

// Returns true when the read ends all-solid (the caller then emits st.seq); *n_fixed = number of rounds that changed
// a base.
template <class Eng>
SGX_HD bool correct_read(Eng& eng, const CorrectParams& P, uint32_t n, const CorrectState& st, uint32_t* n_fixed) {
  *n_fixed = 0;
  const uint32_t k = P.k;
  if (n < k) return false;
  const uint32_t nk = n - k + 1;
  const uint32_t nw32 = (n + 31) / 32;
  uint32_t d0 = 0, d1 = nk;  // k-mers whose count is unknown or stale
  uint32_t rounds = 0;
  for (;;) {
    // ---- count the stale k-mers (src/correct_processor.cpp:110-135)
    for (uint32_t base = d0; base < d1; base += kCorrSlots) {
      const uint32_t bits = eng.qualify([&](int slot) {
        KmerQuery q;
        q.kidx = base + (uint32_t)slot;
        q.have = q.kidx < d1;
        q.sub_pos = kCorrNoSub;
        q.sub_sym = 0;
        q.min_count = (q.have && st.hq && corr_bits_all(st.hq, q.kidx, q.kidx + k - 1)) ? P.high : P.low;
        return q;
      });
      if (eng.leader()) {
        for (int s = 0; s < kCorrSlots; ++s) {
          const uint32_t p = base + (uint32_t)s;
          if (p >= d1) break;
          const uint32_t m = 1u << (p & 31u);
          if ((bits >> s) & 1u) st.ksolid[p >> 5] |= m;
          else st.ksolid[p >> 5] &= ~m;
        }
      }
    }
    eng.sync();
    // ---- solid bases
    bool all_solid = true;
    for (uint32_t w = 0; w < nw32; ++w) {
      uint32_t m = 0;
      for (uint32_t b = 0; b < 32; ++b) {
        const uint32_t j = w * 32 + b;
        if (j >= n) break;
        const uint32_t lo = j + 1 >= k ? j + 1 - k : 0, hi = j < nk - 1 ? j : nk - 1;
        if (!corr_bits_any(st.ksolid, lo, hi)) m |= 1u << b;
      }
      if (eng.leader()) st.unsolid[w] = m;
      if (m) all_solid = false;
    }
    eng.sync();
    // ---- src/correct_processor.cpp:137-144
    if (all_solid) return true;
    if (++rounds > P.rounds) return false;
    // ---- fix the leftmost fixable non-solid base (146-170); kCorrChunk bases are evaluated speculatively at once
    bool corrected = false;
    uint32_t cursor = 0;
    while (!corrected) {
      uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
      int nb = 0;
      while (nb < kCorrChunk && cursor < n) {
        const uint32_t w = cursor >> 5;
        const uint32_t rest = st.unsolid[w] & (0xffffffffu << (cursor & 31u));
        if (!rest) {
          cursor = (w + 1) * 32;
          continue;
        }
        uint32_t b = 0;
        while (!((rest >> b) & 1u)) ++b;
        const uint32_t j = w * 32 + b;
        if (nb == 0) p0 = j;
        else if (nb == 1) p1 = j;
        else if (nb == 2) p2 = j;
        else p3 = j;
        ++nb;
        cursor = j + 1;
      }
      if (nb == 0) break;
      uint32_t res = 0;
      for (int call = 0; call * kCorrSlots < nb * 6; ++call) {
        const uint32_t bits = eng.qualify([&](int slot) {
          const int t = call * kCorrSlots + slot;
          const int b = t / 6, side = (t % 6) / 3, a = t % 3;
          KmerQuery q;
          q.have = b < nb;
          const uint32_t i = b == 0 ? p0 : (b == 1 ? p1 : (b == 2 ? p2 : p3));
          q.kidx = side == 0 ? (i + 1 >= k ? i + 1 - k : 0) : (i < n - k ? i : n - k);
          q.sub_pos = i;
          q.sub_sym = q.have ? (uint32_t)corr_alt(read_rank(st.seq, i), a) : 1u;
          q.min_count = corr_max64(P.count_offset, (q.have && st.hq && corr_bit(st.hq, i)) ? P.high : P.low);
          return q;
        });
        res |= bits << (kCorrSlots * call);
      }
      for (int b = 0; b < nb && !corrected; ++b) {
        const uint32_t i = b == 0 ? p0 : (b == 1 ? p1 : (b == 2 ? p2 : p3));
        for (int side = 0; side < 2 && !corrected; ++side) {
          const uint32_t three = (res >> (6 * b + 3 * side)) & 7u;
          if (three == 1u || three == 2u || three == 4u) {  // exactly one base reaches the count (209-224)
            const int a = three == 1u ? 0 : (three == 2u ? 1 : 2);
            const uint32_t alt = (uint32_t)corr_alt(read_rank(st.seq, i), a);
            eng.sync();
            if (eng.leader()) {
              const uint32_t sh = 2 * (i & 15u);
              st.seq[i >> 4] = (st.seq[i >> 4] & ~(3u << sh)) | ((alt - 1u) << sh);
            }
            eng.sync();
            corrected = true;
            d0 = i + 1 >= k ? i + 1 - k : 0;
            d1 = (i < nk - 1 ? i : nk - 1) + 1;
            ++*n_fixed;
          }
        }
      }
    }
    if (!corrected) return false;  // 172-176
  }
}

// ---------------------------------------------------------------------------------------------------------
// Host engine (tests/hostemu only): the queries one after the other, FMIndex::Interval::get literally.
// ---------------------------------------------------------------------------------------------------------
struct HostCorrEng {
  const FmView* fm;
  const uint32_t* seq;
  uint32_t k;
  uint64_t lookups;
  bool leader() const { return true; }
  void sync() const {}
  uint64_t count(const KmerQuery& q) {
    HostGroup g;
    uint32_t j = q.kidx + k - 1;
    int c = corr_query_sym(q, seq, j);
    pos_t lo = fm->pred[c], hi = fm->pred[c + 1] - 1;
    uint64_t L[5], U[5];
    uint32_t steps = k - 1;
    while (steps > 0 && ival_valid(lo, hi)) {
      --j;
      --steps;
      c = corr_query_sym(q, seq, j);
      fm_occ(*fm, lo - 1, L, g);
      fm_occ(*fm, hi, U, g);
      lookups += 2;
      lo = fm->pred[c] + L[c];
      hi = fm->pred[c] + U[c] - 1;
    }
    return ival_valid(lo, hi) ? hi - lo + 1 : 0;
  }
  template <class Gen>
  uint32_t qualify(Gen gen) {
    uint32_t bits = 0;
    for (int s = 0; s < kCorrSlots; ++s) {
      const KmerQuery q = gen(s);
      if (q.have && count(q) >= q.min_count) bits |= 1u << s;
    }
    return bits;
  }
};

}  // namespace sgx

#endif  // SGX_CORRECT_LOGIC_H_
