// Per-read control logic of the overlap path, written once for an 8-lane device group (product) and for the
// single-lane host stand-in (tests/hostemu).  Every function is "group-uniform": all lanes of a group run the same
// path on the same scalar state; the rank query fm_occ() is the only place lanes divide work.  List manipulation
// that only moves memory is done by the group leader between g.sync() points.
//
// Reference behaviour restated here (paths relative to the chungongyu/siga tree):
//   ov_find      OverlapBlockFinder::find                  src/overlap_builder.cpp:846-906
//   ov_update_*  IntervalPair::updateL/updateR             src/overlap_builder.cpp:95-133
//   ov_filter    SubMaximalBlockFilter::filter/resolve     src/overlap_builder.cpp:919-1082
//   ov_extract   IrreducibleBlockListExtractor::extract    src/overlap_builder.cpp:711-832
//   ov_read      OverlapBuilder::overlap (orchestration)   src/overlap_builder.cpp:1113-1182
#ifndef SGX_OVERLAP_LOGIC_H_
#define SGX_OVERLAP_LOGIC_H_

#include "fm_layout.h"

#ifdef __CUDACC__
#define SGX_LOGIC __device__ __forceinline__
#define SGX_LOGIC_NOINLINE __device__ __noinline__
#else
#define SGX_LOGIC inline
#define SGX_LOGIC_NOINLINE inline
#endif

namespace sgx {

// Status bits raised by the per-read logic (OR-ed into a per-batch word).
enum {
  kOvOverflow = 1,   // a work area was too small (host regrows and retries on the GPU)
  kOvAssert = 2,     // the reference would have hit assert(false) (src/overlap_builder.cpp:1004-1008)
  kOvSubstringAbort = 4  // "substring read found during overlap computation" (src/overlap_builder.cpp:754-758)
};

// AlignFlags of the four searches (src/overlap_builder.cpp:52-55), indexed by search mode:
//   0 suffix-prefix  (seq              on fwd)   1 prefix-prefix (revcomp(seq)    on fwd)
//   2 prefix-suffix  (reverse(seq)     on rev)   3 suffix-suffix (complement(seq) on rev)
SGX_LOGIC uint32_t ov_mode_af(int mode) {
  return mode == 0 ? 0u
       : mode == 1 ? (uint32_t)(SGX_AF_QUERYREV | SGX_AF_QUERYCOMP)
       : mode == 2 ? (uint32_t)(SGX_AF_QUERYREV | SGX_AF_TARGETREV)
                   : (uint32_t)(SGX_AF_TARGETREV | SGX_AF_QUERYCOMP);
}

// k-th symbol (rank) of the string a search mode walks, derived from the stored read.
SGX_LOGIC int ov_sym(const uint32_t* rw, uint32_t l, int mode, uint32_t k) {
  if (mode == 0) return read_rank(rw, k);
  if (mode == 1) return comp_rank(read_rank(rw, l - 1 - k));
  if (mode == 2) return read_rank(rw, l - 1 - k);
  return comp_rank(read_rank(rw, k));
}

SGX_LOGIC void ov_store_block(sgx_block* dst, pos_t c0lo, pos_t c0hi, pos_t c1lo, pos_t c1hi, pos_t r0lo, pos_t r0hi,
                              pos_t r1lo, pos_t r1hi, uint32_t len, uint32_t af) {
  dst->capped[0] = c0lo;
  dst->capped[1] = c0hi;
  dst->capped[2] = c1lo;
  dst->capped[3] = c1hi;
  dst->raw[0] = r0lo;
  dst->raw[1] = r0hi;
  dst->raw[2] = r1lo;
  dst->raw[3] = r1hi;
  dst->length = len;
  dst->af = af;
}

// One OverlapBlockFinder::find pass.  P is the index that is searched (the finder's _fmi), S its opposite.
// The probe ('$' step) and the real step of the reference query the same two positions, so each loop iteration
// costs one pair of rank queries.
template <class G>
SGX_LOGIC void ov_find(const FmView& P, const FmView& S, const uint32_t* rw, uint32_t l, int mode, uint32_t min_ov,
                       sgx_block* cand, uint32_t cap, uint32_t* ncand, sgx_block* contain, uint32_t* ncontain,
                       uint8_t* substring, uint32_t* status, uint64_t* lookups, const G& g) {
  const uint32_t af = ov_mode_af(mode);
  uint32_t n = 0;
  uint64_t nlook = 0;
  int c = ov_sym(rw, l, mode, l - 1);
  // IntervalPair::init (src/overlap_builder.cpp:91-94, src/fmindex.h:91-94)
  pos_t lo0 = P.pred[c], hi0 = P.pred[c + 1] - 1;
  pos_t lo1 = S.pred[c], hi1 = S.pred[c + 1] - 1;
  uint64_t L[5], U[5];
  for (uint32_t i = l - 1; i > 0; --i) {
    // An empty interval stays empty under every further update and emits nothing (the reference keeps stepping;
    // those steps are no-ops), so the walk can stop here.
    if (!ival_valid(lo0, hi0)) break;
    fm_occ(P, lo0 - 1, L, g);
    fm_occ(P, hi0, U, g);
    nlook += 2;
    const uint32_t olen = l - i;
    if (olen >= min_ov) {
      const uint64_t d = U[0] - L[0];
      if (d >= 1) {  // probe[1].valid(): some read starts with seq[i..l)
        if (n < cap) {
          if (g.leader())
            ov_store_block(cand + n, P.pred[0] + L[0], P.pred[0] + U[0] - 1, lo1, lo1 + d - 1, lo0, hi0, lo1, hi1, olen, af);
        } else {
          *status |= kOvOverflow;
        }
        ++n;
      }
    }
    c = ov_sym(rw, l, mode, i - 1);
    uint64_t below = 0;
    for (int b = 0; b < c; ++b) below += U[b] - L[b];
    lo1 = lo1 + below;
    hi1 = lo1 + (U[c] - L[c]) - 1;
    lo0 = P.pred[c] + L[c];
    hi0 = P.pred[c] + U[c] - 1;
  }
  *ncand = n < cap ? n : cap;
  *ncontain = 0;
  if (ival_valid(lo0, hi0)) {
    // Containment handling (src/overlap_builder.cpp:877-905).
    uint64_t L2[5], U2[5];
    fm_occ(P, lo0 - 1, L, g);
    fm_occ(P, hi0, U, g);
    fm_occ(S, lo1 - 1, L2, g);
    fm_occ(S, hi1, U2, g);
    nlook += 4;
    bool dna = false;
    for (int b = 1; b < 5; ++b) dna = dna || (U[b] - L[b]) > 0 || (U2[b] - L2[b]) > 0;
    if (dna) {
      if (g.leader()) *substring = 1;
    } else {
      const uint64_t d = U[0] - L[0];
      pos_t p0lo = P.pred[0] + L[0], p0hi = P.pred[0] + U[0] - 1;
      pos_t p1lo = lo1, p1hi = lo1 + d - 1;
      if (ival_valid(p0lo, p0hi) && ival_valid(p1lo, p1hi)) {
        // probe.updateR('$', S): positions p1lo-1 (== lo1-1, already fetched) and p1hi
        uint64_t U3[5];
        if (p1hi == hi1) {
          for (int b = 0; b < 5; ++b) U3[b] = U2[b];
        } else {
          fm_occ(S, p1hi, U3, g);
          nlook += 1;
        }
        const uint64_t d3 = U3[0] - L2[0];
        p0hi = p0lo + d3 - 1;
        p1lo = S.pred[0] + L2[0];
        p1hi = S.pred[0] + U3[0] - 1;
        if (g.leader()) ov_store_block(contain, p0lo, p0hi, p1lo, p1hi, lo0, hi0, lo1, hi1, l, af);
        *ncontain = 1;
      }
    }
  }
  *lookups += nlook;
}

// ---------------------------------------------------------------------------------------------------------
// Interval-pair updates on a stored block (capped pair), src/overlap_builder.cpp:95-133.
// ---------------------------------------------------------------------------------------------------------

// OverlapBlock::index (src/overlap_builder.cpp:177-179): the index that extends the block to the right.
SGX_LOGIC const FmView& ov_ext_index(uint32_t af, const FmView& fwd, const FmView& rev) {
  return (af & SGX_AF_TARGETREV) ? fwd : rev;
}

// capped.updateR(c, index) given the two rank vectors at capped[1].lower-1 / capped[1].upper.
SGX_LOGIC void ov_apply_updateR(pos_t cp[4], int c, const FmView& idx, const uint64_t L[5], const uint64_t U[5]) {
  uint64_t below = 0;
  for (int b = 0; b < c; ++b) below += U[b] - L[b];
  cp[0] = cp[0] + below;
  cp[1] = cp[0] + (U[c] - L[c]) - 1;
  cp[2] = idx.pred[c] + L[c];
  cp[3] = idx.pred[c] + U[c] - 1;
}

// capped.updateL(c, index) given the rank vectors at capped[0].lower-1 / capped[0].upper.
SGX_LOGIC void ov_apply_updateL(pos_t cp[4], int c, const FmView& idx, const uint64_t L[5], const uint64_t U[5]) {
  uint64_t below = 0;
  for (int b = 0; b < c; ++b) below += U[b] - L[b];
  cp[2] = cp[2] + below;
  cp[3] = cp[2] + (U[c] - L[c]) - 1;
  cp[0] = idx.pred[c] + L[c];
  cp[1] = idx.pred[c] + U[c] - 1;
}

SGX_LOGIC bool ov_pair_valid(const pos_t cp[4]) { return ival_valid(cp[0], cp[1]) && ival_valid(cp[2], cp[3]); }

// Interval::isIntersecting, src/coord.h:42-45.
SGX_LOGIC bool ov_intersect(pos_t s1, pos_t e1, pos_t s2, pos_t e2) { return !(s1 > e2 || s2 > e1); }

SGX_LOGIC void ov_copy_block(sgx_block* dst, const sgx_block* src) { *dst = *src; }

// Stable insertion sort by capped[0].lower ascending (IntervalLeftSorter, src/overlap_builder.cpp:1084-1089).
SGX_LOGIC void ov_sort_left(sgx_block* a, uint32_t n) {
  for (uint32_t i = 1; i < n; ++i) {
    sgx_block t = a[i];
    uint32_t j = i;
    while (j > 0 && a[j - 1].capped[0] > t.capped[0]) {
      a[j] = a[j - 1];
      --j;
    }
    a[j] = t;
  }
}

// Stable insertion sort by length descending (OverlapBlockLengthSorter, src/overlap_builder.cpp:812-817).
SGX_LOGIC void ov_sort_length(sgx_block* a, uint32_t n) {
  for (uint32_t i = 1; i < n; ++i) {
    sgx_block t = a[i];
    uint32_t j = i;
    while (j > 0 && a[j - 1].length < t.length) {
      a[j] = a[j - 1];
      --j;
    }
    a[j] = t;
  }
}

// SubMaximalBlockFilter::resolve (src/overlap_builder.cpp:965-1082).  Appends to R[0..*nR).
template <class G>
SGX_LOGIC_NOINLINE void ov_resolve(const sgx_block* x, const sgx_block* y, sgx_block* R, uint32_t* nR, uint32_t capR,
                                   const FmView& P, const FmView& S, uint32_t* status, uint64_t* lookups, const G& g) {
  const sgx_block* higher = x;
  const sgx_block* lower = y;
  if (higher->length < lower->length) {
    const sgx_block* t = higher;
    higher = lower;
    lower = t;
  }
  uint32_t n = *nR;
  if (n < capR) {
    if (g.leader()) ov_copy_block(R + n, higher);
  } else {
    *status |= kOvOverflow;
  }
  ++n;
  if (higher->length == lower->length) {
    if (higher->capped[0] != lower->capped[0] || higher->capped[1] != lower->capped[1]) *status |= kOvAssert;
  } else if (lower->capped[0] < higher->capped[0] || lower->capped[1] > higher->capped[1]) {
    // Partial overlap: map every reverse-side row of the lower block to its forward row by walking the opposite
    // BWT back to the read start (rare: periodic / low-complexity sequence only).
    const int kMapCap = 32;
    pos_t mapk[kMapCap], mapv[kMapCap];
    int nmap = 0;
    uint64_t L[5], U[5];
    for (pos_t j = lower->capped[2]; j <= lower->capped[3]; ++j) {
      pos_t rg[4] = {lower->raw[0], lower->raw[1], lower->raw[2], lower->raw[3]};
      pos_t tlo = j, thi = j;
      bool done = false;
      while (!done) {
        const int c = fm_sym(S, tlo, g);
        if (c == kDollar) {
          fm_occ(P, rg[0] - 1, L, g);
          fm_occ(P, rg[1], U, g);
          *lookups += 2;
          ov_apply_updateL(rg, kDollar, P, L, U);
          done = true;
        }
        // tracing.update(c, S) (src/fmindex.h:95-99)
        fm_occ(S, tlo - 1, L, g);
        fm_occ(S, thi, U, g);
        tlo = S.pred[c] + L[c];
        thi = S.pred[c] + U[c] - 1;
        // ranges.updateR(c, S)
        fm_occ(S, rg[2] - 1, L, g);
        fm_occ(S, rg[3], U, g);
        *lookups += 4;
        ov_apply_updateR(rg, c, S, L, U);
      }
      pos_t forward;
      if (rg[0] == rg[1]) {
        forward = rg[0];
      } else {
        // duplicated read: hand out the forward rows of its interval one by one
        const pos_t k = rg[0];
        pos_t idx = k;
        int at = -1;
        for (int q = 0; q < nmap; ++q)
          if (mapk[q] == k) at = q;
        if (at >= 0) {
          idx = mapv[at];
          mapv[at] = idx + 1;
        } else if (nmap < kMapCap) {
          mapk[nmap] = k;
          mapv[nmap] = idx + 1;
          ++nmap;
        } else {
          *status |= kOvOverflow;
        }
        forward = idx;
      }
      if (!ov_intersect(forward, forward, higher->capped[0], higher->capped[1])) {
        if (n < capR) {
          if (g.leader()) {
            ov_copy_block(R + n, lower);
            R[n].capped[0] = forward;
            R[n].capped[1] = forward;
            R[n].capped[2] = j;
            R[n].capped[3] = j;
          }
        } else {
          *status |= kOvOverflow;
        }
        ++n;
      }
    }
  }
  *nR = n < capR ? n : capR;
}

// SubMaximalBlockFilter::filter (src/overlap_builder.cpp:919-954) on a[0..*n); R is scratch for resolve().
template <class G>
SGX_LOGIC void ov_filter(sgx_block* a, uint32_t* n_io, uint32_t cap, sgx_block* R, uint32_t capR, const FmView& P,
                         const FmView& S, uint32_t* status, uint64_t* lookups, const G& g) {
  uint32_t n = *n_io;
  if (n == 0) return;
  if (g.leader()) ov_sort_left(a, n);
  g.sync();
  uint32_t prev = 0;
  while (prev + 1 < n) {
    const uint32_t curr = prev + 1;
    if (ov_intersect(a[prev].capped[0], a[prev].capped[1], a[curr].capped[0], a[curr].capped[1])) {
      uint32_t nR = 0;
      ov_resolve(a + prev, a + curr, R, &nR, capR, P, S, status, lookups, g);
      g.sync();
      if (n - 2 + nR > cap) {
        *status |= kOvOverflow;
        break;
      }
      if (g.leader()) {
        ov_sort_left(R, nR);
        // erase curr and prev
        for (uint32_t k = prev; k + 2 < n; ++k) a[k] = a[k + 2];
        const uint32_t m = n - 2;
        // stable merge (list::merge keeps elements of *this before equal elements of the argument)
        uint32_t ia = m, ir = nR, io = m + nR;
        while (ir > 0) {
          if (ia > 0 && a[ia - 1].capped[0] > R[ir - 1].capped[0]) {
            a[--io] = a[--ia];
          } else {
            a[--io] = R[--ir];
          }
        }
      }
      n = n - 2 + nR;
      g.sync();
      prev = 0;
    } else {
      ++prev;
    }
  }
  *n_io = n;
}

// ContainmentBlockRemover (src/overlap_builder.cpp:1094-1111).
template <class G>
SGX_LOGIC void ov_remove_contained(sgx_block* a, uint32_t* n_io, uint32_t seqlen, const G& g) {
  uint32_t n = *n_io, m = 0;
  for (uint32_t k = 0; k < n; ++k) {
    if (a[k].length != seqlen) {
      if (g.leader() && m != k) a[m] = a[k];
      ++m;
    }
  }
  g.sync();
  *n_io = m;
}

// Rank vectors bracketing capped[1] of a block on its right-extension index, complemented to query orientation
// when asked (OverlapBlock::ext, src/overlap_builder.cpp:181-187).
template <class G>
SGX_LOGIC void ov_block_occ(const sgx_block* b, const FmView& fwd, const FmView& rev, uint64_t L[5], uint64_t U[5],
                            uint64_t* lookups, const G& g) {
  const FmView& idx = ov_ext_index(b->af, fwd, rev);
  fm_occ(idx, b->capped[2] - 1, L, g);
  fm_occ(idx, b->capped[3], U, g);
  *lookups += 2;
}

SGX_LOGIC void ov_add_ext(uint64_t exts[5], uint32_t af, const uint64_t L[5], const uint64_t U[5]) {
  if (af & SGX_AF_QUERYCOMP) {
    exts[0] += U[0] - L[0];
    for (int b = 1; b < 5; ++b) exts[b] += U[5 - b] - L[5 - b];
  } else {
    for (int b = 0; b < 5; ++b) exts[b] += U[b] - L[b];
  }
}

// IrreducibleBlockListExtractor::extract (src/overlap_builder.cpp:711-809).
//   st[0..n)   the merged candidate list (already copied to the bottom of the work area) ; st capacity cap blocks
//   gd         group descriptors: 4 arrays of cap_g entries (groups off/cnt, incoming off/cnt)
// The reference keeps a std::list of block groups and walks it with a loop that increments its iterator twice
// per visit and wraps through the list sentinel (src/overlap_builder.cpp:728, 797-801).
// Groups evolve independently, so the emitted SET does not depend on the visiting order -- except when the
// "substring read found" abort fires (src/overlap_builder.cpp:752-758), which keeps whatever was emitted before
// it.  To stay identical in that case too, the very same visiting order is reproduced here.  Children of a
// branching group are bump-allocated above everything else; space of finished groups is not reclaimed (lists are
// a few dozen blocks; if the area overflows the host regrows it and the read is redone on the GPU).
template <class G>
SGX_LOGIC_NOINLINE void ov_extract(sgx_block* st, uint32_t n, uint32_t cap, uint32_t* gd, uint32_t cap_g,
                                   sgx_block* out, uint32_t* nout_io, uint32_t capOut, const FmView& fwd,
                                   const FmView& rev, uint32_t* status, uint64_t* lookups, const G& g) {
  if (n == 0) return;
  uint32_t* g_off = gd;
  uint32_t* g_cnt = gd + cap_g;
  uint32_t* in_off = gd + 2 * (size_t)cap_g;
  uint32_t* in_cnt = gd + 3 * (size_t)cap_g;
  if (g.leader()) {
    ov_sort_length(st, n);
    g_off[0] = 0;
    g_cnt[0] = n;
  }
  g.sync();
  uint32_t k = 1;    // live groups
  uint32_t top = n;  // bump pointer
  uint32_t nout = *nout_io;
  uint64_t L[5], U[5];
  bool stop = false;
  while (k > 0 && !stop) {
    uint32_t nin = 0;
    uint32_t i = 0;
    while (i != k && !stop) {
      const uint32_t s = g_off[i];
      const uint32_t e = s + g_cnt[i];
      bool erase_group = true;
      if (e > s) {
        const uint32_t top_len = st[s].length;
        uint64_t exts[5] = {0, 0, 0, 0, 0};
        uint32_t ntop = 0;
        for (uint32_t j = s; j < e && st[j].length == top_len; ++j) {
          ov_block_occ(st + j, fwd, rev, L, U, lookups, g);
          ov_add_ext(exts, st[j].af, L, U);
          ++ntop;
        }
        if (exts[0] > 0) {
          // The top-level block has ended: emit every top-level block, made terminal by a '$' right-update.
          for (uint32_t j = s; j < s + ntop; ++j) {
            ov_block_occ(st + j, fwd, rev, L, U, lookups, g);
            if (U[0] - L[0] == 0) {
              *status |= kOvSubstringAbort;
              stop = true;  // the reference returns false here and keeps what it already emitted
              break;
            }
            if (nout < capOut) {
              if (g.leader()) {
                sgx_block b = st[j];
                ov_apply_updateR(b.capped, kDollar, ov_ext_index(b.af, fwd, rev), L, U);
                out[nout] = b;
              }
            } else {
              *status |= kOvOverflow;
            }
            ++nout;
          }
        } else {
          for (uint32_t j = s + ntop; j < e; ++j) {
            ov_block_occ(st + j, fwd, rev, L, U, lookups, g);
            ov_add_ext(exts, st[j].af, L, U);
          }
          int nz = 0, first = -1;
          for (int b = 0; b < 5; ++b)
            if (exts[b] > 0) {
              ++nz;
              if (first < 0) first = b;
            }
          if (nz == 1) {
            // Unique extension: update every block in place, dropping the ones that become invalid.
            uint32_t m = s;
            for (uint32_t j = s; j < e; ++j) {
              sgx_block b = st[j];
              const int c = (b.af & SGX_AF_QUERYCOMP) ? comp_rank(first) : first;
              const FmView& idx = ov_ext_index(b.af, fwd, rev);
              ov_block_occ(&b, fwd, rev, L, U, lookups, g);
              ov_apply_updateR(b.capped, c, idx, L, U);
              if (ov_pair_valid(b.capped)) {
                g.sync();
                if (g.leader()) st[m] = b;
                ++m;
              }
            }
            g.sync();
            if (g.leader()) g_cnt[i] = m - s;
            g.sync();
            erase_group = false;
          } else {
            // Branch: one copy of the group per extension symbol, queued behind the current pass.
            for (int x = 0; x < 5 && !stop; ++x) {
              if (exts[x] == 0) continue;
              const uint32_t cstart = top;
              for (uint32_t j = s; j < e; ++j) {
                sgx_block b = st[j];
                const int c = (b.af & SGX_AF_QUERYCOMP) ? comp_rank(x) : x;
                const FmView& idx = ov_ext_index(b.af, fwd, rev);
                ov_block_occ(&b, fwd, rev, L, U, lookups, g);
                ov_apply_updateR(b.capped, c, idx, L, U);
                if (ov_pair_valid(b.capped)) {
                  if (top >= cap) {
                    *status |= kOvOverflow;
                    stop = true;
                    break;
                  }
                  if (g.leader()) st[top] = b;
                  ++top;
                }
              }
              if (stop) break;
              // the reference queues the copy even if every block of it died (it is erased when visited)
              if (nin >= cap_g) {
                *status |= kOvOverflow;
                stop = true;
                break;
              }
              if (g.leader()) {
                in_off[nin] = cstart;
                in_cnt[nin] = top - cstart;
              }
              ++nin;
            }
            g.sync();
          }
        }
      }
      if (stop) break;
      if (erase_group) {
        if (g.leader())
          for (uint32_t q = i; q + 1 < k; ++q) {
            g_off[q] = g_off[q + 1];
            g_cnt[q] = g_cnt[q + 1];
          }
        g.sync();
        --k;  // i now designates the element after the erased one
      } else {
        ++i;
      }
      // the for-loop's own ++i ; incrementing end() of a std::list wraps to begin()
      if (i == k) {
        i = 0;
      } else {
        ++i;
      }
      if (k == 0) break;
    }
    if (stop) break;
    // splice in the newly branched groups
    if (k + nin > cap_g) {
      *status |= kOvOverflow;
      break;
    }
    if (g.leader())
      for (uint32_t q = 0; q < nin; ++q) {
        g_off[k + q] = in_off[q];
        g_cnt[k + q] = in_cnt[q];
      }
    g.sync();
    k += nin;
  }
  *nout_io = nout;
}

// Work area of one group for the post-search stage.
struct OvArena {
  sgx_block* lists;  // 4 lists x list_cap
  uint32_t list_cap;
  sgx_block* R;      // resolve scratch
  uint32_t r_cap;
  sgx_block* st;     // extraction stack
  uint32_t st_cap;
  uint32_t* gd;      // group descriptors, 4 x g_cap entries
  uint32_t g_cap;
};

// Everything OverlapBuilder::overlap does after the four searches (src/overlap_builder.cpp:1135-1179).
//   cand[f], ncand[f]      blocks of search mode f ; contain[f] (ncontain[f] in {0,1}) its containment block
// List naming: mode 0 = suffixfwd, 1 = prefixfwd, 2 = prefixrev, 3 = suffixrev.
template <class G>
SGX_LOGIC void ov_post(const FmView& fwd, const FmView& rev, uint32_t seqlen, int irreducible,
                       const sgx_block* const cand[4], const uint32_t ncand[4], const sgx_block* contain,
                       const uint32_t ncontain[4], const OvArena& ar, sgx_block* out, uint32_t* nout_io,
                       uint32_t capOut, uint32_t* status, uint64_t* lookups, const G& g) {
  uint32_t nl[4];
  // lists = own candidates + the containment blocks of the same index side (fwd: modes 0,1 ; rev: modes 2,3)
  for (int f = 0; f < 4; ++f) {
    sgx_block* dst = ar.lists + (size_t)f * ar.list_cap;
    uint32_t n = 0;
    const int side0 = (f < 2) ? 0 : 2;
    const uint32_t want = ncand[f] + ncontain[side0] + ncontain[side0 + 1];
    if (want > ar.list_cap) {
      *status |= kOvOverflow;
      nl[f] = 0;
      continue;
    }
    if (g.leader()) {
      for (uint32_t k = 0; k < ncand[f]; ++k) dst[n + k] = cand[f][k];
    }
    n += ncand[f];
    for (int q = side0; q < side0 + 2; ++q) {
      if (ncontain[q]) {
        if (g.leader()) dst[n] = contain[q];
        ++n;
      }
    }
    nl[f] = n;
  }
  g.sync();
  // sub-maximal filter: fwd-rank lists with (fwd, rev), rev-rank lists with (rev, fwd)
  for (int f = 0; f < 4; ++f) {
    sgx_block* a = ar.lists + (size_t)f * ar.list_cap;
    if (f < 2) {
      ov_filter(a, &nl[f], ar.list_cap, ar.R, ar.r_cap, fwd, rev, status, lookups, g);
    } else {
      ov_filter(a, &nl[f], ar.list_cap, ar.R, ar.r_cap, rev, fwd, status, lookups, g);
    }
    ov_remove_contained(a, &nl[f], seqlen, g);
  }
  // containments go straight to the output: containfwd then containrev
  uint32_t nout = *nout_io;
  for (int q = 0; q < 4; ++q) {
    if (ncontain[q]) {
      if (nout < capOut) {
        if (g.leader()) out[nout] = contain[q];
      } else {
        *status |= kOvOverflow;
      }
      ++nout;
    }
  }
  if (irreducible) {
    // right extensions: suffixfwd + suffixrev ; left extensions: prefixfwd + prefixrev
    const int pairs[2][2] = {{0, 3}, {1, 2}};
    for (int side = 0; side < 2; ++side) {
      const int fa = pairs[side][0], fb = pairs[side][1];
      const uint32_t n = nl[fa] + nl[fb];
      if (n > ar.st_cap) {
        *status |= kOvOverflow;
        continue;
      }
      if (g.leader()) {
        const sgx_block* a = ar.lists + (size_t)fa * ar.list_cap;
        const sgx_block* b = ar.lists + (size_t)fb * ar.list_cap;
        for (uint32_t k = 0; k < nl[fa]; ++k) ar.st[k] = a[k];
        for (uint32_t k = 0; k < nl[fb]; ++k) ar.st[nl[fa] + k] = b[k];
      }
      g.sync();
      ov_extract(ar.st, n, ar.st_cap, ar.gd, ar.g_cap, out, &nout, capOut, fwd, rev, status, lookups, g);
      g.sync();
    }
  } else {
    const int order[4] = {0, 3, 1, 2};  // suffixfwd, suffixrev, prefixfwd, prefixrev
    for (int q = 0; q < 4; ++q) {
      const sgx_block* a = ar.lists + (size_t)order[q] * ar.list_cap;
      for (uint32_t k = 0; k < nl[order[q]]; ++k) {
        if (nout < capOut) {
          if (g.leader()) out[nout] = a[k];
        } else {
          *status |= kOvOverflow;
        }
        ++nout;
      }
    }
  }
  *nout_io = nout < capOut ? nout : capOut;
}

}  // namespace sgx

#endif  // SGX_OVERLAP_LOGIC_H_
