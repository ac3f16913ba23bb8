// The search kernel: OverlapBlockFinder::find (src/overlap_builder.cpp:846-906 of the reference) for every
// (read, search mode) pair, tuned for issue rate.  Same arithmetic as ov_find() in overlap_logic.h (the validated
// restatement the slow path and the host stand-in run), different work split:
//
//   * FOUR lanes per chain, eight chains per warp.  Lane q of a quad loads uint4 q and q+4 of a rank block, i.e. a
//     quarter of its symbols plus the low and high header word of symbol q+1 (fm_layout.h), for BOTH ends of the
//     interval (lower-1 and upper).  When both ends fall into the same block the second pair of loads is predicated
//     off, so a narrow interval costs one 128-byte line per backward step.
//   * partial popcounts are packed two per register and summed with two xor-shuffles per word; each lane then owns
//     the complete rank of "its" symbol at both ends, and only the three quantities the interval update needs
//     (ranks of the stepped symbol, the '$' difference, the sum of differences below the symbol) are exchanged.
//   * positions are 32-bit when the BWT has < 2^32 symbols (IdxT), 64-bit otherwise; no local memory.
//   * all quads of a warp walk reads of the same length in lock-step, so there is no divergence on fixed-length input.
#ifndef SGX_OVERLAP_SEARCH_CUH_
#define SGX_OVERLAP_SEARCH_CUH_

#include "overlap_logic.h"
#include "quad_rank.cuh"

namespace sgx {

template <class IdxT>
__device__ __forceinline__ void quad_store_block(sgx_block* dst, IdxT c0lo, IdxT c0hi, IdxT c1lo, IdxT c1hi, IdxT r0lo,
                                                 IdxT r0hi, IdxT r1lo, IdxT r1hi, uint32_t len, uint32_t af) {
  // positions are stored as the reference prints them: 64-bit, with -1 staying -1
  const bool wide = sizeof(IdxT) == 8;
  auto ext = [wide](IdxT v) -> uint64_t { return (!wide && v == (IdxT)-1) ? (uint64_t)-1 : (uint64_t)v; };
  ov_store_block(dst, ext(c0lo), ext(c0hi), ext(c1lo), ext(c1hi), ext(r0lo), ext(r0hi), ext(r1lo), ext(r1hi), len, af);
}

// Both indices in one kernel parameter, so that "the searched index" is a dynamic subscript into parameter space
// (a constant-bank load) rather than a pointer to a parameter copy in local memory.
struct FmPair {
  const FmBlockLane* blocks[2];  // 0 = forward index, 1 = reverse index
  uint64_t pred[2][6];
  const uint32_t* inv[2];        // read id -> row of its '$' block entry (inverse .sai / .rsai), when reads are attached
  const uint2* ktab[2];          // k-mer start tables of the two indices (sgx_fm::ktab), or nullptr
  int ktab_k;                    // their k (0: no table)
};

// LEAN: write 12-byte LeanCand records (capped[0], length, flags) instead of full 72-byte blocks.
// list (optional): work item wi searches read list[wi]; its results are stored at item index wi.
//
// Control flow is WARP-uniform: every loop runs while any quad of the warp still has work and finished quads ride
// along with their loads pointed at block 0 and their state frozen.  That allows full-mask shuffles (a quad-mask
// shuffle costs a MATCH/REDUX/VOTE/branch sequence per use to verify convergence) and costs nothing on fixed-length
// input, where all quads of a warp finish together.
template <class IdxT, bool LEAN>
__global__ void __launch_bounds__(256, 4)
ov_search_quad_kernel(FmPair fm, ReadsView reads, uint64_t first, uint64_t count, uint32_t min_ov, int rc,
                      OvBatch bt, const uint32_t* __restrict__ list, unsigned long long* __restrict__ lookups) {
  const unsigned kFull = 0xffffffffu;
  const int wl = threadIdx.x & 31;
  const int q = wl & (kQuad - 1);
  const int qbase = wl & ~(kQuad - 1);
  const uint64_t quad = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kQuad;
  const uint64_t nquads = ((uint64_t)gridDim.x * blockDim.x) / kQuad;
  const int nmodes = rc ? 4 : 2;
  const uint64_t nwork = count * nmodes;
  unsigned long long nlook = 0, nlines = 0;
  const IdxT kNone = (IdxT)-1;

  for (uint64_t wk = quad; __any_sync(kFull, wk < nwork); wk += nquads) {
    const bool have = wk < nwork;
    const uint64_t wi = have ? wk / nmodes : 0;         // where the results are stored
    const uint64_t r = list ? (uint64_t)list[wi] : wi;  // which read (relative to `first`)
    const int mode = rc ? (int)(wk % nmodes) : (int)(wk % nmodes) * 2;  // without rc only modes 0 and 2 run
    const int pi = mode < 2 ? 0 : 1, si = 1 - pi;       // searched index / its opposite
    const FmBlockLane* __restrict__ Pb = fm.blocks[pi];
    const FmBlockLane* __restrict__ Sb = fm.blocks[si];
#define Ppred fm.pred[pi]
#define Spred fm.pred[si]
    const uint32_t* rw = reads.words + reads.woff[first + r];
    const uint32_t l = reads.len[first + r];
    const uint32_t af = ov_mode_af(mode);
    sgx_block* cand = LEAN ? nullptr : bt.cand + ((size_t)wi * 4 + mode) * bt.find_cap;
    LeanCand* lean = LEAN ? bt.lean + ((size_t)wi * 4 + mode) * bt.find_cap : nullptr;
    uint32_t n = 0;
    // the searched string walks the stored read forwards or backwards, complemented or not (ov_sym)
    const bool backwards = mode == 1 || mode == 2;  // string position k reads base l-1-k
    const uint32_t cm = (mode == 1 || mode == 3) ? 3u : 0u;
    // C[] of the searched index in registers (selected by symbol without indexing memory inside the loop)
    const IdxT pc0 = (IdxT)Ppred[0], pc1 = (IdxT)Ppred[1], pc2 = (IdxT)Ppred[2], pc3 = (IdxT)Ppred[3], pc4 = (IdxT)Ppred[4];
    int c = ov_sym(rw, l, mode, l - 1);
    IdxT lo0 = (IdxT)Ppred[c], hi0 = (IdxT)(Ppred[c + 1] - 1);
    IdxT lo1 = (IdxT)Spred[c], hi1 = (IdxT)(Spred[c + 1] - 1);
    // base index of string position l-2 and its step as the string position decreases
    uint32_t j = backwards ? 1u : l - 2;
    const uint32_t dj = backwards ? 1u : 0xffffffffu;
    uint32_t i = have ? l - 1 : 0;  // remaining steps
    bool live = have && i > 0 && hi0 != kNone && hi0 >= lo0;
    int cnext = live ? (int)((((rw[j >> 4] >> (2 * (j & 15))) & 3u) ^ cm) + 1u) : 1;
#pragma unroll 1
    while (__any_sync(kFull, live)) {
      c = cnext;
      j += dj;
      if (live && i > 1) cnext = (int)((((rw[j >> 4] >> (2 * (j & 15))) & 3u) ^ cm) + 1u);  // next step's symbol
      // a finished quad keeps its state and points its loads at block 0
      const QuadRank<IdxT> k = quad_rank<IdxT>(Pb, live ? lo0 : (IdxT)0, live ? (IdxT)(hi0 + 1) : (IdxT)0, q);
      // ranks of the stepped symbol and the sum of differences of the smaller symbols
      const IdxT Lc = __shfl_sync(kFull, k.Lq, qbase + c - 1);
      const IdxT Uc = __shfl_sync(kFull, k.Uq, qbase + c - 1);
      IdxT below = (q < c - 1) ? (IdxT)(k.Uq - k.Lq) : (IdxT)0;
#pragma unroll
      for (int m = 1; m < kQuad; m <<= 1) below += __shfl_xor_sync(kFull, below, m);
      if (live) {
        nlook += 2;
        nlines += (((uint64_t)lo0 >> kBlockShift) != ((uint64_t)(IdxT)(hi0 + 1) >> kBlockShift)) ? 2 : 1;
        const IdxT dd = k.Ud - k.Ld;
        const uint32_t olen = l - i;
        if (olen >= min_ov && dd >= 1) {
          if (n < bt.find_cap && q == 0) {
            if (LEAN) {
              LeanCand lc;
              lc.lo = (uint32_t)(pc0 + k.Ld);
              lc.hi = (uint32_t)(pc0 + k.Ud - 1);
              lc.len_af = olen | (af << 16);
              lean[n] = lc;
            } else {
              quad_store_block<IdxT>(cand + n, pc0 + k.Ld, pc0 + k.Ud - 1, lo1, lo1 + dd - 1, lo0, hi0, lo1, hi1, olen, af);
            }
          }
          ++n;
        }
        below += dd;
        lo1 = lo1 + below;
        hi1 = lo1 + (Uc - Lc) - 1;
        // a chain of selects on registers: the nested form compiles to a branch tree, which diverges whenever the quads of
        // a warp step different symbols (a register-indexed constant load instead was slower, profiles/r1h_summary.md)
        IdxT pc = pc1;
        pc = c == 2 ? pc2 : pc;
        pc = c == 3 ? pc3 : pc;
        pc = c == 4 ? pc4 : pc;
        lo0 = pc + Lc;
        hi0 = pc + Uc - 1;
        --i;
        // an empty interval stays empty and emits nothing (see ov_find): stop stepping
        live = i > 0 && hi0 != kNone && hi0 >= lo0;
      }
    }
    // ---- containment handling (src/overlap_builder.cpp:877-905), again warp-uniform
    const bool tail = have && hi0 != kNone && hi0 >= lo0;  // only reachable with i == 0: an emptied walk has hi0 < lo0
    uint32_t ncontain = 0;
    if (__any_sync(kFull, tail)) {
      const QuadRank<IdxT> kp = quad_rank<IdxT>(Pb, tail ? lo0 : (IdxT)0, tail ? (IdxT)(hi0 + 1) : (IdxT)0, q);
      const QuadRank<IdxT> ks = quad_rank<IdxT>(Sb, tail ? lo1 : (IdxT)0, tail ? (IdxT)(hi1 + 1) : (IdxT)0, q);
      // any non-'$' extension on either side?  (ballot over the warp, then this quad's four bits)
      const bool mydna = (kp.Uq - kp.Lq) > 0 || (ks.Uq - ks.Lq) > 0;
      const bool dna = ((__ballot_sync(kFull, mydna) >> qbase) & 0xfu) != 0;
      const IdxT d = kp.Ud - kp.Ld;
      const IdxT p0lo = pc0 + kp.Ld;
      IdxT p0hi = pc0 + kp.Ud - 1;
      IdxT p1lo = lo1, p1hi = lo1 + d - 1;
      const bool v0 = p0hi != kNone && p0hi >= p0lo, v1 = p1hi != kNone && p1hi >= p1lo;
      const bool want = tail && !dna && v0 && v1;
      const bool need3 = want && p1hi != hi1;
      IdxT U3d = ks.Ud;
      if (__any_sync(kFull, need3)) {
        const QuadRank<IdxT> k3 = quad_rank<IdxT>(Sb, need3 ? lo1 : (IdxT)0, need3 ? (IdxT)(p1hi + 1) : (IdxT)0, q);
        if (need3) {
          U3d = k3.Ud;
          nlook += 1;
          nlines += 1;
        }
      }
      if (tail) {
        nlook += 4;
        nlines += 4;
        if (dna) {
          if (q == 0) bt.substring[r] = 1;
        } else if (want) {
          const IdxT d3 = U3d - ks.Ld;
          p0hi = p0lo + d3 - 1;
          p1lo = (IdxT)Spred[0] + ks.Ld;
          p1hi = (IdxT)Spred[0] + U3d - 1;
          if (q == 0) {
            if (LEAN) {
              LeanCand lc;
              lc.lo = (uint32_t)p0lo;
              lc.hi = (uint32_t)p0hi;
              lc.len_af = l | (af << 16);
              bt.lean_contain[(size_t)wi * 4 + mode] = lc;
            } else {
              quad_store_block<IdxT>(bt.contain + (size_t)wi * 4 + mode, p0lo, p0hi, p1lo, p1hi, lo0, hi0, lo1, hi1, l, af);
            }
          }
          ncontain = 1;
        }
      }
    }
    if (have && q == 0) {
      bt.ncand[wi * 4 + mode] = n < bt.find_cap ? n : bt.find_cap;
      bt.ncontain[wi * 4 + mode] = ncontain;
    }
  }
#undef Ppred
#undef Spred
  // the modes a run without reverse complement skips still need their counters cleared
  if (!rc) {
    for (uint64_t r = quad; r < count; r += nquads) {  // item index
      if (q == 0) {
        bt.ncand[r * 4 + 1] = 0;
        bt.ncand[r * 4 + 3] = 0;
        bt.ncontain[r * 4 + 1] = 0;
        bt.ncontain[r * 4 + 3] = 0;
      }
    }
  }
  if (q == 0 && nlook) {
    atomicAdd(lookups, nlook);
    atomicAdd(lookups + 2, nlines);
  }
}


// ---- lean search (sgx_overlap_edges): the walk the edge pipeline needs, at a third of the instructions --------------
//
// The edge pipeline consumes capped[0] of the '$' blocks only (LeanCand).  Per backward step that takes
//   * the rank of ONE symbol at both interval ends (the next interval): quad_rank_sym's arithmetic -- one popcount per
//     uint4, both ends packed in one register, two xor-shuffles, headers from lane c-1;
//   * once the overlap length has reached min_overlap, the number of '$' rows inside the interval: the in-block '$'
//     counts at both ends and the header difference of the two blocks packed into ONE register (9 + 9 + 14 bits; a
//     quad whose ends lie more than 63 blocks apart takes the full rank query instead), two more shuffles.  Steps
//     shorter than min_overlap skip the '$' probe altogether (30 % of the steps at L = 150, m = 45);
//   * '$' ranks themselves (the block headers summed over the quad) only on the steps that emit a block.
// The other-side interval (IntervalPair's second half, src/overlap_builder.cpp:70-136) is not carried through the walk:
// only the containment tail reads it, for the ~2 % of (read, mode) pairs whose full-length interval is non-empty, and
// there it is recomputed by a plain backward search of the reversed string in the opposite index (the bidirectional
// invariant: it IS the interval of reverse(X) there).
// UNIFORM (every indexed read has the same length, and the queries are the indexed reads): a full-length match is an
// identical read, every row of its interval holds '$' in both indices, so the containment block is the '$' rank of the
// interval's first row and its width -- one rank-block fetch, no walk through the opposite index.  (Modes 0 and 2 always
// end with a full-length match: the read finds itself.)
// Modes 0 and 2 search a read in the index that holds it, so their interval never empties: once it has narrowed to ONE row
// that row is the read's own suffix, no other read shares it, every further '$' probe finds nothing and the full-length
// interval is the read's own row of the '$' block, i.e. its position in the .sai order.  The walk stops there (a read with a
// sequencing error walks on average half its length instead of all of it) and the containment block comes from the inverse
// .sai without a fetch.
// Work order: all (read, mode) pairs that search the forward index first, then those of the reverse index; the quads of
// the grid advance together, so at any time one index (232 MB at config 2) competes for the 126 MB L2, not both.
// The symbol stream is read one 32-bit word per 16 steps.
template <class IdxT, int MINB, bool UNIFORM>
__global__ void __launch_bounds__(256, MINB)
ov_search_lean_kernel(FmPair fm, ReadsView reads, uint64_t first, uint64_t count, uint32_t min_ov, int rc,
                      OvBatch bt, const uint32_t* __restrict__ list, unsigned long long* __restrict__ lookups, int only_pi) {
  const unsigned kFull = 0xffffffffu;
  const int q = threadIdx.x & (kQuad - 1);
  const int qbase = threadIdx.x & 31 & ~(kQuad - 1);
  const int nmodes = rc ? 4 : 2;
  const uint64_t nwork = count * nmodes;
  unsigned long long nlook = 0, nlines = 0;
  const IdxT kNone = (IdxT)-1;
  const bool wide = sizeof(IdxT) == 8;

  // only_pi >= 0: this launch handles the searches of one index only (the launches of both indices cover the batch)
  const uint64_t wk_begin = only_pi > 0 ? nwork / 2 : 0, wk_end = only_pi == 0 ? nwork / 2 : nwork;
  for (uint64_t wk = wk_begin + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kQuad; __any_sync(kFull, wk < wk_end);
       wk += ((uint64_t)gridDim.x * blockDim.x) / kQuad) {
    const bool have = wk < wk_end;
    // index-major order: [0, nwork/2) searches the forward index (modes 0, 1), the rest the reverse index (modes 2, 3)
    const int pi = (have && wk >= nwork / 2) ? 1 : 0;
    const uint64_t within = have ? wk - (pi ? nwork / 2 : 0) : 0;
    const uint64_t wi = rc ? within >> 1 : within;  // where the results are stored
    const int mode = 2 * pi + (rc ? (int)(within & 1) : 0);
    const FmBlockLane* __restrict__ Pb = fm.blocks[pi];
    const uint64_t rd = first + (list ? (uint64_t)list[wi] : wi);  // which read
    const uint32_t* rw = reads.words + reads.woff[rd];
    const uint32_t l = reads.len[rd];
    uint32_t n = 0;
    const bool backwards = mode == 1 || mode == 2;  // string position k reads base l-1-k
    const uint32_t cm = (mode & 1) ? 0xffffffffu : 0u;  // modes 1 and 3 complement: flip both code bits
    // C[] of the searched index: lane q keeps C[q+1], the symbol whose block header it loads, and adds it to the
    // header before the header is handed to the quad (C['$'] = 0)
    const IdxT mypc = (IdxT)fm.pred[pi][q + 1];
    int c = ov_sym(rw, l, mode, l - 1);
    IdxT lo0 = (IdxT)fm.pred[pi][c], hi0 = (IdxT)(fm.pred[pi][c + 1] - 1);
    uint32_t j = backwards ? 1u : l - 2;  // base index of string position l-2; moves by dj as the position decreases
    const uint32_t dj = backwards ? 1u : 0xffffffffu;
    uint32_t i = have ? l - 1 : 0;  // remaining steps
    bool live = have && i > 0 && hi0 != kNone && hi0 >= lo0;
    uint32_t word = live ? (rw[j >> 4] ^ cm) : 0u;  // (complemented) codes of the 16 bases around j
    int cnext = (int)((word >> (2 * (j & 15))) & 3u) + 1;
    uint32_t lines = 0, steps = 0;
    const bool selfmode = (mode & 1) == 0;  // modes 0 and 2: the read is in the searched index
    bool selfonly = false;
#pragma unroll 1
    while (__any_sync(kFull, live)) {
      c = cnext;
      j += dj;
      // next step's symbol: a new word every 16 bases (entered from below when j grows, from above when it falls)
      if (live && i > 1 && (j & 15u) == (backwards ? 0u : 15u)) word = rw[j >> 4] ^ cm;
      cnext = (int)((word >> (2 * (j & 15))) & 3u) + 1;
      const IdxT plo = live ? lo0 : (IdxT)0, phi = live ? (IdxT)(hi0 + 1) : (IdxT)0;
      const uint64_t b1 = (uint64_t)plo >> kBlockShift, b2 = (uint64_t)phi >> kBlockShift;
      const int r1 = (int)(plo & (kBlockSyms - 1)), r2 = (int)(phi & (kBlockSyms - 1));
      const uint4* pa = reinterpret_cast<const uint4*>(Pb + b1 * kGroup + q);
      const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 4);
      uint4 c0 = a0, c1 = a1;
      if (b2 != b1) {
        const uint4* pb = reinterpret_cast<const uint4*>(Pb + b2 * kGroup + q);
        c0 = __ldg(pb);
        c1 = __ldg(pb + 4);
      }
      // prefix masks of this lane's two words at both ends (1 = at or past the prefix end)
      uint32_t ma0, ma1, mc0, mc1;
      asm("shl.b32 %0, %1, %2;" : "=r"(ma0) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * q, 0)));
      asm("shl.b32 %0, %1, %2;" : "=r"(ma1) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * (q + 4), 0)));
      asm("shl.b32 %0, %1, %2;" : "=r"(mc0) : "r"(0xffffffffu), "r"((uint32_t)max(r2 - 32 * q, 0)));
      asm("shl.b32 %0, %1, %2;" : "=r"(mc1) : "r"(0xffffffffu), "r"((uint32_t)max(r2 - 32 * (q + 4), 0)));
      const uint32_t x0 = ((c - 1) & 1) ? 0u : 0xffffffffu, x1 = ((c - 1) & 2) ? 0u : 0xffffffffu;
      uint32_t packed = (uint32_t)__popc(~(a0.z | ma0) & (a0.x ^ x0) & (a0.y ^ x1)) +
                        (uint32_t)__popc(~(a1.z | ma1) & (a1.x ^ x0) & (a1.y ^ x1)) +
                        (((uint32_t)__popc(~(c0.z | mc0) & (c0.x ^ x0) & (c0.y ^ x1)) +
                          (uint32_t)__popc(~(c1.z | mc1) & (c1.x ^ x0) & (c1.y ^ x1))) << 16);
      packed += __shfl_xor_sync(kFull, packed, 1);
      packed += __shfl_xor_sync(kFull, packed, 2);
      // C[c] + header of symbol c at both ends, from lane c-1 of the quad
      const int src = qbase + c - 1;
      IdxT nlo, nhi;
      if (wide) {
        const uint64_t ha = ((uint64_t)a1.w << 32 | a0.w) + mypc, hc = ((uint64_t)c1.w << 32 | c0.w) + mypc;
        nlo = (IdxT)__shfl_sync(kFull, (unsigned long long)ha, src);
        nhi = (IdxT)__shfl_sync(kFull, (unsigned long long)hc, src);
      } else {
        nlo = (IdxT)__shfl_sync(kFull, a0.w + (uint32_t)mypc, src);
        nhi = (IdxT)__shfl_sync(kFull, c0.w + (uint32_t)mypc, src);
      }
      nlo += (IdxT)(packed & 0xffffu);
      nhi += (IdxT)(packed >> 16);
      nhi -= 1;
      const uint32_t olen = l - i;
      const bool elig = live && olen >= min_ov;
      if (__any_sync(kFull, elig)) {
        // '$' rows inside [plo, phi): in-block counts at both ends + what the blocks in between hold
        const uint32_t nb = (uint32_t)min(b2 - b1, (uint64_t)64);
        uint32_t pk = ((uint32_t)__popc(a0.z & ~ma0) + (uint32_t)__popc(a1.z & ~ma1)) |
                      (((uint32_t)__popc(c0.z & ~mc0) + (uint32_t)__popc(c1.z & ~mc1)) << 9) | ((c0.w - a0.w) << 18);
        pk += __shfl_xor_sync(kFull, pk, 1);
        pk += __shfl_xor_sync(kFull, pk, 2);
        const uint32_t dl = pk & 511u;
        IdxT dd = (IdxT)(((pk >> 9) & 511u) - dl + 256u * nb - (pk >> 18));
        IdxT Ld = 0;
        const bool far = elig && nb > 63;  // the 14-bit header field does not hold this distance: full rank query
        if (__any_sync(kFull, far)) {
          const QuadRank<IdxT> k = quad_rank<IdxT>(Pb, far ? plo : (IdxT)0, far ? phi : (IdxT)0, q);
          if (far) {
            dd = k.Ud - k.Ld;
            Ld = k.Ld;
          }
        }
        const bool emit = elig && dd >= 1;
        if (__any_sync(kFull, emit)) {
          // '$' rank at the lower end = rows before it minus the A,C,G,T among them (block headers + in-block part)
          IdxT hs = wide ? (IdxT)((uint64_t)a1.w << 32 | a0.w) : (IdxT)a0.w;
          hs += __shfl_xor_sync(kFull, hs, 1);
          hs += __shfl_xor_sync(kFull, hs, 2);
          if (!far) Ld = (IdxT)(b1 << kBlockShift) - hs + dl;
          if (emit) {
            if (n < bt.find_cap && q == 0) {
              LeanCand lc;  // C['$'] = 0: the '$' rank is the row
              lc.lo = (uint32_t)Ld;
              lc.hi = (uint32_t)(Ld + dd - 1);
              lc.len_af = olen | (ov_mode_af(mode) << 16);
              bt.lean[((size_t)wi * 4 + mode) * bt.find_cap + n] = lc;
            }
            ++n;
          }
        }
      }
      if (live) {
        lines += b2 != b1 ? 2u : 1u;
        lo0 = nlo;
        hi0 = nhi;
        --i;
        ++steps;
        live = i > 0 && hi0 != kNone && hi0 >= lo0;
        if (UNIFORM && selfmode && live && hi0 == lo0) {  // only the read itself is left
          live = false;
          selfonly = true;
        }
      }
    }
    nlook += 2ull * steps;
    // ---- containment handling (src/overlap_builder.cpp:877-905), warp-uniform
    const bool tail = have && !selfonly && hi0 != kNone && hi0 >= lo0;  // only reachable with i == 0
    uint32_t ncontain = 0;
    if (UNIFORM) {
      if (selfonly) {
        if (q == 0) {
          LeanCand lc;
          lc.lo = lc.hi = fm.inv[pi][rd];
          lc.len_af = l | (ov_mode_af(mode) << 16);
          bt.lean_contain[(size_t)wi * 4 + mode] = lc;
        }
        ncontain = 1;
      }
      if (__any_sync(kFull, tail)) {
        const IdxT plo = tail ? lo0 : (IdxT)0;
        const uint64_t b1 = (uint64_t)plo >> kBlockShift;
        const int r1 = (int)(plo & (kBlockSyms - 1));
        const uint4* pa = reinterpret_cast<const uint4*>(Pb + b1 * kGroup + q);
        const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 4);
        uint32_t ma0, ma1;
        asm("shl.b32 %0, %1, %2;" : "=r"(ma0) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * q, 0)));
        asm("shl.b32 %0, %1, %2;" : "=r"(ma1) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * (q + 4), 0)));
        uint32_t dl = (uint32_t)__popc(a0.z & ~ma0) + (uint32_t)__popc(a1.z & ~ma1);
        IdxT hs = wide ? (IdxT)((uint64_t)a1.w << 32 | a0.w) : (IdxT)a0.w;
#pragma unroll
        for (int m = 1; m < kQuad; m <<= 1) {
          dl += __shfl_xor_sync(kFull, dl, m);
          hs += __shfl_xor_sync(kFull, hs, m);
        }
        if (tail) {
          nlook += 1;
          lines += 1;
          if (q == 0) {
            const IdxT Ld = (IdxT)(b1 << kBlockShift) - hs + dl;
            LeanCand lc;
            lc.lo = (uint32_t)Ld;
            lc.hi = (uint32_t)(Ld + (hi0 - lo0));
            lc.len_af = l | (ov_mode_af(mode) << 16);
            bt.lean_contain[(size_t)wi * 4 + mode] = lc;
          }
          ncontain = 1;
        }
      }
    } else if (__any_sync(kFull, tail)) {
      // the other-side interval: reverse(X) searched backwards in the opposite index = X read front to back
      const FmBlockLane* __restrict__ Sb = fm.blocks[1 - pi];
      int cs = tail ? ov_sym(rw, l, mode, 0) : 1;
      IdxT lo1 = (IdxT)fm.pred[1 - pi][cs], hi1 = (IdxT)(fm.pred[1 - pi][cs + 1] - 1);
      uint32_t t = 1;
      bool tl = tail && t < l;
#pragma unroll 1
      while (__any_sync(kFull, tl)) {
        cs = tl ? ov_sym(rw, l, mode, t) : 1;
        IdxT Ls, Us;
        quad_rank_sym<IdxT>(Sb, tl ? lo1 : (IdxT)0, tl ? (IdxT)(hi1 + 1) : (IdxT)0, q, qbase, cs, Ls, Us);
        if (tl) {
          nlook += 2;
          lines += (((uint64_t)lo1 >> kBlockShift) != ((uint64_t)(IdxT)(hi1 + 1) >> kBlockShift)) ? 2u : 1u;
          lo1 = (IdxT)fm.pred[1 - pi][cs] + Ls;
          hi1 = (IdxT)fm.pred[1 - pi][cs] + Us - 1;
          ++t;
          tl = t < l && hi1 != kNone && hi1 >= lo1;
        }
      }
      const QuadRank<IdxT> kp = quad_rank<IdxT>(Pb, tail ? lo0 : (IdxT)0, tail ? (IdxT)(hi0 + 1) : (IdxT)0, q);
      const QuadRank<IdxT> ks = quad_rank<IdxT>(Sb, tail ? lo1 : (IdxT)0, tail ? (IdxT)(hi1 + 1) : (IdxT)0, q);
      const bool mydna = (kp.Uq - kp.Lq) > 0 || (ks.Uq - ks.Lq) > 0;
      const bool dna = ((__ballot_sync(kFull, mydna) >> qbase) & 0xfu) != 0;
      const IdxT d = kp.Ud - kp.Ld;
      const IdxT p0lo = kp.Ld;  // + C['$'] = 0
      IdxT p0hi = kp.Ud - 1;
      const IdxT p1hi = lo1 + d - 1;
      const bool v0 = p0hi != kNone && p0hi >= p0lo, v1 = p1hi != kNone && p1hi >= lo1;
      const bool want = tail && !dna && v0 && v1;
      const bool need3 = want && p1hi != hi1;
      IdxT U3d = ks.Ud;
      if (__any_sync(kFull, need3)) {
        const QuadRank<IdxT> k3 = quad_rank<IdxT>(Sb, need3 ? lo1 : (IdxT)0, need3 ? (IdxT)(p1hi + 1) : (IdxT)0, q);
        if (need3) {
          U3d = k3.Ud;
          nlook += 1;
          lines += 1;
        }
      }
      if (tail) {
        nlook += 4;
        lines += 4;
        if (dna) {
          if (q == 0) bt.substring[list ? (uint64_t)list[wi] : wi] = 1;
        } else if (want) {
          const IdxT d3 = U3d - ks.Ld;
          p0hi = p0lo + d3 - 1;
          if (q == 0) {
            LeanCand lc;
            lc.lo = (uint32_t)p0lo;
            lc.hi = (uint32_t)p0hi;
            lc.len_af = l | (ov_mode_af(mode) << 16);
            bt.lean_contain[(size_t)wi * 4 + mode] = lc;
          }
          ncontain = 1;
        }
      }
    }
    nlines += lines;
    if (have && q == 0) {
      bt.ncand[wi * 4 + mode] = n < bt.find_cap ? n : bt.find_cap;
      bt.ncontain[wi * 4 + mode] = ncontain;
    }
  }
  if (!rc && only_pi <= 0) {
    const uint64_t quad = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kQuad;
    const uint64_t nquads = ((uint64_t)gridDim.x * blockDim.x) / kQuad;
    for (uint64_t r = quad; r < count; r += nquads) {
      if (q == 0) {
        bt.ncand[r * 4 + 1] = 0;
        bt.ncand[r * 4 + 3] = 0;
        bt.ncontain[r * 4 + 1] = 0;
        bt.ncontain[r * 4 + 3] = 0;
      }
    }
  }
  if (q == 0) {
    if (nlook) atomicAdd(lookups, nlook);
    if (nlines) atomicAdd(lookups + 2, nlines);
  }
}


// ---- lean search, persistent quads (reads of one length) -------------------------------------------------------------
//
// ov_search_lean_kernel walks the eight chains of a warp in lock-step until the LAST of them has finished; with the early
// exits above a chain takes anything between a few and L-1 steps (77 on average at config 2), so half of the lanes of a
// warp step were idle riders.  Here a quad that has finished its (read, mode) pair writes its result and takes its next
// pair at once, inside the same loop: the warp executes the finalize-and-fetch block when any of its quads needs it
// (~150 instructions per pair, against ~100 per step), and every step it executes carries eight live chains.
// Same arithmetic and outputs as ov_search_lean_kernel<.., UNIFORM = true>; pairs are handed out in index-major order
// (quad g takes pairs g, g + #quads, ...).
template <class IdxT, int MINB>
__global__ void __launch_bounds__(256, MINB)
ov_search_pq_kernel(FmPair fm, ReadsView reads, uint32_t ulen, uint32_t uwords, uint64_t first, uint64_t count, uint32_t min_ov,
                    int rc, OvBatch bt, const uint32_t* __restrict__ list, unsigned long long* __restrict__ lookups, int only_pi) {
  const unsigned kFull = 0xffffffffu;
  const int q = threadIdx.x & (kQuad - 1);
  const int qbase = threadIdx.x & 31 & ~(kQuad - 1);
  const int nmodes = rc ? 4 : 2;
  const uint64_t nwork = count * nmodes;
  const uint64_t wk_begin = only_pi > 0 ? nwork / 2 : 0, wk_end = only_pi == 0 ? nwork / 2 : nwork;
  const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) / kQuad;
  uint64_t wk = wk_begin + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kQuad;  // next pair of this quad
  unsigned long long nlook = 0, nlines = 0;
  const IdxT kNone = (IdxT)-1;
  const bool wide = sizeof(IdxT) == 8;
  const uint32_t l = ulen;

  // state of the pair in flight
  bool active = false, live = false, selfonly = false, selfmode = false, backwards = false;
  const FmBlockLane* __restrict__ Pb = fm.blocks[0];
  const uint32_t* rw = reads.words;
  IdxT mypc = 0, lo0 = 0, hi0 = 0;
  uint64_t slot = 0;  // wi * 4 + mode: where the results go
  uint32_t rd32 = 0, i = 0, j = 0, word = 0, cm = 0, n = 0, steps = 0, lines = 0;
  int cnext = 1, pi = 0;

  for (;;) {
    const bool need = !live;
    if (__any_sync(kFull, need)) {
      // ---- finish the pair that has just ended: containment block, counters
      const bool tail = need && active && !selfonly && hi0 != kNone && hi0 >= lo0;  // full-length match other than the read itself
      if (__any_sync(kFull, tail)) {
        const IdxT plo = tail ? lo0 : (IdxT)0;
        const uint64_t b1 = (uint64_t)plo >> kBlockShift;
        const int r1 = (int)(plo & (kBlockSyms - 1));
        const uint4* pa = reinterpret_cast<const uint4*>(Pb + b1 * kGroup + q);
        const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 4);
        uint32_t ma0, ma1;
        asm("shl.b32 %0, %1, %2;" : "=r"(ma0) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * q, 0)));
        asm("shl.b32 %0, %1, %2;" : "=r"(ma1) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * (q + 4), 0)));
        uint32_t dl = (uint32_t)__popc(a0.z & ~ma0) + (uint32_t)__popc(a1.z & ~ma1);
        IdxT hs = wide ? (IdxT)((uint64_t)a1.w << 32 | a0.w) : (IdxT)a0.w;
#pragma unroll
        for (int m = 1; m < kQuad; m <<= 1) {
          dl += __shfl_xor_sync(kFull, dl, m);
          hs += __shfl_xor_sync(kFull, hs, m);
        }
        if (tail) {
          nlook += 1;
          nlines += 1;
          if (q == 0) {
            const IdxT Ld = (IdxT)(b1 << kBlockShift) - hs + dl;
            LeanCand lc;
            lc.lo = (uint32_t)Ld;
            lc.hi = (uint32_t)(Ld + (hi0 - lo0));
            lc.len_af = l | (ov_mode_af((int)(slot & 3)) << 16);
            bt.lean_contain[slot] = lc;
          }
        }
      }
      if (need && active) {
        if (selfonly && q == 0) {
          LeanCand lc;
          lc.lo = lc.hi = fm.inv[pi][rd32];
          lc.len_af = l | (ov_mode_af((int)(slot & 3)) << 16);
          bt.lean_contain[slot] = lc;
        }
        if (q == 0) {
          bt.ncand[slot] = n < bt.find_cap ? n : bt.find_cap;
          bt.ncontain[slot] = (selfonly || tail) ? 1u : 0u;
        }
        nlook += 2ull * steps;
        nlines += lines;
        active = false;
      }
      // ---- take the next pair
      if (need && wk < wk_end) {
        active = true;
        pi = wk >= nwork / 2 ? 1 : 0;  // index-major order (see ov_search_lean_kernel)
        const uint64_t within = wk - (pi ? nwork / 2 : 0);
        const uint64_t wi = rc ? within >> 1 : within;
        const int mode = 2 * pi + (rc ? (int)(within & 1) : 0);
        wk += stride;
        slot = wi * 4 + mode;
        Pb = fm.blocks[pi];
        mypc = (IdxT)fm.pred[pi][q + 1];
        const uint64_t rd = first + (list ? (uint64_t)list[wi] : wi);
        rd32 = (uint32_t)rd;
        rw = reads.words + rd * uwords;
        backwards = mode == 1 || mode == 2;
        cm = (mode & 1) ? 0xffffffffu : 0u;
        selfmode = (mode & 1) == 0;
        selfonly = false;
        n = 0;
        steps = 0;
        lines = 0;
        const int c = ov_sym(rw, l, mode, l - 1);
        lo0 = (IdxT)fm.pred[pi][c];
        hi0 = (IdxT)(fm.pred[pi][c + 1] - 1);
        j = backwards ? 1u : l - 2;
        i = l - 1;
        live = i > 0 && hi0 != kNone && hi0 >= lo0;
        word = live ? (rw[j >> 4] ^ cm) : 0u;
        cnext = (int)((word >> (2 * (j & 15))) & 3u) + 1;
      }
      if (!__any_sync(kFull, active)) break;
      if (!__any_sync(kFull, live)) continue;  // pairs that ended before their first step are finished next round
    }
    // ---- one backward step of every live quad (the body of ov_search_lean_kernel)
    const int c = cnext;
    const uint32_t dj = backwards ? 1u : 0xffffffffu;
    j += dj;
    if (live && i > 1 && (j & 15u) == (backwards ? 0u : 15u)) word = rw[j >> 4] ^ cm;
    cnext = (int)((word >> (2 * (j & 15))) & 3u) + 1;
    const IdxT plo = live ? lo0 : (IdxT)0, phi = live ? (IdxT)(hi0 + 1) : (IdxT)0;
    const uint64_t b1 = (uint64_t)plo >> kBlockShift, b2 = (uint64_t)phi >> kBlockShift;
    const int r1 = (int)(plo & (kBlockSyms - 1)), r2 = (int)(phi & (kBlockSyms - 1));
    const uint4* pa = reinterpret_cast<const uint4*>(Pb + b1 * kGroup + q);
    const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 4);
    uint4 c0 = a0, c1 = a1;
    if (b2 != b1) {
      const uint4* pb = reinterpret_cast<const uint4*>(Pb + b2 * kGroup + q);
      c0 = __ldg(pb);
      c1 = __ldg(pb + 4);
    }
    uint32_t ma0, ma1, mc0, mc1;
    asm("shl.b32 %0, %1, %2;" : "=r"(ma0) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * q, 0)));
    asm("shl.b32 %0, %1, %2;" : "=r"(ma1) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * (q + 4), 0)));
    asm("shl.b32 %0, %1, %2;" : "=r"(mc0) : "r"(0xffffffffu), "r"((uint32_t)max(r2 - 32 * q, 0)));
    asm("shl.b32 %0, %1, %2;" : "=r"(mc1) : "r"(0xffffffffu), "r"((uint32_t)max(r2 - 32 * (q + 4), 0)));
    const uint32_t x0 = ((c - 1) & 1) ? 0u : 0xffffffffu, x1 = ((c - 1) & 2) ? 0u : 0xffffffffu;
    uint32_t packed = (uint32_t)__popc(~(a0.z | ma0) & (a0.x ^ x0) & (a0.y ^ x1)) +
                      (uint32_t)__popc(~(a1.z | ma1) & (a1.x ^ x0) & (a1.y ^ x1)) +
                      (((uint32_t)__popc(~(c0.z | mc0) & (c0.x ^ x0) & (c0.y ^ x1)) +
                        (uint32_t)__popc(~(c1.z | mc1) & (c1.x ^ x0) & (c1.y ^ x1))) << 16);
    packed += __shfl_xor_sync(kFull, packed, 1);
    packed += __shfl_xor_sync(kFull, packed, 2);
    const int src = qbase + c - 1;
    IdxT nlo, nhi;
    if (wide) {
      const uint64_t ha = ((uint64_t)a1.w << 32 | a0.w) + mypc, hc = ((uint64_t)c1.w << 32 | c0.w) + mypc;
      nlo = (IdxT)__shfl_sync(kFull, (unsigned long long)ha, src);
      nhi = (IdxT)__shfl_sync(kFull, (unsigned long long)hc, src);
    } else {
      nlo = (IdxT)__shfl_sync(kFull, a0.w + (uint32_t)mypc, src);
      nhi = (IdxT)__shfl_sync(kFull, c0.w + (uint32_t)mypc, src);
    }
    nlo += (IdxT)(packed & 0xffffu);
    nhi += (IdxT)(packed >> 16);
    nhi -= 1;
    const uint32_t olen = l - i;
    const bool elig = live && olen >= min_ov;
    if (__any_sync(kFull, elig)) {
      const uint32_t nb = (uint32_t)min(b2 - b1, (uint64_t)64);
      uint32_t pk = ((uint32_t)__popc(a0.z & ~ma0) + (uint32_t)__popc(a1.z & ~ma1)) |
                    (((uint32_t)__popc(c0.z & ~mc0) + (uint32_t)__popc(c1.z & ~mc1)) << 9) | ((c0.w - a0.w) << 18);
      pk += __shfl_xor_sync(kFull, pk, 1);
      pk += __shfl_xor_sync(kFull, pk, 2);
      const uint32_t dl = pk & 511u;
      IdxT dd = (IdxT)(((pk >> 9) & 511u) - dl + 256u * nb - (pk >> 18));
      IdxT Ld = 0;
      const bool far = elig && nb > 63;
      if (__any_sync(kFull, far)) {
        const QuadRank<IdxT> k = quad_rank<IdxT>(Pb, far ? plo : (IdxT)0, far ? phi : (IdxT)0, q);
        if (far) {
          dd = k.Ud - k.Ld;
          Ld = k.Ld;
        }
      }
      const bool emit = elig && dd >= 1;
      if (__any_sync(kFull, emit)) {
        IdxT hs = wide ? (IdxT)((uint64_t)a1.w << 32 | a0.w) : (IdxT)a0.w;
        hs += __shfl_xor_sync(kFull, hs, 1);
        hs += __shfl_xor_sync(kFull, hs, 2);
        if (!far) Ld = (IdxT)(b1 << kBlockShift) - hs + dl;
        if (emit) {
          if (n < bt.find_cap && q == 0) {
            LeanCand lc;
            lc.lo = (uint32_t)Ld;
            lc.hi = (uint32_t)(Ld + dd - 1);
            lc.len_af = olen | (ov_mode_af((int)(slot & 3)) << 16);
            bt.lean[slot * bt.find_cap + n] = lc;
          }
          ++n;
        }
      }
    }
    if (live) {
      lines += b2 != b1 ? 2u : 1u;
      lo0 = nlo;
      hi0 = nhi;
      --i;
      ++steps;
      live = i > 0 && hi0 != kNone && hi0 >= lo0;
      if (selfmode && live && hi0 == lo0) {  // only the read itself is left
        live = false;
        selfonly = true;
      }
    }
  }
  if (!rc && only_pi <= 0) {
    const uint64_t quad = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kQuad;
    for (uint64_t r = quad; r < count; r += stride) {
      if (q == 0) {
        bt.ncand[r * 4 + 1] = 0;
        bt.ncand[r * 4 + 3] = 0;
        bt.ncontain[r * 4 + 1] = 0;
        bt.ncontain[r * 4 + 3] = 0;
      }
    }
  }
  if (q == 0) {
    if (nlook) atomicAdd(lookups, nlook);
    if (nlines) atomicAdd(lookups + 2, nlines);
  }
}


// ---- persistent quads, lean registers ---------------------------------------------------------------------------------
//
// ov_search_pq_kernel is bound by the latency of its dependent rank-block fetches: one fetch per chain in flight, and the
// number of chains in flight is the number of resident warps, which its 64 registers per thread cap at 32 per SM.  Most of
// those registers hold state that changes once per (read, mode) pair, replicated over the four lanes of the quad.  Here
// that state (result slot, read id, candidate count, counters, the hand-out position) lives in shared memory, one record
// per quad, and what a step can derive is derived (symbol position from the step count, symbol from the cached word):
// the loop keeps the interval, the step count, the cached read word, the index base and C[] of the lane's symbol.
// Same outputs as ov_search_pq_kernel.
struct PqCold {
  unsigned long long wk;      // next pair of this quad
  unsigned long long slot;    // wi * 4 + mode of the pair in flight
  unsigned long long nlook;   // executed rank lookups (two per step)
  unsigned long long nlines;  // rank-block lines those lookups touched
  uint32_t rd;                // read of the pair in flight
  uint32_t n;                 // candidates emitted so far
  uint32_t extra;             // steps whose interval ends lay in different blocks (+ 1 for a start-table fetch)
  uint32_t i0;                // step count the pair started with
};

template <class IdxT, int MINB>
__global__ void __launch_bounds__(256, MINB)
ov_search_pq2_kernel(FmPair fm, ReadsView reads, uint32_t ulen, uint32_t uwords, uint64_t first, uint64_t count, uint32_t min_ov,
                     int rc, OvBatch bt, const uint32_t* __restrict__ list, unsigned long long* __restrict__ lookups, int only_pi,
                     int ktk /* > 0: start every pair from the k-mer table (the caller checked min_ov >= ktk and ulen > ktk) */) {
  __shared__ PqCold cold[256 / kQuad];
  const unsigned kFull = 0xffffffffu;
  const int q = threadIdx.x & (kQuad - 1);
  const int qbase = threadIdx.x & 31 & ~(kQuad - 1);
  PqCold& cs = cold[threadIdx.x / kQuad];
  const uint64_t nwork = count * (rc ? 4 : 2);
  const uint64_t wk_end = only_pi == 0 ? nwork / 2 : nwork;
  const IdxT kNone = (IdxT)-1;
  const bool wide = sizeof(IdxT) == 8;
  const uint32_t l = ulen;
  if (q == 0) {
    cs.wk = (only_pi > 0 ? nwork / 2 : 0) + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kQuad;
    cs.nlook = 0;
    cs.nlines = 0;
  }
  __syncwarp();

  // flags: bit 0 live, 1 active, 2 selfonly, 3 backwards (modes 1, 2), 4 complement (modes 1, 3), 5 index (modes 2, 3)
  uint32_t fl = 0, i = 0, word = 0;
  IdxT lo0 = 0, hi0 = 0, mypc = 0;
  const FmBlockLane* __restrict__ Pb = fm.blocks[0];

  for (;;) {
    if (__any_sync(kFull, !(fl & 1u))) {
      const bool need = !(fl & 1u);
      // ---- finish the pair that has just ended
      const bool tail = need && (fl & 2u) && !(fl & 4u) && hi0 != kNone && hi0 >= lo0;  // full-length match, not just the read itself
      if (__any_sync(kFull, tail)) {
        const IdxT plo = tail ? lo0 : (IdxT)0;
        const uint64_t b1 = (uint64_t)plo >> kBlockShift;
        const int r1 = (int)(plo & (kBlockSyms - 1));
        const uint4* pa = reinterpret_cast<const uint4*>(Pb + b1 * kGroup + q);
        const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 4);
        uint32_t ma0, ma1;
        asm("shl.b32 %0, %1, %2;" : "=r"(ma0) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * q, 0)));
        asm("shl.b32 %0, %1, %2;" : "=r"(ma1) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * (q + 4), 0)));
        uint32_t dl = (uint32_t)__popc(a0.z & ~ma0) + (uint32_t)__popc(a1.z & ~ma1);
        IdxT hs = wide ? (IdxT)((uint64_t)a1.w << 32 | a0.w) : (IdxT)a0.w;
#pragma unroll
        for (int m = 1; m < kQuad; m <<= 1) {
          dl += __shfl_xor_sync(kFull, dl, m);
          hs += __shfl_xor_sync(kFull, hs, m);
        }
        if (tail && q == 0) {
          const IdxT Ld = (IdxT)(b1 << kBlockShift) - hs + dl;
          LeanCand lc;
          lc.lo = (uint32_t)Ld;
          lc.hi = (uint32_t)(Ld + (hi0 - lo0));
          lc.len_af = l | (ov_mode_af((int)(cs.slot & 3)) << 16);
          bt.lean_contain[cs.slot] = lc;
          cs.nlook += 1;
          cs.nlines += 1;
        }
      }
      if (need && (fl & 2u) && q == 0) {
        const uint64_t slot = cs.slot;
        if (fl & 4u) {
          LeanCand lc;
          lc.lo = lc.hi = fm.inv[(fl >> 5) & 1u][cs.rd];
          lc.len_af = l | (ov_mode_af((int)(slot & 3)) << 16);
          bt.lean_contain[slot] = lc;
        }
        bt.ncand[slot] = cs.n < bt.find_cap ? cs.n : bt.find_cap;
        bt.ncontain[slot] = ((fl & 4u) || tail) ? 1u : 0u;
        const uint32_t steps = cs.i0 - i;
        cs.nlook += 2ull * steps;
        cs.nlines += (unsigned long long)steps + cs.extra;
      }
      if (need) fl &= ~2u;
      __syncwarp();
      // ---- take the next pair
      const uint64_t wk = cs.wk;
      __syncwarp();
      if (need && wk < wk_end) {
        const uint32_t pi = wk >= nwork / 2 ? 1u : 0u;  // index-major order (see ov_search_lean_kernel)
        const uint64_t within = wk - (pi ? nwork / 2 : 0);
        const uint64_t wi = rc ? within >> 1 : within;
        const int mode = 2 * (int)pi + (rc ? (int)(within & 1) : 0);
        const uint64_t rd = first + (list ? (uint64_t)list[wi] : wi);
        if (q == 0) {
          cs.wk = wk + ((uint64_t)gridDim.x * blockDim.x) / kQuad;
          cs.slot = wi * 4 + mode;
          cs.rd = (uint32_t)rd;
          cs.n = 0;
          cs.extra = 0;
        }
        Pb = fm.blocks[pi];
        mypc = (IdxT)fm.pred[pi][q + 1];
        const uint32_t* rw = reads.words + rd * uwords;
        fl = 2u | ((mode == 1 || mode == 2) ? 8u : 0u) | ((mode & 1) ? 16u : 0u) | (pi << 5);
        bool alone = false;
        if (ktk > 0) {
          // The first ktk - 1 steps of every chain walk the top of the index, which all chains share; their result depends
          // on the last ktk symbols only and is tabulated (sgx_fm::ktab): ONE fetch instead of ktk - 1 dependent ones.
          // No '$' probe is skipped, because probes start at overlap length min_ov >= ktk.
          uint32_t key = 0;
          for (int t = 0; t < ktk; ++t) key = key * 4u + (uint32_t)(ov_sym(rw, l, mode, l - 1 - (uint32_t)t) - 1);
          const uint2 e = __ldg(fm.ktab[pi] + key);
          lo0 = (IdxT)e.x;
          hi0 = (IdxT)(e.x + e.y - 1u);  // empty entries are {1, 0}
          i = l - (uint32_t)ktk;
          alone = !(fl & 16u) && e.y == 1u;  // modes 0 and 2: only the read itself is left (same rule as after a step)
          if (q == 0) cs.extra = 1;
        } else {
          const int c = ov_sym(rw, l, mode, l - 1);
          lo0 = (IdxT)fm.pred[pi][c];
          hi0 = (IdxT)(fm.pred[pi][c + 1] - 1);
          i = l - 1;
        }
        if (q == 0) cs.i0 = i;
        if (alone) {
          fl |= 4u;
        } else if (i > 0 && hi0 != kNone && hi0 >= lo0) {
          fl |= 1u;
          const uint32_t j = (fl & 8u) ? l - i : i - 1;  // base index of string position i - 1
          word = rw[j >> 4] ^ ((fl & 16u) ? 0xffffffffu : 0u);
        }
      }
      __syncwarp();
      if (!__any_sync(kFull, fl & 2u)) break;
      if (!__any_sync(kFull, fl & 1u)) continue;  // pairs that ended before their first step are finished next round
    }
    // ---- one backward step of every live quad
    const bool live = (fl & 1u) != 0;
    const uint32_t j = (fl & 8u) ? l - i : i - 1;  // base index of string position i - 1, the symbol of this step
    const int c = (int)((word >> (2 * (j & 15))) & 3u) + 1;
    // the symbol of the next step sits in the neighbouring word when this one is at the word's edge
    if (live && i > 1 && (j & 15u) == ((fl & 8u) ? 15u : 0u)) {
      const uint32_t jn = (fl & 8u) ? j + 1 : j - 1;
      word = reads.words[(uint64_t)cs.rd * uwords + (jn >> 4)] ^ ((fl & 16u) ? 0xffffffffu : 0u);
    }
    const IdxT plo = live ? lo0 : (IdxT)0, phi = live ? (IdxT)(hi0 + 1) : (IdxT)0;
    const uint64_t b1 = (uint64_t)plo >> kBlockShift, b2 = (uint64_t)phi >> kBlockShift;
    const int r1 = (int)(plo & (kBlockSyms - 1)), r2 = (int)(phi & (kBlockSyms - 1));
    const uint4* pa = reinterpret_cast<const uint4*>(Pb + b1 * kGroup + q);
    const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 4);
    uint4 c0 = a0, c1 = a1;
    if (b2 != b1) {
      const uint4* pb = reinterpret_cast<const uint4*>(Pb + b2 * kGroup + q);
      c0 = __ldg(pb);
      c1 = __ldg(pb + 4);
    }
    uint32_t ma0, ma1, mc0, mc1;
    asm("shl.b32 %0, %1, %2;" : "=r"(ma0) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * q, 0)));
    asm("shl.b32 %0, %1, %2;" : "=r"(ma1) : "r"(0xffffffffu), "r"((uint32_t)max(r1 - 32 * (q + 4), 0)));
    asm("shl.b32 %0, %1, %2;" : "=r"(mc0) : "r"(0xffffffffu), "r"((uint32_t)max(r2 - 32 * q, 0)));
    asm("shl.b32 %0, %1, %2;" : "=r"(mc1) : "r"(0xffffffffu), "r"((uint32_t)max(r2 - 32 * (q + 4), 0)));
    const uint32_t x0 = ((c - 1) & 1) ? 0u : 0xffffffffu, x1 = ((c - 1) & 2) ? 0u : 0xffffffffu;
    uint32_t packed = (uint32_t)__popc(~(a0.z | ma0) & (a0.x ^ x0) & (a0.y ^ x1)) +
                      (uint32_t)__popc(~(a1.z | ma1) & (a1.x ^ x0) & (a1.y ^ x1)) +
                      (((uint32_t)__popc(~(c0.z | mc0) & (c0.x ^ x0) & (c0.y ^ x1)) +
                        (uint32_t)__popc(~(c1.z | mc1) & (c1.x ^ x0) & (c1.y ^ x1))) << 16);
    packed += __shfl_xor_sync(kFull, packed, 1);
    packed += __shfl_xor_sync(kFull, packed, 2);
    const int src = qbase + c - 1;
    IdxT nlo, nhi;
    if (wide) {
      const uint64_t ha = ((uint64_t)a1.w << 32 | a0.w) + mypc, hc = ((uint64_t)c1.w << 32 | c0.w) + mypc;
      nlo = (IdxT)__shfl_sync(kFull, (unsigned long long)ha, src);
      nhi = (IdxT)__shfl_sync(kFull, (unsigned long long)hc, src);
    } else {
      nlo = (IdxT)__shfl_sync(kFull, a0.w + (uint32_t)mypc, src);
      nhi = (IdxT)__shfl_sync(kFull, c0.w + (uint32_t)mypc, src);
    }
    nlo += (IdxT)(packed & 0xffffu);
    nhi += (IdxT)(packed >> 16);
    nhi -= 1;
    const uint32_t olen = l - i;
    const bool elig = live && olen >= min_ov;
    if (__any_sync(kFull, elig)) {
      const uint32_t nb = (uint32_t)min(b2 - b1, (uint64_t)64);
      uint32_t pk = ((uint32_t)__popc(a0.z & ~ma0) + (uint32_t)__popc(a1.z & ~ma1)) |
                    (((uint32_t)__popc(c0.z & ~mc0) + (uint32_t)__popc(c1.z & ~mc1)) << 9) | ((c0.w - a0.w) << 18);
      pk += __shfl_xor_sync(kFull, pk, 1);
      pk += __shfl_xor_sync(kFull, pk, 2);
      const uint32_t dl = pk & 511u;
      IdxT dd = (IdxT)(((pk >> 9) & 511u) - dl + 256u * nb - (pk >> 18));
      IdxT Ld = 0;
      const bool far = elig && nb > 63;
      if (__any_sync(kFull, far)) {
        const QuadRank<IdxT> k = quad_rank<IdxT>(Pb, far ? plo : (IdxT)0, far ? phi : (IdxT)0, q);
        if (far) {
          dd = k.Ud - k.Ld;
          Ld = k.Ld;
        }
      }
      const bool emit = elig && dd >= 1;
      if (__any_sync(kFull, emit)) {
        IdxT hs = wide ? (IdxT)((uint64_t)a1.w << 32 | a0.w) : (IdxT)a0.w;
        hs += __shfl_xor_sync(kFull, hs, 1);
        hs += __shfl_xor_sync(kFull, hs, 2);
        if (!far) Ld = (IdxT)(b1 << kBlockShift) - hs + dl;
        if (emit && q == 0) {
          const uint32_t n = cs.n;
          const uint64_t slot = cs.slot;
          if (n < bt.find_cap) {
            LeanCand lc;
            lc.lo = (uint32_t)Ld;
            lc.hi = (uint32_t)(Ld + dd - 1);
            lc.len_af = olen | (ov_mode_af((int)(slot & 3)) << 16);
            bt.lean[slot * bt.find_cap + n] = lc;
          }
          cs.n = n + 1;
        }
      }
    }
    if (live) {
      if (b2 != b1 && q == 0) cs.extra += 1;
      lo0 = nlo;
      hi0 = nhi;
      --i;
      if (!(i > 0 && hi0 != kNone && hi0 >= lo0)) {
        fl &= ~1u;
      } else if (!(fl & 16u) && hi0 == lo0) {  // modes 0 and 2 (no complement): only the read itself is left
        fl = (fl & ~1u) | 4u;
      }
    }
  }
  __syncwarp();
  if (!rc && only_pi <= 0) {
    const uint64_t quad = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kQuad;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) / kQuad;
    for (uint64_t r = quad; r < count; r += stride) {
      if (q == 0) {
        bt.ncand[r * 4 + 1] = 0;
        bt.ncand[r * 4 + 3] = 0;
        bt.ncontain[r * 4 + 1] = 0;
        bt.ncontain[r * 4 + 3] = 0;
      }
    }
  }
  if (q == 0) {
    if (cs.nlook) atomicAdd(lookups, cs.nlook);
    if (cs.nlines) atomicAdd(lookups + 2, cs.nlines);
  }
}

}  // namespace sgx

#endif  // SGX_OVERLAP_SEARCH_CUH_
