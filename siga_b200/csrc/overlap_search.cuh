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
  unsigned long long nlook = 0;
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
        }
      }
      if (tail) {
        nlook += 4;
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
  if (q == 0 && nlook) atomicAdd(lookups, nlook);
}

}  // namespace sgx

#endif  // SGX_OVERLAP_SEARCH_CUH_
