// The 4-lane rank query shared by the overlap search (overlap_search.cuh) and the k-mer correction kernel (correct.cu).
//
// Lane q of a quad loads uint4 q and q+4 of a 128-byte rank block (fm_layout.h), i.e. a quarter of its symbols plus the
// low and high header word of symbol q+1, for BOTH ends of an interval; when both ends fall into the same block the
// second pair of loads is predicated off, so a narrow interval costs one 128-byte line per backward-search step.
#ifndef SGX_QUAD_RANK_CUH_
#define SGX_QUAD_RANK_CUH_

#include "fm_layout.h"

namespace sgx {

static const int kQuad = 4;

template <class IdxT>
struct QuadRank {
  IdxT Lq, Uq;      // rank of this lane's symbol (q+1) in [0, plo) and [0, phi)
  IdxT Ld, Ud;      // rank of '$' at both ends (same on all lanes)
};

// fm_lane_counts with the prefix mask built from a clamped PTX shift (shl.b32 yields 0 for shift amounts >= 32).
__device__ __forceinline__ void quad_lane_counts(uint32_t p0, uint32_t p1, uint32_t p2, int k, int r, uint32_t& w1, uint32_t& w2) {
  const int nb = r - 32 * k;
  uint32_t hi_mask;
  asm("shl.b32 %0, %1, %2;" : "=r"(hi_mask) : "r"(0xffffffffu), "r"((uint32_t)max(nb, 0)));
  const uint32_t nd = ~(p2 | hi_mask);
  const uint32_t hi = nd & p1, lo = nd & ~p1;
  w1 = (uint32_t)__popc(lo & ~p0) | ((uint32_t)__popc(lo & p0) << 16);
  w2 = (uint32_t)__popc(hi & ~p0) | ((uint32_t)__popc(hi & p0) << 16);
}

// Ranks of all symbols at prefix lengths plo <= phi, distributed over the quad.  Every lane of the warp calls this
// together (full-mask shuffles); the xor distances 1 and 2 keep the exchange inside each quad.
template <class IdxT>
__device__ __forceinline__ QuadRank<IdxT> quad_rank(const FmBlockLane* __restrict__ blocks, IdxT plo, IdxT phi, int q) {
  const unsigned qmask = 0xffffffffu;
  const uint64_t b1 = (uint64_t)plo >> kBlockShift, b2 = (uint64_t)phi >> kBlockShift;
  const int r1 = (int)(plo & (kBlockSyms - 1)), r2 = (int)(phi & (kBlockSyms - 1));
  const uint4* pa = reinterpret_cast<const uint4*>(blocks + b1 * kGroup + q);
  const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 4);
  uint4 c0 = a0, c1 = a1;
  if (b2 != b1) {
    const uint4* pb = reinterpret_cast<const uint4*>(blocks + b2 * kGroup + q);
    c0 = __ldg(pb);
    c1 = __ldg(pb + 4);
  }
  uint32_t x1, x2, y1, y2, la1, la2, lb1, lb2;
  quad_lane_counts(a0.x, a0.y, a0.z, q, r1, x1, x2);
  quad_lane_counts(a1.x, a1.y, a1.z, q + 4, r1, y1, y2);
  la1 = x1 + y1;
  la2 = x2 + y2;
  quad_lane_counts(c0.x, c0.y, c0.z, q, r2, x1, x2);
  quad_lane_counts(c1.x, c1.y, c1.z, q + 4, r2, y1, y2);
  lb1 = x1 + y1;
  lb2 = x2 + y2;
#pragma unroll
  for (int m = 1; m < kQuad; m <<= 1) {
    la1 += __shfl_xor_sync(qmask, la1, m);
    la2 += __shfl_xor_sync(qmask, la2, m);
    lb1 += __shfl_xor_sync(qmask, lb1, m);
    lb2 += __shfl_xor_sync(qmask, lb2, m);
  }
  // in-block count of this lane's symbol: q=0 A (la1 low), 1 C (la1 high), 2 G (la2 low), 3 T (la2 high)
  const uint32_t wa = (q & 2) ? la2 : la1, wb = (q & 2) ? lb2 : lb1;
  const uint32_t ia = (q & 1) ? (wa >> 16) : (wa & 0xffffu), ib = (q & 1) ? (wb >> 16) : (wb & 0xffffu);
  QuadRank<IdxT> out;
  if (sizeof(IdxT) == 8) {
    out.Lq = (IdxT)(((uint64_t)a1.w << 32 | a0.w) + ia);
    out.Uq = (IdxT)(((uint64_t)c1.w << 32 | c0.w) + ib);
  } else {
    out.Lq = (IdxT)(a0.w + ia);
    out.Uq = (IdxT)(c0.w + ib);
  }
  IdxT sl = out.Lq, su = out.Uq;
#pragma unroll
  for (int m = 1; m < kQuad; m <<= 1) {
    sl += __shfl_xor_sync(qmask, sl, m);
    su += __shfl_xor_sync(qmask, su, m);
  }
  out.Ld = plo - sl;
  out.Ud = phi - su;
  return out;
}

// Rank of ONE symbol c (1..4, the same on the four lanes of a quad) at prefix lengths plo <= phi: Lc = occurrences in
// [0, plo), Uc = in [0, phi).  Same loads as quad_rank, a third of its arithmetic: one popcount per uint4, both ends
// summed in one packed register (two xor-shuffles), the block headers of symbol c fetched from lane c-1 of the quad.
// This is all a plain backward search (FMIndex::Interval::update, src/fmindex.h:96-100) needs.
template <class IdxT>
__device__ __forceinline__ void quad_rank_sym(const FmBlockLane* __restrict__ blocks, IdxT plo, IdxT phi, int q, int qbase,
                                              int c, IdxT& Lc, IdxT& Uc) {
  const unsigned kFull = 0xffffffffu;
  const uint64_t b1 = (uint64_t)plo >> kBlockShift, b2 = (uint64_t)phi >> kBlockShift;
  const int r1 = (int)(plo & (kBlockSyms - 1)), r2 = (int)(phi & (kBlockSyms - 1));
  const uint4* pa = reinterpret_cast<const uint4*>(blocks + b1 * kGroup + q);
  const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 4);
  uint4 c0 = a0, c1 = a1;
  if (b2 != b1) {
    const uint4* pb = reinterpret_cast<const uint4*>(blocks + b2 * kGroup + q);
    c0 = __ldg(pb);
    c1 = __ldg(pb + 4);
  }
  const uint32_t x0 = ((c - 1) & 1) ? 0u : 0xffffffffu, x1 = ((c - 1) & 2) ? 0u : 0xffffffffu;
  uint32_t packed = fm_lane_sym_count(a0.x, a0.y, a0.z, q, r1, x0, x1) + fm_lane_sym_count(a1.x, a1.y, a1.z, q + 4, r1, x0, x1) +
                    ((fm_lane_sym_count(c0.x, c0.y, c0.z, q, r2, x0, x1) + fm_lane_sym_count(c1.x, c1.y, c1.z, q + 4, r2, x0, x1)) << 16);
  packed += __shfl_xor_sync(kFull, packed, 1);
  packed += __shfl_xor_sync(kFull, packed, 2);
  // header words of symbol c sit in lane c-1: .w of uint4 c-1 (low word) and of uint4 c+3 (high word)
  const int src = qbase + c - 1;
  const uint32_t hl1 = __shfl_sync(kFull, a0.w, src), hl2 = __shfl_sync(kFull, c0.w, src);
  if (sizeof(IdxT) == 8) {
    const uint32_t hh1 = __shfl_sync(kFull, a1.w, src), hh2 = __shfl_sync(kFull, c1.w, src);
    Lc = (IdxT)(((uint64_t)hh1 << 32 | hl1) + (packed & 0xffffu));
    Uc = (IdxT)(((uint64_t)hh2 << 32 | hl2) + (packed >> 16));
  } else {
    Lc = (IdxT)(hl1 + (packed & 0xffffu));
    Uc = (IdxT)(hl2 + (packed >> 16));
  }
}

}  // namespace sgx

#endif  // SGX_QUAD_RANK_CUH_
