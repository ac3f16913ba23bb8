// Rank (occ) structure of the device FM-index and the group-cooperative rank query.
//
// Replaces the reference's LargeMarker/SmallMarker + run walk (src/fmindex.cpp:17-112, 178-285; src/fmindex.h:18-55).
// Those markers are never written to disk (src/fmindex.cpp:347-351), so the in-memory layout is ours to choose:
//
//   one 128-byte block per 256 BWT symbols = 8 x uint4; "lane" k of a block holds
//     .x plane0 : low  bit of (rank-1) for symbols [32k, 32k+32) of the block      (A=00 C=01 G=10 T=11)
//     .y plane1 : high bit of (rank-1)
//     .z plane2 : 1 where the symbol is '$'
//     .w header word k of {A_lo, C_lo, G_lo, T_lo, A_hi, C_hi, G_hi, T_hi}: the 64-bit counts of A,C,G,T in all
//        symbols BEFORE this block ('$' before the block = 256*b - A - C - G - T); lanes q and q+4 together hold
//        the full count of symbol q+1, which is what the 4-lane search kernel exploits
//
// A rank query getOcc(pos) (all five symbols, BWT[0..pos] inclusive) is therefore ONE aligned 128-byte fetch: the
// 8 lanes of a group load one uint4 each (a single fully used cache line per group), popcount their own 32
// symbols under a prefix mask, and the partial counts are summed with three xor-shuffles.
#ifndef SGX_FM_LAYOUT_H_
#define SGX_FM_LAYOUT_H_

#include "sgx_common.h"

namespace sgx {

static const int kBlockSyms = 256;
static const int kBlockShift = 8;
static const int kGroup = 8;  // lanes cooperating on one rank query

struct FmBlockLane {
  uint32_t p0, p1, p2, h;
};

struct FmView {
  const FmBlockLane* blocks;  // (n >> 8) + 1 blocks x 8 lanes
  uint64_t n;                 // BWT length
  uint64_t pred[6];           // pred[c] = #symbols with rank < c ; pred[5] = n   (FMIndex::getPC, src/fmindex.h:129)
};

SGX_HD uint64_t fm_num_blocks(uint64_t n) { return (n >> kBlockShift) + 1; }

SGX_HD int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// Partial counts of lane k for the first r symbols of a block: returns A | C<<16 in w1, G | T<<16 in w2.
SGX_HD void fm_lane_counts(uint32_t p0, uint32_t p1, uint32_t p2, int k, int r, uint32_t& w1, uint32_t& w2) {
  int nb = r - 32 * k;
  uint32_t mask = nb >= 32 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << nb) - 1u));
  uint32_t nd = ~p2 & mask;
  uint32_t hi = nd & p1, lo = nd & ~p1;
  w1 = (uint32_t)popc32(lo & ~p0) | ((uint32_t)popc32(lo & p0) << 16);
  w2 = (uint32_t)popc32(hi & ~p0) | ((uint32_t)popc32(hi & p0) << 16);
}

// Single-symbol variant: how many of the first r symbols of a block that lane k holds are the symbol with code bits
// (b0, b1), given x0 = b0 ? 0 : ~0 and x1 = b1 ? 0 : ~0 (so that p ^ x is 1 exactly where the plane bit equals the code bit).
SGX_HD uint32_t fm_lane_sym_count(uint32_t p0, uint32_t p1, uint32_t p2, int k, int r, uint32_t x0, uint32_t x1) {
  const int nb = r - 32 * k;
  uint32_t hi_mask;
#ifdef __CUDA_ARCH__
  asm("shl.b32 %0, %1, %2;" : "=r"(hi_mask) : "r"(0xffffffffu), "r"((uint32_t)(nb > 0 ? nb : 0)));  // 0 for shifts >= 32
#else
  hi_mask = nb >= 32 ? 0u : (nb <= 0 ? 0xffffffffu : (0xffffffffu << nb));
#endif
  return (uint32_t)popc32(~(p2 | hi_mask) & (p0 ^ x0) & (p1 ^ x1));
}

// ---------------------------------------------------------------------------------------------------------
// Group policies.  All control flow that uses a group is "group-uniform": every lane of the group executes the
// same path with the same scalar state; only the rank query splits work across lanes.
// ---------------------------------------------------------------------------------------------------------

#ifdef __CUDACC__
struct DevGroup {
  unsigned mask;  // the 8 lanes of this group inside its warp
  int lane;       // 0..7
  __device__ __forceinline__ DevGroup() {
    int wl = threadIdx.x & 31;
    lane = wl & (kGroup - 1);
    mask = 0xffu << (wl & ~(kGroup - 1));
  }
  __device__ __forceinline__ bool leader() const { return lane == 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

// FMIndex::getOcc(i) (src/fmindex.cpp:320-323, MarkerFind::find 188-231) for all five symbols.
__device__ __forceinline__ void fm_occ(const FmView& fm, pos_t pos, uint64_t out[5], const DevGroup& g) {
  uint64_t p = pos + 1;  // symbols in the prefix; pos == -1 wraps to 0 exactly like the reference's ++i
  if (p == 0) {
    out[0] = out[1] = out[2] = out[3] = out[4] = 0;
    return;
  }
  uint64_t b = p >> kBlockShift;
  int r = (int)(p & (kBlockSyms - 1));
  const uint4 w = __ldg(reinterpret_cast<const uint4*>(fm.blocks + b * kGroup + g.lane));
  uint32_t w1, w2;
  fm_lane_counts(w.x, w.y, w.z, g.lane, r, w1, w2);
#pragma unroll
  for (int m = 1; m < kGroup; m <<= 1) {
    w1 += __shfl_xor_sync(g.mask, w1, m);
    w2 += __shfl_xor_sync(g.mask, w2, m);
  }
  const int base = (threadIdx.x & 31) & ~(kGroup - 1);
  uint32_t h[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) h[k] = __shfl_sync(g.mask, w.w, base + k);
  uint64_t a = ((uint64_t)h[4] << 32 | h[0]) + (w1 & 0xffffu);
  uint64_t c = ((uint64_t)h[5] << 32 | h[1]) + (w1 >> 16);
  uint64_t gg = ((uint64_t)h[6] << 32 | h[2]) + (w2 & 0xffffu);
  uint64_t t = ((uint64_t)h[7] << 32 | h[3]) + (w2 >> 16);
  out[1] = a;
  out[2] = c;
  out[3] = gg;
  out[4] = t;
  out[0] = p - a - c - gg - t;
}

// FMIndex::getChar(i) (src/fmindex.cpp:233-246): rank of BWT[pos].
__device__ __forceinline__ int fm_sym(const FmView& fm, pos_t pos, const DevGroup& g) {
  (void)g;
  const FmBlockLane* l = fm.blocks + (pos >> kBlockShift) * kGroup + ((pos & (kBlockSyms - 1)) >> 5);
  uint32_t bit = (uint32_t)(pos & 31);
  uint32_t p0 = __ldg(&l->p0), p1 = __ldg(&l->p1), p2 = __ldg(&l->p2);
  if ((p2 >> bit) & 1u) return kDollar;
  return 1 + (int)(((p0 >> bit) & 1u) | (((p1 >> bit) & 1u) << 1));
}
#endif  // __CUDACC__

// Single-lane stand-in used by tests/hostemu only (never by the product library).
struct HostGroup {
  bool leader() const { return true; }
  void sync() const {}
};

inline void fm_occ(const FmView& fm, pos_t pos, uint64_t out[5], const HostGroup&) {
  uint64_t p = pos + 1;
  if (p == 0) {
    out[0] = out[1] = out[2] = out[3] = out[4] = 0;
    return;
  }
  uint64_t b = p >> kBlockShift;
  int r = (int)(p & (kBlockSyms - 1));
  const FmBlockLane* blk = fm.blocks + b * kGroup;
  uint32_t w1 = 0, w2 = 0;
  for (int k = 0; k < kGroup; ++k) {
    uint32_t a1, a2;
    fm_lane_counts(blk[k].p0, blk[k].p1, blk[k].p2, k, r, a1, a2);
    w1 += a1;
    w2 += a2;
  }
  uint64_t a = ((uint64_t)blk[4].h << 32 | blk[0].h) + (w1 & 0xffffu);
  uint64_t c = ((uint64_t)blk[5].h << 32 | blk[1].h) + (w1 >> 16);
  uint64_t gg = ((uint64_t)blk[6].h << 32 | blk[2].h) + (w2 & 0xffffu);
  uint64_t t = ((uint64_t)blk[7].h << 32 | blk[3].h) + (w2 >> 16);
  out[1] = a;
  out[2] = c;
  out[3] = gg;
  out[4] = t;
  out[0] = p - a - c - gg - t;
}

inline int fm_sym(const FmView& fm, pos_t pos, const HostGroup&) {
  const FmBlockLane* l = fm.blocks + (pos >> kBlockShift) * kGroup + ((pos & (kBlockSyms - 1)) >> 5);
  uint32_t bit = (uint32_t)(pos & 31);
  if ((l->p2 >> bit) & 1u) return kDollar;
  return 1 + (int)(((l->p0 >> bit) & 1u) | (((l->p1 >> bit) & 1u) << 1));
}

}  // namespace sgx

#endif  // SGX_FM_LAYOUT_H_
