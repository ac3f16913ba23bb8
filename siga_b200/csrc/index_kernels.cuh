// Device kernels of the BWT / read-index construction, shared by the single-device build (index.cu) and the
// bucket-sharded multi-GPU build (index_sharded.cu).  See index.cu for the algorithm.
#ifndef SGX_INDEX_KERNELS_CUH_
#define SGX_INDEX_KERNELS_CUH_

#include "index.cuh"
#include "prims.cuh"

namespace sgx {

// ---- read access ------------------------------------------------------------------------------------------------

struct ReadAcc {
  ReadsView rv;
  int reversed;
  // 2-bit code (A0 C1 G2 T3) of column `col` of read `id` in the indexed orientation
  __device__ __forceinline__ uint32_t code(uint32_t id, uint32_t col, uint32_t l) const {
    const uint32_t j = reversed ? l - 1 - col : col;
    const uint32_t* w = rv.words + rv.woff[id];
    return (w[j >> 4] >> (2 * (j & 15))) & 3u;
  }
};

// ---- 1. read ordering ---------------------------------------------------------------------------------------------

static __global__ void iota_kernel(uint32_t* a, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (uint32_t)i;
}

// keys[k] = bases [col0, col0 + nbases) of read ids[k] in the indexed orientation, first base most significant,
// 2 bits per base, bases past the end of the read = 0.
static __global__ void __launch_bounds__(256) chunk_keys_kernel(ReadAcc acc, const uint32_t* __restrict__ ids, uint64_t n,
                                                               uint32_t col0, uint32_t nbases, uint64_t* __restrict__ keys) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t id = ids[k];
  const uint32_t l = acc.rv.len[id];
  const uint32_t* w = acc.rv.words + acc.rv.woff[id];
  uint64_t key = 0;
  uint32_t cached = 0xffffffffu, word = 0;
  for (uint32_t t = 0; t < nbases; ++t) {
    const uint32_t col = col0 + t;
    uint32_t code = 0;
    if (col < l) {
      const uint32_t j = acc.reversed ? l - 1 - col : col;
      if ((j >> 4) != cached) {
        cached = j >> 4;
        word = w[cached];
      }
      code = (word >> (2 * (j & 15))) & 3u;
    }
    key = (key << 2) | code;
  }
  keys[k] = key;
}
static __global__ void len_keys_kernel(ReadsView rv, const uint32_t* __restrict__ ids, uint64_t n, uint64_t* __restrict__ keys) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) keys[k] = rv.len[ids[k]];
}

// heads[k] = 1 when the k-th read in content order differs from its predecessor.
static __global__ void read_heads_kernel(ReadsView rv, const uint32_t* __restrict__ ids, uint64_t n, uint64_t* __restrict__ heads) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint64_t h = 1;
  if (k > 0) {
    const uint32_t a = ids[k - 1], b = ids[k];
    const uint32_t la = rv.len[a], lb = rv.len[b];
    if (la == lb) {
      const uint32_t* wa = rv.words + rv.woff[a];
      const uint32_t* wb = rv.words + rv.woff[b];
      const uint32_t nw = (la + 15) / 16;
      bool eq = true;
      for (uint32_t w = 0; w < nw && eq; ++w) eq = wa[w] == wb[w];
      if (eq) h = 0;
    }
  }
  heads[k] = h;
}

// rank[ids[k]] = (heads before k) + heads[k]  (dense, 1-based) given the exclusive scan of the head flags.
static __global__ void assign_rank_kernel(const uint32_t* __restrict__ ids, const uint64_t* __restrict__ excl, uint64_t n,
                                   uint64_t total, uint32_t* __restrict__ rank) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint64_t next = k + 1 < n ? excl[k + 1] : total;
  rank[ids[k]] = (uint32_t)next;  // == excl[k] + heads[k]
}

static __global__ void doubling_keys_kernel(const uint32_t* __restrict__ rank, uint64_t n, uint64_t h, int bits,
                                     uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t r2 = i + h < n ? rank[i + h] : 0;  // end of text is smallest
  keys[i] = ((uint64_t)rank[i] << bits) | r2;
  vals[i] = (uint32_t)i;
}

static __global__ void key_heads_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ heads) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  heads[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}

// Sentinel order of model B from the read-start order `sai`: '$' of read i sorts by the suffix that follows it,
// i.e. by the rank of read i+1; the last read's '$' (end of text) is smallest.
static __global__ void find_zero_kernel(const uint32_t* __restrict__ sai, uint64_t n, uint64_t* __restrict__ k0) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && sai[k] == 0) *k0 = k;
}
static __global__ void sentinel_order_kernel(const uint32_t* __restrict__ sai, uint64_t n, const uint64_t* __restrict__ k0p,
                                      uint32_t* __restrict__ sord) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint64_t k0 = *k0p;
  if (k == 0) sord[0] = (uint32_t)(n - 1);
  if (k == k0) return;
  sord[1 + k - (k > k0 ? 1 : 0)] = sai[k] - 1;
}

// ---- 2. column passes ---------------------------------------------------------------------------------------------

static const int kBcrBlock = 256;    // rows per rank checkpoint
static const int kMergeTile = 8192;  // output rows per merge CTA
static const int kMergeThreads = 256;
// The partial BWT is stored 4 bits per symbol (ranks 0..4), two rows per byte, row k in bits 4*(k&1) of byte k>>1.

__device__ __forceinline__ uint32_t nib_get(const uint8_t* __restrict__ B, uint64_t k) {
  return (B[k >> 1] >> (4 * (k & 1))) & 15u;
}

// One thread per byte (two rows) of the first BWT: the N sentinel rows in sentinel order; the symbol of a
// sentinel row is the last base of its read.
template <class PosT>
__global__ void bcr_init_kernel(ReadAcc acc, const uint32_t* __restrict__ sord, uint64_t n, uint8_t* __restrict__ B,
                                uint32_t* __restrict__ ID, PosT* __restrict__ NP, uint8_t* __restrict__ SYM) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t k = 2 * i;
  if (k >= n) return;
  uint32_t byte = 0;
  for (int h = 0; h < 2; ++h) {
    if (k + h < n) {
      const uint32_t id = sord[k + h];
      const uint32_t l = acc.rv.len[id];
      const uint32_t sym = acc.code(id, l - 1, l) + 1;
      byte |= sym << (4 * h);
      ID[k + h] = id;
      NP[k + h] = (PosT)(k + h);
      SYM[k + h] = (uint8_t)sym;  // the symbol of a row is also the first base of the suffix inserted from it
    }
  }
  B[i] = (uint8_t)byte;
}

// A | C<<16 and G | T<<16 counts of the nibbles of a word (8 rows) selected by `keep` (a mask of nibble LSBs).
__device__ __forceinline__ void nib_count_add(uint32_t w, uint32_t keep, uint32_t& accA, uint32_t& accC, uint32_t& accG,
                                              uint32_t& accT) {
  const uint32_t m1 = w & 0x11111111u & keep, m2 = (w >> 1) & 0x11111111u & keep, m4 = (w >> 2) & 0x11111111u & keep;
  accT += m4;
  accG += m1 & m2;
  accA += m1 & ~m2;
  accC += m2 & ~m1;
}
// sum of the eight 4-bit fields of an accumulator (each field <= 15)
__device__ __forceinline__ uint32_t nib_sum(uint32_t acc) {
  const uint32_t v = (acc & 0x0f0f0f0fu) + ((acc >> 4) & 0x0f0f0f0fu);
  return (v * 0x01010101u) >> 24;
}

// Rank checkpoints of a nibble-packed symbol array, two levels:
//   blk[b]             per 256-row block: A | C<<16 , G | T<<16 counts of the rows of its 8192-row tile BEFORE the block
//   tiletot[c*nt + t]  per tile: number of symbol c in tile t (A,C,G,T = c 0..3); the caller's exclusive scan of the
//                      flattened [4][nt] array turns it into (#symbols < c) + (#c before tile t)
// so that rank_c(row) = scanned[c*nt + row/8192] + field_c(blk[row/256]) + (#c in the block before the row), and only
// nt = rows/8192 values per symbol are scanned per column pass.
// The single-device build does not even need blk: the kernel that writes a tile (the merge, or tile_count_kernel for the
// first array) has all of its rows at hand and stores, for every item the tile holds, the item's TILE-LOCAL rank
//   rloc[k] = number of rows of the tile before the item's row that hold the item's symbol
// so that the next pass computes rank_c(row) = scanned[c*nt + row/8192] + rloc[k] without touching the BWT at all.
// Called by all 256 threads of a CTA that holds one tile, thread i owning rows [32 i, 32 i + 32) as four words whose
// nibbles past the end of the array are zero.
__device__ __forceinline__ void tile_checkpoints(const uint32_t out[4], uint64_t tile, uint64_t ntiles, uint32_t* sh_w /* 64 */,
                                                 uint2* __restrict__ blk, uint64_t* __restrict__ tiletot, uint32_t& pre1,
                                                 uint32_t& pre2, uint2* inblock_sm = nullptr /* optional, shared, 256 entries */) {
  // bit planes of the 32 nibbles packed into one word each (word i's plane shifted to bit i of every nibble): symbol
  // ranks are A=1 C=2 G=3 T=4, so A = b0 & ~b1, C = b1 & ~b0, G = b0 & b1, T = b2 -- four popcounts per thread
  const uint32_t M = 0x11111111u;
  const uint32_t p0 = (out[0] & M) | ((out[1] << 1) & (M << 1)) | ((out[2] << 2) & (M << 2)) | ((out[3] << 3) & (M << 3));
  const uint32_t p1 = ((out[0] >> 1) & M) | (out[1] & (M << 1)) | ((out[2] << 1) & (M << 2)) | ((out[3] << 2) & (M << 3));
  const uint32_t p2 = ((out[0] >> 2) & M) | ((out[1] >> 1) & (M << 1)) | (out[2] & (M << 2)) | ((out[3] << 1) & (M << 3));
  const uint32_t w1 = (uint32_t)__popc(p0 & ~p1) | ((uint32_t)__popc(p1 & ~p0) << 16);
  const uint32_t w2 = (uint32_t)__popc(p0 & p1) | ((uint32_t)__popc(p2) << 16);
  // inclusive prefix inside the 8 consecutive threads of a 256-row block
  uint32_t i1 = w1, i2 = w2;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const uint32_t y1 = __shfl_up_sync(0xffffffffu, i1, o, 8), y2 = __shfl_up_sync(0xffffffffu, i2, o, 8);
    if ((int)(threadIdx.x & 7) >= o) {
      i1 += y1;
      i2 += y2;
    }
  }
  if ((threadIdx.x & 7) == 7) {
    sh_w[threadIdx.x >> 3] = i1;
    sh_w[32 + (threadIdx.x >> 3)] = i2;
  }
  // counts of the thread's 256-row block before its 32 rows: with sh_w (after the second barrier) any thread can
  // rebuild the tile prefix of any other thread -- pre of thread i = sh_w[i >> 3] + inblock_sm[i]
  if (inblock_sm) inblock_sm[threadIdx.x] = make_uint2(i1 - w1, i2 - w2);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const uint32_t v1 = sh_w[lane], v2 = sh_w[32 + lane];
    uint32_t x1 = v1, x2 = v2;  // packed 16-bit fields: a tile holds at most 8192 of one symbol
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y1 = __shfl_up_sync(0xffffffffu, x1, o), y2 = __shfl_up_sync(0xffffffffu, x2, o);
      if (lane >= o) {
        x1 += y1;
        x2 += y2;
      }
    }
    if (blk) blk[tile * 32 + lane] = make_uint2(x1 - v1, x2 - v2);
    sh_w[lane] = x1 - v1;  // counts of the tile's rows before this block
    sh_w[32 + lane] = x2 - v2;
    if (lane == 31) {
      tiletot[0 * ntiles + tile] = x1 & 0xffffu;
      tiletot[1 * ntiles + tile] = x1 >> 16;
      tiletot[2 * ntiles + tile] = x2 & 0xffffu;
      tiletot[3 * ntiles + tile] = x2 >> 16;
    }
  }
  __syncthreads();
  // counts of the tile's rows before this thread's 32 rows
  pre1 = sh_w[threadIdx.x >> 3] + i1 - w1;
  pre2 = sh_w[32 + (threadIdx.x >> 3)] + i2 - w2;
}

// Number of nibbles equal to `sym` among the first p (< 32) of a thread's 32 rows.
__device__ __forceinline__ uint32_t rows_eq_before(const uint32_t out[4], uint32_t sym, uint32_t p);

// Checkpoints of an existing array (the first BWT of the column passes): one CTA per 8192-row tile.
static __global__ void __launch_bounds__(256) tile_count_kernel(const uint8_t* __restrict__ B, uint64_t n, uint64_t ntiles,
                                                               uint2* __restrict__ blk, uint64_t* __restrict__ tiletot,
                                                               uint16_t* __restrict__ rloc /* optional: per row */) {
  __shared__ uint32_t sh_w[64];
  const uint64_t row0 = (uint64_t)blockIdx.x * 8192 + (uint64_t)threadIdx.x * 32;
  uint32_t out[4] = {0, 0, 0, 0};
  if (row0 < n) {
    const uint4 v = *reinterpret_cast<const uint4*>(B + row0 / 2);  // the array is padded
    out[0] = v.x;
    out[1] = v.y;
    out[2] = v.z;
    out[3] = v.w;
    const uint32_t valid = (uint32_t)min((uint64_t)32, n - row0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t vw = valid >= 8u * (i + 1) ? 8u : (valid > 8u * i ? valid - 8u * i : 0u);
      out[i] &= vw >= 8 ? 0xffffffffu : ((1u << (4 * vw)) - 1u);
    }
  }
  uint32_t pre1, pre2;
  tile_checkpoints(out, blockIdx.x, ntiles, sh_w, blk, tiletot, pre1, pre2);
  if (rloc && row0 < n) {
    const uint32_t valid = (uint32_t)min((uint64_t)32, n - row0);
    for (uint32_t p = 0; p < valid; ++p) {
      const uint32_t sym = (out[p >> 3] >> (4 * (p & 7))) & 15u;
      uint32_t r = 0;
      if (sym) {
        const uint32_t c = sym - 1, f = c < 2 ? pre1 : pre2;
        r = ((c & 1) ? (f >> 16) : (f & 0xffffu)) + rows_eq_before(out, sym, p);
      }
      rloc[row0 + p] = (uint16_t)r;
    }
  }
}

// number of nibbles equal to the pattern's nibble among the first `nn` (<= 8) nibbles of w
__device__ __forceinline__ uint32_t nib_eq_count(uint32_t w, uint32_t pat, uint32_t nn) {
  uint32_t x = w ^ pat;
  x |= x >> 1;
  x |= x >> 2;
  const uint32_t eq = ~x & 0x11111111u;
  const uint32_t m = nn >= 8 ? 0xffffffffu : ((1u << (4 * nn)) - 1u);
  return (uint32_t)__popc(eq & m);
}

__device__ __forceinline__ uint32_t rows_eq_before(const uint32_t out[4], uint32_t sym, uint32_t p) {
  const uint32_t pat = sym * 0x11111111u, wq = p >> 3;
  uint32_t cnt = 0;
#pragma unroll
  for (uint32_t i = 0; i < 4; ++i) {
    if (i < wq) cnt += nib_eq_count(out[i], pat, 8);
    else if (i == wq) cnt += nib_eq_count(out[i], pat, p & 7);
  }
  return cnt;
}

// Number of nibbles equal to symbol `sym` (1..4) among the first r rows of 256-row block b.  One thread reads the
// block's vectors it needs (those holding rows < r) with all loads issued before the first use: the kernels built on
// this are latency-bound, what counts is independent loads in flight per thread times threads per SM.
__device__ __forceinline__ uint32_t block_rank(const uint8_t* __restrict__ B, uint64_t b, uint32_t r, uint32_t sym) {
  const uint4* v = reinterpret_cast<const uint4*>(B + b * (kBcrBlock / 2));
  uint4 x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) x[q] = (sym && r > 32u * q) ? __ldg(v + q) : make_uint4(0, 0, 0, 0);
  const uint32_t pat = sym * 0x11111111u;
  uint32_t cnt = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (r > 32u * q) {
      const uint32_t left = r - 32u * q;
      cnt += nib_eq_count(x[q].x, pat, left);
      if (left > 8) cnt += nib_eq_count(x[q].y, pat, left - 8);
      if (left > 16) cnt += nib_eq_count(x[q].z, pat, left - 16);
      if (left > 24) cnt += nib_eq_count(x[q].w, pat, left - 24);
    }
  }
  return sym ? cnt : 0;
}
__device__ __forceinline__ uint32_t blk_field(uint2 v, uint32_t c) {
  const uint32_t w = c < 2 ? v.x : v.y;
  return (c & 1) ? (w >> 16) : (w & 0xffffu);
}

// For every active item (in current row order) the row its next suffix takes in the next BWT:
//   base + (#symbols < c in B) + (#c in B[0..P))      with c = SYM[k] - 1 the base that precedes the item's suffix
// (SYM, the symbol the item's row holds, was computed when the row was inserted; 0 = the read has ended: no insert).
// No access to the read store; two dependent memory round trips per item: (SYM, NP) -> (block vectors, checkpoints).
template <class PosT>
__global__ void __launch_bounds__(256) bcr_rank_kernel(const uint8_t* __restrict__ B, const uint2* __restrict__ blk,
                                                       const uint64_t* __restrict__ scanned, uint64_t ntiles, uint64_t base,
                                                       const uint8_t* __restrict__ SYM, const PosT* __restrict__ NP,
                                                       uint64_t nact, PosT* __restrict__ tmpPos) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nact) return;
  const uint32_t sym = (uint32_t)SYM[k];
  if (!sym) return;
  const uint64_t P = (uint64_t)NP[k];
  const uint64_t b = P / kBcrBlock;
  const uint32_t c = sym - 1;
  const uint2 bv = __ldg(blk + b);
  const uint64_t sc = __ldg(scanned + (uint64_t)c * ntiles + (b >> 5));
  const uint32_t inblk = block_rank(B, b, (uint32_t)(P - b * kBcrBlock), sym);
  tmpPos[k] = (PosT)(base + sc + blk_field(bv, c) + inblk);
}

// partition digit of the column passes: the symbol the insert starts with; ended reads (SYM 0) are dropped
struct SymDigit {
  const uint8_t* SYM;
  __device__ __forceinline__ int operator()(uint64_t k) const {
    const int s = (int)SYM[k];
    return s ? s - 1 : 255;
  }
};
// Move functor of the column passes' partition: carries an item to its slot in new-row order and looks up the symbol
// its new row holds (the base before the suffix; '$' when the read starts here).  That base comes from the read store
// -- the one random gather of a pass -- so the 32-bit word it sits in travels with the item (WIN): consecutive passes
// walk through the same word, and only every 16th pass (or every pass's first for a new item) touches the read store.
// The row of the item's next suffix is computed right here (what bcr_rank_kernel did in a launch of its own): the scatter
// already visits every active item once.  The rank needs no access to the BWT: the kernel that wrote the item's row left
// its tile-local rank in rloc (see tile_checkpoints), and the scanned tile totals supply the rest.
template <class PosT>
struct BcrMove {
  ReadAcc acc;
  const uint32_t* ID;
  const uint32_t* WIN;
  uint32_t j;
  uint32_t uniform_len;  // != 0: every read has this length (no gather of len[])
  uint32_t* ID2;
  PosT* NP2;
  uint8_t* SYM2;
  uint32_t* WIN2;
  // rank lookup: scanned tile totals of the previous pass's BWT, the items' rows, symbols and tile-local ranks
  const uint64_t* scanned;
  uint64_t ntiles, base;
  const uint8_t* SYM;
  const PosT* NP;
  const uint16_t* rloc;
  struct Item {
    uint32_t id, win;
    PosT np;
    uint8_t sym;
  };
  __device__ __forceinline__ Item compute(uint64_t k) const {
    Item it;
    const uint32_t id = ID[k];
    const uint32_t c = (uint32_t)SYM[k] - 1;  // SYM is 1..4 here: ended reads are dropped by the partition digit
    const uint64_t tile = (uint64_t)NP[k] / kMergeTile;
    const uint64_t sc = __ldg(scanned + (uint64_t)c * ntiles + tile);
    const uint32_t l = uniform_len ? uniform_len : acc.rv.len[id];
    it.id = id;
    it.np = (PosT)(base + sc + rloc[k]);
    it.sym = 0;
    it.win = 0;
    if (l == j) return it;
    // stored position of column l - j - 1: the walk goes down through a forward read, up through a reversed one
    const uint32_t p = acc.reversed ? j : l - j - 1;
    const bool fresh = j == 1 || (acc.reversed ? (p & 15u) == 0u : (p & 15u) == 15u);
    const uint32_t w = fresh ? acc.rv.words[acc.rv.woff[id] + (p >> 4)] : WIN[k];
    it.sym = (uint8_t)(((w >> (2 * (p & 15u))) & 3u) + 1u);
    it.win = w;
    return it;
  }
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t dest) const {
    const Item it = compute(k);
    ID2[dest] = it.id;
    NP2[dest] = it.np;
    SYM2[dest] = it.sym;
    WIN2[dest] = it.win;
  }
};

// Scatter half of a column pass's partition (same tile walk and order as part_scatter_kernel<4, SymDigit, BcrMove>), with
// the tile's items laid out in bucket order in shared memory first: the four output arrays are then written by consecutive
// threads at consecutive slots -- whole sectors of the byte array instead of eight bytes per warp and bucket.
// The kernel also prepares the NEXT pass's partition: the items it writes are that pass's input, so while writing them it
// counts their symbols per 2048-item tile of the output order (one atomic per distinct (tile, symbol) and warp) into
// cnt_next -- the next pass then only scans those counters, its count kernel is gone -- and clears the counter buffer
// of the pass after next (three buffers rotate: read, accumulate, clear).
template <class PosT>
__global__ void __launch_bounds__(kPartThreads) bcr_scatter_kernel(BcrMove<PosT> mv, uint64_t n, const uint64_t* __restrict__ tileoff,
                                                                   uint32_t ntiles, unsigned long long* __restrict__ cnt_next,
                                                                   uint32_t ntiles_next, uint64_t* __restrict__ clear_buf,
                                                                   uint32_t clear_n) {
  __shared__ uint32_t wc[kPartThreads / 32][4];
  __shared__ uint32_t boff[5];
  __shared__ uint32_t s_id[kPartTile], s_win[kPartTile];
  __shared__ PosT s_np[kPartTile];
  __shared__ uint8_t s_sym[kPartTile];
  if (threadIdx.x < (kPartThreads / 32) * 4) (&wc[0][0])[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t wbase = (uint64_t)blockIdx.x * kPartTile + (uint64_t)w * (kPartRounds * 32);
  int dg[kPartRounds];
  uint32_t lr[kPartRounds];
#pragma unroll
  for (int r = 0; r < kPartRounds; ++r) {
    const uint64_t k = wbase + r * 32 + lane;
    int d = 4;
    if (k < n) {
      const int sy = (int)mv.SYM[k];
      d = sy ? sy - 1 : 4;  // ended reads are dropped
    }
    dg[r] = d;
    lr[r] = 0;
    const unsigned act = __ballot_sync(0xffffffffu, d < 4);
    if (d < 4) {
      const unsigned m = __match_any_sync(act, d);
      const uint32_t before = wc[w][d];
      lr[r] = before + __popc(m & ((1u << lane) - 1u));
      __syncwarp(m);
      if (lane == __ffs(m) - 1) wc[w][d] = before + __popc(m);
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int q = 0; q < kPartThreads / 32; ++q) {
      const uint32_t t = wc[q][d];
      wc[q][d] = run;
      run += t;
    }
    boff[d + 1] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    boff[0] = 0;
    boff[2] += boff[1];
    boff[3] += boff[2];
    boff[4] += boff[3];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kPartRounds; ++r) {
    const int d = dg[r];
    if (d < 4) {
      const uint64_t k = wbase + r * 32 + lane;
      const uint32_t t = boff[d] + wc[w][d] + lr[r];
      const typename BcrMove<PosT>::Item it = mv.compute(k);
      s_id[t] = it.id;
      s_np[t] = it.np;
      s_win[t] = it.win;
      s_sym[t] = it.sym;
    }
  }
  __syncthreads();
  for (uint32_t i = blockIdx.x * kPartThreads + threadIdx.x; i < clear_n; i += gridDim.x * kPartThreads) clear_buf[i] = 0;
  const uint32_t b1 = boff[1], b2 = boff[2], b3 = boff[3], total = boff[4];
  for (uint32_t base = 0; base < total; base += kPartThreads) {  // warp-uniform trip count (the ballots below)
    const uint32_t t = base + threadIdx.x;
    const bool valid = t < total;
    uint32_t sym = 0;
    uint64_t dest = 0;
    if (valid) {
      const int d = (t >= b1) + (t >= b2) + (t >= b3);
      dest = tileoff[(uint64_t)d * ntiles + blockIdx.x] + (t - boff[d]);
      sym = s_sym[t];
      mv.ID2[dest] = s_id[t];
      mv.NP2[dest] = s_np[t];
      mv.SYM2[dest] = (uint8_t)sym;
      mv.WIN2[dest] = s_win[t];
    }
    if (cnt_next) {
      const bool counts = valid && sym != 0;  // ended reads are dropped by the next partition
      const unsigned act = __ballot_sync(0xffffffffu, counts);
      if (counts) {
        const uint32_t key = (uint32_t)(dest / kPartTile) * 4u + (sym - 1u);
        const unsigned m = __match_any_sync(act, key);
        if (lane == __ffs(m) - 1)
          atomicAdd(&cnt_next[(uint64_t)(sym - 1u) * ntiles_next + dest / kPartTile], (unsigned long long)__popc(m));
      }
    }
  }
}

// tile_k0[t] = number of inserts with row < t*kMergeTile, for t in [0, ntiles] (ntiles+1 entries).
template <class PosT>
__global__ void bcr_tile_bounds_kernel(const PosT* __restrict__ NP, uint64_t nact, uint64_t ntiles,
                                       uint64_t* __restrict__ tile_k0) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nact) return;
  // inserts k-1 and k straddle tile starts in (prev_tile, this_tile]
  const uint64_t lo = k == 0 ? 0 : (uint64_t)NP[k - 1] / kMergeTile + 1;
  const uint64_t hi = k == nact ? ntiles : (uint64_t)NP[k] / kMergeTile;
  for (uint64_t t = lo; t <= hi && t <= ntiles; ++t) tile_k0[t] = k;
}

// Streaming merge of the old BWT (n_old rows) with the sorted inserts (rows NP[0..nact) of the NEW array, symbols
// SYM) into Bn (n_new rows), plus both levels of rank checkpoints of Bn.  A tile = 8192 output rows (4 KB), a thread
// 32 rows (one 16-byte store).  The old rows of a tile are staged in shared memory; a thread first takes the 32 old
// nibbles that would land on its rows if it had no inserts (five shared-memory words and four funnel shifts), then
// opens a 4-bit gap for each of its inserts (the bits of its word of the tile's insert mask), which pushes the
// surplus old nibbles out at the top -- exactly the ones the next thread starts with.
//
// The kernel is persistent (a CTA walks tiles blockIdx.x, + gridDim.x, ...) and software-pipelined, because one tile
// is a chain of dependent global accesses (tile bounds -> old rows / insert rows -> symbols) and the first version,
// one CTA per tile, spent its time waiting on that chain (1.6 TB/s):
//   * the insert bounds of the tile after next are loaded while the current tile is processed,
//   * the old rows of the next tile are copied global -> shared by one bulk copy (cp.async.bulk, completion on an
//     mbarrier) into the other half of a double buffer,
//   * the first insert (row, symbol) of every thread for the next tile is loaded into registers.
// Bulk asynchronous copy global -> shared (the non-tensor form of TMA: one instruction moves the whole stretch, the
// copy engine generates the addresses) completing on an mbarrier, and the mbarrier calls around it.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem)),
               "l"(gmem), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

static const int kMergeOldWords = kMergeTile / 8 + 16;  // staged old rows of a tile, 32-bit words

// Stages the old rows of the tile whose inserts are [k0, k1): ONE bulk copy, issued by one thread, from the 32-row-aligned
// position at or below the tile's first old row (the arrays are padded) -- 16-byte granules, at most kMergeTile / 32 + 1 of
// them.  The copy completes on `bar` (expect_tx = its bytes), which every thread of the CTA waits on before it reads.
__device__ __forceinline__ void merge_stage_old(const uint8_t* __restrict__ Bo, uint64_t tile, uint64_t n_new, uint64_t k0,
                                                uint64_t k1, uint32_t* dst, uint64_t* bar) {
  if (threadIdx.x != 0) return;
  const uint64_t t0 = tile * kMergeTile;
  const uint64_t t1 = min(t0 + (uint64_t)kMergeTile, n_new);
  const uint64_t a0 = t0 - k0;
  const uint32_t nold = (uint32_t)((t1 - t0) - (k1 - k0));
  const uint64_t a0al = a0 & ~(uint64_t)31;
  const uint32_t nvec = (nold + (uint32_t)(a0 - a0al) + 31) / 32;
  mbar_arrive_expect_tx(bar, nvec * 16u);
  if (nvec) bulk_copy_g2s(dst, Bo + a0al / 2, nvec * 16u, bar);
}

template <class PosT>
__global__ void __launch_bounds__(kMergeThreads, 8) bcr_merge_kernel(const uint8_t* __restrict__ Bo, uint64_t n_old,
                                                                  const PosT* __restrict__ NP, const uint8_t* __restrict__ SYM,
                                                                  uint64_t nact, const uint64_t* __restrict__ tile_k0,
                                                                  uint8_t* __restrict__ Bn, uint64_t n_new, uint64_t ntiles,
                                                                  uint2* __restrict__ blk, uint64_t* __restrict__ tiletot,
                                                                  uint16_t* __restrict__ rloc /* optional: per insert */) {
  __shared__ __align__(16) uint32_t oldw[2][kMergeOldWords];
  __shared__ uint32_t mask[kMergeTile / 32];
  __shared__ uint8_t symsh[kMergeTile];  // symbols of the tile's inserts, by ordinal
  __shared__ uint32_t ckpt[64];
  __shared__ uint32_t sh[8];
  __shared__ uint32_t outsm[kMergeTile / 8];  // the finished tile (for the tile-local ranks)
  __shared__ uint2 presm[kMergeThreads];      // symbol counts of each thread's 256-row block before its rows
  __shared__ __align__(8) uint64_t stage_bar[2];  // completion of the bulk copy into oldw[0] / oldw[1]
  const uint64_t G = gridDim.x;
  uint64_t tile = blockIdx.x;
  if (tile >= ntiles) return;
  if (threadIdx.x == 0) {
    mbar_init(&stage_bar[0], 1);
    mbar_init(&stage_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  mask[threadIdx.x] = 0;
  __syncthreads();
  uint32_t phases = 0;  // bit s = parity to wait for on stage_bar[s]
  // prologue: bounds of my first two tiles, old rows + first insert of the first
  uint64_t k0 = tile_k0[tile], k1 = tile_k0[tile + 1];
  uint64_t kn0 = 0, kn1 = 0;
  if (tile + G < ntiles) {
    kn0 = tile_k0[tile + G];
    kn1 = tile_k0[tile + G + 1];
  }
  merge_stage_old(Bo, tile, n_new, k0, k1, oldw[0], &stage_bar[0]);
  uint64_t pre_np = 0;
  uint32_t pre_sym = 0;
  if (k0 + threadIdx.x < k1) {
    pre_np = (uint64_t)NP[k0 + threadIdx.x];
    pre_sym = SYM[k0 + threadIdx.x];
  }
  int s = 0;
  for (;; tile += G, s ^= 1) {
    const uint64_t tn = tile + G;
    const bool have_next = tn < ntiles;
    // next tile: stage its old rows into the other buffer (free since the end of the previous iteration)
    if (have_next) merge_stage_old(Bo, tn, n_new, kn0, kn1, oldw[s ^ 1], &stage_bar[s ^ 1]);
    // bounds of the tile after next; first insert of the next tile
    uint64_t knn0 = 0, knn1 = 0;
    if (tn + G < ntiles) {
      knn0 = tile_k0[tn + G];
      knn1 = tile_k0[tn + G + 1];
    }
    uint64_t nxt_np = 0;
    uint32_t nxt_sym = 0;
    if (have_next && kn0 + threadIdx.x < kn1) {
      nxt_np = (uint64_t)NP[kn0 + threadIdx.x];
      nxt_sym = SYM[kn0 + threadIdx.x];
    }

    const uint64_t t0 = tile * kMergeTile;
    const uint64_t t1 = min(t0 + (uint64_t)kMergeTile, n_new);
    const uint64_t a0 = t0 - k0;  // first old row of this tile
    const uint32_t shift = (uint32_t)(a0 & 31);  // nibbles between the staged vector start and a0
    // (mask[] is all zero here: every thread clears its word right after reading it, before the barrier that ends the
    //  iteration; the prologue clears it once)
    // insert mask + symbols by ordinal (the first insert of every thread is already in registers)
    if (k0 + threadIdx.x < k1) {
      const uint32_t r = (uint32_t)(pre_np - t0);
      atomicOr(&mask[r >> 5], 1u << (r & 31));
      symsh[threadIdx.x] = (uint8_t)pre_sym;
    }
    for (uint64_t k = k0 + threadIdx.x + kMergeThreads; k < k1; k += kMergeThreads) {
      const uint32_t r = (uint32_t)((uint64_t)NP[k] - t0);
      atomicOr(&mask[r >> 5], 1u << (r & 31));
      symsh[k - k0] = SYM[k];
    }
    mbar_wait(&stage_bar[s], (phases >> s) & 1u);  // the current tile's old rows (copy issued one iteration ago) have landed
    phases ^= 1u << s;
    __syncthreads();
    const uint32_t m = mask[threadIdx.x];
    mask[threadIdx.x] = 0;  // for the next tile (its atomicOrs come after the barrier at the end of this iteration)
    const uint32_t pre = block_excl_scan32((uint32_t)__popc(m), sh);  // inserts before my first row
    // the 32 old nibbles starting at the one that lands on my first row
    const uint32_t* ow = oldw[s];
    const uint32_t o = shift + threadIdx.x * 32 - pre;
    const uint32_t wi = o >> 3, sft = 4 * (o & 7);
    uint32_t out[4];
    {
      const uint32_t w0 = ow[wi], w1 = ow[wi + 1], w2 = ow[wi + 2], w3 = ow[wi + 3], w4 = ow[wi + 4];
      out[0] = __funnelshift_r(w0, w1, sft);
      out[1] = __funnelshift_r(w1, w2, sft);
      out[2] = __funnelshift_r(w2, w3, sft);
      out[3] = __funnelshift_r(w3, w4, sft);
    }
    // open a gap for every insert, lowest row first
    uint32_t mm = m;
    uint32_t kk = pre;
    while (mm) {
      const uint32_t p = (uint32_t)__ffs(mm) - 1;
      mm &= mm - 1;
      const uint32_t sy = symsh[kk++];
      const uint32_t wq = p >> 3, bit = 4 * (p & 7);
#pragma unroll
      for (int i = 3; i >= 1; --i)
        if ((uint32_t)i > wq) out[i] = (out[i] << 4) | (out[i - 1] >> 28);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if ((uint32_t)i == wq) {
          const uint32_t lowm = (1u << bit) - 1u;
          const uint32_t up = bit < 28 ? ((out[i] >> bit) << (bit + 4)) : 0u;
          out[i] = (out[i] & lowm) | (sy << bit) | up;
        }
      }
    }
    // rows past the end of the array stay zero
    const uint64_t row0 = t0 + (uint64_t)threadIdx.x * 32;
    uint32_t valid = 32;
    if (t1 - t0 != (uint64_t)kMergeTile) {  // only the last tile of the array is partial
      valid = row0 >= t1 ? 0u : (uint32_t)min((uint64_t)32, t1 - row0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t vw = valid >= 8u * (i + 1) ? 8u : (valid > 8u * i ? valid - 8u * i : 0u);
        out[i] &= vw >= 8 ? 0xffffffffu : ((1u << (4 * vw)) - 1u);
      }
    }
    if (valid) *reinterpret_cast<uint4*>(Bn + row0 / 2) = make_uint4(out[0], out[1], out[2], out[3]);
    uint32_t cp1, cp2;
    if (rloc) {  // the finished tile, for the tile-local ranks below (visible after the barriers inside tile_checkpoints)
#pragma unroll
      for (int i = 0; i < 4; ++i) outsm[4 * threadIdx.x + i] = out[i];
    }
    tile_checkpoints(out, tile, ntiles, ckpt, blk, tiletot, cp1, cp2, rloc ? presm : nullptr);
    if (rloc) {
      // Tile-local rank of every insert of the tile: rows of the tile before it that hold the same symbol = the count
      // before its thread's 32 rows (cp1 / cp2) + the matches among the rows of that thread below it.  The inserts are
      // taken by ordinal, one per thread and round, from the finished tile in shared memory: a thread looping over its
      // OWN inserts would leave most lanes idle (a tile holds 8192 / (pass + 1) inserts, under one per thread after
      // pass 31, but the busiest lane of a warp still has two or three).
      for (uint64_t o = threadIdx.x; o < k1 - k0; o += kMergeThreads) {
        const uint32_t r = (uint32_t)((uint64_t)NP[k0 + o] - t0);
        const uint32_t sy = symsh[o];
        uint32_t v = 0;
        if (sy) {
          const uint32_t owner = r >> 5, p = r & 31u, c = sy - 1;
          const uint2 pf = presm[owner];  // counts inside the owner's 256-row block; ckpt[] holds the blocks before it
          const uint32_t f = c < 2 ? ckpt[owner >> 3] + pf.x : ckpt[32 + (owner >> 3)] + pf.y;
          const uint32_t* ow4 = outsm + 4 * owner;
          const uint32_t pat = sy * 0x11111111u, wq = p >> 3;
          uint32_t cnt = nib_eq_count(ow4[wq], pat, p & 7);
          for (uint32_t i = 0; i < wq; ++i) cnt += nib_eq_count(ow4[i], pat, 8);
          v = ((c & 1) ? (f >> 16) : (f & 0xffffu)) + cnt;
        }
        rloc[k0 + o] = (uint16_t)v;
      }
    }
    __syncthreads();  // mask / symsh / oldw[s] / ckpt are free for the next iteration
    if (!have_next) break;
    k0 = kn0;
    k1 = kn1;
    kn0 = knn0;
    kn1 = knn1;
    pre_np = nxt_np;
    pre_sym = nxt_sym;
  }
}

// ---- merge, 64 rows per thread ------------------------------------------------------------------------------------------
//
// bcr_merge_kernel is bound by issue slots (70 % busy, profiles/r2_summary.md), and most of what it issues per tile does
// not depend on the rows: the block-wide scan of the insert counts, the scan of the symbol counts, the pipeline
// bookkeeping (bounds of the tiles ahead, first insert of the next tile, staging), the barriers.  Here a thread owns 64
// rows -- two 32-row groups treated exactly as above, one after the other (old nibbles by funnel shifts, a gap per
// insert) -- so a tile is walked by 128 threads and all of that is paid once per 64 rows instead of once per 32.  The
// per-256-row checkpoints (blk) of the first kernel are not produced (no caller uses them any more).
static const int kM64Threads = 128;

// The 32 old nibbles that land on a 32-row group when it has no inserts, then a gap opened for each insert (mask m, lowest
// row first; kk = ordinal of the group's first insert among the tile's inserts).
__device__ __forceinline__ void merge_rows32(const uint32_t* __restrict__ ow, uint32_t o, uint32_t m, uint32_t kk,
                                             const uint8_t* __restrict__ symsh, uint32_t out[4]) {
  const uint32_t wi = o >> 3, sft = 4 * (o & 7);
  {
    const uint32_t w0 = ow[wi], w1 = ow[wi + 1], w2 = ow[wi + 2], w3 = ow[wi + 3], w4 = ow[wi + 4];
    out[0] = __funnelshift_r(w0, w1, sft);
    out[1] = __funnelshift_r(w1, w2, sft);
    out[2] = __funnelshift_r(w2, w3, sft);
    out[3] = __funnelshift_r(w3, w4, sft);
  }
  while (m) {
    const uint32_t p = (uint32_t)__ffs(m) - 1;
    m &= m - 1;
    const uint32_t sy = symsh[kk++];
    const uint32_t wq = p >> 3, bit = 4 * (p & 7);
    // words above the insert's word move up by a nibble
    if (wq < 3) out[3] = __funnelshift_l(out[2], out[3], 4);
    if (wq < 2) out[2] = __funnelshift_l(out[1], out[2], 4);
    if (wq < 1) out[1] = __funnelshift_l(out[0], out[1], 4);
    // the insert's word: the nibbles from `bit` on move up by one, the symbol takes their place
    const uint32_t w = wq == 0 ? out[0] : wq == 1 ? out[1] : wq == 2 ? out[2] : out[3];
    const uint32_t lowm = (1u << bit) - 1u;
    const uint32_t nw = (w & lowm) | (sy << bit) | ((w & ~lowm) << 4);
    if (wq == 0) out[0] = nw;
    if (wq == 1) out[1] = nw;
    if (wq == 2) out[2] = nw;
    if (wq == 3) out[3] = nw;
  }
}

// A | C<<16 and G | T<<16 counts of the 32 nibbles of four words (bit planes, see tile_checkpoints)
__device__ __forceinline__ void plane_counts32(const uint32_t out[4], uint32_t& w1, uint32_t& w2) {
  const uint32_t M = 0x11111111u;
  const uint32_t p0 = (out[0] & M) | ((out[1] << 1) & (M << 1)) | ((out[2] << 2) & (M << 2)) | ((out[3] << 3) & (M << 3));
  const uint32_t p1 = ((out[0] >> 1) & M) | (out[1] & (M << 1)) | ((out[2] << 1) & (M << 2)) | ((out[3] << 2) & (M << 3));
  const uint32_t p2 = ((out[0] >> 2) & M) | ((out[1] >> 1) & (M << 1)) | (out[2] & (M << 2)) | ((out[3] << 1) & (M << 3));
  w1 = (uint32_t)__popc(p0 & ~p1) | ((uint32_t)__popc(p1 & ~p0) << 16);
  w2 = (uint32_t)__popc(p0 & p1) | ((uint32_t)__popc(p2) << 16);
}

template <class PosT>
__global__ void __launch_bounds__(kM64Threads, 12) bcr_merge64_kernel(const uint8_t* __restrict__ Bo, uint64_t n_old,
                                                                      const PosT* __restrict__ NP, const uint8_t* __restrict__ SYM,
                                                                      uint64_t nact, const uint64_t* __restrict__ tile_k0,
                                                                      uint8_t* __restrict__ Bn, uint64_t n_new, uint64_t ntiles,
                                                                      uint64_t* __restrict__ tiletot, uint16_t* __restrict__ rloc) {
  __shared__ __align__(16) uint32_t oldw[2][kMergeOldWords];
  __shared__ uint32_t mask[kMergeTile / 32];   // thread t owns words 2 t and 2 t + 1
  __shared__ uint8_t symsh[kMergeTile];        // symbols of the tile's inserts, by ordinal
  __shared__ uint32_t sh[4];                   // insert counts of the warps
  __shared__ uint32_t wtot[2][4];              // symbol counts of the warps (A | C << 16, G | T << 16)
  __shared__ uint32_t outsm[kMergeTile / 8];   // the finished tile (for the tile-local ranks)
  __shared__ uint2 presm[2 * kM64Threads];     // symbol counts of the owner's warp before each 32-row group
  __shared__ __align__(8) uint64_t stage_bar[2];
  (void)n_old;
  (void)nact;
  const uint32_t G = gridDim.x;
  const uint32_t T = threadIdx.x, lane = T & 31, wp = T >> 5;
  uint64_t tile = blockIdx.x;
  if (tile >= ntiles) return;
  if (T == 0) {
    mbar_init(&stage_bar[0], 1);
    mbar_init(&stage_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  mask[2 * T] = 0;
  mask[2 * T + 1] = 0;
  __syncthreads();
  uint32_t phases = 0;
  uint64_t k0 = tile_k0[tile], k1 = tile_k0[tile + 1];
  uint64_t kn0 = 0, kn1 = 0;
  if (tile + G < ntiles) {
    kn0 = tile_k0[tile + G];
    kn1 = tile_k0[tile + G + 1];
  }
  merge_stage_old(Bo, tile, n_new, k0, k1, oldw[0], &stage_bar[0]);
  uint64_t pre_np = 0;
  uint32_t pre_sym = 0;
  if (k0 + T < k1) {
    pre_np = (uint64_t)NP[k0 + T];
    pre_sym = SYM[k0 + T];
  }
  int s = 0;
  for (;; tile += G, s ^= 1) {
    const uint64_t tn = tile + G;
    const bool have_next = tn < ntiles;
    if (have_next) merge_stage_old(Bo, tn, n_new, kn0, kn1, oldw[s ^ 1], &stage_bar[s ^ 1]);
    uint64_t knn0 = 0, knn1 = 0;
    if (tn + G < ntiles) {
      knn0 = tile_k0[tn + G];
      knn1 = tile_k0[tn + G + 1];
    }
    uint64_t nxt_np = 0;
    uint32_t nxt_sym = 0;
    if (have_next && kn0 + T < kn1) {
      nxt_np = (uint64_t)NP[kn0 + T];
      nxt_sym = SYM[kn0 + T];
    }
    const uint64_t t0 = tile * kMergeTile;
    const uint64_t t1 = min(t0 + (uint64_t)kMergeTile, n_new);
    const uint32_t nins = (uint32_t)(k1 - k0);
    const uint32_t shift = (uint32_t)((t0 - k0) & 31);  // nibbles between the staged vector start and the tile's first old row
    // insert mask + symbols by ordinal (mask[] is all zero here; the first insert of every thread is in registers)
    if (T < nins) {
      const uint32_t r = (uint32_t)(pre_np - t0);
      atomicOr(&mask[r >> 5], 1u << (r & 31));
      symsh[T] = (uint8_t)pre_sym;
    }
    for (uint32_t o = T + kM64Threads; o < nins; o += kM64Threads) {
      const uint32_t r = (uint32_t)((uint64_t)NP[k0 + o] - t0);
      atomicOr(&mask[r >> 5], 1u << (r & 31));
      symsh[o] = SYM[k0 + o];
    }
    mbar_wait(&stage_bar[s], (phases >> s) & 1u);
    phases ^= 1u << s;
    __syncthreads();
    const uint32_t m0 = mask[2 * T], m1 = mask[2 * T + 1];
    mask[2 * T] = 0;  // for the next tile (its atomicOrs come after the barrier at the end of this iteration)
    mask[2 * T + 1] = 0;
    const uint32_t c0 = (uint32_t)__popc(m0), c1 = (uint32_t)__popc(m1);
    // inserts before my first row: scan across the warp, warp sums through shared memory
    uint32_t incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += y;
    }
    if (lane == 31) sh[wp] = incl;
    __syncthreads();
    uint32_t pre = incl - (c0 + c1);
#pragma unroll
    for (uint32_t q = 0; q < 3; ++q)
      if (q < wp) pre += sh[q];
    const uint32_t* ow = oldw[s];
    uint32_t outA[4], outB[4];
    merge_rows32(ow, shift + T * 64 - pre, m0, pre, symsh, outA);
    merge_rows32(ow, shift + T * 64 + 32 - (pre + c0), m1, pre + c0, symsh, outB);
    const uint64_t row0 = t0 + (uint64_t)T * 64;
    if (t1 - t0 != (uint64_t)kMergeTile) {  // only the last tile of the array is partial: rows past the end stay zero
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t* out = h ? outB : outA;
        const uint64_t r = row0 + 32 * h;
        const uint32_t valid = r >= t1 ? 0u : (uint32_t)min((uint64_t)32, t1 - r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t vw = valid >= 8u * (i + 1) ? 8u : (valid > 8u * i ? valid - 8u * i : 0u);
          out[i] &= vw >= 8 ? 0xffffffffu : ((1u << (4 * vw)) - 1u);
        }
      }
    }
    if (row0 < t1) {
      uint4* dst = reinterpret_cast<uint4*>(Bn + row0 / 2);
      dst[0] = make_uint4(outA[0], outA[1], outA[2], outA[3]);
      if (row0 + 32 < t1) dst[1] = make_uint4(outB[0], outB[1], outB[2], outB[3]);
    }
    // symbol counts: per thread, scanned across the warp; warp sums through shared memory
    uint32_t a1, a2, b1, b2;
    plane_counts32(outA, a1, a2);
    plane_counts32(outB, b1, b2);
    const uint32_t w1 = a1 + b1, w2 = a2 + b2;  // packed 16-bit fields: a tile holds at most 8192 of one symbol
    uint32_t i1 = w1, i2 = w2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y1 = __shfl_up_sync(0xffffffffu, i1, o), y2 = __shfl_up_sync(0xffffffffu, i2, o);
      if ((int)lane >= o) {
        i1 += y1;
        i2 += y2;
      }
    }
    if (lane == 31) {
      wtot[0][wp] = i1;
      wtot[1][wp] = i2;
    }
    if (rloc) {
      reinterpret_cast<uint4*>(outsm)[2 * T] = make_uint4(outA[0], outA[1], outA[2], outA[3]);
      reinterpret_cast<uint4*>(outsm)[2 * T + 1] = make_uint4(outB[0], outB[1], outB[2], outB[3]);
      presm[2 * T] = make_uint2(i1 - w1, i2 - w2);
      presm[2 * T + 1] = make_uint2(i1 - w1 + a1, i2 - w2 + a2);
    }
    __syncthreads();
    if (T == 0) {
      const uint32_t x1 = wtot[0][0] + wtot[0][1] + wtot[0][2] + wtot[0][3], x2 = wtot[1][0] + wtot[1][1] + wtot[1][2] + wtot[1][3];
      tiletot[0 * ntiles + tile] = x1 & 0xffffu;
      tiletot[1 * ntiles + tile] = x1 >> 16;
      tiletot[2 * ntiles + tile] = x2 & 0xffffu;
      tiletot[3 * ntiles + tile] = x2 >> 16;
    }
    if (rloc) {
      // tile-local rank of every insert, one insert per thread and round, from the finished tile in shared memory
      const uint32_t e1[4] = {0u, wtot[0][0], wtot[0][0] + wtot[0][1], wtot[0][0] + wtot[0][1] + wtot[0][2]};
      const uint32_t e2[4] = {0u, wtot[1][0], wtot[1][0] + wtot[1][1], wtot[1][0] + wtot[1][1] + wtot[1][2]};
      for (uint32_t o = T; o < nins; o += kM64Threads) {
        const uint32_t r = (uint32_t)((uint64_t)NP[k0 + o] - t0);
        const uint32_t sy = symsh[o];
        uint32_t v = 0;
        if (sy) {
          const uint32_t grp = r >> 5, p = r & 31u, c = sy - 1, ow_wp = grp >> 6;  // 32-row group, its thread's warp
          const uint2 pf = presm[grp];
          const uint32_t base1 = ow_wp == 0 ? e1[0] : ow_wp == 1 ? e1[1] : ow_wp == 2 ? e1[2] : e1[3];
          const uint32_t base2 = ow_wp == 0 ? e2[0] : ow_wp == 1 ? e2[1] : ow_wp == 2 ? e2[2] : e2[3];
          const uint32_t f = c < 2 ? base1 + pf.x : base2 + pf.y;
          const uint32_t* ow4 = outsm + 4 * grp;
          const uint32_t pat = sy * 0x11111111u, wq = p >> 3;
          uint32_t cnt = nib_eq_count(ow4[wq], pat, p & 7);
          for (uint32_t i = 0; i < wq; ++i) cnt += nib_eq_count(ow4[i], pat, 8);
          v = ((c & 1) ? (f >> 16) : (f & 0xffffu)) + cnt;
        }
        rloc[k0 + o] = (uint16_t)v;
      }
    }
    __syncthreads();  // symsh / oldw[s] / sh / wtot / outsm / presm are free for the next iteration
    if (!have_next) break;
    k0 = kn0;
    k1 = kn1;
    kn0 = knn0;
    kn1 = knn1;
    pre_np = nxt_np;
    pre_sym = nxt_sym;
  }
}

// ---- 3. run-length encoding ------------------------------------------------------------------------------------------

// Run heads of a nibble array, 32 rows (one 16-byte vector) per thread: bit i of the returned mask is set when row
// 32 t + i starts a run (differs from the row before it; row 0 always does).  Rows at or past n are never heads.
__device__ __forceinline__ uint32_t head_mask32(const uint8_t* __restrict__ B, uint64_t t, uint64_t n) {
  const uint64_t row0 = t * 32;
  if (row0 >= n) return 0u;
  const uint4 v = *reinterpret_cast<const uint4*>(B + row0 / 2);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t prev = t ? (uint32_t)(B[row0 / 2 - 1] >> 4) : 0xfu;  // 0xf is no symbol: row 0 is a head
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t x = w[i] ^ ((w[i] << 4) | prev);  // nibble k: row k of the word against the row before it
    prev = w[i] >> 28;
    x |= x >> 1;
    x |= x >> 2;
    x &= 0x11111111u;                 // bit 4 k = row k differs
    x = (x | (x >> 3)) & 0x03030303u; // gather the eight flags of the word into its low byte: pairs ...
    x = (x | (x >> 6)) & 0x000f000fu; // ... nibbles ...
    x = (x | (x >> 12)) & 0xffu;      // ... byte
    m |= x << (8 * i);
  }
  const uint64_t left = n - row0;
  if (left < 32) m &= (1u << (uint32_t)left) - 1u;
  return m;
}

// heads per 8192-row tile (256 threads x 32 rows)
static __global__ void __launch_bounds__(256) head_count_kernel(const uint8_t* __restrict__ B, uint64_t n,
                                                                 uint64_t* __restrict__ tilecnt) {
  __shared__ uint32_t sh[8];
  const uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  uint32_t c = (uint32_t)__popc(head_mask32(B, t, n));
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += sh[q];
    tilecnt[blockIdx.x] = s;
  }
}

// heads[tileoff[tile] + ...] = rows that start a run, in row order
static __global__ void __launch_bounds__(256) head_write_kernel(const uint8_t* __restrict__ B, uint64_t n,
                                                                 const uint64_t* __restrict__ tileoff, uint64_t* __restrict__ heads) {
  __shared__ uint32_t sh[8];
  const uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  uint32_t m = head_mask32(B, t, n);
  uint64_t at = tileoff[blockIdx.x] + block_excl_scan32((uint32_t)__popc(m), sh);
  while (m) {
    const uint32_t i = (uint32_t)__ffs(m) - 1;
    m &= m - 1;
    heads[at++] = t * 32 + i;
  }
}

// Length of run r of a segment of n rows.  Seams of the bucket-sharded build: a run that continues from the previous
// segment belongs to the segment where it starts, so a segment may skip its first run (skip_first) and its last run
// may be longer than its local rows (extend_last = rows of the following segments that continue it).
__device__ __forceinline__ uint64_t run_length(const uint64_t* __restrict__ heads, uint64_t r, uint64_t n_runs, uint64_t n,
                                               int skip_first, uint64_t extend_last) {
  if (r == 0 && skip_first) return 0;
  uint64_t len = (r + 1 < n_runs ? heads[r + 1] : n) - heads[r];
  if (r + 1 == n_runs) len += extend_last;
  return len;
}

// bytes needed by run r: ceil(len / 31)   (src/bwt.cpp:16-27: a run is cut whenever its count reaches 31)
static __global__ void run_bytes_kernel(const uint64_t* __restrict__ heads, uint64_t n_runs, uint64_t n, int skip_first,
                                        uint64_t extend_last, uint64_t* __restrict__ nbytes) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  nbytes[r] = (run_length(heads, r, n_runs, n, skip_first, extend_last) + 30) / 31;
}

static __global__ void run_emit_kernel(const uint8_t* __restrict__ B, const uint64_t* __restrict__ heads,
                                       const uint64_t* __restrict__ boff, uint64_t n_runs, uint64_t n, int skip_first,
                                       uint64_t extend_last, uint8_t* __restrict__ rle) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  uint64_t len = run_length(heads, r, n_runs, n, skip_first, extend_last);
  const uint8_t s = (uint8_t)(nib_get(B, heads[r]) << 5);
  uint64_t o = boff[r];
  while (len > 0) {
    const uint32_t c = len >= 31 ? 31u : (uint32_t)len;
    rle[o++] = s | (uint8_t)c;
    len -= c;
  }
}

// histogram of read lengths: hist[l] += 1  (one atomic per distinct length per warp: read sets are mostly one length)
static __global__ void len_hist_kernel(const uint32_t* __restrict__ len, uint64_t n, uint32_t* __restrict__ hist) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool have = i < n;
  const unsigned act = __ballot_sync(0xffffffffu, have);
  if (!have) return;
  const uint32_t l = len[i];
  const unsigned m = __match_any_sync(act, l);
  if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&hist[l], (uint32_t)__popc(m));
}

// persistent grid of the merge kernel: as many CTAs as are resident at once (asked of the runtime: the register count
// decides it), never more than tiles
template <class PosT>
inline unsigned merge_grid(const sgx_ctx* ctx, uint64_t ntiles) {
  static int per_sm = 0;
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bcr_merge_kernel<PosT>, kMergeThreads, 0) != cudaSuccess || n < 1) n = 4;
    per_sm = n;
  }
  const uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)per_sm;
  return (unsigned)(ntiles < cap ? (ntiles ? ntiles : 1) : cap);
}

template <class PosT>
inline unsigned merge64_grid(const sgx_ctx* ctx, uint64_t ntiles) {
  static int per_sm = 0;
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bcr_merge64_kernel<PosT>, kM64Threads, 0) != cudaSuccess || n < 1) n = 8;
    per_sm = n;
  }
  const uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)per_sm;
  return (unsigned)(ntiles < cap ? (ntiles ? ntiles : 1) : cap);
}

// The merge of one column pass.  SGX_MERGE=32 selects the first kernel (32 rows per thread), the default is the
// 64-rows-per-thread kernel.
template <class PosT>
inline void launch_merge(sgx_ctx* ctx, const uint8_t* Bo, uint64_t n_old, const PosT* NP, const uint8_t* SYM, uint64_t nact,
                         const uint64_t* tile_k0, uint8_t* Bn, uint64_t n_new, uint64_t ntiles, uint64_t* tiletot, uint16_t* rloc) {
  static int which = 0;
  if (which == 0) {
    const char* e = getenv("SGX_MERGE");
    which = (e && atoi(e) == 32) ? 32 : 64;
  }
  if (which == 32) {
    SGX_LAUNCH(ctx, (bcr_merge_kernel<PosT>), merge_grid<PosT>(ctx, ntiles), kMergeThreads, 0, Bo, n_old, NP, SYM, nact, tile_k0, Bn,
               n_new, ntiles, (uint2*)nullptr, tiletot, rloc);
  } else {
    SGX_LAUNCH(ctx, (bcr_merge64_kernel<PosT>), merge64_grid<PosT>(ctx, ntiles), kM64Threads, 0, Bo, n_old, NP, SYM, nact, tile_k0,
               Bn, n_new, ntiles, tiletot, rloc);
  }
}

// ---- host helpers defined in index.cu, shared with index_sharded.cu ------------------------------------------------

// Order of the read-start suffixes (d_sai, the .sai file) and of the N sentinel rows (d_sord) under ordering model
// `order`; both arrays hold N read ids.
void index_read_order(sgx_ctx* ctx, const sgx_reads* reads, const ReadAcc& acc, int order, uint32_t* d_sai, uint32_t* d_sord);
// Stable content sort of n read ids (result in `ids`); scratch arrays hold n elements, part_ws part_ws_elems(n, 256).
void index_content_sort(sgx_ctx* ctx, const sgx_reads* reads, const ReadAcc& acc, uint32_t* ids, uint32_t* ids_alt,
                        uint64_t* keys, uint64_t* keys_alt, uint64_t n, uint64_t* part_ws);
// nact[j] = reads with length >= j (the inserts of column pass j), j in [1, max_len].
std::vector<uint64_t> index_active_counts(sgx_ctx* ctx, const sgx_reads* reads);
// bytes of a 4-bit symbol array of `rows` rows, padded so that whole-tile / whole-vector accesses stay in bounds
inline uint64_t bwt_array_bytes(uint64_t rows) { return (((rows + kMergeTile - 1) / kMergeTile + 2) * kMergeTile) / 2 + 128; }

// Run heads of a segment (rows [0, n) of the nibble array B) and what its neighbours need to know about its ends.
struct RleSegment {
  DevBuf<uint64_t> heads;  // n_runs start rows
  uint64_t n_runs = 0, n_rows = 0;
  uint32_t first_sym = 0, last_sym = 0;
  uint64_t first_len = 0, last_len = 0;  // rows of the first / last run (equal to n_rows when the segment is one run)
};
void rle_heads(sgx_ctx* ctx, const uint8_t* B, uint64_t n, RleSegment* seg);
// Greedy-31 run bytes of the segment (src/bwt.cpp:16-27); see run_length() for skip_first / extend_last.
void rle_emit(sgx_ctx* ctx, const uint8_t* B, const RleSegment& seg, int skip_first, uint64_t extend_last,
              DevBuf<uint8_t>* d_rle, uint64_t* n_bytes);

}  // namespace sgx

#endif  // SGX_INDEX_KERNELS_CUH_
