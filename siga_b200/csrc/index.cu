// BWT / read-index construction on the device (`siga index`).
//
// Replaces SuffixArrayBuilder::build + BWT::BWT of the reference (src/suffix_array_builder.cpp:31-282 `sais`,
// 474-673 `sais2` via 3rdparty/sais.hxx; src/bwt.cpp:7-32).  The reference suffix-sorts the concatenated text on one
// CPU thread with SA-IS.  Here the BWT of the read collection is built directly, without a suffix array, by
// column insertion (the BCR idea the reference vendors in 3rdparty/bcr.c but never calls):
//
//   1. Read ordering (N items): LSD stable partition passes over the read columns give the content order; dense
//      content ranks + prefix doubling over the read sequence give the order of the read-start suffixes under
//      ordering model B (`sais2`: all '$' equal, comparison runs on into the next read).  This order IS the .sai
//      file, and it fixes the order of the N sentinel rows.  Model A (`sais`) orders ties by read index.
//   2. Column passes j = 1..Lmax: the suffixes of length j are inserted into the BWT of all shorter suffixes.  The
//      row of read i's new suffix is N + (#symbols < c before) + rank_c(row of its previous suffix); because that is
//      monotone in the previous row for a fixed c, ordering the inserts is ONE stable 4-way partition by c
//      (count -> scan -> scatter), and the insertion itself is a streaming merge of the old BWT with the sorted
//      inserts.  The merge is the dominant kernel: it reads and writes the partial BWT (4 bits per symbol) once
//      per column (HBM-bound).
//   3. Greedy run-length encoding (runs capped at 31) of the final symbol array into .bwt run bytes.
//
// Positions are 32-bit while the BWT has < 2^32 rows and 64-bit otherwise (PosT template parameter).
#include "index.cuh"

#include <algorithm>

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

__global__ void iota_kernel(uint32_t* a, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (uint32_t)i;
}

// Digit of a read for columns [col0, col0+ncols): base-5 number, leftmost column most significant, '$' (=0) past the
// end of the read so that a read that is a proper prefix of another sorts first.
struct ColDigit {
  ReadAcc acc;
  const uint32_t* ids;
  uint32_t col0;
  int ncols;
  __device__ __forceinline__ int operator()(uint64_t k) const {
    const uint32_t id = ids[k];
    const uint32_t l = acc.rv.len[id];
    int d = 0;
    for (int q = 0; q < ncols; ++q) {
      const uint32_t col = col0 + q;
      const int s = col < l ? (int)acc.code(id, col, l) + 1 : 0;
      d = d * 5 + s;
    }
    return d;
  }
};
struct IdMove {
  const uint32_t* in;
  uint32_t* out;
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t dest) const { out[dest] = in[k]; }
};

// heads[k] = 1 when the k-th read in content order differs from its predecessor.
__global__ void read_heads_kernel(ReadsView rv, const uint32_t* __restrict__ ids, uint64_t n, uint64_t* __restrict__ heads) {
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
__global__ void assign_rank_kernel(const uint32_t* __restrict__ ids, const uint64_t* __restrict__ excl, uint64_t n,
                                   uint64_t total, uint32_t* __restrict__ rank) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint64_t next = k + 1 < n ? excl[k + 1] : total;
  rank[ids[k]] = (uint32_t)next;  // == excl[k] + heads[k]
}

__global__ void doubling_keys_kernel(const uint32_t* __restrict__ rank, uint64_t n, uint64_t h, int bits,
                                     uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t r2 = i + h < n ? rank[i + h] : 0;  // end of text is smallest
  keys[i] = ((uint64_t)rank[i] << bits) | r2;
  vals[i] = (uint32_t)i;
}

__global__ void key_heads_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ heads) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  heads[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}

// Sentinel order of model B from the read-start order `sai`: '$' of read i sorts by the suffix that follows it,
// i.e. by the rank of read i+1; the last read's '$' (end of text) is smallest.
__global__ void find_zero_kernel(const uint32_t* __restrict__ sai, uint64_t n, uint64_t* __restrict__ k0) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && sai[k] == 0) *k0 = k;
}
__global__ void sentinel_order_kernel(const uint32_t* __restrict__ sai, uint64_t n, const uint64_t* __restrict__ k0p,
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
                                uint32_t* __restrict__ ID, PosT* __restrict__ NP) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t k = 2 * i;
  if (k >= n) return;
  uint32_t byte = 0;
  for (int h = 0; h < 2; ++h) {
    if (k + h < n) {
      const uint32_t id = sord[k + h];
      const uint32_t l = acc.rv.len[id];
      byte |= (acc.code(id, l - 1, l) + 1) << (4 * h);
      ID[k + h] = id;
      NP[k + h] = (PosT)(k + h);
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

// Per-256-row counts of A,C,G,T of a nibble-packed symbol array: cnt[c*nb + b].  One warp per block, one word per lane.
__global__ void __launch_bounds__(256) sym_block_count_kernel(const uint8_t* __restrict__ B, uint64_t n, uint64_t nb,
                                                              uint64_t* __restrict__ cnt) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nb) return;
  const uint64_t row0 = warp * kBcrBlock + (uint64_t)lane * 8;
  uint32_t a = 0, c = 0, g = 0, t = 0;
  if (row0 < n) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(B + row0 / 2);  // the array is padded
    const uint64_t left = n - row0;
    const uint32_t keep = left >= 8 ? 0x11111111u : (0x11111111u & ((1u << (4 * (uint32_t)left)) - 1u));
    nib_count_add(w, keep, a, c, g, t);
  }
  uint32_t w1 = nib_sum(a) | (nib_sum(c) << 16), w2 = nib_sum(g) | (nib_sum(t) << 16);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    w1 += __shfl_xor_sync(0xffffffffu, w1, o);
    w2 += __shfl_xor_sync(0xffffffffu, w2, o);
  }
  if (lane == 0) {
    cnt[0 * nb + warp] = w1 & 0xffffu;
    cnt[1 * nb + warp] = w1 >> 16;
    cnt[2 * nb + warp] = w2 & 0xffffu;
    cnt[3 * nb + warp] = w2 >> 16;
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

// For every active read (in current row order): the symbol c that starts its next suffix and the row that suffix
// takes in the next BWT:  N + (#symbols < c in B) + (#c in B[0..P)).
//   scanned = exclusive scan of the flattened [4][nb] block counts, so scanned[c*nb + b] already equals
//   (#symbols < c in B) + (#c before block b).
template <class PosT>
__global__ void __launch_bounds__(256) bcr_rank_kernel(ReadAcc acc, const uint8_t* __restrict__ B, const uint64_t* __restrict__ scanned,
                                                       uint64_t nb, uint64_t n_reads, const uint32_t* __restrict__ ID,
                                                       const PosT* __restrict__ NP, uint64_t nact, uint32_t j,
                                                       uint8_t* __restrict__ tmpC, PosT* __restrict__ tmpPos) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nact) return;
  const uint32_t id = ID[k];
  const uint32_t l = acc.rv.len[id];
  if (l < j) {
    tmpC[k] = 255;  // finished in the previous pass: dropped
    return;
  }
  const uint32_t c = acc.code(id, l - j, l);
  const uint64_t P = (uint64_t)NP[k];
  const uint64_t b = P / kBcrBlock;
  const uint32_t r = (uint32_t)(P - b * kBcrBlock);  // rows of this block before P
  const uint32_t pat = (c + 1) * 0x11111111u;
  uint32_t inblk = 0;
  const uint4* v = reinterpret_cast<const uint4*>(B + b * (kBcrBlock / 2));  // 128 bytes per block
  const uint32_t nvec = (r + 31) / 32;
  for (uint32_t q = 0; q < nvec; ++q) {
    const uint4 x = v[q];
    const uint32_t left = r - q * 32;  // nibbles of this vector that count (may exceed 32)
    inblk += nib_eq_count(x.x, pat, left);
    if (left > 8) inblk += nib_eq_count(x.y, pat, left - 8);
    if (left > 16) inblk += nib_eq_count(x.z, pat, left - 16);
    if (left > 24) inblk += nib_eq_count(x.w, pat, left - 24);
  }
  tmpC[k] = (uint8_t)c;
  tmpPos[k] = (PosT)(n_reads + scanned[(uint64_t)c * nb + b] + inblk);
}

struct BcrDigit {
  const uint8_t* tmpC;
  __device__ __forceinline__ int operator()(uint64_t k) const { return (int)tmpC[k]; }
};
template <class PosT>
struct BcrMove {
  ReadAcc acc;
  const uint32_t* ID;
  const PosT* tmpPos;
  uint32_t j;
  uint32_t* ID2;
  PosT* NP2;
  uint8_t* SYM2;
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t dest) const {
    const uint32_t id = ID[k];
    const uint32_t l = acc.rv.len[id];
    ID2[dest] = id;
    NP2[dest] = tmpPos[k];
    SYM2[dest] = (l == j) ? (uint8_t)0 : (uint8_t)(acc.code(id, l - j - 1, l) + 1);
  }
};

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
// SYM) into Bn (n_new rows), plus the per-256-row symbol counts of Bn.  HBM-bound by design: a CTA produces 8192 rows
// (4 KB), a thread 32 rows (one 16-byte store).  The old rows of the tile are staged in shared memory; a thread
// first takes the 32 old nibbles that would land on its rows if it had no inserts (five shared-memory words and
// four funnel shifts), then opens a 4-bit gap for each of its inserts (the bits of its word of the tile's insert
// mask), which pushes the surplus old nibbles out at the top -- exactly the ones the next thread starts with.
template <class PosT>
__global__ void __launch_bounds__(kMergeThreads) bcr_merge_kernel(const uint8_t* __restrict__ Bo, uint64_t n_old,
                                                                  const PosT* __restrict__ NP, const uint8_t* __restrict__ SYM,
                                                                  uint64_t nact, const uint64_t* __restrict__ tile_k0,
                                                                  uint8_t* __restrict__ Bn, uint64_t n_new, uint64_t nb,
                                                                  uint64_t* __restrict__ cnt) {
  __shared__ __align__(16) uint32_t oldw[kMergeTile / 8 + 16];
  __shared__ uint32_t mask[kMergeTile / 32];
  __shared__ uint64_t sh[9];
  const uint64_t t0 = (uint64_t)blockIdx.x * kMergeTile;
  const uint64_t t1 = min(t0 + (uint64_t)kMergeTile, n_new);
  const uint64_t k0 = tile_k0[blockIdx.x], k1 = tile_k0[blockIdx.x + 1];
  const uint64_t a0 = t0 - k0;  // first old row of this tile
  const uint32_t nold = (uint32_t)((t1 - t0) - (k1 - k0));
  mask[threadIdx.x] = 0;
  // old rows -> shared, 16-byte vectors from the 32-row-aligned position at or below a0 (the arrays are padded)
  const uint64_t a0al = a0 & ~(uint64_t)31;
  const uint32_t shift = (uint32_t)(a0 - a0al);  // nibbles
  const uint32_t nvec = (nold + shift + 31) / 32;
  for (uint32_t q = threadIdx.x; q < nvec; q += kMergeThreads)
    reinterpret_cast<uint4*>(oldw)[q] = *reinterpret_cast<const uint4*>(Bo + a0al / 2 + (uint64_t)q * 16);
  __syncthreads();
  for (uint64_t k = k0 + threadIdx.x; k < k1; k += kMergeThreads) {
    const uint32_t r = (uint32_t)((uint64_t)NP[k] - t0);
    atomicOr(&mask[r >> 5], 1u << (r & 31));
  }
  __syncthreads();
  const uint32_t m = mask[threadIdx.x];
  uint64_t total;
  const uint32_t pre = (uint32_t)block_excl_scan((uint64_t)__popc(m), sh, total);  // inserts before my first row
  // the 32 old nibbles starting at the one that lands on my first row
  const uint32_t o = shift + threadIdx.x * 32 - pre;
  const uint32_t wi = o >> 3, sft = 4 * (o & 7);
  uint32_t out[4];
  {
    const uint32_t w0 = oldw[wi], w1 = oldw[wi + 1], w2 = oldw[wi + 2], w3 = oldw[wi + 3], w4 = oldw[wi + 4];
    out[0] = __funnelshift_r(w0, w1, sft);
    out[1] = __funnelshift_r(w1, w2, sft);
    out[2] = __funnelshift_r(w2, w3, sft);
    out[3] = __funnelshift_r(w3, w4, sft);
  }
  // open a gap for every insert, lowest row first
  uint32_t mm = m;
  uint64_t kk = k0 + pre;
  while (mm) {
    const uint32_t p = (uint32_t)__ffs(mm) - 1;
    mm &= mm - 1;
    const uint32_t s = SYM[kk++];
    const uint32_t wq = p >> 3, bit = 4 * (p & 7);
#pragma unroll
    for (int i = 3; i >= 1; --i)
      if ((uint32_t)i > wq) out[i] = (out[i] << 4) | (out[i - 1] >> 28);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if ((uint32_t)i == wq) {
        const uint32_t lowm = (1u << bit) - 1u;
        const uint32_t up = bit < 28 ? ((out[i] >> bit) << (bit + 4)) : 0u;
        out[i] = (out[i] & lowm) | (s << bit) | up;
      }
    }
  }
  // rows past the end of the array stay zero
  const uint64_t row0 = t0 + (uint64_t)threadIdx.x * 32;
  uint32_t valid = row0 >= t1 ? 0u : (uint32_t)min((uint64_t)32, t1 - row0);
  uint32_t accA = 0, accC = 0, accG = 0, accT = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t vw = valid >= 8u * (i + 1) ? 8u : (valid > 8u * i ? valid - 8u * i : 0u);
    const uint32_t km = vw >= 8 ? 0xffffffffu : ((1u << (4 * vw)) - 1u);
    out[i] &= km;
    nib_count_add(out[i], 0x11111111u, accA, accC, accG, accT);
  }
  if (valid) *reinterpret_cast<uint4*>(Bn + row0 / 2) = make_uint4(out[0], out[1], out[2], out[3]);
  // 256-row block = 8 consecutive threads
  uint32_t w1 = nib_sum(accA) | (nib_sum(accC) << 16), w2 = nib_sum(accG) | (nib_sum(accT) << 16);
#pragma unroll
  for (int oo = 4; oo > 0; oo >>= 1) {
    w1 += __shfl_xor_sync(0xffffffffu, w1, oo);
    w2 += __shfl_xor_sync(0xffffffffu, w2, oo);
  }
  if ((threadIdx.x & 7) == 0) {
    const uint64_t b = row0 / kBcrBlock;
    if (b < nb) {
      cnt[0 * nb + b] = w1 & 0xffffu;
      cnt[1 * nb + b] = w1 >> 16;
      cnt[2 * nb + b] = w2 & 0xffffu;
      cnt[3 * nb + b] = w2 >> 16;
    }
  }
}

// ---- 3. run-length encoding ------------------------------------------------------------------------------------------

struct HeadDigit {
  const uint8_t* B;
  __device__ __forceinline__ int operator()(uint64_t k) const {
    return (k == 0 || nib_get(B, k) != nib_get(B, k - 1)) ? 0 : 1;
  }
};
struct HeadOnlyMove {
  uint64_t* heads;
  uint64_t n_heads;
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t dest) const {
    if (dest < n_heads) heads[dest] = k;  // bucket 1 (non-heads) lands at dest >= n_heads and is dropped
  }
};

// bytes needed by run r: ceil(len / 31)   (src/bwt.cpp:16-27: a run is cut whenever its count reaches 31)
__global__ void run_bytes_kernel(const uint64_t* __restrict__ heads, uint64_t n_runs, uint64_t n, uint64_t* __restrict__ nbytes) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const uint64_t len = (r + 1 < n_runs ? heads[r + 1] : n) - heads[r];
  nbytes[r] = (len + 30) / 31;
}

__global__ void run_emit_kernel(const uint8_t* __restrict__ B, const uint64_t* __restrict__ heads, const uint64_t* __restrict__ boff,
                                uint64_t n_runs, uint64_t n, uint8_t* __restrict__ rle) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const uint64_t start = heads[r];
  uint64_t len = (r + 1 < n_runs ? heads[r + 1] : n) - start;
  const uint8_t s = (uint8_t)(nib_get(B, start) << 5);
  uint64_t o = boff[r];
  while (len > 0) {
    const uint32_t c = len >= 31 ? 31u : (uint32_t)len;
    rle[o++] = s | (uint8_t)c;
    len -= c;
  }
}

// histogram of read lengths clipped to lmax: hist[l] += 1
__global__ void len_hist_kernel(const uint32_t* __restrict__ len, uint64_t n, uint32_t* __restrict__ hist) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&hist[len[i]], 1u);
}

// ---- driver -----------------------------------------------------------------------------------------------------------

template <class PosT>
static void bcr_passes(sgx_ctx* ctx, const sgx_reads* reads, const ReadAcc& acc, const uint32_t* d_sord,
                       const std::vector<uint64_t>& nact_by_pass, uint64_t S, uint8_t* B0, uint8_t* B1,
                       uint8_t** final_B, sgx_index_out* out, KernelClock& k_merge) {
  const uint64_t N = reads->n;
  const uint32_t lmax = reads->max_len;
  const uint64_t nb_max = S / kBcrBlock + 2;
  DevBuf<uint32_t> ID0(N), ID1(N);
  DevBuf<PosT> NP0(N), NP1(N), tmpPos(N);
  DevBuf<uint8_t> SYM(N), tmpC(N);
  DevBuf<uint64_t> cnt(4 * nb_max), scan_ws(scan_ws_elems(4 * nb_max)), part_ws(part_ws_elems(N, 4)),
      tile_k0(S / kMergeTile + 4);
  uint32_t* ID = ID0.p;
  uint32_t* ID2 = ID1.p;
  PosT* NP = NP0.p;
  PosT* NP2 = NP1.p;
  uint8_t* B = B0;
  uint8_t* Bn = B1;

  uint64_t n_cur = N;
  SGX_LAUNCH(ctx, (bcr_init_kernel<PosT>), grid_for((N + 1) / 2, 256), 256, 0, acc, d_sord, N, B, ID, NP);
  uint64_t nb = n_cur / kBcrBlock + 1;
  SGX_LAUNCH(ctx, sym_block_count_kernel, grid_for(nb * 32, 256), 256, 0, (const uint8_t*)B, n_cur, nb, cnt.p);
  scan_exclusive(ctx, cnt.p, 4 * nb, scan_ws.p, (uint64_t*)nullptr);

  uint64_t nact_prev = N;  // items currently held in ID/NP (inserted by the previous pass)
  for (uint32_t j = 1; j <= lmax; ++j) {
    const uint64_t nact = nact_by_pass[j];  // reads with length >= j
    if (nact == 0) break;
    SGX_LAUNCH(ctx, (bcr_rank_kernel<PosT>), grid_for(nact_prev, 256), 256, 0, acc, (const uint8_t*)B,
               (const uint64_t*)cnt.p, nb, N, (const uint32_t*)ID, (const PosT*)NP, nact_prev, j, tmpC.p, tmpPos.p);
    BcrDigit dg{tmpC.p};
    BcrMove<PosT> mv{acc, ID, tmpPos.p, j, ID2, NP2, SYM.p};
    partition_pass<4>(ctx, dg, mv, nact_prev, part_ws.p, (uint64_t*)nullptr);
    const uint64_t n_new = n_cur + nact;
    const uint64_t ntiles = (n_new + kMergeTile - 1) / kMergeTile;
    SGX_LAUNCH(ctx, (bcr_tile_bounds_kernel<PosT>), grid_for(nact + 1, 256), 256, 0, (const PosT*)NP2, nact, ntiles,
               tile_k0.p);
    const uint64_t nb_new = n_new / kBcrBlock + 1;
    // the block past the last full one may not be touched by the merge when n_new is a multiple of 256
    SGX_CUDA(cudaMemsetAsync(cnt.p, 0, 4 * nb_new * sizeof(uint64_t), ctx->stream));
    k_merge.tick();
    SGX_LAUNCH(ctx, (bcr_merge_kernel<PosT>), (unsigned)ntiles, kMergeThreads, 0, (const uint8_t*)B, n_cur,
               (const PosT*)NP2, (const uint8_t*)SYM.p, nact, (const uint64_t*)tile_k0.p, Bn, n_new, nb_new, cnt.p);
    k_merge.tick();
    out->merge_launches++;
    out->merge_bytes += (n_cur + n_new) / 2 + nact * (sizeof(PosT) + 1);
    scan_exclusive(ctx, cnt.p, 4 * nb_new, scan_ws.p, (uint64_t*)nullptr);
    std::swap(B, Bn);
    std::swap(ID, ID2);
    std::swap(NP, NP2);
    n_cur = n_new;
    nb = nb_new;
    nact_prev = nact;
  }
  SGX_REQUIRE(n_cur == S, SGX_ECUDA, "index: column passes did not produce n_suffixes rows");
  *final_B = B;
}

void index_build(sgx_ctx* ctx, const sgx_reads* reads, int order, bool reversed, sgx_index_out* out) {
  memset(out, 0, sizeof(*out));
  const uint64_t N = reads->n;
  const uint64_t S = reads->bases + N;
  const uint32_t lmax = reads->max_len;
  const uint64_t launches0 = ctx->launches;
  EventTimer t_all(ctx->stream), t_sort(ctx->stream), t_bcr(ctx->stream), t_emit(ctx->stream);
  KernelClock k_merge(ctx->stream);
  ReadAcc acc;
  acc.rv = reads->view();
  acc.reversed = reversed ? 1 : 0;

  t_all.start();
  t_sort.start();
  // ---- 1. read ordering
  DevBuf<uint32_t> idsA(N), idsB(N), d_sai(N), d_sord(N);
  DevBuf<uint64_t> part_ws(part_ws_elems(N, 256)), heads(N + 1), scan_ws(scan_ws_elems(N + 1)), d_total(1);
  SGX_LAUNCH(ctx, iota_kernel, grid_for(N, 256), 256, 0, idsA.p, N);
  uint32_t* ids = idsA.p;
  uint32_t* ids2 = idsB.p;
  for (uint32_t c = lmax; c > 0;) {
    const int ncols = c >= 3 ? 3 : (int)c;
    const uint32_t col0 = c - ncols;
    ColDigit dg{acc, ids, col0, ncols};
    IdMove mv{ids, ids2};
    partition_pass<125>(ctx, dg, mv, N, part_ws.p, (uint64_t*)nullptr);
    std::swap(ids, ids2);
    c = col0;
  }
  if (order == SGX_ORDER_SAIS) {
    // model A: equal suffixes tie by read index; the content sort is stable, so `ids` is the read-start order and
    // the sentinel of read i is simply the i-th smallest.
    SGX_CUDA(cudaMemcpyAsync(d_sai.p, ids, N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    SGX_LAUNCH(ctx, iota_kernel, grid_for(N, 256), 256, 0, d_sord.p, N);
  } else {
    // model B: rank reads by (content, rank of the following read, ...): prefix doubling over the read sequence.
    DevBuf<uint32_t> rank(N), vals(N), vals2(N);
    DevBuf<uint64_t> keys(N), keys2(N);
    SGX_LAUNCH(ctx, read_heads_kernel, grid_for(N, 256), 256, 0, acc.rv, (const uint32_t*)ids, N, heads.p);
    scan_exclusive(ctx, heads.p, N, scan_ws.p, d_total.p);
    uint64_t distinct = 0;
    SGX_CUDA(cudaMemcpyAsync(&distinct, d_total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    SGX_LAUNCH(ctx, assign_rank_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)ids, (const uint64_t*)heads.p, N,
               distinct, rank.p);
    const uint32_t* order_ids = ids;
    int bits = 1;
    while (((uint64_t)1 << bits) <= N) ++bits;  // ranks are in [0, N]
    for (uint64_t h = 1; distinct < N; h <<= 1) {
      SGX_LAUNCH(ctx, doubling_keys_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)rank.p, N, h, bits, keys.p, vals.p);
      radix_sort_pairs(ctx, keys.p, vals.p, keys2.p, vals2.p, N, 2 * bits, part_ws.p);
      SGX_LAUNCH(ctx, key_heads_kernel, grid_for(N, 256), 256, 0, (const uint64_t*)keys.p, N, heads.p);
      scan_exclusive(ctx, heads.p, N, scan_ws.p, d_total.p);
      SGX_CUDA(cudaMemcpyAsync(&distinct, d_total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
      SGX_CUDA(cudaStreamSynchronize(ctx->stream));
      SGX_LAUNCH(ctx, assign_rank_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)vals.p, (const uint64_t*)heads.p, N,
                 distinct, rank.p);
      order_ids = vals.p;
      SGX_REQUIRE(h < ((uint64_t)1 << 40), SGX_ECUDA, "index: read-rank doubling did not converge");
    }
    SGX_CUDA(cudaMemcpyAsync(d_sai.p, order_ids, N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    SGX_LAUNCH(ctx, find_zero_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)d_sai.p, N, d_total.p);
    SGX_LAUNCH(ctx, sentinel_order_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)d_sai.p, N,
               (const uint64_t*)d_total.p, d_sord.p);
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));  // keys/vals go out of scope
  }
  t_sort.stop();

  // ---- 2. column passes
  t_bcr.start();
  // reads still active at pass j = reads with length >= j
  std::vector<uint64_t> nact(lmax + 2, 0);
  {
    DevBuf<uint32_t> hist(lmax + 1);
    SGX_CUDA(cudaMemsetAsync(hist.p, 0, (lmax + 1) * sizeof(uint32_t), ctx->stream));
    SGX_LAUNCH(ctx, len_hist_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)reads->len, N, hist.p);
    std::vector<uint32_t> h(lmax + 1);
    SGX_CUDA(cudaMemcpyAsync(h.data(), hist.p, (lmax + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t run = 0;
    for (uint32_t l = lmax; l >= 1; --l) {
      run += h[l];
      nact[l] = run;
    }
  }
  // bytes of a 4-bit array of S rows, padded by a tile and a vector so that whole-vector accesses stay in bounds
  const uint64_t padded = (((S + kMergeTile - 1) / kMergeTile + 2) * kMergeTile) / 2 + 128;
  DevBuf<uint8_t> B0(padded), B1(padded);
  SGX_CUDA(cudaMemsetAsync(B0.p, 0, padded, ctx->stream));
  SGX_CUDA(cudaMemsetAsync(B1.p, 0, padded, ctx->stream));
  uint8_t* Bfinal = nullptr;
  // SGX_FORCE_WIDE=1 (tests): take the 64-bit position path on inputs that would fit 32 bits
  const char* fw = getenv("SGX_FORCE_WIDE");
  if (S < 0xffffffffull && !(fw && fw[0] == '1')) {
    bcr_passes<uint32_t>(ctx, reads, acc, d_sord.p, nact, S, B0.p, B1.p, &Bfinal, out, k_merge);
  } else {
    bcr_passes<uint64_t>(ctx, reads, acc, d_sord.p, nact, S, B0.p, B1.p, &Bfinal, out, k_merge);
  }
  t_bcr.stop();

  // ---- 3. run-length encoding + copy out
  t_emit.start();
  uint64_t n_runs = 0, n_bytes = 0;
  DevBuf<uint8_t> d_rle;
  {
    DevBuf<uint64_t> ws_h(part_ws_elems(S, 2));
    // run heads: positions k with B[k] != B[k-1]
    // (count first so that the heads array is sized exactly)
    const uint32_t ntiles = part_tiles(S);
    HeadDigit hd{Bfinal};
    SGX_LAUNCH(ctx, (part_count_kernel<2, HeadDigit>), ntiles, kPartThreads, 0, hd, S, ws_h.p, ntiles);
    // bucket 0 = heads: total heads = sum of the first ntiles counters
    DevBuf<uint64_t> ws_s(scan_ws_elems((uint64_t)2 * ntiles));
    scan_exclusive(ctx, ws_h.p, (uint64_t)2 * ntiles, ws_s.p, (uint64_t*)nullptr);
    SGX_CUDA(cudaMemcpyAsync(&n_runs, ws_h.p + ntiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    DevBuf<uint64_t> d_heads(n_runs + 1), d_boff(n_runs + 1), ws_r(scan_ws_elems(n_runs + 1));
    HeadOnlyMove hm{d_heads.p, n_runs};
    SGX_LAUNCH(ctx, (part_scatter_kernel<2, HeadDigit, HeadOnlyMove>), ntiles, kPartThreads, 0, hd, hm, S,
               (const uint64_t*)ws_h.p, ntiles);
    SGX_CUDA(cudaMemsetAsync(d_boff.p, 0, (n_runs + 1) * sizeof(uint64_t), ctx->stream));
    SGX_LAUNCH(ctx, run_bytes_kernel, grid_for(n_runs, 256), 256, 0, (const uint64_t*)d_heads.p, n_runs, S, d_boff.p);
    scan_exclusive(ctx, d_boff.p, n_runs + 1, ws_r.p, (uint64_t*)nullptr);
    SGX_CUDA(cudaMemcpyAsync(&n_bytes, d_boff.p + n_runs, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    d_rle.alloc(n_bytes ? n_bytes : 1);
    SGX_LAUNCH(ctx, run_emit_kernel, grid_for(n_runs, 256), 256, 0, (const uint8_t*)Bfinal, (const uint64_t*)d_heads.p,
               (const uint64_t*)d_boff.p, n_runs, S, d_rle.p);
  }
  out->n_strings = N;
  out->n_suffixes = S;
  out->n_runs = n_bytes;
  out->rle = (uint8_t*)pinned_get(ctx, n_bytes ? n_bytes : 1);
  out->sai = (uint32_t*)pinned_get(ctx, N * sizeof(uint32_t));
  SGX_CUDA(cudaMemcpyAsync(out->rle, d_rle.p, n_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaMemcpyAsync(out->sai, d_sai.p, N * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  t_emit.stop();
  t_all.stop();
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  out->ms_total = t_all.ms();
  out->ms_sort = t_sort.ms();
  out->ms_bcr = t_bcr.ms();
  out->ms_emit = t_emit.ms();
  out->ms_merge = k_merge.total_ms();
  out->launches = ctx->launches - launches0;
}

}  // namespace sgx
