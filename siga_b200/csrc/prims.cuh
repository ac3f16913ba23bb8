// Hand-written device-wide primitives used by the index builder and the FM / hits packers:
//   * in-place exclusive scan of a uint64 array (three-phase, recursive over tile sums)
//   * stable multi-way partition pass (count -> scan -> scatter) with arbitrary digit / move functors; one pass of
//     it is a BCR column step, eight-bit passes of it are an LSD radix sort.
// No CUB/Thrust: the passes are fused with their producers/consumers through the functors.
#ifndef SGX_PRIMS_CUH_
#define SGX_PRIMS_CUH_

#include "sgx_device.cuh"

namespace sgx {

static const int kScanThreads = 256;
static const int kScanIpt = 8;
static const int kScanTile = kScanThreads * kScanIpt;  // 2048

// Exclusive scan of one value per thread across a 256-thread block; returns the exclusive prefix, `total` = block sum.
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t* sh /* >= 9 entries */, uint64_t& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint64_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    uint64_t s = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0;
    uint64_t t = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    if (lane < (int)(blockDim.x >> 5)) sh[lane] = t - s;  // exclusive warp offsets
    if (lane == 31) sh[8] = t;
  }
  __syncthreads();
  const uint64_t off = sh[w];
  total = sh[8];
  __syncthreads();
  return off + x - v;
}

// 32-bit variant for per-thread counts whose block total fits 32 bits (two barriers, one shuffle per step).
__device__ __forceinline__ uint32_t block_excl_scan32(uint32_t v, uint32_t* sh /* >= 8 entries */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[w] = x;
  __syncthreads();
  uint32_t off = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < w) off += sh[q];  // 8 warps: a short serial sum beats a second shuffle scan + barrier
  return off + x - v;
}

static __global__ void scan_tile_reduce_kernel(const uint64_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ tilesum) {
  __shared__ uint64_t sh[9];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanIpt;
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanIpt; ++i)
    if (base + i < n) s += in[base + i];
  uint64_t total;
  block_excl_scan(s, sh, total);
  if (threadIdx.x == 0) tilesum[blockIdx.x] = total;
}

// In-place exclusive scan inside each tile, offset by tileoff[tile] (or 0 when tileoff == nullptr).
static __global__ void scan_tile_apply_kernel(uint64_t* __restrict__ data, uint64_t n, const uint64_t* __restrict__ tileoff,
                                       uint64_t* __restrict__ total_out) {
  __shared__ uint64_t sh[9];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanIpt;
  uint64_t v[kScanIpt];
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanIpt; ++i) {
    v[i] = base + i < n ? data[base + i] : 0;
    s += v[i];
  }
  uint64_t total;
  uint64_t run = block_excl_scan(s, sh, total) + (tileoff ? tileoff[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < kScanIpt; ++i) {
    if (base + i < n) data[base + i] = run;
    run += v[i];
  }
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *total_out = run;
}

// Same, with the tile offset summed by the CTA itself from the tile sums (no scan of the tile sums in between): two
// launches instead of three for arrays of up to kScanDirectTiles tiles, which is every scan of a column pass.
static const uint64_t kScanDirectTiles = 4096;
static __global__ void scan_tile_apply_direct_kernel(uint64_t* __restrict__ data, uint64_t n, const uint64_t* __restrict__ tilesum,
                                                     uint64_t* __restrict__ total_out) {
  __shared__ uint64_t sh[9];
  uint64_t before = 0;
  for (uint32_t t = threadIdx.x; t < blockIdx.x; t += kScanThreads) before += tilesum[t];
  uint64_t off_total;
  block_excl_scan(before, sh, off_total);  // off_total = sum of the sums of all earlier tiles
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanIpt;
  uint64_t v[kScanIpt];
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanIpt; ++i) {
    v[i] = base + i < n ? data[base + i] : 0;
    s += v[i];
  }
  uint64_t total;
  uint64_t run = block_excl_scan(s, sh, total) + off_total;
#pragma unroll
  for (int i = 0; i < kScanIpt; ++i) {
    if (base + i < n) data[base + i] = run;
    run += v[i];
  }
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *total_out = run;
}

// Whole array in one CTA of 1024 threads (n <= kScanSmall): a thread sums its contiguous stretch (at most 16 values, all
// loads issued before the first add), the stretch sums are scanned across the block (warp shuffles, the 32 warp sums by
// the first warp), the stretch is rewritten.  One launch instead of two for the short counter arrays of a column pass.
static const uint64_t kScanSmall = 16384;
static const int kScanSmallThreads = 1024;
static const int kScanSmallIpt = (int)(kScanSmall / kScanSmallThreads);  // 16
static __global__ void __launch_bounds__(kScanSmallThreads) scan_small_kernel(uint64_t* __restrict__ data, uint64_t n,
                                                                             uint64_t* __restrict__ total_out) {
  __shared__ uint64_t wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t ipt = (uint32_t)((n + kScanSmallThreads - 1) / kScanSmallThreads);  // <= kScanSmallIpt
  const uint64_t b = (uint64_t)threadIdx.x * ipt;
  uint64_t v[kScanSmallIpt];
#pragma unroll
  for (int i = 0; i < kScanSmallIpt; ++i) v[i] = ((uint32_t)i < ipt && b + i < n) ? data[b + i] : 0;
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanSmallIpt; ++i) s += v[i];
  uint64_t x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  if (w == 0) {
    const uint64_t t = wsum[lane];
    uint64_t z = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t y = __shfl_up_sync(0xffffffffu, z, o);
      if (lane >= o) z += y;
    }
    wsum[lane] = z - t;  // exclusive warp offsets
    if (lane == 31 && total_out) *total_out = z;
  }
  __syncthreads();
  uint64_t run = wsum[w] + x - s;
#pragma unroll
  for (int i = 0; i < kScanSmallIpt; ++i) {
    if ((uint32_t)i < ipt && b + i < n) data[b + i] = run;
    run += v[i];
  }
}

// Number of uint64 workspace elements scan_exclusive needs for an n-element array.
inline uint64_t scan_ws_elems(uint64_t n) {
  uint64_t ws = 0;
  while (n > (uint64_t)kScanTile) {
    n = (n + kScanTile - 1) / kScanTile;
    ws += n;
  }
  return ws + 1;
}

// data[i] <- sum of data[0..i) ; *total_dev (device pointer, optional) <- sum of all.
inline void scan_exclusive(sgx_ctx* ctx, uint64_t* data, uint64_t n, uint64_t* ws, uint64_t* total_dev) {
  if (n == 0) {
    if (total_dev) SGX_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(uint64_t), ctx->stream));
    return;
  }
  const uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
  if (ntiles == 1) {
    SGX_LAUNCH(ctx, scan_tile_apply_kernel, 1, kScanThreads, 0, data, n, (const uint64_t*)nullptr, total_dev);
    return;
  }
  if (n <= kScanSmall) {
    SGX_LAUNCH(ctx, scan_small_kernel, 1, kScanSmallThreads, 0, data, n, total_dev);
    return;
  }
  SGX_LAUNCH(ctx, scan_tile_reduce_kernel, (unsigned)ntiles, kScanThreads, 0, data, n, ws);
  if (ntiles <= kScanDirectTiles) {
    SGX_LAUNCH(ctx, scan_tile_apply_direct_kernel, (unsigned)ntiles, kScanThreads, 0, data, n, (const uint64_t*)ws, total_dev);
    return;
  }
  scan_exclusive(ctx, ws, ntiles, ws + ntiles, total_dev);
  SGX_LAUNCH(ctx, scan_tile_apply_kernel, (unsigned)ntiles, kScanThreads, 0, data, n, (const uint64_t*)ws,
             (uint64_t*)nullptr);
}

// ---------------------------------------------------------------------------------------------------------------
// Stable partition pass.
//   digit(k)  -> bucket of item k in [0, RADIX) ; items whose bucket is >= RADIX are dropped
//   move(k, dest) moves item k to slot dest of the output
// Output order: bucket-major, and inside a bucket the input order (stable).
// Tile = 8 warps x 8 rounds x 32 lanes = 2048 consecutive items; lanes read consecutive items (coalesced).
// ---------------------------------------------------------------------------------------------------------------
static const int kPartThreads = 256;
static const int kPartRounds = 8;
static const int kPartTile = kPartThreads * kPartRounds;

template <int RADIX, class DigitF>
__global__ void __launch_bounds__(kPartThreads) part_count_kernel(DigitF digit, uint64_t n, uint64_t* __restrict__ tilecnt,
                                                                  uint32_t ntiles) {
  __shared__ uint32_t wc[kPartThreads / 32][RADIX];
  for (int i = threadIdx.x; i < (kPartThreads / 32) * RADIX; i += kPartThreads) (&wc[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t wbase = (uint64_t)blockIdx.x * kPartTile + (uint64_t)w * (kPartRounds * 32);
#pragma unroll 1
  for (int r = 0; r < kPartRounds; ++r) {
    const uint64_t k = wbase + r * 32 + lane;
    int d = RADIX;
    if (k < n) d = digit(k);
    const unsigned act = __ballot_sync(0xffffffffu, d < RADIX);
    if (d < RADIX) {
      const unsigned m = __match_any_sync(act, d);
      if (lane == __ffs(m) - 1) wc[w][d] += __popc(m);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RADIX; d += kPartThreads) {
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < kPartThreads / 32; ++q) s += wc[q][d];
    tilecnt[(uint64_t)d * ntiles + blockIdx.x] = s;
  }
}

template <int RADIX, class DigitF, class MoveF>
__global__ void __launch_bounds__(kPartThreads) part_scatter_kernel(DigitF digit, MoveF move, uint64_t n,
                                                                    const uint64_t* __restrict__ tileoff, uint32_t ntiles) {
  __shared__ uint32_t wc[kPartThreads / 32][RADIX];
  for (int i = threadIdx.x; i < (kPartThreads / 32) * RADIX; i += kPartThreads) (&wc[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t wbase = (uint64_t)blockIdx.x * kPartTile + (uint64_t)w * (kPartRounds * 32);
  int dg[kPartRounds];
  uint32_t lr[kPartRounds];
#pragma unroll
  for (int r = 0; r < kPartRounds; ++r) {
    const uint64_t k = wbase + r * 32 + lane;
    int d = RADIX;
    if (k < n) d = digit(k);
    dg[r] = d;
    lr[r] = 0;
    const unsigned act = __ballot_sync(0xffffffffu, d < RADIX);
    if (d < RADIX) {
      const unsigned m = __match_any_sync(act, d);
      const uint32_t before = wc[w][d];
      lr[r] = before + __popc(m & ((1u << lane) - 1u));
      __syncwarp(m);
      if (lane == __ffs(m) - 1) wc[w][d] = before + __popc(m);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RADIX; d += kPartThreads) {
    uint32_t run = 0;
#pragma unroll
    for (int q = 0; q < kPartThreads / 32; ++q) {
      const uint32_t t = wc[q][d];
      wc[q][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kPartRounds; ++r) {
    const int d = dg[r];
    if (d < RADIX) {
      const uint64_t k = wbase + r * 32 + lane;
      const uint64_t dest = tileoff[(uint64_t)d * ntiles + blockIdx.x] + wc[w][d] + lr[r];
      move(k, dest);
    }
  }
}

inline uint32_t part_tiles(uint64_t n) { return (uint32_t)((n + kPartTile - 1) / kPartTile); }
// workspace (uint64 elements): RADIX*ntiles counters + scan workspace
inline uint64_t part_ws_elems(uint64_t n, int radix) {
  const uint64_t c = (uint64_t)radix * part_tiles(n ? n : 1);
  return c + scan_ws_elems(c) + 1;
}

// Runs one stable partition pass.  bucket_starts_dev (optional, device, RADIX+... not provided): callers that need
// the bucket sizes read tilecnt afterwards: after the call ws[d*ntiles] is the start offset of bucket d.
// The two halves of a pass, for callers that want the bucket sizes (ws[d * ntiles] = start of bucket d after the
// scan, *total_dev = items kept) on their way to the host while the scatter is still running.
template <int RADIX, class DigitF>
inline void partition_count(sgx_ctx* ctx, DigitF digit, uint64_t n, uint64_t* ws, uint64_t* total_dev) {
  const uint32_t ntiles = part_tiles(n);
  const uint64_t c = (uint64_t)RADIX * ntiles;
  SGX_LAUNCH(ctx, (part_count_kernel<RADIX, DigitF>), ntiles, kPartThreads, 0, digit, n, ws, ntiles);
  scan_exclusive(ctx, ws, c, ws + c, total_dev);
}
template <int RADIX, class DigitF, class MoveF>
inline void partition_scatter(sgx_ctx* ctx, DigitF digit, MoveF move, uint64_t n, const uint64_t* ws) {
  const uint32_t ntiles = part_tiles(n);
  SGX_LAUNCH(ctx, (part_scatter_kernel<RADIX, DigitF, MoveF>), ntiles, kPartThreads, 0, digit, move, n, ws, ntiles);
}

template <int RADIX, class DigitF, class MoveF>
inline void partition_pass(sgx_ctx* ctx, DigitF digit, MoveF move, uint64_t n, uint64_t* ws, uint64_t* total_dev) {
  if (n == 0) {
    if (total_dev) SGX_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(uint64_t), ctx->stream));
    return;
  }
  partition_count<RADIX, DigitF>(ctx, digit, n, ws, total_dev);
  partition_scatter<RADIX, DigitF, MoveF>(ctx, digit, move, n, ws);
}

// ---------------------------------------------------------------------------------------------------------------
// LSD radix sort of (uint64 key, uint32 value) pairs on bits [0, nbits) of the key, 8 bits per pass.
// Result ends in (keys, vals) ; (keys_alt, vals_alt) is scratch of the same size.
// ---------------------------------------------------------------------------------------------------------------
struct RadixDigit {
  const uint64_t* keys;
  int shift;
  __device__ __forceinline__ int operator()(uint64_t k) const { return (int)((keys[k] >> shift) & 255u); }
};
struct RadixMove {
  const uint64_t* kin;
  const uint32_t* vin;
  uint64_t* kout;
  uint32_t* vout;
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t dest) const {
    kout[dest] = kin[k];
    vout[dest] = vin[k];
  }
};

inline void radix_sort_pairs(sgx_ctx* ctx, uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt,
                             uint64_t n, int nbits, uint64_t* ws) {
  int passes = (nbits + 7) / 8;
  if (passes & 1) ++passes;  // even number of passes so that the result lands in (keys, vals)
  uint64_t* kin = keys;
  uint32_t* vin = vals;
  uint64_t* kout = keys_alt;
  uint32_t* vout = vals_alt;
  for (int p = 0; p < passes; ++p) {
    RadixDigit dg{kin, 8 * p};
    RadixMove mv{kin, vin, kout, vout};
    partition_pass<256>(ctx, dg, mv, n, ws, (uint64_t*)nullptr);
    uint64_t* tk = kin;
    kin = kout;
    kout = tk;
    uint32_t* tv = vin;
    vin = vout;
    vout = tv;
  }
}

}  // namespace sgx

#endif  // SGX_PRIMS_CUH_
