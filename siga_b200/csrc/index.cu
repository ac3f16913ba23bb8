// BWT / read-index construction on the device (`siga index`).
//
// Replaces SuffixArrayBuilder::build + BWT::BWT of the reference (src/suffix_array_builder.cpp:31-282 `sais`,
// 474-673 `sais2` via 3rdparty/sais.hxx; src/bwt.cpp:7-32).  The reference suffix-sorts the concatenated text on one
// CPU thread with SA-IS.  Here the BWT of the read collection is built directly, without a suffix array, by
// column insertion (the BCR idea the reference vendors in 3rdparty/bcr.c but never calls):
//
//   1. Read ordering (N items): an LSD radix sort over 32-base key chunks (index_content_sort) gives the content order;
//      dense content ranks + prefix doubling over the read sequence give the order of the read-start suffixes under
//      ordering model B (`sais2`: all '$' equal, comparison runs on into the next read).  This order IS the .sai
//      file, and it fixes the order of the N sentinel rows.  Model A (`sais`) orders ties by read index.
//   2. Column passes j = 1..Lmax: the suffixes of length j are inserted into the BWT of all shorter suffixes.  The
//      row of read i's new suffix is N + (#symbols < c before) + rank_c(row of its previous suffix); because that is
//      monotone in the previous row for a fixed c, ordering the inserts is ONE stable 4-way partition by c
//      (count -> scan -> scatter), and the insertion itself is a streaming merge of the old BWT with the sorted
//      inserts.  The merge is the dominant kernel: it reads and writes the partial BWT (4 bits per symbol) once
//      per column and rewrites two levels of rank checkpoints (index_kernels.cuh).  An item carries the symbol of its
//      row (= the first base of its next suffix) and the read word the next symbols come from, so the passes touch
//      the read store once per 16 columns.
//   3. Greedy run-length encoding (runs capped at 31) of the final symbol array into .bwt run bytes.
//
// Positions are 32-bit while the BWT has < 2^32 rows and 64-bit otherwise (PosT template parameter).
#include "index.cuh"

#include <algorithm>

#include "index_kernels.cuh"

namespace sgx {

// ---- driver -----------------------------------------------------------------------------------------------------------

template <class PosT>
static void bcr_passes(sgx_ctx* ctx, const sgx_reads* reads, const ReadAcc& acc, const uint32_t* d_sord,
                       const std::vector<uint64_t>& nact_by_pass, uint64_t S, uint8_t* B0, uint8_t* B1,
                       uint8_t** final_B, sgx_index_out* out, KernelClock& k_merge) {
  const uint64_t N = reads->n;
  const uint32_t lmax = reads->max_len;
  const uint32_t uniform_len = reads->min_len == lmax ? lmax : 0;
  const uint64_t nt_max = S / kMergeTile + 2;
  DevBuf<uint32_t> ID0(N), ID1(N);
  DevBuf<PosT> NP0(N), NP1(N);
  DevBuf<uint8_t> SYM0(N), SYM1(N);
  DevBuf<uint32_t> WIN0(N), WIN1(N);
  DevBuf<uint16_t> RL0(N), RL1(N);  // tile-local rank of every item's row (see tile_checkpoints)
  DevBuf<uint64_t> tiletot(4 * nt_max), scan_ws(scan_ws_elems(4 * nt_max)), tile_k0(S / kMergeTile + 4);
  // partition counters of the column passes, three buffers in rotation: pass j reads pw[j % 3] (counted by pass j - 1's
  // scatter, or by the count kernel for the first pass), its scatter accumulates pass j + 1's into pw[(j + 1) % 3] and
  // clears pw[(j + 2) % 3]
  const uint64_t pw_elems = part_ws_elems(N, 4);
  const uint32_t pw_counters = 4 * part_tiles(N ? N : 1);
  DevBuf<uint64_t> pw_all(3 * pw_elems);
  uint64_t* pw[3] = {pw_all.p, pw_all.p + pw_elems, pw_all.p + 2 * pw_elems};
  SGX_CUDA(cudaMemsetAsync(pw[2], 0, pw_counters * sizeof(uint64_t), ctx->stream));
  uint32_t* ID = ID0.p;
  uint32_t* ID2 = ID1.p;
  PosT* NP = NP0.p;
  PosT* NP2 = NP1.p;
  uint8_t* SYM = SYM0.p;
  uint8_t* SYM2 = SYM1.p;
  uint32_t* WIN = WIN0.p;
  uint32_t* WIN2 = WIN1.p;
  uint16_t* RL = RL0.p;
  uint16_t* RL2 = RL1.p;
  uint8_t* B = B0;
  uint8_t* Bn = B1;

  uint64_t n_cur = N;
  SGX_LAUNCH(ctx, (bcr_init_kernel<PosT>), grid_for((N + 1) / 2, 256), 256, 0, acc, d_sord, N, B, ID, NP, SYM);
  uint64_t nt = (n_cur + kMergeTile - 1) / kMergeTile;
  SGX_LAUNCH(ctx, tile_count_kernel, (unsigned)nt, 256, 0, (const uint8_t*)B, n_cur, nt, (uint2*)nullptr, tiletot.p, RL);
  scan_exclusive(ctx, tiletot.p, 4 * nt, scan_ws.p, (uint64_t*)nullptr);

  uint64_t nact_prev = N;  // items currently held in ID/NP/SYM (inserted by the previous pass)
  for (uint32_t j = 1; j <= lmax; ++j) {
    const uint64_t nact = nact_by_pass[j];  // reads with length >= j
    if (nact == 0) break;
    SymDigit dg{SYM};
    BcrMove<PosT> mv{acc, ID, WIN, j, uniform_len, ID2, NP2, SYM2, WIN2, tiletot.p, nt, N, SYM, NP, RL};
    if (nact_prev) {
      uint64_t* cur = pw[j % 3];
      const uint32_t pt = part_tiles(nact_prev);
      if (j == 1) {
        partition_count<4, SymDigit>(ctx, dg, nact_prev, cur, (uint64_t*)nullptr);
      } else {  // counted by the previous pass's scatter
        const uint64_t c = (uint64_t)4 * pt;
        scan_exclusive(ctx, cur, c, cur + c, (uint64_t*)nullptr);
      }
      const bool more = j < lmax && nact_by_pass[j + 1] > 0;
      SGX_LAUNCH(ctx, (bcr_scatter_kernel<PosT>), pt, kPartThreads, 0, mv, nact_prev, (const uint64_t*)cur, pt,
                 more ? (unsigned long long*)pw[(j + 1) % 3] : (unsigned long long*)nullptr, part_tiles(nact ? nact : 1),
                 pw[(j + 2) % 3], pw_counters);
    }
    const uint64_t n_new = n_cur + nact;
    const uint64_t ntiles = (n_new + kMergeTile - 1) / kMergeTile;
    SGX_LAUNCH(ctx, (bcr_tile_bounds_kernel<PosT>), grid_for(nact + 1, 256), 256, 0, (const PosT*)NP2, nact, ntiles,
               tile_k0.p);
    k_merge.tick();
    launch_merge<PosT>(ctx, (const uint8_t*)B, n_cur, (const PosT*)NP2, (const uint8_t*)SYM2, nact, (const uint64_t*)tile_k0.p, Bn,
                       n_new, ntiles, tiletot.p, RL2);
    k_merge.tick();
    out->merge_launches++;
    out->merge_bytes += (n_cur + n_new) / 2 + nact * (sizeof(PosT) + 1);
    scan_exclusive(ctx, tiletot.p, 4 * ntiles, scan_ws.p, (uint64_t*)nullptr);
    std::swap(B, Bn);
    std::swap(ID, ID2);
    std::swap(NP, NP2);
    std::swap(SYM, SYM2);
    std::swap(WIN, WIN2);
    std::swap(RL, RL2);
    n_cur = n_new;
    nt = ntiles;
    nact_prev = nact;
  }
  SGX_REQUIRE(n_cur == S, SGX_ECUDA, "index: column passes did not produce n_suffixes rows");
  *final_B = B;
}

// Content sort of n read ids (stable).  The reads are strings of 32-base (64-bit) key chunks; bases past the end of a read
// count as 'A' in the key, and the read length is the least significant key, so that a read that is a proper prefix of
// another sorts first ('$' < 'A').  A plain LSD radix sort walks all chunks of all reads (38 eight-bit passes at 150
// bases).  Here ALL reads are sorted by their first chunk only; what that leaves tied -- reads that share their first 32
// bases, a quarter of a 100x read set -- is pulled out, sorted LSD by (length,) the remaining chunks and finally by tie
// group, and put back into the places the tied reads held (the groups are contiguous and in order, so the sorted tied
// list maps one to one onto the tied places).  One gather per read and chunk builds the keys; the 8-bit passes inside a
// chunk stream (key, id) pairs.  The result is in `ids`; `keys` holds nothing useful afterwards.
struct TiedDigit {
  const uint64_t* keys;
  uint64_t n;
  __device__ __forceinline__ int operator()(uint64_t k) const {
    const uint64_t v = keys[k];
    return ((k > 0 && keys[k - 1] == v) || (k + 1 < n && keys[k + 1] == v)) ? 0 : 255;
  }
};
struct PlaceMove {
  uint32_t* P;
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t dest) const { P[dest] = (uint32_t)k; }
};
static __global__ void tied_take_kernel(const uint32_t* __restrict__ ids, const uint64_t* __restrict__ keys,
                                        const uint32_t* __restrict__ P, uint64_t nt, uint32_t* __restrict__ T,
                                        uint64_t* __restrict__ heads) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nt) return;
  T[k] = ids[P[k]];
  heads[k] = (k == 0 || keys[P[k]] != keys[P[k - 1]]) ? 1 : 0;  // neighbours in the tied list with one key = one group
}
static __global__ void tied_group_kernel(const uint64_t* __restrict__ excl, const uint64_t* __restrict__ total,
                                         const uint32_t* __restrict__ T, uint64_t nt, uint32_t* __restrict__ group_of) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nt) return;
  const uint64_t incl = k + 1 < nt ? excl[k + 1] : *total;
  group_of[T[k]] = (uint32_t)(incl - 1);
}
static __global__ void group_keys_kernel(const uint32_t* __restrict__ group_of, const uint32_t* __restrict__ T, uint64_t nt,
                                         uint64_t* __restrict__ keys) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nt) keys[k] = group_of[T[k]];
}
static __global__ void tied_put_kernel(const uint32_t* __restrict__ P, const uint32_t* __restrict__ T, uint64_t nt,
                                       uint32_t* __restrict__ ids) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nt) ids[P[k]] = T[k];
}

void index_content_sort(sgx_ctx* ctx, const sgx_reads* reads, const ReadAcc& acc, uint32_t* ids, uint32_t* ids_alt,
                        uint64_t* keys, uint64_t* keys_alt, uint64_t n, uint64_t* part_ws) {
  if (n == 0) return;
  const uint32_t lmax = reads->max_len;
  const bool variable = reads->min_len != lmax;
  const uint32_t nchunks = (lmax + 31) / 32;
  // ---- everything by the first chunk
  const uint32_t nb0 = std::min<uint32_t>(32, lmax);
  SGX_LAUNCH(ctx, chunk_keys_kernel, grid_for(n, 256), 256, 0, acc, (const uint32_t*)ids, n, 0u, nb0, keys);
  radix_sort_pairs(ctx, keys, ids, keys_alt, ids_alt, n, 2 * (int)nb0, part_ws);
  if (nchunks == 1 && !variable) return;
  // ---- the places of the reads that are still tied
  DevBuf<uint32_t> P(n), T2(n), group_of(reads->n);
  DevBuf<uint64_t> heads(n + 1), scan_ws(scan_ws_elems(n + 1)), d_cnt(1);
  TiedDigit td{keys, n};
  PlaceMove pm{P.p};
  partition_pass<1>(ctx, td, pm, n, part_ws, d_cnt.p);
  uint64_t nt = 0;
  SGX_CUDA(cudaMemcpyAsync(&nt, d_cnt.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (nt == 0) return;
  uint32_t* T = ids_alt;
  SGX_LAUNCH(ctx, tied_take_kernel, grid_for(nt, 256), 256, 0, (const uint32_t*)ids, (const uint64_t*)keys, (const uint32_t*)P.p, nt, T,
             heads.p);
  scan_exclusive(ctx, heads.p, nt, scan_ws.p, d_cnt.p);
  SGX_LAUNCH(ctx, tied_group_kernel, grid_for(nt, 256), 256, 0, (const uint64_t*)heads.p, (const uint64_t*)d_cnt.p, (const uint32_t*)T,
             nt, group_of.p);
  // ---- the tied reads: LSD by length, chunks last to second, tie group
  if (variable) {
    int lbits = 1;
    while ((1u << lbits) <= lmax) ++lbits;
    SGX_LAUNCH(ctx, len_keys_kernel, grid_for(nt, 256), 256, 0, acc.rv, (const uint32_t*)T, nt, keys);
    radix_sort_pairs(ctx, keys, T, keys_alt, T2.p, nt, lbits, part_ws);
  }
  for (uint32_t ch = nchunks; ch-- > 1;) {
    const uint32_t nbases = std::min<uint32_t>(32, lmax - 32 * ch);
    SGX_LAUNCH(ctx, chunk_keys_kernel, grid_for(nt, 256), 256, 0, acc, (const uint32_t*)T, nt, 32 * ch, nbases, keys);
    radix_sort_pairs(ctx, keys, T, keys_alt, T2.p, nt, 2 * (int)nbases, part_ws);
  }
  int gbits = 1;
  while (((uint64_t)1 << gbits) < nt) ++gbits;  // group indices are below nt
  SGX_LAUNCH(ctx, group_keys_kernel, grid_for(nt, 256), 256, 0, (const uint32_t*)group_of.p, (const uint32_t*)T, nt, keys);
  radix_sort_pairs(ctx, keys, T, keys_alt, T2.p, nt, gbits, part_ws);
  SGX_LAUNCH(ctx, tied_put_kernel, grid_for(nt, 256), 256, 0, (const uint32_t*)P.p, (const uint32_t*)T, nt, ids);
}

void index_read_order(sgx_ctx* ctx, const sgx_reads* reads, const ReadAcc& acc, int order, uint32_t* d_sai,
                      uint32_t* d_sord) {
  const uint64_t N = reads->n;
  // ---- 1. read ordering
  DevBuf<uint32_t> idsA(N), idsB(N);
  DevBuf<uint64_t> part_ws(part_ws_elems(N, 256)), heads(N + 1), scan_ws(scan_ws_elems(N + 1)), d_total(1);
  DevBuf<uint64_t> keys(N), keys2(N);
  SGX_LAUNCH(ctx, iota_kernel, grid_for(N, 256), 256, 0, idsA.p, N);
  uint32_t* ids = idsA.p;
  index_content_sort(ctx, reads, acc, ids, idsB.p, keys.p, keys2.p, N, part_ws.p);
  if (order == SGX_ORDER_SAIS) {
    // model A: equal suffixes tie by read index; the content sort is stable, so `ids` is the read-start order and
    // the sentinel of read i is simply the i-th smallest.
    SGX_CUDA(cudaMemcpyAsync(d_sai, ids, N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    SGX_LAUNCH(ctx, iota_kernel, grid_for(N, 256), 256, 0, d_sord, N);
  } else {
    // model B: rank reads by (content, rank of the following read, ...): prefix doubling over the read sequence.
    DevBuf<uint32_t> rank(N), vals(N), vals2(N);
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
    SGX_CUDA(cudaMemcpyAsync(d_sai, order_ids, N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    SGX_LAUNCH(ctx, find_zero_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)d_sai, N, d_total.p);
    SGX_LAUNCH(ctx, sentinel_order_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)d_sai, N,
               (const uint64_t*)d_total.p, d_sord);
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));  // keys/vals go out of scope
  }
}

std::vector<uint64_t> index_active_counts(sgx_ctx* ctx, const sgx_reads* reads) {
  const uint64_t N = reads->n;
  const uint32_t lmax = reads->max_len;
  std::vector<uint64_t> nact(lmax + 2, 0);
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
  return nact;
}

void rle_heads(sgx_ctx* ctx, const uint8_t* B, uint64_t n, RleSegment* seg) {
  seg->n_rows = n;
  seg->n_runs = 0;
  if (n == 0) return;
  // run heads: rows k with B[k] != B[k-1], found 32 rows per thread (count per 8192-row tile first, so that the heads
  // array is sized exactly; then the same masks again, written at the scanned offsets)
  const uint64_t ntiles = (n + kMergeTile - 1) / kMergeTile;
  DevBuf<uint64_t> tilecnt(ntiles), ws_s(scan_ws_elems(ntiles)), d_total(1);
  SGX_LAUNCH(ctx, head_count_kernel, (unsigned)ntiles, 256, 0, B, n, tilecnt.p);
  scan_exclusive(ctx, tilecnt.p, ntiles, ws_s.p, d_total.p);
  uint64_t n_runs = 0;
  SGX_CUDA(cudaMemcpyAsync(&n_runs, d_total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  seg->n_runs = n_runs;
  seg->heads.alloc(n_runs + 1);
  SGX_LAUNCH(ctx, head_write_kernel, (unsigned)ntiles, 256, 0, B, n, (const uint64_t*)tilecnt.p, seg->heads.p);
  // ends of the segment: symbols of the first / last row, start of the second / last run
  uint64_t second = n, last_head = 0;
  uint8_t b0 = 0, b1 = 0;
  if (n_runs > 1) {
    SGX_CUDA(cudaMemcpyAsync(&second, seg->heads.p + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaMemcpyAsync(&last_head, seg->heads.p + n_runs - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  SGX_CUDA(cudaMemcpyAsync(&b0, B, 1, cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaMemcpyAsync(&b1, B + (n - 1) / 2, 1, cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  seg->first_sym = b0 & 15u;
  seg->last_sym = (b1 >> (4 * ((n - 1) & 1))) & 15u;
  seg->first_len = second;
  seg->last_len = n - last_head;
}

void rle_emit(sgx_ctx* ctx, const uint8_t* B, const RleSegment& seg, int skip_first, uint64_t extend_last,
              DevBuf<uint8_t>* d_rle, uint64_t* n_bytes) {
  *n_bytes = 0;
  const uint64_t n_runs = seg.n_runs, n = seg.n_rows;
  if (n_runs == 0) {
    d_rle->alloc(1);
    return;
  }
  DevBuf<uint64_t> d_boff(n_runs + 1), ws_r(scan_ws_elems(n_runs + 1));
  SGX_CUDA(cudaMemsetAsync(d_boff.p, 0, (n_runs + 1) * sizeof(uint64_t), ctx->stream));
  SGX_LAUNCH(ctx, run_bytes_kernel, grid_for(n_runs, 256), 256, 0, (const uint64_t*)seg.heads.p, n_runs, n, skip_first,
             extend_last, d_boff.p);
  scan_exclusive(ctx, d_boff.p, n_runs + 1, ws_r.p, (uint64_t*)nullptr);
  SGX_CUDA(cudaMemcpyAsync(n_bytes, d_boff.p + n_runs, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  d_rle->alloc(*n_bytes ? *n_bytes : 1);
  SGX_LAUNCH(ctx, run_emit_kernel, grid_for(n_runs, 256), 256, 0, B, (const uint64_t*)seg.heads.p,
             (const uint64_t*)d_boff.p, n_runs, n, skip_first, extend_last, d_rle->p);
  // d_boff / heads are released stream-ordered after the emit kernel
}

void index_build(sgx_ctx* ctx, const sgx_reads* reads, int order, bool reversed, sgx_index_out* out) {
  memset(out, 0, sizeof(*out));
  const uint64_t N = reads->n;
  const uint64_t S = reads->bases + N;
  const uint64_t launches0 = ctx->launches;
  EventTimer t_all(ctx->stream), t_sort(ctx->stream), t_bcr(ctx->stream), t_emit(ctx->stream);
  KernelClock k_merge(ctx->stream);
  ReadAcc acc;
  acc.rv = reads->view();
  acc.reversed = reversed ? 1 : 0;

  t_all.start();
  t_sort.start();
  DevBuf<uint32_t> d_sai(N), d_sord(N);
  index_read_order(ctx, reads, acc, order, d_sai.p, d_sord.p);
  t_sort.stop();

  // ---- 2. column passes
  t_bcr.start();
  // reads still active at pass j = reads with length >= j
  std::vector<uint64_t> nact = index_active_counts(ctx, reads);
  // bytes of a 4-bit array of S rows, padded by a tile and a vector so that whole-vector accesses stay in bounds
  const uint64_t padded = bwt_array_bytes(S);
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
  uint64_t n_bytes = 0;
  DevBuf<uint8_t> d_rle;
  {
    RleSegment seg;
    rle_heads(ctx, Bfinal, S, &seg);
    rle_emit(ctx, Bfinal, seg, 0, 0, &d_rle, &n_bytes);
  }
  out->n_strings = N;
  out->n_suffixes = S;
  out->n_runs = n_bytes;
  out->seg_bytes = n_bytes;
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
