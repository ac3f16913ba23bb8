// Overlap search on the device: the body of parallel::foreach in OverlapBuilder::build
// (src/overlap_builder.cpp:453-457 of the reference), i.e. OverlapBuilder::overlap per read, as two kernels:
//
//   ov_search_quad_kernel  the four OverlapBlockFinder::find passes of every read, one 4-lane quad per (read, mode)
//                     (overlap_search.cuh); every backward step is one pair of 128-byte rank-block fetches.
//   ov_reduce_kernel  post stage, common case: irreducible extraction by comparing target-read extensions
//                     (overlap_reduce.cuh), one warp per read.
//   ov_post_kernel    post stage, general case: the literal SubMaximalBlockFilter + ContainmentBlockRemover +
//                     IrreducibleBlockListExtractor on a per-group work area (BWT walks; data dependent).
//   ov_gather_kernel  compacts the per-read outputs into one array for the copy out.
//
// The per-read control logic lives in overlap_logic.h and is the same text the host stand-in of the tests runs.
#include "overlap.cuh"

#include <algorithm>

#include "overlap_logic.h"

namespace sgx {

static const int kOvThreads = 256;

struct OvBatch {
  sgx_block* cand;     // [count][4][find_cap]
  uint32_t* ncand;     // [count][4]
  sgx_block* contain;  // [count][4]
  uint32_t* ncontain;  // [count][4]
  uint8_t* substring;  // [count]
  sgx_block* out;      // [count][out_cap]
  uint32_t* nout;      // [count]
  uint32_t find_cap;
  uint32_t out_cap;
};

}  // namespace sgx

#include "overlap_reduce.cuh"
#include "overlap_search.cuh"

namespace sgx {

struct OvArenaCfg {
  sgx_block* blocks;  // per group: 4*list_cap + r_cap + st_cap blocks
  uint32_t* gd;       // per group: 4*g_cap words
  uint32_t list_cap, r_cap, st_cap, g_cap;
};

__global__ void __launch_bounds__(kOvThreads)
ov_post_kernel(FmView fwd, FmView rev, ReadsView reads, uint64_t first, uint64_t count, int irreducible, OvBatch bt,
               OvArenaCfg ac, const uint32_t* __restrict__ list, uint32_t* __restrict__ status,
               unsigned long long* __restrict__ lookups) {
  const DevGroup g;
  const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
  const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) / kGroup;
  const size_t per_group = (size_t)4 * ac.list_cap + ac.r_cap + ac.st_cap;
  OvArena ar;
  ar.lists = ac.blocks + gid * per_group;
  ar.list_cap = ac.list_cap;
  ar.R = ar.lists + (size_t)4 * ac.list_cap;
  ar.r_cap = ac.r_cap;
  ar.st = ar.R + ac.r_cap;
  ar.st_cap = ac.st_cap;
  ar.gd = ac.gd + gid * (size_t)4 * ac.g_cap;
  ar.g_cap = ac.g_cap;
  uint64_t nlook = 0;
  uint32_t st = 0;
  // `count` work items: all reads of the batch, or (list != nullptr) the reads the fast path handed back
  for (uint64_t wi = gid; wi < count; wi += ngroups) {
    const uint64_t r = list ? (uint64_t)list[wi] : wi;
    const uint32_t l = reads.len[first + r];
    const sgx_block* cp[4];
    uint32_t nc[4], nk[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      cp[f] = bt.cand + ((size_t)r * 4 + f) * bt.find_cap;
      nc[f] = bt.ncand[r * 4 + f];
      nk[f] = bt.ncontain[r * 4 + f];
    }
    uint32_t nout = 0;
    ov_post(fwd, rev, l, irreducible, cp, nc, bt.contain + (size_t)r * 4, nk, ar, bt.out + (size_t)r * bt.out_cap, &nout,
            bt.out_cap, &st, &nlook, g);
    g.sync();
    if (g.leader()) bt.nout[r] = nout;
  }
  if (g.leader()) {
    if (nlook) atomicAdd(lookups, (unsigned long long)nlook);
    if (st) atomicOr(status, st);
  }
}

__global__ void ov_widen_kernel(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
  if (i == n) out[i] = 0;
}

// One warp per read: copies its blocks (9 x 8 bytes each) to the compact array.
__global__ void ov_gather_kernel(const sgx_block* __restrict__ out, const uint32_t* __restrict__ nout, uint32_t out_cap,
                                 const uint64_t* __restrict__ off, uint64_t count, sgx_block* __restrict__ dst) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= count) return;
  const uint64_t* src = reinterpret_cast<const uint64_t*>(out + warp * (size_t)out_cap);
  uint64_t* d = reinterpret_cast<uint64_t*>(dst + off[warp]);
  const uint32_t words = nout[warp] * (uint32_t)(sizeof(sgx_block) / 8);
  for (uint32_t w = lane; w < words; w += 32) d[w] = src[w];
}

// all_off[b0 + i] = base + batch_off[i] for i in [0, bn] : per-read block offsets into the compact array
__global__ void ov_offsets_kernel(const uint64_t* __restrict__ batch_off, uint64_t bn, uint64_t base,
                                  uint64_t* __restrict__ all_off) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= bn) all_off[i] = base + batch_off[i];
}

// Device-resident result of the search + post stages (blocks of all reads, compacted).
struct OvCore {
  DevBuf<sgx_block> compact;
  uint64_t n_blocks = 0;
  DevBuf<uint8_t> sub;        // [count] substring flags
  DevBuf<uint64_t> off;       // [count + 1] block offsets
  sgx_hits stats;             // timing / counter fields only
};

static void overlap_core(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                         uint64_t count, uint32_t min_ov, int irreducible, int rc, OvCore& core,
                         KernelClock& k_search, KernelClock& k_post) {
  static_assert(sizeof(sgx_block) == 72, "sgx_block layout");
  SGX_REQUIRE(fwd && rev && reads, SGX_EINVAL, "sgx_overlap_batch: null argument");
  SGX_REQUIRE(first <= reads->n && count <= reads->n - first, SGX_EINVAL, "sgx_overlap_batch: read range out of bounds");
  SGX_REQUIRE(min_ov >= 1, SGX_EINVAL, "sgx_overlap_batch: min_overlap must be >= 1");
  sgx_hits* hits = &core.stats;
  memset(hits, 0, sizeof(*hits));

  const uint32_t lmax = reads->max_len;
  const uint32_t find_cap = lmax > min_ov ? lmax - min_ov : 1;
  uint32_t out_cap = 4 * find_cap + 4;
  // batch size: keep candidate + output regions within ~6 GB
  const size_t per_read = ((size_t)4 * find_cap + 4 + out_cap) * sizeof(sgx_block);
  uint64_t batch = std::max<uint64_t>(1024, std::min<uint64_t>((6ull << 30) / per_read, 1u << 20));
  batch = std::min<uint64_t>(batch, std::max<uint64_t>(count, 1));

  DevBuf<sgx_block> d_cand(batch * 4 * find_cap), d_contain(batch * 4), d_out(batch * out_cap);
  DevBuf<uint32_t> d_ncand(batch * 4), d_ncontain(batch * 4), d_nout(batch), d_status(1);
  DevBuf<uint8_t>& d_sub = core.sub;
  d_sub.alloc(std::max<uint64_t>(count, 1));
  core.off.alloc(count + 1);
  DevBuf<uint64_t> d_off(batch + 1), d_ws(scan_ws_elems(batch + 1)), d_total(1);
  DevBuf<unsigned long long> d_look(2);
  SGX_CUDA(cudaMemsetAsync(d_sub.p, 0, std::max<uint64_t>(count, 1), ctx->stream));
  SGX_CUDA(cudaMemsetAsync(d_look.p, 0, 2 * sizeof(unsigned long long), ctx->stream));

  const FmView vf = fwd->view(), vr = rev->view();
  const ReadsView rv = reads->view();
  const ReadsView rv0 = rv;
  FmPair fmp;
  fmp.blocks[0] = vf.blocks;
  fmp.blocks[1] = vr.blocks;
  for (int i = 0; i < 6; ++i) {
    fmp.pred[0][i] = vf.pred[i];
    fmp.pred[1][i] = vr.pred[i];
  }
  const bool narrow = fwd->n < 0xffffffffull && rev->n < 0xffffffffull;  // 32-bit positions suffice
  // grids: a few resident CTAs per SM, groups stride over the reads of the batch
  const unsigned grid_search = (unsigned)ctx->sm_count * 8;
  const unsigned grid_post = (unsigned)ctx->sm_count * 4;
  const uint64_t post_groups = (uint64_t)grid_post * kOvThreads / kGroup;
  OvArenaCfg ac;
  ac.list_cap = find_cap + 4 + 28;
  ac.r_cap = ac.list_cap;
  ac.st_cap = 4 * ac.list_cap;
  ac.g_cap = 256;
  DevBuf<sgx_block> d_arena;
  DevBuf<uint32_t> d_gd;
  auto alloc_arena = [&]() {
    d_arena.alloc(post_groups * ((size_t)4 * ac.list_cap + ac.r_cap + ac.st_cap));
    d_gd.alloc(post_groups * (size_t)4 * ac.g_cap);
    ac.blocks = d_arena.p;
    ac.gd = d_gd.p;
  };
  alloc_arena();

  // fast post path: needs the read-start orders and the indexed read set on the device (sgx_fm_attach_reads)
  const bool fast = fwd->sai && rev->sai && fwd->indexed && fwd->indexed == rev->indexed &&
                    fwd->indexed->max_len < 65536 && fwd->indexed->ctx == ctx;
  ReadsView treads = rv0;
  SaView sf{nullptr, nullptr}, sr{nullptr, nullptr};
  uint32_t rows_cap = 128, extw = 1;
  size_t reduce_smem = 0;
  unsigned grid_reduce = 1;
  DevBuf<uint32_t> d_slow, d_nslow(1);
  if (fast) {
    treads = fwd->indexed->view();
    sf.sa = fwd->sai;
    sf.inv = fwd->inv;
    sr.sa = rev->sai;
    sr.inv = rev->inv;
    const uint32_t tmax = fwd->indexed->max_len;
    extw = tmax > min_ov ? (tmax - min_ov + 31) / 32 : 1;
    while (rows_cap > 32 && reduce_smem_per_warp(rows_cap, extw) * kRdWarps > (size_t)200 * 1024) rows_cap /= 2;
    reduce_smem = reduce_smem_per_warp(rows_cap, extw) * kRdWarps;
    SGX_CUDA(cudaFuncSetAttribute(ov_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reduce_smem));
    grid_reduce = (unsigned)ctx->sm_count * 16;
    d_slow.alloc(batch);
  }

  DevBuf<sgx_block>& d_compact = core.compact;  // all blocks of all batches
  uint64_t compact_cap = 0, compact_n = 0;
  if (count == 0) SGX_CUDA(cudaMemsetAsync(core.off.p, 0, sizeof(uint64_t), ctx->stream));

  for (uint64_t b0 = 0; b0 < count; b0 += batch) {
    const uint64_t bn = std::min<uint64_t>(batch, count - b0);
    OvBatch bt;
    bt.cand = d_cand.p;
    bt.ncand = d_ncand.p;
    bt.contain = d_contain.p;
    bt.ncontain = d_ncontain.p;
    bt.substring = d_sub.p + b0;
    bt.out = d_out.p;
    bt.nout = d_nout.p;
    bt.find_cap = find_cap;
    bt.out_cap = out_cap;
    SGX_CUDA(cudaMemsetAsync(d_status.p, 0, sizeof(uint32_t), ctx->stream));
    k_search.tick();
    if (narrow) {
      SGX_LAUNCH(ctx, ov_search_quad_kernel<uint32_t>, grid_search, kOvThreads, 0, fmp, rv, first + b0, bn, min_ov, rc,
                 bt, d_look.p);
    } else {
      SGX_LAUNCH(ctx, ov_search_quad_kernel<uint64_t>, grid_search, kOvThreads, 0, fmp, rv, first + b0, bn, min_ov, rc,
                 bt, d_look.p);
    }
    k_search.tick();
    hits->search_launches++;
    // ---- post stage.  Fast path (read store + .sai attached): string-based reduction, the few reads outside its
    //      preconditions are handed back on a list.  Otherwise / for that list: the literal BWT-walking kernel.
    const uint32_t* work_list = nullptr;
    uint64_t work_n = bn;
    if (fast) {
      SGX_CUDA(cudaMemsetAsync(d_nslow.p, 0, sizeof(uint32_t), ctx->stream));
      k_post.tick();
      SGX_LAUNCH(ctx, ov_reduce_kernel, grid_reduce, kRdWarps * 32, reduce_smem, treads, sf, sr, bn, irreducible, bt,
                 rows_cap, extw, d_slow.p, d_nslow.p);
      k_post.tick();
      uint32_t ns = 0;
      SGX_CUDA(cudaMemcpyAsync(&ns, d_nslow.p, sizeof(ns), cudaMemcpyDeviceToHost, ctx->stream));
      SGX_CUDA(cudaStreamSynchronize(ctx->stream));
      work_list = d_slow.p;
      work_n = ns;
      hits->slow_reads += ns;
    } else {
      hits->slow_reads += bn;
    }
    for (int attempt = 0; work_n > 0; ++attempt) {
      if (attempt > 0) SGX_CUDA(cudaMemsetAsync(d_status.p, 0, sizeof(uint32_t), ctx->stream));
      k_post.tick();
      SGX_LAUNCH(ctx, ov_post_kernel, grid_post, kOvThreads, 0, vf, vr, rv, first + b0, work_n, irreducible, bt, ac,
                 work_list, d_status.p, d_look.p + 1);
      k_post.tick();
      uint32_t st = 0;
      SGX_CUDA(cudaMemcpyAsync(&st, d_status.p, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
      SGX_CUDA(cudaStreamSynchronize(ctx->stream));
      SGX_REQUIRE(!(st & kOvAssert), SGX_EINVAL,
                  "overlap: blocks of equal length with different coordinates (the reference asserts here)");
      if (!(st & kOvOverflow)) break;
      // a work area was too small: regrow it and redo the post stage of these reads on the GPU
      SGX_REQUIRE(attempt < 4, SGX_EOVERFLOW, "overlap: work area overflow persists after regrowing");
      ac.list_cap *= 2;
      ac.r_cap *= 4;
      ac.st_cap *= 4;
      ac.g_cap *= 4;
      alloc_arena();
      if (attempt >= 1) {
        // the output regions are too small as well: this also invalidates what the fast path wrote, redo everything
        out_cap *= 2;
        d_out.alloc(batch * (size_t)out_cap);
        bt.out = d_out.p;
        bt.out_cap = out_cap;
        work_list = nullptr;
        work_n = bn;
      }
    }
    // compact
    SGX_LAUNCH(ctx, ov_widen_kernel, grid_for(bn + 1, 256), 256, 0, (const uint32_t*)d_nout.p, bn, d_off.p);
    scan_exclusive(ctx, d_off.p, bn + 1, d_ws.p, (uint64_t*)nullptr);
    uint64_t btotal = 0;
    SGX_CUDA(cudaMemcpyAsync(&btotal, d_off.p + bn, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (compact_n + btotal > compact_cap) {
      uint64_t ncap = std::max<uint64_t>(compact_n + btotal, compact_cap * 2);
      if (b0 == 0 && bn < count) ncap = std::max<uint64_t>(ncap, (uint64_t)((double)btotal / bn * count * 1.1) + 1024);
      DevBuf<sgx_block> nbuf(ncap);
      if (compact_n)
        SGX_CUDA(cudaMemcpyAsync(nbuf.p, d_compact.p, compact_n * sizeof(sgx_block), cudaMemcpyDeviceToDevice, ctx->stream));
      d_compact.swap(nbuf);  // the old area is released when nbuf goes out of scope (stream ordered)
      compact_cap = ncap;
    }
    if (btotal)
      SGX_LAUNCH(ctx, ov_gather_kernel, grid_for(bn * 32, 256), 256, 0, (const sgx_block*)bt.out,
                 (const uint32_t*)d_nout.p, bt.out_cap, (const uint64_t*)d_off.p, bn, d_compact.p + compact_n);
    SGX_LAUNCH(ctx, ov_offsets_kernel, grid_for(bn + 1, 256), 256, 0, (const uint64_t*)d_off.p, bn, compact_n,
               core.off.p + b0);
    compact_n += btotal;
  }
  core.n_blocks = compact_n;
  unsigned long long look[2] = {0, 0};
  SGX_CUDA(cudaMemcpyAsync(look, d_look.p, sizeof(look), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  hits->lookups_search = look[0];
  hits->lookups_extract = look[1];
}

void overlap_run(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                 uint64_t count, uint32_t min_ov, int irreducible, int rc, sgx_hits* hits) {
  SGX_REQUIRE(hits, SGX_EINVAL, "sgx_overlap_batch: null output");
  memset(hits, 0, sizeof(*hits));
  const uint64_t launches0 = ctx->launches;
  EventTimer t_all(ctx->stream);
  KernelClock k_search(ctx->stream), k_post(ctx->stream);
  t_all.start();
  OvCore core;
  overlap_core(ctx, fwd, rev, reads, first, count, min_ov, irreducible, rc, core, k_search, k_post);
  *hits = core.stats;
  const uint64_t compact_n = core.n_blocks;
  DevBuf<sgx_block>& d_compact = core.compact;
  DevBuf<uint8_t>& d_sub = core.sub;

  // copy out
  hits->n_reads = count;
  hits->n_blocks = compact_n;
  hits->substring = (uint8_t*)pinned_get(ctx, std::max<uint64_t>(count, 1));
  hits->block_off = (uint64_t*)pinned_get(ctx, (count + 1) * sizeof(uint64_t));
  hits->blocks = (sgx_block*)pinned_get(ctx, std::max<uint64_t>(compact_n, 1) * sizeof(sgx_block));
  SGX_CUDA(cudaMemcpyAsync(hits->block_off, core.off.p, (count + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (count) SGX_CUDA(cudaMemcpyAsync(hits->substring, d_sub.p, count, cudaMemcpyDeviceToHost, ctx->stream));
  if (compact_n)
    SGX_CUDA(cudaMemcpyAsync(hits->blocks, d_compact.p, compact_n * sizeof(sgx_block), cudaMemcpyDeviceToHost, ctx->stream));
  t_all.stop();
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  hits->ms_total = t_all.ms();
  hits->ms_search = k_search.total_ms();
  hits->ms_extract = k_post.total_ms();
  hits->launches = ctx->launches - launches0;
}

// ---- hits -> edges on the device: Hit2OverlapConverter::convert (src/overlap_builder.cpp:345-375) ------------------
// One thread per query read walks its blocks and the rows of each capped[0] interval.  An edge survives when the
// target is another read (by name), the query name is not smaller than the target name, and it is not a
// containment seen from the reversed query (src/overlap_builder.cpp:358-367).  Names enter as ranks.
template <bool WRITE>
__global__ void ov_edges_kernel(const sgx_block* __restrict__ blocks, const uint64_t* __restrict__ off, uint64_t first,
                                uint64_t count, const uint32_t* __restrict__ qlen, SaView sf, SaView sr,
                                const uint32_t* __restrict__ tlen, const uint32_t* __restrict__ name_rank,
                                uint64_t* __restrict__ cnt, sgx_edge* __restrict__ edges) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= count) return;
  const uint32_t q = (uint32_t)(first + r);
  const uint32_t lq = qlen[q];
  const uint32_t rq = name_rank[q];
  uint64_t n = 0;
  uint64_t at = WRITE ? cnt[r] : 0;
  for (uint64_t b = off[r]; b < off[r + 1]; ++b) {
    const sgx_block blk = blocks[b];
    const uint32_t k = blk.length, af = blk.af;
    const uint32_t* sa = (af & SGX_AF_TARGETREV) ? sr.sa : sf.sa;
    for (uint64_t j = blk.capped[0]; j <= blk.capped[1]; ++j) {
      const uint32_t t = sa[j];
      const uint32_t rt = name_rank[t];
      if (rq == rt) continue;                       // query.name == target.name
      const bool contain = k == lq || k == tlen[t];  // either coordinate spans its whole read
      if (rq < rt || (contain && (af & SGX_AF_QUERYREV))) continue;
      if (WRITE) {
        sgx_edge e;
        e.query = q;
        e.target = t;
        e.length = k;
        e.af = af;
        edges[at + n] = e;
      }
      ++n;
    }
  }
  if (!WRITE) cnt[r] = n;
}

void overlap_edges_run(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                       uint64_t count, uint32_t min_ov, int irreducible, int rc, const uint32_t* name_rank,
                       sgx_edges* out) {
  SGX_REQUIRE(out && name_rank, SGX_EINVAL, "sgx_overlap_edges: null argument");
  SGX_REQUIRE(fwd && rev && fwd->sai && rev->sai && fwd->indexed && fwd->indexed == rev->indexed, SGX_EINVAL,
              "sgx_overlap_edges: both indices need sgx_fm_attach_reads");
  SGX_REQUIRE(reads == fwd->indexed, SGX_EINVAL, "sgx_overlap_edges: the query reads must be the indexed read set");
  memset(out, 0, sizeof(*out));
  const uint64_t launches0 = ctx->launches;
  EventTimer t_all(ctx->stream);
  KernelClock k_search(ctx->stream), k_post(ctx->stream);
  t_all.start();
  OvCore core;
  overlap_core(ctx, fwd, rev, reads, first, count, min_ov, irreducible, rc, core, k_search, k_post);
  const uint64_t n_all = fwd->indexed->n;
  DevBuf<uint32_t> d_rank(n_all);
  SGX_CUDA(cudaMemcpyAsync(d_rank.p, name_rank, n_all * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  DevBuf<uint64_t> d_cnt(count + 1), d_ws(scan_ws_elems(count + 1)), d_total(1);
  SGX_CUDA(cudaMemsetAsync(d_cnt.p, 0, (count + 1) * sizeof(uint64_t), ctx->stream));
  SaView sf{fwd->sai, fwd->inv}, sr{rev->sai, rev->inv};
  SGX_LAUNCH(ctx, ov_edges_kernel<false>, grid_for(count, 256), 256, 0, (const sgx_block*)core.compact.p,
             (const uint64_t*)core.off.p, first, count, (const uint32_t*)reads->len, sf, sr,
             (const uint32_t*)fwd->indexed->len, (const uint32_t*)d_rank.p, d_cnt.p, (sgx_edge*)nullptr);
  scan_exclusive(ctx, d_cnt.p, count + 1, d_ws.p, (uint64_t*)nullptr);
  uint64_t n_edges = 0;
  SGX_CUDA(cudaMemcpyAsync(&n_edges, d_cnt.p + count, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  DevBuf<sgx_edge> d_edges(std::max<uint64_t>(n_edges, 1));
  SGX_LAUNCH(ctx, ov_edges_kernel<true>, grid_for(count, 256), 256, 0, (const sgx_block*)core.compact.p,
             (const uint64_t*)core.off.p, first, count, (const uint32_t*)reads->len, sf, sr,
             (const uint32_t*)fwd->indexed->len, (const uint32_t*)d_rank.p, d_cnt.p, d_edges.p);
  out->n_reads = count;
  out->n_edges = n_edges;
  out->n_blocks = core.n_blocks;
  out->substring = (uint8_t*)pinned_get(ctx, std::max<uint64_t>(count, 1));
  out->edges = (sgx_edge*)pinned_get(ctx, std::max<uint64_t>(n_edges, 1) * sizeof(sgx_edge));
  if (count) SGX_CUDA(cudaMemcpyAsync(out->substring, core.sub.p, count, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_edges)
    SGX_CUDA(cudaMemcpyAsync(out->edges, d_edges.p, n_edges * sizeof(sgx_edge), cudaMemcpyDeviceToHost, ctx->stream));
  t_all.stop();
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  out->ms_total = t_all.ms();
  out->ms_search = k_search.total_ms();
  out->ms_extract = k_post.total_ms();
  out->lookups_search = core.stats.lookups_search;
  out->lookups_extract = core.stats.lookups_extract;
  out->slow_reads = core.stats.slow_reads;
  out->search_launches = core.stats.search_launches;
  out->launches = ctx->launches - launches0;
}

}  // namespace sgx
