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

// Test hooks (tests/test_gpu_parity.py): the named environment variable overrides a sizing decision so that the
// multi-batch loops, the regrow-and-redo branches and the edge-buffer overflow round run on small inputs.
static inline uint64_t env_u64(const char* name, uint64_t dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  const unsigned long long x = strtoull(v, nullptr, 10);
  return x ? (uint64_t)x : dflt;
}

// Candidate block reduced to what the edge pipeline needs: capped[0] (rows of the '$' block, always < 2^32 because
// there are < 2^32 reads), overlap length (low 16 bits) and AlignFlags (bits 16..18).
struct LeanCand {
  uint32_t lo, hi, len_af;
};

struct OvBatch {
  LeanCand* lean;          // [count][4][find_cap]   (lean mode instead of cand)
  LeanCand* lean_contain;  // [count][4]             (lean mode instead of contain)
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
               OvArenaCfg ac, const uint32_t* __restrict__ list, int by_item, uint32_t* __restrict__ status,
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
    const uint64_t rid = list ? (uint64_t)list[wi] : wi;  // which read (relative to `first`)
    const uint64_t r = by_item ? wi : rid;                // where its candidates / outputs are stored
    const uint32_t l = reads.len[first + rid];
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

// ---- shared driver state ------------------------------------------------------------------------------------------

struct OvPlan {
  sgx_ctx* ctx;
  const sgx_fm *fwd, *rev;
  const sgx_reads* reads;
  uint64_t first, count;
  uint32_t min_ov;
  int irreducible, rc;
  FmView vf, vr;
  FmPair fmp;
  ReadsView rv, treads;
  SaView sf, sr;
  bool narrow, fast;
  size_t persist_bytes, persist_window;  // SGX_L2_PERSIST experiment: persisting-L2 carve-out, max window
  uint32_t find_cap, out_cap, rows_cap, extw;
  unsigned grid_search, grid_post, grid_reduce;
  // general-path work area (allocated on first use, regrown on overflow)
  OvArenaCfg ac;
  DevBuf<sgx_block> d_arena;
  DevBuf<uint32_t> d_gd;
  bool arena_ready;
  DevBuf<uint32_t> d_status;
  DevBuf<unsigned long long> d_look;  // [0] search lookups, [1] general-path lookups, [2] rank-block lines the search fetched

  void init(sgx_ctx* c, const sgx_fm* f, const sgx_fm* r, const sgx_reads* rd, uint64_t fst, uint64_t cnt, uint32_t m,
            int irr, int rcomp) {
    static_assert(sizeof(sgx_block) == 72, "sgx_block layout");
    SGX_REQUIRE(f && r && rd, SGX_EINVAL, "sgx_overlap: null argument");
    SGX_REQUIRE(fst <= rd->n && cnt <= rd->n - fst, SGX_EINVAL, "sgx_overlap: read range out of bounds");
    SGX_REQUIRE(m >= 1, SGX_EINVAL, "sgx_overlap: min_overlap must be >= 1");
    ctx = c;
    fwd = f;
    rev = r;
    reads = rd;
    first = fst;
    count = cnt;
    min_ov = m;
    irreducible = irr;
    rc = rcomp;
    vf = f->view();
    vr = r->view();
    rv = rd->view();
    fmp.blocks[0] = vf.blocks;
    fmp.blocks[1] = vr.blocks;
    fmp.inv[0] = f->inv;
    fmp.inv[1] = r->inv;
    fmp.ktab[0] = f->ktab;
    fmp.ktab[1] = r->ktab;
    fmp.ktab_k = (f->ktab && r->ktab && f->ktab_k == r->ktab_k) ? f->ktab_k : 0;
    for (int i = 0; i < 6; ++i) {
      fmp.pred[0][i] = vf.pred[i];
      fmp.pred[1][i] = vr.pred[i];
    }
    const char* fw = getenv("SGX_FORCE_WIDE");  // tests: exercise the 64-bit position path on small inputs
    narrow = f->n < 0xffffffffull && r->n < 0xffffffffull && !(fw && fw[0] == '1');  // 32-bit positions suffice
    const uint32_t lmax = rd->max_len;
    find_cap = lmax > m ? lmax - m : 1;
    out_cap = (uint32_t)env_u64("SGX_OUT_CAP", 4 * find_cap + 4);  // hook: small regions force the redo branch
    grid_search = (unsigned)c->sm_count * 8;
    grid_post = (unsigned)c->sm_count * 4;
    ac.list_cap = find_cap + 4 + 28;
    ac.r_cap = ac.list_cap;
    ac.st_cap = 4 * ac.list_cap;
    ac.g_cap = 256;
    arena_ready = false;
    persist_bytes = persist_window = 0;
    if (env_u64("SGX_L2_PERSIST", 0)) {
      cudaDeviceProp pr;
      SGX_CUDA(cudaGetDeviceProperties(&pr, c->device));
      persist_bytes = (size_t)pr.persistingL2CacheMaxSize;
      persist_window = (size_t)pr.accessPolicyMaxWindowSize;
      SGX_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist_bytes));
    }
    d_status.alloc(1);
    d_look.alloc(3);
    SGX_CUDA(cudaMemsetAsync(d_look.p, 0, 3 * sizeof(unsigned long long), c->stream));
    // fast post path: needs the read-start orders and the indexed read set on the device (sgx_fm_attach_reads)
    fast = f->sai && r->sai && f->indexed && f->indexed == r->indexed && f->indexed->max_len < 65536 &&
           f->indexed->ctx == c && lmax < 65536;
    treads = rv;
    sf = SaView{nullptr, nullptr};
    sr = SaView{nullptr, nullptr};
    rows_cap = 128;
    extw = 1;
    grid_reduce = 1;
    if (fast) {
      treads = f->indexed->view();
      sf = SaView{f->sai, f->inv};
      sr = SaView{r->sai, r->inv};
      const uint32_t tmax = f->indexed->max_len;
      extw = tmax > m ? (tmax - m + 31) / 32 : 1;
      grid_reduce = (unsigned)c->sm_count * 16;
    }
  }

  size_t reduce_smem(bool edges) {
    while (rows_cap > 32 && reduce_smem_per_warp(rows_cap, extw, edges) * kRdWarps > (size_t)200 * 1024) rows_cap /= 2;
    return reduce_smem_per_warp(rows_cap, extw, edges) * kRdWarps;
  }

  void alloc_arena() {
    const uint64_t post_groups = (uint64_t)grid_post * kOvThreads / kGroup;
    d_arena.alloc(post_groups * ((size_t)4 * ac.list_cap + ac.r_cap + ac.st_cap));
    d_gd.alloc(post_groups * (size_t)4 * ac.g_cap);
    ac.blocks = d_arena.p;
    ac.gd = d_gd.p;
    arena_ready = true;
  }

  // lean search: resident CTAs per SM selectable for experiments (SGX_LEAN_MINB = 4 / 5 / 6)
  template <class IdxT>
  void launch_lean(uint64_t rfirst, uint64_t n, const OvBatch& bt, const uint32_t* list) {
    const int minb = (int)env_u64("SGX_LEAN_MINB", 5);
    const unsigned grid = (unsigned)ctx->sm_count * (unsigned)env_u64("SGX_LEAN_WAVES", 1) *
                          (unsigned)env_u64("SGX_LEAN_CTAS", minb == 9 ? 4 : minb);  // CTAs per SM; fewer than resident leaves room for the reduce kernel
    // uniform: the queries are the indexed reads and all of them have one length (see the kernel's header)
    const bool uni = fast && reads == fwd->indexed && reads->min_len == reads->max_len;
    if (minb == 9) {  // experiments: the round-1 kernel (full rank vectors, other-side interval carried along)
      SGX_LAUNCH(ctx, (ov_search_quad_kernel<IdxT, true>), grid_search, kOvThreads, 0, fmp, rv, rfirst, n, min_ov, rc, bt, list, d_look.p);
      return;
    }
    // SGX_L2_PERSIST=1 (experiment): one launch per index, its rank blocks under a persisting-L2 access window
    const bool persist = env_u64("SGX_L2_PERSIST", 0) != 0;
    for (int pass = 0; pass < (persist ? 2 : 1); ++pass) {
      const int only = persist ? pass : -1;
      if (persist) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        const sgx_fm* f = pass == 0 ? fwd : rev;
        av.accessPolicyWindow.base_ptr = (void*)f->blocks;
        av.accessPolicyWindow.num_bytes = std::min<size_t>(fm_num_blocks(f->n) * 128, persist_window);
        av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist_bytes / (double)av.accessPolicyWindow.num_bytes);
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        SGX_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av));
      }
      if (uni && !env_u64("SGX_LEAN_LOCKSTEP", 0)) {  // reads of one length: persistent quads
        const uint32_t ul = reads->max_len, uw = (reads->max_len + 15) / 16;
        // start table: usable when no '$' probe falls inside the skipped steps and a step is left after them
        const int ktk = (fmp.ktab_k >= 2 && min_ov >= (uint32_t)fmp.ktab_k && ul > (uint32_t)fmp.ktab_k && sizeof(IdxT) == 4) ? fmp.ktab_k : 0;
        if (env_u64("SGX_LEAN_PQ2", 1)) {  // cold state in shared memory (default; 0 = the register-heavy variant)
          if (minb == 6) SGX_LAUNCH(ctx, (ov_search_pq2_kernel<IdxT, 6>), grid, kOvThreads, 0, fmp, rv, ul, uw, rfirst, n, min_ov, rc, bt, list, d_look.p, only, ktk);
          else if (minb == 5) SGX_LAUNCH(ctx, (ov_search_pq2_kernel<IdxT, 5>), grid, kOvThreads, 0, fmp, rv, ul, uw, rfirst, n, min_ov, rc, bt, list, d_look.p, only, ktk);
          else SGX_LAUNCH(ctx, (ov_search_pq2_kernel<IdxT, 4>), grid, kOvThreads, 0, fmp, rv, ul, uw, rfirst, n, min_ov, rc, bt, list, d_look.p, only, ktk);
        } else if (minb == 6) SGX_LAUNCH(ctx, (ov_search_pq_kernel<IdxT, 6>), grid, kOvThreads, 0, fmp, rv, ul, uw, rfirst, n, min_ov, rc, bt, list, d_look.p, only);
        else if (minb == 5) SGX_LAUNCH(ctx, (ov_search_pq_kernel<IdxT, 5>), grid, kOvThreads, 0, fmp, rv, ul, uw, rfirst, n, min_ov, rc, bt, list, d_look.p, only);
        else SGX_LAUNCH(ctx, (ov_search_pq_kernel<IdxT, 4>), grid, kOvThreads, 0, fmp, rv, ul, uw, rfirst, n, min_ov, rc, bt, list, d_look.p, only);
      } else if (minb == 6) {
        if (uni) SGX_LAUNCH(ctx, (ov_search_lean_kernel<IdxT, 6, true>), grid, kOvThreads, 0, fmp, rv, rfirst, n, min_ov, rc, bt, list, d_look.p, only);
        else SGX_LAUNCH(ctx, (ov_search_lean_kernel<IdxT, 6, false>), grid, kOvThreads, 0, fmp, rv, rfirst, n, min_ov, rc, bt, list, d_look.p, only);
      } else if (minb == 5) {
        if (uni) SGX_LAUNCH(ctx, (ov_search_lean_kernel<IdxT, 5, true>), grid, kOvThreads, 0, fmp, rv, rfirst, n, min_ov, rc, bt, list, d_look.p, only);
        else SGX_LAUNCH(ctx, (ov_search_lean_kernel<IdxT, 5, false>), grid, kOvThreads, 0, fmp, rv, rfirst, n, min_ov, rc, bt, list, d_look.p, only);
      } else {
        if (uni) SGX_LAUNCH(ctx, (ov_search_lean_kernel<IdxT, 4, true>), grid, kOvThreads, 0, fmp, rv, rfirst, n, min_ov, rc, bt, list, d_look.p, only);
        else SGX_LAUNCH(ctx, (ov_search_lean_kernel<IdxT, 4, false>), grid, kOvThreads, 0, fmp, rv, rfirst, n, min_ov, rc, bt, list, d_look.p, only);
      }
    }
    if (persist) {
      cudaStreamAttrValue av;
      memset(&av, 0, sizeof(av));
      av.accessPolicyWindow.num_bytes = 0;
      SGX_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av));
    }
  }

  // search of `n` work items starting at read `rfirst` (or the reads of `list`, relative to rfirst)
  void search(bool lean, uint64_t rfirst, uint64_t n, const OvBatch& bt, const uint32_t* list) {
    if (n == 0) return;
    if (narrow) {
      if (lean) {
        launch_lean<uint32_t>(rfirst, n, bt, list);
      } else {
        SGX_LAUNCH(ctx, (ov_search_quad_kernel<uint32_t, false>), grid_search, kOvThreads, 0, fmp, rv, rfirst, n, min_ov, rc, bt, list, d_look.p);
      }
    } else {
      if (lean) {
        launch_lean<uint64_t>(rfirst, n, bt, list);
      } else {
        SGX_LAUNCH(ctx, (ov_search_quad_kernel<uint64_t, false>), grid_search, kOvThreads, 0, fmp, rv, rfirst, n, min_ov, rc, bt, list, d_look.p);
      }
    }
  }

  // General (BWT-walking) post stage for `n` work items; returns false when the per-read output regions are too
  // small (the caller enlarges them and calls again); work-area overflows are handled here.
  bool post_general(uint64_t rfirst, uint64_t n, const OvBatch& bt, const uint32_t* list, int by_item, KernelClock& clk) {
    if (n == 0) return true;
    if (!arena_ready) alloc_arena();
    for (int attempt = 0;; ++attempt) {
      SGX_CUDA(cudaMemsetAsync(d_status.p, 0, sizeof(uint32_t), ctx->stream));
      clk.tick();
      SGX_LAUNCH(ctx, ov_post_kernel, grid_post, kOvThreads, 0, vf, vr, rv, rfirst, n, irreducible, bt, ac, list, by_item,
                 d_status.p, d_look.p + 1);
      clk.tick();
      uint32_t st = 0;
      SGX_CUDA(cudaMemcpyAsync(&st, d_status.p, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
      SGX_CUDA(cudaStreamSynchronize(ctx->stream));
      SGX_REQUIRE(!(st & kOvAssert), SGX_EINVAL,
                  "overlap: blocks of equal length with different coordinates (the reference asserts here)");
      if (!(st & kOvOverflow)) return true;
      SGX_REQUIRE(attempt < 5, SGX_EOVERFLOW, "overlap: work area overflow persists after regrowing");
      if (attempt >= 1) return false;  // probably the output regions
      ac.list_cap *= 2;
      ac.r_cap *= 4;
      ac.st_cap *= 4;
      ac.g_cap *= 4;
      alloc_arena();
    }
  }
};

// Device-resident result of the search + post stages (blocks of all reads, compacted).
struct OvCore {
  DevBuf<sgx_block> compact;
  uint64_t n_blocks = 0;
  DevBuf<uint8_t> sub;   // [count] substring flags
  DevBuf<uint64_t> off;  // [count + 1] block offsets
  uint64_t slow_reads = 0, search_launches = 0;
};

// ---- blocks API (sgx_overlap_batch): every block exactly as the reference serialises it ---------------------------

static void overlap_core(OvPlan& pl, OvCore& core, KernelClock& k_search, KernelClock& k_post) {
  sgx_ctx* ctx = pl.ctx;
  const uint64_t count = pl.count;
  // batch size: keep candidate + output regions within ~6 GB
  const size_t per_read = ((size_t)4 * pl.find_cap + 4 + pl.out_cap) * sizeof(sgx_block);
  uint64_t batch = std::max<uint64_t>(1024, std::min<uint64_t>((6ull << 30) / per_read, 1u << 20));
  batch = env_u64("SGX_MAX_BATCH", batch);  // hook: many small batches
  batch = std::min<uint64_t>(batch, std::max<uint64_t>(count, 1));

  DevBuf<sgx_block> d_cand(batch * 4 * pl.find_cap), d_contain(batch * 4), d_out(batch * pl.out_cap);
  DevBuf<uint32_t> d_ncand(batch * 4), d_ncontain(batch * 4), d_nout(batch), d_slow, d_nslow(1);
  core.sub.alloc(std::max<uint64_t>(count, 1));
  core.off.alloc(count + 1);
  DevBuf<uint64_t> d_off(batch + 1), d_ws(scan_ws_elems(batch + 1));
  SGX_CUDA(cudaMemsetAsync(core.sub.p, 0, std::max<uint64_t>(count, 1), ctx->stream));
  if (count == 0) SGX_CUDA(cudaMemsetAsync(core.off.p, 0, sizeof(uint64_t), ctx->stream));
  size_t smem = 0;
  if (pl.fast) {
    smem = pl.reduce_smem(false);
    SGX_CUDA(cudaFuncSetAttribute(ov_reduce_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    d_slow.alloc(batch);
  }
  uint64_t compact_cap = 0, compact_n = 0;
  for (uint64_t b0 = 0; b0 < count; b0 += batch) {
    const uint64_t bn = std::min<uint64_t>(batch, count - b0);
    OvBatch bt;
    memset(&bt, 0, sizeof(bt));
    bt.cand = d_cand.p;
    bt.ncand = d_ncand.p;
    bt.contain = d_contain.p;
    bt.ncontain = d_ncontain.p;
    bt.substring = core.sub.p + b0;
    bt.out = d_out.p;
    bt.nout = d_nout.p;
    bt.find_cap = pl.find_cap;
    bt.out_cap = pl.out_cap;
    k_search.tick();
    pl.search(false, pl.first + b0, bn, bt, nullptr);
    k_search.tick();
    core.search_launches++;
    for (;;) {
      const uint32_t* work_list = nullptr;
      uint64_t work_n = bn;
      if (pl.fast) {
        SGX_CUDA(cudaMemsetAsync(d_nslow.p, 0, sizeof(uint32_t), ctx->stream));
        EdgeSink none;
        memset(&none, 0, sizeof(none));
        k_post.tick();
        SGX_LAUNCH(ctx, (ov_reduce_kernel<false, false>), pl.grid_reduce, kRdWarps * 32, smem, pl.rv, pl.treads, pl.sf, pl.sr,
                   bn, pl.irreducible, bt, pl.rows_cap, pl.extw, (uint64_t)0, d_slow.p, d_nslow.p, none);
        k_post.tick();
        uint32_t ns = 0;
        SGX_CUDA(cudaMemcpyAsync(&ns, d_nslow.p, sizeof(ns), cudaMemcpyDeviceToHost, ctx->stream));
        SGX_CUDA(cudaStreamSynchronize(ctx->stream));
        work_list = d_slow.p;
        work_n = ns;
      }
      core.slow_reads += work_n;
      if (pl.post_general(pl.first + b0, work_n, bt, work_list, 0, k_post)) break;
      // the per-read output regions are too small: enlarge and redo the post stage of this batch
      core.slow_reads -= work_n;
      pl.out_cap *= 2;
      d_out.alloc(batch * (size_t)pl.out_cap);
      bt.out = d_out.p;
      bt.out_cap = pl.out_cap;
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
        SGX_CUDA(cudaMemcpyAsync(nbuf.p, core.compact.p, compact_n * sizeof(sgx_block), cudaMemcpyDeviceToDevice, ctx->stream));
      core.compact.swap(nbuf);  // the old area is released when nbuf goes out of scope (stream ordered)
      compact_cap = ncap;
    }
    if (btotal)
      SGX_LAUNCH(ctx, ov_gather_kernel, grid_for(bn * 32, 256), 256, 0, (const sgx_block*)bt.out, (const uint32_t*)d_nout.p,
                 bt.out_cap, (const uint64_t*)d_off.p, bn, core.compact.p + compact_n);
    SGX_LAUNCH(ctx, ov_offsets_kernel, grid_for(bn + 1, 256), 256, 0, (const uint64_t*)d_off.p, bn, compact_n,
               core.off.p + b0);
    compact_n += btotal;
  }
  core.n_blocks = compact_n;
}

void overlap_run(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                 uint64_t count, uint32_t min_ov, int irreducible, int rc, sgx_hits* hits) {
  SGX_REQUIRE(hits, SGX_EINVAL, "sgx_overlap_batch: null output");
  memset(hits, 0, sizeof(*hits));
  const uint64_t launches0 = ctx->launches;
  EventTimer t_all(ctx->stream);
  KernelClock k_search(ctx->stream), k_post(ctx->stream);
  t_all.start();
  OvPlan pl;
  pl.init(ctx, fwd, rev, reads, first, count, min_ov, irreducible, rc);
  OvCore core;
  overlap_core(pl, core, k_search, k_post);
  hits->n_reads = count;
  hits->n_blocks = core.n_blocks;
  hits->slow_reads = core.slow_reads;
  hits->search_launches = core.search_launches;
  hits->substring = (uint8_t*)pinned_get(ctx, std::max<uint64_t>(count, 1));
  hits->block_off = (uint64_t*)pinned_get(ctx, (count + 1) * sizeof(uint64_t));
  hits->blocks = (sgx_block*)pinned_get(ctx, std::max<uint64_t>(core.n_blocks, 1) * sizeof(sgx_block));
  SGX_CUDA(cudaMemcpyAsync(hits->block_off, core.off.p, (count + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (count) SGX_CUDA(cudaMemcpyAsync(hits->substring, core.sub.p, count, cudaMemcpyDeviceToHost, ctx->stream));
  if (core.n_blocks)
    SGX_CUDA(cudaMemcpyAsync(hits->blocks, core.compact.p, core.n_blocks * sizeof(sgx_block), cudaMemcpyDeviceToHost, ctx->stream));
  unsigned long long look[3] = {0, 0, 0};
  SGX_CUDA(cudaMemcpyAsync(look, pl.d_look.p, sizeof(look), cudaMemcpyDeviceToHost, ctx->stream));
  t_all.stop();
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  hits->lines_search = look[2];
  hits->ms_total = t_all.ms();
  hits->ms_search = k_search.total_ms();
  hits->ms_extract = k_post.total_ms();
  hits->lookups_search = look[0];
  hits->lookups_extract = look[1];
  hits->launches = ctx->launches - launches0;
}

// ---- edges API (sgx_overlap_edges) -------------------------------------------------------------------------------
// Lean pipeline: the search writes 12-byte candidates, the reduce kernel applies the filter of
// Hit2OverlapConverter::convert (src/overlap_builder.cpp:345-375) to the surviving rows and appends sgx_edge records
// straight to the output; no blocks, no compaction, no host round trip per batch.  Reads outside the fast path's
// preconditions are collected and redone at the end through the full-block kernels + ov_edges_items_kernel.

// Deterministic edge order: reads reserve their slice of the edge array with an atomic, so the order of the slices
// depends on warp scheduling; a stable sort by query read (LSD radix over the query index, the edges of one read keep
// the order their warp wrote them in) makes the array -- and the ED section the CLI writes from it -- identical from
// run to run.
__global__ void ov_edge_keys_kernel(const sgx_edge* __restrict__ edges, uint64_t n, uint64_t first, uint64_t* __restrict__ keys,
                                    uint32_t* __restrict__ vals) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (uint64_t)edges[i].query - first;
  vals[i] = (uint32_t)i;
}
__global__ void ov_edge_permute_kernel(const sgx_edge* __restrict__ edges, const uint32_t* __restrict__ order, uint64_t n,
                                       sgx_edge* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 e = *reinterpret_cast<const uint4*>(edges + order[i]);
  *reinterpret_cast<uint4*>(out + i) = e;
}

// One thread per general-path work item: expands the blocks it produced into edges.
__global__ void ov_edges_items_kernel(const sgx_block* __restrict__ out, const uint32_t* __restrict__ nout, uint32_t out_cap,
                                      const uint32_t* __restrict__ list, uint64_t n_items, uint64_t first,
                                      const uint32_t* __restrict__ qlen, SaView sf, SaView sr,
                                      const uint32_t* __restrict__ tlen, EdgeSink sink) {
  const uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= n_items) return;
  const uint64_t q = first + (list ? (uint64_t)list[wi] : wi);
  const uint32_t lq = qlen[q], rq = sink.name_rank[q];
  const sgx_block* blocks = out + wi * (size_t)out_cap;
  const uint32_t nb = nout[wi];
  // count, reserve, write
  unsigned long long n = 0;
  for (uint32_t b = 0; b < nb; ++b) {
    const sgx_block& blk = blocks[b];
    const uint32_t* sa = (blk.af & SGX_AF_TARGETREV) ? sr.sa : sf.sa;
    for (uint64_t j = blk.capped[0]; j <= blk.capped[1]; ++j)
      if (rd_edge_ok(rq, lq, sa[j], blk.length, blk.af, sink.name_rank, tlen)) ++n;
  }
  if (n == 0) return;
  const unsigned long long at = atomicAdd(sink.counter, n);
  if (at + n > sink.cap) {
    atomicOr(sink.overflow, 1u);
    return;
  }
  unsigned long long w = 0;
  for (uint32_t b = 0; b < nb; ++b) {
    const sgx_block& blk = blocks[b];
    const uint32_t* sa = (blk.af & SGX_AF_TARGETREV) ? sr.sa : sf.sa;
    for (uint64_t j = blk.capped[0]; j <= blk.capped[1]; ++j) {
      const uint32_t t = sa[j];
      if (!rd_edge_ok(rq, lq, t, blk.length, blk.af, sink.name_rank, tlen)) continue;
      sgx_edge e;
      e.query = (uint32_t)q;
      e.target = t;
      e.length = blk.length;
      e.af = blk.af;
      sink.edges[at + w] = e;
      ++w;
    }
  }
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
  OvPlan pl;
  pl.init(ctx, fwd, rev, reads, first, count, min_ov, irreducible, rc);
  if (!pl.fast) reads_settle(reads);
  const uint64_t n_all = fwd->indexed->n;
  DevBuf<uint32_t> d_rank(n_all);
  SGX_CUDA(cudaMemcpyAsync(d_rank.p, name_rank, n_all * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  DevBuf<uint8_t> d_sub(std::max<uint64_t>(count, 1));
  DevBuf<uint32_t> d_slow(std::max<uint64_t>(count, 1)), d_flags(2);  // [0] n_slow, [1] overflow
  DevBuf<unsigned long long> d_counter(2);
  DevBuf<sgx_edge> d_edges;
  // edges kept per read: ~3-5 on average at 100x coverage, but the id0 >= id1 rule keeps up to twice the average for a
  // range of reads whose names sort high (a rank of a multi-GPU run); an overflow costs a whole second round
  uint64_t edge_cap = env_u64("SGX_EDGE_CAP", std::max<uint64_t>(count * 12, 1u << 16));  // hook: overflow round
  uint64_t n_edges = 0, n_slow = 0, search_launches = 0;
  double post2_ms = 0;  // reduce kernels that ran on the second stream

  for (int round = 0;; ++round) {
    SGX_REQUIRE(round < 8, SGX_EOVERFLOW, "sgx_overlap_edges: edge buffer overflow persists");
    d_edges.alloc(edge_cap);
    SGX_CUDA(cudaMemsetAsync(d_sub.p, 0, std::max<uint64_t>(count, 1), ctx->stream));
    SGX_CUDA(cudaMemsetAsync(d_flags.p, 0, 2 * sizeof(uint32_t), ctx->stream));
    SGX_CUDA(cudaMemsetAsync(d_counter.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
    SGX_CUDA(cudaMemsetAsync(pl.d_look.p, 0, 3 * sizeof(unsigned long long), ctx->stream));
    EdgeSink sink;
    sink.edges = d_edges.p;
    sink.counter = d_counter.p;
    sink.cap = edge_cap;
    sink.overflow = d_flags.p + 1;
    sink.name_rank = d_rank.p;
    sink.qbase = 0;
    search_launches = 0;
    if (pl.fast) {
      // ---- lean pass over all reads, batch after batch, nothing comes back to the host in between
      const size_t per_read = ((size_t)4 * pl.find_cap + 4) * sizeof(LeanCand) + 32;
      uint64_t batch = std::max<uint64_t>(1024, std::min<uint64_t>((4ull << 30) / per_read, 1u << 21));
      batch = env_u64("SGX_MAX_BATCH", batch);
      batch = std::min<uint64_t>(batch, std::max<uint64_t>(count, 1));
      DevBuf<LeanCand> d_lean(batch * 4 * pl.find_cap), d_lcontain(batch * 4);
      DevBuf<uint32_t> d_ncand(batch * 4), d_ncontain(batch * 4);
      const sgx_reads* idx = fwd->indexed;
      const bool uniform = idx->min_len == idx->max_len;
      uint32_t rows = (uint32_t)env_u64("SGX_REDUCE_ROWS", 128);  // rows of a read held in shared memory at a time
      while (rows > 32 && reduce_lean_smem_per_warp(rows, pl.extw) * kRlWarps > (size_t)200 * 1024) rows /= 2;
      const size_t smem = reduce_lean_smem_per_warp(rows, pl.extw) * kRlWarps;
      SGX_CUDA(cudaFuncSetAttribute(ov_reduce_lean_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SGX_CUDA(cudaFuncSetAttribute(ov_reduce_lean_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const unsigned grid_rl = (unsigned)ctx->sm_count * (unsigned)env_u64("SGX_REDUCE_CTAS", 16);
      const bool old_reduce = env_u64("SGX_REDUCE_OLD", 0) != 0;
      const size_t smem_old = pl.reduce_smem(true);
      SGX_CUDA(cudaFuncSetAttribute(ov_reduce_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_old));
      // SGX_OVERLAP_STREAMS=2 (experiment, off by default): the search of batch b + 1 runs while batch b is reduced, on
      // two streams with two sets of candidate buffers (events order search(b) -> reduce(b) -> search(b + 2)).  Measured
      // at cfg2: the kernels do overlap, each stretches, the pass takes the same 63.5 ms -- both live off the same
      // random-access rate of the memory system (profiles/r2_summary.md).
      const bool two = env_u64("SGX_OVERLAP_STREAMS", 1) >= 2 && count > batch && !old_reduce;
      sgx_ctx* aux = two ? ctx_aux(ctx) : nullptr;
      DevBuf<LeanCand> d_lean2, d_lcontain2;
      DevBuf<uint32_t> d_ncand2, d_ncontain2;
      cudaEvent_t ev_s[2] = {nullptr, nullptr}, ev_r[2] = {nullptr, nullptr};
      if (two) {
        d_lean2.alloc(batch * 4 * pl.find_cap);
        d_lcontain2.alloc(batch * 4);
        d_ncand2.alloc(batch * 4);
        d_ncontain2.alloc(batch * 4);
        for (int k = 0; k < 2; ++k) {
          SGX_CUDA(cudaEventCreateWithFlags(&ev_s[k], cudaEventDisableTiming));
          SGX_CUDA(cudaEventCreateWithFlags(&ev_r[k], cudaEventDisableTiming));
        }
      }
      struct EvGuard {
        cudaEvent_t* a;
        cudaEvent_t* b;
        ~EvGuard() {
          for (int k = 0; k < 2; ++k) {
            if (a[k]) cudaEventDestroy(a[k]);
            if (b[k]) cudaEventDestroy(b[k]);
          }
        }
      } ev_guard{ev_s, ev_r};
      KernelClock k_post2(two ? aux->stream : ctx->stream);
      sgx_ctx* rctx = two ? aux : ctx;  // where the reduce kernels are launched
      const uint64_t aux_launches0 = two ? aux->launches : 0;
      uint64_t bi = 0;
      for (uint64_t b0 = 0; b0 < count; b0 += batch, ++bi) {
        const uint64_t bn = std::min<uint64_t>(batch, count - b0);
        const int set = two ? (int)(bi & 1) : 0;
        OvBatch bt;
        memset(&bt, 0, sizeof(bt));
        bt.lean = set ? d_lean2.p : d_lean.p;
        bt.lean_contain = set ? d_lcontain2.p : d_lcontain.p;
        bt.ncand = set ? d_ncand2.p : d_ncand.p;
        bt.ncontain = set ? d_ncontain2.p : d_ncontain.p;
        bt.substring = d_sub.p + b0;
        bt.find_cap = pl.find_cap;
        bt.out_cap = pl.out_cap;
        if (two && bi >= 2) SGX_CUDA(cudaStreamWaitEvent(ctx->stream, ev_r[set], 0));  // the reduce of batch b - 2 is done with this set
        // a read store that is still arriving (sgx_reads_pack_async): the search needs the query reads of its batch only
        reads_wait_range(ctx->stream, reads, first + b0, first + b0 + bn);
        k_search.tick();
        pl.search(true, first + b0, bn, bt, nullptr);
        k_search.tick();
        ++search_launches;
        sink.qbase = first + b0;
        if (two) {
          SGX_CUDA(cudaEventRecord(ev_s[set], ctx->stream));
          SGX_CUDA(cudaStreamWaitEvent(aux->stream, ev_s[set], 0));
        }
        if (bi == 0) reads_wait_range(ctx->stream, reads, 0, reads->n);  // the reduce reads targets anywhere in the store
        KernelClock& kp = two ? k_post2 : k_post;
        kp.tick();
        if (old_reduce) {  // experiments: the round-1 kernel (one side after the other, staged edges)
          SGX_LAUNCH(ctx, (ov_reduce_kernel<true, true>), pl.grid_reduce, kRdWarps * 32, smem_old, pl.rv, pl.treads, pl.sf, pl.sr, bn,
                     irreducible, bt, pl.rows_cap, pl.extw, b0, d_slow.p, d_flags.p, sink);
        } else if (uniform) {
          SGX_LAUNCH(rctx, (ov_reduce_lean_kernel<true>), grid_rl, kRlWarps * 32, smem, pl.rv, pl.treads, idx->max_len,
                     (idx->max_len + 15) / 16, pl.sf, pl.sr, bn, irreducible, bt, rows, pl.extw, b0, d_slow.p, d_flags.p, sink);
        } else {
          SGX_LAUNCH(rctx, (ov_reduce_lean_kernel<false>), grid_rl, kRlWarps * 32, smem, pl.rv, pl.treads, 0u, 0u, pl.sf,
                     pl.sr, bn, irreducible, bt, rows, pl.extw, b0, d_slow.p, d_flags.p, sink);
        }
        kp.tick();
        if (two) SGX_CUDA(cudaEventRecord(ev_r[set], aux->stream));
      }
      if (two) {
        for (int k = 0; k < 2 && (uint64_t)k < bi; ++k) SGX_CUDA(cudaStreamWaitEvent(ctx->stream, ev_r[k], 0));
        SGX_CUDA(cudaStreamSynchronize(ctx->stream));  // k_post2's events are read below; the buffers go back to the cache
        post2_ms += k_post2.total_ms();
        ctx->launches += aux->launches - aux_launches0;
      }
      uint32_t fl[2] = {0, 0};
      SGX_CUDA(cudaMemcpyAsync(fl, d_flags.p, sizeof(fl), cudaMemcpyDeviceToHost, ctx->stream));
      SGX_CUDA(cudaStreamSynchronize(ctx->stream));
      reads_settle(reads);  // all chunks have arrived by now; a store with a bad base is refused here
      n_slow = fl[0];
      if (fl[1]) {
        edge_cap *= 4;
        continue;
      }
    } else {
      n_slow = count;  // no read store attached: everything takes the general path
    }
    // ---- general path for the reads the lean pass handed back (or all of them): full blocks -> edges
    bool again = false;
    if (n_slow) {
      const size_t per_item = ((size_t)4 * pl.find_cap + 4 + pl.out_cap) * sizeof(sgx_block);
      uint64_t chunk = std::max<uint64_t>(1024, std::min<uint64_t>((4ull << 30) / per_item, 1u << 20));
      chunk = env_u64("SGX_MAX_BATCH", chunk);
      chunk = std::min<uint64_t>(chunk, n_slow);
      DevBuf<sgx_block> d_cand(chunk * 4 * pl.find_cap), d_contain(chunk * 4), d_out(chunk * (size_t)pl.out_cap);
      DevBuf<uint32_t> d_ncand(chunk * 4), d_ncontain(chunk * 4), d_nout(chunk);
      sink.qbase = first;
      for (uint64_t c0 = 0; c0 < n_slow && !again; c0 += chunk) {
        const uint64_t cn = std::min<uint64_t>(chunk, n_slow - c0);
        const uint32_t* list = pl.fast ? d_slow.p + c0 : nullptr;
        const uint64_t rfirst = pl.fast ? first : first + c0;
        OvBatch bt;
        memset(&bt, 0, sizeof(bt));
        bt.cand = d_cand.p;
        bt.contain = d_contain.p;
        bt.ncand = d_ncand.p;
        bt.ncontain = d_ncontain.p;
        bt.substring = pl.fast ? d_sub.p : d_sub.p + c0;
        bt.out = d_out.p;
        bt.nout = d_nout.p;
        bt.find_cap = pl.find_cap;
        bt.out_cap = pl.out_cap;
        k_search.tick();
        pl.search(false, rfirst, cn, bt, list);
        k_search.tick();
        ++search_launches;
        while (!pl.post_general(rfirst, cn, bt, list, 1, k_post)) {
          pl.out_cap *= 2;
          d_out.alloc(chunk * (size_t)pl.out_cap);
          bt.out = d_out.p;
          bt.out_cap = pl.out_cap;
        }
        SGX_LAUNCH(ctx, ov_edges_items_kernel, grid_for(cn, 128), 128, 0, (const sgx_block*)bt.out, (const uint32_t*)d_nout.p,
                   bt.out_cap, list, cn, rfirst, (const uint32_t*)reads->len, pl.sf, pl.sr,
                   (const uint32_t*)fwd->indexed->len, sink);
      }
      uint32_t ov = 0;
      SGX_CUDA(cudaMemcpyAsync(&ov, d_flags.p + 1, sizeof(ov), cudaMemcpyDeviceToHost, ctx->stream));
      SGX_CUDA(cudaStreamSynchronize(ctx->stream));
      if (ov) again = true;
    }
    if (again) {
      edge_cap *= 4;
      continue;
    }
    unsigned long long ne = 0;
    SGX_CUDA(cudaMemcpyAsync(&ne, d_counter.p, sizeof(ne), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    n_edges = ne;
    break;
  }
  if (n_edges > 1 && n_edges < 0xffffffffull && !env_u64("SGX_EDGES_UNSORTED", 0)) {
    static_assert(sizeof(sgx_edge) == 16, "sgx_edge is moved as one 16-byte vector");
    DevBuf<uint64_t> keys(n_edges), keys2(n_edges), ws(part_ws_elems(n_edges, 256));
    DevBuf<uint32_t> vals(n_edges), vals2(n_edges);
    DevBuf<sgx_edge> sorted(n_edges);
    SGX_LAUNCH(ctx, ov_edge_keys_kernel, grid_for(n_edges, 256), 256, 0, (const sgx_edge*)d_edges.p, n_edges, first, keys.p, vals.p);
    int bits = 1;
    while (bits < 32 && ((uint64_t)1 << bits) < std::max<uint64_t>(count, 2)) ++bits;
    radix_sort_pairs(ctx, keys.p, vals.p, keys2.p, vals2.p, n_edges, bits, ws.p);
    SGX_LAUNCH(ctx, ov_edge_permute_kernel, grid_for(n_edges, 256), 256, 0, (const sgx_edge*)d_edges.p, (const uint32_t*)vals.p,
               n_edges, sorted.p);
    d_edges.swap(sorted);
  }
  out->n_reads = count;
  out->n_edges = n_edges;
  out->n_blocks = 0;
  out->slow_reads = n_slow;
  out->search_launches = search_launches;
  out->substring = (uint8_t*)pinned_get(ctx, std::max<uint64_t>(count, 1));
  out->edges = (sgx_edge*)pinned_get(ctx, std::max<uint64_t>(n_edges, 1) * sizeof(sgx_edge));
  if (count) SGX_CUDA(cudaMemcpyAsync(out->substring, d_sub.p, count, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_edges)
    SGX_CUDA(cudaMemcpyAsync(out->edges, d_edges.p, n_edges * sizeof(sgx_edge), cudaMemcpyDeviceToHost, ctx->stream));
  unsigned long long look[3] = {0, 0, 0};
  SGX_CUDA(cudaMemcpyAsync(look, pl.d_look.p, sizeof(look), cudaMemcpyDeviceToHost, ctx->stream));
  t_all.stop();
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  out->lines_search = look[2];
  out->ms_total = t_all.ms();
  out->ms_search = k_search.total_ms();
  out->ms_extract = k_post.total_ms() + post2_ms;
  out->lookups_search = look[0];
  out->lookups_extract = look[1];
  out->launches = ctx->launches - launches0;
}

}  // namespace sgx
