// FM-index construction on the device: run bytes -> symbols -> 128-byte rank blocks (fm_layout.h).
// Replaces FMIndex::initialize / MarkerFill (src/fmindex.cpp:17-161 of the reference).
#include "fm.cuh"

#include <algorithm>

#include "reads.cuh"

namespace sgx {

__global__ void rle_count_kernel(const uint8_t* __restrict__ rle, uint64_t n_runs, uint64_t* __restrict__ cnt) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_runs) cnt[i] = rle[i] & 31u;
}

// One thread per run: writes its (<= 31) symbols.  Neighbouring threads write neighbouring byte ranges.
__global__ void rle_expand_kernel(const uint8_t* __restrict__ rle, const uint64_t* __restrict__ start, uint64_t n_runs,
                                  uint64_t n, uint8_t* __restrict__ sym) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_runs) return;
  const uint8_t r = rle[i];
  const uint8_t s = r >> 5;
  const uint32_t c = r & 31u;
  const uint64_t at = start[i];
  for (uint32_t k = 0; k < c; ++k)
    if (at + k < n) sym[at + k] = s;
}

void rle_expand(sgx_ctx* ctx, const uint8_t* rle_dev, uint64_t n_runs, uint64_t n_suffixes, uint8_t* sym_dev) {
  if (n_runs == 0) {
    SGX_REQUIRE(n_suffixes == 0, SGX_EINVAL, "rle_expand: no runs but n_suffixes != 0");
    return;
  }
  DevBuf<uint64_t> start(n_runs), ws(scan_ws_elems(n_runs)), total(1);
  SGX_LAUNCH(ctx, rle_count_kernel, grid_for(n_runs, 256), 256, 0, rle_dev, n_runs, start.p);
  scan_exclusive(ctx, start.p, n_runs, ws.p, total.p);
  uint64_t tot = 0;
  SGX_CUDA(cudaMemcpyAsync(&tot, total.p, sizeof(tot), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  SGX_REQUIRE(tot == n_suffixes, SGX_EINVAL, "rle_expand: run lengths do not add up to n_suffixes");
  SGX_LAUNCH(ctx, rle_expand_kernel, grid_for(n_runs, 256), 256, 0, rle_dev, (const uint64_t*)start.p, n_runs,
             n_suffixes, sym_dev);
}

// Thread t builds lane t (32 symbols) of the block array and the groups of 8 threads reduce the block's A/C/G/T
// counts into cnt[c*nb + b].
__global__ void __launch_bounds__(256) fm_pack_kernel(const uint8_t* __restrict__ sym, uint64_t n, uint64_t nb,
                                                      FmBlockLane* __restrict__ blocks, uint64_t* __restrict__ cnt) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nlanes = nb * kGroup;
  uint32_t p0 = 0, p1 = 0, p2 = 0;
  uint32_t ca = 0, cc = 0, cg = 0, ct = 0;
  if (t < nlanes) {
    const uint64_t base = t * 32;
    if (base + 32 <= n) {
      const uint4 a = *reinterpret_cast<const uint4*>(sym + base);
      const uint4 b = *reinterpret_cast<const uint4*>(sym + base + 16);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t s = (w[q] >> (8 * j)) & 0xffu;
          const uint32_t bit = 1u << (4 * q + j);
          if (s == 0) {
            p2 |= bit;
          } else {
            const uint32_t code = s - 1;
            if (code & 1) p0 |= bit;
            if (code & 2) p1 |= bit;
          }
        }
      }
    } else {
      for (int j = 0; j < 32; ++j) {
        if (base + j >= n) break;
        const uint32_t s = sym[base + j];
        const uint32_t bit = 1u << j;
        if (s == 0) {
          p2 |= bit;
        } else {
          const uint32_t code = s - 1;
          if (code & 1) p0 |= bit;
          if (code & 2) p1 |= bit;
        }
      }
    }
    uint32_t valid = 0xffffffffu;
    if (base + 32 > n) valid = base >= n ? 0u : ((1u << (uint32_t)(n - base)) - 1u);
    const uint32_t nd = ~p2 & valid;
    ca = __popc(nd & ~p1 & ~p0);
    cc = __popc(nd & ~p1 & p0);
    cg = __popc(nd & p1 & ~p0);
    ct = __popc(nd & p1 & p0);
    FmBlockLane l;
    l.p0 = p0;
    l.p1 = p1;
    l.p2 = p2;
    l.h = 0;
    blocks[t] = l;
  }
  uint32_t w1 = ca | (cc << 16), w2 = cg | (ct << 16);
#pragma unroll
  for (int m = 1; m < kGroup; m <<= 1) {
    w1 += __shfl_xor_sync(0xffffffffu, w1, m);
    w2 += __shfl_xor_sync(0xffffffffu, w2, m);
  }
  if (t < nlanes && (t & (kGroup - 1)) == 0) {
    const uint64_t b = t / kGroup;
    cnt[0 * nb + b] = w1 & 0xffffu;
    cnt[1 * nb + b] = w1 >> 16;
    cnt[2 * nb + b] = w2 & 0xffffu;
    cnt[3 * nb + b] = w2 >> 16;
  }
}

// After the in-place scan of cnt (flattened [4][nb]), header word k of block b = low (k < 4) or high (k >= 4) half of
// the count of symbol k & 3 before block b.
__global__ void fm_header_kernel(FmBlockLane* __restrict__ blocks, uint64_t nb, const uint64_t* __restrict__ scanned) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb * kGroup) return;
  const uint64_t b = t / kGroup;
  const int k = (int)(t & (kGroup - 1));
  const int c = k & 3;
  const uint64_t v = scanned[(uint64_t)c * nb + b] - scanned[(uint64_t)c * nb];
  blocks[t].h = (k >= 4) ? (uint32_t)(v >> 32) : (uint32_t)(v & 0xffffffffu);
}

sgx_fm* fm_build_from_symbols(sgx_ctx* ctx, const uint8_t* sym_dev, uint64_t n) {
  const uint64_t nb = fm_num_blocks(n);
  DevBuf<FmBlockLane> blocks;  // becomes the fm object's storage
  blocks.alloc_owned(nb * kGroup);
  DevBuf<uint64_t> cnt(4 * nb), ws(scan_ws_elems(4 * nb)), total(1);
  SGX_LAUNCH(ctx, fm_pack_kernel, grid_for(nb * kGroup, 256), 256, 0, sym_dev, n, nb, blocks.p, cnt.p);
  scan_exclusive(ctx, cnt.p, 4 * nb, ws.p, total.p);
  SGX_LAUNCH(ctx, fm_header_kernel, grid_for(nb * kGroup, 256), 256, 0, blocks.p, nb, (const uint64_t*)cnt.p);
  uint64_t b4[5];
  for (int c = 0; c < 4; ++c)
    SGX_CUDA(cudaMemcpyAsync(&b4[c], cnt.p + (uint64_t)c * nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaMemcpyAsync(&b4[4], total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  sgx_fm* fm = new sgx_fm();
  fm->ctx = ctx;
  fm->n = n;
  fm->sai = nullptr;
  fm->inv = nullptr;
  fm->n_strings = 0;
  fm->indexed = nullptr;
  uint64_t acgt = 0;
  uint64_t cntc[5];
  for (int c = 0; c < 4; ++c) {
    cntc[c + 1] = b4[c + 1] - b4[c];
    acgt += cntc[c + 1];
  }
  cntc[0] = n - acgt;
  fm->pred[0] = 0;
  for (int c = 1; c <= 5; ++c) fm->pred[c] = fm->pred[c - 1] + cntc[c - 1];
  fm->blocks = blocks.take();
  fm->ktab = nullptr;
  fm->ktab_k = 0;
  // k-mer start table -- an EXPERIMENT, off unless SGX_KTAB=k (2..13) asks for it: measured at config 2 with k = 12 it
  // removes 22 % of the fetched rank-block lines and 1 % of the search time (45.5 -> 45.0 ms): the first steps of a chain
  // walk the top of the index, which stays in L2, and cost next to nothing (profiles/r2_summary.md).
  if (n < 0xffffff00ull) {
    int k = 0;
    if (const char* e = getenv("SGX_KTAB")) k = std::max(0, std::min(13, atoi(e)));
    while (k > 1 && ((uint64_t)1 << (2 * k)) > 4 * n + 64) --k;  // never much larger than the index itself
    if (k >= 2) {
      try {
        fm_build_ktab(ctx, fm, k);
      } catch (...) {
        cudaFreeAsync(fm->blocks, ctx->stream);
        delete fm;
        throw;
      }
    }
  }
  return fm;
}

// Level t of the k-mer table from level t - 1: one 8-lane group per parent string, two rank queries (all four symbols
// at once), four children.  child[4 p + c] extends parent p by base c on the left.
__global__ void __launch_bounds__(256) ktab_level_kernel(FmView fm, const uint2* __restrict__ parent, uint64_t nparent,
                                                         uint2* __restrict__ child) {
  const DevGroup g;
  const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
  if (gid >= nparent) return;
  const uint2 pe = parent[gid];
  uint64_t L[5] = {0, 0, 0, 0, 0}, U[5] = {0, 0, 0, 0, 0};
  if (pe.y) {  // group-uniform
    fm_occ(fm, (pos_t)pe.x - 1, L, g);
    fm_occ(fm, (pos_t)pe.x + pe.y - 1, U, g);
  }
  if (g.lane < 4) {
    const int c = g.lane + 1;
    uint2 ce = make_uint2(1u, 0u);  // empty
    if (pe.y && U[c] > L[c]) ce = make_uint2((uint32_t)(fm.pred[c] + L[c]), (uint32_t)(U[c] - L[c]));
    child[4 * gid + g.lane] = ce;
  }
}

void fm_build_ktab(sgx_ctx* ctx, sgx_fm* fm, int k) {
  SGX_REQUIRE(k >= 1 && k <= 13 && fm->n < 0xffffff00ull, SGX_EINVAL, "fm_build_ktab: k out of range or index too large");
  const uint64_t n_final = (uint64_t)1 << (2 * k);
  DevBuf<uint2> tab, tmp;
  tab.alloc_owned(n_final);
  tmp.alloc(std::max<uint64_t>(n_final / 4, 4));
  // level t lands in `tab` when k - t is even, so that level k does
  uint2 l1[4];
  for (int c = 1; c <= 4; ++c) {
    const uint64_t rows = fm->pred[c + 1] - fm->pred[c];
    l1[c - 1] = rows ? make_uint2((uint32_t)fm->pred[c], (uint32_t)rows) : make_uint2(1u, 0u);
  }
  uint2* cur = ((k - 1) % 2 == 0) ? tab.p : tmp.p;
  SGX_CUDA(cudaMemcpyAsync(cur, l1, sizeof(l1), cudaMemcpyHostToDevice, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));  // l1 is a stack array
  const FmView v = fm->view();
  for (int t = 2; t <= k; ++t) {
    uint2* nxt = ((k - t) % 2 == 0) ? tab.p : tmp.p;
    const uint64_t nparent = (uint64_t)1 << (2 * (t - 1));
    SGX_LAUNCH(ctx, ktab_level_kernel, grid_for(nparent * kGroup, 256), 256, 0, v, (const uint2*)cur, nparent, nxt);
    cur = nxt;
  }
  if (fm->ktab) cudaFreeAsync(fm->ktab, ctx->stream);
  fm->ktab = tab.take();
  fm->ktab_k = k;
}


__global__ void sai_inverse_kernel(const uint32_t* __restrict__ sai, uint64_t n, uint32_t* __restrict__ inv,
                                   uint32_t* __restrict__ bad) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t id = sai[k];
  if (id >= n) {
    atomicOr(bad, 1u);
    return;
  }
  inv[id] = (uint32_t)k;
}

void fm_attach(sgx_ctx* ctx, sgx_fm* fm, const uint32_t* sai, uint64_t n, const sgx_reads* indexed) {
  SGX_REQUIRE(sai && indexed && n == indexed->n, SGX_EINVAL, "sgx_fm_attach_reads: sai / read set size mismatch");
  SGX_REQUIRE(n + indexed->bases == fm->n, SGX_EINVAL, "sgx_fm_attach_reads: read set does not match the BWT length");
  DevBuf<uint32_t> d_sai, d_inv, d_bad(1);  // sai / inv become the fm object's storage
  d_sai.alloc_owned(n);
  d_inv.alloc_owned(n);
  SGX_CUDA(cudaMemcpyAsync(d_sai.p, sai, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  SGX_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(uint32_t), ctx->stream));
  SGX_LAUNCH(ctx, sai_inverse_kernel, grid_for(n, 256), 256, 0, (const uint32_t*)d_sai.p, n, d_inv.p, d_bad.p);
  uint32_t bad = 0;
  SGX_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  SGX_REQUIRE(bad == 0, SGX_EINVAL, "sgx_fm_attach_reads: read id out of range in sai");
  if (fm->sai) cudaFreeAsync(fm->sai, ctx->stream);
  if (fm->inv) cudaFreeAsync(fm->inv, ctx->stream);
  fm->sai = d_sai.take();
  fm->inv = d_inv.take();
  fm->n_strings = n;
  fm->indexed = indexed;
}

// One 8-lane group per query position.
__global__ void __launch_bounds__(256) fm_occ_kernel(FmView fm, const uint64_t* __restrict__ pos, uint64_t n,
                                                     uint64_t* __restrict__ counts) {
  const DevGroup g;
  const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
  const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) / kGroup;
  for (uint64_t i = gid; i < n; i += ngroups) {
    uint64_t c[5];
    fm_occ(fm, pos[i], c, g);
    if (g.leader())
      for (int b = 0; b < 5; ++b) counts[5 * i + b] = c[b];
  }
}

void fm_occ_batch(sgx_ctx* ctx, const sgx_fm* fm, const uint64_t* pos, uint64_t n, uint64_t* counts) {
  if (n == 0) return;
  for (uint64_t i = 0; i < n; ++i)
    SGX_REQUIRE(pos[i] == kNoPos || pos[i] < fm->n, SGX_EINVAL, "sgx_fm_occ: position beyond the BWT");
  DevBuf<uint64_t> d_pos(n), d_cnt(5 * n);
  SGX_CUDA(cudaMemcpyAsync(d_pos.p, pos, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  const unsigned grid = (unsigned)std::min<uint64_t>((n * kGroup + 255) / 256, (uint64_t)ctx->sm_count * 8);
  SGX_LAUNCH(ctx, fm_occ_kernel, grid, 256, 0, fm->view(), (const uint64_t*)d_pos.p, n, d_cnt.p);
  SGX_CUDA(cudaMemcpyAsync(counts, d_cnt.p, 5 * n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
}

// FMIndex::Interval::get + occurrences (src/fmindex.h:67-86): plain backward search, one group per query.
__global__ void __launch_bounds__(256) fm_count_kernel(FmView fm, ReadsView q, uint64_t* __restrict__ counts) {
  const DevGroup g;
  const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
  const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) / kGroup;
  for (uint64_t i = gid; i < q.n; i += ngroups) {
    const uint32_t* rw = q.words + q.woff[i];
    uint32_t j = q.len[i];
    int c = read_rank(rw, j - 1);
    pos_t lo = fm.pred[c], hi = fm.pred[c + 1] - 1;
    uint64_t L[5], U[5];
    while (--j > 0 && ival_valid(lo, hi)) {
      c = read_rank(rw, j - 1);
      fm_occ(fm, lo - 1, L, g);
      fm_occ(fm, hi, U, g);
      lo = fm.pred[c] + L[c];
      hi = fm.pred[c] + U[c] - 1;
    }
    if (g.leader()) counts[i] = ival_valid(lo, hi) ? hi - lo + 1 : 0;
  }
}

void fm_count_batch(sgx_ctx* ctx, const sgx_fm* fm, const char* ascii, const uint64_t* off, uint64_t n, uint64_t* counts) {
  if (n == 0) return;
  sgx_reads* q = reads_pack(ctx, ascii, off, n);
  try {
    DevBuf<uint64_t> d_cnt(n);
    const unsigned grid = (unsigned)std::min<uint64_t>((n * kGroup + 255) / 256, (uint64_t)ctx->sm_count * 8);
    SGX_LAUNCH(ctx, fm_count_kernel, grid, 256, 0, fm->view(), q->view(), d_cnt.p);
    SGX_CUDA(cudaMemcpyAsync(counts, d_cnt.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (...) {
    reads_destroy(q);
    throw;
  }
  reads_destroy(q);
}

}  // namespace sgx
