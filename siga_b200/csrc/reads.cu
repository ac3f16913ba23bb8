// ASCII reads -> 2-bit device read store.  Input side of both hot paths; stands where the reference holds a
// std::vector<DNASeq> of std::strings (src/kseq.cpp:230-257).
#include "reads.cuh"

#include "prims.cuh"

namespace sgx {

// len[i], words-per-read (as uint64 for the scan), and min/max length.
__global__ void reads_len_kernel(const uint64_t* __restrict__ off, uint64_t n, uint32_t* __restrict__ len,
                                 uint64_t* __restrict__ wc, uint32_t* __restrict__ minmax, uint32_t* __restrict__ bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t l = 0;
  bool ok = false;
  if (i < n) {
    const uint64_t d = off[i + 1] - off[i];
    if (off[i + 1] < off[i] || d == 0 || d > 0x7fffffffull) {
      atomicOr(bad, 1u);
      len[i] = 0;
      wc[i] = 0;
    } else {
      ok = true;
      l = (uint32_t)d;
      len[i] = l;
      wc[i] = (l + 15) / 16;
    }
  }
  uint32_t mn = ok ? l : 0xffffffffu, mx = ok ? l : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&minmax[0], mn);
    atomicMax(&minmax[1], mx);
  }
}

// One warp per read; lane k of each round converts 16 consecutive bases into one 32-bit word.
__global__ void __launch_bounds__(256) reads_pack_kernel(const char* __restrict__ ascii, const uint64_t* __restrict__ off,
                                                         const uint64_t* __restrict__ woff, uint64_t n,
                                                         uint32_t* __restrict__ words, uint32_t* __restrict__ bad) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const uint64_t a = off[warp];
  const uint32_t l = (uint32_t)(off[warp + 1] - a);
  const uint32_t nw = (l + 15) / 16;
  uint32_t* dst = words + woff[warp];
  bool ok = true;
  for (uint32_t w = lane; w < nw; w += 32) {
    uint32_t v = 0;
    const uint32_t jb = w * 16;
    const uint32_t je = min(jb + 16, l);
    for (uint32_t j = jb; j < je; ++j) {
      const char ch = ascii[a + j];
      uint32_t code;
      if (ch == 'A') code = 0;
      else if (ch == 'C') code = 1;
      else if (ch == 'G') code = 2;
      else if (ch == 'T') code = 3;
      else {
        code = 0;
        ok = false;
      }
      v |= code << (2 * (j - jb));
    }
    dst[w] = v;
  }
  if (!ok) {
    atomicOr(bad, 2u);
    atomicMin(bad + 1, (uint32_t)warp);  // first offending read, for the error message
  }
}

sgx_reads* reads_pack(sgx_ctx* ctx, const char* ascii, const uint64_t* off, uint64_t n) {
  SGX_REQUIRE(ascii != nullptr && off != nullptr, SGX_EINVAL, "sgx_reads_pack: null input");
  SGX_REQUIRE(n > 0 && n < 0xffffffffull, SGX_EINVAL, "sgx_reads_pack: read count must be in [1, 2^32-1)");
  const uint64_t total = off[n] - off[0];
  SGX_REQUIRE(off[n] >= off[0] && total > 0, SGX_EINVAL, "sgx_reads_pack: bad offsets");
  DevBuf<char> d_ascii(total);
  DevBuf<uint64_t> d_off(n + 1), d_woff, ws(scan_ws_elems(n + 1));
  DevBuf<uint32_t> d_len, d_flags(4);
  d_woff.alloc_owned(n + 1);  // woff / len / words become the read store
  d_len.alloc_owned(n);
  SGX_CUDA(cudaMemcpyAsync(d_ascii.p, ascii + off[0], total, cudaMemcpyHostToDevice, ctx->stream));
  SGX_CUDA(cudaMemcpyAsync(d_off.p, off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  const uint32_t init[4] = {0xffffffffu, 0u, 0u, 0xffffffffu};  // min, max, bad, first read with a bad base
  SGX_CUDA(cudaMemcpyAsync(d_flags.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  SGX_CUDA(cudaMemsetAsync(d_woff.p, 0, (n + 1) * sizeof(uint64_t), ctx->stream));
  SGX_LAUNCH(ctx, reads_len_kernel, grid_for(n, 256), 256, 0, (const uint64_t*)d_off.p, n, d_len.p, d_woff.p,
             d_flags.p, d_flags.p + 2);
  // woff[i] = words before read i ; woff[n] = total words (the scan of n+1 elements with a trailing 0)
  scan_exclusive(ctx, d_woff.p, n + 1, ws.p, (uint64_t*)nullptr);
  uint64_t n_words = 0;
  uint32_t flags[4];
  SGX_CUDA(cudaMemcpyAsync(&n_words, d_woff.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  SGX_REQUIRE((flags[2] & 1u) == 0, SGX_EINVAL, "sgx_reads_pack: empty read or non-monotonic offsets");
  DevBuf<uint32_t> d_words;
  d_words.alloc_owned(n_words + 4);  // padded: word-level readers may touch one word past a read
  SGX_CUDA(cudaMemsetAsync(d_words.p + n_words, 0, 4 * sizeof(uint32_t), ctx->stream));
  // the offsets handed to the pack kernel are relative to the uploaded slice
  const uint64_t base = off[0];
  if (base != 0) {
    // rebase on the host copy is cheaper than a kernel: upload rebased offsets
    std::vector<uint64_t> rb(n + 1);
    for (uint64_t i = 0; i <= n; ++i) rb[i] = off[i] - base;
    SGX_CUDA(cudaMemcpyAsync(d_off.p, rb.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  SGX_LAUNCH(ctx, reads_pack_kernel, grid_for(n * 32, 256), 256, 0, (const char*)d_ascii.p, (const uint64_t*)d_off.p,
             (const uint64_t*)d_woff.p, n, d_words.p, d_flags.p + 2);
  SGX_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flags[2] & 2u) {
    char msg[200];
    snprintf(msg, sizeof(msg), "sgx_reads_pack: base other than A/C/G/T (first in read %u of this call; the hot path takes "
                               "preprocessed reads, as `siga preprocess` writes them)", flags[3]);
    throw Error{SGX_EINVAL, msg};
  }
  sgx_reads* r = new sgx_reads();
  r->ctx = ctx;
  r->n = n;
  r->bases = total;
  r->n_words = n_words;
  r->min_len = flags[0];
  r->max_len = flags[1];
  r->words = d_words.take();
  r->woff = d_woff.take();
  r->len = d_len.take();
  return r;
}

void reads_destroy(sgx_reads* r) {
  if (!r) return;
  cudaFreeAsync(r->words, r->ctx->stream);
  cudaFreeAsync(r->woff, r->ctx->stream);
  cudaFreeAsync(r->len, r->ctx->stream);
  delete r;
}

}  // namespace sgx
