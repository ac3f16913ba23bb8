// ASCII reads -> 2-bit device read store.  Input side of both hot paths; stands where the reference holds a
// std::vector<DNASeq> of std::strings (src/kseq.cpp:230-257).
#include "reads.cuh"

#include <chrono>
#include <string>

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
                                                         const uint64_t* __restrict__ woff, uint64_t first, uint64_t n,
                                                         uint32_t* __restrict__ words, uint32_t* __restrict__ bad) {
  const uint64_t warp = first + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);  // reads [first, n)
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
  d_words.alloc(n_words + 4);  // padded: word-level readers may touch one word past a read
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
             (const uint64_t*)d_woff.p, (uint64_t)0, n, d_words.p, d_flags.p + 2);
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
  r->words_buf.swap(d_words);
  r->words = r->words_buf.p;
  r->woff = d_woff.take();
  r->len = d_len.take();
  return r;
}

// ---- asynchronous pack ---------------------------------------------------------------------------------------------------
// Host -> device streaming of the read store: the ASCII bases cross PCIe in chunks on the
// context's second stream, each chunk packed as soon as it has landed, while the first stream already works -- the
// overlap search of a batch waits only for the chunks that hold its query reads (reads_wait_range), the first kernel
// that needs target reads waits for all of them.
sgx_reads* reads_pack_async(sgx_ctx* ctx, const char* ascii, const uint64_t* off, uint64_t n) {
  SGX_REQUIRE(ascii != nullptr && off != nullptr, SGX_EINVAL, "sgx_reads_pack: null input");
  SGX_REQUIRE(n > 0 && n < 0xffffffffull, SGX_EINVAL, "sgx_reads_pack: read count must be in [1, 2^32-1)");
  const uint64_t total = off[n] - off[0];
  SGX_REQUIRE(off[n] >= off[0] && total > 0, SGX_EINVAL, "sgx_reads_pack: bad offsets");
  const bool trace = getenv("SGX_TRACE_PACK") != nullptr;  // diagnostics: host time of every phase, to stderr
  struct Watch {
    bool on;
    std::chrono::steady_clock::time_point t;
    std::string log;
    void lap(const char* what) {
      if (!on) return;
      const auto now = std::chrono::steady_clock::now();
      char b[64];
      snprintf(b, sizeof(b), " %s %.2f", what, std::chrono::duration<double, std::milli>(now - t).count());
      log += b;
      t = now;
    }
  } watch{trace, std::chrono::steady_clock::now(), std::string()};
  sgx_reads* r = new sgx_reads();
  struct Guard {
    sgx_reads* r;
    ~Guard() {
      if (r) reads_destroy(r);
    }
  } guard{r};
  r->ctx = ctx;
  r->n = n;
  r->bases = total;
  r->words = nullptr;
  r->woff = nullptr;
  r->len = nullptr;
  DevBuf<uint64_t> d_woff, ws(scan_ws_elems(n + 1));
  DevBuf<uint32_t> d_len;
  d_woff.alloc_owned(n + 1);
  d_len.alloc_owned(n);
  watch.lap("alloc_small");
  r->staging.alloc(total);
  watch.lap("alloc_staging");
  r->staging_off.alloc(n + 1);
  r->flags.alloc(4);
  watch.lap("alloc_off");
  const uint64_t base = off[0];
  std::vector<uint64_t> rb;
  const uint64_t* off0 = off;
  if (base != 0) {  // offsets relative to the uploaded slice
    rb.resize(n + 1);
    for (uint64_t i = 0; i <= n; ++i) rb[i] = off[i] - base;
    off0 = rb.data();
  }
  SGX_CUDA(cudaMemcpyAsync(r->staging_off.p, off0, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  const uint32_t init[4] = {0xffffffffu, 0u, 0u, 0xffffffffu};
  SGX_CUDA(cudaMemcpyAsync(r->flags.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  SGX_CUDA(cudaMemsetAsync(d_woff.p, 0, (n + 1) * sizeof(uint64_t), ctx->stream));
  SGX_LAUNCH(ctx, reads_len_kernel, grid_for(n, 256), 256, 0, (const uint64_t*)r->staging_off.p, n, d_len.p, d_woff.p, r->flags.p,
             r->flags.p + 2);
  scan_exclusive(ctx, d_woff.p, n + 1, ws.p, (uint64_t*)nullptr);
  uint64_t n_words = 0;
  uint32_t flags[4];
  SGX_CUDA(cudaMemcpyAsync(&n_words, d_woff.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGX_CUDA(cudaMemcpyAsync(flags, r->flags.p, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
  watch.lap("enqueue_len");
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));  // rb / init are stack or local storage; n_words sizes the store
  watch.lap("sync");
  SGX_REQUIRE((flags[2] & 1u) == 0, SGX_EINVAL, "sgx_reads_pack: empty read or non-monotonic offsets");
  r->words_buf.alloc(n_words + 4);
  SGX_CUDA(cudaMemsetAsync(r->words_buf.p + n_words, 0, 4 * sizeof(uint32_t), ctx->stream));
  r->n_words = n_words;
  r->min_len = flags[0];
  r->max_len = flags[1];
  r->words = r->words_buf.p;
  r->woff = d_woff.take();
  r->len = d_len.take();
  // chunks of about 32 MB of bases on the second stream, which starts after the allocations above
  watch.lap("alloc_words");
  sgx_ctx* aux = ctx_aux(ctx);
  cudaEvent_t ready;
  SGX_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  SGX_CUDA(cudaEventRecord(ready, ctx->stream));
  SGX_CUDA(cudaStreamWaitEvent(aux->stream, ready, 0));
  cudaEventDestroy(ready);
  const uint64_t mean = std::max<uint64_t>(1, total / n);
  r->chunk_reads = std::max<uint64_t>(1024, ((uint64_t)32 << 20) / mean);
  const uint64_t nchunks = (n + r->chunk_reads - 1) / r->chunk_reads;
  r->pending = true;
  for (uint64_t c = 0; c < nchunks; ++c) {
    const uint64_t c0 = c * r->chunk_reads, c1 = std::min(n, c0 + r->chunk_reads);
    const uint64_t b0 = off[c0] - base, b1 = off[c1] - base;
    SGX_CUDA(cudaMemcpyAsync(r->staging.p + b0, ascii + off[c0], b1 - b0, cudaMemcpyHostToDevice, aux->stream));
    SGX_LAUNCH(aux, reads_pack_kernel, grid_for((c1 - c0) * 32, 256), 256, 0, (const char*)r->staging.p,
               (const uint64_t*)r->staging_off.p, (const uint64_t*)r->woff, c0, c1, r->words, r->flags.p + 2);
    cudaEvent_t ev;
    SGX_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    r->chunk_ev.push_back(ev);
    SGX_CUDA(cudaEventRecord(ev, aux->stream));
  }
  ctx->launches += nchunks;
  watch.lap("enqueue_chunks");
  if (trace) fprintf(stderr, "[sgx pack_async]%s\n", watch.log.c_str());
  guard.r = nullptr;
  return r;
}

void reads_wait_range(cudaStream_t stream, const sgx_reads* r, uint64_t lo, uint64_t hi) {
  if (!r->pending || hi <= lo) return;
  const uint64_t c0 = lo / r->chunk_reads, c1 = std::min<uint64_t>((hi - 1) / r->chunk_reads, r->chunk_ev.size() - 1);
  // the chunks are recorded in order on one stream: waiting for the last one of the range covers the earlier ones
  (void)c0;
  SGX_CUDA(cudaStreamWaitEvent(stream, r->chunk_ev[c1], 0));
}

void reads_settle(const sgx_reads* r) {
  if (r->invalid) throw Error{SGX_EINVAL, "sgx_reads: this read store was rejected (base other than A/C/G/T)"};
  if (!r->pending) return;
  SGX_CUDA(cudaEventSynchronize(r->chunk_ev.back()));
  uint32_t flags[4];
  SGX_CUDA(cudaMemcpy(flags, r->flags.p, sizeof(flags), cudaMemcpyDeviceToHost));
  for (cudaEvent_t e : r->chunk_ev) cudaEventDestroy(e);
  r->chunk_ev.clear();
  r->pending = false;
  r->staging.release();
  r->staging_off.release();
  r->flags.release();
  if (flags[2] & 2u) {
    r->invalid = true;
    char msg[200];
    snprintf(msg, sizeof(msg), "sgx_reads_pack: base other than A/C/G/T (first in read %u of this call; the hot path takes "
                               "preprocessed reads, as `siga preprocess` writes them)", flags[3]);
    throw Error{SGX_EINVAL, msg};
  }
}

void reads_destroy(sgx_reads* r) {
  if (!r) return;
  if (r->pending) {  // chunks may still be in flight on the second stream
    if (!r->chunk_ev.empty()) cudaEventSynchronize(r->chunk_ev.back());
    for (cudaEvent_t e : r->chunk_ev) cudaEventDestroy(e);
    r->chunk_ev.clear();
    r->pending = false;
  }
  {
    // the buffers go back to the pool / work-area cache of the store's own context, whatever context the calling
    // thread last used
    sgx_ctx* saved_ctx = g_ctx;
    cudaStream_t saved_stream = g_stream;
    g_ctx = r->ctx;
    g_stream = r->ctx->stream;
    r->staging.release();
    r->staging_off.release();
    r->flags.release();
    r->words_buf.release();
    g_ctx = saved_ctx;
    g_stream = saved_stream;
  }
  if (r->woff) cudaFreeAsync(r->woff, r->ctx->stream);
  if (r->len) cudaFreeAsync(r->len, r->ctx->stream);
  delete r;
}

}  // namespace sgx
