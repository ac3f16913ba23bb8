// Context, error plumbing and small device-memory helpers shared by the .cu files of libsiga_b200.
#ifndef SGX_DEVICE_CUH_
#define SGX_DEVICE_CUH_

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sgx_common.h"

struct sgx_ctx {
  int device;
  int sm_count;
  size_t l2_bytes;
  cudaStream_t stream;
  std::string err;
  uint64_t launches;  // kernels launched since the counter was last reset by an entry point
  void* flush_buf;
  size_t flush_bytes;
  // pinned host buffers handed back by *_free calls, reused by later calls (pinned allocation is slow)
  std::vector<std::pair<void*, size_t> > pinned_free;
  // large device work areas, kept for the lifetime of the context and handed out again to later calls
  struct WsEntry {
    void* p;
    size_t bytes;
    bool in_use;
  };
  std::vector<WsEntry> ws_cache;
  // second stream of the same device with its own work-area cache, for entry points that run two independent jobs
  // side by side (sgx_index_build_pair); created on first use, owned by this context
  sgx_ctx* aux = nullptr;
};

namespace sgx {

struct Error {
  int code;
  std::string msg;
};

#define SGX_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      char _b[512];                                                                                      \
      snprintf(_b, sizeof(_b), "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      throw sgx::Error{_e == cudaErrorMemoryAllocation ? SGX_ENOMEM : SGX_ECUDA, _b};                    \
    }                                                                                                    \
  } while (0)

#define SGX_REQUIRE(cond, code, text)        \
  do {                                       \
    if (!(cond)) throw sgx::Error{code, text}; \
  } while (0)

// Launch bookkeeping: every kernel launch of the library goes through this so that `launches` is a count, not a guess.
#define SGX_LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
  do {                                                                          \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);            \
    (ctx)->launches++;                                                          \
    SGX_CUDA(cudaGetLastError());                                               \
  } while (0)

// Stream of the context whose entry point is executing on this thread (set by SGX_ENTRY_BEGIN); device buffers are
// allocated stream-ordered from the device's memory pool, whose release threshold sgx_init raises so that the
// multi-GB work areas of one call are reused by the next instead of going back to the driver (cudaMalloc/cudaFree
// of those sizes cost up to hundreds of milliseconds each).
extern thread_local cudaStream_t g_stream;
extern thread_local sgx_ctx* g_ctx;
// Work areas of >= kWsMinBytes come from the context's cache (first use: cudaMalloc; afterwards reused, so a
// repeated call allocates nothing); smaller ones from the stream-ordered pool.  Everything runs on the one stream
// of the context, so handing a released area to the next user is ordered by the stream.
static const size_t kWsMinBytes = (size_t)4 << 20;
void* ws_get(size_t bytes);
void ws_put(void* p);

// RAII device buffer.
template <class T>
struct DevBuf {
  T* p;
  size_t n;
  bool cached;
  DevBuf() : p(nullptr), n(0), cached(false) {}
  explicit DevBuf(size_t count) : p(nullptr), n(0), cached(false) { alloc(count); }
  ~DevBuf() { release(); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void alloc(size_t count) {
    release();
    n = count;
    cached = false;
    if (!count) return;
    if (g_ctx && count * sizeof(T) >= kWsMinBytes) {
      p = (T*)ws_get(count * sizeof(T));
      cached = true;
    } else {
      SGX_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), g_stream));
    }
  }
  void release() {
    if (p) {
      if (cached) ws_put(p);
      else cudaFreeAsync(p, g_stream);
    }
    p = nullptr;
    n = 0;
    cached = false;
  }
  // Ownership leaves the buffer (long-lived objects: read store, rank blocks).  Such buffers must come from the
  // stream-ordered pool because they are released with cudaFreeAsync: use alloc_owned() for them.
  T* take() {
    T* q = p;
    p = nullptr;
    n = 0;
    return q;
  }
  void swap(DevBuf& o) {
    T* tp = p;
    p = o.p;
    o.p = tp;
    size_t tn = n;
    n = o.n;
    o.n = tn;
    bool tc = cached;
    cached = o.cached;
    o.cached = tc;
  }
  void alloc_owned(size_t count) {
    release();
    n = count;
    cached = false;
    if (count) SGX_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), g_stream));
  }
};

template <class T>
struct PinnedBuf {
  T* p;
  size_t n;
  PinnedBuf() : p(nullptr), n(0) {}
  ~PinnedBuf() { release(); }
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  void alloc(size_t count) {
    release();
    n = count;
    SGX_CUDA(cudaMallocHost((void**)&p, (count ? count : 1) * sizeof(T)));
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
  T* take() {
    T* q = p;
    p = nullptr;
    n = 0;
    return q;
  }
};

struct EventTimer {
  cudaEvent_t a, b;
  cudaStream_t s;
  explicit EventTimer(cudaStream_t st) : s(st) {
    SGX_CUDA(cudaEventCreate(&a));
    SGX_CUDA(cudaEventCreate(&b));
  }
  ~EventTimer() {
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
  void start() { SGX_CUDA(cudaEventRecord(a, s)); }
  void stop() { SGX_CUDA(cudaEventRecord(b, s)); }
  // Valid after the stream (or event b) has been synchronised.
  double ms() {
    float t = 0;
    SGX_CUDA(cudaEventSynchronize(b));
    SGX_CUDA(cudaEventElapsedTime(&t, a, b));
    return (double)t;
  }
};

// Accumulates the device time of many launches of one kernel family (event pairs recorded around each launch).
struct KernelClock {
  std::vector<cudaEvent_t> ev;
  cudaStream_t s;
  explicit KernelClock(cudaStream_t st) : s(st) {}
  ~KernelClock() {
    for (auto e : ev) cudaEventDestroy(e);
  }
  void tick() {
    cudaEvent_t e;
    SGX_CUDA(cudaEventCreate(&e));
    SGX_CUDA(cudaEventRecord(e, s));
    ev.push_back(e);
  }
  // Sum over consecutive (start, stop) pairs; call after synchronising the stream.
  double total_ms() {
    double t = 0;
    for (size_t i = 0; i + 1 < ev.size(); i += 2) {
      float x = 0;
      SGX_CUDA(cudaEventElapsedTime(&x, ev[i], ev[i + 1]));
      t += x;
    }
    return t;
  }
  size_t pairs() const { return ev.size() / 2; }
};

inline unsigned grid_for(uint64_t n, unsigned per_block) {
  uint64_t g = (n + per_block - 1) / per_block;
  if (g == 0) g = 1;
  return (unsigned)g;
}

}  // namespace sgx

// Entry-point wrapper: turns sgx::Error / std::bad_alloc into an ABI status + message.
#define SGX_ENTRY_BEGIN(ctx)                          \
  try {                                               \
    if (ctx) {                                        \
      SGX_CUDA(cudaSetDevice((ctx)->device));         \
      sgx::g_stream = (ctx)->stream;                  \
      sgx::g_ctx = (ctx);                             \
    }
#define SGX_ENTRY_END(ctx)                      \
  }                                             \
  catch (const sgx::Error& e) {                 \
    sgx_set_error((ctx), e.msg.c_str());        \
    cudaGetLastError();                         \
    return e.code;                              \
  }                                             \
  catch (const std::bad_alloc&) {               \
    sgx_set_error((ctx), "host allocation failed"); \
    return SGX_ENOMEM;                          \
  }                                             \
  return SGX_OK;

void sgx_set_error(sgx_ctx* ctx, const char* msg);

namespace sgx {
// The context's second stream (ctx->aux), created on first use.
sgx_ctx* ctx_aux(sgx_ctx* ctx);
// Pinned host memory pool of a context (outputs of the ABI live here).
void* pinned_get(sgx_ctx* ctx, size_t bytes);
void pinned_put(void* p);  // returns the buffer to the pool of the context it came from
}  // namespace sgx

#endif  // SGX_DEVICE_CUH_
