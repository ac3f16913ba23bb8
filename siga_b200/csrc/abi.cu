// extern "C" entry points of libsiga_b200.so (include/siga_b200.h).  Thin: argument checks, error mapping, timing.
#include <mutex>
#include <new>
#include <thread>
#include <unordered_map>

#include "index.cuh"
#include "correct.cuh"
#include "overlap.cuh"

static std::string g_init_error;
static std::mutex g_pin_mutex;
static std::unordered_map<void*, std::pair<sgx_ctx*, size_t> > g_pin_owner;

void sgx_set_error(sgx_ctx* ctx, const char* msg) {
  if (ctx) {
    ctx->err = msg;
  } else {
    g_init_error = msg;
  }
}

namespace sgx {

thread_local cudaStream_t g_stream = nullptr;
thread_local sgx_ctx* g_ctx = nullptr;

void* ws_get(size_t bytes) {
  sgx_ctx* ctx = g_ctx;
  int best = -1;
  for (size_t i = 0; i < ctx->ws_cache.size(); ++i) {
    const sgx_ctx::WsEntry& e = ctx->ws_cache[i];
    if (!e.in_use && e.bytes >= bytes && e.bytes <= bytes + bytes / 2 + (1u << 20) &&
        (best < 0 || e.bytes < ctx->ws_cache[best].bytes))
      best = (int)i;
  }
  if (best >= 0) {
    ctx->ws_cache[best].in_use = true;
    return ctx->ws_cache[best].p;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    // out of memory: drop every idle cached area and retry once
    cudaGetLastError();
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < ctx->ws_cache.size();) {
      if (!ctx->ws_cache[i].in_use) {
        cudaFree(ctx->ws_cache[i].p);
        ctx->ws_cache.erase(ctx->ws_cache.begin() + i);
      } else {
        ++i;
      }
    }
    SGX_CUDA(cudaMalloc(&p, bytes));
  }
  ctx->ws_cache.push_back(sgx_ctx::WsEntry{p, bytes, true});
  return p;
}

void ws_put(void* p) {
  sgx_ctx* ctx = g_ctx;
  if (!ctx) return;
  for (auto& e : ctx->ws_cache)
    if (e.p == p) e.in_use = false;
}

void* pinned_get(sgx_ctx* ctx, size_t bytes) {
  if (bytes == 0) bytes = 1;
  {
    std::lock_guard<std::mutex> lk(g_pin_mutex);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < ctx->pinned_free.size(); ++i) {
      if (ctx->pinned_free[i].second >= bytes &&
          (best == (size_t)-1 || ctx->pinned_free[i].second < ctx->pinned_free[best].second))
        best = i;
    }
    if (best != (size_t)-1 && ctx->pinned_free[best].second <= 2 * bytes + (1u << 20)) {
      void* p = ctx->pinned_free[best].first;
      ctx->pinned_free.erase(ctx->pinned_free.begin() + best);
      return p;
    }
  }
  void* p = nullptr;
  SGX_CUDA(cudaMallocHost(&p, bytes));
  std::lock_guard<std::mutex> lk(g_pin_mutex);
  g_pin_owner[p] = std::make_pair(ctx, bytes);
  return p;
}

void pinned_put(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pin_mutex);
  auto it = g_pin_owner.find(p);
  if (it == g_pin_owner.end()) return;
  it->second.first->pinned_free.push_back(std::make_pair(p, it->second.second));
}

sgx_ctx* ctx_aux(sgx_ctx* ctx) {
  if (!ctx->aux) {
    sgx_ctx* a = new sgx_ctx();
    a->device = ctx->device;
    a->sm_count = ctx->sm_count;
    a->l2_bytes = ctx->l2_bytes;
    a->launches = 0;
    a->flush_buf = nullptr;
    a->flush_bytes = 0;
    cudaError_t e = cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete a;
      SGX_CUDA(e);
    }
    ctx->aux = a;
  }
  return ctx->aux;
}

__global__ void probe_kernel(int* out) {
  if (out) *out = 1;
}

// Each thread issues `per_thread` dependent-free random 16-byte loads, 8 consecutive lanes covering one 128-byte
// line, i.e. the access shape of a rank query without the arithmetic.
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ buf, uint64_t n_lines, uint64_t per_group,
                                                     uint64_t seed, unsigned long long* __restrict__ sink) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t gid = tid >> 3;
  const int lane = (int)(tid & 7);
  uint64_t x = seed ^ (gid * 0x9E3779B97F4A7C15ull);
  uint32_t acc = 0;
  for (uint64_t i = 0; i < per_group; i += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x ^= x << 13;
      x ^= x >> 7;
      x ^= x << 17;
      const uint64_t line = x % n_lines;
      v[u] = __ldg(buf + line * 8 + lane);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u].x ^ v[u].w;
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

__global__ void fill_kernel(uint4* buf, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = i; k < n; k += stride) buf[k] = make_uint4((uint32_t)k, 1u, 2u, 3u);
}

}  // namespace sgx

using namespace sgx;

extern "C" {

int sgx_abi_version(void) { return SGX_ABI_VERSION; }

int sgx_init(int device, sgx_ctx** out) {
  if (!out) return SGX_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_init_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libsiga_b200 has no CPU fallback)";
    cudaGetLastError();
    return SGX_ENODEV;
  }
  if (device < 0 || device >= count) {
    g_init_error = "device index out of range";
    return SGX_EINVAL;
  }
  sgx_ctx* ctx = nullptr;
  try {
    SGX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SGX_CUDA(cudaGetDeviceProperties(&prop, device));
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, probe_kernel);
    if (e != cudaSuccess) {
      cudaGetLastError();
      char b[512];
      snprintf(b, sizeof(b), "device %d (%s, sm_%d%d) cannot run the sm_100a kernel image: %s (no fallback)", device,
               prop.name, prop.major, prop.minor, cudaGetErrorString(e));
      g_init_error = b;
      return SGX_ENODEV;
    }
    ctx = new sgx_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    ctx->launches = 0;
    ctx->flush_buf = nullptr;
    ctx->flush_bytes = 0;
    SGX_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {
      // keep freed device memory in the pool (see DevBuf)
      cudaMemPool_t pool;
      SGX_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
      uint64_t keep = UINT64_MAX;
      SGX_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    probe_kernel<<<1, 1, 0, ctx->stream>>>(nullptr);
    SGX_CUDA(cudaGetLastError());
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (const sgx::Error& err) {
    g_init_error = err.msg;
    delete ctx;
    cudaGetLastError();
    return err.code == SGX_ECUDA ? SGX_ENODEV : err.code;
  }
  *out = ctx;
  return SGX_OK;
}

void sgx_shutdown(sgx_ctx* ctx) {
  if (!ctx) return;
  if (ctx->aux) {
    sgx_shutdown(ctx->aux);
    ctx->aux = nullptr;
  }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  {
    std::lock_guard<std::mutex> lk(g_pin_mutex);
    for (auto it = g_pin_owner.begin(); it != g_pin_owner.end();) {
      if (it->second.first == ctx) {
        cudaFreeHost(it->first);
        it = g_pin_owner.erase(it);
      } else {
        ++it;
      }
    }
  }
  if (ctx->flush_buf) cudaFreeAsync(ctx->flush_buf, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  for (auto& e : ctx->ws_cache) cudaFree(e.p);
  ctx->ws_cache.clear();
  if (sgx::g_ctx == ctx) sgx::g_ctx = nullptr;
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* sgx_last_error(const sgx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int sgx_sm_count(const sgx_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int sgx_reads_pack(sgx_ctx* ctx, const char* ascii, const uint64_t* off, uint64_t n, sgx_reads** out) {
  if (!ctx || !out) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  *out = reads_pack(ctx, ascii, off, n);
  SGX_ENTRY_END(ctx)
}

int sgx_reads_pack_async(sgx_ctx* ctx, const char* ascii, const uint64_t* off, uint64_t n, sgx_reads** out) {
  if (!ctx || !out) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  *out = reads_pack_async(ctx, ascii, off, n);
  SGX_ENTRY_END(ctx)
}

int sgx_reads_wait(sgx_ctx* ctx, const sgx_reads* reads) {
  if (!ctx || !reads) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  reads_settle(reads);
  SGX_ENTRY_END(ctx)
}

void sgx_reads_free(sgx_reads* reads) {
  if (reads) cudaSetDevice(reads->ctx->device);
  reads_destroy(reads);
}

uint64_t sgx_reads_count(const sgx_reads* reads) { return reads ? reads->n : 0; }
uint64_t sgx_reads_bases(const sgx_reads* reads) { return reads ? reads->bases : 0; }

int sgx_index_build(sgx_ctx* ctx, const sgx_reads* reads, int order, int reversed, sgx_index_out* out) {
  if (!ctx || !reads || !out) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  SGX_REQUIRE(order == SGX_ORDER_SAIS || order == SGX_ORDER_SAIS2, SGX_EINVAL, "sgx_index_build: unknown order");
  reads_settle(reads);
  index_build(ctx, reads, order, reversed != 0, out);
  SGX_ENTRY_END(ctx)
}

// Both indices of a read set (what one `siga index` run builds, src/indexer.cpp:25-77) at the same time: the forward
// build on the context's stream, the build of the reversed reads on a second stream of the same device, driven by a
// second host thread.  The column passes of one build are chains of small dependent kernels; two independent chains
// fill each other's gaps.
int sgx_index_build_pair(sgx_ctx* ctx, const sgx_reads* reads, int order, sgx_index_out* out_fwd, sgx_index_out* out_rev) {
  if (!ctx || !reads || !out_fwd || !out_rev) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  SGX_REQUIRE(order == SGX_ORDER_SAIS || order == SGX_ORDER_SAIS2, SGX_EINVAL, "sgx_index_build_pair: unknown order");
  reads_settle(reads);
  sgx_ctx* aux = ctx_aux(ctx);
  // the second stream starts after everything queued on the first (the read store may still be in flight there)
  cudaEvent_t ready;
  SGX_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  SGX_CUDA(cudaEventRecord(ready, ctx->stream));
  SGX_CUDA(cudaStreamWaitEvent(aux->stream, ready, 0));
  cudaEventDestroy(ready);
  const uint64_t aux_launches0 = aux->launches;
  int rc_rev = SGX_OK;
  std::string err_rev;
  std::thread second([&]() {
    try {
      SGX_CUDA(cudaSetDevice(aux->device));
      sgx::g_stream = aux->stream;
      sgx::g_ctx = aux;
      index_build(aux, reads, order, true, out_rev);
    } catch (const sgx::Error& e) {
      rc_rev = e.code;
      err_rev = e.msg;
      cudaGetLastError();
    } catch (const std::bad_alloc&) {
      rc_rev = SGX_ENOMEM;
      err_rev = "host allocation failed";
    }
  });
  int rc_fwd = SGX_OK;
  std::string err_fwd;
  try {
    index_build(ctx, reads, order, false, out_fwd);
  } catch (const sgx::Error& e) {
    rc_fwd = e.code;
    err_fwd = e.msg;
    cudaGetLastError();
  }
  second.join();
  ctx->launches += aux->launches - aux_launches0;
  if (rc_fwd != SGX_OK || rc_rev != SGX_OK) {
    if (rc_fwd == SGX_OK) sgx_index_out_free(out_fwd);
    if (rc_rev == SGX_OK) sgx_index_out_free(out_rev);
    throw sgx::Error{rc_fwd != SGX_OK ? rc_fwd : rc_rev, rc_fwd != SGX_OK ? err_fwd : err_rev};
  }
  SGX_ENTRY_END(ctx)
}

int sgx_comm_unique_id(uint8_t id[128]) {
  if (!id) return SGX_EINVAL;
  sgx_ctx* none = nullptr;
  SGX_ENTRY_BEGIN(none)
  comm_unique_id(id);
  SGX_ENTRY_END(none)
}

int sgx_comm_init(sgx_ctx* ctx, const uint8_t id[128], int rank, int world, sgx_comm** out) {
  if (!ctx || !id || !out) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  *out = comm_create_nccl(ctx, id, rank, world);
  SGX_ENTRY_END(ctx)
}

int sgx_comm_init_local(sgx_ctx* ctx, int shards, sgx_comm** out) {
  if (!ctx || !out) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  *out = comm_create_local(ctx, shards);
  SGX_ENTRY_END(ctx)
}

void sgx_comm_free(sgx_comm* comm) {
  if (!comm) return;
  cudaSetDevice(comm->ctx->device);
  sgx::g_stream = comm->ctx->stream;  // comm_destroy's farewell barrier allocates a scratch word on the context's stream
  sgx::g_ctx = comm->ctx;
  try {
    comm_destroy(comm);
  } catch (const sgx::Error&) {
    cudaGetLastError();
  }
}

int sgx_comm_world(const sgx_comm* comm) { return comm ? comm->world : 0; }
int sgx_comm_rank(const sgx_comm* comm) { return comm ? comm->rank : -1; }

int sgx_index_build_sharded(sgx_ctx* ctx, sgx_comm* comm, const sgx_reads* reads, int order, int reversed,
                            sgx_index_out* out) {
  if (!ctx || !comm || !reads || !out || comm->ctx != ctx) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  SGX_REQUIRE(order == SGX_ORDER_SAIS || order == SGX_ORDER_SAIS2, SGX_EINVAL, "sgx_index_build_sharded: unknown order");
  reads_settle(reads);
  index_build_sharded(ctx, comm, reads, order, reversed != 0, false, out);
  SGX_ENTRY_END(ctx)
}

int sgx_index_build_sharded_segments(sgx_ctx* ctx, sgx_comm* comm, const sgx_reads* reads, int order, int reversed,
                            sgx_index_out* out) {
  if (!ctx || !comm || !reads || !out || comm->ctx != ctx) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  SGX_REQUIRE(order == SGX_ORDER_SAIS || order == SGX_ORDER_SAIS2, SGX_EINVAL, "sgx_index_build_sharded_segments: unknown order");
  reads_settle(reads);
  index_build_sharded(ctx, comm, reads, order, reversed != 0, true, out);
  SGX_ENTRY_END(ctx)
}

void sgx_index_out_free(sgx_index_out* out) {
  if (!out) return;
  pinned_put(out->rle);
  pinned_put(out->sai);
  out->rle = nullptr;
  out->sai = nullptr;
}

int sgx_fm_create(sgx_ctx* ctx, const uint8_t* rle, uint64_t n_runs, uint64_t n_suffixes, sgx_fm** out) {
  if (!ctx || !out || (!rle && n_runs)) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  DevBuf<uint8_t> d_rle(n_runs ? n_runs : 1), d_sym(n_suffixes ? n_suffixes + 32 : 32);
  if (n_runs) SGX_CUDA(cudaMemcpyAsync(d_rle.p, rle, n_runs, cudaMemcpyHostToDevice, ctx->stream));
  rle_expand(ctx, d_rle.p, n_runs, n_suffixes, d_sym.p);
  *out = fm_build_from_symbols(ctx, d_sym.p, n_suffixes);
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  SGX_ENTRY_END(ctx)
}

void sgx_fm_free(sgx_fm* fm) {
  if (!fm) return;
  cudaSetDevice(fm->ctx->device);
  cudaFreeAsync(fm->blocks, fm->ctx->stream);
  if (fm->sai) cudaFreeAsync(fm->sai, fm->ctx->stream);
  if (fm->inv) cudaFreeAsync(fm->inv, fm->ctx->stream);
  if (fm->ktab) cudaFreeAsync(fm->ktab, fm->ctx->stream);
  delete fm;
}

int sgx_fm_attach_reads(sgx_ctx* ctx, sgx_fm* fm, const uint32_t* sai, uint64_t n_strings, const sgx_reads* indexed_reads) {
  if (!ctx || !fm) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  fm_attach(ctx, fm, sai, n_strings, indexed_reads);
  SGX_ENTRY_END(ctx)
}

uint64_t sgx_fm_length(const sgx_fm* fm) { return fm ? fm->n : 0; }

int sgx_fm_pred(const sgx_fm* fm, uint64_t pred[5]) {
  if (!fm || !pred) return SGX_EINVAL;
  for (int i = 0; i < 5; ++i) pred[i] = fm->pred[i];
  return SGX_OK;
}

int sgx_fm_occ(sgx_ctx* ctx, const sgx_fm* fm, const uint64_t* pos, uint64_t n, uint64_t* counts) {
  if (!ctx || !fm || (!pos && n) || (!counts && n)) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  fm_occ_batch(ctx, fm, pos, n, counts);
  SGX_ENTRY_END(ctx)
}

int sgx_fm_count(sgx_ctx* ctx, const sgx_fm* fm, const char* ascii, const uint64_t* off, uint64_t n, uint64_t* counts) {
  if (!ctx || !fm || !ascii || !off || (!counts && n)) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  fm_count_batch(ctx, fm, ascii, off, n, counts);
  SGX_ENTRY_END(ctx)
}

int sgx_correct_batch(sgx_ctx* ctx, const sgx_fm* fm, const sgx_reads* reads, uint64_t first, uint64_t count, const char* qual,
                      const uint64_t* qual_off, const sgx_correct_params* params, sgx_corrected* out) {
  if (!ctx || !out) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  reads_settle(reads);
  correct_run(ctx, fm, reads, first, count, qual, qual_off, params, out);
  SGX_ENTRY_END(ctx)
}

void sgx_corrected_free(sgx_corrected* out) {
  if (!out) return;
  pinned_put(out->valid);
  pinned_put(out->fixes);
  out->valid = nullptr;
  out->fixes = nullptr;
}

int sgx_overlap_batch(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                      uint64_t count, uint32_t min_overlap, int irreducible, int rc, sgx_hits* out) {
  if (!ctx || !out) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  reads_settle(reads);
  overlap_run(ctx, fwd, rev, reads, first, count, min_overlap, irreducible, rc, out);
  SGX_ENTRY_END(ctx)
}

int sgx_overlap_host(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const char* ascii, const uint64_t* off,
                     uint64_t n, uint32_t min_overlap, int irreducible, int rc, sgx_hits* out) {
  if (!ctx || !out) return SGX_EINVAL;
  sgx_reads* reads = nullptr;
  int rcode = sgx_reads_pack(ctx, ascii, off, n, &reads);
  if (rcode != SGX_OK) return rcode;
  rcode = sgx_overlap_batch(ctx, fwd, rev, reads, 0, n, min_overlap, irreducible, rc, out);
  sgx_reads_free(reads);
  return rcode;
}

int sgx_overlap_edges(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                      uint64_t count, uint32_t min_overlap, int irreducible, int rc, const uint32_t* name_rank,
                      sgx_edges* out) {
  if (!ctx || !out) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  overlap_edges_run(ctx, fwd, rev, reads, first, count, min_overlap, irreducible, rc, name_rank, out);
  SGX_ENTRY_END(ctx)
}

int sgx_edges_gather(sgx_ctx* ctx, sgx_comm* comm, const sgx_edges* local, int root, sgx_edges* out) {
  if (!ctx || !comm || !local || !out || comm->ctx != ctx || root < 0 || root >= comm->world) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  SGX_REQUIRE(!comm->local || comm->world == 1, SGX_EINVAL, "sgx_edges_gather: needs an NCCL communicator");
  memset(out, 0, sizeof(*out));
  const int W = comm->world;
  EventTimer t(ctx->stream);
  t.start();
  // sizes of every rank
  DevBuf<uint64_t> d_sz(2);
  const uint64_t mine[2] = {local->n_reads, local->n_edges};
  SGX_CUDA(cudaMemcpyAsync(d_sz.p, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<uint64_t> sz(2 * (size_t)W);
  const uint64_t* szp = d_sz.p;
  comm_allgather_u64(comm, &szp, 2, sz.data());
  std::vector<uint64_t> fb(W), eb(W);
  uint64_t n_reads = 0, n_edges = 0;
  for (int r = 0; r < W; ++r) {
    fb[r] = sz[2 * r];
    eb[r] = sz[2 * r + 1] * sizeof(sgx_edge);
    n_reads += sz[2 * r];
    n_edges += sz[2 * r + 1];
  }
  const bool is_root = comm->rank == root || comm->local;
  DevBuf<uint8_t> d_f(local->n_reads ? local->n_reads : 1), d_e(local->n_edges ? local->n_edges * sizeof(sgx_edge) : 1);
  if (local->n_reads) SGX_CUDA(cudaMemcpyAsync(d_f.p, local->substring, local->n_reads, cudaMemcpyHostToDevice, ctx->stream));
  if (local->n_edges)
    SGX_CUDA(cudaMemcpyAsync(d_e.p, local->edges, local->n_edges * sizeof(sgx_edge), cudaMemcpyHostToDevice, ctx->stream));
  DevBuf<uint8_t> a_f, a_e;
  if (is_root) {
    a_f.alloc(n_reads ? n_reads : 1);
    a_e.alloc(n_edges ? n_edges * sizeof(sgx_edge) : 1);
  }
  const uint8_t* pf = d_f.p;
  const uint8_t* pe = d_e.p;
  comm_gatherv_bytes(comm, &pf, fb.data(), root, a_f.p);
  comm_gatherv_bytes(comm, &pe, eb.data(), root, a_e.p);
  if (is_root) {
    out->n_reads = n_reads;
    out->n_edges = n_edges;
    out->substring = (uint8_t*)pinned_get(ctx, n_reads ? n_reads : 1);
    out->edges = (sgx_edge*)pinned_get(ctx, n_edges ? n_edges * sizeof(sgx_edge) : 1);
    SGX_CUDA(cudaMemcpyAsync(out->substring, a_f.p, n_reads, cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaMemcpyAsync(out->edges, a_e.p, n_edges * sizeof(sgx_edge), cudaMemcpyDeviceToHost, ctx->stream));
  }
  t.stop();
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  out->ms_total = t.ms();
  SGX_ENTRY_END(ctx)
}

void sgx_edges_free(sgx_edges* edges) {
  if (!edges) return;
  pinned_put(edges->substring);
  pinned_put(edges->edges);
  edges->substring = nullptr;
  edges->edges = nullptr;
}

void sgx_hits_free(sgx_hits* hits) {
  if (!hits) return;
  pinned_put(hits->substring);
  pinned_put(hits->block_off);
  pinned_put(hits->blocks);
  hits->substring = nullptr;
  hits->block_off = nullptr;
  hits->blocks = nullptr;
}

int sgx_bench_gather(sgx_ctx* ctx, uint64_t bytes, uint64_t n_loads, double* fetches_per_s) {
  if (!ctx || !fetches_per_s || bytes < 4096) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  const uint64_t n_lines = bytes / 128;
  DevBuf<uint4> buf(n_lines * 8);
  DevBuf<unsigned long long> sink(1);
  SGX_CUDA(cudaMemsetAsync(sink.p, 0, sizeof(unsigned long long), ctx->stream));
  SGX_LAUNCH(ctx, fill_kernel, (unsigned)ctx->sm_count * 8, 256, 0, buf.p, n_lines * 8);
  const unsigned grid = (unsigned)ctx->sm_count * 8;
  const uint64_t groups = (uint64_t)grid * 256 / 8;
  uint64_t per_group = (n_loads + groups - 1) / groups;
  per_group = (per_group + 3) & ~3ull;
  if (per_group < 4) per_group = 4;
  SGX_LAUNCH(ctx, gather_kernel, grid, 256, 0, (const uint4*)buf.p, n_lines, per_group, 0x1234567ull, sink.p);  // warm-up
  EventTimer t(ctx->stream);
  t.start();
  SGX_LAUNCH(ctx, gather_kernel, grid, 256, 0, (const uint4*)buf.p, n_lines, per_group, 0xABCDEF01ull, sink.p);
  t.stop();
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  *fetches_per_s = (double)(groups * per_group) / (t.ms() * 1e-3);
  SGX_ENTRY_END(ctx)
}

int sgx_flush_l2(sgx_ctx* ctx) {
  if (!ctx) return SGX_EINVAL;
  SGX_ENTRY_BEGIN(ctx)
  const size_t want = std::max<size_t>(2 * ctx->l2_bytes, (size_t)256 << 20);
  if (ctx->flush_bytes < want) {
    if (ctx->flush_buf) cudaFreeAsync(ctx->flush_buf, ctx->stream);
    ctx->flush_buf = nullptr;
    SGX_CUDA(cudaMallocAsync(&ctx->flush_buf, want, ctx->stream));
    ctx->flush_bytes = want;
  }
  SGX_LAUNCH(ctx, fill_kernel, (unsigned)ctx->sm_count * 8, 256, 0, (uint4*)ctx->flush_buf, (uint64_t)(want / 16));
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  SGX_ENTRY_END(ctx)
}

}  // extern "C"
