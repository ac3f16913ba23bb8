// Shard communicator: NCCL (one process per GPU) or local (virtual shards on one context).  See comm.cuh.
#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <mutex>

namespace sgx {

// ---- NCCL, resolved at run time -------------------------------------------------------------------------------------

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mutex;

static NcclApi& nccl_api() {
  std::lock_guard<std::mutex> lk(g_nccl_mutex);
  if (g_nccl.handle) return g_nccl;
  // RTLD_NOLOAD first: reuse the NCCL the host process already brought (torch bundles its own libnccl.so.2)
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    const char* why = dlerror();  // one call: dlerror() clears the message it returns
    throw Error{SGX_ENODEV, std::string("NCCL is not available: ") + (why ? why : "dlopen failed")};
  }
#define SGX_NCCL_SYM(field, name)                                                         \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                              \
  if (!g_nccl.field) throw Error{SGX_ENODEV, std::string("NCCL symbol missing: ") + name};
  SGX_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  SGX_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  SGX_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  SGX_NCCL_SYM(AllGather, "ncclAllGather")
  SGX_NCCL_SYM(AllReduce, "ncclAllReduce")
  SGX_NCCL_SYM(Send, "ncclSend")
  SGX_NCCL_SYM(Recv, "ncclRecv")
  SGX_NCCL_SYM(GroupStart, "ncclGroupStart")
  SGX_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  SGX_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  SGX_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef SGX_NCCL_SYM
  g_nccl.handle = h;
  return g_nccl;
}

#define SGX_NCCL(expr)                                                                                      \
  do {                                                                                                      \
    ncclResult_t _r = (expr);                                                                               \
    if (_r != ncclSuccess) {                                                                                \
      char _b[512];                                                                                         \
      snprintf(_b, sizeof(_b), "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, nccl_api().GetErrorString(_r)); \
      throw sgx::Error{SGX_ECUDA, _b};                                                                      \
    }                                                                                                       \
  } while (0)

// ncclGroupStart ... ncclGroupEnd with the group closed on every path: a throw between the two calls must not leave the
// NCCL group of this thread open (the next collective would silently queue into it).
struct NcclGroup {
  NcclApi& api;
  bool open;
  explicit NcclGroup(NcclApi& a) : api(a), open(false) {
    SGX_NCCL(api.GroupStart());
    open = true;
  }
  void end() {
    open = false;
    SGX_NCCL(api.GroupEnd());
  }
  ~NcclGroup() {
    if (open) api.GroupEnd();
  }
};

void comm_unique_id(uint8_t id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  SGX_NCCL(nccl_api().GetUniqueId(&u));
  memcpy(id, &u, 128);
}

static void comm_init_staging(sgx_comm* cm) {
  cm->ag_dev = nullptr;
  cm->ag_host = nullptr;
  cm->ag_cap = 0;
  SGX_CUDA(cudaEventCreateWithFlags(&cm->ag_done, cudaEventDisableTiming));
  cm->side = nullptr;
  cm->side_go = nullptr;
  if (!cm->local) {
    SGX_CUDA(cudaStreamCreateWithFlags(&cm->side, cudaStreamNonBlocking));
    SGX_CUDA(cudaEventCreateWithFlags(&cm->side_go, cudaEventDisableTiming));
  }
  cm->peer_state = 0;
  cm->peer_bytes = 0;
  cm->epoch = 0;
  for (int i = 0; i < kMaxShards; ++i) cm->peer_mine[i] = cm->peer_addr[i] = nullptr;
}

static void peer_release(sgx_comm* cm) {
  for (int g = 0; g < cm->world; ++g) {
    const bool mine = cm->local || g == cm->rank;
    if (!mine && cm->peer_addr[g]) cudaIpcCloseMemHandle(cm->peer_addr[g]);
    cm->peer_addr[g] = nullptr;
  }
  for (int li = 0; li < cm->n_local(); ++li) {
    if (cm->peer_mine[li]) cudaFree(cm->peer_mine[li]);
    cm->peer_mine[li] = nullptr;
  }
  cm->peer_bytes = 0;
}

sgx_comm* comm_create_nccl(sgx_ctx* ctx, const uint8_t id[128], int rank, int world) {
  SGX_REQUIRE(world >= 1 && world <= kMaxShards && rank >= 0 && rank < world, SGX_EINVAL, "sgx_comm_init: bad rank/world");
  ncclUniqueId u;
  memcpy(&u, id, 128);
  ncclComm_t c = nullptr;
  SGX_NCCL(nccl_api().CommInitRank(&c, world, u, rank));
  sgx_comm* cm = new sgx_comm();
  cm->ctx = ctx;
  cm->world = world;
  cm->rank = rank;
  cm->local = false;
  cm->nccl = c;
  comm_init_staging(cm);
  return cm;
}

sgx_comm* comm_create_local(sgx_ctx* ctx, int shards) {
  SGX_REQUIRE(shards >= 1 && shards <= kMaxShards, SGX_EINVAL, "sgx_comm_init_local: 1..16 shards");
  sgx_comm* cm = new sgx_comm();
  cm->ctx = ctx;
  cm->world = shards;
  cm->rank = 0;
  cm->local = true;
  cm->nccl = nullptr;
  comm_init_staging(cm);
  return cm;
}

void comm_destroy(sgx_comm* cm) {
  if (!cm) return;
  cudaStreamSynchronize(cm->ctx->stream);
  if (cm->side) cudaStreamSynchronize(cm->side);
  // peers may still be storing into this shard's region or reading their mapping of it: leave together
  if (cm->nccl && cm->peer_bytes) {
    try {
      comm_barrier(cm);
    } catch (const Error&) {
    }
  }
  peer_release(cm);
  if (cm->nccl) nccl_api().CommDestroy((ncclComm_t)cm->nccl);
  if (cm->side) cudaStreamDestroy(cm->side);
  if (cm->side_go) cudaEventDestroy(cm->side_go);
  if (cm->ag_dev) cudaFree(cm->ag_dev);
  if (cm->ag_host) cudaFreeHost(cm->ag_host);
  cudaEventDestroy(cm->ag_done);
  delete cm;
}

// ---- primitives --------------------------------------------------------------------------------------------------------

static void ag_reserve(sgx_comm* cm, uint64_t elems) {
  if (cm->ag_cap >= elems) return;
  SGX_CUDA(cudaStreamSynchronize(cm->ctx->stream));
  if (cm->ag_dev) cudaFree(cm->ag_dev);
  if (cm->ag_host) cudaFreeHost(cm->ag_host);
  cm->ag_dev = nullptr;
  cm->ag_host = nullptr;
  cm->ag_cap = 0;
  const uint64_t cap = std::max<uint64_t>(elems, 1024);
  SGX_CUDA(cudaMalloc((void**)&cm->ag_dev, cap * sizeof(uint64_t)));
  SGX_CUDA(cudaMallocHost((void**)&cm->ag_host, cap * sizeof(uint64_t)));
  cm->ag_cap = cap;
}

void comm_allgather_u64_begin(sgx_comm* cm, const uint64_t* const* dev_vals, uint64_t n) {
  cudaStream_t st = cm->ctx->stream;
  const uint64_t total = (uint64_t)cm->world * n;
  ag_reserve(cm, total);
  if (cm->local) {
    for (int s = 0; s < cm->world; ++s)
      SGX_CUDA(cudaMemcpyAsync(cm->ag_host + (uint64_t)s * n, dev_vals[s], n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SGX_CUDA(cudaEventRecord(cm->ag_done, st));
  } else {
    // on the side stream, behind everything enqueued on the main stream so far: what the caller enqueues next on the
    // main stream (the scatter and the push of a column pass) runs beside the collective instead of behind it
    SGX_CUDA(cudaEventRecord(cm->side_go, st));
    SGX_CUDA(cudaStreamWaitEvent(cm->side, cm->side_go, 0));
    SGX_NCCL(nccl_api().AllGather(dev_vals[0], cm->ag_dev, n, ncclUint64, (ncclComm_t)cm->nccl, cm->side));
    SGX_CUDA(cudaMemcpyAsync(cm->ag_host, cm->ag_dev, total * sizeof(uint64_t), cudaMemcpyDeviceToHost, cm->side));
    SGX_CUDA(cudaEventRecord(cm->ag_done, cm->side));
  }
}

void comm_allgather_u64_end(sgx_comm* cm, uint64_t n, uint64_t* out) {
  SGX_CUDA(cudaEventSynchronize(cm->ag_done));
  memcpy(out, cm->ag_host, (uint64_t)cm->world * n * sizeof(uint64_t));
}

static __global__ void add_u64_kernel(uint64_t* __restrict__ acc, const uint64_t* __restrict__ x, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] += x[i];
}

void comm_allreduce_sum_u64(sgx_comm* cm, uint64_t* const* dev_vals, uint64_t n) {
  cudaStream_t st = cm->ctx->stream;
  if (cm->local) {
    for (int s = 1; s < cm->world; ++s) SGX_LAUNCH(cm->ctx, add_u64_kernel, grid_for(n, 256), 256, 0, dev_vals[0], (const uint64_t*)dev_vals[s], n);
    for (int s = 1; s < cm->world; ++s)
      SGX_CUDA(cudaMemcpyAsync(dev_vals[s], dev_vals[0], n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return;
  }
  SGX_NCCL(nccl_api().AllReduce(dev_vals[0], dev_vals[0], n, ncclUint64, ncclSum, (ncclComm_t)cm->nccl, st));
}

void comm_alltoallv(sgx_comm* cm, const void* const* send, void* const* recv, size_t elem_bytes, const uint64_t* cnt,
                    const uint64_t* send_off, const uint64_t* recv_off) {
  cudaStream_t st = cm->ctx->stream;
  const int W = cm->world;
  if (cm->local) {
    for (int s = 0; s < W; ++s)
      for (int d = 0; d < W; ++d) {
        const uint64_t c = cnt[s * W + d];
        if (!c) continue;
        SGX_CUDA(cudaMemcpyAsync((uint8_t*)recv[d] + recv_off[d * W + s] * elem_bytes,
                                 (const uint8_t*)send[s] + send_off[s * W + d] * elem_bytes, c * elem_bytes,
                                 cudaMemcpyDeviceToDevice, st));
      }
    return;
  }
  const int me = cm->rank;
  NcclApi& api = nccl_api();
  // own block: plain copy; the rest as one grouped send/recv set
  if (cnt[me * W + me])
    SGX_CUDA(cudaMemcpyAsync((uint8_t*)recv[0] + recv_off[me * W + me] * elem_bytes,
                             (const uint8_t*)send[0] + send_off[me * W + me] * elem_bytes, cnt[me * W + me] * elem_bytes,
                             cudaMemcpyDeviceToDevice, st));
  NcclGroup grp(api);
  for (int p = 0; p < W; ++p) {
    if (p == me) continue;
    const uint64_t cs = cnt[me * W + p], cr = cnt[p * W + me];
    if (cs)
      SGX_NCCL(api.Send((const uint8_t*)send[0] + send_off[me * W + p] * elem_bytes, cs * elem_bytes, ncclUint8, p,
                        (ncclComm_t)cm->nccl, st));
    if (cr)
      SGX_NCCL(api.Recv((uint8_t*)recv[0] + recv_off[me * W + p] * elem_bytes, cr * elem_bytes, ncclUint8, p,
                        (ncclComm_t)cm->nccl, st));
  }
  grp.end();
}

void comm_gatherv_bytes(sgx_comm* cm, const uint8_t* const* send, const uint64_t* bytes, int root, uint8_t* root_recv) {
  cudaStream_t st = cm->ctx->stream;
  const int W = cm->world;
  std::vector<uint64_t> off(W + 1, 0);
  for (int s = 0; s < W; ++s) off[s + 1] = off[s] + bytes[s];
  if (cm->local) {
    for (int s = 0; s < W; ++s)
      if (bytes[s]) SGX_CUDA(cudaMemcpyAsync(root_recv + off[s], send[s], bytes[s], cudaMemcpyDeviceToDevice, st));
    return;
  }
  NcclApi& api = nccl_api();
  const int me = cm->rank;
  if (me == root) {
    if (bytes[me]) SGX_CUDA(cudaMemcpyAsync(root_recv + off[me], send[0], bytes[me], cudaMemcpyDeviceToDevice, st));
    NcclGroup grp(api);
    for (int s = 0; s < W; ++s)
      if (s != me && bytes[s]) SGX_NCCL(api.Recv(root_recv + off[s], bytes[s], ncclUint8, s, (ncclComm_t)cm->nccl, st));
    grp.end();
  } else if (bytes[me]) {
    SGX_NCCL(api.Send(send[0], bytes[me], ncclUint8, root, (ncclComm_t)cm->nccl, st));
  }
}

void comm_allgatherv(sgx_comm* cm, const void* const* send, const uint64_t* counts, size_t elem_bytes, void* recv) {
  cudaStream_t st = cm->ctx->stream;
  const int W = cm->world;
  std::vector<uint64_t> off(W + 1, 0);
  for (int s = 0; s < W; ++s) off[s + 1] = off[s] + counts[s];
  uint8_t* dst = (uint8_t*)recv;
  if (cm->local) {
    for (int s = 0; s < W; ++s)
      if (counts[s])
        SGX_CUDA(cudaMemcpyAsync(dst + off[s] * elem_bytes, send[s], counts[s] * elem_bytes, cudaMemcpyDeviceToDevice, st));
    return;
  }
  NcclApi& api = nccl_api();
  const int me = cm->rank;
  if (counts[me])
    SGX_CUDA(cudaMemcpyAsync(dst + off[me] * elem_bytes, send[0], counts[me] * elem_bytes, cudaMemcpyDeviceToDevice, st));
  NcclGroup grp(api);
  for (int p = 0; p < W; ++p) {
    if (p == me) continue;
    if (counts[me]) SGX_NCCL(api.Send(send[0], counts[me] * elem_bytes, ncclUint8, p, (ncclComm_t)cm->nccl, st));
    if (counts[p]) SGX_NCCL(api.Recv(dst + off[p] * elem_bytes, counts[p] * elem_bytes, ncclUint8, p, (ncclComm_t)cm->nccl, st));
  }
  grp.end();
}

bool comm_peer_reserve(sgx_comm* cm, size_t bytes) {
  if (cm->peer_state < 0) return false;
  if (cm->peer_state == 0) {
    const char* x = getenv("SGX_SHARD_XCHG");
    if (x && !strcmp(x, "nccl")) {
      cm->peer_state = -1;
      return false;
    }
  }
  if (cm->peer_bytes >= bytes) return true;
  cudaStream_t st = cm->ctx->stream;
  SGX_CUDA(cudaStreamSynchronize(st));
  if (cm->peer_bytes) {
    comm_barrier(cm);  // nobody is still using the old regions
    peer_release(cm);
  }
  const size_t want = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  for (int li = 0; li < cm->n_local(); ++li) {
    SGX_CUDA(cudaMalloc((void**)&cm->peer_mine[li], want));
    SGX_CUDA(cudaMemsetAsync(cm->peer_mine[li], 0, kPeerHeaderBytes, st));
  }
  SGX_CUDA(cudaStreamSynchronize(st));
  if (cm->local) {
    for (int g = 0; g < cm->world; ++g) cm->peer_addr[g] = cm->peer_mine[g];
    cm->peer_bytes = want;
    cm->peer_state = 1;
    return true;
  }
  // exchange the IPC handles through the communicator and map every peer's region
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  const int W = cm->world, me = cm->rank;
  const uint64_t hw = sizeof(cudaIpcMemHandle_t) / sizeof(uint64_t);
  ag_reserve(cm, (uint64_t)(W + 1) * hw);
  cudaIpcMemHandle_t mine;
  uint64_t ok = cudaIpcGetMemHandle(&mine, cm->peer_mine[0]) == cudaSuccess ? 1 : 0;
  cudaGetLastError();
  std::vector<cudaIpcMemHandle_t> all(W);
  SGX_CUDA(cudaMemcpyAsync(cm->ag_dev + (uint64_t)W * hw, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  SGX_NCCL(nccl_api().AllGather(cm->ag_dev + (uint64_t)W * hw, cm->ag_dev, hw, ncclUint64, (ncclComm_t)cm->nccl, st));
  SGX_CUDA(cudaMemcpyAsync(all.data(), cm->ag_dev, (uint64_t)W * sizeof(mine), cudaMemcpyDeviceToHost, st));
  SGX_CUDA(cudaStreamSynchronize(st));
  cm->peer_addr[me] = cm->peer_mine[0];
  for (int g = 0; g < W && ok; ++g) {
    if (g == me) continue;
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    cm->peer_addr[g] = (uint8_t*)p;
  }
  // everybody or nobody: the processes must agree on the transport
  DevBuf<uint64_t> agree(1);
  SGX_CUDA(cudaMemcpyAsync(agree.p, &ok, sizeof(ok), cudaMemcpyHostToDevice, st));
  SGX_NCCL(nccl_api().AllReduce(agree.p, agree.p, 1, ncclUint64, ncclMin, (ncclComm_t)cm->nccl, st));
  SGX_CUDA(cudaMemcpyAsync(&ok, agree.p, sizeof(ok), cudaMemcpyDeviceToHost, st));
  SGX_CUDA(cudaStreamSynchronize(st));
  cm->peer_bytes = want;
  if (!ok) {
    peer_release(cm);
    cm->peer_state = -1;
    return false;
  }
  cm->peer_state = 1;
  return true;
}

void comm_barrier(sgx_comm* cm) {
  if (cm->local) {
    SGX_CUDA(cudaStreamSynchronize(cm->ctx->stream));
    return;
  }
  DevBuf<uint64_t> x(1);
  SGX_CUDA(cudaMemsetAsync(x.p, 0, sizeof(uint64_t), cm->ctx->stream));
  SGX_NCCL(nccl_api().AllReduce(x.p, x.p, 1, ncclUint64, ncclSum, (ncclComm_t)cm->nccl, cm->ctx->stream));
  SGX_CUDA(cudaStreamSynchronize(cm->ctx->stream));
}

}  // namespace sgx
