// Shard communicator of the multi-GPU paths (include/siga_b200.h: sgx_comm).
//
// Two transports behind the same four primitives:
//   * NCCL  -- one process per GPU (torch.distributed / mpirun style launch); the unique id travels through the
//              caller's own rendezvous (sgx_comm_unique_id on rank 0 -> broadcast -> sgx_comm_init everywhere).
//              libnccl.so.2 is resolved with dlopen at first use, so the library itself has no link-time NCCL
//              dependency and picks up the NCCL already loaded by the host process (e.g. torch's).
//   * local -- `world` virtual shards inside ONE context (one device, one stream).  The sharded algorithms run
//              unchanged, with device-to-device copies where NCCL would send; this is how the bucket-sharded index
//              build is parity-tested on a single GPU.
// A process drives `n_local` shards: `world` of them in local mode, exactly one (its rank) in NCCL mode.
#ifndef SGX_COMM_CUH_
#define SGX_COMM_CUH_

#include "sgx_device.cuh"

namespace sgx {
static const int kMaxShards = 16;
static const size_t kPeerHeaderBytes = 4096;  // start of a peer region: one arrival flag (uint64) per sender
}

struct sgx_comm {
  sgx_ctx* ctx;
  int world;    // number of shards
  int rank;     // NCCL mode: this process's shard; local mode: 0
  bool local;   // true: all shards live in this process
  void* nccl;   // ncclComm_t
  // small persistent staging of comm_allgather_u64 (device gather buffer, pinned host copy, completion event)
  uint64_t* ag_dev;
  uint64_t* ag_host;
  uint64_t ag_cap;  // uint64 elements
  cudaEvent_t ag_done;
  // NCCL mode: the per-pass count all-gather runs on a stream of its own, beside the scatter / push of the main stream
  cudaStream_t side;
  cudaEvent_t side_go;
  // Peer-memory exchange (comm_peer_reserve): one region per shard, every region mapped into every process of the
  // communicator (cudaIpc between processes; plain pointers between the virtual shards of local mode).  Senders store
  // straight into the receiver's region over NVLink and raise the receiver's arrival flag; nothing goes through NCCL.
  int peer_state;                          // 0 not tried yet, 1 mapped, -1 unavailable (the NCCL send/recv path is used)
  uint8_t* peer_mine[sgx::kMaxShards];     // regions of the local shards (by local index)
  uint8_t* peer_addr[sgx::kMaxShards];     // region of shard g as this process addresses it
  size_t peer_bytes;
  uint64_t epoch;                          // exchanges done on this communicator; the flag value of the next one
  int n_local() const { return local ? world : 1; }
  int shard_of(int li) const { return local ? li : rank; }  // global shard index of local shard li
};

namespace sgx {

void comm_unique_id(uint8_t id[128]);
sgx_comm* comm_create_nccl(sgx_ctx* ctx, const uint8_t id[128], int rank, int world);
sgx_comm* comm_create_local(sgx_ctx* ctx, int shards);
void comm_destroy(sgx_comm* cm);

// out[(s * n) + i] = value i of shard s, for all `world` shards, on the HOST of every process.  `dev_vals[li]` is a
// device array of n uint64 values of local shard li.  Waits for the gathered values only (an event), not for work
// enqueued on the stream afterwards by comm_allgather_u64_begin's caller:
//   comm_allgather_u64_begin(...); <enqueue more kernels>; comm_allgather_u64_end(cm, n, out);
void comm_allgather_u64_begin(sgx_comm* cm, const uint64_t* const* dev_vals, uint64_t n);
void comm_allgather_u64_end(sgx_comm* cm, uint64_t n, uint64_t* out);
inline void comm_allgather_u64(sgx_comm* cm, const uint64_t* const* dev_vals, uint64_t n, uint64_t* out) {
  comm_allgather_u64_begin(cm, dev_vals, n);
  comm_allgather_u64_end(cm, n, out);
}

// Element-wise sum over shards of an n-element uint64 device array (in place, every local shard gets the sum).
void comm_allreduce_sum_u64(sgx_comm* cm, uint64_t* const* dev_vals, uint64_t n);

// Personalised all-to-all of `elem_bytes`-sized items: local shard li sends items [send_off[s][d], send_off[s][d] +
// cnt[s][d]) of send[li] to shard d, which stores them at item recv_off[d][s] of its recv buffer (s = shard_of(li)).
// cnt / send_off / recv_off are world x world host matrices (row-major) known to every process.
void comm_alltoallv(sgx_comm* cm, const void* const* send, void* const* recv, size_t elem_bytes, const uint64_t* cnt,
                    const uint64_t* send_off, const uint64_t* recv_off);

// Concatenation of byte segments on shard `root`: shard s's `bytes[s]` bytes land at offset sum(bytes[0..s)) of
// `root_recv` (a device buffer of the local shard that is `root`; ignored elsewhere).
void comm_gatherv_bytes(sgx_comm* cm, const uint8_t* const* send, const uint64_t* bytes, int root, uint8_t* root_recv);

// Concatenation, on EVERY process, of one block per shard: shard s contributes counts[s] elements of elem_bytes from
// send[local index of s]; `recv` (one buffer per process) receives them in shard order.
void comm_allgatherv(sgx_comm* cm, const void* const* send, const uint64_t* counts, size_t elem_bytes, void* recv);

void comm_barrier(sgx_comm* cm);

// Collective.  Makes sure every shard owns a peer region of at least `bytes` (kPeerHeaderBytes of flags + payload) that
// all processes of the communicator can store into; returns false when peer mapping is not possible here (cudaIpc
// refused, SGX_SHARD_XCHG=nccl): the caller then exchanges through comm_alltoallv.  Regions live until comm_destroy.
bool comm_peer_reserve(sgx_comm* cm, size_t bytes);
// The next exchange's flag value (all processes call this in lock step, so they agree on it).
inline uint64_t comm_next_epoch(sgx_comm* cm) { return ++cm->epoch; }
// Arrival flag of sender `s` in the region of shard `d`.
inline uint64_t* comm_peer_flag(sgx_comm* cm, int d, int s) { return (uint64_t*)cm->peer_addr[d] + s; }

}  // namespace sgx

#endif  // SGX_COMM_CUH_
