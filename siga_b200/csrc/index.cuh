// Index construction driver (index.cu).
#ifndef SGX_INDEX_CUH_
#define SGX_INDEX_CUH_

#include "comm.cuh"
#include "reads.cuh"

namespace sgx {
// Fills `out` (pinned host buffers from the context pool) with the run bytes and the read-start order.
void index_build(sgx_ctx* ctx, const sgx_reads* reads, int order, bool reversed, sgx_index_out* out);
// Same result with the BWT rows sharded over a communicator (index_sharded.cu).
void index_build_sharded(sgx_ctx* ctx, sgx_comm* cm, const sgx_reads* reads, int order, bool reversed, bool segments,
                         sgx_index_out* out);
}  // namespace sgx

#endif  // SGX_INDEX_CUH_
