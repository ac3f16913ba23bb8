// Index construction driver (index.cu).
#ifndef SGX_INDEX_CUH_
#define SGX_INDEX_CUH_

#include "reads.cuh"

namespace sgx {
// Fills `out` (pinned host buffers from the context pool) with the run bytes and the read-start order.
void index_build(sgx_ctx* ctx, const sgx_reads* reads, int order, bool reversed, sgx_index_out* out);
}  // namespace sgx

#endif  // SGX_INDEX_CUH_
