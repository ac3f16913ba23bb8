// Overlap-search driver (overlap.cu).
#ifndef SGX_OVERLAP_CUH_
#define SGX_OVERLAP_CUH_

#include "fm.cuh"
#include "reads.cuh"

namespace sgx {
// Fills `hits` (pinned host buffers from the context pool) for reads [first, first+count).
void overlap_run(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                 uint64_t count, uint32_t min_ov, int irreducible, int rc, sgx_hits* hits);
}  // namespace sgx

#endif  // SGX_OVERLAP_CUH_
