// Overlap-search driver (overlap.cu).
#ifndef SGX_OVERLAP_CUH_
#define SGX_OVERLAP_CUH_

#include "fm.cuh"
#include "reads.cuh"

namespace sgx {
// Fills `hits` (pinned host buffers from the context pool) for reads [first, first+count).
void overlap_run(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                 uint64_t count, uint32_t min_ov, int irreducible, int rc, sgx_hits* hits);
// Same search, but the hits -> edge conversion of Hit2OverlapConverter (src/overlap_builder.cpp:345-375) also runs on
// the device and only compact edge records come back.
void overlap_edges_run(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                       uint64_t count, uint32_t min_ov, int irreducible, int rc, const uint32_t* name_rank,
                       sgx_edges* out);
}  // namespace sgx

#endif  // SGX_OVERLAP_CUH_
