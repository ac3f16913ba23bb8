// k-mer error-correction driver (correct.cu).
#ifndef SGX_CORRECT_CUH_
#define SGX_CORRECT_CUH_

#include "fm.cuh"
#include "reads.cuh"

namespace sgx {
// KmerCorrector::process (src/correct_processor.cpp:72-232) for reads [first, first+count) against the forward index.
// qual / qual_off: host quality strings of exactly those reads (nullptr: FASTA input, phred 15 everywhere).
void correct_run(sgx_ctx* ctx, const sgx_fm* fm, const sgx_reads* reads, uint64_t first, uint64_t count, const char* qual,
                 const uint64_t* qual_off, const sgx_correct_params* prm, sgx_corrected* out);
}  // namespace sgx

#endif  // SGX_CORRECT_CUH_
