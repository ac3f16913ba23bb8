// Device FM-index object and its builders (from run bytes of a .bwt file, or straight from a device symbol array).
#ifndef SGX_FM_CUH_
#define SGX_FM_CUH_

#include "fm_layout.h"
#include "prims.cuh"

struct sgx_reads;

struct sgx_fm {
  sgx_ctx* ctx;
  sgx::FmBlockLane* blocks;  // device
  uint64_t n;
  uint64_t pred[6];
  // optional (sgx_fm_attach_reads): read-start order of this index, its inverse, and the indexed read set
  uint32_t* sai;   // device, n_strings
  uint32_t* inv;   // device, inv[sai[k]] = k
  uint64_t n_strings;
  const sgx_reads* indexed;
  // k-mer start table (fm_build_ktab): interval {first row, rows} of every string of ktab_k bases, indexed by the string
  // read backwards as a base-4 number (last base most significant) -- the state of a backward search after its first
  // ktab_k steps.  Only for indices with < 2^32 rows; nullptr / 0 otherwise.
  uint2* ktab;
  int ktab_k;
  sgx::FmView view() const {
    sgx::FmView v;
    v.blocks = blocks;
    v.n = n;
    for (int i = 0; i < 6; ++i) v.pred[i] = pred[i];
    return v;
  }
};

namespace sgx {

// Builds the rank blocks of `sym` (n device bytes, ranks 0..4).  Returns a new sgx_fm (caller owns).
sgx_fm* fm_build_from_symbols(sgx_ctx* ctx, const uint8_t* sym_dev, uint64_t n);
// Expands run bytes (device) into a fresh device symbol array of n_suffixes bytes.
void rle_expand(sgx_ctx* ctx, const uint8_t* rle_dev, uint64_t n_runs, uint64_t n_suffixes, uint8_t* sym_dev);

// Builds fm->ktab for strings of k bases (k >= 1; level by level, 4^k / 3 pairs of rank queries in all).
void fm_build_ktab(sgx_ctx* ctx, sgx_fm* fm, int k);
// Uploads the read-start order and builds its inverse.
void fm_attach(sgx_ctx* ctx, sgx_fm* fm, const uint32_t* sai, uint64_t n, const sgx_reads* indexed);
// Batched FMIndex::getOcc(i): host positions in, host counts (5 per position) out.
void fm_occ_batch(sgx_ctx* ctx, const sgx_fm* fm, const uint64_t* pos, uint64_t n, uint64_t* counts);
// Batched FMIndex::Interval::occurrences of host query strings.
void fm_count_batch(sgx_ctx* ctx, const sgx_fm* fm, const char* ascii, const uint64_t* off, uint64_t n, uint64_t* counts);

}  // namespace sgx

#endif  // SGX_FM_CUH_
