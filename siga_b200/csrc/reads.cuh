// Device read store: 2 bits per base, every read starting on a 32-bit word (sgx_common.h: ReadsView).
#ifndef SGX_READS_CUH_
#define SGX_READS_CUH_

#include "sgx_device.cuh"

struct sgx_reads {
  sgx_ctx* ctx;
  uint32_t* words;  // device
  uint64_t* woff;   // device, n + 1
  uint32_t* len;    // device, n
  uint64_t n;
  uint64_t bases;
  uint64_t n_words;
  uint32_t max_len;
  uint32_t min_len;
  sgx::ReadsView view() const {
    sgx::ReadsView v;
    v.words = words;
    v.woff = woff;
    v.len = len;
    v.n = n;
    return v;
  }
};

namespace sgx {
// Packs host reads (ascii + offsets) into a new device store.  Throws sgx::Error (SGX_EINVAL on bad input).
sgx_reads* reads_pack(sgx_ctx* ctx, const char* ascii, const uint64_t* off, uint64_t n);
void reads_destroy(sgx_reads* r);
}  // namespace sgx

#endif  // SGX_READS_CUH_
