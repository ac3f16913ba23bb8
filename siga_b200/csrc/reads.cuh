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
  // sgx_reads_pack_async: the bases arrive in chunks on the context's second stream while the caller goes on; chunk c
  // holds reads [c * chunk_reads, (c + 1) * chunk_reads) and chunk_ev[c] fires when its words are in place.  `pending`
  // until a consumer has waited for all chunks and looked at the validity flags (reads_settle).
  mutable bool pending = false;
  mutable bool invalid = false;
  mutable std::vector<cudaEvent_t> chunk_ev;
  uint64_t chunk_reads = 0;
  // The packed words come from the context's work-area cache when they are large (a stream-ordered pool allocation of
  // ~100 MB was seen to stall the host for 50 - 230 ms every few calls, profiles/r2_summary.md); `words` points into it.
  sgx::DevBuf<uint32_t> words_buf;
  mutable sgx::DevBuf<char> staging;        // uploaded ASCII, released when settled
  mutable sgx::DevBuf<uint64_t> staging_off;
  mutable sgx::DevBuf<uint32_t> flags;      // [0] min len, [1] max len, [2] bad bits, [3] first read with a bad base
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
// Same, returning as soon as the lengths are known: the bases are uploaded and packed chunk by chunk on the context's
// second stream.  `ascii` must stay valid until a consumer has settled the store.
sgx_reads* reads_pack_async(sgx_ctx* ctx, const char* ascii, const uint64_t* off, uint64_t n);
// Makes `stream` wait for the chunks that hold reads [lo, hi) (no-op for a settled store).
void reads_wait_range(cudaStream_t stream, const sgx_reads* r, uint64_t lo, uint64_t hi);
// Waits for every chunk, releases the upload staging and throws SGX_EINVAL if a base other than A/C/G/T was seen.
void reads_settle(const sgx_reads* r);
void reads_destroy(sgx_reads* r);
}  // namespace sgx

#endif  // SGX_READS_CUH_
