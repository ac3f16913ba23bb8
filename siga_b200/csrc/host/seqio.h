// Host-side sequence input and SGA-compatible output writers of the `siga` front-end.
//
// Mirrors, re-implemented on zlib, the behaviour of the reference's readers/writers on the index/overlap path:
//   FASTA/FASTQ readers        src/kseq.cpp:127-228  (format chosen by the first byte; multi-line FASTA joined; name =
//                              text up to the first blank/tab, the rest is the comment, src/kseq.cpp:71-79)
//   Utils::stem                src/utils.cpp:128-135 (strip .gz/.bz2, directory and one extension)
//   .bwt / .sai writers        src/bwt.cpp:121-178, src/suffix_array.cpp:16-43
//   .bwt / .sai readers        src/bwt.cpp:59-98, src/suffix_array.cpp:56-95
#ifndef SGX_HOST_SEQIO_H_
#define SGX_HOST_SEQIO_H_

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

namespace sgxhost {

// Byte buffer of the read set: like std::string, except that resize_uninit() does not touch the new bytes, so that the
// threads of the parallel reader fault in and fill their own parts (zero-filling 460 MB on one thread cost more than
// parsing it on eight).
class ByteBuf {
 public:
  ByteBuf() : p_(nullptr), n_(0), cap_(0) {}
  ~ByteBuf() { free(p_); }
  ByteBuf(const ByteBuf&) = delete;
  ByteBuf& operator=(const ByteBuf&) = delete;
  const char* data() const { return p_ ? p_ : ""; }
  char* data() { return p_; }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  char& operator[](size_t i) { return p_[i]; }
  const char& operator[](size_t i) const { return p_[i]; }
  void clear() { n_ = 0; }
  void release() {
    free(p_);
    p_ = nullptr;
    n_ = cap_ = 0;
  }
  bool reserve(size_t cap) {
    if (cap <= cap_) return true;
    char* q = (char*)realloc(p_, cap);
    if (!q) return false;
    p_ = q;
    cap_ = cap;
    return true;
  }
  bool resize_uninit(size_t n) {
    if (!reserve(n)) return false;
    n_ = n;
    return true;
  }
  void append(const char* s, size_t len) {
    if (n_ + len > cap_ && !reserve(std::max(n_ + len, cap_ + cap_ / 2 + 4096))) throw std::bad_alloc();
    memcpy(p_ + n_, s, len);
    n_ += len;
  }
  bool operator==(const ByteBuf& o) const { return n_ == o.n_ && (n_ == 0 || memcmp(p_, o.p_, n_) == 0); }

 private:
  char* p_;
  size_t n_, cap_;
};

struct ReadSet {
  std::vector<std::string> names;
  std::vector<std::string> comments;
  ByteBuf bases;                // all reads back to back
  std::vector<uint64_t> off;    // n + 1 offsets into bases
  bool keep_quals = false;      // set before reading: keep the FASTQ quality strings (`correct` needs them)
  ByteBuf quals;                // with keep_quals and FASTQ input: qualities, same offsets as bases; else empty
  size_t size() const { return names.size(); }
};

// Returns false (with a message in *err) when the file cannot be read or is not FASTA/FASTQ.
// Plain files of 16 MB and more are parsed on all cores (same result as the serial reader, which handles .gz input,
// small files and anything irregular).
bool read_sequences(const std::string& path, ReadSet* out, std::string* err);
bool read_sequences_serial(const std::string& path, ReadSet* out, std::string* err);

std::string stem(const std::string& path);

bool write_bwt(const std::string& path, uint64_t n_strings, uint64_t n_suffixes, const uint8_t* runs, uint64_t n_runs);
// One .bwt file written by several processes together (sgx_index_build_sharded_segments): each writes its `seg_bytes` run
// bytes at their place behind the 30-byte header; the one with `header` set also writes the header and gives the file its
// final length (so a longer stale file of the same name does not survive).  No truncation on open: the writers race.
bool write_bwt_segment(const std::string& path, uint64_t n_strings, uint64_t n_suffixes, uint64_t n_runs, const uint8_t* seg,
                       uint64_t seg_offset, uint64_t seg_bytes, bool header);
bool write_sai(const std::string& path, const uint32_t* ids, uint64_t n);
bool read_bwt(const std::string& path, uint64_t* n_strings, uint64_t* n_suffixes, std::vector<uint8_t>* runs);
bool read_sai(const std::string& path, std::vector<uint32_t>* ids);

// Rank of every name in std::string order (equal names share a rank).
std::vector<uint32_t> name_ranks(const std::vector<std::string>& names);

}  // namespace sgxhost

#endif  // SGX_HOST_SEQIO_H_
