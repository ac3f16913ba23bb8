// Shared host/device definitions of the siga_b200 hot path.
//
// Everything in this header compiles both under nvcc (device code of the product library) and under plain g++
// (tests/hostemu, a test-only single-lane instantiation of the same control logic used to debug it without a GPU).
#ifndef SGX_COMMON_H_
#define SGX_COMMON_H_

#include <stdint.h>

#include "../../include/siga_b200.h"

#ifdef __CUDACC__
#define SGX_HD __host__ __device__ __forceinline__
#define SGX_D __device__ __forceinline__
#else
#define SGX_HD inline
#define SGX_D inline
#endif

namespace sgx {

typedef uint64_t pos_t;
static const pos_t kNoPos = (pos_t)-1;

// Symbol ranks, src/alphabet.h:13-39.
enum { kDollar = 0, kA = 1, kC = 2, kG = 3, kT = 4, kSigma = 5 };

// Complement in rank space: A<->T, C<->G; '$' (and anything unknown) stays 0, which is what the reference's
// make_dna_complement('$') -> '\0' -> torank 0 amounts to (src/kseq.cpp:18-42, src/alphabet.h:19-39).
SGX_HD int comp_rank(int r) { return r == 0 ? 0 : 5 - r; }

// FMIndex::Interval (src/fmindex.h:63-114).
struct Ival {
  pos_t lo, hi;
};
SGX_HD bool ival_valid(pos_t lo, pos_t hi) { return hi != kNoPos && hi >= lo; }

// 2-bit packed read store: read i occupies 32-bit words [woff[i], woff[i+1]) ; base j sits in word j>>4 at
// bits 2*(j&15) with code A=0 C=1 G=2 T=3 (rank = code + 1).
struct ReadsView {
  const uint32_t* words;
  const uint64_t* woff;
  const uint32_t* len;
  uint64_t n;
};

SGX_HD int read_rank(const uint32_t* w, uint32_t j) { return (int)((w[j >> 4] >> (2 * (j & 15))) & 3u) + 1; }

}  // namespace sgx

#endif  // SGX_COMMON_H_
