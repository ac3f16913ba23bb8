// Post stage, common case: irreducible-overlap extraction from the TARGET READS instead of BWT walks.
//
// What the reference does (IrreducibleBlockListExtractor::extract, src/overlap_builder.cpp:711-809) is extend every
// candidate block base by base through the FM-index until the longest block of a group reaches the end of its read;
// at 100x coverage / 1 % errors that is ~3400 rank queries per read, 3.7x the cost of the search itself.  The
// reference also loads .sai/.rsai (src/overlap_builder.cpp:466), which map a row of a candidate's capped[0] interval
// to the target read.  With the read store on the device the extension string of a candidate row is simply the rest
// of that read, so the whole group process collapses to string comparisons:
//
//   rows      every row r of every candidate block on one side (right: suffix-prefix + suffix-suffix lists,
//             left: prefix-prefix + prefix-suffix lists); target t = sa[r]; overlap k = block length;
//             ext(r) = X[k..len_t) in query orientation, X = t (forward-indexed targets) or reverse(t)
//             (TARGETREV), complemented under QUERYCOMP.
//   rule      when all targets of the side have the same length (top-level blocks then end first, no abort, no
//             '$' branching): row a is transitive  <=>  some row b with k_b > k_a has ext(b) as a prefix of ext(a).
//             (Group semantics: a group is dropped when its longest block ends; b ends while a is still in b's
//             group exactly when ext(b) is a prefix of ext(a).)
//   output    rows of one block with identical ext are identical reads and contiguous; each surviving class is one
//             output block: capped[0] = its row range, capped[1] = the rank range of those reads in the opposite
//             index (inverse .sai), raw/length/af copied -- bit-identical to what the '$'-terminated walk yields.
//
// Anything outside these preconditions (interval intersections inside a list = periodic sequence, targets of
// different lengths, more rows than the shared-memory table holds) is NOT approximated: the read is put on the
// slow list and redone by ov_post_kernel, the literal BWT-walking implementation.
#ifndef SGX_OVERLAP_REDUCE_CUH_
#define SGX_OVERLAP_REDUCE_CUH_

#include "overlap_logic.h"

namespace sgx {

struct SaView {
  const uint32_t* sa;   // row of the '$' block -> read id
  const uint32_t* inv;  // read id -> row
};

static const int kRdWarps = 4;  // warps (reads in flight) per CTA

// Bytes of shared memory one warp needs for `rows_cap` rows of `extw` 64-bit extension words.
__host__ __device__ inline size_t reduce_smem_per_warp(uint32_t rows_cap, uint32_t extw) {
  return (size_t)rows_cap * extw * 8 + (size_t)rows_cap * (4 + 4 + 2 + 2 + 2 + 2);
}

__device__ __forceinline__ bool rd_intersect(uint64_t s1, uint64_t e1, uint64_t s2, uint64_t e2) {
  return !(s1 > e2 || s2 > e1);
}

__global__ void __launch_bounds__(kRdWarps * 32)
ov_reduce_kernel(ReadsView treads, SaView sf, SaView sr, uint64_t count, int irreducible, OvBatch bt, uint32_t rows_cap,
                 uint32_t extw, uint32_t* __restrict__ slow_list, uint32_t* __restrict__ n_slow) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  unsigned char* base = smem_raw + (size_t)wib * reduce_smem_per_warp(rows_cap, extw);
  uint64_t* ext = reinterpret_cast<uint64_t*>(base);                       // [rows_cap][extw]
  uint32_t* r_row = reinterpret_cast<uint32_t*>(ext + (size_t)rows_cap * extw);  // row in the '$' block
  uint32_t* r_tgt = r_row + rows_cap;                                      // target read id
  uint16_t* r_ent = reinterpret_cast<uint16_t*>(r_tgt + rows_cap);         // entry index within the side
  uint16_t* r_k = r_ent + rows_cap;                                        // overlap length
  uint16_t* r_n = r_k + rows_cap;                                          // extension length (bases)
  uint16_t* r_flag = r_n + rows_cap;                                       // bit0 transitive, bit1 class head

  const uint64_t warp0 = (uint64_t)blockIdx.x * kRdWarps + wib;
  const uint64_t nwarps = (uint64_t)gridDim.x * kRdWarps;
  for (uint64_t r = warp0; r < count; r += nwarps) {
    uint32_t nc[4], nk[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      nc[f] = bt.ncand[r * 4 + f];
      nk[f] = bt.ncontain[r * 4 + f];
    }
    const sgx_block* cand = bt.cand + (size_t)r * 4 * bt.find_cap;
    const sgx_block* contain = bt.contain + (size_t)r * 4;
    sgx_block* out = bt.out + (size_t)r * bt.out_cap;
    bool slow = false;

    // ---- SubMaximalBlockFilter precondition: no two capped[0] intervals of a list intersect (then the filter only
    //      reorders, src/overlap_builder.cpp:929-953).  Lists carry the containment blocks of their index side.
#pragma unroll 1
    for (int f = 0; f < 4; ++f) {
      const int q0 = f < 2 ? 0 : 2;
      const uint32_t n = nc[f] + nk[q0] + nk[q0 + 1];
      const sgx_block* lst = cand + (size_t)f * bt.find_cap;
      for (uint32_t i = lane; i < n; i += 32) {
        const sgx_block* bi = i < nc[f] ? lst + i : contain + (i - nc[f] == 0 && nk[q0] ? q0 : q0 + 1);
        const uint64_t lo = bi->capped[0], hi = bi->capped[1];
        for (uint32_t j = i + 1; j < n; ++j) {
          const sgx_block* bj = j < nc[f] ? lst + j : contain + (j - nc[f] == 0 && nk[q0] ? q0 : q0 + 1);
          if (rd_intersect(lo, hi, bj->capped[0], bj->capped[1])) slow = true;
        }
      }
    }
    slow = __any_sync(0xffffffffu, slow);

    uint32_t nout = 0;
    if (!slow) {
      // containment blocks go straight out: containfwd then containrev (src/overlap_builder.cpp:1160-1162)
      for (int q = 0; q < 4; ++q) {
        if (nk[q]) {
          if (lane == 0) out[nout] = contain[q];
          ++nout;
        }
      }
    }
    if (!slow && !irreducible) {
      // exhaustive mode: suffixfwd, suffixrev, prefixfwd, prefixrev (src/overlap_builder.cpp:1175-1178)
      const int order[4] = {0, 3, 1, 2};
      for (int q = 0; q < 4; ++q) {
        const sgx_block* lst = cand + (size_t)order[q] * bt.find_cap;
        const uint32_t n = nc[order[q]];
        if (nout + n > bt.out_cap) {
          slow = true;
          break;
        }
        for (uint32_t i = lane; i < n; i += 32) out[nout + i] = lst[i];
        nout += n;
      }
    }
    if (!slow && irreducible) {
#pragma unroll 1
      for (int side = 0; side < 2 && !slow; ++side) {
        const int fa = side == 0 ? 0 : 1, fb = side == 0 ? 3 : 2;
        const sgx_block* la = cand + (size_t)fa * bt.find_cap;
        const sgx_block* lb = cand + (size_t)fb * bt.find_cap;
        const uint32_t na = nc[fa], ne = nc[fa] + nc[fb];
        // ---- rows: one per (entry, row of its capped[0] interval)
        uint32_t R = 0;
        for (uint32_t e0 = 0; e0 < ne && !slow; e0 += 32) {
          const uint32_t e = e0 + lane;
          uint64_t lo = 0, w = 0;
          if (e < ne) {
            const sgx_block* b = e < na ? la + e : lb + (e - na);
            lo = b->capped[0];
            w = b->capped[1] - lo + 1;
            if (w > rows_cap) w = rows_cap + 1;
          }
          uint32_t incl = (uint32_t)w;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
          }
          const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
          if (R + tot > rows_cap) {
            slow = true;
            break;
          }
          const uint32_t at = R + incl - (uint32_t)w;
          for (uint32_t x = 0; x < (uint32_t)w; ++x) {
            r_row[at + x] = (uint32_t)(lo + x);
            r_ent[at + x] = (uint16_t)e;
          }
          R += tot;
        }
        if (slow) break;
        __syncwarp();
        // ---- per row: target read, extension string in query orientation
        uint32_t len0 = 0;
        bool mixed = false;
        for (uint32_t s = lane; s < R; s += 32) {
          const uint32_t e = r_ent[s];
          const sgx_block* b = e < na ? la + e : lb + (e - na);
          const uint32_t k = b->length, af = b->af;
          const bool tr = (af & SGX_AF_TARGETREV) != 0;
          const uint32_t t = (tr ? sr.sa : sf.sa)[r_row[s]];
          const uint32_t lt = treads.len[t];
          const uint32_t n = lt - k;
          if (n > extw * 32) mixed = true;  // cannot happen when extw covers max_len - min_overlap
          r_tgt[s] = t;
          r_k[s] = (uint16_t)k;
          r_n[s] = (uint16_t)n;
          r_flag[s] = (uint16_t)lt;  // borrowed until the length check below
          const uint32_t* w = treads.words + treads.woff[t];
          const uint32_t cmask = (af & SGX_AF_QUERYCOMP) ? 3u : 0u;
          uint64_t acc = 0;
          uint64_t* dst = ext + (size_t)s * extw;
          for (uint32_t p = 0; p < n && !mixed; ++p) {
            const uint32_t j = tr ? lt - 1 - k - p : k + p;
            const uint64_t code = ((w[j >> 4] >> (2 * (j & 15))) & 3u) ^ cmask;
            acc |= code << (2 * (p & 31));
            if ((p & 31) == 31) {
              dst[p >> 5] = acc;
              acc = 0;
            }
          }
          if (!mixed) {
            if (n & 31) dst[n >> 5] = acc;
            for (uint32_t q = (n + 31) >> 5; q < extw; ++q) dst[q] = 0;
          }
        }
        __syncwarp();
        if (R) len0 = r_flag[0];
        for (uint32_t s = lane; s < R; s += 32)
          if (r_flag[s] != len0) mixed = true;
        if (__any_sync(0xffffffffu, mixed)) {
          slow = true;  // targets of different lengths: group semantics are not the simple rule -> literal path
          break;
        }
        __syncwarp();
        // ---- transitive rows and class heads
        for (uint32_t s = lane; s < R; s += 32) {
          const uint32_t ka = r_k[s];
          const uint64_t* ea = ext + (size_t)s * extw;
          bool red = false;
          for (uint32_t b2 = 0; b2 < R && !red; ++b2) {
            if (r_k[b2] <= ka) continue;
            const uint32_t nb = r_n[b2];
            const uint64_t* eb = ext + (size_t)b2 * extw;
            bool eq = true;
            const uint32_t full = nb >> 5;
            for (uint32_t q = 0; q < full && eq; ++q) eq = ea[q] == eb[q];
            if (eq && (nb & 31)) {
              const uint64_t m = (1ull << (2 * (nb & 31))) - 1ull;
              eq = ((ea[full] ^ eb[full]) & m) == 0;
            }
            red = eq;
          }
          bool head = true;
          if (s > 0 && r_ent[s - 1] == r_ent[s]) {
            const uint64_t* ep = ext + (size_t)(s - 1) * extw;
            bool same = true;
            for (uint32_t q = 0; q < extw && same; ++q) same = ea[q] == ep[q];
            head = !same;
          }
          r_flag[s] = (uint16_t)((red ? 1u : 0u) | (head ? 2u : 0u));
        }
        __syncwarp();
        // ---- one output block per surviving class
        for (uint32_t s0 = 0; s0 < R; s0 += 32) {
          const uint32_t s = s0 + lane;
          const bool emit = s < R && (r_flag[s] & 3u) == 2u;
          const unsigned em = __ballot_sync(0xffffffffu, emit);
          if (emit) {
            const uint32_t at = nout + __popc(em & ((1u << lane) - 1u));
            if (at < bt.out_cap) {
              const uint32_t e = r_ent[s];
              const sgx_block* b = e < na ? la + e : lb + (e - na);
              const bool tr = (b->af & SGX_AF_TARGETREV) != 0;
              const uint32_t* inv = tr ? sf.inv : sr.inv;  // opposite (right-extension) index
              uint32_t size = 1;
              uint32_t mn = inv[r_tgt[s]];
              while (s + size < R && !(r_flag[s + size] & 2u)) {
                mn = min(mn, inv[r_tgt[s + size]]);
                ++size;
              }
              sgx_block o = *b;
              o.capped[0] = r_row[s];
              o.capped[1] = (uint64_t)r_row[s] + size - 1;
              o.capped[2] = mn;
              o.capped[3] = (uint64_t)mn + size - 1;
              out[at] = o;
            }
          }
          nout += __popc(em);
        }
        if (nout > bt.out_cap) slow = true;
        __syncwarp();
      }
    }
    if (lane == 0) {
      if (slow) {
        slow_list[atomicAdd(n_slow, 1u)] = (uint32_t)r;
      } else {
        bt.nout[r] = nout;
      }
    }
    __syncwarp();
  }
}

}  // namespace sgx

#endif  // SGX_OVERLAP_REDUCE_CUH_
