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
//   output    BLOCKS: rows of one block with identical ext are identical reads and contiguous; each surviving class
//             is one output block: capped[0] = its row range, capped[1] = the rank range of those reads in the
//             opposite index (inverse .sai), raw/length/af copied -- bit-identical to what the '$'-terminated walk
//             yields.  EDGES: every surviving row is one candidate edge; the filter of Hit2OverlapConverter::convert
//             (src/overlap_builder.cpp:358-367) is applied on the spot and only sgx_edge records leave the kernel.
//
// Anything outside these preconditions (interval intersections inside a list = periodic sequence, targets of
// different lengths, more rows / edges than the shared-memory tables hold) is NOT approximated: the read is put on
// the slow list and redone by the literal BWT-walking implementation (ov_post_kernel).
#ifndef SGX_OVERLAP_REDUCE_CUH_
#define SGX_OVERLAP_REDUCE_CUH_

#include "overlap_logic.h"

namespace sgx {

struct SaView {
  const uint32_t* sa;   // row of the '$' block -> read id
  const uint32_t* inv;  // read id -> row
};

// Where EDGES mode puts its output.
struct EdgeSink {
  sgx_edge* edges;
  unsigned long long* counter;  // edges written so far
  uint64_t cap;
  uint32_t* overflow;           // set when cap was too small (host regrows and reruns)
  const uint32_t* name_rank;    // per indexed read
  uint64_t qbase;               // global index of the first read of this launch
};

static const int kRdWarps = 4;     // warps (reads in flight) per CTA
static const int kRdEdgeCap = 256; // staged edges per read in EDGES mode

// Bytes of shared memory one warp needs for `rows_cap` rows of `extw` 64-bit extension words.
__host__ __device__ inline size_t reduce_smem_per_warp(uint32_t rows_cap, uint32_t extw, bool edges) {
  return (size_t)rows_cap * extw * 8 + (size_t)rows_cap * (4 + 4 + 2 + 2 + 2 + 2) + (edges ? (size_t)kRdEdgeCap * 8 : 0);
}

__device__ __forceinline__ bool rd_intersect(uint64_t s1, uint64_t e1, uint64_t s2, uint64_t e2) {
  return !(s1 > e2 || s2 > e1);
}

struct RdEnt {
  uint32_t lo, hi, len, af;
};

template <bool LEAN>
__device__ __forceinline__ RdEnt rd_cand(const OvBatch& bt, uint64_t r, int f, uint32_t i) {
  RdEnt e;
  if (LEAN) {
    const LeanCand c = bt.lean[((size_t)r * 4 + f) * bt.find_cap + i];
    e.lo = c.lo;
    e.hi = c.hi;
    e.len = c.len_af & 0xffffu;
    e.af = c.len_af >> 16;
  } else {
    const sgx_block* b = bt.cand + ((size_t)r * 4 + f) * bt.find_cap + i;
    e.lo = (uint32_t)b->capped[0];
    e.hi = (uint32_t)b->capped[1];
    e.len = b->length;
    e.af = b->af;
  }
  return e;
}

template <bool LEAN>
__device__ __forceinline__ RdEnt rd_contain(const OvBatch& bt, uint64_t r, int q) {
  RdEnt e;
  if (LEAN) {
    const LeanCand c = bt.lean_contain[(size_t)r * 4 + q];
    e.lo = c.lo;
    e.hi = c.hi;
    e.len = c.len_af & 0xffffu;
    e.af = c.len_af >> 16;
  } else {
    const sgx_block* b = bt.contain + (size_t)r * 4 + q;
    e.lo = (uint32_t)b->capped[0];
    e.hi = (uint32_t)b->capped[1];
    e.len = b->length;
    e.af = b->af;
  }
  return e;
}

// Hit2OverlapConverter::convert's filter (src/overlap_builder.cpp:358-367) for query (rank rq, length lq) and
// target t with overlap k.
__device__ __forceinline__ bool rd_edge_ok(uint32_t rq, uint32_t lq, uint32_t t, uint32_t k, uint32_t af,
                                           const uint32_t* __restrict__ name_rank, const uint32_t* __restrict__ tlen) {
  const uint32_t rt = name_rank[t];
  if (rq == rt) return false;                    // query.name == target.name
  const bool contain = k == lq || k == tlen[t];  // either coordinate spans its whole read
  return !(rq < rt || (contain && (af & SGX_AF_QUERYREV)));
}

// 16 bases (32 bits) of a 2-bit packed read starting at base `pos` (the read store is padded by two words).
__device__ __forceinline__ uint32_t rd_get16(const uint32_t* __restrict__ w, uint32_t pos) {
  return __funnelshift_r(w[pos >> 4], w[(pos >> 4) + 1], 2 * (pos & 15));
}
// reverse the order of the sixteen 2-bit groups of a word
__device__ __forceinline__ uint32_t rd_rev16(uint32_t x) {
  x = __brev(x);
  return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}
// Bases [16*d, 16*d+16) of the extension string of a row: forward targets continue after the overlap
// (t[k + p]); TARGETREV targets are read backwards (t[lt-1-k-p]).  Garbage beyond the extension is masked by the caller.
__device__ __forceinline__ uint32_t rd_ext16(const uint32_t* __restrict__ w, uint32_t lt, uint32_t k, bool tr, uint32_t d) {
  if (!tr) return rd_get16(w, k + 16 * d);
  const int s0 = (int)lt - (int)k - 16 * (int)d - 16;  // first forward base of this chunk
  const uint32_t x = s0 >= 0 ? rd_get16(w, (uint32_t)s0) : (s0 > -16 ? (w[0] << (2 * (uint32_t)(-s0))) : 0u);
  return rd_rev16(x);
}

// Warp-cooperative append of up to 32 staged edges (one per lane with `ok`).  Returns false on staging overflow.
__device__ __forceinline__ bool rd_stage(bool ok, uint32_t t, uint32_t k, uint32_t af, uint32_t* e_tgt, uint32_t* e_kaf,
                                         uint32_t& ecount, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  const uint32_t n = __popc(m);
  if (ecount + n > (uint32_t)kRdEdgeCap) return false;
  if (ok) {
    const uint32_t at = ecount + __popc(m & ((1u << lane) - 1u));
    e_tgt[at] = t;
    e_kaf[at] = k | (af << 16);
  }
  ecount += n;
  return true;
}

template <bool LEAN, bool EDGES>
__global__ void __launch_bounds__(kRdWarps * 32)
ov_reduce_kernel(ReadsView qreads, ReadsView treads, SaView sf, SaView sr, uint64_t count, int irreducible, OvBatch bt,
                 uint32_t rows_cap, uint32_t extw, uint64_t slow_base, uint32_t* __restrict__ slow_list,
                 uint32_t* __restrict__ n_slow, EdgeSink sink) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  unsigned char* base = smem_raw + (size_t)wib * reduce_smem_per_warp(rows_cap, extw, EDGES);
  uint64_t* ext = reinterpret_cast<uint64_t*>(base);                             // [rows_cap][extw]
  uint32_t* r_row = reinterpret_cast<uint32_t*>(ext + (size_t)rows_cap * extw);  // row in the '$' block
  uint32_t* r_tgt = r_row + rows_cap;                                            // target read id
  uint16_t* r_ent = reinterpret_cast<uint16_t*>(r_tgt + rows_cap);               // entry index within the side
  uint16_t* r_k = r_ent + rows_cap;                                              // overlap length
  uint16_t* r_n = r_k + rows_cap;                                                // extension length (bases)
  uint16_t* r_flag = r_n + rows_cap;                                             // bit0 transitive, bit1 class head
  uint32_t* e_tgt = reinterpret_cast<uint32_t*>(r_flag + rows_cap);              // EDGES: staged edges of this read
  uint32_t* e_kaf = e_tgt + kRdEdgeCap;

  const uint64_t warp0 = (uint64_t)blockIdx.x * kRdWarps + wib;
  const uint64_t nwarps = (uint64_t)gridDim.x * kRdWarps;
  for (uint64_t r = warp0; r < count; r += nwarps) {
    uint32_t nc[4], nk[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      nc[f] = bt.ncand[r * 4 + f];
      nk[f] = bt.ncontain[r * 4 + f];
    }
    sgx_block* out = EDGES ? nullptr : bt.out + (size_t)r * bt.out_cap;
    bool slow = false;
    // query side of the edge filter (EDGES)
    const uint64_t qglob = sink.qbase + r;
    uint32_t lq = 0, rq = 0;
    if (EDGES) {
      lq = qreads.len[qglob];
      rq = sink.name_rank[qglob];
    }
    uint32_t ecount = 0;

    // The SubMaximalBlockFilter precondition (no two capped[0] intervals of a list intersect, so that the filter
    // only reorders, src/overlap_builder.cpp:929-953) is checked on the row table of each side below: lists carry
    // the containment blocks of their index side, and two entries of a list intersect iff they share a row.
    uint32_t nout = 0;
    if (!slow) {
      // containment blocks go straight out: containfwd then containrev (src/overlap_builder.cpp:1160-1162)
      for (int q = 0; q < 4 && !slow; ++q) {
        if (!nk[q]) continue;
        if (EDGES) {
          const RdEnt c = rd_contain<LEAN>(bt, r, q);
          const uint32_t* sa = (c.af & SGX_AF_TARGETREV) ? sr.sa : sf.sa;
          for (uint32_t x0 = c.lo; x0 <= c.hi && !slow; x0 += 32) {
            const uint32_t x = x0 + lane;
            bool ok = false;
            uint32_t t = 0;
            if (x <= c.hi) {
              t = sa[x];
              ok = rd_edge_ok(rq, lq, t, c.len, c.af, sink.name_rank, treads.len);
            }
            if (!rd_stage(ok, t, c.len, c.af, e_tgt, e_kaf, ecount, lane)) slow = true;
            if (x0 > 0xffffffffu - 32) break;
          }
        } else {
          if (lane == 0) out[nout] = bt.contain[(size_t)r * 4 + q];
          ++nout;
        }
      }
    }
    if (!slow) {
#pragma unroll 1
      for (int side = 0; side < 2 && !slow; ++side) {
        const int fa = side == 0 ? 0 : 1, fb = side == 0 ? 3 : 2;
        const uint32_t na = nc[fa], ne = nc[fa] + nc[fb];
        // ---- rows: one per (entry, row of its capped[0] interval)
        uint32_t R = 0;
        for (uint32_t e0 = 0; e0 < ne && !slow; e0 += 32) {
          const uint32_t e = e0 + lane;
          uint32_t lo = 0, w = 0;
          if (e < ne) {
            const RdEnt b = e < na ? rd_cand<LEAN>(bt, r, fa, e) : rd_cand<LEAN>(bt, r, fb, e - na);
            lo = b.lo;
            const uint64_t ww = (uint64_t)b.hi - b.lo + 1;
            w = ww > rows_cap ? rows_cap + 1 : (uint32_t)ww;
          }
          uint32_t incl = w;
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
          const uint32_t at = R + incl - w;
          for (uint32_t x = 0; x < w; ++x) {
            r_row[at + x] = lo + x;
            r_ent[at + x] = (uint16_t)e;
          }
          R += tot;
        }
        if (slow) break;
        // ---- SubMaximalBlockFilter precondition, on the entries: the capped[0] intervals of one list (its candidate
        //      blocks plus the containment blocks of its index side, which belong to both lists of that side) must be
        //      pairwise disjoint
        bool inter = false;
        for (int h = 0; h < 2; ++h) {
          const int f = h == 0 ? fa : fb;
          const int q0 = f < 2 ? 0 : 2;
          const uint32_t nE = nc[f];
          for (uint32_t i = lane; i < nE + 2; i += 32) {
            RdEnt a;
            if (i < nE) a = rd_cand<LEAN>(bt, r, f, i);
            else if (nk[q0 + (i - nE)]) a = rd_contain<LEAN>(bt, r, q0 + (int)(i - nE));
            else continue;
            for (uint32_t i2 = 0; i2 < i; ++i2) {
              RdEnt c;
              if (i2 < nE) c = rd_cand<LEAN>(bt, r, f, i2);
              else if (nk[q0 + (i2 - nE)]) c = rd_contain<LEAN>(bt, r, q0 + (int)(i2 - nE));
              else continue;
              if (rd_intersect(a.lo, a.hi, c.lo, c.hi)) inter = true;
            }
          }
        }
        if (__any_sync(0xffffffffu, inter)) {
          slow = true;  // intervals of one list intersect (periodic sequence): SubMaximalBlockFilter must run
          break;
        }
        const uint32_t Rc = R;  // candidate rows
        __syncwarp();
        // ---- per candidate row: target read, extension string in query orientation (word level)
        //      (exhaustive mode, src/overlap_builder.cpp:1175-1178, only needs the rows and the intersection test)
        uint32_t len0 = 0;
        bool mixed = false;
        for (uint32_t s = lane; s < Rc && irreducible; s += 32) {
          const uint32_t e = r_ent[s];
          const RdEnt b = e < na ? rd_cand<LEAN>(bt, r, fa, e) : rd_cand<LEAN>(bt, r, fb, e - na);
          const uint32_t k = b.len, af = b.af;
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
          const uint32_t cm = (af & SGX_AF_QUERYCOMP) ? 0xffffffffu : 0u;
          uint64_t* dst = ext + (size_t)s * extw;
          if (!mixed) {
            for (uint32_t q = 0; q < extw; ++q) {
              uint64_t v = 0;
              if (32 * q < n) {
                const uint32_t lo = rd_ext16(w, lt, k, tr, 2 * q) ^ cm;
                const uint32_t hi = (32 * q + 16 < n) ? (rd_ext16(w, lt, k, tr, 2 * q + 1) ^ cm) : 0u;
                v = ((uint64_t)hi << 32) | lo;
                const uint32_t left = n - 32 * q;
                if (left < 32) v &= (1ull << (2 * left)) - 1ull;
              }
              dst[q] = v;
            }
          }
        }
        __syncwarp();
        if (Rc && irreducible) len0 = r_flag[0];
        for (uint32_t s = lane; s < Rc && irreducible; s += 32)
          if (r_flag[s] != len0) mixed = true;
        if (__any_sync(0xffffffffu, mixed)) {
          slow = true;  // targets of different lengths: group semantics are not the simple rule -> literal path
          break;
        }
        __syncwarp();
        // ---- transitive rows and class heads.  Rows with a longer overlap are scanned from the end of the table
        //      (entries of a list come in increasing overlap length): the longest overlaps have the shortest
        //      extensions and are the likeliest prefixes, so the scan usually stops after a few rows.
        for (uint32_t s = lane; s < Rc && irreducible; s += 32) {
          const uint32_t es = r_ent[s];
          const uint32_t ka = r_k[s];
          const uint64_t* ea = ext + (size_t)s * extw;
          bool red = false;
          for (uint32_t b2 = Rc; b2-- > 0 && !red;) {
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
          if (!EDGES && s > 0 && r_ent[s - 1] == es) {
            const uint64_t* ep = ext + (size_t)(s - 1) * extw;
            bool same = true;
            for (uint32_t q = 0; q < extw && same; ++q) same = ea[q] == ep[q];
            head = !same;
          }
          r_flag[s] = (uint16_t)((red ? 1u : 0u) | (head ? 2u : 0u));
        }
        if (!irreducible)
          for (uint32_t s = lane; s < Rc; s += 32) r_flag[s] = 2u;
        __syncwarp();
        if (EDGES) {
          // ---- every surviving row is one candidate edge
          for (uint32_t s0 = 0; s0 < R && !slow; s0 += 32) {
            const uint32_t s = s0 + lane;
            bool ok = false;
            uint32_t t = 0, k = 0, af = 0;
            if (s < R && !(r_flag[s] & 1u)) {
              const uint32_t e = r_ent[s];
              const RdEnt b = e < na ? rd_cand<LEAN>(bt, r, fa, e) : rd_cand<LEAN>(bt, r, fb, e - na);
              t = irreducible ? r_tgt[s] : ((b.af & SGX_AF_TARGETREV) ? sr.sa : sf.sa)[r_row[s]];
              k = b.len;
              af = b.af;
              ok = rd_edge_ok(rq, lq, t, k, af, sink.name_rank, treads.len);
            }
            if (!rd_stage(ok, t, k, af, e_tgt, e_kaf, ecount, lane)) slow = true;
          }
        } else {
          // ---- one output block per surviving class (exhaustive mode: the candidate blocks themselves)
          const sgx_block* la = bt.cand + ((size_t)r * 4 + fa) * bt.find_cap;
          const sgx_block* lb = bt.cand + ((size_t)r * 4 + fb) * bt.find_cap;
          if (!irreducible) {
            if (nout + ne > bt.out_cap) {
              slow = true;
              break;
            }
            for (uint32_t i = lane; i < ne; i += 32) out[nout + i] = i < na ? la[i] : lb[i - na];
            nout += ne;
          }
          for (uint32_t s0 = 0; s0 < R && irreducible; s0 += 32) {
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
        }
        __syncwarp();
      }
    }
    if (EDGES && !slow) {
      // one reservation per read, then a coalesced copy of the staged edges
      __syncwarp();
      unsigned long long at = 0;
      if (lane == 0 && ecount) at = atomicAdd(sink.counter, (unsigned long long)ecount);
      at = __shfl_sync(0xffffffffu, at, 0);
      if (at + ecount > sink.cap) {
        if (lane == 0 && ecount) atomicOr(sink.overflow, 1u);
      } else {
        for (uint32_t i = lane; i < ecount; i += 32) {
          sgx_edge e;
          e.query = (uint32_t)qglob;
          e.target = e_tgt[i];
          e.length = e_kaf[i] & 0xffffu;
          e.af = e_kaf[i] >> 16;
          sink.edges[at + i] = e;
        }
      }
    }
    if (lane == 0) {
      if (slow) {
        slow_list[atomicAdd(n_slow, 1u)] = (uint32_t)(slow_base + r);
      } else if (!EDGES) {
        bt.nout[r] = nout;
      }
    }
    __syncwarp();
  }
}


// ---- lean reduce (sgx_overlap_edges): the same rule on the 12-byte candidates, reorganised for latency ------------
//
// ov_reduce_kernel above serves both output forms; for the edge pipeline it walks a read's two sides one after the other,
// each a chain of dependent gathers (candidate -> .sai row -> length -> word offset -> read words -> name rank), and
// re-reads the candidates from global memory in every phase.  This kernel does the work of one read in ONE pass:
//   * the candidate entries of all four lists and the containment entries are copied to shared memory once (coalesced);
//   * the rows of both sides and of the containment blocks go into one table, so that the gathers of every row of the
//     read (.sai, read words, name rank) are in flight together: four dependent round trips per read instead of twelve;
//   * reads of one common length (known when the read store was packed) need no gather of len[] / woff[] at all;
//   * edges are counted with a ballot, reserved with one atomic per read and written straight from the row table (no
//     staging area), which shrinks the shared-memory footprint per warp.
// Preconditions and fallbacks are those of ov_reduce_kernel (anything else goes to the slow list, nothing is approximated);
// a read whose rows do not fit the table in one pass takes its sides one at a time.
static const int kRlWarps = 4;
static const int kRlEntCap = 96;   // candidate + containment entries per read held in shared memory

__host__ __device__ inline size_t reduce_lean_smem_per_warp(uint32_t rows_cap, uint32_t extw) {
  return (size_t)rows_cap * extw * 8 + (size_t)rows_cap * (4 + 4 + 2 + 1 + 1) + (size_t)kRlEntCap * 12;
}

template <bool UNIFORM>
__global__ void __launch_bounds__(kRlWarps * 32)
ov_reduce_lean_kernel(ReadsView qreads, ReadsView treads, uint32_t ulen, uint32_t uwords, SaView sf, SaView sr,
                      uint64_t count, int irreducible, OvBatch bt, uint32_t rows_cap, uint32_t extw, uint64_t slow_base,
                      uint32_t* __restrict__ slow_list, uint32_t* __restrict__ n_slow, EdgeSink sink) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  unsigned char* base = smem_raw + (size_t)wib * reduce_lean_smem_per_warp(rows_cap, extw);
  uint64_t* ext = reinterpret_cast<uint64_t*>(base);                             // [rows_cap][extw]
  uint32_t* r_row = reinterpret_cast<uint32_t*>(ext + (size_t)rows_cap * extw);  // row in the '$' block
  uint32_t* r_tgt = r_row + rows_cap;                                            // target read id
  uint16_t* r_n = reinterpret_cast<uint16_t*>(r_tgt + rows_cap);                 // extension length (bases)
  uint8_t* r_ent = reinterpret_cast<uint8_t*>(r_n + rows_cap);                   // entry index
  uint8_t* r_flag = r_ent + rows_cap;                                            // bit0 transitive, bit1 passes the edge filter
  uint32_t* e_lo = reinterpret_cast<uint32_t*>(r_flag + rows_cap);               // entries: capped[0], length | af << 16
  uint32_t* e_hi = e_lo + kRlEntCap;
  uint32_t* e_kaf = e_hi + kRlEntCap;

  const uint64_t warp0 = (uint64_t)blockIdx.x * kRlWarps + wib;
  const uint64_t nwarps = (uint64_t)gridDim.x * kRlWarps;
  for (uint64_t r = warp0; r < count; r += nwarps) {
    // ---- entries: lists 0..3 (search modes), then the containment entry of each mode that has one
    const uint4 nc4 = *reinterpret_cast<const uint4*>(bt.ncand + r * 4);
    const uint4 nk4 = *reinterpret_cast<const uint4*>(bt.ncontain + r * 4);
    const uint32_t nc[4] = {nc4.x, nc4.y, nc4.z, nc4.w};
    const uint32_t nk[4] = {nk4.x, nk4.y, nk4.z, nk4.w};
    const uint32_t eb1 = nc[0], eb2 = eb1 + nc[1], eb3 = eb2 + nc[2], Ec = eb3 + nc[3];  // list f starts at eb_f
    const uint32_t E = Ec + 4;  // containment slots Ec + q; absent ones are empty intervals
    const uint64_t qglob = sink.qbase + r;
    const uint32_t lq = UNIFORM ? ulen : qreads.len[qglob];
    const uint32_t rq = sink.name_rank[qglob];
    bool slow = E > (uint32_t)kRlEntCap;
    if (!slow) {
      for (uint32_t e = lane; e < E; e += 32) {
        LeanCand c;
        if (e < Ec) {
          const int f = (e >= eb1) + (e >= eb2) + (e >= eb3);
          const uint32_t i = e - (f == 0 ? 0u : f == 1 ? eb1 : f == 2 ? eb2 : eb3);
          c = bt.lean[((size_t)r * 4 + f) * bt.find_cap + i];
        } else if (nk[e - Ec]) {
          c = bt.lean_contain[(size_t)r * 4 + (e - Ec)];
        } else {
          c.lo = 1;
          c.hi = 0;
          c.len_af = 0;
        }
        e_lo[e] = c.lo;
        e_hi[e] = c.hi;
        e_kaf[e] = c.len_af;
      }
      __syncwarp();
      // ---- SubMaximalBlockFilter precondition (src/overlap_builder.cpp:929-953): inside one list -- its candidates plus the
      //      containment entries of its index side -- no two capped[0] intervals may intersect
      bool inter = false;
      for (uint32_t e = lane; e < Ec; e += 32) {
        const int f = (e >= eb1) + (e >= eb2) + (e >= eb3);
        const uint32_t f0 = f == 0 ? 0u : f == 1 ? eb1 : f == 2 ? eb2 : eb3;
        const uint32_t lo = e_lo[e], hi = e_hi[e];
        for (uint32_t e2 = f0; e2 < e; ++e2) inter |= rd_intersect(lo, hi, e_lo[e2], e_hi[e2]);
        const uint32_t q0 = Ec + (f < 2 ? 0u : 2u);
        inter |= (e_hi[q0] >= e_lo[q0] && rd_intersect(lo, hi, e_lo[q0], e_hi[q0])) ||
                 (e_hi[q0 + 1] >= e_lo[q0 + 1] && rd_intersect(lo, hi, e_lo[q0 + 1], e_hi[q0 + 1]));
      }
      if (lane < 2) {  // the two containment entries of an index side against each other
        const uint32_t q0 = Ec + 2 * lane;
        inter |= e_hi[q0] >= e_lo[q0] && e_hi[q0 + 1] >= e_lo[q0 + 1] && rd_intersect(e_lo[q0], e_hi[q0], e_lo[q0 + 1], e_hi[q0 + 1]);
      }
      if (__any_sync(kFull, inter)) slow = true;  // periodic sequence: the filter must run (literal path)
    }
    // ---- row counts per side (lists 0 and 3 extend to the right, 1 and 2 to the left) and of the containment entries
    unsigned long long at = 0;
    if (!slow) {
      uint32_t w[3] = {0, 0, 0};
      for (uint32_t e = lane; e < E; e += 32) {
        const uint64_t ww = e_hi[e] >= e_lo[e] ? (uint64_t)e_hi[e] - e_lo[e] + 1 : 0;
        const uint32_t wc = ww > rows_cap ? rows_cap + 1 : (uint32_t)ww;
        const int g = e >= Ec ? 2 : ((e < eb1 || e >= eb3) ? 0 : 1);
        w[g] = min(w[g] + wc, rows_cap + 1);
      }
#pragma unroll
      for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w[g] = min(w[g] + __shfl_xor_sync(kFull, w[g], o), rows_cap + 1);
      }
      // passes: everything at once when it fits, else one group (side / containment) at a time
      // (one group at a time only for reads of one length: there no check can fail after a pass has written edges)
      const bool one_pass = w[0] + w[1] + w[2] <= rows_cap;
      if (!one_pass && (!UNIFORM || w[0] > rows_cap || w[1] > rows_cap || w[2] > rows_cap)) slow = true;
      const int npass = one_pass ? 1 : 3;
      for (int pass = 0; pass < npass && !slow; ++pass) {
        // ---- rows of this pass: (entry, row of its interval), entries in list order 0,3 | 1,2 | containment
        uint32_t R = 0, side_end[3] = {0, 0, 0};
#pragma unroll 1
        for (int g = 0; g < 3; ++g) {
          if (one_pass || g == pass) {
            // entries of group g in the order the reference concatenates the lists: g0 = list 0 then 3, g1 = 1 then 2
            const uint32_t aS = g == 0 ? 0u : g == 1 ? eb1 : Ec, aE = g == 0 ? eb1 : g == 1 ? eb2 : E;
            const uint32_t bS = g == 0 ? eb3 : g == 1 ? eb2 : E, bE = g == 0 ? Ec : g == 1 ? eb3 : E;
            const uint32_t ne = (aE - aS) + (bE - bS);
            for (uint32_t i0 = 0; i0 < ne; i0 += 32) {
              const uint32_t i = i0 + lane;
              uint32_t e = 0, wd = 0;
              if (i < ne) {
                e = i < aE - aS ? aS + i : bS + (i - (aE - aS));
                wd = e_hi[e] >= e_lo[e] ? e_hi[e] - e_lo[e] + 1 : 0;
              }
              uint32_t incl = wd;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += y;
              }
              const uint32_t a = R + incl - wd;
              for (uint32_t x = 0; x < wd; ++x) {
                r_row[a + x] = e_lo[e] + x;
                r_ent[a + x] = (uint8_t)e;
              }
              R += __shfl_sync(kFull, incl, 31);
            }
          }
          side_end[g] = R;
        }
        __syncwarp();
        // ---- per row: target read, edge filter, extension string in query orientation (word level)
        bool mixed = false;
        uint32_t len_first[2] = {0xffffffffu, 0xffffffffu};
        for (uint32_t s = lane; s < R; s += 32) {
          const uint32_t e = r_ent[s];
          const uint32_t kaf = e_kaf[e];
          const uint32_t k = kaf & 0xffffu, af = kaf >> 16;
          const bool tr = (af & SGX_AF_TARGETREV) != 0;
          const uint32_t t = (tr ? sr.sa : sf.sa)[r_row[s]];
          const uint32_t lt = UNIFORM ? ulen : treads.len[t];
          const uint32_t rt = sink.name_rank[t];
          r_tgt[s] = t;
          const bool contain = k == lq || k == lt;
          const bool okf = rq != rt && !(rq < rt || (contain && (af & SGX_AF_QUERYREV)));
          uint8_t flag = okf ? 2 : 0;
          if (s < side_end[1] && irreducible) {
            const int g = s < side_end[0] ? 0 : 1;
            if (len_first[g] == 0xffffffffu) len_first[g] = lt;
            if (len_first[g] != lt) mixed = true;
            const uint32_t n = lt - k;
            if (n > extw * 32) mixed = true;  // cannot happen when extw covers max_len - min_overlap
            r_n[s] = (uint16_t)n;
            const uint32_t* wsrc = treads.words + (UNIFORM ? (uint64_t)t * uwords : treads.woff[t]);
            const uint32_t cm = (af & SGX_AF_QUERYCOMP) ? 0xffffffffu : 0u;
            uint64_t* dst = ext + (size_t)s * extw;
            if (!mixed) {
              for (uint32_t q = 0; q < extw; ++q) {
                uint64_t v = 0;
                if (32 * q < n) {
                  const uint32_t lo = rd_ext16(wsrc, lt, k, tr, 2 * q) ^ cm;
                  const uint32_t hi = (32 * q + 16 < n) ? (rd_ext16(wsrc, lt, k, tr, 2 * q + 1) ^ cm) : 0u;
                  v = ((uint64_t)hi << 32) | lo;
                  const uint32_t left = n - 32 * q;
                  if (left < 32) v &= (1ull << (2 * left)) - 1ull;
                }
                dst[q] = v;
              }
            }
          }
          r_flag[s] = flag;
        }
        if (!UNIFORM && irreducible) {
          // all targets of a side must have one length (else the group semantics are not the simple rule)
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t mn = __reduce_min_sync(kFull, len_first[g]);
            if (len_first[g] != 0xffffffffu && len_first[g] != mn) mixed = true;
          }
        }
        if (__any_sync(kFull, mixed)) {
          slow = true;
          break;
        }
        __syncwarp();
        // ---- transitive rows: some row of the same side with a longer overlap has its whole extension as a prefix.
        //      Rows with longer overlaps are scanned from the end of the side (entries come in increasing overlap length).
        for (uint32_t s = lane; s < side_end[1] && irreducible; s += 32) {
          const uint32_t s0 = s < side_end[0] ? 0u : side_end[0], s1 = s < side_end[0] ? side_end[0] : side_end[1];
          const uint32_t ka = e_kaf[r_ent[s]] & 0xffffu;
          const uint64_t* ea = ext + (size_t)s * extw;
          bool red = false;
          for (uint32_t b2 = s1; b2-- > s0 && !red;) {
            if ((e_kaf[r_ent[b2]] & 0xffffu) <= ka) continue;
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
          if (red) r_flag[s] |= 1;
        }
        __syncwarp();
        // ---- edges of this pass: count, reserve, write
        uint32_t cnt = 0;
        for (uint32_t s = lane; s < R; s += 32) cnt += r_flag[s] == 2 ? 1u : 0u;
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += y;
        }
        const uint32_t tot = __shfl_sync(kFull, incl, 31);
        if (tot) {
          if (lane == 0) at = atomicAdd(sink.counter, (unsigned long long)tot);
          at = __shfl_sync(kFull, at, 0);
          if (at + tot > sink.cap) {
            if (lane == 0) atomicOr(sink.overflow, 1u);
          } else {
            unsigned long long o = at + incl - cnt;
            for (uint32_t s = lane; s < R; s += 32) {
              if (r_flag[s] == 2) {
                const uint32_t kaf = e_kaf[r_ent[s]];
                sgx_edge ed;
                ed.query = (uint32_t)qglob;
                ed.target = r_tgt[s];
                ed.length = kaf & 0xffffu;
                ed.af = kaf >> 16;
                sink.edges[o++] = ed;
              }
            }
          }
        }
        __syncwarp();
      }
    }
    // every check that can fail (entry / row capacity, intersections, mixed target lengths) fails before the first edge
    // of the read is written: in one-pass mode they all precede the write, group-at-a-time mode is for uniform reads only
    if (lane == 0 && slow) slow_list[atomicAdd(n_slow, 1u)] = (uint32_t)(slow_base + r);
    __syncwarp();
  }
}

}  // namespace sgx

#endif  // SGX_OVERLAP_REDUCE_CUH_
