// `siga correct` on the device: KmerCorrector::process (src/correct_processor.cpp:72-232) for a batch of reads.
//
// One warp per read (reads are handed out dynamically: a read that needs ten correction rounds costs many times a clean
// one).  The control logic is correct_logic.h, executed uniformly by the 32 lanes on per-warp state in shared memory
// (working sequence, quality bits, solid / decided bits, remembered rare strings).  Its two parallel primitives --
// certificate searches over runs of undecided k-mers (windows) and "does this k-mer with one base substituted reach
// this count?" (qualify) -- each give up to eight queries to the eight 4-lane quads of the warp.  A quad runs a plain
// backward search (FMIndex::Interval::update, src/fmindex.h:96-100) on the 128-byte rank blocks with quad_rank_sym
// (quad_rank.cuh): the rank of the ONE stepped symbol at both interval ends from one or two 128-byte lines.
// Bound: random 128-byte access, like the overlap search; algorithmic work = executed rank lookups x 128 B.  The
// reference's way of counting costs 2 (k-1) (n-k+1) lookups per counting round (7200 at n = 150, k = 31, eight times
// the four overlap searches of the same read); the certificates settle the same solid bits with about a third
// (correct_logic.h).  Measured: profiles/r1h_summary.md -- the kernel is issue-bound, hence the branch-free loops.
//
// Output is compact: one flag per read (written or dropped by PostCorrector, src/correct_processor.cpp:258-264) and
// the list of substituted bases of the written reads.
#include "correct.cuh"

#include <algorithm>

#include "correct_logic.h"
#include "quad_rank.cuh"

namespace sgx {

template <class IdxT>
struct DevCorrEng {
  const FmBlockLane* __restrict__ blocks;
  const uint64_t* pred;          // C[] of the index by rank (pred[5] = n): the kernel PARAMETER itself, so that pred[c] is
                                 // one constant-bank load with a register index (selecting among five registers
                                 // compiles to a branch tree)
  const uint32_t* seq;           // working sequence of this warp's read (shared memory)
  uint32_t k;
  IdxT low, high;                // the two support thresholds of the certificate searches (clamp_count)
  bool two;                      // qualities present: k-mers of both threshold classes exist
  int lane, q, qbase, slot;
  unsigned long long nlook;

  __device__ __forceinline__ bool leader() const { return lane == 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ uint32_t lane_id() const { return (uint32_t)lane; }
  __device__ __forceinline__ uint32_t num_lanes() const { return 32u; }
  __device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t m) const { atomicOr(p, m); }
  __device__ __forceinline__ IdxT pcv(int c) const { return (IdxT)pred[c]; }

  // One backward-search step of every quad of the warp: the interval [lo, lo + size) of the quad's string extended by
  // symbol c (FMIndex::Interval::update, src/fmindex.h:96-100; an empty interval is size 0, which is what the
  // reference's upper = lower - 1 amounts to).  Idle quads ride along with their loads pointed at block 0, so that every
  // shuffle is a full-mask shuffle (see ov_search_quad_kernel).
  __device__ __forceinline__ void step(bool active, int c, IdxT& lo, IdxT& size) {
    IdxT Lc, Uc;
    quad_rank_sym<IdxT>(blocks, active ? lo : (IdxT)0, active ? (IdxT)(lo + size) : (IdxT)0, q, qbase, c, Lc, Uc);
    // branch-free: the kernel is bound by issue slots, and every divergent region costs a BSSY/BSYNC pair
    nlook += active ? 2u : 0u;
    lo = active ? (IdxT)(pcv(c) + Lc) : lo;
    size = active ? (IdxT)(Uc - Lc) : size;
  }
  // Counts are compared in the position type: with 32-bit positions every count is below 2^32 - 1, so a larger
  // threshold behaves like 2^32 - 1 (never reached).
  __device__ __forceinline__ static IdxT clamp_count(uint64_t v) {
    return (sizeof(IdxT) == 4 && v > 0xffffffffull) ? (IdxT)0xffffffffu : (IdxT)v;
  }

  // try2Correct queries: all 32 lanes call this together; quad s answers query gen(s).
  template <class Gen>
  __device__ __forceinline__ uint32_t qualify(Gen gen) {
    const unsigned kFull = 0xffffffffu;
    const KmerQuery kq = gen(slot);
    const IdxT need = clamp_count(kq.min_count);
    uint32_t j = kq.have ? kq.kidx + k - 1 : 0u;  // read position of the last k-mer symbol: the search starts there
    int c = corr_query_sym(kq, seq, j);
    IdxT lo = pcv(c), size = (IdxT)(pcv(c + 1) - lo);
    uint32_t steps = k - 1;
    // the interval only shrinks: the search can stop as soon as it is smaller than the count asked for
    bool live = kq.have && steps > 0 && size > 0 && size >= need;
#pragma unroll 1
    while (__any_sync(kFull, live)) {
      j -= live ? 1u : 0u;  // a finished query keeps reading its last symbol (always a valid position)
      c = corr_query_sym(kq, seq, j);
      step(live, c, lo, size);
      steps -= live ? 1u : 0u;
      live = live && steps > 0 && size > 0 && size >= need;
    }
    const unsigned b = __ballot_sync(kFull, kq.have && size >= need && q == 0);
    uint32_t out = 0;
#pragma unroll
    for (int s = 0; s < kCorrSlots; ++s) out |= ((b >> (4 * s)) & 1u) << s;
    return out;
  }

  // Certificate searches (correct_logic.h: WinQuery): quad s searches read[a..e] of query gen(s) from e towards a and
  // stores {a, e, drop_low, drop_high} at out[4 s].  The caller synchronises before reading `out`.
  template <class Gen>
  __device__ __forceinline__ void windows(Gen gen, uint32_t* out) {
    const unsigned kFull = 0xffffffffu;
    const WinQuery wq = gen(slot);
    uint32_t pos = wq.have ? wq.e : 0u, dl = kCorrNone, dh = kCorrNone;
    bool fl = false, fh = !two;
    int c = read_rank(seq, pos);
    IdxT lo = pcv(c), size = (IdxT)(pcv(c + 1) - lo);
    bool alive = wq.have;
#pragma unroll 1
    for (;;) {
      const bool nl = alive && !fl && size < low, nh = alive && !fh && size < high;
      dl = nl ? pos : dl;
      dh = nh ? pos : dh;
      fl = fl || nl;
      fh = fh || nh;
      alive = alive && size != 0 && !(fl && fh) && pos != wq.a;
      if (!__any_sync(kFull, alive)) break;
      pos -= alive ? 1u : 0u;  // a finished search keeps reading its last symbol (always a valid position)
      c = read_rank(seq, pos);
      step(alive, c, lo, size);
    }
    if (wq.have && q == 0) {
      out[4 * slot] = wq.a;
      out[4 * slot + 1] = wq.e;
      out[4 * slot + 2] = dl;
      out[4 * slot + 3] = dh;
    }
  }
};

// Where the quality bits of read g (index in the read store) start: reads own ceil(len/16) consecutive 32-bit words
// from woff[g] and need ceil(len/32) quality words; (woff[g] + g) >> 1 leaves room for exactly that.
__device__ __forceinline__ uint64_t corr_hq_off(const ReadsView& reads, uint64_t g) { return (reads.woff[g] + g) >> 1; }

// bit j of a read's quality words: phred(j) >= 20 (Quality::Phred::fromchar, src/quality.h:30-34).  One warp per read.
__global__ void __launch_bounds__(256) corr_hq_kernel(const char* __restrict__ qual, const uint64_t* __restrict__ off,
                                                      ReadsView reads, uint64_t first, uint64_t count, uint64_t hbase,
                                                      uint32_t* __restrict__ hq, uint32_t* __restrict__ bad) {
  const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= count) return;
  const uint64_t g = first + i;
  const uint32_t l = reads.len[g];
  if (off[i + 1] - off[i] != (uint64_t)l) {
    if (lane == 0) atomicOr(bad, 1u);
    return;
  }
  const char* src = qual + (off[i] - off[0]);
  uint32_t* dst = hq + (corr_hq_off(reads, g) - hbase);
  for (uint32_t w = lane; w < (l + 31) / 32; w += 32) {
    uint32_t v = 0;
    const uint32_t je = min(32u, l - w * 32);
    for (uint32_t b = 0; b < je; ++b)
      if ((int)(uint8_t)src[w * 32 + b] - kCorrPhredBase >= kCorrPhredCutoff) v |= 1u << b;
    dst[w] = v;
  }
}

// Shared-memory words of one warp: sequence | quality bits, khigh, known, ksolid, unsolid, tmp | certificate results |
// remembered rare strings.
__host__ __device__ inline size_t corr_smem_words(uint32_t w16, uint32_t w32) {
  return (size_t)w16 + 6 * (size_t)w32 + kCorrWinWords + kCorrDeadCap;
}

struct CorrCounters {
  unsigned long long next_read, n_fixes, lookups, fixed_reads;
};

template <class IdxT>
__global__ void __launch_bounds__(256, 4)
corr_kernel(const __grid_constant__ FmView fm, ReadsView reads, uint64_t first, uint64_t count, const uint32_t* __restrict__ hq, uint64_t hbase,
            CorrectParams P, uint32_t w16, uint32_t w32, uint8_t* __restrict__ valid, sgx_fix* __restrict__ fixes,
            CorrCounters* __restrict__ ctr) {
  extern __shared__ uint32_t corr_smem[];
  const int lane = threadIdx.x & 31;
  uint32_t* mine = corr_smem + (size_t)(threadIdx.x >> 5) * corr_smem_words(w16, w32);
  CorrectState st;
  st.seq = mine;
  uint32_t* hqs = mine + w16;
  st.hq = hq ? hqs : nullptr;
  st.khigh = hqs + w32;
  st.known = st.khigh + w32;
  st.ksolid = st.known + w32;
  st.unsolid = st.ksolid + w32;
  st.tmp = st.unsolid + w32;
  st.win = st.tmp + w32;
  st.dead = st.win + kCorrWinWords;
  DevCorrEng<IdxT> eng;
  eng.blocks = fm.blocks;
  eng.pred = fm.pred;
  eng.seq = st.seq;
  eng.k = P.k;
  eng.low = DevCorrEng<IdxT>::clamp_count(P.low);
  eng.high = DevCorrEng<IdxT>::clamp_count(P.high);
  eng.two = hq != nullptr;
  eng.lane = lane;
  eng.q = lane & (kQuad - 1);
  eng.qbase = lane & ~(kQuad - 1);
  eng.slot = lane >> 2;
  eng.nlook = 0;
  unsigned long long fixed_reads = 0;
  for (;;) {
    unsigned long long r = 0;
    if (lane == 0) r = atomicAdd(&ctr->next_read, 1ull);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= count) break;
    const uint64_t g = first + r;
    const uint32_t n = reads.len[g];
    const uint32_t* __restrict__ src = reads.words + reads.woff[g];
    for (uint32_t w = lane; w < (n + 15) / 16; w += 32) st.seq[w] = src[w];
    if (hq) {
      const uint32_t* __restrict__ hsrc = hq + (corr_hq_off(reads, g) - hbase);
      for (uint32_t w = lane; w < (n + 31) / 32; w += 32) hqs[w] = hsrc[w];
    }
    __syncwarp();
    uint32_t n_fixed = 0;
    const bool ok = correct_read(eng, P, n, st, &n_fixed);
    __syncwarp();
    if (lane == 0) {
      valid[r] = ok ? 1 : 0;
      if (ok && n_fixed) {
        uint32_t cnt = 0;
        for (uint32_t w = 0; w < (n + 15) / 16; ++w) {
          const uint32_t d = src[w] ^ st.seq[w];
          cnt += __popc((d | (d >> 1)) & 0x55555555u);
        }
        if (cnt) {
          ++fixed_reads;
          sgx_fix* dst = fixes + atomicAdd(&ctr->n_fixes, (unsigned long long)cnt);
          for (uint32_t w = 0; w < (n + 15) / 16; ++w) {
            uint32_t d = src[w] ^ st.seq[w];
            d = (d | (d >> 1)) & 0x55555555u;
            while (d) {
              const int b = (__ffs(d) - 1) >> 1;
              d &= d - 1;
              const uint32_t pos = w * 16 + (uint32_t)b;
              dst->read = (uint32_t)r;
              dst->pos_code = (pos << 2) | ((st.seq[w] >> (2 * b)) & 3u);
              ++dst;
            }
          }
        }
      }
    }
    __syncwarp();
  }
  // every quad counted the steps of its own queries
  unsigned long long nl = eng.q == 0 ? eng.nlook : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nl += __shfl_xor_sync(0xffffffffu, nl, o);
  if (lane == 0) {
    if (nl) atomicAdd(&ctr->lookups, nl);
    if (fixed_reads) atomicAdd(&ctr->fixed_reads, fixed_reads);
  }
}

void correct_run(sgx_ctx* ctx, const sgx_fm* fm, const sgx_reads* reads, uint64_t first, uint64_t count, const char* qual,
                 const uint64_t* qual_off, const sgx_correct_params* prm, sgx_corrected* out) {
  SGX_REQUIRE(out, SGX_EINVAL, "sgx_correct_batch: null output");
  memset(out, 0, sizeof(*out));
  SGX_REQUIRE(fm && reads && prm, SGX_EINVAL, "sgx_correct_batch: null argument");
  SGX_REQUIRE(first <= reads->n && count <= reads->n - first, SGX_EINVAL, "sgx_correct_batch: read range out of bounds");
  SGX_REQUIRE(prm->kmer_size >= 1, SGX_EINVAL, "sgx_correct_batch: kmer-size must be at least 1");
  SGX_REQUIRE(reads->max_len <= 32768, SGX_EINVAL, "sgx_correct_batch: reads longer than 32768 bases are not supported");
  SGX_REQUIRE((qual == nullptr) == (qual_off == nullptr), SGX_EINVAL, "sgx_correct_batch: qual and qual_off go together");
  out->n_reads = count;
  if (count == 0) {
    out->valid = (uint8_t*)pinned_get(ctx, 1);
    out->fixes = (sgx_fix*)pinned_get(ctx, sizeof(sgx_fix));
    return;
  }
  const uint64_t launches0 = ctx->launches;
  EventTimer t_all(ctx->stream);
  KernelClock clk(ctx->stream);
  t_all.start();
  CorrectParams P;
  P.k = prm->kmer_size;
  // every round fixes one base or ends the read, and a fix makes its k-mer solid, so real reads stop after a handful of
  // rounds; the cap only bounds the kernel's run time should fixes ever chase each other in a cycle
  P.rounds = std::min<uint32_t>(prm->kmer_rounds, 1u << 20);
  P.low = (uint64_t)(int64_t)prm->kmer_threshold;  // int -> size_t like CorrectThreshold::requiredSupport
  P.high = (uint64_t)((int64_t)prm->kmer_threshold + 1);
  P.count_offset = prm->kmer_count_offset;
  const uint32_t w16 = (reads->max_len + 15) / 16, w32 = (reads->max_len + 31) / 32;

  // quality bits
  DevBuf<uint32_t> d_hq, d_bad(1);
  uint64_t hbase = 0;
  if (qual) {
    uint64_t wo[2];
    SGX_CUDA(cudaMemcpyAsync(&wo[0], reads->woff + first, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaMemcpyAsync(&wo[1], reads->woff + first + count, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    hbase = (wo[0] + first) >> 1;
    const uint64_t hwords = ((wo[1] + first + count) >> 1) - hbase + 1;
    const uint64_t qbytes = qual_off[count] - qual_off[0];
    SGX_REQUIRE(qual_off[count] >= qual_off[0] && qbytes > 0, SGX_EINVAL, "sgx_correct_batch: bad quality offsets");
    d_hq.alloc(hwords);
    DevBuf<char> d_q(qbytes);
    DevBuf<uint64_t> d_off(count + 1);
    SGX_CUDA(cudaMemcpyAsync(d_q.p, qual + qual_off[0], qbytes, cudaMemcpyHostToDevice, ctx->stream));
    SGX_CUDA(cudaMemcpyAsync(d_off.p, qual_off, (count + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    SGX_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(uint32_t), ctx->stream));
    SGX_LAUNCH(ctx, corr_hq_kernel, grid_for(count * 32, 256), 256, 0, (const char*)d_q.p, (const uint64_t*)d_off.p,
               reads->view(), first, count, hbase, d_hq.p, d_bad.p);
    uint32_t bad = 0;
    SGX_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));  // also keeps d_q / d_off alive until the kernel is done
    SGX_REQUIRE(bad == 0, SGX_EINVAL, "sgx_correct_batch: a quality string differs in length from its read");
  }

  const uint64_t per_read = std::min<uint64_t>(P.rounds, reads->max_len);
  DevBuf<sgx_fix> d_fix(std::max<uint64_t>(count * per_read, 1));
  DevBuf<uint8_t> d_valid(count);
  DevBuf<CorrCounters> d_ctr(1);
  SGX_CUDA(cudaMemsetAsync(d_ctr.p, 0, sizeof(CorrCounters), ctx->stream));
  // one warp per read, up to 8 warps per block; long reads (33 KB of state per warp at 32768 bases) get fewer warps so
  // that a block stays within the 227 KB of shared memory an SM can give one block
  const size_t per_warp = corr_smem_words(w16, w32) * sizeof(uint32_t);
  const unsigned warps = (unsigned)std::max<size_t>(1, std::min<size_t>(8, ((size_t)200 << 10) / per_warp));
  const size_t smem = warps * per_warp;
  const char* fw = getenv("SGX_FORCE_WIDE");  // tests: exercise the 64-bit position path on small inputs
  const bool wide = fm->n >= 0xffffffffull || (fw && fw[0] == '1');
  if (smem > 48 * 1024) {
    if (wide) SGX_CUDA(cudaFuncSetAttribute(corr_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else SGX_CUDA(cudaFuncSetAttribute(corr_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const unsigned grid = (unsigned)std::min<uint64_t>((count + warps - 1) / warps, (uint64_t)ctx->sm_count * 4 * (8 / warps));
  clk.tick();
  if (wide) {
    SGX_LAUNCH(ctx, corr_kernel<uint64_t>, grid, 32 * warps, smem, fm->view(), reads->view(), first, count, (const uint32_t*)d_hq.p,
               hbase, P, w16, w32, d_valid.p, d_fix.p, d_ctr.p);
  } else {
    SGX_LAUNCH(ctx, corr_kernel<uint32_t>, grid, 32 * warps, smem, fm->view(), reads->view(), first, count, (const uint32_t*)d_hq.p,
               hbase, P, w16, w32, d_valid.p, d_fix.p, d_ctr.p);
  }
  clk.tick();
  CorrCounters c;
  try {
    out->valid = (uint8_t*)pinned_get(ctx, count);
    SGX_CUDA(cudaMemcpyAsync(&c, d_ctr.p, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaMemcpyAsync(out->valid, d_valid.p, count, cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    out->n_fixes = c.n_fixes;
    out->fixes = (sgx_fix*)pinned_get(ctx, std::max<uint64_t>(c.n_fixes, 1) * sizeof(sgx_fix));
    if (c.n_fixes)
      SGX_CUDA(cudaMemcpyAsync(out->fixes, d_fix.p, c.n_fixes * sizeof(sgx_fix), cudaMemcpyDeviceToHost, ctx->stream));
    t_all.stop();
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (...) {
    pinned_put(out->valid);
    pinned_put(out->fixes);
    memset(out, 0, sizeof(*out));
    throw;
  }
  out->ms_total = t_all.ms();
  out->ms_kernel = clk.total_ms();
  out->lookups = c.lookups;
  out->fixed_reads = c.fixed_reads;
  out->launches = ctx->launches - launches0;
}

}  // namespace sgx
