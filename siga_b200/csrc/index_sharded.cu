// Bucket-sharded BWT / read-index construction across several GPUs (sgx_index_build_sharded).
//
// Same column-insertion algorithm as index.cu, with the ROWS of the BWT partitioned by the leading k-mer of their
// suffix (BASELINE.json north_star: "suffixes are partitioned by leading k-mer bucket, exchanged with NCCL
// all-to-all over NVLink, and concatenated into the global BWT"):
//
//   * Bucket key of a suffix = its first 8 symbols as a base-5 number ('$' = 0 ends the key).  Keys are contiguous
//     in suffix order, so a shard owning the key range [bound[g], bound[g+1]) owns a contiguous row range of every
//     partial BWT.  The bounds come from a k-mer histogram of a read sample (each shard counts its slice, NCCL
//     all-reduce); they only balance the load -- any bounds give the same BWT.  Exact final rows per shard are
//     counted the same way (slice + all-reduce) to size the arrays.
//   * The packed reads are replicated.  The read-start order (.sai) and the sentinel order are computed sharded too
//     (sharded_read_order below: every shard sorts the reads whose first 8-mer it owns) and then all-gathered.
//   * Column pass j: every shard LF-maps the suffixes it inserted in pass j-1 through ITS rows (local rank of the
//     preceding base c), finds the owner of the new suffix from its key, and partitions the inserts by (destination
//     shard, c).  One small all-gather (NCCL, on a side stream) carries the send counts and each shard's A/C/G/T totals
//     to the hosts; the inserts themselves are written by the scatter kernel of that partition STRAIGHT INTO THE
//     RECEIVERS' MEMORY over NVLink -- every shard owns a receive region that all processes have mapped (cudaIpc), with
//     one worst-case-sized slot per sender -- followed by the sender's arrival flag; the receiver's stream spins on its
//     flags and goes on (section "exchange through peer memory" below).  Where the regions cannot be mapped, all ranks
//     fall back to a send buffer + grouped ncclSend/ncclRecv.  The receiver turns the sender-local rank part into its
//     own local row with a per-(sender, symbol) constant
//         row = part + N + sum_{c'<c} T[c'] + sum_{s'<s} tot[s'][c] - sum_{c'<c} tot[s][c'] - first_row_of(receiver)
//     The send partition is by (destination, c), so every sender's block arrives ordered by (c, sender row) and the
//     receiver's row order -- for one c the new row grows with the old global row, i.e. with (sender, sender row) --
//     is a permutation known from the gathered counts: one placement kernel, no counting pass.  Then the same
//     streaming merge.
//   * Run-length encoding per shard; a run crossing a seam belongs to the shard where it starts (run_length() in
//     index_kernels.cuh); the segments are concatenated on shard 0 (sgx_index_build_sharded) or stay where they are,
//     each with its offset (sgx_index_build_sharded_segments).
//
// The communicator is sgx_comm: one process per GPU, or virtual shards of one context (the mode the single-GPU parity
// tests run: same kernels, the "peers" in the same HBM).
#include <algorithm>
#include <chrono>

#include "comm.cuh"
#include "index_kernels.cuh"
#include "shard_plan.h"

namespace sgx {

static const int kKeySyms = 8;
static const uint32_t kKeySpace = kPlanKeySpace;  // 5^8
static const uint32_t kKeyTop = kPlanKeyTop;      // 5^7
static const int kSendBuckets = kPlanSendBuckets;  // (destination shard, first base of the new suffix)
static const int kMetaN = kPlanMetaN;              // per-shard values gathered every pass: send counts + A/C/G/T totals
static_assert(kPlanMaxShards == kMaxShards, "shard_plan.h and comm.cuh disagree on the shard limit");

static inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct ShardBounds {
  uint32_t b[kMaxShards + 1];
  int world;
  // bounds are non-decreasing; equal neighbours make empty shards, which then own nothing
  __device__ __forceinline__ int owner(uint32_t key) const {
    int g = 0;
#pragma unroll
    for (int i = 1; i < kMaxShards; ++i) g += (i < world && key >= b[i]) ? 1 : 0;
    return g;
  }
};

__device__ __forceinline__ uint32_t suffix_key(const ReadAcc& acc, uint32_t id, uint32_t col, uint32_t l) {
  uint32_t key = 0;
#pragma unroll
  for (int t = 0; t < kKeySyms; ++t) {
    const uint32_t c = col + t;
    key = key * 5u + (c < l ? acc.code(id, c, l) + 1u : 0u);
  }
  return key;
}

// hist[key] += 1 for every suffix (all l+1 of them, the bare '$' included) of the sampled reads
// first, first + step, ... ; the key is rolled from the end of the read.
static __global__ void __launch_bounds__(256) key_hist_kernel(ReadAcc acc, uint64_t n, uint64_t first, uint64_t step,
                                                              unsigned long long* __restrict__ hist) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t id64 = first + t * step;
  if (id64 >= n) return;
  const uint32_t id = (uint32_t)id64;
  const uint32_t l = acc.rv.len[id];
  const int lane = threadIdx.x & 31;
  // one atomic per distinct key among the lanes that are here together: the keys of the shortest suffixes ("$", "c$",
  // ...) are the same in every lane, and a million atomics on one address serialise
  auto add = [&](uint32_t key) {
    const unsigned m = __match_any_sync(__activemask(), key);
    if (lane == __ffs(m) - 1) atomicAdd(&hist[key], (unsigned long long)__popc(m));
  };
  uint32_t key = 0;
  add(0u);
  for (uint32_t col = l; col-- > 0;) {
    key = (acc.code(id, col, l) + 1u) * kKeyTop + key / 5u;
    add(key);
  }
}

// rows[g] += number of suffixes of reads [r0, r1) owned by shard g.  A read has at most 2^16 suffixes here (longer reads
// take the 64-bit counters of the tail loop), so the sixteen per-thread counters are 16-bit fields of four registers.
static __global__ void __launch_bounds__(256) shard_rows_kernel(ReadAcc acc, uint64_t r0, uint64_t r1, ShardBounds bounds,
                                                                unsigned long long* __restrict__ rows) {
  __shared__ unsigned long long s_rows[kMaxShards];
  if (threadIdx.x < kMaxShards) s_rows[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t id64 = r0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int world = bounds.world;
  uint64_t c4[4] = {0, 0, 0, 0};  // counter g = bits [16 (g & 3), 16 (g & 3) + 16) of c4[g >> 2]
  if (id64 < r1) {
    const uint32_t id = (uint32_t)id64;
    const uint32_t l = acc.rv.len[id];
    uint32_t key = 0;
    c4[0] = 1;  // the '$' suffix
    uint32_t since = 1;
    for (uint32_t col = l; col-- > 0;) {
      key = (acc.code(id, col, l) + 1u) * kKeyTop + key / 5u;
      int o = 0;
      for (int i = 1; i < world; ++i) o += key >= bounds.b[i] ? 1 : 0;
      const uint64_t inc = 1ull << (16 * (o & 3));
      const int w = o >> 2;
      c4[0] += w == 0 ? inc : 0ull;
      c4[1] += w == 1 ? inc : 0ull;
      c4[2] += w == 2 ? inc : 0ull;
      c4[3] += w == 3 ? inc : 0ull;
      if (++since == 0xffffu) {  // flush before a field can overflow
#pragma unroll
        for (int g = 0; g < kMaxShards; ++g) {
          const unsigned long long v = (c4[g >> 2] >> (16 * (g & 3))) & 0xffffull;
          if (v) atomicAdd(&s_rows[g], v);
        }
        c4[0] = c4[1] = c4[2] = c4[3] = 0;
        since = 0;
      }
    }
  }
#pragma unroll
  for (int g = 0; g < kMaxShards; ++g) {
    uint32_t v = (uint32_t)((c4[g >> 2] >> (16 * (g & 3))) & 0xffffull);
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_rows[g], (unsigned long long)v);
  }
  __syncthreads();
  if (threadIdx.x < kMaxShards && s_rows[threadIdx.x]) atomicAdd(&rows[threadIdx.x], s_rows[threadIdx.x]);
}

// Column pass, sender side: for every suffix inserted here by the previous pass, the destination shard of the suffix
// one base longer and the sender-local part of its row: (#symbols < c in my rows) + (#c in my rows before mine).
// The key of the new suffix is rolled from the key of the suffix it extends (c * 5^7 + key / 5), which travels with
// the item, so the kernel does not touch the read store.
template <class PosT>
__global__ void __launch_bounds__(256) shard_rank_kernel(const uint16_t* __restrict__ rloc,
                                                         const uint64_t* __restrict__ scanned, uint64_t ntiles, ShardBounds bounds,
                                                         const PosT* __restrict__ NP, const uint8_t* __restrict__ SYM,
                                                         const uint32_t* __restrict__ KEY, uint64_t held,
                                                         uint8_t* __restrict__ tmpDest, PosT* __restrict__ tmpPos) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= held) return;
  const uint32_t sym = (uint32_t)SYM[k];  // 0: the read ended in the previous pass
  if (!sym) {
    tmpDest[k] = 255;
    return;
  }
  const uint64_t P = (uint64_t)NP[k];
  const uint32_t key = sym * kKeyTop + KEY[k] / 5u;
  const uint32_t c = sym - 1;
  // rank of c before the item's row in this shard's rows: scanned tile totals + the tile-local rank the kernel that
  // wrote the row left behind (tile_checkpoints, index_kernels.cuh) -- no access to the BWT
  const uint64_t sc = __ldg(scanned + (uint64_t)c * ntiles + P / kMergeTile);
  tmpDest[k] = (uint8_t)(bounds.owner(key) * 4 + (int)c);  // send bucket: destination-major, then c
  tmpPos[k] = (PosT)(sc + rloc[k]);
}

// One insert on the wire: the sender-local rank part, the read, the bucket key of the new suffix (rolled from the key of
// the suffix it extends, so neither side touches the read store for it) and the 16-base word of the read the walk is in
// (the receiver takes the symbol of the new row out of it; a fresh word is fetched once every 16 passes).
template <class PosT>
struct WireItem;
template <>
struct __align__(16) WireItem<uint32_t> {
  uint32_t pos, id, key, win;
};
template <>
struct __align__(8) WireItem<uint64_t> {
  uint64_t pos;
  uint32_t id, key, win, pad;
};

struct DestDigit {
  const uint8_t* tmpDest;
  __device__ __forceinline__ int operator()(uint64_t k) const { return (int)tmpDest[k]; }  // 255 -> dropped
};
template <class PosT>
struct SendMove {
  const uint32_t* ID;
  const PosT* tmpPos;
  const uint8_t* SYM;
  const uint32_t* KEY;
  const uint32_t* WIN;
  WireItem<PosT>* send;
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t dest) const {
    WireItem<PosT> w;
    w.pos = tmpPos[k];
    w.id = ID[k];
    w.key = (uint32_t)SYM[k] * kKeyTop + KEY[k] / 5u;
    w.win = WIN[k];
    send[dest] = w;
  }
};

// meta[4 d + c] = items for shard d whose new suffix starts with c, meta[kSendBuckets + c] = symbol c in my rows.
//   bucket_start: the partition workspace after its scan (bucket e starts at bucket_start[e * ptiles]); kept = items sent.
static __global__ void shard_meta_kernel(const uint64_t* __restrict__ bucket_start, uint32_t ptiles, int have_items,
                                         const uint64_t* __restrict__ kept, const uint64_t* __restrict__ scanned, uint64_t ntiles,
                                         const uint64_t* __restrict__ sym_total, uint64_t* __restrict__ meta) {
  const int t = threadIdx.x;
  if (t < kSendBuckets) {
    uint64_t c = 0;
    if (have_items) {
      const uint64_t a = bucket_start[(uint64_t)t * ptiles];
      const uint64_t e = t + 1 < kSendBuckets ? bucket_start[(uint64_t)(t + 1) * ptiles] : *kept;
      c = e - a;
    }
    meta[t] = c;
  } else if (t < kSendBuckets + 4) {
    const int c = t - kSendBuckets;
    const uint64_t e = c < 3 ? scanned[(uint64_t)(c + 1) * ntiles] : *sym_total;
    meta[t] = e - scanned[(uint64_t)c * ntiles];
  }
}

// ---- exchange through peer memory -------------------------------------------------------------------------------------
// Every shard owns a region that all processes of the communicator have mapped (comm_peer_reserve): a header of arrival
// flags and, per parity of the pass, one slot per sender sized for the worst case (every read of the set), so a sender
// needs nothing from its peers to know where its block goes.  The push kernel streams the partitioned send buffer --
// ordered by (destination, c) -- into the destinations' slots with coalesced 16-byte stores over NVLink, and the CTA
// that finishes last raises this sender's flag at every destination (release at system scope behind the CTAs' own
// fences).  The receiver's stream holds a one-warp kernel that spins on its flags (acquire) before the placement
// kernel: no NCCL call, no host round trip on the data path.  Two parities: a sender may be one pass ahead of a
// receiver that is still placing the previous pass's block; it cannot be two ahead, because its own next pass needs
// that receiver's block of the pass in between.
struct PushPlan {
  uint8_t* slot[kMaxShards];   // this sender's slot in the region of destination d (current parity)
  uint64_t* flag[kMaxShards];  // this sender's arrival flag at destination d
  int world;
};

// What the peer transport reads of a held item.  Destination bucket and sender-local rank part are computed where they
// are needed -- in the count kernel (bucket only) and in the send kernel -- instead of by a kernel of their own that
// writes them out for the other two to read back (shard_rank_kernel, which the NCCL transport still uses).
template <class PosT>
struct HeldItems {
  const uint32_t* ID;
  const PosT* NP;
  const uint8_t* SYM;
  const uint32_t* KEY;
  const uint32_t* WIN;
  const uint16_t* rloc;
  const uint64_t* scanned;  // scanned tile totals of the shard's rows, [4][ntiles]
  uint64_t ntiles;
  ShardBounds bounds;
  // send bucket (destination-major, then c) of item k, 255 when the read ended in the previous pass; key = rolled key
  __device__ __forceinline__ int bucket(uint64_t k, uint32_t& key) const {
    const uint32_t sym = (uint32_t)SYM[k];
    if (!sym) return 255;
    key = sym * kKeyTop + KEY[k] / 5u;
    int o = 0;
    for (int i = 1; i < bounds.world; ++i) o += key >= bounds.b[i] ? 1 : 0;
    return o * 4 + (int)sym - 1;
  }
  __device__ __forceinline__ PosT rank_part(uint64_t k, int c) const {
    return (PosT)(__ldg(scanned + (uint64_t)c * ntiles + (uint64_t)NP[k] / kMergeTile) + rloc[k]);
  }
};
template <class PosT>
struct HeldDigit {
  HeldItems<PosT> h;
  __device__ __forceinline__ int operator()(uint64_t k) const {
    uint32_t key;
    return h.bucket(k, key);
  }
};

// Scatter half of the send partition, fused with the push.  Same tile walk as part_scatter_kernel (prims.cuh) and the
// same stable (destination, c) order, but the tile's 2048 wire items are first laid out in bucket order in shared
// memory and then written by consecutive threads: every run of one bucket becomes one stretch of coalesced 16-byte
// stores straight into the destination's slot (a warp of the plain scatter spreads its 32 items over up to 16 buckets,
// which is tolerable in local HBM and wasteful over NVLink).  No send buffer, no second pass over the items.
template <class PosT>
__global__ void __launch_bounds__(kPartThreads) shard_send_kernel(HeldItems<PosT> held, uint64_t n,
                                                                  const uint64_t* __restrict__ tileoff, uint32_t ntiles,
                                                                  PushPlan pp, uint64_t epoch, unsigned* __restrict__ done) {
  extern __shared__ __align__(16) uint8_t send_smem[];
  WireItem<PosT>* stage = (WireItem<PosT>*)send_smem;  // kPartTile items
  __shared__ uint32_t wc[kPartThreads / 32][kSendBuckets];
  __shared__ uint32_t boff[kSendBuckets + 1];
  __shared__ uint64_t gbase[kSendBuckets];  // slot-relative position of the tile's first item of bucket e
  __shared__ uint8_t bkt[kPartTile];
  __shared__ WireItem<PosT>* s_slot[kMaxShards];
  for (int i = threadIdx.x; i < (kPartThreads / 32) * kSendBuckets; i += kPartThreads) (&wc[0][0])[i] = 0;
  if ((int)threadIdx.x < pp.world) s_slot[threadIdx.x] = (WireItem<PosT>*)pp.slot[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t wbase = (uint64_t)blockIdx.x * kPartTile + (uint64_t)w * (kPartRounds * 32);
  int dg[kPartRounds];
  uint32_t lr[kPartRounds], keys[kPartRounds];
#pragma unroll
  for (int r = 0; r < kPartRounds; ++r) {
    const uint64_t k = wbase + r * 32 + lane;
    int d = kSendBuckets;
    keys[r] = 0;
    if (k < n) d = held.bucket(k, keys[r]);  // 255: the read has ended, nothing to send
    dg[r] = d;
    lr[r] = 0;
    const unsigned act = __ballot_sync(0xffffffffu, d < kSendBuckets);
    if (d < kSendBuckets) {
      const unsigned m = __match_any_sync(act, d);
      const uint32_t before = wc[w][d];
      lr[r] = before + __popc(m & ((1u << lane) - 1u));
      __syncwarp(m);
      if (lane == __ffs(m) - 1) wc[w][d] = before + __popc(m);
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < kSendBuckets) {
    const int e = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int q = 0; q < kPartThreads / 32; ++q) {
      const uint32_t t = wc[q][e];
      wc[q][e] = run;
      run += t;
    }
    boff[e + 1] = run;  // bucket total of the tile, scanned below
    gbase[e] = n ? tileoff[(uint64_t)e * ntiles + blockIdx.x] - tileoff[(uint64_t)(e & ~3) * ntiles] : 0;
  }
  __syncthreads();
  if (w == 0) {  // exclusive scan of the 64 bucket totals (two per lane)
    const uint32_t a = boff[2 * lane + 1], b2 = boff[2 * lane + 2];
    uint32_t x = a + b2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    __syncwarp();
    boff[2 * lane + 1] = x - b2;
    boff[2 * lane + 2] = x;
    if (lane == 0) boff[0] = 0;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kPartRounds; ++r) {
    const int d = dg[r];
    if (d < kSendBuckets) {
      const uint64_t k = wbase + r * 32 + lane;
      const uint32_t t = boff[d] + wc[w][d] + lr[r];
      WireItem<PosT> it;
      it.pos = held.rank_part(k, d & 3);
      it.id = held.ID[k];
      it.key = keys[r];
      it.win = held.WIN[k];
      stage[t] = it;
      bkt[t] = (uint8_t)d;
    }
  }
  __syncthreads();
  const uint32_t total = boff[kSendBuckets];
  for (uint32_t t = threadIdx.x; t < total; t += kPartThreads) {
    const int e = bkt[t];
    s_slot[e >> 2][gbase[e] + (t - boff[e])] = stage[t];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(done, 1u);
    if (t == gridDim.x - 1) {
      *done = 0;  // the next pass on this stream starts from zero
      __threadfence_system();
      for (int d = 0; d < pp.world; ++d) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pp.flag[d]), "l"(epoch) : "memory");
    }
  }
}

// One warp: lane s waits until sender s has raised its flag to this pass's epoch.  A peer that never arrives (its
// process died) must not hang the device: after kPushTimeoutNs the kernel traps and the build fails with a CUDA error.
static const unsigned long long kPushTimeoutNs = 30ull * 1000000000ull;
static __global__ void shard_wait_kernel(const uint64_t* __restrict__ flags, int world, uint64_t epoch) {
  if ((int)threadIdx.x < world) {
    const uint64_t* f = flags + threadIdx.x;
    unsigned long long t0 = 0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      uint64_t v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= epoch) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > kPushTimeoutNs) __trap();
      __nanosleep(200);
    }
  }
  __syncwarp();
  __threadfence_system();
}

// Receiver side.  Every sender's block arrives ordered by (c, sender row), so the inserts of this shard in row order
// are, for c = A..T, the (sender 0, c) run, the (sender 1, c) run, ...: a permutation known from the gathered counts.
// Entry e = c * world + s of the plan: first output slot, first receive-buffer slot, and the constant that turns the
// sender-local rank part into a local row.

// One thread per insert, in output (row) order: find its run, fetch the wire item, write id / local row / key / the
// symbol its row holds (the base before the suffix, '$' when the read starts here) / the word that symbol came from.
template <class PosT>
__global__ void __launch_bounds__(256) shard_place_kernel(ReadAcc acc, const WireItem<PosT>* __restrict__ recv, uint64_t n,
                                                          uint32_t j, uint32_t uniform_len, RecvPlan plan,
                                                          uint32_t* __restrict__ ID2, PosT* __restrict__ NP2,
                                                          uint8_t* __restrict__ SYM2, uint32_t* __restrict__ KEY2,
                                                          uint32_t* __restrict__ WIN2) {
  // the plan (at most 64 runs) in shared memory: a search with thread-dependent indices into kernel parameters would go
  // through a local-memory copy of the whole structure in every thread
  __shared__ uint64_t s_out0[kSendBuckets + 1], s_src0[kSendBuckets];
  __shared__ long long s_K[kSendBuckets];
  for (int e = threadIdx.x; e <= plan.entries; e += blockDim.x) {
    s_out0[e] = plan.out0[e];
    if (e < plan.entries) {
      s_src0[e] = plan.src0[e];
      s_K[e] = plan.K[e];
    }
  }
  __syncthreads();
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  int lo = 0, hi = plan.entries;  // last entry with out0 <= d
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_out0[mid] <= d) lo = mid;
    else hi = mid;
  }
  const WireItem<PosT> w = recv[s_src0[lo] + (d - s_out0[lo])];
  const uint32_t id = w.id;
  const uint32_t l = uniform_len ? uniform_len : acc.rv.len[id];
  ID2[d] = id;
  NP2[d] = (PosT)((long long)w.pos + s_K[lo]);
  KEY2[d] = w.key;
  if (l == j) {  // the read starts here: its row holds '$' and nothing follows
    SYM2[d] = 0;
    return;
  }
  // stored position of column l - j - 1 (the walk goes down through a forward read, up through a reversed one); its
  // word travelled with the item unless the walk has just entered a new one
  const uint32_t p = acc.reversed ? j : l - j - 1;
  const bool fresh = j == 1 || (acc.reversed ? (p & 15u) == 0u : (p & 15u) == 15u);
  uint32_t word = w.win;
  if (fresh) {
    const uint64_t wo = uniform_len ? (uint64_t)id * ((uniform_len + 15u) >> 4) : acc.rv.woff[id];  // reads start on words
    word = acc.rv.words[wo + (p >> 4)];
  }
  SYM2[d] = (uint8_t)(((word >> (2 * (p & 15u))) & 3u) + 1u);
  WIN2[d] = word;
}

template <class PosT>
struct Shard {
  int g = 0;
  uint64_t rows_final = 0;
  DevBuf<uint8_t> B0, B1, SYM, tmpDest;
  DevBuf<uint32_t> ID0, ID1, KEY0, KEY1, WIN0, WIN1;
  DevBuf<PosT> NP0, NP1, tmpPos;
  DevBuf<WireItem<PosT> > send, recv;  // NCCL transport only (the peer transport scatters into the receivers' regions)
  DevBuf<uint16_t> RL;  // tile-local rank of every held item's row
  DevBuf<uint64_t> tiletot, scan_ws, part_ws, tile_k0, meta, kept, sym_total;
  DevBuf<unsigned> push_done;
  uint8_t* B = nullptr;
  uint8_t* Bn = nullptr;
  uint32_t* ID = nullptr;
  uint32_t* ID2 = nullptr;
  uint32_t* KEY = nullptr;
  uint32_t* KEY2 = nullptr;
  uint32_t* WIN = nullptr;
  uint32_t* WIN2 = nullptr;
  PosT* NP = nullptr;
  PosT* NP2 = nullptr;
  uint64_t n_cur = 0, nt = 1, held = 0;
  RleSegment seg;
  DevBuf<uint8_t> rle;
  uint64_t rle_bytes = 0;

  void alloc(sgx_ctx* ctx, uint64_t N, bool peer) {
    const uint64_t padded = bwt_array_bytes(rows_final);
    B0.alloc(padded);
    B1.alloc(padded);
    SGX_CUDA(cudaMemsetAsync(B0.p, 0, padded, ctx->stream));
    SGX_CUDA(cudaMemsetAsync(B1.p, 0, padded, ctx->stream));
    SYM.alloc(N);
    if (!peer) tmpDest.alloc(N);
    ID0.alloc(N);
    ID1.alloc(N);
    KEY0.alloc(N);
    KEY1.alloc(N);
    WIN0.alloc(N);
    WIN1.alloc(N);
    SGX_CUDA(cudaMemsetAsync(KEY0.p, 0, N * sizeof(uint32_t), ctx->stream));  // the sentinel rows: key of "$"
    SGX_CUDA(cudaMemsetAsync(WIN0.p, 0, N * sizeof(uint32_t), ctx->stream));  // pass 1 fetches its own words
    if (!peer) {
      send.alloc(N);
      recv.alloc(N);
    }
    NP0.alloc(N);
    NP1.alloc(N);
    if (!peer) tmpPos.alloc(N);
    const uint64_t nt_max = rows_final / kMergeTile + 2;
    RL.alloc(N);
    tiletot.alloc(4 * nt_max);
    scan_ws.alloc(scan_ws_elems(4 * nt_max));
    part_ws.alloc(part_ws_elems(N, kSendBuckets));
    tile_k0.alloc(rows_final / kMergeTile + 4);
    meta.alloc(kMetaN);
    kept.alloc(1);
    sym_total.alloc(1);
    push_done.alloc(1);
    SGX_CUDA(cudaMemsetAsync(push_done.p, 0, sizeof(unsigned), ctx->stream));
    B = B0.p;
    Bn = B1.p;
    ID = ID0.p;
    ID2 = ID1.p;
    KEY = KEY0.p;
    KEY2 = KEY1.p;
    WIN = WIN0.p;
    WIN2 = WIN1.p;
    NP = NP0.p;
    NP2 = NP1.p;
  }
};

template <class PosT>
static void sharded_build(sgx_ctx* ctx, sgx_comm* cm, const sgx_reads* reads, const ReadAcc& acc, const uint32_t* d_sord,
                          const ShardBounds& bounds, const std::vector<uint64_t>& rows_final, sgx_index_out* out,
                          KernelClock& k_merge, EventTimer& t_bcr, EventTimer& t_emit, bool segments, uint8_t** rle_host,
                          uint64_t* rle_total, uint64_t* seg_offset, uint64_t* seg_len) {
  const uint64_t N = reads->n;
  const uint32_t lmax = reads->max_len;
  const uint32_t uniform_len = reads->min_len == lmax ? lmax : 0;
  const int W = cm->world;
  const int nl = cm->n_local();
  // Transport of the inserts: peer memory when every process could map every region, NCCL send/recv otherwise.  A slot
  // holds the worst case, all N reads going from one sender to one receiver (the first passes, whose keys are mostly
  // '$', do send everything to a few shards).
  const uint64_t slot_cap = N;
  const size_t slot_bytes = (size_t)slot_cap * sizeof(WireItem<PosT>);
  const bool peer = comm_peer_reserve(cm, kPeerHeaderBytes + 2 * (size_t)W * slot_bytes);
  std::vector<Shard<PosT> > sh(nl);
  for (int li = 0; li < nl; ++li) {
    sh[li].g = cm->shard_of(li);
    sh[li].rows_final = rows_final[sh[li].g];
    sh[li].alloc(ctx, N, peer);
  }
  std::vector<uint64_t> n_cur_all(W, 0);
  n_cur_all[0] = N;

  t_bcr.start();
  // pass 0: the N sentinel rows, all on shard 0
  for (int li = 0; li < nl; ++li) {
    Shard<PosT>& s = sh[li];
    if (s.g == 0) {
      SGX_LAUNCH(ctx, (bcr_init_kernel<PosT>), grid_for((N + 1) / 2, 256), 256, 0, acc, d_sord, N, s.B, s.ID, s.NP, s.SYM.p);
      s.n_cur = N;
      s.held = N;
    }
    s.nt = std::max<uint64_t>(1, (s.n_cur + kMergeTile - 1) / kMergeTile);
    SGX_LAUNCH(ctx, tile_count_kernel, (unsigned)s.nt, 256, 0, (const uint8_t*)s.B, s.n_cur, s.nt, (uint2*)nullptr, s.tiletot.p, s.RL.p);
    scan_exclusive(ctx, s.tiletot.p, 4 * s.nt, s.scan_ws.p, s.sym_total.p);
  }

  // SGX_SHARD_TRACE=1: device time per phase of the column passes, printed to stderr (diagnostics only)
  const char* tr = getenv("SGX_SHARD_TRACE");
  const bool trace = tr && tr[0] == '1';
  KernelClock k_send(ctx->stream), k_gather(ctx->stream), k_xchg(ctx->stream), k_recv(ctx->stream);
  double host_wait_ms = 0;
#define SGX_TRACE_TICK(clock) \
  if (trace) (clock).tick()
  std::vector<uint64_t> M((size_t)W * kMetaN), cntm((size_t)W * W), soff((size_t)W * W), roff((size_t)W * W);
  std::vector<const uint64_t*> meta_ptr(nl);
  std::vector<const void*> send_ptr(nl);
  std::vector<void*> recv_ptr(nl);
  const size_t send_smem = (size_t)kPartTile * sizeof(WireItem<PosT>);
  if (peer)
    SGX_CUDA(cudaFuncSetAttribute(shard_send_kernel<PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)send_smem));
  for (uint32_t j = 1; j <= lmax; ++j) {
    // ---- sender side: ranks + destinations, bucket sizes; the sizes travel (all-gather, copy to the host) while
    //      the scatter into the send buffer and the push into the peers' regions run
    SGX_TRACE_TICK(k_send);
    for (int li = 0; li < nl; ++li) {
      Shard<PosT>& s = sh[li];
      uint32_t ntiles = 1;
      if (s.held && peer) {
        HeldDigit<PosT> hd{HeldItems<PosT>{s.ID, s.NP, s.SYM.p, s.KEY, s.WIN, s.RL.p, s.tiletot.p, s.nt, bounds}};
        partition_count<kSendBuckets, HeldDigit<PosT> >(ctx, hd, s.held, s.part_ws.p, s.kept.p);
        ntiles = part_tiles(s.held);
      } else if (s.held) {
        SGX_LAUNCH(ctx, (shard_rank_kernel<PosT>), grid_for(s.held, 256), 256, 0, (const uint16_t*)s.RL.p,
                   (const uint64_t*)s.tiletot.p, s.nt, bounds, (const PosT*)s.NP, (const uint8_t*)s.SYM.p,
                   (const uint32_t*)s.KEY, s.held, s.tmpDest.p, s.tmpPos.p);
        DestDigit dg{s.tmpDest.p};
        partition_count<kSendBuckets, DestDigit>(ctx, dg, s.held, s.part_ws.p, s.kept.p);
        ntiles = part_tiles(s.held);
      }
      SGX_LAUNCH(ctx, shard_meta_kernel, 1, 96, 0, (const uint64_t*)s.part_ws.p, ntiles, s.held ? 1 : 0,
                 (const uint64_t*)s.kept.p, (const uint64_t*)s.tiletot.p, s.nt, (const uint64_t*)s.sym_total.p, s.meta.p);
      meta_ptr[li] = s.meta.p;
    }
    SGX_TRACE_TICK(k_send);
    SGX_TRACE_TICK(k_gather);
    comm_allgather_u64_begin(cm, meta_ptr.data(), kMetaN);
    const uint64_t epoch = peer ? comm_next_epoch(cm) : 0;
    const size_t parity_off = kPeerHeaderBytes + (size_t)(epoch & 1) * W * slot_bytes;
    for (int li = 0; li < nl; ++li) {
      Shard<PosT>& s = sh[li];
      if (!peer) {
        if (s.held) {
          DestDigit dg{s.tmpDest.p};
          SendMove<PosT> mv{s.ID, s.tmpPos.p, s.SYM.p, s.KEY, s.WIN, s.send.p};
          partition_scatter<kSendBuckets, DestDigit, SendMove<PosT> >(ctx, dg, mv, s.held, s.part_ws.p);
        }
        continue;
      }
      // peer transport: scatter straight into the destinations' slots; a shard with nothing to send still raises its flags
      PushPlan pp;
      memset(&pp, 0, sizeof(pp));
      pp.world = W;
      for (int d = 0; d < W; ++d) {
        pp.slot[d] = cm->peer_addr[d] + parity_off + (size_t)s.g * slot_bytes;
        pp.flag[d] = comm_peer_flag(cm, d, s.g);
      }
      SGX_LAUNCH(ctx, (shard_send_kernel<PosT>), part_tiles(s.held ? s.held : 1), kPartThreads, send_smem,
                 (HeldItems<PosT>{s.ID, s.NP, s.SYM.p, s.KEY, s.WIN, s.RL.p, s.tiletot.p, s.nt, bounds}), s.held,
                 (const uint64_t*)s.part_ws.p, part_tiles(s.held ? s.held : 1), pp, epoch, s.push_done.p);
    }
    SGX_TRACE_TICK(k_gather);
    const double hw0 = trace ? now_ms() : 0;
    comm_allgather_u64_end(cm, kMetaN, M.data());
    if (trace) host_wait_ms += now_ms() - hw0;
    std::vector<uint64_t> n_in(W, 0), n_new(W, 0), first_row(W + 1, 0);
    const uint64_t moved = shard_pass_counts(M.data(), W, n_cur_all.data(), cntm.data(), soff.data(), roff.data(), n_in.data(),
                                             n_new.data(), first_row.data());
    if (moved == 0) break;  // every read has ended
    SGX_TRACE_TICK(k_xchg);
    if (peer) {
      // the receive "buffer" of shard d is the current parity of its region: sender s's block starts at slot s
      for (int d = 0; d < W; ++d)
        for (int q = 0; q < W; ++q) roff[d * W + q] = (uint64_t)q * slot_cap;
      for (int li = 0; li < nl; ++li)
        if (n_in[sh[li].g])
          SGX_LAUNCH(ctx, shard_wait_kernel, 1, 32, 0, (const uint64_t*)cm->peer_mine[li], W, epoch);
    } else {
      for (int li = 0; li < nl; ++li) {
        send_ptr[li] = sh[li].send.p;
        recv_ptr[li] = sh[li].recv.p;
      }
      comm_alltoallv(cm, send_ptr.data(), recv_ptr.data(), sizeof(WireItem<PosT>), cntm.data(), soff.data(), roff.data());
    }
    SGX_TRACE_TICK(k_xchg);
    SGX_TRACE_TICK(k_recv);

    // ---- receiver side
    for (int li = 0; li < nl; ++li) {
      Shard<PosT>& s = sh[li];
      const int g = s.g;
      const uint64_t held_new = n_in[g];
      SGX_REQUIRE(n_new[g] <= s.rows_final, SGX_ECUDA, "sharded index: a shard received more rows than were counted for it");
      if (held_new) {
        RecvPlan plan;
        const uint64_t out_run = shard_recv_plan(M.data(), W, g, N, roff.data(), first_row.data(), &plan);
        SGX_REQUIRE(out_run == held_new, SGX_ECUDA, "sharded index: receive plan does not cover the received inserts");
        const WireItem<PosT>* rbuf = peer ? (const WireItem<PosT>*)(cm->peer_mine[li] + parity_off) : (const WireItem<PosT>*)s.recv.p;
        SGX_LAUNCH(ctx, (shard_place_kernel<PosT>), grid_for(held_new, 256), 256, 0, acc, rbuf, held_new, j, uniform_len, plan,
                   s.ID2, s.NP2, s.SYM.p, s.KEY2, s.WIN2);
        const uint64_t rows_new = s.n_cur + held_new;
        const uint64_t ntiles = (rows_new + kMergeTile - 1) / kMergeTile;
        SGX_LAUNCH(ctx, (bcr_tile_bounds_kernel<PosT>), grid_for(held_new + 1, 256), 256, 0, (const PosT*)s.NP2, held_new,
                   ntiles, s.tile_k0.p);
        k_merge.tick();
        launch_merge<PosT>(ctx, (const uint8_t*)s.B, s.n_cur, (const PosT*)s.NP2, (const uint8_t*)s.SYM.p, held_new,
                           (const uint64_t*)s.tile_k0.p, s.Bn, rows_new, ntiles, s.tiletot.p, s.RL.p);
        k_merge.tick();
        out->merge_launches++;
        out->merge_bytes += (s.n_cur + rows_new) / 2 + held_new * (sizeof(PosT) + 1);
        scan_exclusive(ctx, s.tiletot.p, 4 * ntiles, s.scan_ws.p, s.sym_total.p);
        std::swap(s.B, s.Bn);
        std::swap(s.ID, s.ID2);
        std::swap(s.NP, s.NP2);
        std::swap(s.KEY, s.KEY2);
        std::swap(s.WIN, s.WIN2);
        s.n_cur = rows_new;
        s.nt = ntiles;
      }
      // without inserts the rows and their scanned counts stay as they are
      s.held = held_new;
    }
    SGX_TRACE_TICK(k_recv);
    for (int d = 0; d < W; ++d) n_cur_all[d] = n_new[d];
  }
  for (int li = 0; li < nl; ++li)
    SGX_REQUIRE(sh[li].n_cur == sh[li].rows_final, SGX_ECUDA, "sharded index: column passes did not fill the shard's rows");
  t_bcr.stop();
  if (trace) {
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    // an unmatched tick is left when the loop ended on `moved == 0`; total_ms() only sums complete pairs
    fprintf(stderr, "[sgx shard %d/%d] %s transport, passes: sender %.2f ms, scatter+push %.2f ms, %s %.2f ms, receiver %.2f ms (merge %.2f ms); host waited %.2f ms for counts\n",
            cm->shard_of(0), W, peer ? "peer-memory" : "NCCL", k_send.total_ms(), k_gather.total_ms(), peer ? "flag wait" : "all-to-all",
            k_xchg.total_ms(), k_recv.total_ms(), k_merge.total_ms(), host_wait_ms);
  }
#undef SGX_TRACE_TICK

  // ---- run-length encoding with seam handling, segments concatenated on shard 0
  t_emit.start();
  const int kSeamN = 5;
  std::vector<uint64_t> seam((size_t)W * kSeamN);
  {
    std::vector<DevBuf<uint64_t> > d_seam(nl);
    std::vector<const uint64_t*> ptr(nl);
    for (int li = 0; li < nl; ++li) {
      Shard<PosT>& s = sh[li];
      rle_heads(ctx, s.B, s.n_cur, &s.seg);
      const uint64_t v[kSeamN] = {s.seg.n_rows, s.seg.first_sym, s.seg.first_len, s.seg.last_sym, s.seg.last_len};
      d_seam[li].alloc(kSeamN);
      SGX_CUDA(cudaMemcpyAsync(d_seam[li].p, v, sizeof(v), cudaMemcpyHostToDevice, ctx->stream));
      SGX_CUDA(cudaStreamSynchronize(ctx->stream));  // v is a stack array
      ptr[li] = d_seam[li].p;
    }
    comm_allgather_u64(cm, ptr.data(), kSeamN, seam.data());
  }
  std::vector<uint64_t> seg_bytes(W, 0);
  {
    std::vector<DevBuf<uint64_t> > d_b(nl);
    std::vector<const uint64_t*> ptr(nl);
    for (int li = 0; li < nl; ++li) {
      Shard<PosT>& s = sh[li];
      const int g = s.g;
      int skip_first = 0;
      uint64_t extend = 0;
      shard_seam_rule(seam.data(), W, g, s.seg.n_runs, &skip_first, &extend);
      rle_emit(ctx, s.B, s.seg, skip_first, extend, &s.rle, &s.rle_bytes);
      d_b[li].alloc(1);
      SGX_CUDA(cudaMemcpyAsync(d_b[li].p, &s.rle_bytes, sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
      SGX_CUDA(cudaStreamSynchronize(ctx->stream));
      ptr[li] = d_b[li].p;
    }
    comm_allgather_u64(cm, ptr.data(), 1, seg_bytes.data());
  }
  uint64_t total = 0;
  for (int g = 0; g < W; ++g) total += seg_bytes[g];
  *rle_total = total;
  *rle_host = nullptr;
  *seg_offset = 0;
  *seg_len = total;
  if (segments && !cm->local) {
    // every rank keeps its own segment: nothing is gathered, the copies to the hosts run side by side
    const int g = cm->rank;
    for (int q = 0; q < g; ++q) *seg_offset += seg_bytes[q];
    *seg_len = seg_bytes[g];
    *rle_host = (uint8_t*)pinned_get(ctx, seg_bytes[g] ? seg_bytes[g] : 1);
    if (seg_bytes[g]) SGX_CUDA(cudaMemcpyAsync(*rle_host, sh[0].rle.p, seg_bytes[g], cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  } else {
    const bool root_here = cm->local || cm->rank == 0;
    DevBuf<uint8_t> d_all;
    if (root_here) d_all.alloc(total ? total : 1);
    std::vector<const uint8_t*> ptr(nl);
    for (int li = 0; li < nl; ++li) ptr[li] = sh[li].rle.p;
    comm_gatherv_bytes(cm, ptr.data(), seg_bytes.data(), 0, d_all.p);
    if (root_here) {
      *rle_host = (uint8_t*)pinned_get(ctx, total ? total : 1);
      SGX_CUDA(cudaMemcpyAsync(*rle_host, d_all.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      *seg_len = 0;
    }
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  t_emit.stop();
}

// ---- sharded read ordering -----------------------------------------------------------------------------------------
// The order of the read-start suffixes (.sai) and of the sentinel rows, with the N reads split over the shards by the
// same leading-k-mer bounds as the BWT rows: a shard sorts the reads whose first 8-mer it owns (identical reads share a
// shard, shard order = lexicographic order), dense ranks get a per-shard offset from an all-gather of the distinct
// counts, and the prefix-doubling rounds of ordering model B (ties decided by the following reads) sort only the
// shard's own (rank, rank of the read h further) pairs; the full rank array is rebuilt on every process by an
// all-gather of (read id, rank) after each round.

struct OwnerDigit {
  ReadAcc acc;
  ShardBounds bounds;
  int g;
  __device__ __forceinline__ int operator()(uint64_t k) const {
    const uint32_t id = (uint32_t)k;
    return bounds.owner(suffix_key(acc, id, 0, acc.rv.len[id])) == g ? 0 : 255;
  }
};
struct IotaMove {
  uint32_t* out;
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t dest) const { out[dest] = (uint32_t)k; }
};
// out[k] = offset + (number of heads in [0, k]) given the exclusive scan of the head flags
static __global__ void local_rank_kernel(const uint64_t* __restrict__ excl, uint64_t n, uint64_t total, uint64_t offset,
                                         uint32_t* __restrict__ out) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = (uint32_t)(offset + (k + 1 < n ? excl[k + 1] : total));
}
static __global__ void scatter_rank_kernel(const uint32_t* __restrict__ ids, const uint32_t* __restrict__ ranks, uint64_t n,
                                           uint32_t* __restrict__ rank) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) rank[ids[k]] = ranks[k];
}
static __global__ void doubling_keys_owned_kernel(const uint32_t* __restrict__ rank, uint64_t n_all, const uint32_t* __restrict__ ids,
                                                  uint64_t n, uint64_t h, int bits, uint64_t* __restrict__ keys,
                                                  uint32_t* __restrict__ vals) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint64_t i = ids[k];
  const uint64_t r2 = i + h < n_all ? rank[i + h] : 0;  // end of text is smallest
  keys[k] = ((uint64_t)rank[i] << bits) | r2;
  vals[k] = (uint32_t)i;
}

struct OrderShard {
  int g = 0;
  uint64_t n = 0;
  DevBuf<uint32_t> ids, ids2, lrank;
  DevBuf<uint64_t> keys, keys2, heads, scan_ws, part_ws, total;
};

static void sharded_read_order(sgx_ctx* ctx, sgx_comm* cm, const sgx_reads* reads, const ReadAcc& acc, int order,
                               const ShardBounds& bounds, uint32_t* d_sai, uint32_t* d_sord) {
  const uint64_t N = reads->n;
  const int W = cm->world, nl = cm->n_local();
  // SGX_SHARD_TRACE=1: wall time per phase (the stream is drained at every mark; diagnostics only)
  const char* tr = getenv("SGX_SHARD_TRACE");
  const bool trace = tr && tr[0] == '1';
  double t_mark = trace ? now_ms() : 0, t_select = 0, t_content = 0, t_publish = 0, t_double = 0;
  int rounds = 0;
  auto mark = [&](double& acc_ms) {
    if (!trace) return;
    cudaStreamSynchronize(ctx->stream);
    const double t = now_ms();
    acc_ms += t - t_mark;
    t_mark = t;
  };
  std::vector<OrderShard> sh(nl);
  std::vector<uint64_t> cnt(W, 0), distinct(W, 0);
  std::vector<const uint64_t*> u64_ptr(nl);
  std::vector<const void*> blk_ptr(nl);
  DevBuf<uint64_t> sel_ws(part_ws_elems(N, 1));
  // ---- the reads every shard owns (in read-index order: the sort below is stable, model A ties by read index)
  for (int li = 0; li < nl; ++li) {
    OrderShard& s = sh[li];
    s.g = cm->shard_of(li);
    s.total.alloc(1);
    OwnerDigit dg{acc, bounds, s.g};
    partition_count<1, OwnerDigit>(ctx, dg, N, sel_ws.p, s.total.p);
    SGX_CUDA(cudaMemcpyAsync(&s.n, s.total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint64_t cap = s.n ? s.n : 1;
    s.ids.alloc(cap);
    s.ids2.alloc(cap);
    s.lrank.alloc(cap);
    s.keys.alloc(cap);
    s.keys2.alloc(cap);
    s.heads.alloc(cap + 1);
    s.scan_ws.alloc(scan_ws_elems(cap + 1));
    s.part_ws.alloc(part_ws_elems(cap, 256));
    IotaMove mv{s.ids.p};
    partition_scatter<1, OwnerDigit, IotaMove>(ctx, dg, mv, N, sel_ws.p);
    u64_ptr[li] = s.total.p;
  }
  comm_allgather_u64(cm, u64_ptr.data(), 1, cnt.data());
  uint64_t sum = 0;
  for (int g = 0; g < W; ++g) sum += cnt[g];
  SGX_REQUIRE(sum == N, SGX_ECUDA, "sharded index: the shards' read sets do not add up to the read count");
  mark(t_select);
  // ---- content order + dense content ranks
  for (int li = 0; li < nl; ++li) {
    OrderShard& s = sh[li];
    index_content_sort(ctx, reads, acc, s.ids.p, s.ids2.p, s.keys.p, s.keys2.p, s.n, s.part_ws.p);
    if (s.n) SGX_LAUNCH(ctx, read_heads_kernel, grid_for(s.n, 256), 256, 0, acc.rv, (const uint32_t*)s.ids.p, s.n, s.heads.p);
    scan_exclusive(ctx, s.heads.p, s.n, s.scan_ws.p, s.total.p);
  }
  mark(t_content);
  DevBuf<uint32_t> rank, ranks_by_pos;
  uint32_t* order_ids = d_sai;  // concatenation of the shards' sorted ids = the global order
  auto publish = [&](bool need_ranks) {
    // offsets of the dense ranks from the distinct counts, then (ids, ranks) of all shards to every process
    comm_allgather_u64(cm, u64_ptr.data(), 1, distinct.data());
    uint64_t off = 0, total = 0;
    for (int g = 0; g < W; ++g) total += distinct[g];
    for (int li = 0; li < nl; ++li) {
      OrderShard& s = sh[li];
      off = 0;
      for (int g = 0; g < s.g; ++g) off += distinct[g];
      if (s.n && need_ranks)
        SGX_LAUNCH(ctx, local_rank_kernel, grid_for(s.n, 256), 256, 0, (const uint64_t*)s.heads.p, s.n, distinct[s.g], off,
                   s.lrank.p);
      blk_ptr[li] = s.ids.p;
    }
    comm_allgatherv(cm, blk_ptr.data(), cnt.data(), sizeof(uint32_t), order_ids);
    if (need_ranks) {
      for (int li = 0; li < nl; ++li) blk_ptr[li] = sh[li].lrank.p;
      comm_allgatherv(cm, blk_ptr.data(), cnt.data(), sizeof(uint32_t), ranks_by_pos.p);
      SGX_LAUNCH(ctx, scatter_rank_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)order_ids, (const uint32_t*)ranks_by_pos.p, N,
                 rank.p);
    }
    return total;
  };
  if (order == SGX_ORDER_SAIS) {
    publish(false);
    SGX_LAUNCH(ctx, iota_kernel, grid_for(N, 256), 256, 0, d_sord, N);
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    return;
  }
  rank.alloc(N);
  ranks_by_pos.alloc(N);
  uint64_t n_distinct = publish(true);
  mark(t_publish);
  int bits = 1;
  while (((uint64_t)1 << bits) <= N) ++bits;  // ranks are in [0, N]
  for (uint64_t h = 1; n_distinct < N; h <<= 1) {
    for (int li = 0; li < nl; ++li) {
      OrderShard& s = sh[li];
      if (s.n) {
        SGX_LAUNCH(ctx, doubling_keys_owned_kernel, grid_for(s.n, 256), 256, 0, (const uint32_t*)rank.p, N,
                   (const uint32_t*)s.ids.p, s.n, h, bits, s.keys.p, s.ids2.p);
        radix_sort_pairs(ctx, s.keys.p, s.ids2.p, s.keys2.p, s.ids.p, s.n, 2 * bits, s.part_ws.p);
        // sorted ids are in ids2 (keys in keys); make them the shard's current order
        SGX_CUDA(cudaMemcpyAsync(s.ids.p, s.ids2.p, s.n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        SGX_LAUNCH(ctx, key_heads_kernel, grid_for(s.n, 256), 256, 0, (const uint64_t*)s.keys.p, s.n, s.heads.p);
      }
      scan_exclusive(ctx, s.heads.p, s.n, s.scan_ws.p, s.total.p);
    }
    mark(t_double);
    ++rounds;
    n_distinct = publish(true);
    mark(t_publish);
    SGX_REQUIRE(h < ((uint64_t)1 << 40), SGX_ECUDA, "index: read-rank doubling did not converge");
  }
  uint64_t* k0 = sh[0].total.p;  // scratch
  SGX_LAUNCH(ctx, find_zero_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)d_sai, N, k0);
  SGX_LAUNCH(ctx, sentinel_order_kernel, grid_for(N, 256), 256, 0, (const uint32_t*)d_sai, N, (const uint64_t*)k0, d_sord);
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (trace) {
    double t_final = 0;
    mark(t_final);
    fprintf(stderr, "[sgx shard %d/%d] read order: select own reads %.2f ms, content sort %.2f ms, %d doubling round(s) %.2f ms, "
            "publish (all-gather ids + ranks) %.2f ms, sentinel order %.2f ms\n", cm->shard_of(0), W, t_select, t_content, rounds,
            t_double, t_publish, t_final);
  }
}

void index_build_sharded(sgx_ctx* ctx, sgx_comm* cm, const sgx_reads* reads, int order, bool reversed, bool segments,
                         sgx_index_out* out) {
  memset(out, 0, sizeof(*out));
  const uint64_t N = reads->n;
  const uint64_t S = reads->bases + N;
  const int W = cm->world;
  const uint64_t launches0 = ctx->launches;
  EventTimer t_all(ctx->stream), t_sort(ctx->stream), t_bcr(ctx->stream), t_emit(ctx->stream);
  KernelClock k_merge(ctx->stream);
  ReadAcc acc;
  acc.rv = reads->view();
  acc.reversed = reversed ? 1 : 0;

  t_all.start();
  t_sort.start();
  const double t_host0 = now_ms();
  DevBuf<uint32_t> d_sai(N), d_sord(N);
  // ---- bucket bounds from the k-mer histogram of a read sample; exact rows per shard
  const int nslices = cm->local ? 1 : W;
  const int slice = cm->local ? 0 : cm->rank;
  ShardBounds bounds;
  memset(&bounds, 0, sizeof(bounds));
  bounds.world = W;
  std::vector<uint64_t> rows_final(W, 0);
  {
    DevBuf<uint64_t> hist(kKeySpace);
    SGX_CUDA(cudaMemsetAsync(hist.p, 0, kKeySpace * sizeof(uint64_t), ctx->stream));
    const uint64_t stride = std::max<uint64_t>(1, N >> 19);  // about half a million sampled reads in total
    const uint64_t n_sample = (N + stride - 1) / stride;
    const uint64_t mine = (n_sample + nslices - 1 - slice) / nslices;  // samples slice, slice + nslices, ...
    if (mine)
      SGX_LAUNCH(ctx, key_hist_kernel, grid_for(mine, 256), 256, 0, acc, N, (uint64_t)slice * stride,
                 (uint64_t)nslices * stride, (unsigned long long*)hist.p);
    if (!cm->local) {
      uint64_t* hp = hist.p;
      comm_allreduce_sum_u64(cm, &hp, kKeySpace);
    }
    std::vector<uint64_t> h(kKeySpace);
    SGX_CUDA(cudaMemcpyAsync(h.data(), hist.p, kKeySpace * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    shard_bounds_from_hist(h.data(), W, bounds.b);

    DevBuf<uint64_t> rows(kMaxShards);
    SGX_CUDA(cudaMemsetAsync(rows.p, 0, kMaxShards * sizeof(uint64_t), ctx->stream));
    const uint64_t r0 = N / nslices * slice + std::min<uint64_t>(N % nslices, slice);
    const uint64_t r1 = r0 + N / nslices + ((uint64_t)slice < N % nslices ? 1 : 0);
    if (r1 > r0)
      SGX_LAUNCH(ctx, shard_rows_kernel, grid_for(r1 - r0, 256), 256, 0, acc, r0, r1, bounds, (unsigned long long*)rows.p);
    if (!cm->local) {
      uint64_t* rp = rows.p;
      comm_allreduce_sum_u64(cm, &rp, kMaxShards);
    }
    uint64_t hr[kMaxShards];
    SGX_CUDA(cudaMemcpyAsync(hr, rows.p, sizeof(hr), cudaMemcpyDeviceToHost, ctx->stream));
    SGX_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t sum = 0;
    for (int g = 0; g < W; ++g) {
      rows_final[g] = hr[g];
      sum += hr[g];
    }
    SGX_REQUIRE(sum == S, SGX_ECUDA, "sharded index: shard row counts do not add up to the number of suffixes");
  }
  {
    const char* tr = getenv("SGX_SHARD_TRACE");
    if (tr && tr[0] == '1') fprintf(stderr, "[sgx shard %d/%d] bucket bounds + rows per shard: %.2f ms\n", cm->shard_of(0), W, now_ms() - t_host0);
  }
  // ---- read order (.sai) and sentinel order, the reads sharded by the same bounds
  sharded_read_order(ctx, cm, reads, acc, order, bounds, d_sai.p, d_sord.p);
  t_sort.stop();

  uint8_t* rle_host = nullptr;
  uint64_t rle_total = 0, seg_offset = 0, seg_len = 0;
  // Positions inside the sharded build are shard-local rows and sender-local rank parts (the global offsets live in
  // the 64-bit constants of the receive plan), so 32 bits suffice whenever every SHARD has < 2^32 rows -- config 4
  // (5.03 G symbols) runs the 32-bit kernels on 2+ GPUs.  SGX_FORCE_WIDE=1 (tests) takes the 64-bit path anyway.
  uint64_t max_rows = 0;
  for (int g = 0; g < W; ++g) max_rows = std::max(max_rows, rows_final[g]);
  const char* fw = getenv("SGX_FORCE_WIDE");
  if (max_rows < 0xffffff00ull && !(fw && fw[0] == '1')) {
    sharded_build<uint32_t>(ctx, cm, reads, acc, d_sord.p, bounds, rows_final, out, k_merge, t_bcr, t_emit, segments, &rle_host, &rle_total,
                            &seg_offset, &seg_len);
  } else {
    sharded_build<uint64_t>(ctx, cm, reads, acc, d_sord.p, bounds, rows_final, out, k_merge, t_bcr, t_emit, segments, &rle_host, &rle_total,
                            &seg_offset, &seg_len);
  }
  out->n_strings = N;
  out->n_suffixes = S;
  out->n_runs = rle_total;
  out->seg_offset = seg_offset;
  out->seg_bytes = seg_len;
  out->rle = rle_host ? rle_host : (uint8_t*)pinned_get(ctx, 1);  // only shard 0 holds the concatenated run bytes
  out->sai = (uint32_t*)pinned_get(ctx, N * sizeof(uint32_t));
  SGX_CUDA(cudaMemcpyAsync(out->sai, d_sai.p, N * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  t_all.stop();
  SGX_CUDA(cudaStreamSynchronize(ctx->stream));
  out->ms_total = t_all.ms();
  out->ms_sort = t_sort.ms();
  out->ms_bcr = t_bcr.ms();
  out->ms_emit = t_emit.ms();
  out->ms_merge = k_merge.total_ms();
  out->launches = ctx->launches - launches0;
}

}  // namespace sgx
