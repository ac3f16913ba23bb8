// Host-side planning of the bucket-sharded index build (index_sharded.cu): everything the ranks compute from the
// values they all-gather, in plain C++ so that the same code is exercised on CPU by tests/hostemu (numpy emulation of
// the sharded column passes, single process and two gloo ranks) -- no CUDA, no NCCL in this header.
#ifndef SGX_SHARD_PLAN_H_
#define SGX_SHARD_PLAN_H_

#include <stdint.h>
#include <string.h>

namespace sgx {

static const int kPlanMaxShards = 16;
static const int kPlanSendBuckets = kPlanMaxShards * 4;   // (destination shard, first base of the new suffix)
static const int kPlanMetaN = kPlanSendBuckets + 4;       // gathered per shard and pass: send counts + A/C/G/T totals
static const uint32_t kPlanKeySpace = 390625;             // 5^8 leading-8-mer keys ('$' = 0 ends a key)
static const uint32_t kPlanKeyTop = 78125;                // 5^7

// Receiver side of a column pass.  Entry e = c * world + s: the run of inserts that sender s sends for symbol c.
struct RecvPlan {
  uint64_t out0[kPlanSendBuckets + 1];  // first output slot of the run (non-decreasing; out0[entries] = inserts)
  uint64_t src0[kPlanSendBuckets];      // first slot of the run in the receive buffer
  long long K[kPlanSendBuckets];        // local row = sender-local rank part + K
  int entries;
  int world;
};

// Shard bounds b[0..world] from a key histogram: shard g owns keys [b[g], b[g+1]).  Balanced by cumulative count;
// key 0 (the N sentinel rows) always stays on shard 0; equal neighbours make empty shards.
inline void shard_bounds_from_hist(const uint64_t* hist, int world, uint32_t* b /* kPlanMaxShards + 1 */) {
  uint64_t total = 0;
  for (uint32_t k = 0; k < kPlanKeySpace; ++k) total += hist[k];
  b[0] = 0;
  uint64_t cum = 0;
  uint32_t key = 0;
  for (int g = 1; g < world; ++g) {
    const uint64_t want = total / world * g + (total % world) * g / world;
    while (key < kPlanKeySpace && cum < want) cum += hist[key++];
    uint32_t v = key > 1u ? key : 1u;
    if (v < b[g - 1]) v = b[g - 1];
    b[g] = v;
  }
  for (int g = world; g <= kPlanMaxShards; ++g) b[g] = kPlanKeySpace;
}

inline int shard_owner(const uint32_t* b, int world, uint32_t key) {
  int g = 0;
  for (int i = 1; i < world; ++i) g += key >= b[i] ? 1 : 0;
  return g;
}

// From the gathered meta matrix M[world][kPlanMetaN] and the rows every shard holds before the pass:
//   cnt[s*W+d]  items s sends to d          soff[s*W+d]  their first slot in s's send buffer (bucket order: d, then c)
//   roff[d*W+s] first slot of s's block in d's receive buffer (sender-major)
//   n_in[d]     inserts d receives          n_new[d]     rows of d after the pass
//   first_row[d] global row of d's first row after the pass (first_row[W] = all rows)
// Returns the number of items moved by the pass (0 = every read has ended).
inline uint64_t shard_pass_counts(const uint64_t* M, int W, const uint64_t* n_cur, uint64_t* cnt, uint64_t* soff,
                                  uint64_t* roff, uint64_t* n_in, uint64_t* n_new, uint64_t* first_row) {
  uint64_t moved = 0;
  for (int s = 0; s < W; ++s) {
    uint64_t run = 0;
    for (int d = 0; d < W; ++d) {
      uint64_t c4 = 0;
      for (int c = 0; c < 4; ++c) c4 += M[(size_t)s * kPlanMetaN + d * 4 + c];
      cnt[s * W + d] = c4;
      soff[s * W + d] = run;
      run += c4;
    }
    moved += run;
  }
  first_row[0] = 0;
  for (int d = 0; d < W; ++d) {
    uint64_t run = 0;
    for (int s = 0; s < W; ++s) {
      roff[d * W + s] = run;
      run += cnt[s * W + d];
    }
    n_in[d] = run;
    n_new[d] = n_cur[d] + run;
    first_row[d + 1] = first_row[d] + n_new[d];
  }
  return moved;
}

// Placement plan of receiver g.  Every sender's block arrives ordered by (c, sender row); for one c the new row grows
// with the old global row, i.e. with (sender, sender row), so the inserts in row order are, for c = A..T, the run of
// sender 0, of sender 1, ...  The sender-local rank part (symbols < c in the sender's rows + c's before the item) turns
// into a local row by adding
//     N + sum_{c'<c} T[c'] + sum_{s'<s} tot[s'][c] - sum_{c'<c} tot[s][c'] - first_row[g]
// (N sentinel rows, T = global symbol totals, tot = per-shard totals: the last kPlanMetaN - kPlanSendBuckets values).
// Returns the number of inserts covered (must equal n_in[g]).
inline uint64_t shard_recv_plan(const uint64_t* M, int W, int g, uint64_t n_reads, const uint64_t* roff,
                                const uint64_t* first_row, RecvPlan* plan) {
  memset(plan, 0, sizeof(*plan));
  plan->world = W;
  plan->entries = 4 * W;
  uint64_t T[4] = {0, 0, 0, 0};
  for (int s = 0; s < W; ++s)
    for (int c = 0; c < 4; ++c) T[c] += M[(size_t)s * kPlanMetaN + kPlanSendBuckets + c];
  long long less_global = (long long)n_reads;  // rows of suffixes starting with a symbol < c
  uint64_t out_run = 0;
  for (int c = 0; c < 4; ++c) {
    long long before = 0;  // symbols c in the rows of shards before the sender
    for (int q = 0; q < W; ++q) {
      const uint64_t* mq = &M[(size_t)q * kPlanMetaN];
      long long less_local = 0;
      uint64_t src = roff[g * W + q];  // the sender's block, then its runs of smaller symbols
      for (int c2 = 0; c2 < c; ++c2) {
        less_local += (long long)mq[kPlanSendBuckets + c2];
        src += mq[g * 4 + c2];
      }
      const int e = c * W + q;
      plan->out0[e] = out_run;
      plan->src0[e] = src;
      plan->K[e] = less_global + before - less_local - (long long)first_row[g];
      out_run += mq[g * 4 + c];
      before += (long long)mq[kPlanSendBuckets + c];
    }
    less_global += (long long)T[c];
  }
  plan->out0[4 * W] = out_run;
  return out_run;
}

// Seam handling of the run-length encoding.  seam[s*5 + {0: rows, 1: first symbol, 2: rows of the first run,
// 3: last symbol, 4: rows of the last run}] of every shard; a run that crosses seams belongs to the shard where it
// starts: shard g skips its first run when it continues the last run of the previous non-empty shard, and its last
// run is extended by the rows of the following shards that continue it.
inline void shard_seam_rule(const uint64_t* seam, int W, int g, uint64_t my_runs, int* skip_first, uint64_t* extend_last) {
  *skip_first = 0;
  *extend_last = 0;
  const uint64_t* me = &seam[(size_t)g * 5];
  if (me[0] == 0) return;
  for (int p = g - 1; p >= 0; --p) {
    if (seam[(size_t)p * 5 + 0] == 0) continue;
    *skip_first = seam[(size_t)p * 5 + 3] == me[1] ? 1 : 0;
    break;
  }
  if (my_runs == 1 && *skip_first) return;  // the whole shard is the tail of an earlier run
  for (int q = g + 1; q < W; ++q) {
    const uint64_t* sq = &seam[(size_t)q * 5];
    if (sq[0] == 0) continue;
    if (sq[1] != me[3]) break;
    *extend_last += sq[2];
    if (sq[2] != sq[0]) break;  // that shard has further runs
  }
}

}  // namespace sgx

#endif  // SGX_SHARD_PLAN_H_
