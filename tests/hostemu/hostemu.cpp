// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or called from the product library.
//
// Single-lane host instantiation of siga_b200/csrc/overlap_logic.h (the exact per-read control logic the CUDA
// kernels run, with sgx::HostGroup in place of the 8-lane device group) over the same 128-byte rank-block layout.
// There is no GPU in the build container, so this is how the intricate list logic (sub-maximal filter,
// irreducible extraction) is debugged against the compiled reference before a GPU run is spent on it.
// The product path (libsiga_b200.so) has no CPU fallback and fails loudly without a device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../siga_b200/csrc/correct_logic.h"
#include "../../siga_b200/csrc/overlap_logic.h"
#include "../../siga_b200/csrc/shard_plan.h"

using namespace sgx;

struct HemuFm {
  std::vector<FmBlockLane> blocks;
  FmView view;
};

static HemuFm* build_from_syms(const std::vector<uint8_t>& sym) {
  HemuFm* fm = new HemuFm();
  const uint64_t n = sym.size();
  const uint64_t nb = fm_num_blocks(n);
  fm->blocks.assign(nb * kGroup, FmBlockLane{0, 0, 0, 0});
  uint64_t cnt[5] = {0, 0, 0, 0, 0};
  for (uint64_t b = 0; b < nb; ++b) {
    FmBlockLane* blk = &fm->blocks[b * kGroup];
    for (int c = 0; c < 4; ++c) {
      blk[c].h = (uint32_t)(cnt[c + 1] & 0xffffffffu);
      blk[c + 4].h = (uint32_t)(cnt[c + 1] >> 32);
    }
    for (int j = 0; j < kBlockSyms; ++j) {
      const uint64_t i = b * kBlockSyms + j;
      if (i >= n) break;
      const int s = sym[i];
      FmBlockLane& l = blk[j >> 5];
      const uint32_t bit = 1u << (j & 31);
      if (s == 0) {
        l.p2 |= bit;
      } else {
        if ((s - 1) & 1) l.p0 |= bit;
        if ((s - 1) & 2) l.p1 |= bit;
      }
      cnt[s]++;
    }
  }
  fm->view.blocks = fm->blocks.data();
  fm->view.n = n;
  fm->view.pred[0] = 0;
  for (int c = 1; c <= 5; ++c) fm->view.pred[c] = fm->view.pred[c - 1] + cnt[c - 1];
  return fm;
}

extern "C" {

void* hemu_fm_from_rle(const uint8_t* rle, uint64_t n_runs) {
  std::vector<uint8_t> sym;
  for (uint64_t i = 0; i < n_runs; ++i) {
    const int s = rle[i] >> 5, c = rle[i] & 31;
    for (int k = 0; k < c; ++k) sym.push_back((uint8_t)s);
  }
  return build_from_syms(sym);
}

void hemu_fm_free(void* p) { delete (HemuFm*)p; }

uint64_t hemu_fm_length(void* p) { return ((HemuFm*)p)->view.n; }

// getOcc(c, pos) for one symbol through fm_lane_sym_count (the per-lane arithmetic of quad_rank_sym): the eight lane
// words of the block summed the way the four lanes of a quad sum them (lane q: words q and q + 4).
uint64_t hemu_fm_occ_sym(void* p, uint64_t pos, int c) {
  const FmView& fm = ((HemuFm*)p)->view;
  const uint64_t pl = pos + 1;
  if (pl == 0) return 0;
  const FmBlockLane* blk = fm.blocks + (pl >> kBlockShift) * kGroup;
  const int r = (int)(pl & (kBlockSyms - 1));
  const uint32_t x0 = ((c - 1) & 1) ? 0u : 0xffffffffu, x1 = ((c - 1) & 2) ? 0u : 0xffffffffu;
  uint32_t packed = 0;
  for (int q = 0; q < 4; ++q)
    packed += fm_lane_sym_count(blk[q].p0, blk[q].p1, blk[q].p2, q, r, x0, x1) +
              fm_lane_sym_count(blk[q + 4].p0, blk[q + 4].p1, blk[q + 4].p2, q + 4, r, x0, x1);
  return ((uint64_t)blk[c + 3].h << 32 | blk[c - 1].h) + packed;
}

void hemu_fm_occ(void* p, const uint64_t* pos, uint64_t n, uint64_t* counts) {
  HostGroup g;
  for (uint64_t i = 0; i < n; ++i) fm_occ(((HemuFm*)p)->view, pos[i], counts + 5 * i, g);
}

// Same contract as sgx_overlap_host, but memory comes from malloc (release with hemu_hits_free).
int hemu_overlap(void* fwd_, void* rev_, const char* ascii, const uint64_t* off, uint64_t n, uint32_t min_ov,
                 int irreducible, int rc, sgx_hits* out) {
  const FmView& fwd = ((HemuFm*)fwd_)->view;
  const FmView& rev = ((HemuFm*)rev_)->view;
  HostGroup g;
  uint32_t lmax = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t l = off[i + 1] - off[i];
    if (l == 0) return SGX_EINVAL;
    if (l > lmax) lmax = (uint32_t)l;
  }
  const uint32_t find_cap = lmax > min_ov ? lmax - min_ov : 1;
  std::vector<sgx_block> all;
  std::vector<uint64_t> boff(n + 1, 0);
  std::vector<uint8_t> sub(n, 0);
  uint64_t lookups_s = 0, lookups_e = 0;
  uint32_t status = 0;
  uint32_t list_cap = find_cap + 4 + 64;
  uint32_t st_cap = 4 * list_cap + 256;
  for (int attempt = 0; attempt < 6; ++attempt) {
    status = 0;
    all.clear();
    lookups_s = lookups_e = 0;
    std::vector<sgx_block> cand(4 * (size_t)find_cap), contain(4), lists(4 * (size_t)list_cap), R(list_cap),
        st(st_cap), outb(st_cap);
    std::vector<uint32_t> gd(4 * (size_t)st_cap);
    for (uint64_t i = 0; i < n; ++i) {
      const uint32_t l = (uint32_t)(off[i + 1] - off[i]);
      std::vector<uint32_t> rw((l + 15) / 16, 0);
      for (uint32_t j = 0; j < l; ++j) {
        int code;
        switch (ascii[off[i] + j]) {
          case 'A': code = 0; break;
          case 'C': code = 1; break;
          case 'G': code = 2; break;
          case 'T': code = 3; break;
          default: return SGX_EINVAL;
        }
        rw[j >> 4] |= (uint32_t)code << (2 * (j & 15));
      }
      uint32_t ncand[4] = {0, 0, 0, 0}, ncontain[4] = {0, 0, 0, 0};
      uint8_t substring = 0;
      for (int mode = 0; mode < 4; ++mode) {
        if (!rc && (mode == 1 || mode == 3)) continue;
        const FmView& P = mode < 2 ? fwd : rev;
        const FmView& S = mode < 2 ? rev : fwd;
        ov_find(P, S, rw.data(), l, mode, min_ov, cand.data() + (size_t)mode * find_cap, find_cap, &ncand[mode],
                &contain[mode], &ncontain[mode], &substring, &status, &lookups_s, g);
      }
      OvArena ar;
      ar.lists = lists.data();
      ar.list_cap = list_cap;
      ar.R = R.data();
      ar.r_cap = list_cap;
      ar.st = st.data();
      ar.st_cap = st_cap;
      ar.gd = gd.data();
      ar.g_cap = st_cap;
      const sgx_block* cp[4] = {cand.data(), cand.data() + find_cap, cand.data() + 2 * (size_t)find_cap,
                                cand.data() + 3 * (size_t)find_cap};
      uint32_t nout = 0;
      ov_post(fwd, rev, l, irreducible, cp, ncand, contain.data(), ncontain, ar, outb.data(), &nout, st_cap, &status,
              &lookups_e, g);
      sub[i] = substring;
      boff[i] = all.size();
      all.insert(all.end(), outb.begin(), outb.begin() + nout);
    }
    boff[n] = all.size();
    if (!(status & kOvOverflow)) break;
    list_cap *= 4;
    st_cap *= 4;
  }
  if (status & kOvOverflow) return SGX_EOVERFLOW;
  memset(out, 0, sizeof(*out));
  out->n_reads = n;
  out->n_blocks = all.size();
  out->substring = (uint8_t*)malloc(n ? n : 1);
  memcpy(out->substring, sub.data(), n);
  out->block_off = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
  memcpy(out->block_off, boff.data(), (n + 1) * sizeof(uint64_t));
  out->blocks = (sgx_block*)malloc((all.size() ? all.size() : 1) * sizeof(sgx_block));
  memcpy(out->blocks, all.data(), all.size() * sizeof(sgx_block));
  out->lookups_search = lookups_s;
  out->lookups_extract = lookups_e;
  out->launches = status;  // status bits are surfaced here for the tests
  return SGX_OK;
}

void hemu_hits_free(sgx_hits* h) {
  free(h->substring);
  free(h->block_off);
  free(h->blocks);
  memset(h, 0, sizeof(*h));
}

// ---- k-mer error correction (siga_b200/csrc/correct_logic.h) on the host engine.  `out` receives the reads with the
//      fixes applied (same offsets); valid[i] = the read ends all-solid, i.e. the reference would write it.
static uint64_t g_corr_stats[4];
// {certificate calls, their warp steps (longest query per call), try2Correct calls, their warp steps} of the last hemu_correct
void hemu_correct_stats(uint64_t out[4]) { memcpy(out, g_corr_stats, sizeof(g_corr_stats)); }

int hemu_correct(void* fm_, const char* ascii, const char* qual, const uint64_t* off, uint64_t n, uint32_t k, int threshold,
                 uint32_t rounds, uint64_t count_offset, char* out, uint8_t* valid, uint32_t* n_fixed, uint64_t* lookups) {
  const FmView& fm = ((HemuFm*)fm_)->view;
  if (k == 0) return SGX_EINVAL;
  CorrectParams P;
  P.k = k;
  P.rounds = rounds;
  P.low = (uint64_t)(int64_t)threshold;
  P.high = (uint64_t)(int64_t)(threshold + 1);
  P.count_offset = count_offset;
  uint64_t look = 0;
  memset(g_corr_stats, 0, sizeof(g_corr_stats));
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t l = (uint32_t)(off[i + 1] - off[i]);
    if (l == 0) return SGX_EINVAL;
    std::vector<uint32_t> seq((l + 15) / 16, 0), hq((l + 31) / 32, 0), ks((l + 31) / 32, 0), us((l + 31) / 32, 0),
        kh((l + 31) / 32, 0), kn((l + 31) / 32, 0), win(kCorrWinWords, 0), dead(kCorrDeadCap, 0), tmp((l + 31) / 32, 0);
    for (uint32_t j = 0; j < l; ++j) {
      int code;
      switch (ascii[off[i] + j]) {
        case 'A': code = 0; break;
        case 'C': code = 1; break;
        case 'G': code = 2; break;
        case 'T': code = 3; break;
        default: return SGX_EINVAL;
      }
      seq[j >> 4] |= (uint32_t)code << (2 * (j & 15));
      if (qual && (int)(uint8_t)qual[off[i] + j] - kCorrPhredBase >= kCorrPhredCutoff) hq[j >> 5] |= 1u << (j & 31);
    }
    CorrectState st;
    st.seq = seq.data();
    st.hq = qual ? hq.data() : nullptr;
    st.khigh = kh.data();
    st.known = kn.data();
    st.ksolid = ks.data();
    st.unsolid = us.data();
    st.win = win.data();
    st.dead = dead.data();
    st.tmp = tmp.data();
    HostCorrEng eng;
    eng.fm = &fm;
    eng.seq = seq.data();
    eng.k = k;
    eng.low = P.low;
    eng.high = P.high;
    eng.lookups = 0;
    eng.win_calls = eng.win_steps = eng.qual_calls = eng.qual_steps = 0;
    uint32_t nf = 0;
    valid[i] = correct_read(eng, P, l, st, &nf) ? 1 : 0;
    if (n_fixed) n_fixed[i] = nf;
    look += eng.lookups;
    g_corr_stats[0] += eng.win_calls;
    g_corr_stats[1] += eng.win_steps;
    g_corr_stats[2] += eng.qual_calls;
    g_corr_stats[3] += eng.qual_steps;
    for (uint32_t j = 0; j < l; ++j) out[off[i] + j] = valid[i] ? "ACGT"[read_rank(seq.data(), j) - 1] : ascii[off[i] + j];
  }
  if (lookups) *lookups = look;
  return SGX_OK;
}

// ---- host-side planning of the bucket-sharded index build (siga_b200/csrc/shard_plan.h), for the CPU emulation in
//      tests/test_cpu_sharded_index.py
int hemu_plan_consts(int* max_shards, int* send_buckets, int* meta_n, uint32_t* key_space, uint32_t* key_top) {
  *max_shards = kPlanMaxShards;
  *send_buckets = kPlanSendBuckets;
  *meta_n = kPlanMetaN;
  *key_space = kPlanKeySpace;
  *key_top = kPlanKeyTop;
  return (int)sizeof(RecvPlan);
}
void hemu_shard_bounds(const uint64_t* hist, int world, uint32_t* b) { shard_bounds_from_hist(hist, world, b); }
int hemu_shard_owner(const uint32_t* b, int world, uint32_t key) { return shard_owner(b, world, key); }
uint64_t hemu_shard_pass_counts(const uint64_t* M, int W, const uint64_t* n_cur, uint64_t* cnt, uint64_t* soff, uint64_t* roff,
                                uint64_t* n_in, uint64_t* n_new, uint64_t* first_row) {
  return shard_pass_counts(M, W, n_cur, cnt, soff, roff, n_in, n_new, first_row);
}
// plan flattened for ctypes: out0[entries+1], src0[entries], K[entries]
uint64_t hemu_shard_recv_plan(const uint64_t* M, int W, int g, uint64_t n_reads, const uint64_t* roff, const uint64_t* first_row,
                              uint64_t* out0, uint64_t* src0, long long* K) {
  RecvPlan plan;
  const uint64_t n = shard_recv_plan(M, W, g, n_reads, roff, first_row, &plan);
  for (int e = 0; e < plan.entries; ++e) {
    out0[e] = plan.out0[e];
    src0[e] = plan.src0[e];
    K[e] = plan.K[e];
  }
  out0[plan.entries] = plan.out0[plan.entries];
  return n;
}
void hemu_shard_seam_rule(const uint64_t* seam, int W, int g, uint64_t my_runs, int* skip_first, uint64_t* extend_last) {
  shard_seam_rule(seam, W, g, my_runs, skip_first, extend_last);
}

}  // extern "C"
