// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or shipped with the product (siga_b200/libsiga_b200.so,
// siga_b200/bin/siga have no CPU path and fail without a B200).
//
// A host stand-in for the part of the C ABI (include/siga_b200.h) that `siga index | overlap | rmdup | correct | match` call, built on the
// single-lane instantiations in hostemu.cpp.  Linked with the UNMODIFIED front-end sources
// (siga_b200/csrc/host/siga_main.cpp, seqio.cpp) into build/siga_hostemu, it lets the CPU test-suite run the
// front-end's own host logic -- option parsing, FASTQ qualities, fix application, ASQG / FASTA formatting, the fork /
// part file / concatenation path of --gpus N -- against the reference binary without a GPU.  The index build is a
// brute-force suffix sort (small inputs); the NCCL-sharded build has no host stand-in and returns SGX_ENODEV.
#include "hostemu.cpp"

#include <algorithm>
#include <unistd.h>

#include <string>

struct sgx_ctx {
  std::string err;
};
struct sgx_reads {
  std::string ascii;
  std::vector<uint64_t> off;  // rebased to 0
};
struct sgx_fm {
  void* fm;
  std::vector<uint32_t> sai;  // sgx_fm_attach_reads
};
struct sgx_comm {
  int rank, world;
};

namespace {
// SGX_STUB_FAKE=1: no index is built and no search runs; correct writes every read unchanged and overlap returns three
// synthetic edges per read.  Only for timing the front-end's own host phases (load / format / compress / write) at sizes
// the host engines would take hours for.
bool fake_mode() {
  static const bool f = getenv("SGX_STUB_FAKE") && getenv("SGX_STUB_FAKE")[0] == '1';
  return f;
}
}  // namespace

extern "C" {

// Failure injection for the tests of the front-end's worker coordination (`--gpus N`): the rank whose device index equals
// SGX_STUB_FAIL_DEVICE fails in sgx_reads_pack; the rank on SGX_STUB_HANG_DEVICE never returns from it (a peer stuck in
// a collective).
static int g_device = 0;
int sgx_init(int device, sgx_ctx** out) {
  g_device = device;
  *out = new sgx_ctx();
  return SGX_OK;
}
void sgx_shutdown(sgx_ctx* ctx) { delete ctx; }
const char* sgx_last_error(const sgx_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }

int sgx_reads_pack(sgx_ctx* ctx, const char* ascii, const uint64_t* off, uint64_t n, sgx_reads** out) {
  if (!ctx || !ascii || !off || n == 0) return SGX_EINVAL;
  if (getenv("SGX_STUB_HANG_DEVICE") && atoi(getenv("SGX_STUB_HANG_DEVICE")) == g_device)
    for (;;) pause();
  if (getenv("SGX_STUB_FAIL_DEVICE") && atoi(getenv("SGX_STUB_FAIL_DEVICE")) == g_device) {
    ctx->err = "sgx_reads_pack: injected failure";
    return SGX_ENOMEM;
  }
  sgx_reads* r = new sgx_reads();
  r->ascii.assign(ascii + off[0], off[n] - off[0]);
  r->off.resize(n + 1);
  for (uint64_t i = 0; i <= n; ++i) r->off[i] = off[i] - off[0];
  for (char ch : r->ascii)
    if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') {
      ctx->err = "sgx_reads_pack: base other than A/C/G/T";
      delete r;
      return SGX_EINVAL;
    }
  *out = r;
  return SGX_OK;
}
void sgx_reads_free(sgx_reads* r) { delete r; }

int sgx_fm_create(sgx_ctx* ctx, const uint8_t* rle, uint64_t n_runs, uint64_t n_suffixes, sgx_fm** out) {
  (void)ctx;
  sgx_fm* f = new sgx_fm();
  if (fake_mode()) {
    f->fm = nullptr;
    *out = f;
    return SGX_OK;
  }
  f->fm = hemu_fm_from_rle(rle, n_runs);
  if (hemu_fm_length(f->fm) != n_suffixes) {
    hemu_fm_free(f->fm);
    delete f;
    return SGX_EINVAL;
  }
  *out = f;
  return SGX_OK;
}
void sgx_fm_free(sgx_fm* f) {
  if (!f) return;
  if (f->fm) hemu_fm_free(f->fm);
  delete f;
}
int sgx_fm_attach_reads(sgx_ctx*, sgx_fm* f, const uint32_t* sai, uint64_t n, const sgx_reads*) {
  f->sai.assign(sai, sai + n);
  return SGX_OK;
}

int sgx_correct_batch(sgx_ctx* ctx, const sgx_fm* fm, const sgx_reads* reads, uint64_t first, uint64_t count, const char* qual,
                      const uint64_t* qual_off, const sgx_correct_params* prm, sgx_corrected* out) {
  if (!ctx || !fm || !reads || !prm || !out || first + count > reads->off.size() - 1) return SGX_EINVAL;
  memset(out, 0, sizeof(*out));
  out->n_reads = count;
  out->valid = (uint8_t*)calloc(count ? count : 1, 1);
  if (fake_mode()) {
    memset(out->valid, 1, count);
    out->fixes = (sgx_fix*)malloc(sizeof(sgx_fix));
    return SGX_OK;
  }
  std::vector<uint64_t> off(count + 1);
  for (uint64_t i = 0; i <= count; ++i) off[i] = reads->off[first + i] - reads->off[first];
  const char* in = reads->ascii.data() + reads->off[first];
  std::string q;
  if (qual) {
    for (uint64_t i = 0; i < count; ++i)
      if (qual_off[i + 1] - qual_off[i] != off[i + 1] - off[i]) {
        ctx->err = "sgx_correct_batch: a quality string differs in length from its read";
        free(out->valid);
        out->valid = nullptr;
        return SGX_EINVAL;
      }
    q.assign(qual + qual_off[0], qual_off[count] - qual_off[0]);
  }
  std::string fixed(off[count], '\0');
  std::vector<uint32_t> nfix(count ? count : 1);
  uint64_t lookups = 0;
  const int rc = count ? hemu_correct(fm->fm, in, qual ? q.data() : nullptr, off.data(), count, prm->kmer_size, prm->kmer_threshold,
                                      prm->kmer_rounds, prm->kmer_count_offset, &fixed[0], out->valid, nfix.data(), &lookups)
                       : SGX_OK;
  if (rc != SGX_OK) {
    free(out->valid);
    out->valid = nullptr;
    return rc;
  }
  std::vector<sgx_fix> fixes;
  for (uint64_t i = 0; i < count; ++i) {
    if (!out->valid[i]) continue;
    bool any = false;
    for (uint64_t j = off[i]; j < off[i + 1]; ++j)
      if (fixed[j] != in[j]) {
        uint32_t code = fixed[j] == 'A' ? 0 : (fixed[j] == 'C' ? 1 : (fixed[j] == 'G' ? 2 : 3));
        fixes.push_back(sgx_fix{(uint32_t)i, (uint32_t)((j - off[i]) << 2) | code});
        any = true;
      }
    out->fixed_reads += any ? 1 : 0;
  }
  out->n_fixes = fixes.size();
  out->fixes = (sgx_fix*)malloc((fixes.size() ? fixes.size() : 1) * sizeof(sgx_fix));
  if (!fixes.empty()) memcpy(out->fixes, fixes.data(), fixes.size() * sizeof(sgx_fix));
  out->lookups = lookups;
  return SGX_OK;
}
void sgx_corrected_free(sgx_corrected* out) {
  if (!out) return;
  free(out->valid);
  free(out->fixes);
  out->valid = nullptr;
  out->fixes = nullptr;
}

// `siga rmdup`: hit blocks of a read range through overlap_logic.h on the host lane
int sgx_overlap_batch(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first, uint64_t count,
                      uint32_t min_overlap, int irreducible, int rc, sgx_hits* out) {
  if (!ctx || !fwd || !rev || !reads || !out || first + count > reads->off.size() - 1) return SGX_EINVAL;
  if (fake_mode()) {  // no blocks at all: every read is kept
    memset(out, 0, sizeof(*out));
    out->n_reads = count;
    out->substring = (uint8_t*)calloc(count ? count : 1, 1);
    out->block_off = (uint64_t*)calloc(count + 1, sizeof(uint64_t));
    out->blocks = (sgx_block*)malloc(sizeof(sgx_block));
    return SGX_OK;
  }
  return hemu_overlap(fwd->fm, rev->fm, reads->ascii.data(), reads->off.data() + first, count, min_overlap, irreducible, rc, out);
}
void sgx_hits_free(sgx_hits* h) { hemu_hits_free(h); }

// `siga overlap`: the same blocks turned into edge records by Hit2OverlapConverter::convert + OverlapBlock::overlap
// (src/overlap_builder.cpp:158-175, 345-375), restated for the host (the product does this in ov_edges_kernel /
// ov_reduce_kernel on the device)
int sgx_overlap_edges(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first, uint64_t count,
                      uint32_t min_overlap, int irreducible, int rc, const uint32_t* name_rank, sgx_edges* out) {
  if (!ctx || !fwd || !rev || !reads || !name_rank || !out || first + count > reads->off.size() - 1) return SGX_EINVAL;
  if (fwd->sai.size() != reads->off.size() - 1 || rev->sai.size() != reads->off.size() - 1) return SGX_EINVAL;
  if (fake_mode()) {
    const uint64_t n = reads->off.size() - 1;
    memset(out, 0, sizeof(*out));
    out->n_reads = count;
    out->substring = (uint8_t*)calloc(count ? count : 1, 1);
    out->n_edges = 3 * count;
    out->edges = (sgx_edge*)malloc((out->n_edges ? out->n_edges : 1) * sizeof(sgx_edge));
    uint64_t x = 88172645463325252ull;
    for (uint64_t k = 0; k < out->n_edges; ++k) {
      x ^= x << 13;
      x ^= x >> 7;
      x ^= x << 17;
      out->edges[k] = sgx_edge{(uint32_t)(first + k / 3), (uint32_t)(x % n), min_overlap + (uint32_t)(x >> 40) % 50, (uint32_t)(x >> 60) & 7u};
    }
    return SGX_OK;
  }
  sgx_hits h;
  const int st = hemu_overlap(fwd->fm, rev->fm, reads->ascii.data(), reads->off.data() + first, count, min_overlap, irreducible, rc, &h);
  if (st != SGX_OK) return st;
  std::vector<sgx_edge> edges;
  for (uint64_t k = 0; k < count; ++k) {
    const uint64_t q = first + k, lq = reads->off[q + 1] - reads->off[q];
    for (uint64_t b = h.block_off[k]; b < h.block_off[k + 1]; ++b) {
      const sgx_block& blk = h.blocks[b];
      const std::vector<uint32_t>& sa = (blk.af & SGX_AF_TARGETREV) ? rev->sai : fwd->sai;
      for (uint64_t j = blk.capped[0]; j <= blk.capped[1]; ++j) {
        const uint32_t t = sa[j];
        const uint64_t lt = reads->off[t + 1] - reads->off[t];
        if (name_rank[q] == name_rank[t]) continue;  // same name: a self hit
        const bool contained = blk.length == lq || blk.length == lt;
        if (name_rank[q] < name_rank[t] || (contained && (blk.af & SGX_AF_QUERYREV))) continue;
        edges.push_back(sgx_edge{(uint32_t)q, t, blk.length, blk.af});
      }
    }
  }
  memset(out, 0, sizeof(*out));
  out->n_reads = count;
  out->substring = (uint8_t*)malloc(count ? count : 1);
  memcpy(out->substring, h.substring, count);
  out->n_edges = edges.size();
  out->edges = (sgx_edge*)malloc((edges.size() ? edges.size() : 1) * sizeof(sgx_edge));
  if (!edges.empty()) memcpy(out->edges, edges.data(), edges.size() * sizeof(sgx_edge));
  out->n_blocks = h.n_blocks;
  hemu_hits_free(&h);
  return SGX_OK;
}
void sgx_edges_free(sgx_edges* e) {
  if (!e) return;
  free(e->substring);
  free(e->edges);
  e->substring = nullptr;
  e->edges = nullptr;
}

// FMIndex::Interval::occurrences (src/fmindex.h:67-86) per query on the scalar rank query
int sgx_fm_count(sgx_ctx* ctx, const sgx_fm* fm, const char* ascii, const uint64_t* off, uint64_t n, uint64_t* counts) {
  if (!ctx || !fm || !ascii || !off || (!counts && n)) return SGX_EINVAL;
  const FmView& v = ((HemuFm*)fm->fm)->view;
  HostGroup g;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t j = off[i + 1] - off[i];
    if (j == 0) return SGX_EINVAL;
    int rk[256];
    memset(rk, 0, sizeof(rk));
    rk[(int)'A'] = 1;
    rk[(int)'C'] = 2;
    rk[(int)'G'] = 3;
    rk[(int)'T'] = 4;
    const char* w = ascii + off[i];
    for (uint64_t t = 0; t < j; ++t)
      if (!rk[(unsigned char)w[t]]) {
        ctx->err = "sgx_reads_pack: base other than A/C/G/T";
        return SGX_EINVAL;
      }
    int c = rk[(unsigned char)w[j - 1]];
    pos_t lo = v.pred[c], hi = v.pred[c + 1] - 1;
    uint64_t L[5], U[5];
    while (--j > 0 && ival_valid(lo, hi)) {
      c = rk[(unsigned char)w[j - 1]];
      fm_occ(v, lo - 1, L, g);
      fm_occ(v, hi, U, g);
      lo = v.pred[c] + L[c];
      hi = v.pred[c] + U[c] - 1;
    }
    counts[i] = ival_valid(lo, hi) ? hi - lo + 1 : 0;
  }
  return SGX_OK;
}

int sgx_fm_pred(const sgx_fm* fm, uint64_t pred[5]) {
  if (!fm || !pred) return SGX_EINVAL;
  const FmView& v = ((HemuFm*)fm->fm)->view;
  for (int c = 0; c < 5; ++c) pred[c] = v.pred[c];
  return SGX_OK;
}

int sgx_fm_occ(sgx_ctx* ctx, const sgx_fm* fm, const uint64_t* pos, uint64_t n, uint64_t* counts) {
  if (!ctx || !fm || !pos || !counts) return SGX_EINVAL;
  hemu_fm_occ(fm->fm, pos, n, counts);
  return SGX_OK;
}

// ---- no host stand-in
// `siga index`: brute-force suffix sorting of the two ordering models (the statement
// oracle/pyoracle.py checks against the reference's golden vectors), small inputs only.
//   model B (sais2): plain suffix array of T = R0 $ R1 $ ... with every '$' the smallest symbol and end-of-text smaller still;
//   model A (sais):  suffixes ordered by (R_i[j:] $, i).
// BWT symbol of a row = the symbol before the suffix, '$' at a read start; runs are cut at 31 (src/bwt.cpp:16-27).
int sgx_index_build(sgx_ctx* ctx, const sgx_reads* reads, int order, int reversed, sgx_index_out* out) {
  if (!ctx || !reads || !out || (order != SGX_ORDER_SAIS && order != SGX_ORDER_SAIS2)) return SGX_EINVAL;
  const uint64_t n = reads->off.size() - 1;
  if (fake_mode()) {  // run bytes of a realistic volume (one per ~5 symbols) and the identity read order
    memset(out, 0, sizeof(*out));
    out->n_strings = n;
    out->n_suffixes = reads->ascii.size() + n;
    out->n_runs = out->n_suffixes / 5;
    out->seg_bytes = out->n_runs;
    out->rle = (uint8_t*)malloc(out->n_runs ? out->n_runs : 1);
    memset(out->rle, 0x41, out->n_runs);
    out->sai = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    for (uint64_t i = 0; i < n; ++i) out->sai[i] = (uint32_t)i;
    return SGX_OK;
  }
  std::vector<uint8_t> text;  // ranks, 0 = '$'
  std::vector<uint64_t> start(n + 1);
  std::vector<uint32_t> owner;
  for (uint64_t i = 0; i < n; ++i) {
    start[i] = text.size();
    const uint64_t a = reads->off[i], l = reads->off[i + 1] - a;
    for (uint64_t j = 0; j < l; ++j) {
      const char ch = reads->ascii[reversed ? a + l - 1 - j : a + j];
      text.push_back((uint8_t)(ch == 'A' ? 1 : ch == 'C' ? 2 : ch == 'G' ? 3 : 4));
      owner.push_back((uint32_t)i);
    }
    text.push_back(0);
    owner.push_back((uint32_t)i);
  }
  start[n] = text.size();
  const uint64_t S = text.size();
  std::vector<uint64_t> sa(S);
  for (uint64_t q = 0; q < S; ++q) sa[q] = q;
  if (order == SGX_ORDER_SAIS2) {
    std::sort(sa.begin(), sa.end(), [&](uint64_t x, uint64_t y) {
      const uint64_t m = std::min(S - x, S - y);
      const int c = memcmp(&text[x], &text[y], m);
      return c != 0 ? c < 0 : x > y;  // the shorter suffix (end of text) first
    });
  } else {
    std::sort(sa.begin(), sa.end(), [&](uint64_t x, uint64_t y) {
      const uint64_t ex = start[owner[x] + 1], ey = start[owner[y] + 1];  // one past the read's own '$'
      const uint64_t m = std::min(ex - x, ey - y);
      const int c = memcmp(&text[x], &text[y], m);  // both strings end with their '$', so equal prefixes mean equal strings
      if (c != 0) return c < 0;
      if (ex - x != ey - y) return ex - x < ey - y;
      return owner[x] < owner[y];
    });
  }
  std::vector<uint8_t> rle;
  std::vector<uint32_t> sai;
  int run_sym = -1, run_len = 0;
  for (uint64_t k = 0; k < S; ++k) {
    const uint64_t q = sa[k];
    const bool at_start = q == start[owner[q]];
    const int sym = at_start ? 0 : text[q - 1];
    if (at_start) sai.push_back(owner[q]);
    if (sym == run_sym && run_len < 31) {
      ++run_len;
    } else {
      if (run_len) rle.push_back((uint8_t)((run_sym << 5) | run_len));
      run_sym = sym;
      run_len = 1;
    }
  }
  if (run_len) rle.push_back((uint8_t)((run_sym << 5) | run_len));
  memset(out, 0, sizeof(*out));
  out->n_strings = n;
  out->n_suffixes = S;
  out->n_runs = rle.size();
  out->seg_bytes = out->n_runs;
  out->rle = (uint8_t*)malloc(rle.size() ? rle.size() : 1);
  memcpy(out->rle, rle.data(), rle.size());
  out->sai = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
  memcpy(out->sai, sai.data(), n * sizeof(uint32_t));
  return SGX_OK;
}
void sgx_index_out_free(sgx_index_out* o);
int sgx_index_build_pair(sgx_ctx* ctx, const sgx_reads* reads, int order, sgx_index_out* out_fwd, sgx_index_out* out_rev) {
  int rc = sgx_index_build(ctx, reads, order, 0, out_fwd);
  if (rc == SGX_OK && (rc = sgx_index_build(ctx, reads, order, 1, out_rev)) != SGX_OK) sgx_index_out_free(out_fwd);
  return rc;
}

void sgx_index_out_free(sgx_index_out* o) {
  if (!o) return;
  free(o->rle);
  free(o->sai);
  o->rle = nullptr;
  o->sai = nullptr;
}
// Stand-in communicator: no transport at all -- every rank computes the whole index and keeps its share of it.  Enough to
// run the front-end's --gpus N code of `siga index` (id hand-out over pipes, all ranks of a strand group writing one .bwt
// together with pwrite) on CPU processes.
int sgx_comm_unique_id(uint8_t* id) {
  memset(id, 0x5a, 128);
  return SGX_OK;
}
int sgx_comm_init(sgx_ctx* ctx, const uint8_t* id, int rank, int world, sgx_comm** out) {
  if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world || id[0] != 0x5a) return SGX_EINVAL;
  *out = new sgx_comm{rank, world};
  return SGX_OK;
}
void sgx_comm_free(sgx_comm* c) { delete c; }
int sgx_index_build_sharded(sgx_ctx* ctx, sgx_comm* comm, const sgx_reads* reads, int order, int reversed, sgx_index_out* out) {
  if (!comm) return SGX_EINVAL;
  const int rc = sgx_index_build(ctx, reads, order, reversed, out);
  if (rc == SGX_OK && comm->rank != 0) {  // the run bytes are complete on rank 0 only
    out->seg_bytes = 0;
    out->rle[0] = 0;
  }
  return rc;
}
int sgx_index_build_sharded_segments(sgx_ctx* ctx, sgx_comm* comm, const sgx_reads* reads, int order, int reversed,
                                     sgx_index_out* out) {
  if (!comm) return SGX_EINVAL;
  const int rc = sgx_index_build(ctx, reads, order, reversed, out);
  if (rc != SGX_OK) return rc;
  // rank r keeps bytes [n r / W, n (r + 1) / W) of the run bytes (uneven on purpose when n is not a multiple of W)
  const uint64_t n = out->n_runs, a = n * (uint64_t)comm->rank / comm->world, b = n * (uint64_t)(comm->rank + 1) / comm->world;
  memmove(out->rle, out->rle + a, b - a);
  out->seg_offset = a;
  out->seg_bytes = b - a;
  return SGX_OK;
}

}  // extern "C"
