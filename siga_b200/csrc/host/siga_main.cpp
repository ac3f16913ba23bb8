// `siga` front-end of the B200 build: the reference's `index` and `overlap` sub-commands with the same options and
// the same SGA-compatible outputs, the work itself done by libsiga_b200 (include/siga_b200.h).
//
//   siga index   [-a sais|sais2] [-t N] [-p PREFIX] [--no-forward] [--no-reverse] [--device D] READS
//       mirrors Indexer::run (src/indexer.cpp:25-104): writes PREFIX.sai/.bwt (and .rsai/.rbwt for the reversed reads)
//   siga overlap [-m LEN] [-t N] [--batch-size N] [-p PREFIX] [-x] [--no-opposite-strand] [--device D] READS
//       mirrors Overlapping::run + OverlapBuilder::build (src/overlap.cpp:24-54, src/overlap_builder.cpp:423-509):
//       loads PREFIX.bwt/.rbwt/.sai/.rsai, writes PREFIX.asqg.gz (HT, VT in input order, then ED)
//
// Option defaults are the reference's CODE defaults (min-overlap 10, algorithm sais2).
// -t/--threads only sizes the host-side output formatting/compression pool; -c/--log4cxx is accepted and ignored (there
// is no log4cxx here); -s/--ini FILE supplies options as `key=value` lines which the command line overrides.  The per-thread .hits.gz files of the reference are an internal detail
// nothing downstream reads and are not produced.
#include <fcntl.h>
#include <getopt.h>
#include <signal.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>
#include <strings.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/siga_b200.h"
#include "fastgz.h"
#include "seqio.h"

using namespace sgxhost;

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void info(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void info(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "INFO siga - ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}

int fail(sgx_ctx* ctx, const char* what, int rc) {
  fprintf(stderr, "ERROR siga - %s failed (%d): %s\n", what, rc, sgx_last_error(ctx));
  return -1;
}

// After sgx_reads_pack has refused the input: name the first read with a byte other than A/C/G/T.
// Success exit of a sub-command, called once its output files are closed and checked.  Leaving right here skips the
// destructors of the read set (millions of strings), the explicit release of device and pinned memory and the tear-down
// of the CUDA context by the runtime's exit handlers: a quarter of a second and more at config-2 size that frees nothing
// the end of the process does not free anyway.
[[noreturn]] void finish_ok() {
  fflush(nullptr);
  _exit(0);
}

void explain_bad_bases(const ReadSet& rs, uint64_t first = 0, uint64_t count = (uint64_t)-1) {
  const uint64_t end = count == (uint64_t)-1 ? rs.size() : first + count;
  for (uint64_t i = first; i < end; ++i)
    for (uint64_t k = rs.off[i]; k < rs.off[i + 1]; ++k) {
      const char ch = rs.bases[k];
      if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') {
        fprintf(stderr, "ERROR siga - read %llu (%s) has '%c' at base %llu: reads must hold A/C/G/T only (run them through "
                        "`siga preprocess` first)\n", (unsigned long long)i, rs.names[i].c_str(), ch, (unsigned long long)(k - rs.off[i]));
        return;
      }
    }
}

// The worker processes of a `--gpus N` run (rank 0 is the calling process, ranks 1..N-1 are forked before any CUDA call).
// A rank that fails leaves through _exit(1) while its peers may sit inside an NCCL collective that will never complete,
// so the parent watches its children: the first one that ends badly gets the others killed, the part files removed and
// the parent itself out with an error, wherever its main thread is blocked.  The same clean-up runs when the parent
// returns early on an error of its own.  Children die with the parent (PR_SET_PDEATHSIG).
class Gang {
 public:
  ~Gang() { abort_all(); }
  // Forks ranks 1..world-1; returns this process's rank (0 in the parent), -1 when a fork failed.
  int spawn(int world) {
    fflush(nullptr);
    const pid_t parent = getpid();
    for (int r = 1; r < world; ++r) {
      const pid_t pid = fork();
      if (pid < 0) {
        fprintf(stderr, "ERROR siga - fork failed\n");
        abort_all();
        return -1;
      }
      if (pid == 0) {
        kids_.clear();  // not ours
        prctl(PR_SET_PDEATHSIG, SIGKILL);
        if (getppid() != parent) _exit(1);  // the parent went away between fork and prctl
        return r;
      }
      kids_.push_back(pid);
    }
    done_.assign(kids_.size(), 0);
    if (!kids_.empty()) monitor_ = std::thread([this]() { watch(); });
    return 0;
  }
  void add_temp(const std::string& path) {
    std::lock_guard<std::mutex> lk(m_);
    temps_.push_back(path);
  }
  // Parent: waits for every child; true when all of them exited with status 0.
  bool wait_all() {
    if (monitor_.joinable()) monitor_.join();
    return !failed_;
  }
  void remove_temps() {
    std::lock_guard<std::mutex> lk(m_);
    for (const auto& t : temps_) remove(t.c_str());
    temps_.clear();
  }

 private:
  void watch() {
    size_t left = kids_.size();
    while (left > 0) {
      for (size_t i = 0; i < kids_.size(); ++i) {
        if (done_[i]) continue;
        int st = 0;
        std::unique_lock<std::mutex> lk(km_);  // a pid is signalled only while it is known not to be reaped
        const pid_t w = waitpid(kids_[i], &st, WNOHANG);
        if (w == 0) continue;
        done_[i] = 1;
        lk.unlock();
        --left;
        if (w < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) {
          failed_ = true;
          if (!aborting_) {
            // a worker failed while the others (and possibly this process's main thread) wait for it in a collective
            fprintf(stderr, "ERROR siga - GPU worker process %d failed; stopping the other workers\n", (int)kids_[i]);
            for (size_t k = 0; k < kids_.size(); ++k)
              if (!done_[k]) kill(kids_[k], SIGKILL);
            for (size_t k = 0; k < kids_.size(); ++k)
              if (!done_[k]) waitpid(kids_[k], &st, 0);
            remove_temps();
            fflush(nullptr);
            _exit(1);
          }
        }
      }
      if (left) std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
  }
  void abort_all() {
    aborting_ = true;
    if (monitor_.joinable()) {
      {
        std::lock_guard<std::mutex> lk(km_);
        for (size_t k = 0; k < kids_.size(); ++k)
          if (!done_[k]) kill(kids_[k], SIGKILL);
      }
      monitor_.join();
    }
    kids_.clear();
    remove_temps();
  }
  std::vector<pid_t> kids_;
  std::vector<char> done_;
  std::vector<std::string> temps_;
  std::thread monitor_;
  std::mutex m_, km_;
  std::atomic<bool> failed_{false}, aborting_{false};
};

// One gzip member per chunk, compressed on worker threads, written in order (multi-member gzip is what zlib's and
// boost's readers, hence `siga assemble`, accept).
class GzChunkWriter {
 public:
  GzChunkWriter(const std::string& path, unsigned threads, int level)
      : f_(fopen(path.c_str(), "wb")), threads_(threads ? threads : 1), level_(level) {}
  ~GzChunkWriter() {
    if (f_) fclose(f_);
  }
  bool ok() const { return f_ != NULL; }
  // Flushes and closes the file; false when the last buffered bytes could not be written (ENOSPC, EIO).  Callers fold
  // this into their result: a destructor cannot report it.
  bool close() {
    if (!f_) return false;
    const bool good = fflush(f_) == 0 && !ferror(f_);
    const bool closed = fclose(f_) == 0;
    f_ = NULL;
    return good && closed;
  }
  // produce(i, &text) fills the text of chunk i in [0, n); chunks are emitted in index order.  encoder(i) picks the
  // member encoder of chunk i: zlib at the configured level, or one of the two writers of fastgz.h (literal-only for
  // sequence records, greedy LZ77 for records with repeated fields).
  enum Encoder { kZlib = 0, kLiteral = 1, kFastLz = 2 };
  template <class F>
  bool write_chunks(size_t n, F produce) {
    return write_chunks(n, produce, [](size_t) { return kZlib; });
  }
  template <class F, class L>
  bool write_chunks(size_t n, F produce, L encoder) {
    // a pool of workers takes chunk indices in order; the calling thread writes finished members in index order while
    // later chunks are still being produced (at most 2 x threads chunks are in flight, which bounds the memory)
    const size_t window = 2 * (size_t)threads_;
    std::vector<std::string> slot(window);
    std::vector<char> ready(window, 0);
    std::mutex mu;
    std::condition_variable cv;
    size_t next = 0, written = 0;
    bool good = true;
    auto worker = [&]() {
      for (;;) {
        size_t i;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&]() { return !good || next >= n || next < written + window; });
          if (!good || next >= n) return;
          i = next++;
        }
        std::string text;
        produce(i, &text);
        const int enc = encoder(i);
        std::string gz = enc == kLiteral ? gzip_member_huffman(text.data(), text.size())
                         : enc == kFastLz ? gzip_member_lz(text.data(), text.size()) : deflate_member(text, level_);
        if (gz.empty() && enc != kZlib) gz = deflate_member(text, level_);  // the fast encoders return "" on an internal mismatch
        {
          std::lock_guard<std::mutex> lk(mu);
          if (gz.empty()) good = false;  // a gzip member is never empty (18 bytes of header + trailer at least)
          slot[i % window].swap(gz);
          ready[i % window] = 1;
        }
        cv.notify_all();
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads_ && t < n; ++t) pool.emplace_back(worker);
    while (written < n) {
      std::string gz;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return ready[written % window] != 0; });
        gz.swap(slot[written % window]);
        ready[written % window] = 0;
      }
      const bool ok_write = fwrite(gz.data(), 1, gz.size(), f_) == gz.size();
      {
        std::lock_guard<std::mutex> lk(mu);
        if (!ok_write) good = false;
        ++written;
      }
      cv.notify_all();
      if (!ok_write) break;
    }
    {
      std::lock_guard<std::mutex> lk(mu);
      if (written < n) good = false;  // stop the workers after a failed write
    }
    cv.notify_all();
    for (auto& th : pool) th.join();
    return good;
  }

 private:
  static std::string deflate_member(const std::string& text, int level) {
    // "" = failure (the caller treats an empty member as an error)
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (text.size() > 0x7fffffffu) return std::string();  // chunks are a few MB; avail_in is 32 bits
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return std::string();
    const uLong bound = deflateBound(&zs, (uLong)text.size()) + 64;
    if (bound > 0xffffffffu) {
      deflateEnd(&zs);
      return std::string();
    }
    std::string out(bound, '\0');
    zs.next_in = (Bytef*)text.data();
    zs.avail_in = (uInt)text.size();
    zs.next_out = (Bytef*)&out[0];
    zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, Z_FINISH);
    out.resize(rc == Z_STREAM_END ? zs.total_out : 0);
    deflateEnd(&zs);
    return out;
  }
  FILE* f_;
  unsigned threads_;
  int level_;
};

void append_uint(std::string* s, uint64_t v) {
  char tmp[24];
  char* p = tmp + sizeof(tmp);
  do {
    *--p = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  s->append(p, tmp + sizeof(tmp) - p);
}

// Decimal digits of v at p, returns the end; values below 10 000 (coordinates inside a read) take no division loop.
inline char* put_uint(char* p, uint64_t v) {
  static const char kPairs[201] =
      "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
      "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
  if (v < 10) {
    *p++ = (char)('0' + v);
    return p;
  }
  if (v < 100) {
    memcpy(p, kPairs + 2 * v, 2);
    return p + 2;
  }
  if (v < 1000) {
    *p++ = (char)('0' + v / 100);
    memcpy(p, kPairs + 2 * (v % 100), 2);
    return p + 2;
  }
  if (v < 10000) {
    memcpy(p, kPairs + 2 * (v / 100), 2);
    memcpy(p + 2, kPairs + 2 * (v % 100), 2);
    return p + 4;
  }
  char tmp[24];
  char* q = tmp + sizeof(tmp);
  do {
    *--q = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  const size_t n = (size_t)(tmp + sizeof(tmp) - q);
  memcpy(p, q, n);
  return p + n;
}

// CR:i / BX:Z / EX:Z tags carried in the read comment (src/overlap_builder.cpp:304-316, src/asqg.h:29-60).
void append_comment_tags(const std::string& comment, std::string* line) {
  if (comment.empty()) return;
  std::string cr, bx, ex;
  size_t b = 0;
  while (b <= comment.size()) {
    size_t e = comment.find(' ', b);
    if (e == std::string::npos) e = comment.size();
    const std::string tok = comment.substr(b, e - b);
    b = e + 1;
    if (tok.size() < 2) continue;
    const std::string key = tok.substr(0, 2);
    if (key != "CR" && key != "BX" && key != "EX") continue;
    // TagValue::fromstring: exactly KEY:type:value with the right type code
    size_t c1 = tok.find(':'), c2 = c1 == std::string::npos ? c1 : tok.find(':', c1 + 1);
    if (c1 == std::string::npos || c2 == std::string::npos || tok.find(':', c2 + 1) != std::string::npos) continue;
    if (c2 - c1 != 2) continue;
    const char type = tok[c1 + 1];
    const std::string val = tok.substr(c2 + 1);
    if (key == "CR") {
      if (type != 'i') continue;
      cr = "CR:i:" + std::to_string(atoi(val.c_str()));
    } else {
      if (type != 'Z') continue;
      (key == "BX" ? bx : ex) = key + ":Z:" + val;
    }
  }
  // writer order: SS, CR, BX, EX (src/asqg.cpp:171-186)
  for (const std::string* t : {&cr, &bx, &ex}) {
    if (!t->empty()) {
      *line += '\t';
      *line += *t;
    }
  }
}

struct CommonOpts {
  std::string prefix;
  int threads = 0;     // host-side text / gzip work; 0 = all hardware threads (the search itself runs on the GPU)
  int device = 0;
  int gzip_level = 1;  // zlib level of the .gz members that go through zlib
  bool gzip_fast = true;  // sequence records (VT lines, corrected reads) use the literal-only encoder and ED records the
                          // greedy-LZ77 one (fastgz.h); --gzip-level N sends everything through zlib at level N
  int gpus = 1;        // --gpus N: one forked process per GPU (devices device .. device + N - 1)
  bool help = false;
};

// Index files and reads of `overlap` / `rmdup`, loaded concurrently with the CUDA context start-up.
struct LoadedInput {
  uint64_t ns = 0, nf = 0, rns = 0, rnf = 0;
  std::vector<uint8_t> runs, rruns;
  std::vector<uint32_t> sai, rsai;
  ReadSet rs;
  std::string err;
  sgx_ctx* ctx = nullptr;
  int ctx_rc = SGX_OK;
  bool ok_bwt = false, ok_rbwt = false, ok_sai = false, ok_rsai = false, ok_reads = false;
  double t_part[4] = {0, 0, 0, 0};  // seconds of the concurrent loaders: CUDA start, .bwt, .sai, reads
};

bool load_input(const std::string& input, const std::string& output, int device, LoadedInput* in, bool with_ctx = true) {
  std::vector<std::thread> pool;
  auto timed = [in](int slot, const std::function<void()>& f) {
    const double t = now_s();
    f();
    in->t_part[slot] = now_s() - t;
  };
  if (with_ctx) pool.emplace_back([&]() { timed(0, [&]() { in->ctx_rc = sgx_init(device, &in->ctx); }); });
  pool.emplace_back([&]() { timed(1, [&]() { in->ok_bwt = read_bwt(output + ".bwt", &in->ns, &in->nf, &in->runs); }); });
  pool.emplace_back([&]() { in->ok_rbwt = read_bwt(output + ".rbwt", &in->rns, &in->rnf, &in->rruns); });
  pool.emplace_back([&]() { timed(2, [&]() { in->ok_sai = read_sai(output + ".sai", &in->sai); }); });
  pool.emplace_back([&]() { in->ok_rsai = read_sai(output + ".rsai", &in->rsai); });
  pool.emplace_back([&]() { timed(3, [&]() { in->ok_reads = read_sequences(input, &in->rs, &in->err); }); });
  for (auto& t : pool) t.join();
  info("load (concurrent): CUDA start %.3f s, .bwt %.3f s, .sai %.3f s, reads %.3f s", in->t_part[0], in->t_part[1], in->t_part[2],
       in->t_part[3]);
  if (!in->ok_bwt || !in->ok_rbwt) {
    fprintf(stderr, "ERROR siga - Failed to load FMIndex from %s\n", input.c_str());
    return false;
  }
  if (!in->ok_sai || !in->ok_rsai) {
    fprintf(stderr, "ERROR siga - failed to load suffix array index %s\n", output.c_str());
    return false;
  }
  if (!in->ok_reads) {
    fprintf(stderr, "ERROR siga - %s\n", in->err.c_str());
    return false;
  }
  if (in->rs.size() != in->ns || in->rs.size() != in->sai.size() || in->rs.size() != in->rsai.size()) {
    fprintf(stderr, "ERROR siga - %s holds %zu reads but the index was built from %llu\n", input.c_str(), in->rs.size(),
            (unsigned long long)in->ns);
    return false;
  }
  if (in->ctx_rc != SGX_OK) {
    fail(nullptr, "sgx_init", in->ctx_rc);
    return false;
  }
  return true;
}

// Writes the HT line / the VT lines of reads [first, first + ed.n_reads) / the ED lines of `ed` as gzip members.
// One .asqg.gz = HT, all VT (input order), all ED (src/bigraph.cpp:532-616 requires this order); concatenated gzip
// members are a valid gzip file, which is how the multi-GPU run assembles its output from per-rank parts.
bool write_asqg(const std::string& path, const CommonOpts& o, const ReadSet& rs, const sgx_edges& ed, uint64_t first,
                unsigned min_overlap, bool with_ht, bool with_vt, bool with_ed) {
  GzChunkWriter gz(path, (unsigned)o.threads, o.gzip_level);
  if (!gz.ok()) return false;
  const size_t n = with_vt ? (size_t)ed.n_reads : 0;
  const size_t vt_chunk = 1 << 16, ed_chunk = 1 << 18;
  const size_t n_ht = with_ht ? 1 : 0;
  const size_t n_vt = (n + vt_chunk - 1) / vt_chunk, n_ed = with_ed ? (ed.n_edges + ed_chunk - 1) / ed_chunk : 0;
  // worst-case bytes of one ED record: "ED\t" + two names + separator + six numbers no longer than the longest read's
  // length (coordinates inside a read) + the flags
  size_t ed_bound = 0, max_name = 0;
  if (n_ed) {
    uint64_t max_len = 0;
    for (size_t i = 0; i < rs.size(); ++i) {
      max_name = std::max(max_name, rs.names[i].size());
      max_len = std::max<uint64_t>(max_len, rs.off[i + 1] - rs.off[i]);
    }
    char digits[24];
    const size_t nd = (size_t)(put_uint(digits, max_len) - digits);
    ed_bound = 3 + 2 * max_name + 1 + 6 * (1 + nd) + 5;
  }
  const bool wrote = gz.write_chunks(n_ht + n_vt + n_ed, [&](size_t ci, std::string* text) {
    if (ci < n_ht) {
      *text = "HT\tVN:i:1\tOL:i:" + std::to_string(min_overlap) + "\tCN:i:1\n";  // IN/ER/TE are never set on this path
    } else if (ci < n_ht + n_vt) {
      const size_t b = (ci - n_ht) * vt_chunk, e = std::min(n, b + vt_chunk);
      text->reserve((e - b) * 200);
      for (size_t k = b; k < e; ++k) {
        const size_t i = first + k;
        *text += "VT\t";
        *text += rs.names[i];
        *text += '\t';
        text->append(rs.bases.data() + rs.off[i], rs.off[i + 1] - rs.off[i]);
        *text += ed.substring[k] ? "\tSS:i:1" : "\tSS:i:0";
        append_comment_tags(rs.comments[i], text);
        *text += '\n';
      }
    } else {
      const size_t b = (ci - n_ht - n_vt) * ed_chunk, e = std::min<size_t>(ed.n_edges, b + ed_chunk);
      // Written through a raw pointer into a buffer sized for the worst case: the target's name and length are two
      // random accesses per record (prefetched a few records ahead), everything else is digits.
      const size_t any_record = 3 + 2 * max_name + 1 + 6 * 21 + 5;  // whatever the six numbers are
      text->resize((e - b) * ed_bound + any_record);
      char* p0 = &(*text)[0];
      char* p = p0;
      const size_t ahead = 16;
      for (size_t k = b; k < e; ++k) {
        // never true for edges whose coordinates lie inside their reads (every record then fits ed_bound); an edge that
        // broke that promise gets room instead of running over the buffer
        if ((size_t)(p - p0) + any_record > text->size()) {
          const size_t at = (size_t)(p - p0);
          text->resize(text->size() * 2);
          p0 = &(*text)[0];
          p = p0 + at;
        }
        if (k + ahead < e) {
          const uint32_t t = ed.edges[k + ahead].target;
          __builtin_prefetch(&rs.names[t]);
          __builtin_prefetch(&rs.off[t]);
        }
        const sgx_edge& g = ed.edges[k];
        const uint64_t lq = rs.off[g.query + 1] - rs.off[g.query], lt = rs.off[g.target + 1] - rs.off[g.target];
        uint64_t s0 = lq - g.length, e0 = lq - 1, s1 = 0, e1 = g.length - 1;
        if (g.af & SGX_AF_QUERYREV) {
          const uint64_t t = s0;
          s0 = lq - 1 - e0;
          e0 = lq - 1 - t;
        }
        if (g.af & SGX_AF_TARGETREV) {
          const uint64_t t = s1;
          s1 = lt - 1 - e1;
          e1 = lt - 1 - t;
        }
        memcpy(p, "ED\t", 3);
        p += 3;
        const std::string& qn = rs.names[g.query];
        memcpy(p, qn.data(), qn.size());
        p += qn.size();
        *p++ = ' ';
        const std::string& tn = rs.names[g.target];
        memcpy(p, tn.data(), tn.size());
        p += tn.size();
        const uint64_t v[6] = {s0, e0, lq, s1, e1, lt};
        for (int q = 0; q < 6; ++q) {
          *p++ = ' ';
          p = put_uint(p, v[q]);
        }
        memcpy(p, (g.af & SGX_AF_QUERYCOMP) ? " 1 0\n" : " 0 0\n", 5);
        p += 5;
      }
      text->resize((size_t)(p - p0));
    }
  }, [&](size_t ci) {  // VT records are mostly read sequence; HT / ED records repeat their fields
    if (!o.gzip_fast) return GzChunkWriter::kZlib;
    return ci >= n_ht && ci < n_ht + n_vt ? GzChunkWriter::kLiteral : GzChunkWriter::kFastLz;
  });
  return gz.close() && wrote;
}

bool append_file(FILE* dst, const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<char> buf(1 << 22);
  bool ok = true;
  size_t k;
  while (ok && (k = fread(buf.data(), 1, buf.size(), f)) > 0) ok = fwrite(buf.data(), 1, k, dst) == k;
  fclose(f);
  return ok;
}

// Writes the strings back to back into `path`, every string by its own thread at its own offset (the copy into the page
// cache is what a single writer spends its time on at hundreds of MB).
bool write_parts_parallel(const std::string& path, const std::vector<std::string>& parts) {
  const int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return false;
  std::vector<uint64_t> at(parts.size() + 1, 0);
  for (size_t i = 0; i < parts.size(); ++i) at[i + 1] = at[i] + parts[i].size();
  std::vector<char> good(parts.size(), 1);
  std::vector<std::thread> pool;
  for (size_t i = 0; i < parts.size(); ++i)
    pool.emplace_back([&, i]() {
      const char* p = parts[i].data();
      size_t left = parts[i].size();
      uint64_t off = at[i];
      while (left > 0) {
        const ssize_t k = pwrite(fd, p, left, (off_t)off);
        if (k <= 0) {
          good[i] = 0;
          return;
        }
        p += k;
        left -= (size_t)k;
        off += (uint64_t)k;
      }
    });
  for (auto& th : pool) th.join();
  bool ok = close(fd) == 0;
  for (char g : good) ok = ok && g;
  return ok;
}

int usage_index() {
  printf("siga index [OPTION] ... READSFILE\n"
         "Index the reads in READSFILE using a suffixarray/bwt (built on the GPU)\n\n"
         "      -h, --help                       display this help and exit\n\n"
         "      -a, --algorithm=STR              suffix ordering of the BWT. STR can be:\n"
         "                                       sais  - equal suffixes ordered by read index\n"
         "                                       sais2 - plain suffix order of the concatenated reads (default)\n"
         "      -t, --threads=NUM                host threads for output formatting (default: 1)\n"
         "      -p, --prefix=PREFIX              write index to file using PREFIX instead of prefix of READSFILE\n"
         "          --no-forward                 suppress construction of the forward BWT.\n"
         "          --no-reverse                 suppress construction of the reverse BWT.\n"
         "          --gpus=N                     build each index with N GPUs together (rows sharded by leading k-mer, NCCL)\n"
         "          --device=NUM                 CUDA device (default: 0)\n\n");
  return 256;
}

int usage_overlap() {
  printf("siga overlap [OPTION] ... READSFILE\n"
         "Compute pairwise overlap between all the sequences in READS (searched on the GPU)\n\n"
         "      -h, --help                       display this help and exit\n\n"
         "      -t, --threads=NUM                host threads for output formatting/compression (default: all)\n"
         "          --gzip-level=N               zlib level of the .asqg.gz members, 1..9 (default: 1)\n"
         "          --gpus=N                     shard the reads over N GPUs (devices --device .. --device + N - 1), one process each\n"
         "          --batch-size=NUM             accepted for compatibility (batching is internal)\n"
         "      -m, --min-overlap=LEN            minimum overlap required between two reads (default: 10)\n"
         "      -p, --prefix=PREFIX              read the index from / write the graph to PREFIX instead of prefix of READSFILE\n"
         "      -x, --exhaustive                 output all overlaps, including transitive edges\n"
         "          --no-opposite-strand         treat all reads as forward strand\n"
         "          --device=NUM                 CUDA device (default: 0)\n\n");
  return 256;
}

int cmd_index(int argc, char** argv) {
  enum { OPT_NO_REVERSE = 1000, OPT_NO_FORWARD, OPT_DEVICE, OPT_GPUS };
  static const option longopts[] = {{"log4cxx", required_argument, nullptr, 'c'}, {"ini", required_argument, nullptr, 's'},
                                    {"prefix", required_argument, nullptr, 'p'},  {"threads", required_argument, nullptr, 't'},
                                    {"algorithm", required_argument, nullptr, 'a'}, {"no-reverse", no_argument, nullptr, OPT_NO_REVERSE},
                                    {"no-forward", no_argument, nullptr, OPT_NO_FORWARD}, {"device", required_argument, nullptr, OPT_DEVICE},
                                    {"gpus", required_argument, nullptr, OPT_GPUS},
                                    {"help", no_argument, nullptr, 'h'},          {nullptr, 0, nullptr, 0}};
  CommonOpts o;
  std::string algorithm = "sais2";
  bool fwd = true, rev = true;
  int c;
  while ((c = getopt_long(argc, argv, "c:s:a:t:p:g:h", longopts, nullptr)) != -1) {
    switch (c) {
      case 'a': algorithm = optarg; break;
      case 't': o.threads = atoi(optarg); break;
      case 'p': o.prefix = optarg; break;
      case OPT_NO_REVERSE: rev = false; break;
      case OPT_NO_FORWARD: fwd = false; break;
      case OPT_GPUS: o.gpus = std::max(1, std::min(16, atoi(optarg))); break;
      case OPT_DEVICE: o.device = atoi(optarg); break;
      case 'h': o.help = true; break;
      default: break;
    }
  }
  if (o.help || argc - optind != 1) return usage_index();
  const std::string input = argv[optind];
  const std::string output = o.prefix.empty() ? stem(input) : o.prefix;
  info("input: %s", input.c_str());
  info("output: %s.(.sai|.bwt|.rsai|.rbwt)", output.c_str());
  info("algorithm: %s", algorithm.c_str());
  int order;
  if (strcasecmp(algorithm.c_str(), "sais") == 0) order = SGX_ORDER_SAIS;
  else if (strcasecmp(algorithm.c_str(), "sais2") == 0) order = SGX_ORDER_SAIS2;
  else {
    // the reference accepts "rope" and silently writes nothing; refuse instead
    fprintf(stderr, "ERROR siga - Failed to create suffix array builder algorithm %s\n", algorithm.c_str());
    return -1;
  }
  const double t0 = now_s();
  ReadSet rs;
  std::string err;
  info("loading reads: %s", input.c_str());
  sgx_ctx* ctx = nullptr;
  int rc_init = SGX_OK;
  int rc = SGX_OK;
  int world = o.gpus;
  int my = 0;
  int group = -1, sub_world = 1, sub_rank = 0;  // --gpus N: strand group of this rank (see below)
  double t_cuda = 0;  // seconds inside sgx_init (driver + context start-up)
  Gang gang;
  sgx_comm* comm = nullptr;
  if (world == 1) {
    std::thread init_thread([&]() {  // CUDA start-up overlaps the parse
      const double tc = now_s();
      rc_init = sgx_init(o.device, &ctx);
      t_cuda = now_s() - tc;
    });
    const bool parsed = read_sequences(input, &rs, &err);
    init_thread.join();
    if (!parsed) {
      fprintf(stderr, "ERROR siga - %s\n", err.c_str());
      return -1;
    }
    if (rc_init != SGX_OK) return fail(nullptr, "sgx_init", rc_init);
  } else {
    // --gpus N: parse once, fork one process per GPU before any CUDA call, carry rank 0's NCCL id through pipes; all
    // ranks then build each index together (rows sharded by leading k-mer bucket) and rank 0 writes the files.
    if (!read_sequences(input, &rs, &err)) {
      fprintf(stderr, "ERROR siga - %s\n", err.c_str());
      return -1;
    }
    // A strand is sharded over several GPUs only when it is big: a column pass of a config-2-sized build (463 M rows)
    // takes 0.3 ms on one GPU, less than the collectives of a sharded pass cost.  Below the
    // threshold one GPU per wanted strand does the work and the other devices are not used.
    {
      uint64_t min_rows = (uint64_t)1 << 31;
      if (const char* e = getenv("SIGA_SHARD_MIN_ROWS")) min_rows = strtoull(e, nullptr, 10);
      const uint64_t rows = rs.bases.size() + rs.size();
      const int strands = (fwd ? 1 : 0) + (rev ? 1 : 0);
      if (rows < min_rows && world > strands) {
        info("%llu rows per index: one GPU per strand (sharding starts at %llu rows), using %d of the %d GPUs",
             (unsigned long long)rows, (unsigned long long)min_rows, strands, world);
        world = strands;
      }
    }
    std::vector<int> rd(world, -1), wr(world, -1);
    for (int r = 1; r < world; ++r) {
      int fds[2];
      if (pipe(fds) != 0) {
        fprintf(stderr, "ERROR siga - pipe failed\n");
        return -1;
      }
      rd[r] = fds[0];
      wr[r] = fds[1];
    }
    if ((my = gang.spawn(world)) < 0) return -1;
    if ((rc_init = sgx_init(o.device + my, &ctx)) != SGX_OK) {
      fail(nullptr, "sgx_init", rc_init);
      if (my) _exit(1);
      return -1;
    }
    // Strand groups: the two indices are independent jobs, so with both wanted and an even number of GPUs the ranks split
    // into one group per strand (ranks [0, N/2) forward, [N/2, N) reverse) -- the first factor of two costs no
    // communication -- and only inside a group of more than one GPU are the rows sharded over NCCL.  Each group's first
    // rank holds the finished index and writes its two files.
    if (fwd && rev && world % 2 == 0) {
      sub_world = world / 2;
      group = my / sub_world;
      sub_rank = my % sub_world;
    } else {
      sub_world = world;
      group = -1;  // one group builds the wanted strands one after the other
      sub_rank = my;
    }
    if (sub_world > 1) {
      const int leader = group > 0 ? sub_world : 0;
      uint8_t id[128];
      bool ok_id = true;
      if (my == leader) {
        ok_id = sgx_comm_unique_id(id) == SGX_OK;
        for (int r = leader + 1; r < leader + sub_world; ++r) ok_id = write(wr[r], id, sizeof(id)) == (ssize_t)sizeof(id) && ok_id;
      } else {
        ok_id = read(rd[my], id, sizeof(id)) == (ssize_t)sizeof(id);
      }
      if (!ok_id || (rc = sgx_comm_init(ctx, id, sub_rank, sub_world, &comm)) != SGX_OK) {
        fail(ok_id ? ctx : nullptr, "sgx_comm_init", rc);
        if (my) _exit(1);
        return -1;
      }
    }
  }
  const double t_load = now_s();
  sgx_reads* reads = nullptr;
  rc = sgx_reads_pack(ctx, rs.bases.data(), rs.off.data(), rs.size(), &reads);
  if (rc != SGX_OK) {
    fail(ctx, "sgx_reads_pack", rc);
    if (rc == SGX_EINVAL && my == 0) explain_bad_bases(rs);
    if (my) _exit(1);
    return -1;
  }
  // the files of one strand are written on a host thread while the device builds the other strand; on one GPU with both
  // strands wanted, the two builds run side by side on the device (sgx_index_build_pair) and all four files are written
  // at once
  sgx_index_out outs[2];
  std::thread writers[2];
  bool wrote[2] = {true, true};
  auto start_writer = [&](int pass) {
    sgx_index_out& out = outs[pass];
    const std::string sa = output + (pass == 0 ? ".sai" : ".rsai"), bw = output + (pass == 0 ? ".bwt" : ".rbwt");
    info("%s: %llu strings, %llu suffixes, %llu runs; device %.1f ms (read order %.1f, column passes %.1f, run-length coding + copy "
         "out %.1f; %llu kernel launches)", bw.c_str(), (unsigned long long)out.n_strings, (unsigned long long)out.n_suffixes,
         (unsigned long long)out.n_runs, out.ms_total, out.ms_sort, out.ms_bcr, out.ms_emit, (unsigned long long)out.launches);
    writers[pass] = std::thread([&outs, &wrote, pass, sa, bw]() {
      const sgx_index_out& o2 = outs[pass];
      bool a = false, b = false;
      std::thread t([&]() { a = write_sai(sa, o2.sai, o2.n_strings); });
      b = write_bwt(bw, o2.n_strings, o2.n_suffixes, o2.rle, o2.n_runs);
      t.join();
      wrote[pass] = a && b;
    });
  };
  // sharded build: every GPU of the strand group writes the run bytes of its own shard into the one .bwt file
  // (sgx_index_build_sharded_segments: nothing is gathered); the group's first rank adds the header and the .sai
  auto start_segment_writer = [&](int pass, bool leader) {
    sgx_index_out& out = outs[pass];
    const std::string sa = output + (pass == 0 ? ".sai" : ".rsai"), bw = output + (pass == 0 ? ".bwt" : ".rbwt");
    if (leader)
      info("%s: %llu strings, %llu suffixes, %llu runs; device %.1f ms (read order %.1f, column passes %.1f, run-length coding + copy "
           "out %.1f; %llu kernel launches)", bw.c_str(), (unsigned long long)out.n_strings, (unsigned long long)out.n_suffixes,
           (unsigned long long)out.n_runs, out.ms_total, out.ms_sort, out.ms_bcr, out.ms_emit, (unsigned long long)out.launches);
    writers[pass] = std::thread([&outs, &wrote, pass, sa, bw, leader]() {
      const sgx_index_out& o2 = outs[pass];
      bool a = true, b = false;
      std::thread t([&]() {
        if (leader) a = write_sai(sa, o2.sai, o2.n_strings);
      });
      b = write_bwt_segment(bw, o2.n_strings, o2.n_suffixes, o2.n_runs, o2.rle, o2.seg_offset, o2.seg_bytes, leader);
      t.join();
      wrote[pass] = a && b;
    });
  };
  if (world == 1 && fwd && rev) {
    info("building forward and reverse index side by side on device %d", o.device);
    rc = sgx_index_build_pair(ctx, reads, order, &outs[0], &outs[1]);
    if (rc != SGX_OK) return fail(ctx, "sgx_index_build_pair", rc);
    start_writer(0);
    start_writer(1);
  } else {
    for (int pass = 0; pass < 2; ++pass) {
      if ((pass == 0 && !fwd) || (pass == 1 && !rev)) continue;
      if (group >= 0 && pass != group) continue;  // the other strand group builds this one
      if (sub_rank == 0)
        info("building %s index on %d GPU(s) from device %d", pass == 0 ? "forward" : "reverse", sub_world, o.device + my);
      sgx_index_out& out = outs[pass];
      rc = comm ? sgx_index_build_sharded_segments(ctx, comm, reads, order, pass, &out) : sgx_index_build(ctx, reads, order, pass, &out);
      if (rc != SGX_OK) {
        for (auto& w : writers)
          if (w.joinable()) w.join();
        fail(ctx, comm ? "sgx_index_build_sharded_segments" : "sgx_index_build", rc);
        if (my) _exit(1);
        return -1;
      }
      if (comm) {
        start_segment_writer(pass, sub_rank == 0);
        continue;
      }
      if (sub_rank) {  // a strand below the sharding threshold is built by its group's first rank alone
        sgx_index_out_free(&out);
        continue;
      }
      start_writer(pass);
    }
  }
  const double t_built = now_s();
  for (int pass = 0; pass < 2; ++pass) {
    if (!writers[pass].joinable()) continue;
    writers[pass].join();
    sgx_index_out_free(&outs[pass]);
    if (!wrote[pass]) {
      fprintf(stderr, "ERROR siga - failed to write the %s index files of %s\n", pass == 0 ? "forward" : "reverse", output.c_str());
      if (my) _exit(1);
      return -1;
    }
  }
  if (comm) sgx_comm_free(comm);
  if (my) {
    sgx_reads_free(reads);
    // (the context goes with the process: see main)
    _exit(0);
  }
  const bool kids_ok = gang.wait_all();
  if (!kids_ok) {
    fprintf(stderr, "ERROR siga - a GPU worker process failed\n");
    return -1;
  }
  const double t1 = now_s();
  info("wall: %.3f sec (load + CUDA start %.3f [CUDA start alone %.3f], pack + build %.3f, rest of the writes %.3f)", t1 - t0,
       t_load - t0, t_cuda, t_built - t_load, t1 - t_built);
  finish_ok();  // reads, context and host buffers go with the process
}

int cmd_overlap(int argc, char** argv) {
  enum { OPT_BATCH_SIZE = 1000, OPT_NO_RC, OPT_DEVICE, OPT_GZIP_LEVEL, OPT_GPUS };
  static const option longopts[] = {{"log4cxx", required_argument, nullptr, 'c'},   {"ini", required_argument, nullptr, 's'},
                                    {"prefix", required_argument, nullptr, 'p'},    {"threads", required_argument, nullptr, 't'},
                                    {"batch-size", required_argument, nullptr, OPT_BATCH_SIZE},
                                    {"min-overlap", required_argument, nullptr, 'm'}, {"exhaustive", no_argument, nullptr, 'x'},
                                    {"no-opposite-strand", no_argument, nullptr, OPT_NO_RC},
                                    {"device", required_argument, nullptr, OPT_DEVICE}, {"help", no_argument, nullptr, 'h'},
                                    {"gzip-level", required_argument, nullptr, OPT_GZIP_LEVEL},
                                    {"gpus", required_argument, nullptr, OPT_GPUS}, {nullptr, 0, nullptr, 0}};
  CommonOpts o;
  unsigned min_overlap = 10;
  bool exhaustive = false, rcomp = true;
  int c;
  while ((c = getopt_long(argc, argv, "c:s:t:p:m:xh", longopts, nullptr)) != -1) {
    switch (c) {
      case 't': o.threads = atoi(optarg); break;
      case 'p': o.prefix = optarg; break;
      case 'm': min_overlap = (unsigned)atoi(optarg); break;
      case 'x': exhaustive = true; break;
      case OPT_NO_RC: rcomp = false; break;
      case OPT_GZIP_LEVEL:
        o.gzip_level = std::max(1, std::min(9, atoi(optarg)));
        o.gzip_fast = false;
        break;
      case OPT_GPUS: o.gpus = std::max(1, std::min(16, atoi(optarg))); break;
      case OPT_DEVICE: o.device = atoi(optarg); break;
      case 'h': o.help = true; break;
      default: break;
    }
  }
  if (o.help || argc - optind != 1) return usage_overlap();
  if (o.threads <= 0) o.threads = (int)std::thread::hardware_concurrency();
  const std::string input = argv[optind];
  const std::string output = o.prefix.empty() ? stem(input) : o.prefix;
  info("input: %s", input.c_str());
  info("output: %s.asqg.gz", output.c_str());
  const double t0 = now_s();

  // --gpus N: the index files and reads are loaded once, then one process per GPU is forked (before any CUDA call);
  // rank r searches the contiguous read range r of N against its own copy of both FM-indices and writes the VT / ED
  // members of its range into part files, which the parent concatenates (HT, all VT in input order, all ED).
  const int world = o.gpus;
  LoadedInput in;
  if (!load_input(input, output, o.device, &in, /*with_ctx=*/world == 1)) return -1;
  ReadSet& rs = in.rs;
  std::vector<uint8_t>&runs = in.runs, &rruns = in.rruns;
  std::vector<uint32_t>&sai = in.sai, &rsai = in.rsai;
  const uint64_t nf = in.nf, rnf = in.rnf;
  std::vector<uint32_t> rank;
  int my = 0;
  Gang gang;
  if (world > 1) {
    rank = name_ranks(rs.names);  // before the fork: every rank needs it
    if ((my = gang.spawn(world)) < 0) return -1;
    if (my == 0)
      for (int r = 0; r < world; ++r)
        for (const char* ext : {".vt", ".ed"}) gang.add_temp(output + ".asqg.gz.part" + std::to_string(r) + ext);
    if ((in.ctx_rc = sgx_init(o.device + my, &in.ctx)) != SGX_OK) {
      fail(nullptr, "sgx_init", in.ctx_rc);
      if (my) _exit(1);
      return -1;
    }
    o.threads = std::max(1, o.threads / world);
  }
  sgx_ctx* ctx = in.ctx;
  int rc = SGX_OK;
  const double t_load = now_s();
  const uint64_t first = rs.size() * (uint64_t)my / world, count = rs.size() * (uint64_t)(my + 1) / world - first;

  std::thread rank_thread([&]() {
    if (rank.empty()) rank = name_ranks(rs.names);  // host sort of the names while the device loads the index
  });
  struct Joiner {
    std::thread& t;
    ~Joiner() {
      if (t.joinable()) t.join();
    }
  } rank_joiner{rank_thread};
  sgx_reads* reads = nullptr;
  sgx_fm *fwd = nullptr, *rev = nullptr;
  // a failing worker rank leaves through _exit so that the parent's watcher sees a non-zero status
  auto bail = [&](const char* what, int code) {
    fail(ctx, what, code);
    if (my) {
      if (rank_thread.joinable()) rank_thread.join();
      _exit(1);
    }
    return -1;
  };
  const double tp0 = now_s();
  if ((rc = sgx_reads_pack(ctx, rs.bases.data(), rs.off.data(), rs.size(), &reads)) != SGX_OK) {
    if (rc == SGX_EINVAL && my == 0) explain_bad_bases(rs);
    return bail("sgx_reads_pack", rc);
  }
  const double tp1 = now_s();
  if ((rc = sgx_fm_create(ctx, runs.data(), runs.size(), nf, &fwd)) != SGX_OK) return bail("sgx_fm_create", rc);
  if ((rc = sgx_fm_create(ctx, rruns.data(), rruns.size(), rnf, &rev)) != SGX_OK) return bail("sgx_fm_create", rc);
  const double tp2 = now_s();
  if ((rc = sgx_fm_attach_reads(ctx, fwd, sai.data(), sai.size(), reads)) != SGX_OK) return bail("sgx_fm_attach_reads", rc);
  if ((rc = sgx_fm_attach_reads(ctx, rev, rsai.data(), rsai.size(), reads)) != SGX_OK) return bail("sgx_fm_attach_reads", rc);
  const double tp3 = now_s();
  info("device set-up: reads %.3f s, rank blocks of both indices %.3f s, .sai attach %.3f s", tp1 - tp0, tp2 - tp1, tp3 - tp2);
  info("fm-index: %llu symbols in %zu runs (%.4f symbols per run)", (unsigned long long)nf, runs.size(),
       runs.empty() ? 0.0 : (double)nf / runs.size());
  sgx_edges ed;
  rank_thread.join();
  rc = sgx_overlap_edges(ctx, fwd, rev, reads, first, count, min_overlap, exhaustive ? 0 : 1, rcomp ? 1 : 0, rank.data(), &ed);
  if (rc != SGX_OK) {
    fail(ctx, "sgx_overlap_edges", rc);
    if (my) _exit(1);
    return -1;
  }
  const double t_search = now_s();
  info("processed %llu sequences: %llu blocks, %llu edges; device %.1f ms (search %.1f ms, post %.1f ms; %llu rank lookups)",
       (unsigned long long)count, (unsigned long long)ed.n_blocks, (unsigned long long)ed.n_edges, ed.ms_total, ed.ms_search, ed.ms_extract,
       (unsigned long long)(ed.lookups_search + ed.lookups_extract));

  // ---- ASQG: HT, all VT (input order), all ED (src/bigraph.cpp:532-616 requires this order)
  const std::string final_path = output + ".asqg.gz";
  bool good;
  if (world == 1) {
    good = write_asqg(final_path, o, rs, ed, 0, min_overlap, true, true, true);
  } else {
    const std::string part = final_path + ".part" + std::to_string(my);
    good = write_asqg(part + ".vt", o, rs, ed, first, min_overlap, my == 0, true, false) &&
           write_asqg(part + ".ed", o, rs, ed, first, min_overlap, false, false, true);
    if (my) {
      sgx_edges_free(&ed);
      // (the context goes with the process: see main)
      _exit(good ? 0 : 1);
    }
    if (!gang.wait_all()) good = false;
    if (good) {
      FILE* f = fopen(final_path.c_str(), "wb");
      good = f != nullptr;
      for (int pass = 0; pass < 2 && good; ++pass)
        for (int r = 0; r < world && good; ++r)
          good = append_file(f, final_path + ".part" + std::to_string(r) + (pass == 0 ? ".vt" : ".ed"));
      if (f && fclose(f) != 0) good = false;
    }
    gang.remove_temps();
  }
  if (!good) {
    fprintf(stderr, "ERROR siga - write to %s.asqg.gz failed\n", output.c_str());
    return -1;
  }
  const double t1 = now_s();
  info("wall: %.3f sec (load %.3f, gpu %.3f, write %.3f)", t1 - t0, t_load - t0, t_search - t_load, t1 - t_search);
  finish_ok();  // edges, indices, reads, context and host buffers go with the process
}

int usage_rmdup() {
  printf("siga rmdup [OPTION] ... READSFILE\n"
         "Remove duplicated reads from the data set\n"
         "\n"
         "      -h, --help                       display this help and exit\n"
         "      -p, --prefix=PREFIX              use PREFIX instead of the prefix of the reads filename for the input/output files\n"
         "      -t, --threads=N                  accepted for compatibility (the search runs on the GPU)\n"
         "      -d, --sample-rate=N              accepted for compatibility (the rank layout is fixed: one 128-byte block per 256 symbols)\n"
         "          --device=N                   CUDA device (default 0)\n"
         "\n");
  return 256;
}

// `siga rmdup` (src/rmdup.cpp:23-54, OverlapBuilder::rmdup / duplicate, src/overlap_builder.cpp:623-704, 1184-1195):
// every read is searched at min-overlap = |read|, which leaves its containment blocks; the two searches duplicate()
// runs are the seq / forward-index one (AlignFlags 0) and the complement / reverse-index one (TARGETREV|QUERYCOMP).
// Output: PREFIX.rmdup.fa and PREFIX.rmdup.dups.fa, records as Hits2FastaConverter::convert writes them (563-616);
// record order follows the input (the reference's follows its per-thread hits files).
int cmd_rmdup(int argc, char** argv) {
  enum { OPT_DEVICE = 1000 };
  static const option longopts[] = {{"log4cxx", required_argument, nullptr, 'c'}, {"ini", required_argument, nullptr, 's'},
                                    {"prefix", required_argument, nullptr, 'p'},  {"threads", required_argument, nullptr, 't'},
                                    {"sample-rate", required_argument, nullptr, 'd'}, {"device", required_argument, nullptr, OPT_DEVICE},
                                    {"help", no_argument, nullptr, 'h'},          {nullptr, 0, nullptr, 0}};
  CommonOpts o;
  int c;
  while ((c = getopt_long(argc, argv, "c:s:t:p:d:h", longopts, nullptr)) != -1) {
    switch (c) {
      case 't': o.threads = atoi(optarg); break;
      case 'p': o.prefix = optarg; break;
      case OPT_DEVICE: o.device = atoi(optarg); break;
      case 'h': o.help = true; break;
      default: break;
    }
  }
  if (o.help || argc - optind != 1) return usage_rmdup();
  const std::string input = argv[optind];
  const std::string output = o.prefix.empty() ? stem(input) : o.prefix;
  info("input: %s", input.c_str());
  info("output: %s", output.c_str());
  const double t0 = now_s();
  LoadedInput in;
  if (!load_input(input, output, o.device, &in)) return -1;
  const double t_load = now_s();
  ReadSet& rs = in.rs;
  std::vector<uint8_t>&runs = in.runs, &rruns = in.rruns;
  std::vector<uint32_t>&sai = in.sai, &rsai = in.rsai;
  const uint64_t nf = in.nf, rnf = in.rnf;
  sgx_ctx* ctx = in.ctx;
  int rc = SGX_OK;
  sgx_reads* reads = nullptr;
  sgx_fm *fwd = nullptr, *rev = nullptr;
  if ((rc = sgx_reads_pack(ctx, rs.bases.data(), rs.off.data(), rs.size(), &reads)) != SGX_OK) {
    if (rc == SGX_EINVAL) explain_bad_bases(rs);
    return fail(ctx, "sgx_reads_pack", rc);
  }
  if ((rc = sgx_fm_create(ctx, runs.data(), runs.size(), nf, &fwd)) != SGX_OK) return fail(ctx, "sgx_fm_create", rc);
  if ((rc = sgx_fm_create(ctx, rruns.data(), rruns.size(), rnf, &rev)) != SGX_OK) return fail(ctx, "sgx_fm_create", rc);
  if ((rc = sgx_fm_attach_reads(ctx, fwd, sai.data(), sai.size(), reads)) != SGX_OK) return fail(ctx, "sgx_fm_attach_reads", rc);
  if ((rc = sgx_fm_attach_reads(ctx, rev, rsai.data(), rsai.size(), reads)) != SGX_OK) return fail(ctx, "sgx_fm_attach_reads", rc);
  uint32_t max_len = 0;
  for (size_t i = 0; i < rs.size(); ++i) max_len = std::max<uint32_t>(max_len, (uint32_t)(rs.off[i + 1] - rs.off[i]));
  // min-overlap = the longest read: no read has a proper overlap that long, only the full-length (containment) blocks remain
  sgx_hits hits;
  rc = sgx_overlap_batch(ctx, fwd, rev, reads, 0, rs.size(), max_len, /*irreducible=*/0, /*rc=*/1, &hits);
  if (rc != SGX_OK) return fail(ctx, "sgx_overlap_batch", rc);
  info("processed %zu sequences; device %.1f ms", rs.size(), hits.ms_total);
  const double t_gpu = now_s();

  // records are formatted per range of reads on all threads; both files keep the input order
  if (o.threads <= 0) o.threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const size_t n_parts = std::max<size_t>(1, std::min<size_t>((size_t)o.threads, (rs.size() + 65535) / 65536));
  std::vector<std::string> keep(n_parts), dups(n_parts);
  std::vector<uint64_t> part_dups(n_parts, 0);
  {
    std::vector<std::thread> pool;
    for (size_t part = 0; part < n_parts; ++part)
      pool.emplace_back([&, part]() {
        const size_t q0 = rs.size() * part / n_parts, q1 = rs.size() * (part + 1) / n_parts;
        keep[part].reserve((size_t)(rs.off[q1] - rs.off[q0]) + (q1 - q0) * 48);
        for (size_t q = q0; q < q1; ++q) {
          const std::string& qn = rs.names[q];
          uint64_t copies = 0;
          bool contained = hits.substring[q] != 0;
          for (uint64_t b = hits.block_off[q]; b < hits.block_off[q + 1]; ++b) {
            const sgx_block& blk = hits.blocks[b];
            if (blk.af != 0 && blk.af != (SGX_AF_TARGETREV | SGX_AF_QUERYCOMP)) continue;
            const std::vector<uint32_t>& sa = (blk.af & SGX_AF_TARGETREV) ? rsai : sai;
            copies += blk.capped[1] - blk.capped[0] + 1;
            for (uint64_t j = blk.capped[0]; j <= blk.capped[1] && !contained; ++j) {
              const std::string& tn = rs.names[sa[j]];
              if (tn != qn && !(qn < tn)) contained = true;  // an identical / reverse-complement read with a smaller name stays
            }
          }
          std::string* dst = contained ? &dups[part] : &keep[part];
          *dst += '>';
          *dst += qn;
          if (contained) {
            *dst += ",seqrank=";
            append_uint(dst, q);
            ++part_dups[part];
          }
          *dst += ' ';
          *dst += qn;
          *dst += " NumDuplicates=";
          append_uint(dst, copies);
          *dst += '\n';
          dst->append(rs.bases.data() + rs.off[q], rs.off[q + 1] - rs.off[q]);
          *dst += '\n';
        }
      });
    for (auto& th : pool) th.join();
  }
  uint64_t n_dups = 0;
  for (uint64_t d : part_dups) n_dups += d;
  for (int k = 0; k < 2; ++k) {
    const std::string path = output + (k == 0 ? ".rmdup.fa" : ".rmdup.dups.fa");
    const bool good = write_parts_parallel(path, k == 0 ? keep : dups);
    if (!good) {
      fprintf(stderr, "ERROR siga - Failed to create FASTA %s\n", path.c_str());
      return -1;
    }
  }
  info("%zu reads kept, %llu duplicates / substrings removed", rs.size() - (size_t)n_dups, (unsigned long long)n_dups);
  const double t1 = now_s();
  info("wall: %.3f sec (load %.3f, gpu %.3f, write %.3f)", t1 - t0, t_load - t0, t_gpu - t_load, t1 - t_gpu);
  finish_ok();  // hits, indices, reads, context and host buffers go with the process
}

int usage_correct() {
  printf("siga correct [OPTION] ... READSFILE\n"
         "Correct sequencing errors in all the reads in READSFILE\n"
         "\n"
         "      -h, --help                       display this help and exit\n"
         "\n"
         "      -p, --prefix=PREFIX              use PREFIX instead of prefix of READSFILE for the names of the index files\n"
         "      -o, --outfile=FILE               write the corrected reads to FILE (default READFILE.ec.fa)\n"
         "      -t, --threads=NUM                host threads for the output text (the correction runs on the GPU)\n"
         "      -a, --algorithm=STR              specify the correction algorithm to use. STR must be kmer (default: kmer)\n"
         "\n"
         "      -k, --kmer-size=N                the length of the kmer to user (default: 31)\n"
         "      -x, --kmer-threshold=N           attempt to correct kmers that are seen less than N times (default: 3)\n"
         "      -i, --kmer-rounds=N              perform up to N rounds of kmer correction (default: 10)\n"
         "      -O, --kmer-count-offset=N        when correcting a kmer, require the count of the new kmer is at least +N higher than the count of the old kmer. (default: 1)\n"
         "          --device=N                   CUDA device (default 0)\n"
         "          --gpus=N                     correct on N GPUs (devices D .. D+N-1): reads sharded by contiguous range\n"
         "          --gzip-level=N               zlib level of a .gz OUTFILE (default 1)\n"
         "\n");
  return 256;
}

// Text of the written reads of [first, first + count) in input order (PostCorrector, src/correct_processor.cpp:258-264;
// record format: operator<<, src/kseq.cpp:106-125), produced on `threads` threads; plain file or gzip members.
// valid[i - first] tells whether read i is written.
bool write_corrected(const std::string& path, bool gz, int threads, int gzip_level, bool gzip_fast, const ReadSet& rs,
                     const uint8_t* valid, size_t first, size_t count) {
  const bool fastq = !rs.quals.empty();
  const size_t chunk = 1 << 16, n_chunks = (count + chunk - 1) / chunk;
  auto produce = [&](size_t ci, std::string* text) {
    const size_t b = first + ci * chunk, e = std::min(first + count, b + chunk);
    text->reserve(text->size() + (size_t)(rs.off[e] - rs.off[b]) * (fastq ? 2 : 1) + (e - b) * 64);
    for (size_t i = b; i < e; ++i) {
      if (!valid[i - first]) continue;
      *text += fastq ? '@' : '>';
      *text += rs.names[i];
      if (!rs.comments[i].empty()) {
        *text += ' ';
        *text += rs.comments[i];
      }
      *text += '\n';
      text->append(rs.bases.data() + rs.off[i], rs.off[i + 1] - rs.off[i]);
      *text += '\n';
      if (fastq) {
        *text += "+\n";
        text->append(rs.quals.data() + rs.off[i], rs.off[i + 1] - rs.off[i]);
        *text += '\n';
      }
    }
  };
  if (threads < 1) threads = 1;
  if (gz) {
    GzChunkWriter w(path, (unsigned)threads, gzip_level);
    if (!w.ok()) return false;
    const bool wrote = w.write_chunks(std::max<size_t>(n_chunks, 1), [&](size_t ci, std::string* text) {
      if (ci < n_chunks) produce(ci, text);
    }, [&](size_t) { return gzip_fast ? GzChunkWriter::kLiteral : GzChunkWriter::kZlib; });
    return w.close() && wrote;
  }
  // plain file: one range of reads per thread, formatted and then written at its own offset
  const size_t n_parts = std::max<size_t>(1, std::min<size_t>((size_t)threads, n_chunks));
  std::vector<std::string> text(n_parts);
  {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n_parts; ++t)
      pool.emplace_back([&, t]() {
        const size_t c0 = n_chunks * t / n_parts, c1 = n_chunks * (t + 1) / n_parts;
        const size_t r0 = first + c0 * chunk, r1 = std::min(first + count, first + c1 * chunk);
        if (r1 > r0) text[t].reserve((size_t)(rs.off[r1] - rs.off[r0]) * (fastq ? 2 : 1) + (r1 - r0) * 64);
        for (size_t ci = c0; ci < c1; ++ci) produce(ci, &text[t]);
      });
    for (auto& th : pool) th.join();
  }
  return write_parts_parallel(path, text);
}

// `siga correct` (src/correct.cpp:22-59, CorrectProcessor::process src/correct_processor.cpp:269-316): loads PREFIX.bwt
// only, corrects every read with sgx_correct_batch and writes, in input order, the reads that end all-solid
// (PostCorrector, src/correct_processor.cpp:258-264) in the format of the input (operator<<, src/kseq.cpp:106-125).
// --gpus N: reads are independent units; rank r corrects the contiguous read range r of N against its own copy of the
// forward index (no collective) and writes the records of its range into a part file; the parent concatenates the
// parts in rank order, which is input order.
int cmd_correct(int argc, char** argv) {
  enum { OPT_DEVICE = 1000, OPT_GPUS, OPT_GZIP_LEVEL };
  static const option longopts[] = {{"log4cxx", required_argument, nullptr, 'c'},     {"ini", required_argument, nullptr, 's'},
                                    {"prefix", required_argument, nullptr, 'p'},      {"outfile", required_argument, nullptr, 'o'},
                                    {"threads", required_argument, nullptr, 't'},     {"algorithm", required_argument, nullptr, 'a'},
                                    {"kmer-size", required_argument, nullptr, 'k'},   {"kmer-threshold", required_argument, nullptr, 'x'},
                                    {"kmer-rounds", required_argument, nullptr, 'i'}, {"kmer-count-offset", required_argument, nullptr, 'O'},
                                    {"device", required_argument, nullptr, OPT_DEVICE}, {"gpus", required_argument, nullptr, OPT_GPUS},
                                    {"gzip-level", required_argument, nullptr, OPT_GZIP_LEVEL}, {"help", no_argument, nullptr, 'h'},
                                    {nullptr, 0, nullptr, 0}};
  CommonOpts o;
  std::string outfile, algorithm = "kmer";
  sgx_correct_params prm;
  prm.kmer_size = 31;
  prm.kmer_threshold = 3;
  prm.kmer_rounds = 10;
  prm.kmer_count_offset = 1;
  int c;
  while ((c = getopt_long(argc, argv, "c:s:p:o:t:a:k:x:i:O:h", longopts, nullptr)) != -1) {
    switch (c) {
      case 't': o.threads = atoi(optarg); break;
      case 'p': o.prefix = optarg; break;
      case 'o': outfile = optarg; break;
      case 'a': algorithm = optarg; break;
      case 'k': prm.kmer_size = (uint32_t)std::min<unsigned long long>(strtoull(optarg, nullptr, 10), 0xffffffffull); break;
      case 'x': prm.kmer_threshold = atoi(optarg); break;
      case 'i': prm.kmer_rounds = (uint32_t)std::min<unsigned long long>(strtoull(optarg, nullptr, 10), 0xffffffffull); break;
      case 'O': prm.kmer_count_offset = strtoull(optarg, nullptr, 10); break;
      case OPT_DEVICE: o.device = atoi(optarg); break;
      case OPT_GPUS: o.gpus = std::max(1, std::min(16, atoi(optarg))); break;
      case OPT_GZIP_LEVEL:
        o.gzip_level = std::max(1, std::min(9, atoi(optarg)));
        o.gzip_fast = false;
        break;
      case 'h': o.help = true; break;
      default: break;
    }
  }
  if (o.help || argc - optind != 1) return usage_correct();
  if (algorithm != "kmer") {
    // the reference's "overlap" corrector is an empty stub (src/correct_processor.cpp:234-243) and any other name fails
    fprintf(stderr, "ERROR siga - Failed to do error correction: algorithm '%s' is not available (kmer only)\n", algorithm.c_str());
    return -1;
  }
  const std::string input = argv[optind];
  const std::string st = stem(input);
  if (outfile.empty()) outfile = st + ".ec.fa";
  const std::string prefix = o.prefix.empty() ? st : o.prefix;
  info("input: %s", input.c_str());
  if (outfile.size() > 4 && outfile.compare(outfile.size() - 4, 4, ".bz2") == 0) {
    fprintf(stderr, "ERROR siga - Failed to create DNASeqWriter %s (bzip2 output is not supported by this build)\n", outfile.c_str());
    return -1;
  }
  const bool gz = outfile.size() > 3 && outfile.compare(outfile.size() - 3, 3, ".gz") == 0;
  const double t0 = now_s();
  if (o.threads <= 0) o.threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const int world = o.gpus;

  ReadSet rs;
  rs.keep_quals = true;
  std::string err;
  uint64_t ns = 0, nf = 0;
  std::vector<uint8_t> runs;
  sgx_ctx* ctx = nullptr;
  int ctx_rc = SGX_OK;
  bool ok_bwt = false, ok_reads = false;
  {
    std::vector<std::thread> pool;
    if (world == 1) pool.emplace_back([&]() { ctx_rc = sgx_init(o.device, &ctx); });  // no CUDA call before a fork
    pool.emplace_back([&]() { ok_bwt = read_bwt(prefix + ".bwt", &ns, &nf, &runs); });
    pool.emplace_back([&]() { ok_reads = read_sequences(input, &rs, &err); });
    for (auto& t : pool) t.join();
  }
  if (!ok_bwt) {
    fprintf(stderr, "ERROR siga - Failed to load FMIndex from %s\n", prefix.c_str());
    return -1;
  }
  if (!ok_reads) {
    fprintf(stderr, "ERROR siga - %s\n", err.c_str());
    return -1;
  }
  int my = 0;
  Gang gang;
  if (world > 1) {
    if ((my = gang.spawn(world)) < 0) return -1;
    if (my == 0)
      for (int r = 0; r < world; ++r) gang.add_temp(outfile + ".part" + std::to_string(r));
    ctx_rc = sgx_init(o.device + my, &ctx);
    o.threads = std::max(1, o.threads / world);
  }
  // a failing rank of a multi-GPU run leaves through _exit so that the parent sees a non-zero status
  auto bail = [&](const char* what, int rc) {
    fail(ctx_rc == SGX_OK ? ctx : nullptr, what, rc);
    if (my) _exit(1);
    return -1;
  };
  if (ctx_rc != SGX_OK) return bail("sgx_init", ctx_rc);
  const bool fastq = !rs.quals.empty();
  const double t_load = now_s();
  const uint64_t first = rs.size() * (uint64_t)my / world, count = rs.size() * (uint64_t)(my + 1) / world - first;
  sgx_corrected res;
  memset(&res, 0, sizeof(res));
  sgx_reads* reads = nullptr;
  sgx_fm* fm = nullptr;
  int rc = SGX_OK;
  if (count > 0) {
    // only this rank's range goes to the device (offsets need not start at zero)
    if ((rc = sgx_reads_pack(ctx, rs.bases.data(), rs.off.data() + first, count, &reads)) != SGX_OK) {
      if (rc == SGX_EINVAL) explain_bad_bases(rs, first, count);
      return bail("sgx_reads_pack", rc);
    }
    if ((rc = sgx_fm_create(ctx, runs.data(), runs.size(), nf, &fm)) != SGX_OK) return bail("sgx_fm_create", rc);
    rc = sgx_correct_batch(ctx, fm, reads, 0, count, fastq ? rs.quals.data() : nullptr, fastq ? rs.off.data() + first : nullptr, &prm, &res);
    if (rc != SGX_OK) return bail("sgx_correct_batch", rc);
    info("processed %llu sequences; device %.1f ms (kernel %.1f ms, %.0f rank lookups per read), %llu bases fixed in %llu reads",
         (unsigned long long)count, res.ms_total, res.ms_kernel, (double)res.lookups / (double)count, (unsigned long long)res.n_fixes,
         (unsigned long long)res.fixed_reads);
    for (uint64_t k = 0; k < res.n_fixes; ++k)
      rs.bases[rs.off[first + res.fixes[k].read] + (res.fixes[k].pos_code >> 2)] = "ACGT"[res.fixes[k].pos_code & 3];
  }
  const double t_gpu = now_s();

  bool good;
  if (world == 1) {
    good = write_corrected(outfile, gz, o.threads, o.gzip_level, o.gzip_fast, rs, res.valid, 0, rs.size());
  } else {
    good = write_corrected(outfile + ".part" + std::to_string(my), gz, o.threads, o.gzip_level, o.gzip_fast, rs, res.valid, first, count);
    if (my) {
      sgx_corrected_free(&res);
      // (the context goes with the process: see main)
      _exit(good ? 0 : 1);
    }
    if (!gang.wait_all()) good = false;
    if (good) {
      FILE* f = fopen(outfile.c_str(), "wb");
      good = f != nullptr;
      for (int r = 0; r < world && good; ++r) good = append_file(f, outfile + ".part" + std::to_string(r));
      if (f && fclose(f) != 0) good = false;
    }
    gang.remove_temps();
  }
  if (!good) {
    fprintf(stderr, "ERROR siga - Failed to create DNASeqWriter %s\n", outfile.c_str());
    return -1;
  }
  uint64_t written = 0;
  for (size_t i = 0; i < count; ++i) written += res.valid[i] ? 1 : 0;
  info("%llu of %llu reads%s written to %s", (unsigned long long)written, (unsigned long long)count,
       world > 1 ? " of rank 0's range" : "", outfile.c_str());
  const double t1 = now_s();
  info("wall: %.3f sec (load %.3f, gpu %.3f, write %.3f)", t1 - t0, t_load - t0, t_gpu - t_load, t1 - t_gpu);
  finish_ok();  // results, index, reads, context and host buffers go with the process
}

int usage_match() {
  printf("siga match [OPTION] ... READSFILE\n"
         "Match reads in READSFILE with ref\n"
         "\n"
         "      -h, --help                       display this help and exit\n"
         "\n"
         "      -p, --prefix=PREFIX              use PREFIX instead of prefix of READSFILE for the names of the index files\n"
         "      -l, --max-length=N               match only the first and the last N bases of longer reads\n"
         "          --no-opposite-strand         treat all reads as forward strand\n"
         "          --device=N                   CUDA device (default 0)\n"
         "\n");
  return 256;
}

// `siga match` (src/match.cpp:22-76): for every read of every READSFILE the number of occurrences of the read -- or of
// its first and last max-length bases -- in the forward index PREFIX.bwt, plus those of the reverse complement unless
// --no-opposite-strand; one "VT\t<0|1>\t<name>\t<seq>\t<count>" line per query on stdout, input order.  The counts are
// FMIndex::Interval::occurrences (src/fmindex.h:67-86) = one sgx_fm_count call per file.
int cmd_match(int argc, char** argv) {
  enum { OPT_NO_RC = 1000, OPT_DEVICE };
  static const option longopts[] = {{"log4cxx", required_argument, nullptr, 'c'},    {"ini", required_argument, nullptr, 's'},
                                    {"prefix", required_argument, nullptr, 'p'},     {"threads", required_argument, nullptr, 't'},
                                    {"max-length", required_argument, nullptr, 'l'}, {"no-opposite-strand", no_argument, nullptr, OPT_NO_RC},
                                    {"device", required_argument, nullptr, OPT_DEVICE}, {"help", no_argument, nullptr, 'h'},
                                    {nullptr, 0, nullptr, 0}};
  CommonOpts o;
  size_t max_len = (size_t)-1;
  bool rcomp = true;
  int c;
  while ((c = getopt_long(argc, argv, "c:s:p:t:l:h", longopts, nullptr)) != -1) {
    switch (c) {
      case 'p': o.prefix = optarg; break;
      case 'l': max_len = (size_t)strtoull(optarg, nullptr, 10); break;
      case OPT_NO_RC: rcomp = false; break;
      case OPT_DEVICE: o.device = atoi(optarg); break;
      case 'h': o.help = true; break;
      default: break;
    }
  }
  if (o.help || argc - optind < 1) return usage_match();
  const std::string prefix = o.prefix.empty() ? stem(argv[optind]) : o.prefix;
  uint64_t ns = 0, nf = 0;
  std::vector<uint8_t> runs;
  if (!read_bwt(prefix + ".bwt", &ns, &nf, &runs)) {
    fprintf(stderr, "ERROR siga - Failed to load FMIndex from %s\n", prefix.c_str());
    return -1;
  }
  sgx_ctx* ctx = nullptr;
  int rc = sgx_init(o.device, &ctx);
  if (rc != SGX_OK) return fail(nullptr, "sgx_init", rc);
  sgx_fm* fm = nullptr;
  if ((rc = sgx_fm_create(ctx, runs.data(), runs.size(), nf, &fm)) != SGX_OK) return fail(ctx, "sgx_fm_create", rc);
  int status = 0;
  for (int a = optind; a < argc && status == 0; ++a) {
    ReadSet rs;
    std::string err;
    if (!read_sequences(argv[a], &rs, &err)) {
      fprintf(stderr, "ERROR siga - %s\n", err.c_str());
      status = -1;
      break;
    }
    // queries of read i: the read (or its first max-length bases) and, for longer reads, its last max-length bases;
    // each followed by its reverse complement when both strands count
    std::string q;
    std::vector<uint64_t> qoff(1, 0);
    auto add = [&](const char* p, size_t len) {
      if (len) {  // an empty query occurs 0 times (Interval::get leaves the interval invalid)
        q.append(p, len);
        qoff.push_back(q.size());
      }
      if (len && rcomp) {
        for (size_t k = 0; k < len; ++k) {
          const char ch = p[len - 1 - k];
          q += ch == 'A' ? 'T' : (ch == 'C' ? 'G' : (ch == 'G' ? 'C' : (ch == 'T' ? 'A' : ch)));
        }
        qoff.push_back(q.size());
      }
    };
    for (size_t i = 0; i < rs.size(); ++i) {
      const char* seq = rs.bases.data() + rs.off[i];
      const size_t len = rs.off[i + 1] - rs.off[i];
      if (len > max_len) {
        add(seq, max_len);
        add(seq + len - max_len, max_len);
      } else {
        add(seq, len);
      }
    }
    // Queries with a byte other than A/C/G/T: the reference searches them with that byte mapped to rank 0, the
    // sentinel (DNAAlphabet::torank, src/alphabet.h:19-39; Interval::get, src/fmindex.h:67-78).  They are rare
    // (un-preprocessed input), so they are stepped here on the host through sgx_fm_occ with exactly that mapping and
    // only the clean queries go to the batched device search (sgx_reads_pack rejects other bytes).
    std::vector<uint64_t> counts(qoff.size() - 1);
    std::string qc;
    std::vector<uint64_t> qcoff(1, 0);
    std::vector<size_t> clean_ix;
    uint64_t pred[5] = {0, 0, 0, 0, 0};
    sgx_fm_pred(fm, pred);
    bool occ_fail = false;
    for (size_t k = 0; k + 1 < qoff.size() && !occ_fail; ++k) {
      const char* w = q.data() + qoff[k];
      const size_t len = qoff[k + 1] - qoff[k];
      bool clean = true;
      for (size_t j = 0; j < len && clean; ++j) clean = w[j] == 'A' || w[j] == 'C' || w[j] == 'G' || w[j] == 'T';
      if (clean) {
        qc.append(w, len);
        qcoff.push_back(qc.size());
        clean_ix.push_back(k);
        continue;
      }
      auto rank_of = [](char ch) { return ch == 'A' ? 1 : ch == 'C' ? 2 : ch == 'G' ? 3 : ch == 'T' ? 4 : 0; };
      uint64_t lower = 0, upper = (uint64_t)-1;
      for (size_t j = len; j-- > 0;) {
        const int r = rank_of(w[j]);
        uint64_t pos[2], occ[10];
        if (j + 1 == len) {  // Interval::init
          pos[0] = pos[1] = nf - 1;
          if (sgx_fm_occ(ctx, fm, pos, 1, occ) != SGX_OK) occ_fail = true;
          lower = pred[r];
          upper = lower + occ[r] - 1;
        } else {             // Interval::update
          pos[0] = lower - 1;
          pos[1] = upper;
          if (sgx_fm_occ(ctx, fm, pos, 2, occ) != SGX_OK) occ_fail = true;
          lower = pred[r] + occ[r];
          upper = pred[r] + occ[5 + r] - 1;
        }
        if (occ_fail || !(upper != (uint64_t)-1 && upper >= lower)) break;
      }
      counts[k] = (upper != (uint64_t)-1 && upper >= lower) ? upper - lower + 1 : 0;
    }
    std::vector<uint64_t> ccounts(clean_ix.size());
    if (occ_fail || (!ccounts.empty() && (rc = sgx_fm_count(ctx, fm, qc.data(), qcoff.data(), ccounts.size(), ccounts.data())) != SGX_OK)) {
      fail(ctx, occ_fail ? "sgx_fm_occ" : "sgx_fm_count", rc);
      status = -1;
      break;
    }
    for (size_t k = 0; k < clean_ix.size(); ++k) counts[clean_ix[k]] = ccounts[k];
    std::string text;
    size_t k = 0;
    const size_t per = rcomp ? 2 : 1;
    for (size_t i = 0; i < rs.size(); ++i) {
      const size_t len = rs.off[i + 1] - rs.off[i];
      const int parts = len > max_len ? 2 : 1;
      const size_t qlen = len > max_len ? max_len : len;
      for (int part = 0; part < parts; ++part) {
        uint64_t n = 0;
        if (qlen)
          for (size_t j = 0; j < per; ++j) n += counts[k++];
        text += part == 0 ? "VT\t0\t" : "VT\t1\t";
        text += rs.names[i];
        text += '\t';
        text.append(rs.bases.data() + rs.off[i], len);
        text += '\t';
        append_uint(&text, n);
        text += '\n';
      }
    }
    if (fwrite(text.data(), 1, text.size(), stdout) != text.size()) status = -1;
  }
  fflush(stdout);
  sgx_fm_free(fm);
  // (the context goes with the process: see main)
  return status;
}

}  // namespace

// -s / --ini FILE (src/main.cpp:63-77): `key=value` lines of an INI file are options, command-line options override
// them.  The file's entries become `--key=value` arguments placed before the command line's own, and every option loop
// here lets the last occurrence win.  A file that cannot be read is an error, as in the reference.
bool expand_ini(int argc, char** argv, std::vector<std::string>* store, std::vector<char*>* out) {
  std::string file;
  for (int a = 2; a < argc; ++a) {
    const std::string s = argv[a];
    if ((s == "-s" || s == "--ini") && a + 1 < argc) file = argv[a + 1];
    else if (s.compare(0, 6, "--ini=") == 0) file = s.substr(6);
    else if (s.size() > 2 && s[0] == '-' && s[1] == 's') file = s.substr(2);
  }
  out->assign(argv, argv + argc);
  if (file.empty()) return true;
  FILE* f = fopen(file.c_str(), "r");
  if (!f) {
    fprintf(stderr, "ERROR siga - load %s failed(cannot open file).\n", file.c_str());
    return false;
  }
  char line[4096];
  store->reserve(256);  // the pointers handed out below must stay valid
  while (fgets(line, sizeof(line), f) && store->size() < 256) {
    std::string s = line;
    const size_t cpos = s.find_first_of(";#");
    if (cpos == 0) continue;
    while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
    size_t b = 0;
    while (b < s.size() && isspace((unsigned char)s[b])) ++b;
    s = s.substr(b);
    if (s.empty() || s[0] == '[') continue;  // sections only group keys; the option names are flat
    const size_t eq = s.find('=');
    std::string key = eq == std::string::npos ? s : s.substr(0, eq), val = eq == std::string::npos ? "" : s.substr(eq + 1);
    while (!key.empty() && isspace((unsigned char)key.back())) key.pop_back();
    size_t vb = 0;
    while (vb < val.size() && isspace((unsigned char)val[vb])) ++vb;
    val = val.substr(vb);
    if (key.empty() || key == "ini") continue;
    store->push_back(val.empty() ? "--" + key : "--" + key + "=" + val);
  }
  fclose(f);
  out->clear();
  out->push_back(argv[0]);
  out->push_back(argv[1]);
  for (auto& s : *store) out->push_back(&s[0]);
  for (int a = 2; a < argc; ++a) out->push_back(argv[a]);
  return true;
}

int main(int argc0, char** argv0) {
  std::vector<std::string> ini_store;
  std::vector<char*> args;
  if (argc0 >= 2 && !expand_ini(argc0, argv0, &ini_store, &args)) return 1;
  if (argc0 < 2) args.assign(argv0, argv0 + argc0);
  args.push_back(nullptr);
  const int argc = (int)args.size() - 1;
  char** argv = args.data();
  if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) {
    printf("siga (B200 build) -- Usage: siga <command> [<args>]\n\n"
           "   index       build the BWT and FM-index for a set of reads\n"
           "   overlap     compute overlaps between reads\n"
           "   rmdup       remove duplicated reads from the data set\n"
           "   correct     correct sequencing errors in a set of reads\n"
           "   match       count the occurrences of reads in an indexed read set\n\n"
           "See 'siga <command> -h' to read about a specific subcommand.\n");
    return 256;
  }
  const std::string cmd = argv[1];
  // Every output file is closed (and its close checked) inside the command.  Leaving through _exit skips the tear-down
  // of the CUDA context and of gigabytes of cached device work areas by the runtime's exit handlers, which costs up to
  // a second and a half and frees nothing the process exit does not free anyway.
  auto leave = [](int rc) {
    fflush(nullptr);
    _exit(rc & 0xff);
    return rc;
  };
  if (cmd == "index") return leave(cmd_index(argc - 1, argv + 1));
  if (cmd == "overlap") return leave(cmd_overlap(argc - 1, argv + 1));
  if (cmd == "rmdup") return leave(cmd_rmdup(argc - 1, argv + 1));
  if (cmd == "correct") return leave(cmd_correct(argc - 1, argv + 1));
  if (cmd == "match") return leave(cmd_match(argc - 1, argv + 1));
  fprintf(stderr, "siga: '%s' is not provided by this build (only index / overlap / rmdup / correct / match are)\n", cmd.c_str());
  return 256;
}
