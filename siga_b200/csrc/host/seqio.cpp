#include "seqio.h"

#include <thread>

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace sgxhost {

namespace {

bool ends_with(const std::string& s, const std::string& p) {
  return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0;
}

void trim(std::string& s) {
  size_t b = 0, e = s.size();
  while (b < e && isspace((unsigned char)s[b])) ++b;
  while (e > b && isspace((unsigned char)s[e - 1])) --e;
  if (b > 0 || e < s.size()) s = s.substr(b, e - b);
}

// Line reader over gzread (plain files pass through zlib unchanged).
class LineReader {
 public:
  explicit LineReader(const std::string& path) : buf_(1 << 20), pos_(0), len_(0) {
    f_ = gzopen(path.c_str(), "rb");
    if (f_) gzbuffer(f_, 1 << 20);
  }
  ~LineReader() {
    if (f_) gzclose(f_);
  }
  bool ok() const { return f_ != NULL; }
  int peek() {
    if (pos_ >= len_ && !fill()) return -1;
    return (unsigned char)buf_[pos_];
  }
  bool getline(std::string& line) {
    line.clear();
    bool any = false;
    for (;;) {
      if (pos_ >= len_ && !fill()) return any;
      any = true;
      const char* p = &buf_[pos_];
      const char* nl = (const char*)memchr(p, '\n', len_ - pos_);
      if (nl) {
        line.append(p, nl - p);
        pos_ += (nl - p) + 1;
        return true;
      }
      line.append(p, len_ - pos_);
      pos_ = len_;
    }
  }

 private:
  bool fill() {
    int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
    if (n <= 0) return false;
    pos_ = 0;
    len_ = (size_t)n;
    return true;
  }
  gzFile f_;
  std::vector<char> buf_;
  size_t pos_, len_;
};

void split_name(const std::string& header, std::string* name, std::string* comment) {
  size_t i = header.find_first_of(" \t");
  if (i != std::string::npos) {
    *name = header.substr(0, i);
    *comment = header.substr(i + 1);
  } else {
    *name = header;
    comment->clear();
  }
}

void push(ReadSet* rs, const std::string& header, const std::string& seq, const std::string* qual = nullptr) {
  std::string name, comment;
  split_name(header, &name, &comment);
  rs->names.push_back(name);
  rs->comments.push_back(comment);
  rs->bases.append(seq.data(), seq.size());
  if (qual && rs->keep_quals) rs->quals.append(qual->data(), qual->size());
  rs->off.push_back(rs->bases.size());
}

}  // namespace

bool read_sequences_serial(const std::string& path, ReadSet* out, std::string* err) {
  LineReader in(path);
  if (!in.ok()) {
    *err = "Failed to open input file " + path;
    return false;
  }
  out->names.clear();
  out->comments.clear();
  out->bases.clear();
  out->quals.release();
  out->off.assign(1, 0);
  const int first = in.peek();
  std::string line;
  if (first == '>') {
    // FASTAReader::read (src/kseq.cpp:191-228) keys everything on the text behind '>' being non-empty: a header that
    // follows a named record without sequence ends the input (207-210); a bare ">" is no name at all -- the sequence lines
    // behind it stay in the buffer and join the NEXT named record, and are dropped at the end of the file.
    std::string name, seq;
    while (in.getline(line)) {
      trim(line);
      if (line.empty()) continue;
      if (line[0] == '>') {
        if (!seq.empty() && !name.empty()) {
          push(out, name, seq);
          seq.clear();
        } else if (!name.empty()) {
          break;
        }
        name = line.substr(1);
      } else {
        seq += line;
      }
    }
    if (!seq.empty() && !name.empty()) push(out, name, seq);
    return true;
  }
  if (first == '@') {
    int state = 0;
    std::string name, seq;
    while (in.getline(line)) {
      trim(line);
      if (line.empty()) continue;
      if (state == 0) {
        if (line[0] != '@') break;
        name = line.substr(1);
        state = 1;
      } else if (state == 1) {
        seq = line;
        state = 2;
      } else if (state == 2) {
        if (!(line[0] == '+' && (line.size() == 1 || ends_with(line, name)))) break;
        state = 3;
      } else {
        if (line.size() != seq.size()) break;
        push(out, name, seq, &line);
        state = 0;
      }
    }
    return true;
  }
  *err = "failed to create dna reader (the file starts with neither '>' nor '@'): " + path;
  return false;
}

// ---- parallel parse of plain (uncompressed) files ------------------------------------------------------------------------
// The file is mapped, cut into one range per thread at record starts, every range is parsed with the same line rules
// as the serial reader (lines trimmed, empty lines skipped, multi-line FASTA joined) and the parts are concatenated.
// Anything the serial reader would stop at (a header after an empty FASTA record, a malformed FASTQ record) makes the
// whole call fall back to the serial reader, so the result is always the serial one.
namespace {

struct Line {
  const char* b;
  const char* e;
};
// next trimmed, non-empty line of [p, end); false at the end
inline bool next_line(const char*& p, const char* end, Line* ln) {
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    const char* b = p;
    const char* e = le;
    p = nl ? nl + 1 : end;
    while (b < e && isspace((unsigned char)*b)) ++b;
    while (e > b && isspace((unsigned char)e[-1])) --e;
    if (b < e) {
      ln->b = b;
      ln->e = e;
      return true;
    }
  }
  return false;
}

// One pass over the records of [p, end) with the line rules of the serial reader; the sink sees every record as
// header(hb, he), then seq(b, e) per sequence line (and qual(b, e) for FASTQ), then done().  false = something the
// serial reader treats specially (a header after an empty FASTA record, a malformed FASTQ record): redo serially.
template <class Sink>
bool scan_fasta_range(const char* p, const char* end, Sink& sink) {
  Line ln;
  bool have = false, any_seq = false;
  while (next_line(p, end, &ln)) {
    if (*ln.b == '>') {
      if (ln.e - ln.b == 1) return false;  // a bare ">": the serial reader has the reference's rule for it
      if (have && any_seq) sink.done();
      else if (have) return false;
      sink.header(ln.b + 1, ln.e);
      have = true;
      any_seq = false;
    } else {
      if (!have) return false;
      sink.seq(ln.b, ln.e);
      any_seq = true;
    }
  }
  if (have && any_seq) sink.done();
  else if (have) return false;
  return true;
}

template <class Sink>
bool scan_fastq_range(const char* p, const char* end, Sink& sink) {
  Line h, sq, pl, ql;
  while (next_line(p, end, &h)) {
    if (*h.b != '@') return false;
    if (!next_line(p, end, &sq) || !next_line(p, end, &pl) || !next_line(p, end, &ql)) return false;
    const size_t nlen = (size_t)(h.e - h.b - 1);
    const bool plus_ok = *pl.b == '+' && (pl.e - pl.b == 1 || ((size_t)(pl.e - pl.b) >= nlen && memcmp(pl.e - nlen, h.b + 1, nlen) == 0));
    if (!plus_ok || (ql.e - ql.b) != (sq.e - sq.b)) return false;
    sink.header(h.b + 1, h.e);
    sink.seq(sq.b, sq.e);
    sink.qual(ql.b, ql.e);
    sink.done();
  }
  return true;
}

// first pass: how many records and bases a range holds
struct CountSink {
  size_t records = 0, bases = 0;
  void header(const char*, const char*) {}
  void seq(const char* b, const char* e) { bases += (size_t)(e - b); }
  void qual(const char*, const char*) {}
  void done() { ++records; }
};

// second pass: every range writes its records straight into their final places
struct FillSink {
  ReadSet* out;
  size_t rec, pos;  // next record index, next byte of bases / quals
  bool quals;
  void header(const char* hb, const char* he) {
    const char* sp = hb;
    while (sp < he && *sp != ' ' && *sp != '\t') ++sp;
    out->names[rec].assign(hb, sp);
    out->comments[rec].assign(sp < he ? sp + 1 : he, he);
  }
  void seq(const char* b, const char* e) {
    memcpy(out->bases.data() + pos, b, (size_t)(e - b));
    pos += (size_t)(e - b);
  }
  void qual(const char* b, const char* e) {
    if (quals) memcpy(out->quals.data() + pos - (size_t)(e - b), b, (size_t)(e - b));  // same offsets as the bases
  }
  void done() { out->off[++rec] = pos; }
};

}  // namespace

namespace {
// bzip2 input (the reference reads .bz2 through boost::iostreams, src/utils.cpp:92-95).  There are no bzlib headers in
// the build image, only the runtime library, so libbz2.so.1 is resolved with dlopen and the stream interface declared
// here (stable since bzip2 1.0).  The file is decompressed into a temporary plain file, which the normal readers parse.
struct BzStream {
  char* next_in;
  unsigned int avail_in, total_in_lo32, total_in_hi32;
  char* next_out;
  unsigned int avail_out, total_out_lo32, total_out_hi32;
  void* state;
  void* (*bzalloc)(void*, int, int);
  void (*bzfree)(void*, void*);
  void* opaque;
};

bool bunzip2_to_temp(const std::string& path, std::string* tmp_path, std::string* err) {
  void* h = dlopen("libbz2.so.1.0", RTLD_NOW);
  if (!h) h = dlopen("libbz2.so.1", RTLD_NOW);
  if (!h) {
    *err = "cannot read " + path + ": bzip2 input needs libbz2.so.1, which is not installed (decompress the file first)";
    return false;
  }
  typedef int (*InitF)(BzStream*, int, int);
  typedef int (*StepF)(BzStream*);
  InitF init = (InitF)dlsym(h, "BZ2_bzDecompressInit");
  StepF step = (StepF)dlsym(h, "BZ2_bzDecompress"), end = (StepF)dlsym(h, "BZ2_bzDecompressEnd");
  if (!init || !step || !end) {
    *err = "cannot read " + path + ": libbz2 lacks the decompression interface";
    return false;
  }
  FILE* in = fopen(path.c_str(), "rb");
  if (!in) {
    *err = "cannot open " + path;
    return false;
  }
  const char* tdir = getenv("TMPDIR");
  std::string tmpl = std::string(tdir && *tdir ? tdir : "/tmp") + "/siga_bz2_XXXXXX";
  std::vector<char> name(tmpl.begin(), tmpl.end());
  name.push_back('\0');
  const int fd = mkstemp(name.data());
  FILE* outf = fd >= 0 ? fdopen(fd, "wb") : nullptr;
  if (!outf) {
    fclose(in);
    *err = "cannot create a temporary file for " + path;
    return false;
  }
  *tmp_path = name.data();
  std::vector<char> ibuf(1 << 20), obuf(1 << 22);
  BzStream bz;
  memset(&bz, 0, sizeof(bz));
  bool ok = init(&bz, 0, 0) == 0, open_stream = ok;
  while (ok) {
    if (bz.avail_in == 0) {
      bz.avail_in = (unsigned int)fread(ibuf.data(), 1, ibuf.size(), in);
      bz.next_in = ibuf.data();
      if (bz.avail_in == 0) break;  // end of file (between streams when open_stream is false)
    }
    if (!open_stream) {  // concatenated streams
      ok = init(&bz, 0, 0) == 0;
      open_stream = ok;
      if (!ok) break;
    }
    bz.next_out = obuf.data();
    bz.avail_out = (unsigned int)obuf.size();
    const int rc = step(&bz);
    const size_t got = obuf.size() - bz.avail_out;
    if (got && fwrite(obuf.data(), 1, got, outf) != got) ok = false;
    if (rc == 4) {  // BZ_STREAM_END
      char* ni = bz.next_in;
      const unsigned int ai = bz.avail_in;
      end(&bz);
      memset(&bz, 0, sizeof(bz));
      bz.next_in = ni;
      bz.avail_in = ai;
      open_stream = false;
    } else if (rc != 0) {
      ok = false;
    }
  }
  if (open_stream) {
    end(&bz);
    ok = false;  // the file ended inside a stream
  }
  fclose(in);
  if (fclose(outf) != 0) ok = false;
  if (!ok) {
    remove(tmp_path->c_str());
    *err = "cannot read " + path + ": not a valid bzip2 file";
  }
  return ok;
}
}  // namespace

bool read_sequences(const std::string& path, ReadSet* out, std::string* err) {
  if (ends_with(path, ".bz2")) {
    std::string tmp;
    if (!bunzip2_to_temp(path, &tmp, err)) return false;
    const bool ok = read_sequences(tmp, out, err);
    remove(tmp.c_str());
    return ok;
  }
  size_t kMinParallel = (size_t)16 << 20;
  if (const char* mp = getenv("SIGA_PARSE_MIN_PARALLEL")) kMinParallel = (size_t)strtoull(mp, nullptr, 10);  // tests
  unsigned T = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (ends_with(path, ".gz") || T < 2) return read_sequences_serial(path, out, err);
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return read_sequences_serial(path, out, err);  // the serial reader produces the error text
  struct stat st;
  if (fstat(fd, &st) != 0 || (size_t)st.st_size < kMinParallel) {
    close(fd);
    return read_sequences_serial(path, out, err);
  }
  const size_t size = (size_t)st.st_size;
  const char* base = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (base == MAP_FAILED) return read_sequences_serial(path, out, err);
  const char* end = base + size;
  const bool fasta = base[0] == '>', fastq = base[0] == '@';
  bool ok = fasta || fastq;
  std::vector<const char*> cut;
  if (ok) {
    cut.push_back(base);
    for (unsigned t = 1; t < T; ++t) {
      const char* p = base + size / T * t;
      if (p <= cut.back()) continue;
      const char* found = nullptr;
      while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!nl || nl + 1 >= end) break;
        p = nl + 1;
        if (fasta) {
          if (*p == '>') {
            found = p;
            break;
          }
        } else if (*p == '@') {
          // a header, not a quality line: the line after next starts with '+'
          const char* l2 = (const char*)memchr(p, '\n', (size_t)(end - p));
          const char* l3 = l2 ? (const char*)memchr(l2 + 1, '\n', (size_t)(end - l2 - 1)) : nullptr;
          if (l3 && l3 + 1 < end && l3[1] == '+') {
            found = p;
            break;
          }
        }
      }
      if (found && found > cut.back()) cut.push_back(found);
    }
    cut.push_back(end);
    const size_t parts = cut.size() - 1;
    // pass 1: sizes of every range
    std::vector<CountSink> cnt(parts);
    std::vector<char> good(parts, 0);
    {
      std::vector<std::thread> pool;
      for (size_t i = 0; i < parts; ++i)
        pool.emplace_back([&, i]() {
          good[i] = (fasta ? scan_fasta_range(cut[i], cut[i + 1], cnt[i]) : scan_fastq_range(cut[i], cut[i + 1], cnt[i])) ? 1 : 0;
        });
      for (auto& th : pool) th.join();
    }
    for (size_t i = 0; i < parts; ++i) ok = ok && good[i];
    if (ok) {
      std::vector<size_t> n0(parts + 1, 0), b0(parts + 1, 0);
      for (size_t i = 0; i < parts; ++i) {
        n0[i + 1] = n0[i] + cnt[i].records;
        b0[i + 1] = b0[i] + cnt[i].bases;
      }
      const bool quals = fastq && out->keep_quals;
      out->names.assign(n0[parts], std::string());
      out->comments.assign(n0[parts], std::string());
      out->off.assign(n0[parts] + 1, 0);
      out->quals.release();
      ok = out->bases.resize_uninit(b0[parts]) && (!quals || out->quals.resize_uninit(b0[parts]));
      if (ok) {
        // pass 2: every range fills its own slice (and faults in its own pages)
        std::vector<std::thread> pool;
        for (size_t i = 0; i < parts; ++i)
          pool.emplace_back([&, i]() {
            FillSink fs{out, n0[i], b0[i], quals};
            out->off[n0[i]] = b0[i];
            if (fasta) scan_fasta_range(cut[i], cut[i + 1], fs);
            else scan_fastq_range(cut[i], cut[i + 1], fs);
          });
        for (auto& th : pool) th.join();
      }
    }
  }
  munmap((void*)base, size);
  if (!ok) return read_sequences_serial(path, out, err);
  return true;
}

std::string stem(const std::string& path) {
  if (ends_with(path, ".gz")) return stem(path.substr(0, path.size() - 3));
  if (ends_with(path, ".bz2")) return stem(path.substr(0, path.size() - 4));
  std::string f = path;
  size_t k = f.find_last_of('/');
  if (k != std::string::npos) f = f.substr(k + 1);
  k = f.find_last_of('.');
  if (k != std::string::npos && k != 0) f = f.substr(0, k);
  return f;
}

bool write_bwt(const std::string& path, uint64_t n_strings, uint64_t n_suffixes, const uint8_t* runs, uint64_t n_runs) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const uint16_t magic = 0xCACA;
  const int32_t flag = 0;  // BWF_NOFMI
  bool ok = fwrite(&magic, 2, 1, f) == 1 && fwrite(&n_strings, 8, 1, f) == 1 && fwrite(&n_suffixes, 8, 1, f) == 1 &&
            fwrite(&n_runs, 8, 1, f) == 1 && fwrite(&flag, 4, 1, f) == 1;
  if (ok && n_runs) ok = fwrite(runs, 1, n_runs, f) == n_runs;
  return fclose(f) == 0 && ok;
}

bool write_bwt_segment(const std::string& path, uint64_t n_strings, uint64_t n_suffixes, uint64_t n_runs, const uint8_t* seg,
                       uint64_t seg_offset, uint64_t seg_bytes, bool header) {
  const int fd = open(path.c_str(), O_WRONLY | O_CREAT, 0644);
  if (fd < 0) return false;
  bool ok = true;
  auto put = [&](const void* p, uint64_t n, uint64_t at) {
    const uint8_t* b = (const uint8_t*)p;
    while (n && ok) {
      const ssize_t w = pwrite(fd, b, (size_t)std::min<uint64_t>(n, (uint64_t)1 << 30), (off_t)at);
      if (w <= 0) ok = false;
      else b += w, n -= (uint64_t)w, at += (uint64_t)w;
    }
  };
  if (header) {
    uint8_t h[30];
    const uint16_t magic = 0xCACA;
    const int32_t flag = 0;  // BWF_NOFMI
    memcpy(h, &magic, 2);
    memcpy(h + 2, &n_strings, 8);
    memcpy(h + 10, &n_suffixes, 8);
    memcpy(h + 18, &n_runs, 8);
    memcpy(h + 26, &flag, 4);
    ok = ftruncate(fd, (off_t)(30 + n_runs)) == 0;
    put(h, 30, 0);
  }
  put(seg, seg_bytes, 30 + seg_offset);
  return close(fd) == 0 && ok;
}

bool write_sai(const std::string& path, const uint32_t* ids, uint64_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  std::string buf;
  buf.reserve(1 << 22);
  char tmp[64];
  int k = snprintf(tmp, sizeof(tmp), "%u\n%llu\n%llu\n", 0xCACAu, (unsigned long long)n, (unsigned long long)n);
  buf.append(tmp, k);
  bool ok = true;
  for (uint64_t i = 0; i < n && ok; ++i) {
    // "<read> 0\n" : only full-length suffixes are listed (src/suffix_array.cpp:36-43)
    char* p = tmp + sizeof(tmp);
    *--p = '\n';
    *--p = '0';
    *--p = ' ';
    uint32_t v = ids[i];
    do {
      *--p = (char)('0' + v % 10);
      v /= 10;
    } while (v);
    buf.append(p, tmp + sizeof(tmp) - p);
    if (buf.size() > (1 << 22) - 64) {
      ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
      buf.clear();
    }
  }
  if (ok && !buf.empty()) ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
  return fclose(f) == 0 && ok;
}

bool read_bwt(const std::string& path, uint64_t* n_strings, uint64_t* n_suffixes, std::vector<uint8_t>* runs) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  uint16_t magic = 0;
  uint64_t n_runs = 0;
  int32_t flag = 0;
  bool ok = fread(&magic, 2, 1, f) == 1 && magic == 0xCACA && fread(n_strings, 8, 1, f) == 1 &&
            fread(n_suffixes, 8, 1, f) == 1 && fread(&n_runs, 8, 1, f) == 1 && fread(&flag, 4, 1, f) == 1;
  if (ok) {
    runs->resize(n_runs);
    if (n_runs) ok = fread(runs->data(), 1, n_runs, f) == n_runs;
  }
  fclose(f);
  return ok;
}

bool read_sai(const std::string& path, std::vector<uint32_t>* ids) {
  // "<magic>\n<n>\n<n>\n" then n lines "<read> <pos>\n" (src/suffix_array.cpp:56-95); parsed from one buffer
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::string buf;
  if (fseek(f, 0, SEEK_END) == 0) {
    const long sz = ftell(f);
    if (sz > 0) buf.resize((size_t)sz);
    fseek(f, 0, SEEK_SET);
  }
  const bool got = buf.empty() || fread(&buf[0], 1, buf.size(), f) == buf.size();
  fclose(f);
  if (!got) return false;
  const char* p = buf.data();
  const char* e = p + buf.size();
  auto next_uint = [&](unsigned long long* v) -> bool {
    while (p < e && (*p < '0' || *p > '9')) ++p;
    if (p >= e) return false;
    unsigned long long x = 0;
    while (p < e && *p >= '0' && *p <= '9') x = x * 10 + (unsigned long long)(*p++ - '0');
    *v = x;
    return true;
  };
  unsigned long long magic = 0, n = 0, m = 0;
  if (!next_uint(&magic) || !next_uint(&n) || !next_uint(&m) || magic != 0xCACA) return false;
  ids->resize(n);
  for (unsigned long long i = 0; i < n; ++i) {
    unsigned long long a = 0, b2 = 0;
    if (!next_uint(&a) || !next_uint(&b2)) return false;
    (*ids)[i] = (uint32_t)a;
  }
  return true;
}

std::vector<uint32_t> name_ranks(const std::vector<std::string>& names) {
  // index sort by name: chunks sorted on worker threads, then merged pairwise (3 M names: ~0.2 s instead of ~1.5 s)
  const size_t n = names.size();
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  auto less = [&](uint32_t a, uint32_t b) { return names[a] < names[b]; };
  unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < 100000) T = 1;
  unsigned P = 1;
  while (P * 2 <= T) P *= 2;  // power of two parts
  std::vector<size_t> cut(P + 1);
  for (unsigned i = 0; i <= P; ++i) cut[i] = n * i / P;
  {
    std::vector<std::thread> pool;
    for (unsigned i = 0; i < P; ++i)
      pool.emplace_back([&, i]() { std::sort(order.begin() + cut[i], order.begin() + cut[i + 1], less); });
    for (auto& t : pool) t.join();
  }
  for (unsigned w = 1; w < P; w *= 2) {
    std::vector<std::thread> pool;
    for (unsigned i = 0; i + w < P + 0 && i + w <= P; i += 2 * w)
      pool.emplace_back([&, i, w]() {
        std::inplace_merge(order.begin() + cut[i], order.begin() + cut[i + w], order.begin() + cut[std::min(i + 2 * w, P)], less);
      });
    for (auto& t : pool) t.join();
  }
  std::vector<uint32_t> rank(n);
  uint32_t r = 0;
  for (size_t i = 0; i < n; ++i) {
    if (i > 0 && names[order[i]] != names[order[i - 1]]) ++r;
    rank[order[i]] = r;
  }
  return rank;
}

}  // namespace sgxhost
