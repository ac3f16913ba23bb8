// Oracle build shim (test infrastructure only).
// Implements the interface declared by the reference's src/utils.h (Utils::{srand,rand,cputime,maxrss,ifstream,
// ofstream,stem}) on top of zlib, because the reference's own src/utils.cpp needs boost::iostreams which is not
// installed in this image.  Behaviour mirrored: ".gz" names are (de)compressed, an input stream can seek back to
// offset 0 (the reference re-reads its input, src/overlap_builder.cpp:334), stem() strips the directory and one
// extension after first stripping .gz/.bz2 (src/utils.cpp:128-135).
#include "utils.h"

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <streambuf>
#include <string>
#include <vector>

#include <sys/resource.h>
#include <zlib.h>

namespace {

bool has_suffix(const std::string& s, const std::string& p) {
  return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0;
}

class GzInBuf : public std::streambuf {
 public:
  explicit GzInBuf(const std::string& fn) : buf_(1 << 16) { f_ = gzopen(fn.c_str(), "rb"); if (f_) gzbuffer(f_, 1 << 18); }
  ~GzInBuf() { if (f_) gzclose(f_); }
  bool ok() const { return f_ != NULL; }
 protected:
  int_type underflow() {
    if (gptr() < egptr()) return traits_type::to_int_type(*gptr());
    int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
    if (n <= 0) return traits_type::eof();
    setg(buf_.data(), buf_.data(), buf_.data() + n);
    return traits_type::to_int_type(*gptr());
  }
  pos_type seekoff(off_type off, std::ios_base::seekdir way, std::ios_base::openmode) {
    if (off == 0 && way == std::ios_base::beg) { gzrewind(f_); setg(buf_.data(), buf_.data(), buf_.data()); return pos_type(0); }
    return pos_type(off_type(-1));
  }
  pos_type seekpos(pos_type p, std::ios_base::openmode m) { return seekoff(off_type(p), std::ios_base::beg, m); }
 private:
  gzFile f_;
  std::vector<char> buf_;
};

class GzOutBuf : public std::streambuf {
 public:
  explicit GzOutBuf(const std::string& fn) : buf_(1 << 16) {
    f_ = gzopen(fn.c_str(), "wb");
    setp(buf_.data(), buf_.data() + buf_.size());
  }
  ~GzOutBuf() { sync(); if (f_) gzclose(f_); }
  bool ok() const { return f_ != NULL; }
 protected:
  int_type overflow(int_type c) {
    if (flush_buf() < 0) return traits_type::eof();
    if (c != traits_type::eof()) { *pptr() = traits_type::to_char_type(c); pbump(1); }
    return traits_type::not_eof(c);
  }
  int sync() { return flush_buf(); }
 private:
  int flush_buf() {
    std::ptrdiff_t n = pptr() - pbase();
    if (n > 0 && f_) { if (gzwrite(f_, pbase(), (unsigned)n) != n) return -1; }
    setp(buf_.data(), buf_.data() + buf_.size());
    return 0;
  }
  gzFile f_;
  std::vector<char> buf_;
};

template <class Buf, class Base>
class OwningStream : public Base {
 public:
  explicit OwningStream(Buf* b) : Base(b), b_(b) {}
  ~OwningStream() { this->rdbuf(NULL); delete b_; }
 private:
  Buf* b_;
};

}  // namespace

namespace Utils {

void srand() { ::srand(time(NULL)); }
int rand() { return ::rand(); }

double cputime() {
  struct rusage r;
  getrusage(RUSAGE_SELF, &r);
  return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}

double maxrss() {
  struct rusage r;
  getrusage(RUSAGE_SELF, &r);
  return (r.ru_maxrss * 1024.0) / 1073741824.0;
}

std::istream* ifstream(const std::string& filename) {
  // gzopen/gzread transparently pass plain files through, so one path serves both.
  GzInBuf* b = new GzInBuf(filename);
  if (!b->ok()) { delete b; return NULL; }
  return new OwningStream<GzInBuf, std::istream>(b);
}

std::ostream* ofstream(const std::string& filename) {
  if (has_suffix(filename, ".gz")) {
    GzOutBuf* b = new GzOutBuf(filename);
    if (!b->ok()) { delete b; return NULL; }
    return new OwningStream<GzOutBuf, std::ostream>(b);
  }
  std::ofstream* f = new std::ofstream(filename.c_str(), std::ios_base::out | std::ios_base::binary);
  if (!*f) { delete f; return NULL; }
  return f;
}

std::string stem(const std::string& filename) {
  if (has_suffix(filename, ".gz")) return stem(filename.substr(0, filename.size() - 3));
  if (has_suffix(filename, ".bz2")) return stem(filename.substr(0, filename.size() - 4));
  std::string f = filename;
  size_t k = f.find_last_of('/');
  if (k != std::string::npos) f = f.substr(k + 1);
  k = f.find_last_of('.');
  if (k != std::string::npos && k != 0) f = f.substr(0, k);
  return f;
}

}  // namespace Utils
