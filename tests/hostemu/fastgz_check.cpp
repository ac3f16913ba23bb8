// TEST INFRASTRUCTURE ONLY: writes gzip_member_huffman(file) (or, after a first argument `lz`, gzip_member_lz(file)) for every file given on the command line to <file>.gz and
// prints "<file> <input bytes> <output bytes> <MB/s>"; the Python test decompresses the results with zlib.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>

#include "../../siga_b200/csrc/host/fastgz.h"

int main(int argc, char** argv) {
  // first argument "lz": the LZ77 variant (gzip_member_lz) instead of the literal-only one
  const bool lz = argc > 1 && std::string(argv[1]) == "lz";
  for (int i = lz ? 2 : 1; i < argc; ++i) {
    std::ifstream f(argv[i], std::ios::binary);
    const std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const auto t0 = std::chrono::steady_clock::now();
    const std::string gz = lz ? sgxhost::gzip_member_lz(data.data(), data.size()) : sgxhost::gzip_member_huffman(data.data(), data.size());
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::ofstream o(std::string(argv[i]) + ".gz", std::ios::binary);
    o.write(gz.data(), (std::streamsize)gz.size());
    printf("%s %zu %zu %.0f\n", argv[i], data.size(), gz.size(), data.size() / 1e6 / (dt > 0 ? dt : 1e-9));
  }
  return 0;
}
