// Literal-only gzip member writer for the text outputs of the `siga` front-end (the VT records of .asqg.gz, which
// are ~60 % of the file and consist of read sequences: zlib's level-1 matcher finds next to nothing in them and
// costs ~20 ns per byte).  The output is an ordinary gzip member; members are concatenated by GzChunkWriter
// (siga_main.cpp), which is what zlib's and boost::iostreams' readers -- hence `siga assemble` -- accept.
#ifndef SGX_HOST_FASTGZ_H_
#define SGX_HOST_FASTGZ_H_

#include <stddef.h>

#include <string>

namespace sgxhost {

// One complete gzip member (header, deflate stream, CRC-32, size) that decodes to data[0, n).
std::string gzip_member_huffman(const char* data, size_t n);

// The same with a greedy single-probe LZ77 match finder in front (4..258-byte matches within 32 KB): for record text
// with repeated fields (the ED lines), where literal-only coding would give away a third of the ratio.
std::string gzip_member_lz(const char* data, size_t n);

}  // namespace sgxhost

#endif  // SGX_HOST_FASTGZ_H_
