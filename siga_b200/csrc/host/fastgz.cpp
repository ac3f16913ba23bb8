// Literal-only gzip writer of the `siga` front-end (see fastgz.h).
//
// RFC 1951 dynamic-Huffman blocks (BTYPE = 2) that contain literals and the end-of-block symbol only: per block one
// byte histogram, length-limited canonical Huffman codes, then one table look-up and one bit append per input byte.
// No hash chains, no match search: DNA text has no repeats a level-1 matcher finds, so the output is within 2 % of
// zlib level 1 at several times its speed; any gzip reader (zlib, boost::iostreams, `siga assemble`) decodes it.
#include "fastgz.h"

#include <zlib.h>

#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <vector>

namespace sgxhost {

namespace {

const size_t kBlock = 256 * 1024;  // input bytes per deflate block (one Huffman table each)

// Huffman code lengths of `n` symbols with frequencies `freq`, none longer than `limit`; unused symbols get length 0.
// Plain Huffman (two-queue merge over the sorted used symbols), then the overflow fix-up every deflate writer uses:
// lengths above the limit are folded into it and the Kraft sum is repaired by lengthening the cheapest shorter codes.
void code_lengths(const uint32_t* freq, int n, int limit, uint8_t* len) {
  struct Node {
    uint64_t w;
    int left, right;  // children (-1: leaf)
  };
  std::vector<int> used;
  for (int i = 0; i < n; ++i) {
    len[i] = 0;
    if (freq[i]) used.push_back(i);
  }
  if (used.empty()) return;
  if (used.size() == 1) {  // a lone symbol still needs a one-bit code
    len[used[0]] = 1;
    return;
  }
  std::sort(used.begin(), used.end(), [&](int a, int b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
  const int m = (int)used.size();
  std::vector<Node> nodes(2 * m);
  for (int i = 0; i < m; ++i) nodes[i] = Node{freq[used[i]], -1, -1};
  int leaf = 0, inner = m, next = m;  // queues: leaves [leaf, m), inner nodes [inner, next)
  auto pop = [&]() {
    if (leaf < m && (inner >= next || nodes[leaf].w <= nodes[inner].w)) return leaf++;
    return inner++;
  };
  while ((m - leaf) + (next - inner) > 1) {
    const int a = pop(), b = pop();
    nodes[next] = Node{nodes[a].w + nodes[b].w, a, b};
    ++next;
  }
  // depths, root = nodes[next - 1]
  std::vector<int> depth(next, 0);
  for (int i = next - 1; i >= m; --i) {
    depth[nodes[i].left] = depth[i] + 1;
    depth[nodes[i].right] = depth[i] + 1;
  }
  std::vector<int> count(64, 0);
  for (int i = 0; i < m; ++i) count[std::min(depth[i], 63)]++;
  // fold the over-long codes into `limit` and repair the Kraft sum
  for (int d = limit + 1; d < 64; ++d) {
    count[limit] += count[d];
    count[d] = 0;
  }
  uint64_t total = 0;
  for (int d = limit; d >= 1; --d) total += (uint64_t)count[d] << (limit - d);
  while (total != (1ull << limit)) {
    count[limit]--;
    for (int d = limit - 1; d >= 1; --d)
      if (count[d]) {
        count[d]--;
        count[d + 1] += 2;
        break;
      }
    total--;
  }
  // rarest symbols get the longest codes: `used` is sorted by rising frequency
  int k = 0;
  for (int d = limit; d >= 1; --d)
    for (int c = 0; c < count[d]; ++c) len[used[k++]] = (uint8_t)d;
}

// Canonical codes of RFC 1951 section 3.2.2, bit-reversed so that they can be appended LSB first.
void canonical_codes(const uint8_t* len, int n, uint16_t* code) {
  int bl_count[16] = {0};
  for (int i = 0; i < n; ++i) bl_count[len[i]]++;
  bl_count[0] = 0;
  int next_code[16] = {0};
  int c = 0;
  for (int b = 1; b < 16; ++b) {
    c = (c + bl_count[b - 1]) << 1;
    next_code[b] = c;
  }
  for (int i = 0; i < n; ++i) {
    code[i] = 0;
    if (!len[i]) continue;
    unsigned v = (unsigned)next_code[len[i]]++, r = 0;
    for (int b = 0; b < len[i]; ++b) r |= ((v >> b) & 1u) << (len[i] - 1 - b);
    code[i] = (uint16_t)r;
  }
}

// Appends bit strings LSB first into a caller-sized buffer: whole bytes of the 64-bit accumulator are stored with one
// unaligned 8-byte write per append (the bytes beyond the valid ones are overwritten by the next append).
struct BitWriter {
  unsigned char* p;
  uint64_t acc;
  int n;  // valid bits in acc, < 8 between appends
  explicit BitWriter(unsigned char* dst) : p(dst), acc(0), n(0) {}
  inline void put(uint64_t v, int bits) {  // bits <= 56
    acc |= v << n;
    n += bits;
    memcpy(p, &acc, 8);  // little-endian host
    p += n >> 3;
    acc >>= n & ~7;
    n &= 7;
  }
  unsigned char* finish() {  // pads the last byte with zero bits
    if (n > 0) {
      *p++ = (unsigned char)(acc & 0xff);
      acc = 0;
      n = 0;
    }
    return p;
  }
};

// Code tables of one block and the exact number of bits it will take.
struct BlockPlan {
  uint8_t len[258];   // literal / length code lengths (257 used) + the one distance code (length 0: no matches)
  uint16_t code[257];
  uint8_t cllen[19];  // code-length alphabet
  uint16_t clcode[19];
  uint64_t bits;
};

void plan_block(const unsigned char* p, size_t n, BlockPlan* bp) {
  // four interleaved histograms: consecutive equal bytes (homopolymers) would otherwise serialise on one counter
  uint32_t h4[4][256];
  memset(h4, 0, sizeof(h4));
  size_t j = 0;
  for (; j + 4 <= n; j += 4) {
    h4[0][p[j]]++;
    h4[1][p[j + 1]]++;
    h4[2][p[j + 2]]++;
    h4[3][p[j + 3]]++;
  }
  for (; j < n; ++j) h4[0][p[j]]++;
  uint32_t freq[257];
  for (int s = 0; s < 256; ++s) freq[s] = h4[0][s] + h4[1][s] + h4[2][s] + h4[3][s];
  freq[256] = 1;  // end of block
  code_lengths(freq, 257, 15, bp->len);
  canonical_codes(bp->len, 257, bp->code);
  bp->len[257] = 0;
  // the 258 code lengths are sent one by one with the code-length alphabet (no repeat symbols: 258 x ~4 bits per 256 KB)
  uint32_t clfreq[19];
  memset(clfreq, 0, sizeof(clfreq));
  for (int i = 0; i < 258; ++i) clfreq[bp->len[i]]++;
  code_lengths(clfreq, 19, 7, bp->cllen);
  canonical_codes(bp->cllen, 19, bp->clcode);
  uint64_t bits = 3 + 5 + 5 + 4 + 19 * 3;
  for (int i = 0; i < 258; ++i) bits += bp->cllen[bp->len[i]];
  for (int s = 0; s < 257; ++s) bits += (uint64_t)freq[s] * bp->len[s];
  bp->bits = bits;
}

// G literals per append; the caller picks G so that G code words fit into 56 bits.
template <int G>
size_t encode_groups(BitWriter& bw, const unsigned char* p, size_t n, const uint32_t* tab) {
  size_t i = 0;
  for (; i + G <= n; i += G) {
    uint64_t v = 0;
    int nb = 0;
#pragma GCC unroll 8
    for (int t = 0; t < G; ++t) {
      const uint32_t e = tab[p[i + t]];
      v |= (uint64_t)(e & 0xffffu) << nb;
      nb += (int)(e >> 16);
    }
    bw.put(v, nb);
  }
  return i;
}

void write_block(BitWriter& bw, const unsigned char* p, size_t n, bool last, const BlockPlan& bp) {
  static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  bw.put(last ? 1u : 0u, 1);
  bw.put(2u, 2);         // BTYPE = dynamic Huffman
  bw.put(257 - 257, 5);  // HLIT
  bw.put(1 - 1, 5);      // HDIST: one distance code (of zero bits)
  bw.put(19 - 4, 4);     // HCLEN: all 19 code-length code lengths
  for (int i = 0; i < 19; ++i) bw.put(bp.cllen[order[i]], 3);
  for (int i = 0; i < 258; ++i) bw.put(bp.clcode[bp.len[i]], bp.cllen[bp.len[i]]);
  // payload: as many literals per append as fit into 56 bits at this block's longest code
  uint32_t tab[256];  // code | length << 16
  int maxlen = 1;
  for (int s = 0; s < 256; ++s) {
    tab[s] = (uint32_t)bp.code[s] | ((uint32_t)bp.len[s] << 16);
    if (bp.len[s] > maxlen) maxlen = bp.len[s];
  }
  size_t i;
  if (maxlen <= 7) i = encode_groups<8>(bw, p, n, tab);
  else if (maxlen <= 9) i = encode_groups<6>(bw, p, n, tab);
  else if (maxlen <= 11) i = encode_groups<5>(bw, p, n, tab);
  else if (maxlen <= 14) i = encode_groups<4>(bw, p, n, tab);
  else i = encode_groups<3>(bw, p, n, tab);
  for (; i < n; ++i) bw.put(bp.code[p[i]], bp.len[p[i]]);
  bw.put(bp.code[256], bp.len[256]);
}

}  // namespace

std::string gzip_member_huffman(const char* data, size_t n) {
  const unsigned char* p = (const unsigned char*)data;
  const size_t n_blocks = (n + kBlock - 1) / kBlock;
  std::vector<BlockPlan> plan(n_blocks);
  uint64_t bits = n ? 0 : 10;  // empty input: one final block with the fixed code (BFINAL = 1, BTYPE = 01, end-of-block)
  for (size_t b = 0; b < n_blocks; ++b) {
    plan_block(p + b * kBlock, std::min(kBlock, n - b * kBlock), &plan[b]);
    bits += plan[b].bits;
  }
  const size_t body = (size_t)((bits + 7) / 8);
  std::string out(10 + body + 8 + 8, '\0');  // header, deflate stream, trailer, slack for the writer's 8-byte stores
  static const unsigned char header[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 4 /* fastest */, 3 /* unix */};
  memcpy(&out[0], header, 10);
  BitWriter bw((unsigned char*)&out[10]);
  if (n == 0) {
    bw.put(1u, 1);
    bw.put(1u, 2);
    bw.put(0u, 7);
  }
  for (size_t b = 0; b < n_blocks; ++b) {
    const size_t m = std::min(kBlock, n - b * kBlock);
    write_block(bw, p + b * kBlock, m, b + 1 == n_blocks, plan[b]);
  }
  unsigned char* end = bw.finish();
  if ((size_t)(end - (unsigned char*)&out[10]) != body) return std::string();  // cannot happen: sizes are computed from the same tables
  uLong crc = crc32(0L, Z_NULL, 0);
  for (size_t b = 0; b < n; b += (size_t)1 << 30) crc = crc32(crc, p + b, (uInt)std::min<size_t>((size_t)1 << 30, n - b));
  const uint32_t tail[2] = {(uint32_t)crc, (uint32_t)(n & 0xffffffffu)};
  memcpy(end, tail, 8);
  out.resize(10 + body + 8);
  return out;
}

}  // namespace sgxhost
