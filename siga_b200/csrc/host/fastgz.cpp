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

// ---------------------------------------------------------------------------------------------------------
// Greedy single-probe LZ77 for record text with repeated fields (the ED lines of an ASQG file): one hash-table probe
// per position, matches of 4..258 bytes within the 32 KB window, no lazy evaluation, no chains.
// ---------------------------------------------------------------------------------------------------------
const int kHashBits = 15;
const uint32_t kNoPosition = 0xffffffffu;

inline uint32_t load32(const unsigned char* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}

// token: literal byte, or 0x80000000 | (length - 3) << 15 | (distance - 1)
inline void length_symbol_slow(uint32_t l3, int* sym, int* nbits, uint32_t* extra) {  // l3 = length - 3 in [0, 255]
  if (l3 < 8) {
    *sym = 257 + (int)l3;
    *nbits = 0;
    *extra = 0;
  } else if (l3 == 255) {
    *sym = 285;
    *nbits = 0;
    *extra = 0;
  } else {
    const int nb = 31 - __builtin_clz(l3);
    *sym = 257 + 4 * (nb - 1) + (int)((l3 >> (nb - 2)) & 3u);
    *nbits = nb - 2;
    *extra = l3 & ((1u << (nb - 2)) - 1u);
  }
}
struct LengthTable {
  uint16_t sym[256];
  uint8_t nbits[256];
  uint8_t extra[256];
  LengthTable() {
    for (uint32_t l3 = 0; l3 < 256; ++l3) {
      int s, nb;
      uint32_t ex;
      length_symbol_slow(l3, &s, &nb, &ex);
      sym[l3] = (uint16_t)s;
      nbits[l3] = (uint8_t)nb;
      extra[l3] = (uint8_t)ex;
    }
  }
};
const LengthTable kLength;
inline void length_symbol(uint32_t l3, int* sym, int* nbits, uint32_t* extra) {
  *sym = kLength.sym[l3];
  *nbits = kLength.nbits[l3];
  *extra = kLength.extra[l3];
}
inline void distance_symbol(uint32_t d1, int* sym, int* nbits, uint32_t* extra) {  // d1 = distance - 1 in [0, 32767]
  if (d1 < 4) {
    *sym = (int)d1;
    *nbits = 0;
    *extra = 0;
  } else {
    const int nb = 31 - __builtin_clz(d1);
    *sym = 2 * nb + (int)((d1 >> (nb - 1)) & 1u);
    *nbits = nb - 1;
    *extra = d1 & ((1u << (nb - 1)) - 1u);
  }
}

struct LzBlock {
  std::vector<uint32_t> tokens;  // kBlock + 1 slots, n_tokens used
  size_t n_tokens;
  uint32_t lfreq[286], dfreq[30];
  uint64_t extra_bits;
};

// Parses data[from, to) into tokens; matches may reach back into data[0, from) (one deflate stream per member).
void lz_parse(const unsigned char* data, size_t from, size_t to, size_t n, uint32_t* table, LzBlock* blk) {
  if (blk->tokens.size() < kBlock + 1) blk->tokens.resize(kBlock + 1);
  uint32_t* tok = blk->tokens.data();
  size_t nt = 0;
  memset(blk->lfreq, 0, sizeof(blk->lfreq));
  memset(blk->dfreq, 0, sizeof(blk->dfreq));
  blk->extra_bits = 0;
  size_t i = from;
  while (i < to) {
    if (i + 4 <= n) {
      const uint32_t v = load32(data + i);
      const uint32_t h = (v * 2654435761u) >> (32 - kHashBits);
      const uint32_t cand = table[h];
      table[h] = (uint32_t)i;
      if (cand != kNoPosition && i - cand <= 32768 && load32(data + cand) == v) {
        const size_t limit = std::min<size_t>(258, to - i);  // a match does not cross the block end
        size_t len = 4;
        while (len + 8 <= limit) {  // eight bytes per comparison
          uint64_t x, y;
          memcpy(&x, data + cand + len, 8);
          memcpy(&y, data + i + len, 8);
          if (x != y) {
            len += (size_t)(__builtin_ctzll(x ^ y) >> 3);
            goto extended;
          }
          len += 8;
        }
        while (len < limit && data[cand + len] == data[i + len]) ++len;
      extended:
        if (len <= limit) {
          int sym, nb;
          uint32_t ex;
          length_symbol((uint32_t)len - 3, &sym, &nb, &ex);
          blk->lfreq[sym]++;
          blk->extra_bits += (uint64_t)nb;
          distance_symbol((uint32_t)(i - cand) - 1, &sym, &nb, &ex);
          blk->dfreq[sym]++;
          blk->extra_bits += (uint64_t)nb;
          tok[nt++] = 0x80000000u | (((uint32_t)len - 3) << 15) | ((uint32_t)(i - cand) - 1);
          // index a second position inside the match (cheap, and it is what later records tend to repeat from)
          if (i + len / 2 + 4 <= n) table[(load32(data + i + len / 2) * 2654435761u) >> (32 - kHashBits)] = (uint32_t)(i + len / 2);
          i += len;
          continue;
        }
      }
    }
    blk->lfreq[data[i]]++;
    tok[nt++] = data[i];
    ++i;
  }
  blk->lfreq[256]++;
  blk->n_tokens = nt;
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

namespace sgxhost {

std::string gzip_member_lz(const char* data, size_t n) {
  const unsigned char* p = (const unsigned char*)data;
  std::string out(10 + n / 3 + 65536, '\0');
  static const unsigned char header[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 4 /* fastest */, 3 /* unix */};
  memcpy(&out[0], header, 10);
  size_t pos = 10;  // bytes of `out` that are final
  uint64_t pend_acc = 0;
  int pend_n = 0;   // bits of the last, partial byte carried from block to block
  std::vector<uint32_t> table((size_t)1 << kHashBits, kNoPosition);
  LzBlock blk;
  static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  const size_t n_blocks = n ? (n + kBlock - 1) / kBlock : 1;
  for (size_t b = 0; b < n_blocks; ++b) {
    const size_t from = b * kBlock, to = std::min(n, from + kBlock);
    lz_parse(p, from, to, n, table.data(), &blk);
    uint8_t llen[286], dlen[30], cllen[19];
    uint16_t lcode[286], dcode[30], clcode[19];
    code_lengths(blk.lfreq, 286, 15, llen);
    canonical_codes(llen, 286, lcode);
    code_lengths(blk.dfreq, 30, 15, dlen);
    canonical_codes(dlen, 30, dcode);
    uint32_t clfreq[19];
    memset(clfreq, 0, sizeof(clfreq));
    for (int i = 0; i < 286; ++i) clfreq[llen[i]]++;
    for (int i = 0; i < 30; ++i) clfreq[dlen[i]]++;
    code_lengths(clfreq, 19, 7, cllen);
    canonical_codes(cllen, 19, clcode);
    uint64_t bits = 3 + 5 + 5 + 4 + 19 * 3 + blk.extra_bits;
    for (int i = 0; i < 286; ++i) bits += cllen[llen[i]] + (uint64_t)blk.lfreq[i] * llen[i];
    for (int i = 0; i < 30; ++i) bits += cllen[dlen[i]] + (uint64_t)blk.dfreq[i] * dlen[i];
    const size_t need = pos + (size_t)((bits + pend_n + 7) / 8) + 16;
    if (out.size() < need) out.resize(std::max(need, out.size() + out.size() / 2));
    BitWriter bw((unsigned char*)&out[pos]);
    bw.acc = pend_acc;
    bw.n = pend_n;
    bw.put(b + 1 == n_blocks ? 1u : 0u, 1);
    bw.put(2u, 2);
    bw.put(286 - 257, 5);
    bw.put(30 - 1, 5);
    bw.put(19 - 4, 4);
    for (int i = 0; i < 19; ++i) bw.put(cllen[order[i]], 3);
    for (int i = 0; i < 286; ++i) bw.put(clcode[llen[i]], cllen[llen[i]]);
    for (int i = 0; i < 30; ++i) bw.put(clcode[dlen[i]], cllen[dlen[i]]);
    for (size_t ti = 0; ti < blk.n_tokens; ++ti) {
      const uint32_t t = blk.tokens[ti];
      if (!(t & 0x80000000u)) {
        bw.put(lcode[t], llen[t]);
      } else {
        int sym, nb, dsym, dnb;
        uint32_t ex, dex;
        length_symbol((t >> 15) & 0xffu, &sym, &nb, &ex);
        distance_symbol(t & 0x7fffu, &dsym, &dnb, &dex);
        // length code + extra (<= 20 bits), distance code + extra (<= 28 bits): 48 bits at most in one append
        uint64_t v = lcode[sym];
        int k = llen[sym];
        v |= (uint64_t)ex << k;
        k += nb;
        v |= (uint64_t)dcode[dsym] << k;
        k += dlen[dsym];
        v |= (uint64_t)dex << k;
        k += dnb;
        bw.put(v, k);
      }
    }
    bw.put(lcode[256], llen[256]);
    pos = (size_t)(bw.p - (unsigned char*)&out[0]);
    pend_acc = bw.acc;
    pend_n = bw.n;
  }
  if (pend_n > 0) out[pos++] = (char)(pend_acc & 0xff);
  uLong crc = crc32(0L, Z_NULL, 0);
  for (size_t b = 0; b < n; b += (size_t)1 << 30) crc = crc32(crc, p + b, (uInt)std::min<size_t>((size_t)1 << 30, n - b));
  const uint32_t tail[2] = {(uint32_t)crc, (uint32_t)(n & 0xffffffffu)};
  if (out.size() < pos + 8) out.resize(pos + 8);
  memcpy(&out[pos], tail, 8);
  out.resize(pos + 8);
  return out;
}

}  // namespace sgxhost
