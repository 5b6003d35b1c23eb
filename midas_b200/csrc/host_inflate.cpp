// Raw DEFLATE decoder for BGZF blocks (host side of the ingest, reference midas/run/snps.py:186 pysam.AlignmentFile ->
// htslib bgzf_read). The stage from a BAM on disk is bounded by the host's inflate once the counting is on the GPU;
// zlib's inflate() does ~290 MB/s per thread on BAM data. This decoder is written for the one case that matters here
// - a whole block of at most 64 KiB in memory, output size known - with a 64-bit bit buffer refilled eight bytes at a
// time, two-level decode tables (10 + up to 5 bits for literal/length codes, 8 + up to 7 for distances) whose entries
// carry base value and extra-bit count, and word-wise match copies. The caller verifies the block's CRC-32 (as it
// did with zlib) and falls back to zlib for any block this decoder refuses, so a decoding slip cannot go unnoticed.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

// (declared in common.h too; repeated here so that the file builds on its own for tests/native/inflate_harness.cpp)
namespace mb {
bool fast_inflate(const uint8_t* in, size_t in_len, size_t in_slack, uint8_t* out, size_t out_len);
}

namespace {

constexpr int LIT_BITS = 10, DIST_BITS = 8;
constexpr uint32_t F_LITERAL = 1u << 31, F_EOB = 1u << 30, F_SUB = 1u << 29, F_BAD = 1u << 28;
// entry: bits 0-7 code length (bits to consume; for F_SUB: primary bits), bits 8-15 extra bits (F_SUB: subtable bits),
//        bits 16-27 + flags: literal value / base value / subtable offset (base values up to 32768 need 16 bits: the
//        flags sit above bit 28, base in bits 12-27 for distances - see pack()).
inline uint32_t pack(uint32_t flags, uint32_t value, uint32_t extra, uint32_t len) {
  return flags | (value << 11) | (extra << 6) | len;      // len < 64, extra < 32, value < 65536 (bits 11-26)
}
inline uint32_t e_len(uint32_t e) { return e & 63u; }
inline uint32_t e_extra(uint32_t e) { return (e >> 6) & 31u; }
inline uint32_t e_value(uint32_t e) { return (e >> 11) & 0xffffu; }

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t bitrev(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

struct Tables {
  uint32_t lit[(1 << LIT_BITS) + 1024];     // primary + subtables (at most 286 long codes, each subtable <= 32 entries)
  uint32_t dist[(1 << DIST_BITS) + 512];
};

// Canonical Huffman table from code lengths (RFC 1951 3.2.2), LSB-first lookup. `is_lit`: symbols 0-255 literal, 256
// end of block, 257.. lengths; else distances. Returns false for an over-subscribed or (non-trivially) incomplete code.
bool build(const uint8_t* lens, int n_sym, bool is_lit, uint32_t* tab, int primary_bits, int tab_cap) {
  int count[16] = {0};
  for (int i = 0; i < n_sym; ++i) count[lens[i]]++;
  count[0] = 0;
  int left = 1, n_codes = 0;
  for (int l = 1; l <= 15; ++l) {
    left <<= 1;
    left -= count[l];
    if (left < 0) return false;
    n_codes += count[l];
  }
  const uint32_t bad = pack(F_BAD, 0, 0, 1);
  const int psize = 1 << primary_bits;
  for (int i = 0; i < psize; ++i) tab[i] = bad;
  if (n_codes == 0) return !is_lit;                      // no distance codes: a block of literals only
  if (left > 0 && !(n_codes == 1 && !is_lit)) return false;   // incomplete (one distance code of length 1 is allowed)
  uint32_t next_code[16];
  {
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
  }
  auto entry_of = [&](int sym, int len_field) -> uint32_t {
    if (is_lit) {
      if (sym < 256) return pack(F_LITERAL, (uint32_t)sym, 0, (uint32_t)len_field);
      if (sym == 256) return pack(F_EOB, 0, 0, (uint32_t)len_field);
      if (sym - 257 >= 29) return pack(F_BAD, 0, 0, (uint32_t)len_field);
      return pack(0, kLenBase[sym - 257], kLenExtra[sym - 257], (uint32_t)len_field);
    }
    if (sym >= 30) return pack(F_BAD, 0, 0, (uint32_t)len_field);
    return pack(0, kDistBase[sym], kDistExtra[sym], (uint32_t)len_field);
  };
  // pass 1: how many subtable bits every primary prefix needs
  int sub_bits[1 << LIT_BITS];
  for (int i = 0; i < psize; ++i) sub_bits[i] = 0;
  {
    uint32_t nc[16];
    memcpy(nc, next_code, sizeof(nc));
    for (int s = 0; s < n_sym; ++s) {
      const int l = lens[s];
      if (l <= primary_bits) { if (l) nc[l]++; continue; }
      const uint32_t rev = bitrev(nc[l]++, l);
      const int prefix = (int)(rev & (uint32_t)(psize - 1));
      if (l - primary_bits > sub_bits[prefix]) sub_bits[prefix] = l - primary_bits;
    }
  }
  int sub_off[1 << LIT_BITS];
  int used = psize;
  for (int i = 0; i < psize; ++i) {
    sub_off[i] = 0;
    if (sub_bits[i]) {
      sub_off[i] = used;
      used += 1 << sub_bits[i];
      if (used > tab_cap) return false;
      for (int k = 0; k < (1 << sub_bits[i]); ++k) tab[sub_off[i] + k] = bad;
      tab[i] = pack(F_SUB, (uint32_t)sub_off[i], (uint32_t)sub_bits[i], (uint32_t)primary_bits);
    }
  }
  // pass 2: fill
  for (int s = 0; s < n_sym; ++s) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t rev = bitrev(next_code[l]++, l);
    if (l <= primary_bits) {
      const uint32_t e = entry_of(s, l);
      for (uint32_t k = rev; k < (uint32_t)psize; k += 1u << l) tab[k] = e;
    } else {
      const int prefix = (int)(rev & (uint32_t)(psize - 1));
      const int sb = sub_bits[prefix], rest = l - primary_bits;
      const uint32_t e = entry_of(s, rest);
      for (uint32_t k = rev >> primary_bits; k < (1u << sb); k += 1u << rest) tab[sub_off[prefix] + k] = e;
    }
  }
  return true;
}

struct Fixed {
  Tables t;
  bool ok;
  Fixed() {
    uint8_t l[288];
    for (int i = 0; i < 144; ++i) l[i] = 8;
    for (int i = 144; i < 256; ++i) l[i] = 9;
    for (int i = 256; i < 280; ++i) l[i] = 7;
    for (int i = 280; i < 288; ++i) l[i] = 8;
    ok = build(l, 288, true, t.lit, LIT_BITS, (int)(sizeof(t.lit) / 4));
    uint8_t d[32];                                         // 32 five-bit codes; 30 and 31 never occur in valid data
    for (int i = 0; i < 32; ++i) d[i] = 5;
    ok = ok && build(d, 32, false, t.dist, DIST_BITS, (int)(sizeof(t.dist) / 4));
  }
};

inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

}  // namespace

namespace mb {

// Inflate one raw DEFLATE stream of `in_len` bytes into exactly `out_len` bytes. `in_slack`: bytes readable past the
// end of `in` (0-8; the caller knows what follows the block in its mapping). Returns true if the stream decoded
// cleanly to exactly out_len bytes; false = refused (the caller falls back to zlib). Never writes outside
// [out, out + out_len), never reads outside [in, in + in_len + in_slack).
bool fast_inflate(const uint8_t* in, size_t in_len, size_t in_slack, uint8_t* out, size_t out_len) {
  static const Fixed fixed;
  if (!fixed.ok) return false;
  const uint8_t* ip = in;
  const uint8_t* const in_end = in + in_len;
  const uint8_t* const in_fast_end = in_len + in_slack >= 8 ? in + (in_len + in_slack - 8) : in;   // load64 allowed at <= this
  uint8_t* op = out;
  uint8_t* const out_end = out + out_len;
  uint64_t bb = 0;
  int bc = 0;
  Tables dyn;
  size_t over = 0;                                          // zero bytes fed past the end of the input
  const bool fast_ok = in_len + in_slack >= 8;
  auto refill = [&]() {
    if (fast_ok && ip <= in_fast_end) {
      // eight bytes at once; the bits above `bc | 56` are loaded again (to the same positions) by the next refill
      bb |= load64(ip) << bc;
      ip += (63 - bc) >> 3;
      bc |= 56;
    } else {
      if (bc < 64) bb &= ((uint64_t)1 << bc) - 1;           // drop what a word-wise refill left above the valid bits
      while (bc <= 56) {
        if (ip < in_end) bb |= (uint64_t)(*ip++) << bc;
        else ++over;
        bc += 8;
      }
    }
  };
  // bytes of the last refill that lie past in_end do not count as consumed input: a valid stream never needs them,
  // a corrupt one is caught by the CRC / length checks
  for (;;) {
    refill();
    const uint32_t final_block = (uint32_t)(bb & 1u), type = (uint32_t)((bb >> 1) & 3u);
    bb >>= 3; bc -= 3;
    const Tables* T = nullptr;
    if (type == 0) {
      // stored: drop to the byte boundary, LEN / NLEN, copy
      const int drop = bc & 7;
      bb >>= drop; bc -= drop;
      // un-read the whole bytes still in the bit buffer
      size_t back = (size_t)(bc >> 3);
      if (over) { const size_t z = over < back ? over : back; back -= z; over -= z; }
      ip -= back;
      bb = 0; bc = 0;
      if (in_end - ip < 4) return false;
      const uint32_t len = (uint32_t)ip[0] | ((uint32_t)ip[1] << 8), nlen = (uint32_t)ip[2] | ((uint32_t)ip[3] << 8);
      ip += 4;
      if ((len ^ 0xffffu) != nlen || (size_t)(in_end - ip) < len || (size_t)(out_end - op) < len) return false;
      memcpy(op, ip, len);
      op += len; ip += len;
      if (final_block) break;
      continue;
    } else if (type == 1) {
      T = &fixed.t;
    } else if (type == 2) {
      const uint32_t hlit = (uint32_t)(bb & 31u) + 257, hdist = (uint32_t)((bb >> 5) & 31u) + 1, hclen = (uint32_t)((bb >> 10) & 15u) + 4;
      bb >>= 14; bc -= 14;
      if (hlit > 286 || hdist > 30) return false;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t cl[19] = {0};
      for (uint32_t i = 0; i < hclen; ++i) {
        if (bc < 3) refill();
        cl[order[i]] = (uint8_t)(bb & 7u);
        bb >>= 3; bc -= 3;
      }
      // the code-length code (at most 7 bits): one flat 128-entry table; it must be complete (zlib's rule)
      uint8_t clsym[128], cllen[128];
      {
        int count[8] = {0};
        for (int i = 0; i < 19; ++i) count[cl[i]]++;
        count[0] = 0;
        int left = 1;
        for (int l = 1; l <= 7; ++l) { left <<= 1; left -= count[l]; if (left < 0) return false; }
        if (left != 0) return false;
        uint32_t nc[9], code = 0;
        for (int l = 1; l <= 7; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; nc[l] = code; }
        memset(cllen, 0, sizeof(cllen));
        memset(clsym, 0, sizeof(clsym));
        for (int s = 0; s < 19; ++s) {
          const int l = cl[s];
          if (!l) continue;
          const uint32_t rev = bitrev(nc[l]++, l);
          for (uint32_t k = rev; k < 128; k += 1u << l) { clsym[k] = (uint8_t)s; cllen[k] = (uint8_t)l; }
        }
      }
      uint8_t lens[286 + 30 + 138];
      uint32_t n = 0;
      while (n < hlit + hdist) {
        if (bc < 14) refill();
        const uint32_t k = (uint32_t)(bb & 127u);
        const uint32_t l = cllen[k], s = clsym[k];
        if (!l) return false;
        bb >>= l; bc -= (int)l;
        if (s < 16) { lens[n++] = (uint8_t)s; continue; }
        uint32_t rep, val = 0;
        if (s == 16) { if (!n) return false; val = lens[n - 1]; rep = 3 + (uint32_t)(bb & 3u); bb >>= 2; bc -= 2; }
        else if (s == 17) { rep = 3 + (uint32_t)(bb & 7u); bb >>= 3; bc -= 3; }
        else { rep = 11 + (uint32_t)(bb & 127u); bb >>= 7; bc -= 7; }
        if (n + rep > hlit + hdist) return false;
        memset(lens + n, (int)val, rep);
        n += rep;
      }
      if (lens[256] == 0) return false;                      // no end-of-block code
      if (!build(lens, (int)hlit, true, dyn.lit, LIT_BITS, (int)(sizeof(dyn.lit) / 4))) return false;
      if (!build(lens + hlit, (int)hdist, false, dyn.dist, DIST_BITS, (int)(sizeof(dyn.dist) / 4))) return false;
      T = &dyn;
    } else {
      return false;
    }
    // ---- symbols
    const uint32_t* const lt = T->lit;
    const uint32_t* const dt = T->dist;
    bool eob = false;
    // fast loop: while a whole refill word can be read and the longest match plus the copy overshoot fits, nothing
    // inside needs a bounds check but the match distance
    constexpr uint32_t LMASK = (1u << LIT_BITS) - 1, DMASK = (1u << DIST_BITS) - 1;
    // Invariant at the top: the bit buffer has just been refilled (>= 56 bits) and `e` is the entry its low bits select,
    // looked up before the previous symbol's output was written - the table load runs under the copy.
    if (fast_ok && ip <= in_fast_end && (size_t)(out_end - op) >= 258 + 16) {
#define MB_REFILL() do { bb |= load64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; } while (0)
      MB_REFILL();
      uint32_t e = lt[bb & LMASK];
      while (ip <= in_fast_end && (size_t)(out_end - op) >= 258 + 16) {
        if (e & F_LITERAL) {                                 // up to three literals (<= 30 bits), then the next entry
          bb >>= e_len(e); bc -= (int)e_len(e);
          *op++ = (uint8_t)e_value(e);
          e = lt[bb & LMASK];
          if (e & F_LITERAL) {
            bb >>= e_len(e); bc -= (int)e_len(e);
            *op++ = (uint8_t)e_value(e);
            e = lt[bb & LMASK];
            if (e & F_LITERAL) {
              bb >>= e_len(e); bc -= (int)e_len(e);
              *op++ = (uint8_t)e_value(e);
              e = lt[bb & LMASK];
            }
          }
          MB_REFILL();                                       // (the low bits `e` was selected by are unchanged)
          continue;
        }
        if (e & (F_SUB | F_EOB | F_BAD)) {                   // the rare cases behind one test: a length code is none of them
          if (e & F_SUB) {
            bb >>= LIT_BITS; bc -= LIT_BITS;
            e = lt[e_value(e) + (uint32_t)(bb & ((1u << e_extra(e)) - 1))];
            if (e & F_LITERAL) {
              bb >>= e_len(e); bc -= (int)e_len(e);
              *op++ = (uint8_t)e_value(e);
              e = lt[bb & LMASK];
              MB_REFILL();
              continue;
            }
          }
          if (e & (F_EOB | F_BAD)) {
            if (e & F_BAD) return false;
            bb >>= e_len(e); bc -= (int)e_len(e);
            eob = true;
            break;
          }
        }
        bb >>= e_len(e); bc -= (int)e_len(e);
        const uint32_t length = e_value(e) + (uint32_t)(bb & ((1u << e_extra(e)) - 1));
        bb >>= e_extra(e); bc -= (int)e_extra(e);
        uint32_t d = dt[bb & DMASK];
        if (d & (F_SUB | F_BAD)) {
          if (d & F_SUB) {
            bb >>= DIST_BITS; bc -= DIST_BITS;
            d = dt[e_value(d) + (uint32_t)(bb & ((1u << e_extra(d)) - 1))];
          }
          if (d & F_BAD) return false;
        }
        bb >>= e_len(d); bc -= (int)e_len(d);
        const uint32_t dist = e_value(d) + (uint32_t)(bb & ((1u << e_extra(d)) - 1));
        bb >>= e_extra(d); bc -= (int)e_extra(d);
        MB_REFILL();
        e = lt[bb & LMASK];                                  // the next symbol's entry: in flight during the copy
        if (dist > (size_t)(op - out)) return false;
        const uint8_t* src = op - dist;
        uint8_t* const stop = op + length;
        if (dist >= 8) {
          memcpy(op, src, 8); memcpy(op + 8, src + 8, 8);    // most matches are short: two words, then the rest
          if (length > 16) {
            op += 16; src += 16;
            do { memcpy(op, src, 8); op += 8; src += 8; } while (op < stop);
          }
        } else if (dist == 1) {
          uint64_t v = *src;
          v *= 0x0101010101010101ull;
          do { memcpy(op, &v, 8); op += 8; } while (op < stop);
        } else {
          do { *op++ = *src++; } while (op < stop);
        }
        op = stop;
      }
#undef MB_REFILL
    }
    if (!eob && bc < 64) bb &= ((uint64_t)1 << bc) - 1;      // the careful loop below ORs bytes in: clear what lies above
    if (!eob) for (;;) {
      refill();                                              // >= 56 bits: one literal/length (<= 20) + one distance (<= 28)
      uint32_t e = lt[bb & ((1u << LIT_BITS) - 1)];
      if (e & F_SUB) {
        bb >>= LIT_BITS; bc -= LIT_BITS;
        e = lt[e_value(e) + (uint32_t)(bb & ((1u << e_extra(e)) - 1))];
      }
      bb >>= e_len(e); bc -= (int)e_len(e);
      if (e & F_LITERAL) {
        if (op >= out_end) return false;
        *op++ = (uint8_t)e_value(e);
        // a second and third literal out of the same refill (<= 45 bits used)
        e = lt[bb & ((1u << LIT_BITS) - 1)];
        if ((e & (F_LITERAL | F_SUB)) == F_LITERAL && op < out_end) {
          bb >>= e_len(e); bc -= (int)e_len(e);
          *op++ = (uint8_t)e_value(e);
          e = lt[bb & ((1u << LIT_BITS) - 1)];
          if ((e & (F_LITERAL | F_SUB)) == F_LITERAL && op < out_end) {
            bb >>= e_len(e); bc -= (int)e_len(e);
            *op++ = (uint8_t)e_value(e);
          }
        }
        continue;
      }
      if (e & (F_EOB | F_BAD)) {
        if (e & F_BAD) return false;
        break;
      }
      const uint32_t length = e_value(e) + (uint32_t)(bb & ((1u << e_extra(e)) - 1));
      bb >>= e_extra(e); bc -= (int)e_extra(e);
      uint32_t d = dt[bb & ((1u << DIST_BITS) - 1)];
      if (d & F_SUB) {
        bb >>= DIST_BITS; bc -= DIST_BITS;
        d = dt[e_value(d) + (uint32_t)(bb & ((1u << e_extra(d)) - 1))];
      }
      if (d & F_BAD) return false;
      bb >>= e_len(d); bc -= (int)e_len(d);
      const uint32_t dist = e_value(d) + (uint32_t)(bb & ((1u << e_extra(d)) - 1));
      bb >>= e_extra(d); bc -= (int)e_extra(d);
      if (bc < 0) return false;                              // ran past the bits that were there
      if (dist > (size_t)(op - out) || length > (size_t)(out_end - op)) return false;
      const uint8_t* src = op - dist;
      if (dist >= 8 && (size_t)(out_end - op) >= length + 8) {
        uint8_t* const stop = op + length;                   // word-wise, may write up to 7 bytes past `stop` (inside the block)
        do { memcpy(op, src, 8); op += 8; src += 8; } while (op < stop);
        op = stop;
      } else if (dist == 1) {
        memset(op, *src, length);
        op += length;
      } else {
        for (uint32_t k = 0; k < length; ++k) op[k] = src[k];
        op += length;
      }
    }
    if (final_block) break;
  }
  if (op != out_end) return false;
  // input accounting: whole unread bytes in the bit buffer are handed back; what was consumed must lie inside the
  // stream (the fast refill may have read the bytes behind it, the slow one fed zeros: neither may have been used)
  const size_t back = (size_t)(bc >> 3);
  const size_t taken = (size_t)(ip - in) + over;
  if (taken < back || taken - back > in_len) return false;
  return true;
}

}  // namespace mb
