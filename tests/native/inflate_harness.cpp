// Harness of csrc/host_inflate.cpp (built by tests/test_inflate_native.py with -fsanitize=address,undefined):
// 1. raw DEFLATE streams made by zlib at every level from four kinds of data and five sizes must decode to the
//    source bytes, with guard bytes on both sides of the output untouched and the input buffer exactly as long as
//    the stream plus the declared slack (so that any over-read is an ASAN error);
// 2. 20 000 corrupted streams (bit flips, truncation, wrong output length) must be refused or decode to something
//    (the caller's CRC check judges that) without touching the guards.
#include <zlib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <stdint.h>
namespace mb { bool fast_inflate(const uint8_t* in, size_t in_len, size_t in_slack, uint8_t* out, size_t out_len); }
int main() {

  int bad = 0, tot = 0;
  for (int level = 0; level <= 9; ++level) for (int kind = 0; kind < 4; ++kind) for (int sz : {1, 10, 300, 5000, 65280}) {
    std::vector<uint8_t> src(sz);
    srand(level * 100 + kind * 10 + sz);
    for (int i = 0; i < sz; ++i) {
      if (kind == 0) src[i] = rand() & 255;
      else if (kind == 1) src[i] = "ACGT"[rand() & 3];
      else if (kind == 2) src[i] = (i % 97 < 50) ? 'A' + (i % 7) : rand() & 15;
      else src[i] = 'x';
    }
    std::vector<uint8_t> comp(sz * 2 + 1024);
    z_stream zs; memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, kind == 3 && level == 5 ? Z_FIXED : Z_DEFAULT_STRATEGY);
    zs.next_in = src.data(); zs.avail_in = sz; zs.next_out = comp.data(); zs.avail_out = comp.size();
    deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out; deflateEnd(&zs);
    for (size_t slack : {(size_t)0, (size_t)8}) {
      std::vector<uint8_t> in(clen + slack, 0xAB);
      memcpy(in.data(), comp.data(), clen);
      std::vector<uint8_t> out(sz + 16, 0xCD);
      bool ok = mb::fast_inflate(in.data(), clen, slack, out.data() + 8, sz);
      bool same = ok && !memcmp(out.data() + 8, src.data(), sz);
      bool guard = true; for (int i = 0; i < 8; ++i) guard = guard && out[i] == 0xCD && out[8 + sz + i] == 0xCD;
      ++tot;
      if (!same || !guard) { ++bad; if (bad < 12) printf("FAIL level %d kind %d sz %d slack %zu ok=%d guard=%d clen=%zu\n", level, kind, sz, slack, ok, guard, clen); }
    }
  }
    printf("valid streams: %d of %d failed\n", bad, tot);
  // 1. corrupted streams: never crash, never touch the guards
  int crashes = 0, accepted = 0, n = 0;
  for (int it = 0; it < 20000; ++it) {
    srand(it);
    int sz = 1 + rand() % 20000;
    std::vector<uint8_t> src(sz);
    int kind = rand() % 3;
    for (int i = 0; i < sz; ++i) src[i] = kind == 0 ? rand() & 255 : (kind == 1 ? "ACGT"[rand() & 3] : ((i % 53 < 30) ? 'A' + i % 5 : rand() & 7));
    std::vector<uint8_t> comp(sz * 2 + 1024);
    z_stream zs; memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, rand() % 10, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = src.data(); zs.avail_in = sz; zs.next_out = comp.data(); zs.avail_out = comp.size();
    deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out; deflateEnd(&zs);
    // corrupt: flip bits, truncate, or wrong out_len
    std::vector<uint8_t> in(comp.begin(), comp.begin() + clen);
    int mode = rand() % 4;
    size_t out_len = sz;
    if (mode == 0) for (int k = 0; k < 1 + rand() % 4; ++k) in[rand() % clen] ^= 1 << (rand() % 8);
    else if (mode == 1) in.resize(rand() % clen);
    else if (mode == 2) out_len = sz > 3 ? sz - 1 - rand() % 3 : sz + 1;
    else out_len = sz + 1 + rand() % 5;
    size_t slack = (rand() & 1) ? 8 : 0;
    size_t il = in.size();
    in.resize(il + slack, 0x5A);
    std::vector<uint8_t> out(out_len + 32, 0xCD);
    bool ok = mb::fast_inflate(in.data(), il, slack, out.data() + 16, out_len);
    for (int i = 0; i < 16; ++i) if (out[i] != 0xCD || out[16 + out_len + i] != 0xCD) { ++crashes; break; }
    accepted += ok; ++n;
  }
  printf("corrupted: %d runs, %d guard violations, %d accepted (CRC would judge those)\n", n, crashes, accepted);
  printf("corrupted streams: %d runs, %d guard violations, %d accepted\n", n, crashes, accepted);
  return bad != 0 || crashes != 0;
}
