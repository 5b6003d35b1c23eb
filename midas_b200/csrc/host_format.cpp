// Text rows of <species>.snps.gz (reference midas/run/snps.py:179-182 header, :201-210 rows):
//   ref_id \t ref_pos (1-based) \t ref_allele \t depth \t count_a \t count_c \t count_g \t count_t \n
// The reference builds each row with '\t'.join(str(x)) in a per-position Python loop; integers
// print the same in any language, so this formatter only has to be fast. Gzip framing is left to
// the caller (any member layout is valid; only the decompressed bytes are compared).
#include <cstring>

#include "common.h"

namespace {

inline char* put_u32(char* p, uint32_t v) {
  char tmp[10];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

inline char* put_u64(char* p, uint64_t v) {
  char tmp[20];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

}  // namespace

extern "C" uint64_t mb_format_snps_bound(uint64_t name_len, uint64_t n_rows) {
  return n_rows * (name_len + 1 + 20 + 1 + 1 + 1 + 5 * 11 + 1);
}

extern "C" int64_t mb_format_snps_rows(const char* contig_name, const uint8_t* ref_bases, uint64_t first_pos0,
                                       uint64_t n_rows, const uint32_t* count_a, const uint32_t* count_c,
                                       const uint32_t* count_g, const uint32_t* count_t, char* out, uint64_t cap) {
  if (!contig_name || !ref_bases || !count_a || !count_c || !count_g || !count_t || !out)
    return mb::fail(MB_EINVAL, "mb_format_snps_rows: null argument");
  const size_t nl = strlen(contig_name);
  if (cap < mb_format_snps_bound(nl, n_rows))
    return mb::fail(MB_EINVAL, "mb_format_snps_rows: output buffer smaller than mb_format_snps_bound");
  char* p = out;
  for (uint64_t i = 0; i < n_rows; ++i) {
    memcpy(p, contig_name, nl);
    p += nl;
    *p++ = '\t';
    p = put_u64(p, first_pos0 + i + 1);                                  // ref_pos = i+1 (snps.py:202)
    *p++ = '\t';
    *p++ = (char)ref_bases[i];                                           // ref_allele (snps.py:203)
    *p++ = '\t';
    const uint32_t a = count_a[i], c = count_c[i], g = count_g[i], t = count_t[i];
    p = put_u64(p, (uint64_t)a + c + g + t);                             // depth (snps.py:204)
    *p++ = '\t';
    p = put_u32(p, a);
    *p++ = '\t';
    p = put_u32(p, c);
    *p++ = '\t';
    p = put_u32(p, g);
    *p++ = '\t';
    p = put_u32(p, t);
    *p++ = '\n';
  }
  return (int64_t)(p - out);
}
