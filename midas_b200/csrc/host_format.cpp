// Text rows of <species>.snps.gz (reference midas/run/snps.py:179-182 header, :201-210 rows):
//   ref_id \t ref_pos (1-based) \t ref_allele \t depth \t count_a \t count_c \t count_g \t count_t \n
// The reference builds each row with '\t'.join(str(x)) in a per-position Python loop; integers
// print the same in any language, so this formatter only has to be fast. Gzip framing is left to
// the caller (any member layout is valid; only the decompressed bytes are compared).
#include <cmath>
#include <cstdio>
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

// ---- rows of snps_freq.txt / snps_depth.txt of merge_midas.py snps (reference midas/merge/snps.py:196-201):
//   site_id \t v(sample 0) \t v(sample 1) ... \n   with '{0:.3g}'.format(maf) resp. str(depth).
// The values stay where the kernel wrote them (sample-major, `stride` sites per sample): row r is site
// site_index[r] (0-based; the id printed is site_index[r] + 1, snps.py:352). printf's %.3g is the same
// correctly rounded three-significant-digit conversion as Python's format (two-digit exponents included).
extern "C" uint64_t mb_format_merge_bound(uint64_t n_rows, uint32_t n_samples) {
  return n_rows * (21 + (uint64_t)n_samples * 17 + 1) + 32;
}

extern "C" int64_t mb_format_merge_freq_rows(const uint64_t* site_index, uint64_t n_rows, uint32_t n_samples,
                                             uint64_t stride, const double* maf, uint64_t id_base, char* out,
                                             uint64_t cap) {
  if (!site_index || !maf || !out) return mb::fail(MB_EINVAL, "mb_format_merge_freq_rows: null argument");
  if (cap < mb_format_merge_bound(n_rows, n_samples))
    return mb::fail(MB_EINVAL, "mb_format_merge_freq_rows: output buffer smaller than mb_format_merge_bound");
  char* p = out;
  for (uint64_t r = 0; r < n_rows; ++r) {
    const uint64_t g = site_index[r];
    p = put_u64(p, id_base + g + 1);
    for (uint32_t s = 0; s < n_samples; ++s) {
      *p++ = '\t';
      const double v = maf[(size_t)s * stride + g];
      if (v == 0.0 && !std::signbit(v)) *p++ = '0';
      else if (v == 1.0) *p++ = '1';
      else p += snprintf(p, 17, "%.3g", v);
    }
    *p++ = '\n';
  }
  return (int64_t)(p - out);
}

extern "C" int64_t mb_format_merge_depth_rows(const uint64_t* site_index, uint64_t n_rows, uint32_t n_samples,
                                              uint64_t stride, const uint32_t* depth, uint64_t id_base, char* out,
                                              uint64_t cap) {
  if (!site_index || !depth || !out) return mb::fail(MB_EINVAL, "mb_format_merge_depth_rows: null argument");
  if (cap < mb_format_merge_bound(n_rows, n_samples))
    return mb::fail(MB_EINVAL, "mb_format_merge_depth_rows: output buffer smaller than mb_format_merge_bound");
  char* p = out;
  for (uint64_t r = 0; r < n_rows; ++r) {
    const uint64_t g = site_index[r];
    p = put_u64(p, id_base + g + 1);
    for (uint32_t s = 0; s < n_samples; ++s) {
      *p++ = '\t';
      p = put_u32(p, depth[(size_t)s * stride + g]);
    }
    *p++ = '\n';
  }
  return (int64_t)(p - out);
}
