// Host side of the boundary: reference layout and the re-layout of decoded BAM
// records (mb_plain_reads) into the packed struct-of-arrays the kernels stream.
//
// This is the step pysam/htslib performs for the reference before its loops see an
// AlignedSegment (reference midas/run/snps.py:186,194; midas/run/genes.py:175-180):
// it derives, per read, the two CIGAR-only quantities keep_read needs
//   alen  = len(aln.query_alignment_sequence)   (snps.py:145, genes.py:154)
//   qlen  = aln.query_length                    (snps.py:146)
// and validates the inputs on which the reference would raise. No counting happens here.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

#include "common.h"

namespace mb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

}  // namespace mb

extern "C" const char* mb_version(void) { return "midas_b200 0.1.0 (sm_100a)"; }
extern "C" const char* mb_last_error(void) { return mb::g_err; }

// ------------------------------------------------------------------------------ layout
static std::atomic<uint64_t> g_layout_id{1};

extern "C" int mb_layout_create(const mb_ref* ref, mb_layout** out) {
  if (!ref || !out) return mb::fail(MB_EINVAL, "mb_layout_create: null argument");
  if (ref->n_species == 0 && ref->n_contigs != 0)
    return mb::fail(MB_EINVAL, "mb_layout_create: contigs without species");
  mb_layout* L = new mb_layout();
  L->id = g_layout_id.fetch_add(1);
  L->genes_only = false;
  L->n_contigs = ref->n_contigs;
  L->n_species = ref->n_species;
  L->contig_len.assign(ref->contig_len, ref->contig_len + ref->n_contigs);
  L->contig_species.assign(ref->contig_species, ref->contig_species + ref->n_contigs);
  L->contig_goff.assign(ref->n_contigs, 0);
  L->species_goff.assign(ref->n_species + 1, 0);
  L->species_len.assign(ref->n_species, 0);
  for (uint32_t c = 0; c < ref->n_contigs; ++c) {
    int32_t s = ref->contig_species[c];
    if (s < 0 || (uint32_t)s >= ref->n_species) {
      delete L;
      return mb::fail(MB_EINVAL, "mb_layout_create: contig %u has species %d outside [0,%u)", c, s,
                      ref->n_species);
    }
    if (ref->contig_len[c] == 0) {  // pysam count_coverage raises ValueError("interval of size 0")
      delete L;
      return mb::fail(MB_EINVAL, "mb_layout_create: contig %u has length 0", c);
    }
    L->species_len[s] += ref->contig_len[c];
  }
  // species-major, each species region starts on a tile boundary
  uint64_t cur = 0;
  std::vector<uint64_t> next(ref->n_species);
  for (uint32_t s = 0; s < ref->n_species; ++s) {
    cur = (cur + MB_TILE - 1) / MB_TILE * MB_TILE;
    L->species_goff[s] = (uint32_t)cur;
    next[s] = cur;
    cur += L->species_len[s];
    if (cur > 0xFFFF0000ull) {
      delete L;
      return mb::fail(MB_EOVERFLOW, "mb_layout_create: padded genome exceeds 32-bit coordinates");
    }
  }
  cur = (cur + MB_TILE - 1) / MB_TILE * MB_TILE;
  L->species_goff[ref->n_species] = (uint32_t)cur;
  L->gpad = cur;
  L->n_tiles = (uint32_t)(cur / MB_TILE);
  for (uint32_t c = 0; c < ref->n_contigs; ++c) {
    int32_t s = ref->contig_species[c];
    L->contig_goff[c] = (uint32_t)next[s];
    next[s] += ref->contig_len[c];
  }
  // contigs in global-coordinate order (species-major, given order inside a species)
  L->order.resize(ref->n_contigs);
  std::iota(L->order.begin(), L->order.end(), 0u);
  std::stable_sort(L->order.begin(), L->order.end(),
                   [&](uint32_t a, uint32_t b) { return L->contig_species[a] < L->contig_species[b]; });
  L->identity_order = true;
  for (uint32_t c = 0; c < ref->n_contigs; ++c) L->identity_order = L->identity_order && L->order[c] == c;
  L->tile_species.assign(L->n_tiles, -1);
  for (uint32_t s = 0; s < ref->n_species; ++s) {
    uint32_t t0 = L->species_goff[s] / MB_TILE;
    uint32_t t1 = (uint32_t)((L->species_goff[s] + L->species_len[s] + MB_TILE - 1) / MB_TILE);
    for (uint32_t t = t0; t < t1; ++t) L->tile_species[t] = (int32_t)s;
  }
  *out = L;
  return MB_OK;
}

extern "C" int mb_layout_create_genes(uint32_t n_genes, mb_layout** out) {
  if (!out) return mb::fail(MB_EINVAL, "mb_layout_create_genes: null argument");
  mb_layout* L = new mb_layout();
  L->id = g_layout_id.fetch_add(1);
  L->genes_only = true;
  L->n_contigs = n_genes;
  L->n_species = 0;
  L->gpad = 0;
  L->n_tiles = 0;
  L->identity_order = true;
  L->species_goff.assign(1, 0);
  *out = L;
  return MB_OK;
}

extern "C" void mb_layout_destroy(mb_layout* L) { delete L; }
extern "C" uint64_t mb_layout_gpad(const mb_layout* L) { return L ? L->gpad : 0; }
extern "C" uint32_t mb_layout_n_tiles(const mb_layout* L) { return L ? L->n_tiles : 0; }
extern "C" const uint32_t* mb_layout_contig_goff(const mb_layout* L) { return L->contig_goff.data(); }
extern "C" const uint32_t* mb_layout_species_goff(const mb_layout* L) { return L->species_goff.data(); }
extern "C" const uint64_t* mb_layout_species_len(const mb_layout* L) { return L->species_len.data(); }

// -------------------------------------------------------------------------------- pack
namespace {

struct ReadGeom {
  uint32_t loc;
  uint16_t alen, q_start, ref_span, nm;
  uint8_t flags;
  uint8_t max_q;          // largest Phred of the read
  uint32_t n_blk;         // MB_BLOCK-base blocks
  uint32_t n_cplx_words;  // 0 for simple reads, else 1 + n_ops_after_clip
  uint8_t need_words;     // snps: 32-position words a staged row of this read needs (saturated at 255)
};

// Analyse one record. Returns MB_OK or an error (message set).
int analyse(const mb_plain_reads* in, const mb_layout* L, int mode, uint64_t i, ReadGeom* g) {
  const int64_t lq = in->l_qseq[i];
  if (lq <= 0) return mb::fail(MB_EINVAL, "read %llu: l_qseq=%lld (reference: ZeroDivisionError snps.py:157)",
                               (unsigned long long)i, (long long)lq);
  if (lq > 65535) return mb::fail(MB_EOVERFLOW, "read %llu: l_qseq=%lld exceeds the 65535-base short-read limit",
                                  (unsigned long long)i, (long long)lq);
  if ((int64_t)(in->seq_off[i + 1] - in->seq_off[i]) < lq)
    return mb::fail(MB_EMALFORMED, "read %llu: seq_off span < l_qseq", (unsigned long long)i);
  if (in->seq_nibbles && (in->seq_off[i] & 1))
    return mb::fail(MB_EMALFORMED, "read %llu: seq_off must be even with seq_nibbles", (unsigned long long)i);
  if (in->nm[i] < 0)
    return mb::fail(MB_EINVAL, "read %llu: NM tag missing (reference: KeyError snps.py:148)", (unsigned long long)i);
  if (in->qual[in->seq_off[i]] == 0xFF)
    return mb::fail(MB_EINVAL, "read %llu: qualities missing (reference: np.mean(None) snps.py:151)",
                    (unsigned long long)i);
  const uint32_t* cig = in->cigar + in->cigar_off[i];
  const int64_t n_cig = (int64_t)(in->cigar_off[i + 1] - in->cigar_off[i]);
  if (n_cig > 65535) return mb::fail(MB_EOVERFLOW, "read %llu: more than 65535 CIGAR ops", (unsigned long long)i);
  int64_t qsum = 0, span = 0, n_match_ops = 0, n_other = 0;
  for (int64_t k = 0; k < n_cig; ++k) {
    uint32_t op = cig[k] & 0xF;
    int64_t len = cig[k] >> 4;
    if (op == mb::OP_P || op > mb::OP_X)
      return mb::fail(MB_EMALFORMED, "read %llu: unsupported CIGAR op %u", (unsigned long long)i, op);
    if (op == mb::OP_M || op == mb::OP_EQ || op == mb::OP_X) { qsum += len; span += len; ++n_match_ops; }
    else if (op == mb::OP_I) { qsum += len; ++n_other; }
    else if (op == mb::OP_S) { qsum += len; }
    else if (op == mb::OP_D || op == mb::OP_N) { span += len; ++n_other; }
  }
  if (n_cig > 0 && qsum != lq)
    return mb::fail(MB_EMALFORMED, "read %llu: CIGAR consumes %lld query bases, l_qseq=%lld", (unsigned long long)i,
                    (long long)qsum, (long long)lq);
  // pysam getQueryStart / getQueryEnd (the end walk never visits op 0)
  int64_t lead = 0, trail = 0;
  for (int64_t k = 0; k < n_cig; ++k) {
    uint32_t op = cig[k] & 0xF;
    if (op == mb::OP_H) continue;
    if (op == mb::OP_S) lead += cig[k] >> 4; else break;
  }
  for (int64_t k = n_cig - 1; k >= 1; --k) {
    uint32_t op = cig[k] & 0xF;
    if (op == mb::OP_H) continue;
    if (op == mb::OP_S) trail += cig[k] >> 4; else break;
  }
  int64_t alen = lq - lead - trail;
  if (alen <= 0) return mb::fail(MB_EINVAL, "read %llu: aligned length 0 (reference: ZeroDivisionError snps.py:148)",
                                 (unsigned long long)i);
  g->alen = (uint16_t)alen;
  g->flags = 0;
  uint32_t nm = (uint32_t)in->nm[i];
  if (nm > 65535u) { nm = 65535u; g->flags |= MB_RF_NMSAT; }
  g->nm = (uint16_t)nm;
  g->n_blk = (uint32_t)((lq + MB_BLOCK - 1) / MB_BLOCK);
  g->n_cplx_words = 0;
  {
    // or of all qualities, eight at a time: bit 7 set = a quality above 127, bits 6-7 clear = six planes suffice
    const uint8_t* q = in->qual + in->seq_off[i];
    uint64_t acc = 0;
    int64_t k = 0;
    for (; k + 8 <= lq; k += 8) { uint64_t w; memcpy(&w, q + k, 8); acc |= w; }
    for (; k < lq; ++k) acc |= q[k];
    acc |= acc >> 32; acc |= acc >> 16; acc |= acc >> 8;
    if (acc & 0x80u)
      return mb::fail(MB_EINVAL, "read %llu: a base quality above 127 is not representable", (unsigned long long)i);
    g->max_q = (uint8_t)(acc & 0xffu);                      // an upper bound with the same leading bit as the maximum
  }
  const int32_t rid = in->ref_id[i];
  if (mode == 1) {  // genes: loc = gene index, the pileup geometry is unused
    if (rid < 0 || (uint32_t)rid >= L->n_contigs)
      return mb::fail(MB_EINVAL, "read %llu: ref_id %d outside [0,%u) (reference: KeyError genes.py:180)",
                      (unsigned long long)i, rid, L->n_contigs);
    g->loc = (uint32_t)rid;
    g->q_start = (uint16_t)lead;
    g->ref_span = 0;
    g->need_words = 0;
    g->flags |= MB_RF_SIMPLE;
    return MB_OK;
  }
  if (rid < 0 || (uint32_t)rid >= L->n_contigs)
    return mb::fail(MB_EINVAL, "read %llu: ref_id %d outside [0,%u)", (unsigned long long)i, rid, L->n_contigs);
  const int64_t pos = in->pos[i];
  const int64_t clen = L->contig_len[rid];
  if (pos < 0 || pos >= clen)
    return mb::fail(MB_EMALFORMED, "read %llu: pos %lld outside contig of length %lld", (unsigned long long)i,
                    (long long)pos, (long long)clen);
  g->loc = L->contig_goff[rid] + (uint32_t)pos;
  const bool simple = (n_other == 0 && n_match_ops <= 1);
  int64_t span_clipped = std::min<int64_t>(span, clen - pos);  // pysam ignores refpos >= contig_len
  if (span_clipped > 65535)
    return mb::fail(MB_EOVERFLOW, "read %llu: reference span %lld exceeds 65535", (unsigned long long)i,
                    (long long)span_clipped);
  g->ref_span = (uint16_t)span_clipped;
  {  // row words: the aligned span placed at bit (loc & 31), and one lane per 32 query bases
    const int64_t by_span = (((int64_t)(g->loc & 31u)) + span_clipped + 31) / 32;
    const int64_t by_query = (lq + 31) / 32;
    g->need_words = (uint8_t)std::min<int64_t>(255, std::max(by_span, by_query));
  }
  if (simple) {
    // query index of the first aligned base = all S before the single M-like op
    int64_t qs = 0;
    for (int64_t k = 0; k < n_cig; ++k) {
      uint32_t op = cig[k] & 0xF;
      if (op == mb::OP_S) qs += cig[k] >> 4;
      else if (op != mb::OP_H) break;
    }
    g->q_start = (uint16_t)qs;
    g->flags |= MB_RF_SIMPLE;
  } else {
    g->q_start = (uint16_t)lead;
    g->n_cplx_words = (uint32_t)n_cig;  // the raw CIGAR words: the kernels walk them
  }
  return MB_OK;
}

// The M/=/X runs of a read after clipping at the contig end: dst = {n_runs, then for the first two runs
// (q0 | len << 16, x0)} = query index of the run's first base, its length, and its reference offset from
// the read's start.
void count_runs(const uint32_t* cig, int64_t n_cig, int64_t room /* contig_len - pos */, uint32_t* dst) {
  uint32_t n = 0;
  int64_t q = 0, r = 0;
  for (int64_t k = 0; k < n_cig && r < room; ++k) {
    uint32_t op = cig[k] & 0xF;
    int64_t len = cig[k] >> 4;
    if (op == mb::OP_M || op == mb::OP_EQ || op == mb::OP_X) {
      int64_t take = std::min<int64_t>(len, room - r);
      if (take > 0) {
        if (n < 2) {
          dst[1 + 2 * n] = (uint32_t)q | ((uint32_t)take << 16);
          dst[2 + 2 * n] = (uint32_t)r;
        }
        ++n;
      }
      q += len;
      r += len;
    } else if (op == mb::OP_I || op == mb::OP_S) {
      q += len;
    } else if (op == mb::OP_D || op == mb::OP_N) {
      r += len;
    }
  }
  dst[0] = n;
}

constexpr uint64_t kLsb = 0x0101010101010101ull;
// bit 0 of each of eight bytes -> the eight low bits of the result (byte i -> bit i)
static inline uint32_t gather8(uint64_t m) { return (uint32_t)((m * 0x0102040810204080ull) >> 56); }
// 1 in bit 0 of every byte of v (bytes < 0x80) that is not zero
static inline uint64_t nonzero_bytes(uint64_t v) { return (((v + 0x7f7f7f7f7f7f7f7full) | v) >> 7) & kLsb; }


template <class F>
void parallel_for(uint64_t n, int n_threads, F f) {
  if (n_threads <= 1 || n < 4096) { f(0, n, 0); return; }
  std::vector<std::thread> th;
  uint64_t per = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    uint64_t lo = std::min<uint64_t>(n, per * t), hi = std::min<uint64_t>(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=] { f(lo, hi, t); });
  }
  for (auto& x : th) x.join();
}

void* alloc_host(size_t bytes, bool pinned) {
  if (bytes == 0) bytes = 16;
  void* p = nullptr;
  if (pinned) {
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) return p;
    cudaGetLastError();
    return nullptr;
  }
  if (posix_memalign(&p, 256, bytes) != 0) return nullptr;
  return p;
}

}  // namespace

extern "C" void mb_packed_free(mb_packed_reads* p) {
  if (!p) return;
  mb_packed_owner* o = reinterpret_cast<mb_packed_owner*>(p);
  for (void* b : o->buf) {
    if (!b) continue;
    if (o->pinned) cudaFreeHost(b); else free(b);
  }
  delete o;
}

extern "C" int mb_pack_reads(const mb_plain_reads* in, const mb_layout* L, int mode, int sort, int pinned,
                             int n_threads, mb_packed_reads** out) {
  if (!in || !L || !out) return mb::fail(MB_EINVAL, "mb_pack_reads: null argument");
  if (mode != 0 && mode != 1) return mb::fail(MB_EINVAL, "mb_pack_reads: mode must be 0 (snps) or 1 (genes)");
  if (mode == 0 && L->genes_only) return mb::fail(MB_EINVAL, "mb_pack_reads: snps mode needs a layout with coordinates");
  const uint64_t n = in->n_reads;
  if (n >= 0xFFFFFFFFull) return mb::fail(MB_EOVERFLOW, "mb_pack_reads: more than 2^32-2 reads in one batch");
  if (n_threads <= 0) n_threads = (int)std::min<unsigned>(32, std::max(1u, std::thread::hardware_concurrency()));

  const bool timing = getenv("MB_PACK_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_phase = now();
  auto phase = [&](const char* name) { if (timing) { const double t = now(); fprintf(stderr, "mb_pack_reads: %-10s %8.1f ms\n", name, t - t_phase); t_phase = t; } };
  std::vector<ReadGeom> geom(n);
  std::vector<int> rc(n_threads + 1, MB_OK);
  std::vector<std::string> msg(n_threads + 1);
  parallel_for(n, n_threads, [&](uint64_t lo, uint64_t hi, int t) {
    for (uint64_t i = lo; i < hi; ++i) {
      int r = analyse(in, L, mode, i, &geom[i]);
      if (r != MB_OK) { rc[t] = r; msg[t] = mb_last_error(); return; }
    }
  });
  for (int t = 0; t <= n_threads; ++t)
    if (rc[t] != MB_OK) return mb::fail(rc[t], "%s", msg[t].c_str());

  phase("analyse");
  // order
  std::vector<uint32_t> perm;
  bool is_sorted = true;
  if (mode == 0) {
    std::vector<int> unsorted(n_threads + 1, 0);
    parallel_for(n, n_threads, [&](uint64_t lo, uint64_t hi, int t) {
      for (uint64_t i = std::max<uint64_t>(lo, 1); i < hi; ++i)
        if (geom[i].loc < geom[i - 1].loc) { unsorted[t] = 1; return; }
    });
    for (int t = 0; t <= n_threads; ++t) is_sorted = is_sorted && !unsorted[t];
    if (!is_sorted) {
      if (!sort) return mb::fail(MB_EUNSORTED, "mb_pack_reads: snps reads are not sorted by (layout order, pos)");
      perm.resize(n);
      std::iota(perm.begin(), perm.end(), 0u);
      std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return geom[a].loc < geom[b].loc; });
    }
  }
  auto src = [&](uint64_t j) -> uint64_t { return perm.empty() ? j : perm[j]; };

  // offsets
  uint8_t max_q = 0;
  for (uint64_t i = 0; i < n; ++i) max_q = std::max(max_q, geom[i].max_q);
  const uint32_t q_planes = max_q <= 63 ? 6u : 7u;
  const uint32_t NP = mode == 0 ? 3u + q_planes : q_planes;     // words per block
  const uint32_t P0 = mode == 0 ? 3u : 0u;                      // first Phred plane
  std::vector<uint32_t> unit_off(n + 1), cg_off(n + 1);
  uint64_t nb = 0, ncw = 0, tot_bases = 0;
  uint32_t max_span = 0, max_units = 0;
  {  // groups of MB_GROUP reads, units interleaved: units per read of a group = the largest of its reads
    const uint64_t n_groups = (n + MB_GROUP - 1) / MB_GROUP;
    std::vector<uint32_t> g_units(n_groups);
    parallel_for(n_groups, n_threads, [&](uint64_t lo, uint64_t hi, int) {
      for (uint64_t gi = lo; gi < hi; ++gi) {
        const uint64_t j1 = std::min<uint64_t>(n, (gi + 1) * MB_GROUP);
        uint64_t units = 0;
        for (uint64_t j = gi * MB_GROUP; j < j1; ++j)
          units = std::max<uint64_t>(units, ((uint64_t)geom[src(j)].n_blk * NP + 7) / 8);
        g_units[gi] = (uint32_t)units;
      }
    });
    std::vector<uint64_t> g_off(n_groups + 1, 0);
    for (uint64_t gi = 0; gi < n_groups; ++gi) g_off[gi + 1] = g_off[gi] + (uint64_t)g_units[gi] * MB_GROUP;
    for (uint64_t gi = 0; gi < n_groups; ++gi) max_units = std::max(max_units, g_units[gi]);
    nb = g_off[n_groups];
    if (nb > 0xFFFFFFF0ull) return mb::fail(MB_EOVERFLOW, "mb_pack_reads: batch too large for 32-bit unit offsets");
    parallel_for(n, n_threads, [&](uint64_t lo, uint64_t hi, int) {
      for (uint64_t j = lo; j < hi; ++j) unit_off[j] = (uint32_t)(g_off[j / MB_GROUP] + j % MB_GROUP);
    });
  }
  for (uint64_t j = 0; j < n; ++j) {
    const ReadGeom& g = geom[src(j)];
    cg_off[j] = (uint32_t)ncw;
    ncw += g.n_cplx_words;
    tot_bases += (uint64_t)in->l_qseq[src(j)];
    max_span = std::max<uint32_t>(max_span, g.ref_span);
    if (ncw > 0xFFFFFFF0ull)
      return mb::fail(MB_EOVERFLOW, "mb_pack_reads: batch too large for 32-bit run-record offsets");
  }
  unit_off[n] = (uint32_t)nb;
  cg_off[n] = (uint32_t)ncw;

  // row width of the tile kernel: the smallest W that all but 0.1 % of the reads fit
  uint32_t row_words = 0;
  uint64_t n_long = 0;
  if (mode == 0) {
    uint64_t hist[256] = {0};
    {
      std::vector<uint64_t> part((size_t)(n_threads + 1) * 256, 0);
      parallel_for(n, n_threads, [&](uint64_t lo, uint64_t hi, int t) {
        uint64_t* h = part.data() + (size_t)t * 256;
        for (uint64_t i = lo; i < hi; ++i) h[geom[i].need_words]++;
      });
      for (int t = 0; t <= n_threads; ++t)
        for (int w = 0; w < 256; ++w) hist[w] += part[(size_t)t * 256 + w];
    }
    uint64_t above = 0;
    for (int w = 255; w > (int)MB_ROW_WORDS_MAX; --w) above += hist[w];
    row_words = MB_ROW_WORDS_MAX;
    for (uint32_t w = MB_ROW_WORDS_MAX; w > 2; --w) {  // try shrinking to w-1
      if (above + hist[w] > n / 1000) break;
      above += hist[w];
      row_words = w - 1;
    }
    for (int w = (int)row_words + 1; w < 256; ++w) n_long += hist[w];
    if (n_long)
      parallel_for(n, n_threads, [&](uint64_t lo, uint64_t hi, int) {
        for (uint64_t i = lo; i < hi; ++i)
          if (geom[i].need_words > row_words) geom[i].flags |= MB_RF_LONG;
      });
  }

  phase("offsets");
  mb_packed_owner* o = new mb_packed_owner();
  memset(o, 0, sizeof(*o));
  o->pinned = pinned != 0;
  size_t sizes[7] = {sizeof(mb_read_hdr) * (size_t)n, sizeof(uint32_t) * (size_t)(n + 1), sizeof(uint32_t) * (size_t)n,
                     sizeof(uint32_t) * (size_t)ncw, (size_t)nb * 32, 0,
                     sizeof(uint32_t) * (size_t)n_long};
  for (int k = 0; k < 7; ++k) {
    o->buf[k] = alloc_host(sizes[k], o->pinned);
    if (!o->buf[k] && o->pinned) {  // no device / pin failure: fall back to pageable memory
      for (int q = 0; q < k; ++q) { cudaFreeHost(o->buf[q]); o->buf[q] = nullptr; }
      o->pinned = false;
      k = -1;
      continue;
    }
    if (!o->buf[k]) { mb_packed_free(&o->pub); return mb::fail(MB_ENOMEM, "mb_pack_reads: out of host memory"); }
  }
  mb_read_hdr* hdr = (mb_read_hdr*)o->buf[0];
  uint32_t* voff = (uint32_t*)o->buf[1];
  uint32_t* cplx = (uint32_t*)o->buf[2];
  uint32_t* cigar = (uint32_t*)o->buf[3];
  uint32_t* planes = (uint32_t*)o->buf[4];
  uint32_t* long_idx = (uint32_t*)o->buf[6];
  phase("alloc");
  memcpy(voff, unit_off.data(), sizeof(uint32_t) * (n + 1));
  parallel_for(nb, n_threads, [&](uint64_t lo, uint64_t hi, int) { memset(planes + lo * 8, 0, (hi - lo) * 32); });
  if (n_long) {
    uint64_t k = 0;
    for (uint64_t j = 0; j < n; ++j)
      if (geom[src(j)].flags & MB_RF_LONG) long_idx[k++] = (uint32_t)j;
  }

  parallel_for(n, n_threads, [&](uint64_t lo, uint64_t hi, int) {
    for (uint64_t j = lo; j < hi; ++j) {
      const uint64_t i = src(j);
      const ReadGeom& g = geom[i];
      mb_read_hdr h;
      h.loc = g.loc;
      h.l_qseq = (uint16_t)in->l_qseq[i];
      h.alen = g.alen;
      h.q_start = g.q_start;
      h.nm = g.nm;
      h.ref_span = g.ref_span;
      h.mapq = in->mapq[i];
      h.flags = g.flags;
      hdr[j] = h;
      if (g.n_cplx_words) {
        const uint32_t* cig = in->cigar + in->cigar_off[i];
        const int64_t n_cig = (int64_t)(in->cigar_off[i + 1] - in->cigar_off[i]);
        const int64_t room = (int64_t)L->contig_len[in->ref_id[i]] - in->pos[i];
        memcpy(cigar + cg_off[j], cig, sizeof(uint32_t) * (size_t)n_cig);   // raw ops, walked by the kernels
        cplx[j] = cg_off[j];
        // the common indel read = two runs (after clipping at the contig end): carried in the cplx word itself
        uint32_t rec[5] = {0, 0, 0, 0, 0};
        count_runs(cig, n_cig, room, rec);
        if (rec[0] == 2) {
          const uint32_t q0 = rec[1] & 0xffffu, len0 = rec[1] >> 16, x0 = rec[2];
          const uint32_t q1 = rec[3] & 0xffffu, len1 = rec[3] >> 16, x1 = rec[4];
          const int64_t ins = (int64_t)q1 - (int64_t)(q0 + len0), del = (int64_t)x1 - (int64_t)len0;
          if (q0 == g.q_start && x0 == 0 && ins >= 0 && ins <= 255 && del >= 0 && del <= 255 &&
              x1 + len1 == g.ref_span) {
            cplx[j] = len0 | ((uint32_t)ins << 16) | ((uint32_t)del << 24);
            h.flags |= MB_RF_RUNS2;
            hdr[j].flags = h.flags;
          }
        }
      } else {
        cplx[j] = 0xFFFFFFFFu;
      }
      const uint8_t* s = in->seq + in->seq_off[i];
      const uint8_t* q = in->qual + in->seq_off[i];
      const uint32_t lq = h.l_qseq;
      // word w of this read: unit w/8 of the read, units of a group interleaved (stride MB_GROUP units)
      uint32_t* reg = planes + (size_t)unit_off[j] * 8;
      auto word = [&](uint32_t w) -> uint32_t& { return reg[(size_t)(w >> 3) * (8 * MB_GROUP) + (w & 7u)]; };
      for (uint32_t b = 0; b < g.n_blk; ++b) {
        uint32_t w[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};        // lo, hi, acgt, Phred bits 0..6
        const uint32_t nk = std::min<uint32_t>(MB_BLOCK, lq - b * MB_BLOCK);
        uint8_t sb[MB_BLOCK] = {0}, qb[MB_BLOCK] = {0};          // the block's bases, zero padded
        if (in->seq_nibbles) {                                   // two codes per byte, even element in the high nibble
          const uint8_t* s4 = in->seq + (in->seq_off[i] >> 1) + b * (MB_BLOCK / 2);
          for (uint32_t k = 0; k < nk; ++k) sb[k] = (k & 1) ? (s4[k >> 1] & 15) : (s4[k >> 1] >> 4);
        } else {
          memcpy(sb, s + b * MB_BLOCK, nk);
        }
        memcpy(qb, q + b * MB_BLOCK, nk);
        for (uint32_t part = 0; part < 4; ++part) {              // eight bases at a time, byte-parallel in a u64
          uint64_t sw, qw;
          memcpy(&sw, sb + 8 * part, 8);
          memcpy(&qw, qb + 8 * part, 8);
          const uint32_t sh = 8 * part;
          for (uint32_t k = 0; k < q_planes; ++k) w[3 + k] |= gather8((qw >> k) & kLsb) << sh;
          // BAM nibble -> 2-bit code: A=1 C=2 G=4 T=8 are the nibbles with exactly one bit set
          const uint64_t nib = sw & 0x0f0f0f0f0f0f0f0full;
          const uint64_t t = nib & ((nib | 0x8080808080808080ull) - kLsb) & 0x0f0f0f0f0f0f0f0full;   // nib & (nib - 1)
          const uint64_t acgt = nonzero_bytes(nib) & ~nonzero_bytes(t);
          const uint64_t b1 = (nib >> 1) & kLsb, b2 = (nib >> 2) & kLsb, b3 = (nib >> 3) & kLsb;
          w[0] |= gather8(acgt & (b1 | b3)) << sh;               // C, T
          w[1] |= gather8(acgt & (b2 | b3)) << sh;               // G, T
          w[2] |= gather8(acgt) << sh;
        }
        if (mode == 0) { word(b * NP) = w[0]; word(b * NP + 1) = w[1]; word(b * NP + 2) = w[2]; }
        for (uint32_t k = 0; k < q_planes; ++k) word(b * NP + P0 + k) = w[3 + k];
      }
      if (mode == 0) {
        // candidate plane: keep only the bases of M/=/X runs whose reference position is inside the contig
        // (S, I never count; pysam ignores refpos >= contig_len)
        const uint32_t* cig = in->cigar + in->cigar_off[i];
        const int64_t n_cig = (int64_t)(in->cigar_off[i + 1] - in->cigar_off[i]);
        const int64_t room = (int64_t)L->contig_len[in->ref_id[i]] - in->pos[i];
        uint32_t inrun[(65535 + 31) / 32 + 1];
        for (uint32_t b = 0; b < g.n_blk; ++b) inrun[b] = 0;
        int64_t qq = 0, rr = 0;
        for (int64_t k = 0; k < n_cig; ++k) {
          const uint32_t op = cig[k] & 0xF;
          const int64_t len = cig[k] >> 4;
          if (op == mb::OP_M || op == mb::OP_EQ || op == mb::OP_X) {
            const int64_t take = std::max<int64_t>(0, std::min<int64_t>(len, room - rr));
            for (int64_t x = qq; x < qq + take;) {               // bits [qq, qq + take), a word at a time
              const int64_t wi = x >> 5, lo_b = x & 31, n_b = std::min<int64_t>(32 - lo_b, qq + take - x);
              inrun[wi] |= (n_b == 32 ? 0xffffffffu : ((1u << n_b) - 1u)) << lo_b;
              x += n_b;
            }
            qq += len;
            rr += len;
          } else if (op == mb::OP_I || op == mb::OP_S) {
            qq += len;
          } else if (op == mb::OP_D || op == mb::OP_N) {
            rr += len;
          }
        }
        for (uint32_t b = 0; b < g.n_blk; ++b) word(b * NP + 2) &= inrun[b];
      }
    }
  });

  phase("fill");
  o->pub.n_reads = n;
  o->pub.n_vec = nb;
  o->pub.n_cigar_words = ncw;
  o->pub.total_bases = tot_bases;
  o->pub.max_ref_span = max_span;
  o->pub.sorted = (mode == 0) ? 1u : 0u;
  o->pub.row_words = row_words;
  o->pub.n_long = (uint32_t)n_long;
  o->pub.q_planes = q_planes;
  o->pub.n_planes = NP;
  o->pub.max_units = max_units;
  o->pub.long_idx = long_idx;
  o->pub.hdr = hdr;
  o->pub.vec_off = voff;
  o->pub.cplx = cplx;
  o->pub.cigar = cigar;
  o->pub.planes = planes;
  *out = &o->pub;
  return MB_OK;
}
