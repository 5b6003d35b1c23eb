// Pooled allele call across samples on sm_100a ("next" row f.3, merge_midas.py snps).
//
// Replaces, for every genomic site of a species at once, the per-site arithmetic of the
// reference's GenomicSite (midas/merge/snps.py): compute_pooled_counts :42-47, call_alleles
// :49-76, compute_per_sample_mafs :78-91, compute_prevalence :93-104 (called per site from
// build_sharded_tables :324-360). Input is the stack of per-sample count planes exactly as
// mb_snps_pileup writes them: counts[s][plane][site]. Pure streaming, HBM-bound:
// 16 B x n_samples read per site, 12 B x n_samples + 25 B written.
#include <cstdlib>

#include "common.h"

namespace {

constexpr int MERGE_THREADS = 128;

struct MergeParams {
  const uint32_t* counts;        // [S][4][G]
  const double* mean_depth;      // [S]
  uint64_t n_sites;
  uint32_t n_samples;
  double snp_freq, site_ratio;
  int32_t site_depth;
  uint32_t* pooled;              // [4][G]
  uint8_t* call;                 // [G]
  uint32_t* count_samples;       // [G]
  uint32_t* sample_depth;        // [S][G]
  double* sample_maf;            // [S][G]
  const unsigned long long* sd_max;   // [S] smallest per-sample depth that fails the site_ratio test (k_merge_thresholds)
  uint32_t n_shared;             // samples KS .. KS + n_shared - 1 keep their packed counts in shared memory
};

// Per sample, the largest per-sample depth sd that still passes the ratio test of compute_prevalence
// (snps.py:99-100: fail if sd / mean_depth > site_ratio). The quotient is a correctly rounded double
// division, monotone in sd, so the threshold is found once per call with the same division and the site
// kernel compares integers: exact, one division per sample instead of one per sample and site.
__global__ void k_merge_thresholds(const double* __restrict__ mean_depth, uint32_t n_samples, double site_ratio,
                                   unsigned long long* __restrict__ sd_max) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_samples) return;
  const double md = mean_depth[t];
  unsigned long long lo = 0, hi = 0x100000000ull;          // first sd in [0, 2^32] that fails
  while (lo < hi) {
    const unsigned long long mid = lo + ((hi - lo) >> 1);
    if (__ddiv_rn((double)mid, md) > site_ratio) hi = mid; else lo = mid + 1;
  }
  sd_max[t] = lo;                                           // sd passes iff sd < sd_max[t]
}

// One thread per site, two passes over the site's n_samples x 4 counts: pooled counts and the allele
// call, then per-sample depth / MAF / prevalence from the major and minor planes. The counts of the first
// KS samples stay in registers between the passes, packed four 8-bit counts to a word (a site where any of
// them exceeds 255 falls back to re-reading, like the samples past KS): the second pass then costs no HBM
// traffic for them. Both passes are unrolled four samples deep (sixteen / eight loads in flight per thread)
// and the MAF division is taken only where a sample has minor-allele reads.
template <int KS, bool SHARED>
__global__ void __launch_bounds__(MERGE_THREADS) k_merge_sites(const MergeParams p) {
  extern __shared__ uint32_t s_keep[];                       // [n_shared][MERGE_THREADS]
  const uint64_t G = p.n_sites;
  const uint32_t S = p.n_samples;
  for (uint64_t g = (uint64_t)blockIdx.x * MERGE_THREADS + threadIdx.x; g < G; g += (uint64_t)gridDim.x * MERGE_THREADS) {
    // pooled counts (snps.py:42-47)
    unsigned long long pc[4] = {0, 0, 0, 0};
    uint32_t kept[KS];
    uint32_t wide = 0;                                   // or of the kept counts: > 255 means they do not fit
#pragma unroll
    for (int s0 = 0; s0 < KS; s0 += 4) {
      if ((uint32_t)s0 < S) {
        uint32_t v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t* c = p.counts + (size_t)(s0 + u) * 4 * G + g;
          const bool in = (uint32_t)(s0 + u) < S;
          v[u][0] = in ? __ldg(c) : 0u; v[u][1] = in ? __ldg(c + G) : 0u;
          v[u][2] = in ? __ldg(c + 2 * G) : 0u; v[u][3] = in ? __ldg(c + 3 * G) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int k = 0; k < 4; ++k) pc[k] += v[u][k];
          wide |= v[u][0] | v[u][1] | v[u][2] | v[u][3];
          kept[s0 + u] = v[u][0] | (v[u][1] << 8) | (v[u][2] << 16) | (v[u][3] << 24);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) kept[s0 + u] = 0;
      }
    }
    // samples past the registers: packed the same way into a thread-private column of shared memory
    // (s_keep[sample - KS][thread]: conflict-free), as many as the launch reserved; the rest is re-read
    const uint32_t s_end = SHARED ? KS + p.n_shared : KS;
    uint32_t s = KS;
    for (; s + 4 <= S; s += 4) {
      uint32_t v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t* c = p.counts + (size_t)(s + u) * 4 * G + g;
        v[u][0] = __ldg(c); v[u][1] = __ldg(c + G); v[u][2] = __ldg(c + 2 * G); v[u][3] = __ldg(c + 3 * G);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) pc[k] += v[u][k];
        if (SHARED && s + u < s_end) {
          wide |= v[u][0] | v[u][1] | v[u][2] | v[u][3];
          s_keep[(s + u - KS) * MERGE_THREADS + threadIdx.x] = v[u][0] | (v[u][1] << 8) | (v[u][2] << 16) | (v[u][3] << 24);
        }
      }
    }
    for (; s < S; ++s) {
      const uint32_t* c = p.counts + (size_t)s * 4 * G + g;
      uint32_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { v[k] = __ldg(c + (size_t)k * G); pc[k] += v[k]; }
      if (SHARED && s < s_end) {
        wide |= v[0] | v[1] | v[2] | v[3];
        s_keep[(s - KS) * MERGE_THREADS + threadIdx.x] = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
      }
    }
    const bool packed = wide < 256u;
    const unsigned long long depth = pc[0] + pc[1] + pc[2] + pc[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) p.pooled[(size_t)k * G + g] = (uint32_t)(pc[k] > 0xffffffffull ? 0xffffffffull : pc[k]);
    // call_alleles (snps.py:49-76): stable sort by frequency, descending (ties keep A<C<G<T): a sorting
    // network on keys count << 2 | (3 - allele), all in registers
    unsigned long long key[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) key[k] = (pc[k] << 2) | (unsigned long long)(3 - k);
#define MB_CSWAP(a, b) { const unsigned long long hi_ = key[a] > key[b] ? key[a] : key[b], lo_ = key[a] > key[b] ? key[b] : key[a]; key[a] = hi_; key[b] = lo_; }
    MB_CSWAP(0, 1) MB_CSWAP(2, 3) MB_CSWAP(0, 2) MB_CSWAP(1, 3) MB_CSWAP(1, 2)
#undef MB_CSWAP
    uint32_t call = 0;
    bool has_major = false, has_minor = false;
    uint32_t major = 0, minor = 0;
    if (depth > 0) {
      const double dd = (double)depth;
      double fr[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) fr[k] = __ddiv_rn((double)(key[k] >> 2), dd);
      if (fr[0] > 0) { has_major = true; major = 3u - (uint32_t)(key[0] & 3ull); }
      if (fr[1] > 0) { has_minor = true; minor = 3u - (uint32_t)(key[1] & 3ull); }
      uint32_t snp = 0;                      // 0 none, 1 mono, 2 bi, 3 tri, 4 quad
      if (fr[3] >= p.snp_freq) snp = 4;
      else if (fr[2] >= p.snp_freq) snp = 3;
      else if (fr[1] >= p.snp_freq) snp = 2;
      else if (fr[0] >= p.snp_freq) snp = 1;
      call = major | (minor << 2) | (snp << 4) | (has_minor ? MB_CALL_HAS_MINOR : 0u);
    }
    p.call[g] = (uint8_t)call;
    // compute_per_sample_mafs (snps.py:78-91) + compute_prevalence (snps.py:93-104)
    uint32_t pass = 0;
    const uint32_t* cmaj = p.counts + (size_t)major * G + g;
    const uint32_t* cmin = p.counts + (size_t)minor * G + g;
    auto finish = [&](const uint32_t t, const uint32_t cm, const uint32_t cn) {
      uint32_t sd = 0;
      double maf = 0.0;
      if (has_major) {
        sd = has_minor ? cm + cn : cm;
        if (has_minor && cn > 0) maf = __ddiv_rn((double)cn, (double)sd);   // cn == 0: 0/sd = 0.0 exactly
      }
      p.sample_depth[(size_t)t * G + g] = sd;
      p.sample_maf[(size_t)t * G + g] = maf;
      const bool fail = ((int64_t)sd < (int64_t)p.site_depth) || ((unsigned long long)sd >= __ldg(p.sd_max + t));
      pass += fail ? 0u : 1u;
    };
    const uint32_t shm = 8u * major, shn = 8u * minor;
#pragma unroll
    for (int t0 = 0; t0 < KS; t0 += 4) {
      if ((uint32_t)t0 < S) {
        uint32_t cm[4], cn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (packed) {
            cm[u] = has_major ? (kept[t0 + u] >> shm) & 255u : 0u;
            cn[u] = has_minor ? (kept[t0 + u] >> shn) & 255u : 0u;
          } else {
            const bool in = (uint32_t)(t0 + u) < S;
            cm[u] = has_major && in ? __ldg(cmaj + (size_t)(t0 + u) * 4 * G) : 0u;
            cn[u] = has_minor && in ? __ldg(cmin + (size_t)(t0 + u) * 4 * G) : 0u;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if ((uint32_t)(t0 + u) < S) finish(t0 + u, cm[u], cn[u]);
      }
    }
    uint32_t t = KS;
    if (SHARED && packed) {                                    // shared-memory samples: no HBM traffic
      const uint32_t t_end = s_end < S ? s_end : S;
      for (; t < t_end; ++t) {
        const uint32_t w = s_keep[(t - KS) * MERGE_THREADS + threadIdx.x];
        finish(t, has_major ? (w >> shm) & 255u : 0u, has_minor ? (w >> shn) & 255u : 0u);
      }
    }
    for (; t + 4 <= S; t += 4) {
      uint32_t cm[4], cn[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cm[u] = has_major ? __ldg(cmaj + (size_t)(t + u) * 4 * G) : 0u;
        cn[u] = has_minor ? __ldg(cmin + (size_t)(t + u) * 4 * G) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) finish(t + u, cm[u], cn[u]);
    }
    for (; t < S; ++t)
      finish(t, has_major ? __ldg(cmaj + (size_t)t * 4 * G) : 0u, has_minor ? __ldg(cmin + (size_t)t * 4 * G) : 0u);
    p.count_samples[g] = pass;
  }
}

}  // namespace

extern "C" int mb_merge_sites_dev(mb_ctx* ctx, const uint32_t* counts_dev, uint32_t n_samples, uint64_t n_sites,
                                  const double* mean_depth_dev, const mb_merge_params* mp, uint32_t* pooled_dev,
                                  uint8_t* call_dev, uint32_t* count_samples_dev, uint32_t* sample_depth_dev,
                                  double* sample_maf_dev) {
  if (!ctx || !counts_dev || !mean_depth_dev || !mp || !pooled_dev || !call_dev || !count_samples_dev ||
      !sample_depth_dev || !sample_maf_dev)
    return mb::fail(MB_EINVAL, "mb_merge_sites_dev: null argument");
  if (n_samples == 0) return mb::fail(MB_EINVAL, "mb_merge_sites_dev: no samples (reference: ZeroDivisionError snps.py:104)");
  MB_CUDA_TRY(cudaSetDevice(ctx->device));
  if (n_sites == 0) return MB_OK;
  MergeParams p;
  p.counts = counts_dev;
  p.mean_depth = mean_depth_dev;
  p.n_sites = n_sites;
  p.n_samples = n_samples;
  p.snp_freq = mp->snp_freq;
  p.site_ratio = mp->site_ratio;
  p.site_depth = mp->site_depth;
  p.pooled = pooled_dev;
  p.call = call_dev;
  p.count_samples = count_samples_dev;
  p.sample_depth = sample_depth_dev;
  p.sample_maf = sample_maf_dev;
  void* thr = nullptr;
  { int rc = mb::ctx_reserve(ctx, 14, sizeof(unsigned long long) * (size_t)n_samples, &thr); if (rc) return rc; }
  p.sd_max = (const unsigned long long*)thr;
  k_merge_thresholds<<<(n_samples + 127) / 128, 128, 0, ctx->stream>>>(mean_depth_dev, n_samples, mp->site_ratio,
                                                                       (unsigned long long*)thr);
  ctx->launches++;
  const uint64_t want = (n_sites + MERGE_THREADS - 1) / MERGE_THREADS;
  // 32 samples stay in registers between the passes (80 registers; 64 samples = 128 registers measured slower
  // at every sample count). MB_MERGE_KS=4 selects the re-reading variant for comparison.
  int ks = 32;
  if (const char* e = getenv("MB_MERGE_KS")) { if (atoi(e) == 4) ks = 4; }
  void (*fn)(const MergeParams) = ks == 4 ? k_merge_sites<4, false>
                                          : (n_samples > 32 ? k_merge_sites<32, true> : k_merge_sites<32, false>);
  // samples past the 32 in registers keep their packed counts in shared memory: up to 96 of them (48 KB per
  // CTA, four CTAs per SM); MB_MERGE_SHARED overrides the count (experiments)
  uint32_t n_shared = (ks == 32 && n_samples > 32u) ? std::min<uint32_t>(n_samples - 32u, 96u) : 0u;
  if (const char* e = getenv("MB_MERGE_SHARED")) n_shared = std::min<uint32_t>(n_shared, (uint32_t)std::max(0, atoi(e)));
  p.n_shared = n_shared;
  const size_t smem = sizeof(uint32_t) * (size_t)n_shared * MERGE_THREADS;
  if (smem > 48 * 1024) MB_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  MB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, MERGE_THREADS, smem));
  const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * (occ > 0 ? occ : 1));
  { int trc = mb::timing_begin(ctx); if (trc) return trc; }
  fn<<<grid, MERGE_THREADS, smem, ctx->stream>>>(p);
  { int trc = mb::timing_end(ctx); if (trc) return trc; }
  ctx->launches++;
  MB_CUDA_TRY(cudaGetLastError());
  return MB_OK;
}
