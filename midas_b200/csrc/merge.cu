// Pooled allele call across samples on sm_100a ("next" row f.3, merge_midas.py snps).
//
// Replaces, for every genomic site of a species at once, the per-site arithmetic of the
// reference's GenomicSite (midas/merge/snps.py): compute_pooled_counts :42-47, call_alleles
// :49-76, compute_per_sample_mafs :78-91, compute_prevalence :93-104 (called per site from
// build_sharded_tables :324-360). Input is the stack of per-sample count planes exactly as
// mb_snps_pileup writes them: counts[s][plane][site]. Pure streaming, HBM-bound:
// 16 B x n_samples read per site, 12 B x n_samples + 24 B written.
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
};

// One thread per site. STAGED = true: the site's n_samples x 4 counts are read from HBM once and kept
// in a thread-private column of shared memory for the second (per-sample) pass; otherwise the second
// pass re-reads them from HBM (any n_samples).
template <bool STAGED>
__global__ void __launch_bounds__(MERGE_THREADS) k_merge_sites(const MergeParams p) {
  extern __shared__ uint32_t s_cnt[];          // [n_samples][4][MERGE_THREADS]
  const uint64_t G = p.n_sites;
  const uint32_t S = p.n_samples;
  for (uint64_t g = (uint64_t)blockIdx.x * MERGE_THREADS + threadIdx.x; g < G; g += (uint64_t)gridDim.x * MERGE_THREADS) {
    // pooled counts (snps.py:42-47)
    unsigned long long pc[4] = {0, 0, 0, 0};
    uint32_t s = 0;
    for (; s + 4 <= S; s += 4) {
      uint32_t v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t* c = p.counts + (size_t)(s + u) * 4 * G + g;
        v[u][0] = c[0]; v[u][1] = c[G]; v[u][2] = c[2 * G]; v[u][3] = c[3 * G];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          pc[k] += v[u][k];
          if (STAGED) s_cnt[((s + u) * 4 + k) * MERGE_THREADS + threadIdx.x] = v[u][k];
        }
    }
    for (; s < S; ++s) {
      const uint32_t* c = p.counts + (size_t)s * 4 * G + g;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t x = c[(size_t)k * G];
        pc[k] += x;
        if (STAGED) s_cnt[(s * 4 + k) * MERGE_THREADS + threadIdx.x] = x;
      }
    }
    const unsigned long long depth = pc[0] + pc[1] + pc[2] + pc[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) p.pooled[(size_t)k * G + g] = (uint32_t)(pc[k] > 0xffffffffull ? 0xffffffffull : pc[k]);
    // call_alleles (snps.py:49-76): stable sort by frequency, descending (ties keep A<C<G<T)
    int ord[4] = {0, 1, 2, 3};
#pragma unroll
    for (int a = 1; a < 4; ++a)
#pragma unroll
      for (int b = a; b > 0; --b)
        if (pc[ord[b]] > pc[ord[b - 1]]) { const int t = ord[b]; ord[b] = ord[b - 1]; ord[b - 1] = t; }
    uint32_t call = 0;
    bool has_major = false, has_minor = false;
    int major = 0, minor = 0;
    if (depth > 0) {
      const double dd = (double)depth;
      double fr[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) fr[k] = __ddiv_rn((double)pc[ord[k]], dd);
      if (fr[0] > 0) { has_major = true; major = ord[0]; }
      if (fr[1] > 0) { has_minor = true; minor = ord[1]; }
      uint32_t snp = 0;                      // 0 none, 1 mono, 2 bi, 3 tri, 4 quad
      if (fr[3] >= p.snp_freq) snp = 4;
      else if (fr[2] >= p.snp_freq) snp = 3;
      else if (fr[1] >= p.snp_freq) snp = 2;
      else if (fr[0] >= p.snp_freq) snp = 1;
      call = (uint32_t)major | ((uint32_t)minor << 2) | (snp << 4) | (has_minor ? MB_CALL_HAS_MINOR : 0u);
    }
    p.call[g] = (uint8_t)call;
    // compute_per_sample_mafs (snps.py:78-91) + compute_prevalence (snps.py:93-104)
    uint32_t pass = 0;
    for (uint32_t t = 0; t < S; ++t) {
      uint32_t sd = 0;
      double maf = 0.0;
      if (has_major) {
        const uint32_t cm = STAGED ? s_cnt[(t * 4 + major) * MERGE_THREADS + threadIdx.x]
                                   : p.counts[(size_t)t * 4 * G + (size_t)major * G + g];
        if (!has_minor) sd = cm;
        else {
          const uint32_t cn = STAGED ? s_cnt[(t * 4 + minor) * MERGE_THREADS + threadIdx.x]
                                     : p.counts[(size_t)t * 4 * G + (size_t)minor * G + g];
          sd = cm + cn;
          maf = sd > 0 ? __ddiv_rn((double)cn, (double)sd) : 0.0;
        }
      }
      p.sample_depth[(size_t)t * G + g] = sd;
      p.sample_maf[(size_t)t * G + g] = maf;
      const bool fail = ((int64_t)sd < (int64_t)p.site_depth) || (__ddiv_rn((double)sd, p.mean_depth[t]) > p.site_ratio);
      pass += fail ? 0u : 1u;
    }
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
  const uint64_t want = (n_sites + MERGE_THREADS - 1) / MERGE_THREADS;
  const size_t smem = sizeof(uint32_t) * 4 * (size_t)n_samples * MERGE_THREADS;
  // staging the counts in shared memory halves the HBM reads but leaves 4-8 warps per SM: measured 1.3-3x
  // slower than re-reading (profiles/r01_merge_bench.json); kept for small sample counts only
  const bool staged = smem <= 16 * 1024;
  int occ = 1;
  if (staged) {
    MB_CUDA_TRY(cudaFuncSetAttribute(k_merge_sites<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_merge_sites<true>, MERGE_THREADS, smem));
  } else {
    MB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_merge_sites<false>, MERGE_THREADS, 0));
  }
  const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * (occ > 0 ? occ : 1));
  { int trc = mb::timing_begin(ctx); if (trc) return trc; }
  if (staged) k_merge_sites<true><<<grid, MERGE_THREADS, smem, ctx->stream>>>(p);
  else k_merge_sites<false><<<grid, MERGE_THREADS, 0, ctx->stream>>>(p);
  { int trc = mb::timing_end(ctx); if (trc) return trc; }
  ctx->launches++;
  MB_CUDA_TRY(cudaGetLastError());
  return MB_OK;
}
