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

__global__ void __launch_bounds__(256) k_merge_sites(const MergeParams p) {
  const uint64_t G = p.n_sites;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (uint64_t)gridDim.x * blockDim.x) {
    // pooled counts (snps.py:42-47)
    unsigned long long pc[4] = {0, 0, 0, 0};
    for (uint32_t s = 0; s < p.n_samples; ++s) {
      const uint32_t* c = p.counts + (size_t)s * 4 * G + g;
      pc[0] += c[0]; pc[1] += c[G]; pc[2] += c[2 * G]; pc[3] += c[3 * G];
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
    for (uint32_t s = 0; s < p.n_samples; ++s) {
      const uint32_t* c = p.counts + (size_t)s * 4 * G + g;
      uint32_t sd = 0;
      double maf = 0.0;
      if (has_major) {
        const uint32_t cm = c[(size_t)major * G];
        if (!has_minor) sd = cm;
        else {
          const uint32_t cn = c[(size_t)minor * G];
          sd = cm + cn;
          maf = sd > 0 ? __ddiv_rn((double)cn, (double)sd) : 0.0;
        }
      }
      p.sample_depth[(size_t)s * G + g] = sd;
      p.sample_maf[(size_t)s * G + g] = maf;
      const bool fail = ((int64_t)sd < (int64_t)p.site_depth) || (__ddiv_rn((double)sd, p.mean_depth[s]) > p.site_ratio);
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
  const uint64_t want = (n_sites + 255) / 256;
  const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * 8);
  { int trc = mb::timing_begin(ctx); if (trc) return trc; }
  k_merge_sites<<<grid, 256, 0, ctx->stream>>>(p);
  { int trc = mb::timing_end(ctx); if (trc) return trc; }
  ctx->launches++;
  MB_CUDA_TRY(cudaGetLastError());
  return MB_OK;
}
