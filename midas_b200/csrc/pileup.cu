// SNP pileup on sm_100a: per-reference-position A/C/G/T counts + per-species stats.
//
// Replaces, for all species of a sample at once (reference paths):
//   midas/run/snps.py:141-162  keep_read        -> header filters + mean-quality test, fused
//   midas/run/snps.py:194-199  count_coverage   -> CIGAR walk + per-base quality test + count
//   midas/run/snps.py:201-213  per-position loop -> depth / covered_bases / total_depth, fused
//                                                  into the tile flush (no second pass)
//
// Data-parallel decomposition: the padded genome is cut into MB_TILE-position tiles. Reads are
// sorted by global coordinate, so the reads overlapping a tile are one contiguous index range.
// A CTA owns a tile: it accumulates the tile's four count planes in shared memory (ATOMS),
// then writes them to HBM with plain coalesced 128-bit stores. No global atomics on counters,
// no pre-zeroing of the output, every output element written exactly once. Reads that straddle
// a tile boundary are visited by both tiles, each counting only its own positions; the tile
// containing the read's start "owns" it for aligned_reads/mapped_reads.
#include <cstdio>

#include "common.h"

namespace {

constexpr uint32_t TILE = MB_TILE;
constexpr int PILE_THREADS = 512;
constexpr int PILE_WARPS = PILE_THREADS / 32;
constexpr int CHUNK = 1024;  // reads whose header filters are evaluated per phase

struct PileParams {
  const uint4* hdr;
  const uint32_t* blk_off;
  const uint32_t* cplx;
  const uint32_t* cigar;
  const uint32_t* seq2;
  const uint8_t* qual;
  const uint32_t* tile_rng;
  const int32_t* tile_species;
  uint32_t* counts;
  unsigned long long* stats;
  uint32_t* sched;
  uint64_t gpad;
  uint32_t n_tiles;
  double mapid, aln_cov;
  int readq, mapq;
  uint32_t qthr;  // 0x80 + baseq: byte >= qthr <=> literal ACGT and Phred >= baseq
};

// Shared-memory index of tile-relative position x inside a plane. Within each block of 128
// positions the 4x32 matrix is transposed, so that lanes holding positions x0+4*lane+k (k fixed)
// hit 32 distinct banks.
__device__ __forceinline__ uint32_t swz(uint32_t x) { return (x & ~127u) | ((x & 3u) << 5) | ((x >> 2) & 31u); }

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void k_init_stats(unsigned long long* stats, const uint64_t* species_len, uint32_t n_species,
                             uint32_t* sched) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) sched[0] = 0;
  if (s < n_species) {
    stats[5 * s + 0] = species_len[s];  // genome_length (snps.py:211)
    stats[5 * s + 1] = 0;
    stats[5 * s + 2] = 0;
    stats[5 * s + 3] = 0;
    stats[5 * s + 4] = 0;
  }
}

// For every tile the contiguous range of reads that can overlap it. Conservative lower end
// (tile_start - max_ref_span); the pileup kernel skips non-overlapping reads by header.
__global__ void k_tile_ranges(const uint4* __restrict__ hdr, uint32_t n_reads, uint32_t max_span, uint32_t n_tiles,
                              uint32_t* __restrict__ rng) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const uint32_t ts = t * TILE, te = ts + TILE;
  const uint32_t lo_key = ts > max_span ? ts - max_span : 0u;
  uint32_t lo = 0, hi = n_reads;
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&hdr[mid].x) < lo_key) lo = mid + 1; else hi = mid;
  }
  const uint32_t r0 = lo;
  hi = n_reads;
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&hdr[mid].x) < te) lo = mid + 1; else hi = mid;
  }
  rng[2 * t] = r0;
  rng[2 * t + 1] = lo;
}

__global__ void __launch_bounds__(PILE_THREADS, 2) k_pileup_v1(const PileParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* hist = smem;                                         // [4][TILE]
  uint32_t* s_blk = hist + 4 * TILE;                             // [CHUNK]
  uint8_t* s_hk = reinterpret_cast<uint8_t*>(s_blk + CHUNK);     // [CHUNK]
  __shared__ unsigned long long s_red[4];                        // covered, depth, aligned, mapped
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_tile = atomicAdd(p.sched, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.n_tiles) break;
    const uint32_t ts = tile * TILE;
    const uint32_t r0 = p.tile_rng[2 * tile], r1 = p.tile_rng[2 * tile + 1];

    for (int i = tid; i < (int)TILE; i += PILE_THREADS) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0, 0, 0, 0);
    if (tid < 4) s_red[tid] = 0;
    uint32_t my_aligned = 0, my_mapped = 0;

    for (uint32_t c0 = r0; c0 < r1; c0 += CHUNK) {
      const int n = (int)min((uint32_t)CHUNK, r1 - c0);
      __syncthreads();
      // ---- phase 0: header-only filters, one thread per read (snps.py:148,154,157)
      for (int i = tid; i < n; i += PILE_THREADS) {
        const uint4 h = __ldg(&p.hdr[c0 + i]);
        const uint32_t loc = h.x, lq = h.y & 0xffffu, alen = h.y >> 16, nm = h.z >> 16;
        const uint32_t span = h.w & 0xffffu, mq = (h.w >> 16) & 0xffu, fl = h.w >> 24;
        uint32_t code;
        if (loc < ts && loc + span <= ts) {
          code = 2;  // neither owned by nor overlapping this tile
        } else {
          double pid = __ddiv_rn((double)(100 * ((int)alen - (int)nm)), (double)alen);
          if (fl & MB_RF_NMSAT) pid = -__longlong_as_double(0x7ff0000000000000ll);
          bool ok = !(pid < p.mapid);
          ok = ok && !((int)mq < p.mapq);
          ok = ok && !(__ddiv_rn((double)alen, (double)lq) < p.aln_cov);
          code = ok ? 1u : 0u;
        }
        s_hk[i] = (uint8_t)code;
        s_blk[i] = __ldg(&p.blk_off[c0 + i]);
      }
      __syncthreads();
      // ---- phase 1: one warp per read
      for (int i = warp; i < n; i += PILE_WARPS) {
        const uint32_t code = s_hk[i];
        if (code & 2u) continue;
        const uint4 h = __ldg(&p.hdr[c0 + i]);
        const uint32_t loc = h.x, lq = h.y & 0xffffu, qstart = h.z & 0xffffu;
        const uint32_t span = h.w & 0xffffu, fl = h.w >> 24;
        const uint32_t blk0 = s_blk[i];
        const uint32_t nwords = ((lq + 15u) >> 4) << 2;
        const uint32_t* __restrict__ qw = reinterpret_cast<const uint32_t*>(p.qual) + (size_t)blk0 * 4;
        const bool owner = loc >= ts;
        if (!(code & 1u)) {  // failed a header filter: the qualities are never needed
          if (lane == 0 && owner) ++my_aligned;
          continue;
        }
        // mean-quality filter (snps.py:151): mean(q) < readq  <=>  sum(q) < readq * l_qseq (exact)
        uint32_t w0 = (uint32_t)lane < nwords ? __ldg(qw + lane) : 0u;
        uint32_t w1 = (uint32_t)lane + 32u < nwords ? __ldg(qw + lane + 32) : 0u;
        uint32_t s = __vsadu4(w0 & 0x7f7f7f7fu, 0u) + __vsadu4(w1 & 0x7f7f7f7fu, 0u);
        for (uint32_t j = lane + 64; j < nwords; j += 32) s += __vsadu4(__ldg(qw + j) & 0x7f7f7f7fu, 0u);
        s = warp_sum(s);
        const bool keep = (long long)s >= (long long)p.readq * (long long)lq;
        if (lane == 0 && owner) { ++my_aligned; my_mapped += keep ? 1u : 0u; }
        if (!keep) continue;
        const int x0 = (int)(loc - ts);  // tile-relative position of the first aligned base (may be < 0)
        if (fl & MB_RF_SIMPLE) {
          for (uint32_t j = lane; j < nwords; j += 32) {
            const uint32_t w = j < 32 ? w0 : (j < 64 ? w1 : __ldg(qw + j));
            const uint32_t sw = __ldg(p.seq2 + blk0 + (j >> 2));
            const uint32_t sub = (j & 3u) << 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int rel = (int)(4 * j + k) - (int)qstart;
              const int x = x0 + rel;
              const uint32_t qb = (w >> (8 * k)) & 0xffu;
              if ((uint32_t)rel < span && (uint32_t)x < TILE && qb >= p.qthr) {
                const uint32_t bi = sub + k;
                const uint32_t cd = ((sw >> bi) & 1u) | (((sw >> (16 + bi)) & 1u) << 1);
                atomicAdd(&hist[cd * TILE + swz((uint32_t)x)], 1u);
              }
            }
          }
        } else {
          const uint32_t* __restrict__ rec = p.cigar + __ldg(p.cplx + c0 + i);
          const uint32_t n_ops = __ldg(rec);
          int qpos = 0, rpos = x0;
          for (uint32_t o = 0; o < n_ops; ++o) {
            const uint32_t cw = __ldg(rec + 1 + o);
            const uint32_t op = cw & 0xfu;
            const int len = (int)(cw >> 4);
            if (op == mb::OP_M || op == mb::OP_EQ || op == mb::OP_X) {
              for (int q = qpos + lane; q < qpos + len; q += 32) {
                const int x = rpos + (q - qpos);
                if ((uint32_t)x < TILE) {
                  const uint32_t qb = __ldg(p.qual + (size_t)blk0 * 16 + q);
                  if (qb >= p.qthr) {
                    const uint32_t sw = __ldg(p.seq2 + blk0 + (q >> 4));
                    const uint32_t bi = q & 15;
                    const uint32_t cd = ((sw >> bi) & 1u) | (((sw >> (16 + bi)) & 1u) << 1);
                    atomicAdd(&hist[cd * TILE + swz((uint32_t)x)], 1u);
                  }
                }
              }
              qpos += len;
              rpos += len;
            } else if (op == mb::OP_I || op == mb::OP_S) {
              qpos += len;
            } else if (op == mb::OP_D || op == mb::OP_N) {
              rpos += len;
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- flush: smem planes -> HBM (coalesced 128-bit stores) + fused per-position epilogue
    uint32_t covered = 0;
    unsigned long long depth_sum = 0;
    for (uint32_t x4 = tid * 4; x4 < TILE; x4 += PILE_THREADS * 4) {
      const uint32_t base = (x4 & ~127u) | ((x4 >> 2) & 31u);
      uint32_t d[4] = {0, 0, 0, 0};
#pragma unroll
      for (int pl = 0; pl < 4; ++pl) {
        uint4 v;
        v.x = hist[pl * TILE + base];
        v.y = hist[pl * TILE + base + 32];
        v.z = hist[pl * TILE + base + 64];
        v.w = hist[pl * TILE + base + 96];
        d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
        *reinterpret_cast<uint4*>(p.counts + (size_t)pl * p.gpad + ts + x4) = v;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { covered += d[k] > 0; depth_sum += d[k]; }  // snps.py:204,212-213
    }
    covered = warp_sum(covered);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) depth_sum += __shfl_xor_sync(0xffffffffu, depth_sum, o);
    if (lane == 0) {
      if (covered) atomicAdd(&s_red[0], (unsigned long long)covered);
      if (depth_sum) atomicAdd(&s_red[1], depth_sum);
      if (my_aligned) atomicAdd(&s_red[2], (unsigned long long)my_aligned);
      if (my_mapped) atomicAdd(&s_red[3], (unsigned long long)my_mapped);
    }
    __syncthreads();
    if (tid < 4) {
      const int sp = p.tile_species[tile];
      const unsigned long long v = s_red[tid];
      if (sp >= 0 && v) atomicAdd(&p.stats[5 * sp + 1 + tid], v);
    }
  }
}

constexpr size_t kPileSmem = sizeof(uint32_t) * 4 * TILE + sizeof(uint32_t) * CHUNK + CHUNK;

}  // namespace

extern "C" int mb_snps_pileup_dev(mb_ctx* ctx, const mb_dev_reads* reads, const mb_layout* L, const mb_filters* f,
                                  uint32_t* counts_dev, mb_species_stats* stats_dev) {
  if (!ctx || !reads || !L || !f || !counts_dev || !stats_dev)
    return mb::fail(MB_EINVAL, "mb_snps_pileup_dev: null argument");
  if (!reads->sorted) return mb::fail(MB_EUNSORTED, "mb_snps_pileup_dev: reads were not packed in snps (sorted) mode");
  if (f->baseq < 0 || f->baseq > 127) return mb::fail(MB_EINVAL, "mb_snps_pileup_dev: baseq %d outside [0,127]", f->baseq);
  MB_CUDA_TRY(cudaSetDevice(ctx->device));
  mb_dev_layout dl;
  int rc = mb::get_dev_layout(ctx, L, &dl);
  if (rc) return rc;
  const uint32_t n_tiles = L->n_tiles;
  if (ctx->tile_rng_cap < (size_t)2 * n_tiles + 2) {
    if (ctx->tile_rng) { MB_CUDA_TRY(cudaStreamSynchronize(ctx->stream)); MB_CUDA_TRY(cudaFree(ctx->tile_rng)); }
    ctx->tile_rng = nullptr;
    ctx->tile_rng_cap = 0;
    MB_CUDA_TRY(cudaMalloc(&ctx->tile_rng, sizeof(uint32_t) * ((size_t)2 * n_tiles + 2)));
    ctx->tile_rng_cap = (size_t)2 * n_tiles + 2;
  }
  if (!ctx->pile_attr_set) {
    MB_CUDA_TRY(cudaFuncSetAttribute(k_pileup_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPileSmem));
    ctx->pile_attr_set = true;
  }
  cudaStream_t st = ctx->stream;
  const uint32_t ns = L->n_species;
  k_init_stats<<<ns ? (ns + 255) / 256 : 1, 256, 0, st>>>(
      reinterpret_cast<unsigned long long*>(stats_dev), dl.species_len, ns, ctx->sched);
  ctx->launches++;
  if (n_tiles == 0) { MB_CUDA_TRY(cudaGetLastError()); return MB_OK; }
  k_tile_ranges<<<(n_tiles + 255) / 256, 256, 0, st>>>(reinterpret_cast<const uint4*>(reads->hdr),
                                                       (uint32_t)reads->n_reads, reads->max_ref_span, n_tiles,
                                                       ctx->tile_rng);
  ctx->launches++;
  PileParams p;
  p.hdr = reinterpret_cast<const uint4*>(reads->hdr);
  p.blk_off = reads->blk_off;
  p.cplx = reads->cplx;
  p.cigar = reads->cigar;
  p.seq2 = reads->seq2;
  p.qual = reads->qual;
  p.tile_rng = ctx->tile_rng;
  p.tile_species = dl.tile_species;
  p.counts = counts_dev;
  p.stats = reinterpret_cast<unsigned long long*>(stats_dev);
  p.sched = ctx->sched;
  p.gpad = L->gpad;
  p.n_tiles = n_tiles;
  p.mapid = f->mapid;
  p.aln_cov = f->aln_cov;
  p.readq = f->readq;
  p.mapq = f->mapq;
  p.qthr = 0x80u + (uint32_t)f->baseq;
  const uint32_t grid = min(n_tiles, (uint32_t)ctx->sm_count * 2u);
  { int trc = mb::timing_begin(ctx); if (trc) return trc; }
  k_pileup_v1<<<grid, PILE_THREADS, kPileSmem, st>>>(p);
  { int trc = mb::timing_end(ctx); if (trc) return trc; }
  ctx->launches++;
  MB_CUDA_TRY(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_snps_pileup(mb_ctx* ctx, const mb_packed_reads* reads, const mb_layout* L, const mb_filters* f,
                              uint32_t* counts_host, mb_species_stats* stats_host) {
  if (!ctx || !reads || !L || !f || !counts_host || !stats_host)
    return mb::fail(MB_EINVAL, "mb_snps_pileup: null argument");
  MB_CUDA_TRY(cudaSetDevice(ctx->device));
  mb_dev_reads dr;
  mb_dev_reads* d = &dr;
  int rc = mb::upload_reads(ctx, reads, d, /*staged=*/true);
  if (rc) return rc;
  void *counts = nullptr, *stats = nullptr;
  const size_t cbytes = sizeof(uint32_t) * 4 * (size_t)L->gpad, sbytes = sizeof(mb_species_stats) * (size_t)L->n_species;
  rc = mb::ctx_reserve(ctx, 0, cbytes, &counts);
  if (!rc) rc = mb::ctx_reserve(ctx, 1, sbytes, &stats);
  if (!rc) rc = mb_snps_pileup_dev(ctx, d, L, f, (uint32_t*)counts, (mb_species_stats*)stats);
  if (!rc && cbytes) {
    cudaError_t e = cudaMemcpyAsync(counts_host, counts, cbytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && sbytes) e = cudaMemcpyAsync(stats_host, stats, sbytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = mb::fail(MB_ECUDA, "mb_snps_pileup: %s", cudaGetErrorString(e));
  }
  return rc;
}
