// Gene coverage on sm_100a: per-gene aligned/mapped reads and mapped bp.
//
// Replaces the record loop of count_mapped_bp (reference midas/run/genes.py:178-189) and its
// keep_read (genes.py:153-169). Reads arrive in BAM file order (unsorted); each read touches one
// gene, so the kernel is a streaming scan (16 B header + the read's qualities) followed by two
// L2-resident atomics per read into the 24-byte per-gene record.
//
// Per warp, 32 consecutive reads at a time:
//   A. lane i owns read i: coalesced header + block-offset loads, header-only filters
//      (identity genes.py:157, mapq :163, alignment coverage :166), one thread per read;
//   B. the qualities of the reads that survived A are summed 16 lanes per read with 128-bit loads
//      (mean-quality test genes.py:160; exact integer form sum(q) >= readq*l_qseq);
//   C. lane i issues read i's atomics: (aligned,mapped) as one packed u64 add, mapped_bp as another.
#include <cub/device/device_radix_sort.cuh>

#include "common.h"

namespace {

constexpr int GENE_THREADS = 256;

struct GeneParams {
  const uint4* hdr;
  const uint32_t* blk_off;
  const uint8_t* qual;
  uint64_t n_reads;
  unsigned long long* out;  // mb_gene_out as 3 x u64: [aligned|mapped<<32], mapped_bp, depth bits
  uint32_t* key;            // exact order: per read, gene index if kept else 0xFFFFFFFF (may be null)
  uint32_t* val;            // exact order: per read, aligned length
  uint32_t n_genes;
  double mapid, aln_cov;
  int readq, mapq;
};

__global__ void __launch_bounds__(GENE_THREADS) k_genes_v1(const GeneParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp_global = ((uint64_t)blockIdx.x * GENE_THREADS + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * GENE_THREADS) >> 5;
  const int sub = lane >> 4, sl = lane & 15;  // two 16-lane groups
  for (uint64_t base = warp_global * 32; base < p.n_reads; base += n_warps * 32) {
    const uint64_t r = base + lane;
    const bool valid = r < p.n_reads;
    uint32_t gene = 0, lq = 0, alen = 0, blk0 = 0;
    bool ok = false;
    if (valid) {
      const uint4 h = __ldg(&p.hdr[r]);
      gene = h.x;
      lq = h.y & 0xffffu;
      alen = h.y >> 16;
      const uint32_t nm = h.z >> 16, mq = (h.w >> 16) & 0xffu, fl = h.w >> 24;
      blk0 = __ldg(&p.blk_off[r]);
      double pid = __ddiv_rn((double)(100 * ((int)alen - (int)nm)), (double)alen);
      if (fl & MB_RF_NMSAT) pid = -__longlong_as_double(0x7ff0000000000000ll);
      ok = !(pid < p.mapid) && !((int)mq < p.mapq) && !(__ddiv_rn((double)alen, (double)lq) < p.aln_cov);
    }
    // B: quality sums, two reads per step
    uint32_t need = __ballot_sync(0xffffffffu, ok);
    uint32_t my_sum = 0;
    while (need) {
      // pick up to two reads: lowest set bits
      const int i0 = __ffs(need) - 1;
      need &= need - 1;
      int i1 = -1;
      if (need) { i1 = __ffs(need) - 1; need &= need - 1; }
      const int src = sub == 0 ? i0 : i1;
      const int srcc = src < 0 ? 0 : src;
      const uint32_t b0 = __shfl_sync(0xffffffffu, blk0, srcc);
      const uint32_t l = __shfl_sync(0xffffffffu, lq, srcc);
      uint32_t s = 0;
      if (src >= 0) {
        const uint32_t nblk = (l + 15u) >> 4;
        const uint4* __restrict__ q = reinterpret_cast<const uint4*>(p.qual) + b0;
        for (uint32_t b = sl; b < nblk; b += 16) {
          const uint4 v = __ldg(q + b);
          s += __vsadu4(v.x & 0x7f7f7f7fu, 0u) + __vsadu4(v.y & 0x7f7f7f7fu, 0u) +
               __vsadu4(v.z & 0x7f7f7f7fu, 0u) + __vsadu4(v.w & 0x7f7f7f7fu, 0u);
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      // hand the sums back to the owning lanes
      const uint32_t s0 = __shfl_sync(0xffffffffu, s, 0);
      const uint32_t s1 = __shfl_sync(0xffffffffu, s, 16);
      if (lane == i0) my_sum = s0;
      if (lane == i1) my_sum = s1;
    }
    // C
    if (valid && gene < p.n_genes) {
      const bool keep = ok && ((long long)my_sum >= (long long)p.readq * (long long)lq);
      unsigned long long* o = p.out + (size_t)gene * 3;
      atomicAdd(o, 1ull | ((unsigned long long)(keep ? 1u : 0u) << 32));  // genes.py:181-182,187-188
      if (keep) atomicAdd(o + 1, (unsigned long long)alen);                  // numerator of genes.py:189
      if (p.key) { p.key[r] = keep ? gene : 0xFFFFFFFFu; p.val[r] = alen; }
    } else if (valid && p.key) {
      p.key[r] = 0xFFFFFFFFu;
      p.val[r] = 0;
    }
  }
}

__global__ void k_genes_finalize(unsigned long long* out, const uint32_t* __restrict__ gene_len, uint32_t n_genes) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genes) return;
  const unsigned long long bp = out[(size_t)g * 3 + 1];
  const uint32_t gl = gene_len[g];
  const double depth = gl ? __ddiv_rn((double)bp, (double)gl) : 0.0;
  out[(size_t)g * 3 + 2] = (unsigned long long)__double_as_longlong(depth);
}

// exact_order: the reference adds len(query_alignment_sequence)/float(gene.length) to gene.depth
// once per kept record in BAM file order (genes.py:178-189). After a stable sort by gene the kept
// reads of a gene are contiguous and still in file order; one thread per gene repeats the same
// sequence of IEEE-754 double divisions and additions.
__global__ void k_genes_exact(const uint32_t* __restrict__ key, const uint32_t* __restrict__ val, uint64_t n_reads,
                              unsigned long long* out, const uint32_t* __restrict__ gene_len, uint32_t n_genes) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genes) return;
  uint64_t lo = 0, hi = n_reads;
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (key[mid] < g) lo = mid + 1; else hi = mid;
  }
  const double gl = (double)gene_len[g];
  double depth = 0.0;
  for (uint64_t i = lo; i < n_reads && key[i] == g; ++i) depth = __dadd_rn(depth, __ddiv_rn((double)val[i], gl));
  out[(size_t)g * 3 + 2] = (unsigned long long)__double_as_longlong(depth);
}

}  // namespace

static_assert(sizeof(mb_gene_out) == 24, "mb_gene_out must be 24 bytes");

extern "C" int mb_genes_coverage_dev(mb_ctx* ctx, const mb_dev_reads* reads, const uint32_t* gene_len_dev,
                                     uint32_t n_genes, const mb_filters* f, int exact_order, mb_gene_out* out_dev) {
  if (!ctx || !reads || !gene_len_dev || !f || !out_dev)
    return mb::fail(MB_EINVAL, "mb_genes_coverage_dev: null argument");
  MB_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  uint32_t *key = nullptr, *val = nullptr, *key2 = nullptr, *val2 = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  const uint64_t n = reads->n_reads;
  if (exact_order && n) {
    if (n > 0x7fffffffull) return mb::fail(MB_EOVERFLOW, "mb_genes_coverage_dev: exact_order supports < 2^31 reads");
    void* b[4];
    for (int k = 0; k < 4; ++k) { int rc = mb::ctx_reserve(ctx, 4 + k, sizeof(uint32_t) * n, &b[k]); if (rc) return rc; }
    key = (uint32_t*)b[0]; val = (uint32_t*)b[1]; key2 = (uint32_t*)b[2]; val2 = (uint32_t*)b[3];
    MB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key2, val, val2, (int)n, 0, 32, st));
    void* t = nullptr;
    int rc = mb::ctx_reserve(ctx, 15, tmp_bytes, &t);
    if (rc) return rc;
    tmp = t;
  }
  MB_CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(mb_gene_out) * (size_t)n_genes, st));
  GeneParams p;
  p.hdr = reinterpret_cast<const uint4*>(reads->hdr);
  p.blk_off = reads->blk_off;
  p.qual = reads->qual;
  p.n_reads = reads->n_reads;
  p.out = reinterpret_cast<unsigned long long*>(out_dev);
  p.key = key;
  p.val = val;
  p.n_genes = n_genes;
  p.mapid = f->mapid;
  p.aln_cov = f->aln_cov;
  p.readq = f->readq;
  p.mapq = f->mapq;
  { int trc = mb::timing_begin(ctx); if (trc) return trc; }
  if (reads->n_reads) {
    uint64_t want = (reads->n_reads + GENE_THREADS - 1) / GENE_THREADS;
    uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * 8);
    k_genes_v1<<<grid, GENE_THREADS, 0, st>>>(p);
    ctx->launches++;
  }
  { int trc = mb::timing_end(ctx); if (trc) return trc; }
  if (n_genes) {
    k_genes_finalize<<<(n_genes + 255) / 256, 256, 0, st>>>(p.out, gene_len_dev, n_genes);
    ctx->launches++;
    if (exact_order && n) {
      MB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key2, val, val2, (int)n, 0, 32, st));
      k_genes_exact<<<(n_genes + 127) / 128, 128, 0, st>>>(key2, val2, n, p.out, gene_len_dev, n_genes);
      ctx->launches++;
    }
  }
  MB_CUDA_TRY(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_genes_coverage(mb_ctx* ctx, const mb_packed_reads* reads, const uint32_t* gene_len_host,
                                 uint32_t n_genes, const mb_filters* f, int exact_order, mb_gene_out* out_host) {
  if (!ctx || !reads || !gene_len_host || !f || !out_host)
    return mb::fail(MB_EINVAL, "mb_genes_coverage: null argument");
  MB_CUDA_TRY(cudaSetDevice(ctx->device));
  mb_dev_reads dr;
  mb_dev_reads* d = &dr;
  int rc = mb::upload_reads(ctx, reads, d, /*staged=*/true);
  if (rc) return rc;
  void *gl = nullptr, *out = nullptr;
  rc = mb::ctx_reserve(ctx, 2, sizeof(uint32_t) * (size_t)n_genes, &gl);
  if (!rc) rc = mb::ctx_reserve(ctx, 3, sizeof(mb_gene_out) * (size_t)n_genes, &out);
  if (!rc && n_genes) {
    cudaError_t e = cudaMemcpyAsync(gl, gene_len_host, sizeof(uint32_t) * (size_t)n_genes, cudaMemcpyHostToDevice,
                                    ctx->stream);
    if (e != cudaSuccess) rc = mb::fail(MB_ECUDA, "mb_genes_coverage: %s", cudaGetErrorString(e));
  }
  if (!rc) rc = mb_genes_coverage_dev(ctx, d, (const uint32_t*)gl, n_genes, f, exact_order, (mb_gene_out*)out);
  if (!rc) {
    cudaError_t e = cudaSuccess;
    if (n_genes)
      e = cudaMemcpyAsync(out_host, out, sizeof(mb_gene_out) * (size_t)n_genes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = mb::fail(MB_ECUDA, "mb_genes_coverage: %s", cudaGetErrorString(e));
  }
  return rc;
}
