// Gene coverage on sm_100a: per-gene aligned/mapped reads and mapped bp.
//
// Replaces the record loop of count_mapped_bp (reference midas/run/genes.py:178-189) and its
// keep_read (genes.py:153-169). Reads arrive in BAM file order (unsorted); each read touches one
// gene, so the kernel is a streaming scan (16 B header + the read's qualities) followed by two
// L2-resident atomics per read into the 24-byte per-gene record.
//
// One thread per read: header-only filters (identity genes.py:157, mapq :163, alignment coverage :166),
// then the Phred bit planes of the read (genes-mode packing holds nothing else) as independent 128-bit
// loads: sum(q) = sum_k 2^k popc(plane k) (mean-quality test genes.py:160 in its exact integer form
// sum(q) >= readq*l_qseq), then the read's atomics: (aligned,mapped) as one packed u64 add, mapped_bp
// as another.
#include <cub/device/device_radix_sort.cuh>

#include <cmath>

#include "common.h"

namespace {

constexpr int GENE_THREADS = 256;

__device__ __forceinline__ void ldg256(const uint32_t* __restrict__ p, uint32_t* r) {
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "l"(p));
}

struct GeneParams {
  const uint4* hdr;
  const uint32_t* vec_off;
  const uint32_t* planes;   // genes-mode packing: q_planes words per 32-base block, regions in 32-byte units
  const uint32_t* zero_unit;  // 32 bytes of zeros (the slack behind planes[]): what a load past a read's end fetches
  uint64_t n_reads;
  unsigned long long* out;  // mb_gene_out as 3 x u64: [aligned|mapped<<32], mapped_bp, depth bits
  uint32_t* key;            // exact order: per read, gene index if kept else 0xFFFFFFFF (may be null)
  uint32_t* val;            // exact order: per read, aligned length
  uint32_t n_genes;
  const int32_t* t_pid;     // smallest 100*(alen-NM) that passes mapid, per alen (ctx.cu)
  const int32_t* t_cov;     // smallest alen that passes aln_cov, per l_qseq
  uint32_t pid_always, readq_eff;
  int mapq;
};

template <int NQ>
__global__ void __launch_bounds__(GENE_THREADS) k_genes_v2(const GeneParams p) {
  // One thread per read: its Phred planes are up to NQ independent 256-bit loads per step (a 150-base
  // read is 4 of them, all in flight at once); the units of 32 consecutive reads are interleaved in
  // memory, so each of a warp's loads reads 1 KB of consecutive bytes.
  // Per read three dependent round trips would be exposed (header -> filter tables -> planes); instead the header
  // of the thread's next read is requested one iteration ahead and the first NQ units of the planes are requested as
  // soon as the header is there, before the filter tables are consulted (a filtered read's planes are fetched and
  // dropped: the byte model counts every read's planes anyway).
  const uint64_t stride = (uint64_t)gridDim.x * GENE_THREADS;
  uint64_t r = (uint64_t)blockIdx.x * GENE_THREADS + threadIdx.x;
  uint4 nh = make_uint4(0, 0, 0, 0);
  uint32_t nv0 = 0;
  if (r < p.n_reads) { nh = __ldg(&p.hdr[r]); nv0 = __ldg(&p.vec_off[r]); }
  for (; r < p.n_reads; r += stride) {
    const uint4 h = nh;
    const uint32_t v0 = nv0;
    if (r + stride < p.n_reads) { nh = __ldg(&p.hdr[r + stride]); nv0 = __ldg(&p.vec_off[r + stride]); }
    const uint32_t gene = h.x, lq = h.y & 0xffffu, alen = h.y >> 16;
    const uint32_t nm = h.z >> 16, mq = (h.w >> 16) & 0xffu, fl = h.w >> 24;
    const uint32_t nv8 = (((lq + 31u) >> 5) * (uint32_t)NQ + 7u) >> 3;
    const uint32_t* __restrict__ q = p.planes + (size_t)v0 * 8;
    uint32_t v[8 * NQ];                                   // NQ 256-bit loads = 8 blocks: plane of word i is i % NQ
#pragma unroll
    for (int u = 0; u < NQ; ++u) ldg256((uint32_t)u < nv8 ? q + (size_t)u * (8 * MB_GROUP) : p.zero_unit, v + 8 * u);
    // exact integer forms of the two double-precision tests of keep_read (tables built in ctx.cu)
    bool ok = (fl & MB_RF_NMSAT) ? (p.pid_always != 0)
                                 : (100 * ((int)alen - (int)nm) >= __ldg(p.t_pid + alen));   // genes.py:157
    ok = ok && (int)mq >= p.mapq;                                                             // genes.py:163
    ok = ok && (int)alen >= __ldg(p.t_cov + lq);                                              // genes.py:166
    bool keep = false;
    if (ok) {
      // mean-quality test genes.py:160: mean(q) < readq  <=>  sum(q) < readq * l_qseq (exact). The two top planes
      // alone bound the sum from below and settle all but poor reads.
      const uint32_t thr = p.readq_eff * lq;
      uint32_t pc[NQ];
#pragma unroll
      for (int k = 0; k < NQ; ++k) pc[k] = 0;
#pragma unroll
      for (int i = 0; i < 8 * NQ; ++i)
        if (i % NQ >= NQ - 2) pc[i % NQ] += __popc(v[i]);
      keep = (pc[NQ - 1] << (NQ - 1)) + (pc[NQ - 2] << (NQ - 2)) >= thr;   // (of the first 256 bases: still a lower bound)
      if (!keep) {
#pragma unroll
        for (int i = 0; i < 8 * NQ; ++i)
          if (i % NQ < NQ - 2) pc[i % NQ] += __popc(v[i]);
        for (uint32_t b = NQ; b < nv8; b += NQ) {          // reads longer than 256 bases
#pragma unroll
          for (int u = 0; u < NQ; ++u) ldg256(b + (uint32_t)u < nv8 ? q + (size_t)(b + u) * (8 * MB_GROUP) : p.zero_unit, v + 8 * u);
#pragma unroll
          for (int i = 0; i < 8 * NQ; ++i) pc[i % NQ] += __popc(v[i]);
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < NQ; ++k) s += pc[k] << k;
        keep = s >= thr;
      }
    }
    if (gene < p.n_genes) {
      unsigned long long* o = p.out + (size_t)gene * 3;
      atomicAdd(o, 1ull | ((unsigned long long)(keep ? 1u : 0u) << 32));  // genes.py:181-182,187-188
      if (keep) atomicAdd(o + 1, (unsigned long long)alen);                  // numerator of genes.py:189
      if (p.key) { p.key[r] = keep ? gene : p.n_genes; p.val[r] = alen; }   // dropped reads sort behind every gene
    } else if (p.key) {
      p.key[r] = p.n_genes;
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
  if (reads->n_reads && (reads->n_planes != reads->q_planes || (reads->q_planes != 6 && reads->q_planes != 7)))
    return mb::fail(MB_EINVAL, "mb_genes_coverage_dev: reads were not packed in genes mode (Phred planes only)");
  MB_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  uint32_t *key = nullptr, *val = nullptr, *key2 = nullptr, *val2 = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  const uint64_t n = reads->n_reads;
  int key_bits = 1;                                   // sort keys are gene indices 0..n_genes (n_genes = dropped read)
  while (key_bits < 32 && (n_genes >> key_bits)) ++key_bits;
  if (exact_order && n) {
    if (n > 0x7fffffffull) return mb::fail(MB_EOVERFLOW, "mb_genes_coverage_dev: exact_order supports < 2^31 reads");
    void* b[4];
    for (int k = 0; k < 4; ++k) { int rc = mb::ctx_reserve(ctx, 4 + k, sizeof(uint32_t) * n, &b[k]); if (rc) return rc; }
    key = (uint32_t*)b[0]; val = (uint32_t*)b[1]; key2 = (uint32_t*)b[2]; val2 = (uint32_t*)b[3];
    MB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key2, val, val2, (int)n, 0, key_bits, st));
    void* t = nullptr;
    int rc = mb::ctx_reserve(ctx, 15, tmp_bytes, &t);
    if (rc) return rc;
    tmp = t;
  }
  MB_CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(mb_gene_out) * (size_t)n_genes, st));
  GeneParams p;
  p.hdr = reinterpret_cast<const uint4*>(reads->hdr);
  p.vec_off = reads->vec_off;
  p.planes = reads->planes;
  p.zero_unit = reads->planes + (size_t)reads->n_vec * 8;       // the uploads keep 256 zero bytes behind every array
  p.n_reads = reads->n_reads;
  p.out = reinterpret_cast<unsigned long long*>(out_dev);
  p.key = key;
  p.val = val;
  p.n_genes = n_genes;
  { int frc = mb::get_filter_tables(ctx, f, &p.t_pid, &p.t_cov); if (frc) return frc; }
  p.pid_always = (f->mapid != f->mapid || f->mapid == -HUGE_VAL) ? 1u : 0u;
  p.readq_eff = f->readq <= 0 ? 0u : (f->readq > 128 ? 128u : (uint32_t)f->readq);
  p.mapq = f->mapq;
  { int trc = mb::timing_begin(ctx); if (trc) return trc; }
  if (reads->n_reads) {
    uint64_t want = (reads->n_reads + GENE_THREADS - 1) / GENE_THREADS;
    uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * 8);
    if (reads->q_planes == 6) k_genes_v2<6><<<grid, GENE_THREADS, 0, st>>>(p);
    else k_genes_v2<7><<<grid, GENE_THREADS, 0, st>>>(p);
    ctx->launches++;
  }
  { int trc = mb::timing_end(ctx); if (trc) return trc; }
  if (n_genes) {
    k_genes_finalize<<<(n_genes + 255) / 256, 256, 0, st>>>(p.out, gene_len_dev, n_genes);
    ctx->launches++;
    if (exact_order && n) {
      MB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key2, val, val2, (int)n, 0, key_bits, st));
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

extern "C" int mb_genes_coverage_plain(mb_ctx* ctx, const mb_plain_reads* plain_host, const uint32_t* gene_len_host,
                                       uint32_t n_genes, const mb_filters* f, int exact_order, mb_gene_out* out_host) {
  if (!ctx || !plain_host || !gene_len_host || !f || !out_host)
    return mb::fail(MB_EINVAL, "mb_genes_coverage_plain: null argument");
  MB_CUDA_TRY(cudaSetDevice(ctx->device));
  mb_layout* L = nullptr;
  int rc = mb_layout_create_genes(n_genes, &L);
  if (rc) return rc;
  mb_plain_reads dev;
  rc = mb::upload_plain(ctx, plain_host, &dev);                         // H2D of the decoded records
  mb_dev_reads dr;
  if (!rc) rc = mb::pack_on_device(ctx, &dev, L, 1, /*staged=*/true, &dr);   // the re-layout, on the GPU
  mb_layout_destroy(L);
  if (rc) return rc;
  void *gl = nullptr, *out = nullptr;
  rc = mb::ctx_reserve(ctx, 2, sizeof(uint32_t) * (size_t)n_genes, &gl);
  if (!rc) rc = mb::ctx_reserve(ctx, 3, sizeof(mb_gene_out) * (size_t)n_genes, &out);
  if (!rc && n_genes) {
    cudaError_t e = cudaMemcpyAsync(gl, gene_len_host, sizeof(uint32_t) * (size_t)n_genes, cudaMemcpyHostToDevice,
                                    ctx->stream);
    if (e != cudaSuccess) rc = mb::fail(MB_ECUDA, "mb_genes_coverage_plain: %s", cudaGetErrorString(e));
  }
  if (!rc) rc = mb_genes_coverage_dev(ctx, &dr, (const uint32_t*)gl, n_genes, f, exact_order, (mb_gene_out*)out);
  if (!rc) {
    cudaError_t e = cudaSuccess;
    if (n_genes)
      e = cudaMemcpyAsync(out_host, out, sizeof(mb_gene_out) * (size_t)n_genes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = mb::fail(MB_ECUDA, "mb_genes_coverage_plain: %s", cudaGetErrorString(e));
  }
  return rc;
}
