// Context, device-resident reads and device mirrors of the reference layout.
#include <cstdio>
#include <cstring>
#include <climits>

#include <chrono>

#include "common.h"

namespace mb {

int ctx_reserve(mb_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes == 0) bytes = 256;
  if (ctx->stage_cap[slot] < bytes) {
    if (ctx->stage[slot]) {
      MB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      MB_CUDA_TRY(cudaFree(ctx->stage[slot]));
      ctx->stage[slot] = nullptr;
      ctx->stage_cap[slot] = 0;
    }
    size_t cap = bytes + bytes / 8;
    cap = (cap + 255) / 256 * 256;
    MB_CUDA_TRY(cudaMalloc(&ctx->stage[slot], cap));
    ctx->stage_cap[slot] = cap;
  }
  *out = ctx->stage[slot];
  return MB_OK;
}

int get_dev_layout(mb_ctx* ctx, const mb_layout* L, mb_dev_layout* out) {
  auto it = ctx->layouts.find(L->id);
  if (it != ctx->layouts.end()) { *out = it->second; return MB_OK; }
  mb_dev_layout d{};
  size_t nt = L->n_tiles ? L->n_tiles : 1, ns = L->n_species ? L->n_species : 1, nc = L->n_contigs ? L->n_contigs : 1;
  MB_CUDA_TRY(cudaMalloc(&d.tile_species, sizeof(int32_t) * nt));
  MB_CUDA_TRY(cudaMalloc(&d.species_len, sizeof(uint64_t) * ns));
  MB_CUDA_TRY(cudaMalloc(&d.contig_len, sizeof(uint32_t) * nc));
  MB_CUDA_TRY(cudaMalloc(&d.contig_goff, sizeof(uint32_t) * nc));
  MB_CUDA_TRY(cudaMalloc(&d.order, sizeof(uint32_t) * nc));
  if (L->n_contigs && !L->genes_only) {
    MB_CUDA_TRY(cudaMemcpyAsync(d.contig_len, L->contig_len.data(), sizeof(uint32_t) * L->n_contigs, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA_TRY(cudaMemcpyAsync(d.contig_goff, L->contig_goff.data(), sizeof(uint32_t) * L->n_contigs, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA_TRY(cudaMemcpyAsync(d.order, L->order.data(), sizeof(uint32_t) * L->n_contigs, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (L->n_tiles)
    MB_CUDA_TRY(cudaMemcpyAsync(d.tile_species, L->tile_species.data(), sizeof(int32_t) * L->n_tiles,
                                cudaMemcpyHostToDevice, ctx->stream));
  if (L->n_species)
    MB_CUDA_TRY(cudaMemcpyAsync(d.species_len, L->species_len.data(), sizeof(uint64_t) * L->n_species,
                                cudaMemcpyHostToDevice, ctx->stream));
  // the host vectors outlive the copy only if the layout does: make the upload complete now
  MB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->layouts[L->id] = d;
  *out = d;
  return MB_OK;
}

// The reference's two floating-point read filters compare an IEEE-754 double quotient of small
// integers with a user threshold (snps.py:148,157):
//     100*(alen-NM)/float(alen) < mapid        alen/float(l_qseq) < aln_cov
// Both quotients are monotone in their integer numerator, so for every denominator there is one
// smallest numerator that passes. The host finds it with the same correctly-rounded division the
// reference's interpreter performs, and the kernels compare integers:
//     t_pid[alen] = min num  with (double)num / alen  >= mapid   -> keep iff 100*(alen-NM) >= t_pid[alen]
//     t_cov[lq]   = min alen with (double)alen / lq   >= aln_cov -> keep iff alen >= t_cov[lq]
int get_filter_tables(mb_ctx* c, const mb_filters* f, const int32_t** t_pid, const int32_t** t_cov) {
  const int N = 65536;
  if (!c->ftab) MB_CUDA_TRY(cudaMalloc(&c->ftab, sizeof(int32_t) * 2 * N));
  const bool same = c->ftab_valid && memcmp(&c->ftab_mapid, &f->mapid, sizeof(double)) == 0 &&
                    memcmp(&c->ftab_cov, &f->aln_cov, sizeof(double)) == 0;
  if (!same) {
    std::vector<int32_t> h(2 * N);
    const double mapid = f->mapid, cov = f->aln_cov;
    h[0] = INT32_MAX;  // alen = 0 never reaches the kernels (MB_EINVAL at pack time)
    h[N] = INT32_MAX;
    for (int d = 1; d < N; ++d) {
      const double dd = (double)d;
      {  // numerators 100*(alen-NM), NM in [0,65535]
        const long long lo = 100LL * (d - 65535), hi = 100LL * d;
        int32_t t;
        if (mapid != mapid || (double)lo / dd >= mapid) t = INT32_MIN;
        else if (!((double)hi / dd >= mapid)) t = INT32_MAX;
        else {
          long long a = lo, b = hi;  // a fails, b passes
          while (b - a > 1) { const long long m = a + (b - a) / 2; if ((double)m / dd >= mapid) b = m; else a = m; }
          t = (int32_t)b;
        }
        h[d] = t;
      }
      {  // numerators alen in [0,65535]
        int32_t t;
        if (cov != cov || 0.0 / dd >= cov) t = 0;
        else if (!(65535.0 / dd >= cov)) t = INT32_MAX;
        else {
          long long a = 0, b = 65535;
          while (b - a > 1) { const long long m = a + (b - a) / 2; if ((double)m / dd >= cov) b = m; else a = m; }
          t = (int32_t)b;
        }
        h[N + d] = t;
      }
    }
    MB_CUDA_TRY(cudaStreamSynchronize(c->stream));  // a previous launch may still read the old tables
    MB_CUDA_TRY(cudaMemcpy(c->ftab, h.data(), sizeof(int32_t) * 2 * N, cudaMemcpyHostToDevice));
    c->ftab_mapid = f->mapid;
    c->ftab_cov = f->aln_cov;
    c->ftab_valid = true;
  }
  *t_pid = c->ftab;
  *t_cov = c->ftab + N;
  return MB_OK;
}

int timing_begin(mb_ctx* c) {
  const int slot = (int)(c->ring_n % mb_ctx::kRing);
  if (c->ring_n >= (uint64_t)mb_ctx::kRing) {  // fold the pair about to be overwritten
    float ms = 0.f;
    MB_CUDA_TRY(cudaEventSynchronize(c->ring[slot][1]));
    MB_CUDA_TRY(cudaEventElapsedTime(&ms, c->ring[slot][0], c->ring[slot][1]));
    c->ring_acc_ms += ms;
    c->ring_acc_n += 1;
  }
  MB_CUDA_TRY(cudaEventRecord(c->ring[slot][0], c->stream));
  return MB_OK;
}

int timing_end(mb_ctx* c) {
  const int slot = (int)(c->ring_n % mb_ctx::kRing);
  MB_CUDA_TRY(cudaEventRecord(c->ring[slot][1], c->stream));
  c->ring_n += 1;
  return MB_OK;
}

}  // namespace mb

extern "C" int mb_ctx_create(int device, mb_ctx** out) {
  if (!out) return mb::fail(MB_EINVAL, "mb_ctx_create: null argument");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return mb::fail(MB_ECUDA, "mb_ctx_create: no CUDA device (%s); libmidas_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return mb::fail(MB_EINVAL, "mb_ctx_create: device %d not in [0,%d)", device, n);
  MB_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  MB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return mb::fail(MB_ECUDA, "mb_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
  mb_ctx* c = new mb_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->ring_n = 0;
  c->ring_acc_ms = 0.0;
  c->ring_acc_n = 0;
  c->launches = 0;
  c->pile_smem = 0;
  c->pile_occ = 1;
  c->tile_rng = nullptr;
  c->tile_rng_cap = 0;
  c->rng_for = nullptr;
  c->rng_W = 0;
  c->rng_ntiles = 0;
  c->bam_resident = nullptr;
  c->units = nullptr;
  c->units_cap = 0;
  c->units_layout = ~0ull;
  c->units_len = 0;
  c->n_units = 0;
  c->sched = nullptr;
  c->stats_acc = nullptr;
  c->stats_acc_cap = 0;
  c->ftab = nullptr;
  c->ftab_valid = false;
  memset(c->stage, 0, sizeof(c->stage));
  memset(c->stage_cap, 0, sizeof(c->stage_cap));
  c->pack_meta = nullptr;
  c->pack_meta_host = nullptr;
  c->pack_timed = false;
  c->ovf = nullptr;
  c->ovf_host = nullptr;
  c->stats_mirror = nullptr;
  MB_CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  MB_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i <= mb_ctx::kChunks; ++i) MB_CUDA_TRY(cudaEventCreateWithFlags(&c->chunk_ev[i], cudaEventDisableTiming));
  MB_CUDA_TRY(cudaMalloc(&c->pack_meta, 2048));
  MB_CUDA_TRY(cudaHostAlloc(&c->pack_meta_host, 2048, cudaHostAllocDefault));
  MB_CUDA_TRY(cudaEventCreate(&c->pack_ev[0]));
  MB_CUDA_TRY(cudaEventCreate(&c->pack_ev[1]));
  MB_CUDA_TRY(cudaMalloc(&c->ovf, 256));
  MB_CUDA_TRY(cudaMemset(c->ovf, 0, 256));
  MB_CUDA_TRY(cudaHostAlloc((void**)&c->ovf_host, 256, cudaHostAllocDefault));
  c->ovf_host[0] = 0;
  for (int i = 0; i < mb_ctx::kRing; ++i) {
    MB_CUDA_TRY(cudaEventCreate(&c->ring[i][0]));
    MB_CUDA_TRY(cudaEventCreate(&c->ring[i][1]));
  }
  MB_CUDA_TRY(cudaMalloc(&c->sched, 256));
  MB_CUDA_TRY(cudaMemset(c->sched, 0, 256));        // the pileup kernel's last CTA leaves the counters at zero again
  *out = c;
  return MB_OK;
}

extern "C" void mb_ctx_destroy(mb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  const bool timing = getenv("MB_CTX_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = now();
  auto phase = [&](const char* what) { if (timing) { const double t = now(); fprintf(stderr, "mb_ctx_destroy: %-10s %8.1f ms\n", what, t - t0); t0 = t; } };
  for (auto& kv : c->layouts) {
    cudaFree(kv.second.tile_species);
    cudaFree(kv.second.species_len);
    cudaFree(kv.second.contig_len);
    cudaFree(kv.second.contig_goff);
    cudaFree(kv.second.order);
  }
  phase("layouts");
  for (int i = 0; i < mb_ctx::kStage; ++i)
    if (c->stage[i]) {
      cudaFree(c->stage[i]);
      if (timing) { const double t = now(); fprintf(stderr, "mb_ctx_destroy: stage %2d (%7.1f MB) %8.1f ms\n", i, c->stage_cap[i] / 1e6, t - t0); t0 = t; }
    }
  cudaStreamDestroy(c->copy_stream);
  for (int i = 0; i <= mb_ctx::kChunks; ++i) cudaEventDestroy(c->chunk_ev[i]);
  if (c->pack_meta) cudaFree(c->pack_meta);
  if (c->pack_meta_host) cudaFreeHost(c->pack_meta_host);
  cudaEventDestroy(c->pack_ev[0]);
  cudaEventDestroy(c->pack_ev[1]);
  if (c->ovf) cudaFree(c->ovf);
  if (c->ovf_host) cudaFreeHost(c->ovf_host);
  if (c->tile_rng) cudaFree(c->tile_rng);
  if (c->units) cudaFree(c->units);
  if (c->sched) cudaFree(c->sched);
  if (c->stats_acc) cudaFree(c->stats_acc);
  if (c->ftab) cudaFree(c->ftab);
  for (int i = 0; i < mb_ctx::kRing; ++i) {
    cudaEventDestroy(c->ring[i][0]);
    cudaEventDestroy(c->ring[i][1]);
  }
  cudaStreamDestroy(c->own_stream);
  phase("rest");
  delete c;
}

extern "C" int mb_ctx_set_stream(mb_ctx* c, void* s) {
  if (!c) return mb::fail(MB_EINVAL, "mb_ctx_set_stream: null context");
  MB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return MB_OK;
}

extern "C" int mb_ctx_sync(mb_ctx* c) {
  if (!c) return mb::fail(MB_EINVAL, "mb_ctx_sync: null context");
  MB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return MB_OK;
}

extern "C" int mb_ctx_device(const mb_ctx* c) { return c ? c->device : -1; }

namespace mb {
int check_overflow(mb_ctx* c, const char* who) {
  MB_CUDA_TRY(cudaMemcpyAsync(c->ovf_host, c->ovf, 4, cudaMemcpyDeviceToHost, c->stream));
  MB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (c->ovf_host[0]) {
    c->ovf_host[0] = 0;
    MB_CUDA_TRY(cudaMemsetAsync(c->ovf, 0, 4, c->stream));
    return mb::fail(MB_EOVERFLOW, "%s: a per-position u32 counter overflowed (saturated at 4294967295)", who);
  }
  return MB_OK;
}
}  // namespace mb

extern "C" int mb_ctx_overflowed(mb_ctx* c, int* flag) {
  if (!c || !flag) return mb::fail(MB_EINVAL, "mb_ctx_overflowed: null argument");
  MB_CUDA_TRY(cudaSetDevice(c->device));
  const int rc = mb::check_overflow(c, "mb_ctx_overflowed");
  if (rc == MB_EOVERFLOW) { *flag = 1; return MB_OK; }
  *flag = 0;
  return rc;
}

extern "C" int mb_ctx_set_stats_mirror(mb_ctx* c, mb_species_stats* peer_rows) {
  if (!c) return mb::fail(MB_EINVAL, "mb_ctx_set_stats_mirror: null context");
  c->stats_mirror = peer_rows;
  return MB_OK;
}

extern "C" int mb_host_alloc(size_t bytes, void** out) {
  if (!out) return mb::fail(MB_EINVAL, "mb_host_alloc: null argument");
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); return mb::fail(MB_ECUDA, "mb_host_alloc(%zu): %s", bytes, cudaGetErrorString(e)); }
  return MB_OK;
}
extern "C" void mb_host_free(void* p) { if (p) cudaFreeHost(p); }
extern "C" uint64_t mb_ctx_launch_count(const mb_ctx* c) { return c ? c->launches : 0; }

extern "C" int mb_ctx_last_kernel_ms(mb_ctx* c, float* ms) {
  if (!c || !ms) return mb::fail(MB_EINVAL, "mb_ctx_last_kernel_ms: null argument");
  if (c->ring_n == 0) return mb::fail(MB_EINVAL, "mb_ctx_last_kernel_ms: no kernel has been timed yet");
  const int slot = (int)((c->ring_n - 1) % mb_ctx::kRing);
  MB_CUDA_TRY(cudaEventSynchronize(c->ring[slot][1]));
  MB_CUDA_TRY(cudaEventElapsedTime(ms, c->ring[slot][0], c->ring[slot][1]));
  return MB_OK;
}

extern "C" int mb_ctx_timing_reset(mb_ctx* c) {
  if (!c) return mb::fail(MB_EINVAL, "mb_ctx_timing_reset: null context");
  MB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->ring_n = 0;
  c->ring_acc_ms = 0.0;
  c->ring_acc_n = 0;
  return MB_OK;
}

extern "C" int mb_ctx_timing_read(mb_ctx* c, double* total_ms, uint64_t* n) {
  if (!c || !total_ms || !n) return mb::fail(MB_EINVAL, "mb_ctx_timing_read: null argument");
  MB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  double acc = c->ring_acc_ms;
  uint64_t cnt = c->ring_acc_n;
  const uint64_t live = c->ring_n < (uint64_t)mb_ctx::kRing ? c->ring_n : (uint64_t)mb_ctx::kRing;
  for (uint64_t k = 0; k < live; ++k) {
    const int slot = (int)((c->ring_n - 1 - k) % mb_ctx::kRing);
    float ms = 0.f;
    MB_CUDA_TRY(cudaEventElapsedTime(&ms, c->ring[slot][0], c->ring[slot][1]));
    acc += ms;
    cnt += 1;
  }
  *total_ms = acc;
  *n = cnt;
  return MB_OK;
}

// ------------------------------------------------------------------ device-resident reads
namespace mb {
// Copy a packed batch to the device. staged: into context-owned buffers that are reused by the
// next host-buffer call (no cudaMalloc/cudaFree on the e2e path); otherwise fresh allocations.
int upload_reads(mb_ctx* c, const mb_packed_reads* h, mb_dev_reads* d, bool staged) {
  memset(d, 0, sizeof(*d));
  d->owned = !staged;
  d->n_reads = h->n_reads;
  d->n_vec = h->n_vec;
  d->q_planes = h->q_planes;
  d->n_planes = h->n_planes;
  d->max_units = h->max_units;
  d->n_cigar_words = h->n_cigar_words;
  d->total_bases = h->total_bases;
  d->max_ref_span = h->max_ref_span;
  d->sorted = h->sorted;
  d->row_words = h->row_words;
  d->n_long = h->n_long;
  struct Part { void** dst; const void* src; size_t bytes; } parts[6] = {
      {(void**)&d->long_idx, h->long_idx, sizeof(uint32_t) * (size_t)h->n_long},
      {(void**)&d->hdr, h->hdr, sizeof(mb_read_hdr) * (size_t)h->n_reads},
      {(void**)&d->vec_off, h->vec_off, sizeof(uint32_t) * (size_t)(h->n_reads + 1)},
      {(void**)&d->cplx, h->cplx, sizeof(uint32_t) * (size_t)h->n_reads},
      {(void**)&d->cigar, h->cigar, sizeof(uint32_t) * (size_t)h->n_cigar_words},
      {(void**)&d->planes, h->planes, (size_t)h->n_vec * 32}};
  int slot = 8;
  for (auto& p : parts) {
    // +256 B of zero slack so vectorised tail loads never leave the allocation
    if (staged) {
      int rc = ctx_reserve(c, slot++, p.bytes + 256, p.dst);
      if (rc) return rc;
    } else {
      cudaError_t e = cudaMalloc(p.dst, p.bytes + 256);
      if (e != cudaSuccess)
        return mb::fail(MB_ENOMEM, "mb_reads_upload: cudaMalloc(%zu) failed: %s", p.bytes, cudaGetErrorString(e));
    }
    MB_CUDA_TRY(cudaMemsetAsync((char*)*p.dst + p.bytes, 0, 256, c->stream));
    if (p.bytes) MB_CUDA_TRY(cudaMemcpyAsync(*p.dst, p.src, p.bytes, cudaMemcpyHostToDevice, c->stream));
    d->bytes += p.bytes;
  }
  return MB_OK;
}
}  // namespace mb

extern "C" int mb_reads_upload(mb_ctx* c, const mb_packed_reads* h, mb_dev_reads** out) {
  if (!c || !h || !out) return mb::fail(MB_EINVAL, "mb_reads_upload: null argument");
  MB_CUDA_TRY(cudaSetDevice(c->device));
  mb_dev_reads* d = new mb_dev_reads();
  int rc = mb::upload_reads(c, h, d, false);
  if (rc) { mb_reads_free(c, d); return rc; }
  *out = d;
  return MB_OK;
}

extern "C" void mb_reads_free(mb_ctx* c, mb_dev_reads* d) {
  if (!d) return;
  if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); if (c->rng_for == d) c->rng_for = nullptr; }
  mb::free_dev_reads_arrays(d);
  delete d;
}

extern "C" uint64_t mb_reads_bytes(const mb_dev_reads* d) { return d ? d->bytes : 0; }
extern "C" uint64_t mb_reads_total_bases(const mb_dev_reads* d) { return d ? d->total_bases : 0; }
extern "C" uint64_t mb_reads_n(const mb_dev_reads* d) { return d ? d->n_reads : 0; }
