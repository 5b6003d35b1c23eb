// Device packer: decoded BAM records (mb_plain_reads) -> the packed struct-of-arrays the pileup kernels
// stream (mb_packed_reads layout, include/midas_b200.h), entirely on the GPU.
//
// This is the step pysam/htslib performs for the reference before its loops see an AlignedSegment
// (reference midas/run/snps.py:186,194; midas/run/genes.py:175-180): per read the two CIGAR-only
// quantities keep_read needs (alen = len(query_alignment_sequence) snps.py:145, query_length :146),
// validation of the inputs on which the reference raises, and the re-layout. csrc/host_pack.cpp is the
// host version of the same re-layout; tests/test_pack_gpu.py holds the two byte-identical.
//
//   k_pack_blocks    contig -> range of reads (binary search in the coordinate-sorted ref_id column) and
//                    where that block goes in layout order: a BAM is sorted by (tid, pos), the layout is
//                    species-major, so whole contigs move and nothing is sorted
//   k_pack_analyse   one thread per record: CIGAR walk (alen, q_start, ref_span clipped at the contig end,
//                    single-indel run word), validation, 16-byte header, need-words histogram
//   k_pack_qor       OR of every quality byte (6 or 7 Phred planes; a byte above 127 is an error)
//   k_pack_finish    row width of the tile kernel (all but 0.1 % of the reads fit), plane counts
//   k_pack_units / k_pack_scan{1,2,3}  32-byte units per group of 32 reads -> vec_off (exclusive scan)
//   k_pack_long      MB_RF_LONG flags + their index list
//   k_pack_fill      one warp per read, one lane per base: every plane word of a 32-base block is one
//                    __ballot_sync (sequence low / high bit, candidate, Phred bits), zero padding included
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.h"

namespace {

enum PackErr : uint32_t {
  PE_NONE = 0, PE_LQ0, PE_LQBIG, PE_SEQOFF, PE_NM, PE_NOQUAL, PE_NCIG, PE_OP, PE_QSUM, PE_ALEN0, PE_REFID, PE_POS,
  PE_SPAN, PE_CIGOFF, PE_SEQALIGN
};

struct PackMeta {
  unsigned long long err;          // min over failing records of (index << 8 | PackErr); ~0 = none
  unsigned long long total_bases;
  unsigned long long n_vec;
  unsigned long long n_cig;        // cigar_off[n]
  uint32_t unsorted, q_or, max_span, n_long, row_words, q_planes, n_planes, long_fill, max_units, pad_;
  uint32_t hist[256];
};

struct Plain {  // device view of mb_plain_reads
  uint64_t n;
  const int32_t *ref_id, *pos, *nm, *l_qseq;
  const uint8_t* mapq;
  const uint64_t *cigar_off, *seq_off;
  const uint64_t* qual_off;   // where a read's qualities start in qual[] (= seq_off, except for records parsed out of a BAM stream)
  const uint32_t* cigar;
  const uint8_t *seq, *qual;
  uint32_t nib;
  uint32_t qual_later;   // 1: the quality bytes are still in flight (overlapped transfer): k_pack_fill checks for 0xFF
};

__global__ void k_pack_meta_init(PackMeta* m) {   // a kernel, not a copy: nothing of the packer queues behind bulk transfers
  uint32_t* w = reinterpret_cast<uint32_t*>(m);
  for (uint32_t i = threadIdx.x; i < sizeof(PackMeta) / 4; i += blockDim.x) w[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) m->err = ~0ull;
}

__device__ __forceinline__ void report(PackMeta* m, uint64_t i, uint32_t code) {
  atomicMin(&m->err, (unsigned long long)((i << 8) | code));
}

// ---- contig blocks: source range of every contig and its destination in layout order (one block)
__global__ void __launch_bounds__(1024) k_pack_blocks(const int32_t* __restrict__ ref_id, uint64_t n, uint32_t n_contigs,
                                                      const uint32_t* __restrict__ order, uint32_t* blk_src,
                                                      uint32_t* blk_dst) {
  __shared__ unsigned long long s_scan[1024];
  __shared__ unsigned long long s_carry;
  for (uint32_t c = threadIdx.x; c < n_contigs; c += blockDim.x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (ref_id[mid] < (int32_t)c) lo = mid + 1; else hi = mid; }
    const uint64_t a = lo;
    hi = n;
    while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (ref_id[mid] <= (int32_t)c) lo = mid + 1; else hi = mid; }
    blk_src[c] = (uint32_t)a;
    blk_dst[c] = (uint32_t)(lo - a);   // count for now
  }
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_contigs; base += blockDim.x) {
    const uint32_t k = base + threadIdx.x;
    const uint32_t c = k < n_contigs ? order[k] : 0u;
    const unsigned long long v = k < n_contigs ? blk_dst[c] : 0ull;
    s_scan[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t o = 1; o < blockDim.x; o <<= 1) {
      const unsigned long long t = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0ull;
      __syncthreads();
      s_scan[threadIdx.x] += t;
      __syncthreads();
    }
    const unsigned long long incl = s_scan[threadIdx.x], carry = s_carry;
    __syncthreads();
    if (k < n_contigs) blk_dst[c] = (uint32_t)(carry + incl - v);
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + incl;
    __syncthreads();
  }
}

// ---- per record: geometry + validation (the checks and their order follow host_pack.cpp::analyse)
template <int MODE>
__global__ void __launch_bounds__(256) k_pack_analyse(const Plain in, const uint32_t n_contigs,
                                                      const uint32_t* __restrict__ contig_len,
                                                      const uint32_t* __restrict__ contig_goff,
                                                      const uint32_t* __restrict__ blk_src,
                                                      const uint32_t* __restrict__ blk_dst, uint4* hdr, uint32_t* cplx,
                                                      uint32_t* src_of, uint16_t* nblk, uint8_t* need, PackMeta* meta) {
  __shared__ uint32_t s_hist[256];
  __shared__ unsigned long long s_bases;
  __shared__ uint32_t s_span;
  for (int k = threadIdx.x; k < 256; k += blockDim.x) s_hist[k] = 0;
  if (threadIdx.x == 0) { s_bases = 0; s_span = 0; }
  __syncthreads();
  unsigned long long my_bases = 0;
  uint32_t my_span = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < in.n; i += stride) {
    const int64_t lq = in.l_qseq[i];
    const int32_t rid = in.ref_id[i];
    const int64_t pos = in.pos[i];
    if (MODE == 0 && i > 0) {   // BAM coordinate order
      const int32_t prid = in.ref_id[i - 1];
      if (prid > rid || (prid == rid && (int64_t)in.pos[i - 1] > pos)) meta->unsorted = 1;
    }
    if (lq <= 0) { report(meta, i, PE_LQ0); continue; }
    if (lq > 65535) { report(meta, i, PE_LQBIG); continue; }
    const uint64_t so = in.seq_off[i];
    if ((int64_t)(in.seq_off[i + 1] - so) < lq) { report(meta, i, PE_SEQOFF); continue; }
    if (in.nib && (so & 1)) { report(meta, i, PE_SEQALIGN); continue; }
    const int32_t nm_raw = in.nm[i];
    if (nm_raw < 0) { report(meta, i, PE_NM); continue; }
    if (!in.qual_later && in.qual[in.qual_off[i]] == 0xFF) { report(meta, i, PE_NOQUAL); continue; }
    const uint64_t co = in.cigar_off[i];
    const int64_t n_cig = (int64_t)(in.cigar_off[i + 1] - co);
    if (n_cig > 65535) { report(meta, i, PE_NCIG); continue; }
    if (co + (uint64_t)n_cig > 0xFFFFFFF0ull) { report(meta, i, PE_CIGOFF); continue; }
    const uint32_t* __restrict__ cig = in.cigar + co;
    int64_t qsum = 0, span = 0, n_match = 0, n_other = 0, lead = 0;
    bool bad_op = false, in_lead = true;
    for (int64_t k = 0; k < n_cig; ++k) {
      const uint32_t w = cig[k], op = w & 15u;
      const int64_t len = w >> 4;
      if (op == mb::OP_P || op > mb::OP_X) { bad_op = true; break; }
      if (op == mb::OP_M || op == mb::OP_EQ || op == mb::OP_X) { qsum += len; span += len; ++n_match; }
      else if (op == mb::OP_I) { qsum += len; ++n_other; }
      else if (op == mb::OP_S) { qsum += len; }
      else if (op == mb::OP_D || op == mb::OP_N) { span += len; ++n_other; }
      // pysam getQueryStart: soft clips before the first op that is neither S nor H
      if (in_lead) { if (op == mb::OP_S) lead += len; else if (op != mb::OP_H) in_lead = false; }
    }
    if (bad_op) { report(meta, i, PE_OP); continue; }
    if (n_cig > 0 && qsum != lq) { report(meta, i, PE_QSUM); continue; }
    int64_t trail = 0;   // pysam getQueryEnd: the walk from the end never visits op 0
    for (int64_t k = n_cig - 1; k >= 1; --k) {
      const uint32_t op = cig[k] & 15u;
      if (op == mb::OP_H) continue;
      if (op == mb::OP_S) trail += cig[k] >> 4; else break;
    }
    const int64_t alen = lq - lead - trail;
    if (alen <= 0) { report(meta, i, PE_ALEN0); continue; }
    uint32_t flags = 0;
    uint32_t nm = (uint32_t)nm_raw;
    if (nm > 65535u) { nm = 65535u; flags |= MB_RF_NMSAT; }
    if (rid < 0 || (uint32_t)rid >= n_contigs) { report(meta, i, PE_REFID); continue; }
    uint32_t loc, ref_span = 0, cw = 0xFFFFFFFFu, need_w = 0;
    uint64_t j = i;
    if (MODE == 1) {
      loc = (uint32_t)rid;
      flags |= MB_RF_SIMPLE;
    } else {
      const int64_t clen = contig_len[rid];
      if (pos < 0 || pos >= clen) { report(meta, i, PE_POS); continue; }
      loc = contig_goff[rid] + (uint32_t)pos;
      const int64_t room = clen - pos;
      const int64_t sc = span < room ? span : room;   // pysam ignores refpos >= contig_len
      if (sc > 65535) { report(meta, i, PE_SPAN); continue; }
      ref_span = (uint32_t)sc;
      const int64_t by_span = ((int64_t)(loc & 31u) + sc + 31) / 32, by_query = (lq + 31) / 32;
      const int64_t nw = by_span > by_query ? by_span : by_query;
      need_w = (uint32_t)(nw > 255 ? 255 : nw);
      if (n_other == 0 && n_match <= 1) {
        flags |= MB_RF_SIMPLE;
      } else {
        // the runs after clipping at the contig end: exactly two of the single-indel shape ride in the word
        int64_t q = 0, r = 0;
        uint32_t n_runs = 0, q0 = 0, len0 = 0, x0 = 0, q1 = 0, len1 = 0, x1 = 0;
        for (int64_t k = 0; k < n_cig && r < room; ++k) {
          const uint32_t w = cig[k], op = w & 15u;
          const int64_t len = w >> 4;
          if (op == mb::OP_M || op == mb::OP_EQ || op == mb::OP_X) {
            const int64_t take = len < room - r ? len : room - r;
            if (take > 0) {
              if (n_runs == 0) { q0 = (uint32_t)q; len0 = (uint32_t)take; x0 = (uint32_t)r; }
              else if (n_runs == 1) { q1 = (uint32_t)q; len1 = (uint32_t)take; x1 = (uint32_t)r; }
              ++n_runs;
            }
            q += len; r += len;
          } else if (op == mb::OP_I || op == mb::OP_S) q += len;
          else if (op == mb::OP_D || op == mb::OP_N) r += len;
        }
        cw = (uint32_t)co;
        if (n_runs == 2) {
          const int64_t ins = (int64_t)q1 - (int64_t)(q0 + len0), del = (int64_t)x1 - (int64_t)len0;
          if (q0 == (uint32_t)lead && x0 == 0 && ins >= 0 && ins <= 255 && del >= 0 && del <= 255 &&
              x1 + len1 == ref_span) {
            cw = len0 | ((uint32_t)ins << 16) | ((uint32_t)del << 24);
            flags |= MB_RF_RUNS2;
          }
        }
      }
      j = (uint64_t)blk_dst[rid] + (i - (uint64_t)blk_src[rid]);
      if (j >= in.n) { meta->unsorted = 1; continue; }   // ref_id column not ascending: the block table is meaningless
    }
    hdr[j] = make_uint4(loc, (uint32_t)lq | ((uint32_t)alen << 16), (uint32_t)lead | (nm << 16),
                        ref_span | ((uint32_t)in.mapq[i] << 16) | (flags << 24));
    cplx[j] = cw;
    src_of[j] = (uint32_t)i;
    nblk[j] = (uint16_t)((lq + MB_BLOCK - 1) / MB_BLOCK);
    if (MODE == 0) { need[j] = (uint8_t)need_w; atomicAdd(&s_hist[need_w], 1u); }
    my_bases += (unsigned long long)lq;
    my_span = max(my_span, ref_span);
  }
  atomicAdd(&s_bases, my_bases);
  atomicMax(&s_span, my_span);
  __syncthreads();
  if (MODE == 0)
    for (int k = threadIdx.x; k < 256; k += blockDim.x)
      if (s_hist[k]) atomicAdd(&meta->hist[k], s_hist[k]);
  if (threadIdx.x == 0) {
    if (s_bases) atomicAdd(&meta->total_bases, s_bases);
    if (s_span) atomicMax(&meta->max_span, s_span);
  }
}

// ---- OR of the whole quality array [seq_off[0], seq_off[n]) as one flat stream (padding elements included: they
// are zero from every producer in this repository; if a foreign producer left garbage there, bits 6-7 show up
// and the exact per-record pass below decides)
__global__ void __launch_bounds__(256) k_pack_qor_flat(const Plain in, PackMeta* meta, uint64_t r_lo, uint64_t r_hi) {
  const uintptr_t a0 = (uintptr_t)(in.qual + in.seq_off[r_lo]), a1 = (uintptr_t)(in.qual + in.seq_off[r_hi]);
  const uintptr_t v0 = (a0 + 15) & ~(uintptr_t)15, v1 = a1 & ~(uintptr_t)15;
  uint32_t acc = 0;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
  if (v1 > v0) {
    const uint4* __restrict__ p = reinterpret_cast<const uint4*>(v0);
    const uint64_t nv = (v1 - v0) >> 4;
    for (uint64_t k = t; k < nv; k += nt) { const uint4 w = __ldg(p + k); acc |= w.x | w.y | w.z | w.w; }
    if (t < 32) {                                   // the unaligned ends, byte by byte
      for (uintptr_t b = a0 + t; b < v0 && b < a1; b += 32) acc |= *reinterpret_cast<const uint8_t*>(b);
      for (uintptr_t b = v1 + t; b < a1; b += 32) acc |= *reinterpret_cast<const uint8_t*>(b);
    }
  } else if (t < 32) {
    for (uintptr_t b = a0 + t; b < a1; b += 32) acc |= *reinterpret_cast<const uint8_t*>(b);
  }
  acc |= acc >> 16;
  acc |= acc >> 8;
  acc &= 0xffu;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicOr(&meta->q_or, acc);
}

// ---- OR of every quality byte of every record: one warp per record, aligned 32-bit words, ends masked
__global__ void __launch_bounds__(256) k_pack_qor(const Plain in, PackMeta* meta) {
  const int lane = threadIdx.x & 31;
  const uint64_t wg = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t acc = 0;
  for (uint64_t i = wg; i < in.n; i += nw) {
    const int64_t lq = in.l_qseq[i];
    if (lq <= 0 || lq > 65535) continue;
    const uintptr_t a0 = (uintptr_t)(in.qual + in.qual_off[i]), a1 = a0 + (uintptr_t)lq;
    const uintptr_t w0 = a0 & ~(uintptr_t)3, w1 = (a1 + 3) & ~(uintptr_t)3;
    for (uintptr_t w = w0 + 4 * (uintptr_t)lane; w < w1; w += 128) {
      uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(w));
      if (w < a0) v &= 0xffffffffu << (8 * (uint32_t)(a0 - w));
      if (w + 4 > a1) v &= 0xffffffffu >> (8 * (uint32_t)(w + 4 - a1));
      acc |= v;
    }
  }
  acc |= acc >> 16;
  acc |= acc >> 8;
  acc &= 0xffu;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && acc) atomicOr(&meta->q_or, acc);
}

// ---- row width: the smallest W that all but 0.1 % of the reads fit (as host_pack.cpp); plane counts
__global__ void k_pack_finish(PackMeta* meta, uint64_t n, int mode, const uint64_t* __restrict__ cigar_off) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  meta->n_cig = n ? cigar_off[n] : 0ull;
  const uint32_t q_planes = (meta->q_or & 0x40u) ? 7u : 6u;
  meta->q_planes = q_planes;
  meta->n_planes = mode == 0 ? 3u + q_planes : q_planes;
  uint32_t row_words = 0;
  unsigned long long n_long = 0;
  if (mode == 0) {
    unsigned long long above = 0;
    for (int w = 255; w > (int)MB_ROW_WORDS_MAX; --w) above += meta->hist[w];
    row_words = MB_ROW_WORDS_MAX;
    for (uint32_t w = MB_ROW_WORDS_MAX; w > 2; --w) {
      if (above + meta->hist[w] > n / 1000) break;
      above += meta->hist[w];
      row_words = w - 1;
    }
    for (int w = (int)row_words + 1; w < 256; ++w) n_long += meta->hist[w];
  }
  meta->row_words = row_words;
  meta->n_long = (uint32_t)n_long;
}

// ---- 32-byte units per read of a group (= the largest of its 32 reads): one warp per group
__global__ void __launch_bounds__(256) k_pack_units(const uint16_t* __restrict__ nblk, uint64_t n, uint64_t n_groups,
                                                    PackMeta* meta, uint32_t* g_units) {
  const uint32_t NP = meta->n_planes;
  uint32_t umax = 0;
  const int lane = threadIdx.x & 31;
  const uint64_t wg = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t g = wg; g < n_groups; g += nw) {
    const uint64_t j = g * MB_GROUP + lane;
    uint32_t u = j < n ? ((uint32_t)nblk[j] * NP + 7u) / 8u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) u = max(u, __shfl_xor_sync(0xffffffffu, u, o));
    if (lane == 0) g_units[g] = u;
    umax = max(umax, u);
  }
  if (lane == 0 && umax) atomicMax(&meta->max_units, umax);
}

constexpr int SCAN_T = 1024, SCAN_PER = 4, SCAN_CHUNK = SCAN_T * SCAN_PER;

__device__ __forceinline__ unsigned long long block_incl_scan(unsigned long long v, unsigned long long* s_warp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) s_warp[wid] = v;
  __syncthreads();
  if (wid == 0) {
    unsigned long long w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  const unsigned long long r = v + (wid ? s_warp[wid - 1] : 0ull);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_T) k_pack_scan1(const uint32_t* __restrict__ g_units, uint64_t n_groups,
                                                       unsigned long long* bsum) {
  __shared__ unsigned long long s_warp[32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * SCAN_PER;
  unsigned long long v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER; ++k) if (base + k < n_groups) v += g_units[base + k];
  const unsigned long long incl = block_incl_scan(v, s_warp);
  if (threadIdx.x == SCAN_T - 1) bsum[blockIdx.x] = incl;
}

__global__ void __launch_bounds__(SCAN_T) k_pack_scan2(unsigned long long* bsum, uint32_t n_blocks, PackMeta* meta) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_blocks; base += SCAN_T) {
    const uint32_t k = base + threadIdx.x;
    const unsigned long long v = k < n_blocks ? bsum[k] : 0ull;
    const unsigned long long incl = block_incl_scan(v, s_warp);
    const unsigned long long carry = s_carry;
    __syncthreads();
    if (k < n_blocks) bsum[k] = carry + incl - v;
    if (threadIdx.x == SCAN_T - 1) s_carry = carry + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) meta->n_vec = s_carry * MB_GROUP;
}

// vec_off[j] = 32 * (units of all groups before j's) + j % 32; vec_off[n] = n_vec
__global__ void __launch_bounds__(SCAN_T) k_pack_scan3(const uint32_t* __restrict__ g_units, uint64_t n_groups,
                                                       const unsigned long long* __restrict__ bpre, uint64_t n,
                                                       const PackMeta* __restrict__ meta, uint32_t* vec_off) {
  __shared__ unsigned long long s_warp[32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * SCAN_PER;
  unsigned long long u[SCAN_PER], v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER; ++k) { u[k] = base + k < n_groups ? g_units[base + k] : 0u; v += u[k]; }
  unsigned long long excl = block_incl_scan(v, s_warp) - v + bpre[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_PER; ++k) {
    const uint64_t g = base + k;
    if (g < n_groups) {
      const uint64_t j0 = g * MB_GROUP, j1 = min(n, j0 + MB_GROUP);
      for (uint64_t j = j0; j < j1; ++j) vec_off[j] = (uint32_t)(excl * MB_GROUP + (j - j0));
    }
    excl += u[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) vec_off[n] = (uint32_t)meta->n_vec;
}

__global__ void __launch_bounds__(256) k_pack_long(const uint8_t* __restrict__ need, uint64_t n, PackMeta* meta,
                                                   uint4* hdr, uint32_t* long_idx) {
  const uint32_t rw = meta->row_words;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    if (need[j] > rw) {
      hdr[j].w |= (uint32_t)MB_RF_LONG << 24;
      long_idx[atomicAdd(&meta->long_fill, 1u)] = (uint32_t)j;
    }
  }
}

// ---- the planes: one warp per group of 32 reads. The lanes first fetch the 32 reads' metadata side by side (one
// latency chain per group); then read after read, eight 32-base blocks at a time. Lane l loads base l of every
// block of the chunk and packs one byte (quality) or one nibble (sequence class) per block into a 32-bit word;
// the warp's 32 words are a 32 x 32 bit matrix, and ONE warp-level bit transposition (5 shuffle steps) turns it
// into plane words: lane 8 s + k then holds Phred bit k of block s, lane 4 s + k the sequence plane k of block s.
// Padding words and the unused slots of the last group are written as zeros.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, const uint32_t lane) {
  // 32 x 32 bit matrix over the lanes of a warp: on return bit l of lane j = bit j of lane l on entry
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const uint32_t m = d == 16 ? 0x0000ffffu : d == 8 ? 0x00ff00ffu : d == 4 ? 0x0f0f0f0fu : d == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, d);
    x = (lane & d) ? (((y >> d) & m) | (x & ~m)) : ((x & m) | ((y << d) & ~m));
  }
  return x;
}

__device__ __forceinline__ uint32_t range_mask(int lo, int hi) {   // bits [lo, hi) of a word, clamped to [0, 32)
  lo = max(lo, 0);
  hi = min(max(hi, 0), 32);
  uint32_t up, dn;
  asm("shl.b32 %0, %1, %2;" : "=r"(up) : "r"(0xffffffffu), "r"((uint32_t)lo));
  asm("shr.u32 %0, %1, %2;" : "=r"(dn) : "r"(0xffffffffu), "r"((uint32_t)(32 - hi)));
  return hi > lo ? (up & dn) : 0u;
}

template <int MODE, int NQ>
__global__ void __launch_bounds__(256) k_pack_fill(const Plain in, const uint4* __restrict__ hdr,
                                                   const uint32_t* __restrict__ src_of,
                                                   const uint32_t* __restrict__ vec_off, uint64_t n,
                                                   PackMeta* meta, uint32_t* planes, uint64_t g_lo, uint64_t g_hi) {
  constexpr uint32_t P0 = MODE == 0 ? 3u : 0u, NP = P0 + NQ;
  const uint32_t n_vec = (uint32_t)meta->n_vec;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t wg = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t qs = lane >> 3, qk = lane & 7u;      // after a quality transposition: block qs of four, Phred bit qk
  const uint32_t ss = lane >> 2, sk = lane & 3u;      // after a sequence transposition: block ss of eight, plane sk
  for (uint64_t g = g_lo + wg; g < g_hi; g += nw) {
    const uint64_t g0 = g * MB_GROUP, g1 = g0 + MB_GROUP;
    const uint32_t v0 = __ldg(&vec_off[g0]);
    const uint32_t units = ((g1 < n ? __ldg(&vec_off[g1]) : n_vec) - v0) / MB_GROUP;
    // my read's metadata (lane l <-> read g0 + l)
    const uint64_t jm = g0 + lane;
    uint4 hm = make_uint4(0, 0, 0, 0);
    uint64_t so_m = 0, co_m = 0, qo_m = 0;
    if (jm < n) {
      hm = __ldg(&hdr[jm]);
      const uint64_t im = src_of[jm];
      so_m = in.seq_off[im];
      qo_m = in.qual_off[im];
      co_m = in.cigar_off[im];
    }
    for (uint32_t rr = 0; rr < MB_GROUP; ++rr) {
      uint32_t* __restrict__ reg = planes + ((size_t)v0 + rr) * 8;
      auto word = [&](const uint32_t w) -> uint32_t& { return reg[(size_t)(w >> 3) * (8 * MB_GROUP) + (w & 7u)]; };
      if (g0 + rr >= n) {                              // the unused slots of the last group: zeros, as the host packer
        for (uint32_t w = lane; w < units * 8; w += 32) word(w) = 0u;
        continue;
      }
      const uint32_t hy = __shfl_sync(0xffffffffu, hm.y, rr), hz = __shfl_sync(0xffffffffu, hm.z, rr);
      const uint32_t hw = __shfl_sync(0xffffffffu, hm.w, rr);
      const uint64_t so = __shfl_sync(0xffffffffu, so_m, rr), co = __shfl_sync(0xffffffffu, co_m, rr);
      const uint64_t qo = __shfl_sync(0xffffffffu, qo_m, rr);
      const uint32_t lq = hy & 0xffffu, qstart = hz & 0xffffu, span = hw & 0xffffu, fl = hw >> 24;
      const uint32_t n_blk = (lq + MB_BLOCK - 1) / MB_BLOCK;
      // my base of every block: quality byte at qb[32 b], sequence nibble in byte sb[16 b] (seq_off is even with nibble
      // packing, so the nibble within the byte depends on the lane alone)
      if (in.qual_later && lane == 0 && in.qual[qo] == 0xFF) report(meta, src_of[g0 + rr], PE_NOQUAL);
      const uint8_t* __restrict__ qb = in.qual + qo + lane;
      const uint8_t* __restrict__ sb = in.nib ? in.seq + ((so + lane) >> 1) : in.seq + so + lane;
      const uint32_t sb_step = in.nib ? 16u : 32u, nib_shift = (in.nib && !(lane & 1u)) ? 4u : 0u;
      const bool simple = (fl & MB_RF_SIMPLE) != 0;
      for (uint32_t b0 = 0; b0 < n_blk; b0 += 8) {
        // ---- loads: base `lane` of each of the chunk's (up to eight) blocks, all in flight at once
        uint32_t qa = 0, qb4 = 0, sn = 0;
        const uint8_t* __restrict__ qc = qb + (size_t)b0 * MB_BLOCK;
        const uint8_t* __restrict__ sc = sb + (size_t)b0 * sb_step;
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) {
          if (b0 + s8 >= n_blk) break;                                     // uniform: the read has no such block
          if ((b0 + s8) * MB_BLOCK + lane < lq) {
            const uint32_t q = qc[s8 * (int)MB_BLOCK];
            if (s8 < 4) qa |= q << (8 * s8); else qb4 |= q << (8 * (s8 - 4));
            if (MODE == 0) {
              const uint32_t nib = ((uint32_t)sc[s8 * sb_step] >> nib_shift) & 15u;
              // BAM nibble -> {low bit, high bit, is A/C/G/T}: A=1 C=2 G=4 T=8 -> 4, 5, 6, 7; anything else 0
              const uint32_t cls = nib == 8u ? 7u : (nib < 8u ? (0x00060540u >> (4 * nib)) & 7u : 0u);
              sn |= cls << (4 * s8);
            }
          }
        }
        uint32_t qb_ = qb4;
        // ---- quality planes: lane 8 s + k <- Phred bit k of block b0 + s (first four blocks), b0 + 4 + s (last four)
        qa = warp_transpose32(qa, lane);
        if (qk < (uint32_t)NQ && b0 + qs < n_blk) word((b0 + qs) * NP + P0 + qk) = qa;
        if (b0 + 4 < n_blk) {
          qb_ = warp_transpose32(qb_, lane);
          if (qk < (uint32_t)NQ && b0 + 4 + qs < n_blk) word((b0 + 4 + qs) * NP + P0 + qk) = qb_;
        }
        if (MODE == 0) {
          // ---- sequence planes: lane 4 s + k <- plane k of block b0 + s; the candidate plane (k = 2) keeps only the
          // bases of M/=/X runs whose reference position is inside the contig (clipped where it reaches ref_span)
          sn = warp_transpose32(sn, lane);
          const uint32_t blk = b0 + ss;
          if (sk == 2u && blk < n_blk) {
            const int base = (int)(blk * MB_BLOCK);
            uint32_t inrun = 0;
            if (simple) {
              inrun = range_mask((int)qstart - base, (int)(qstart + span) - base);
            } else {
              const uint32_t* __restrict__ cig = in.cigar + co;
              uint32_t qq = 0, rx = 0;
              for (uint32_t ci = 0; qq < lq; ++ci) {
                const uint32_t w = __ldg(cig + ci), op = w & 15u, len = w >> 4;
                if (op == mb::OP_M || op == mb::OP_EQ || op == mb::OP_X) {
                  const uint32_t take = rx < span ? min(len, span - rx) : 0u;
                  inrun |= range_mask((int)qq - base, (int)(qq + take) - base);
                  qq += len; rx += len;
                } else if (op == mb::OP_I || op == mb::OP_S) qq += len;
                else if (op == mb::OP_D || op == mb::OP_N) rx += len;
              }
            }
            sn &= inrun;
          }
          if (sk < 3u && blk < n_blk) word(blk * NP + sk) = sn;
        }
      }
      for (uint32_t w = n_blk * NP + lane; w < units * 8; w += 32) word(w) = 0u;   // padding up to the group's units
    }
  }
}

typedef void (*FillFn)(const Plain, const uint4*, const uint32_t*, const uint32_t*, uint64_t, PackMeta*, uint32_t*, uint64_t,
                       uint64_t);

const char* pack_err_text(uint32_t code, int* status) {
  switch (code) {
    case PE_LQ0: *status = MB_EINVAL; return "l_qseq <= 0 (reference: ZeroDivisionError snps.py:157)";
    case PE_LQBIG: *status = MB_EOVERFLOW; return "l_qseq exceeds the 65535-base short-read limit";
    case PE_SEQOFF: *status = MB_EMALFORMED; return "seq_off span < l_qseq";
    case PE_SEQALIGN: *status = MB_EMALFORMED; return "seq_off must be even with seq_nibbles";
    case PE_NM: *status = MB_EINVAL; return "NM tag missing (reference: KeyError snps.py:148)";
    case PE_NOQUAL: *status = MB_EINVAL; return "qualities missing (reference: np.mean(None) snps.py:151)";
    case PE_NCIG: *status = MB_EOVERFLOW; return "more than 65535 CIGAR ops";
    case PE_OP: *status = MB_EMALFORMED; return "unsupported CIGAR op";
    case PE_QSUM: *status = MB_EMALFORMED; return "CIGAR query length != l_qseq";
    case PE_ALEN0: *status = MB_EINVAL; return "aligned length 0 (reference: ZeroDivisionError snps.py:148)";
    case PE_REFID: *status = MB_EINVAL; return "ref_id outside the layout";
    case PE_POS: *status = MB_EMALFORMED; return "pos outside its contig";
    case PE_SPAN: *status = MB_EOVERFLOW; return "reference span exceeds 65535";
    case PE_CIGOFF: *status = MB_EOVERFLOW; return "batch too large for 32-bit CIGAR offsets";
    default: *status = MB_EINVAL; return "invalid record";
  }
}

// array `slot` of the packed batch: an own allocation (+256 B of zero slack) or a context staging buffer
int pack_alloc(mb_ctx* c, bool staged, int slot, size_t bytes, void** out) {
  if (staged) {
    int rc = mb::ctx_reserve(c, slot, bytes + 256, out);
    if (rc) return rc;
  } else {
    cudaError_t e = cudaMalloc(out, bytes + 256);
    if (e != cudaSuccess) return mb::fail(MB_ENOMEM, "mb_reads_pack: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  }
  MB_CUDA_TRY(cudaMemsetAsync((char*)*out + bytes, 0, 256, c->stream));
  return MB_OK;
}

}  // namespace

namespace mb {

void free_dev_reads_arrays(mb_dev_reads* d) {
  if (!d || !d->owned) return;
  cudaFree(d->hdr); cudaFree(d->vec_off); cudaFree(d->cplx); cudaFree(d->cigar); cudaFree(d->planes); cudaFree(d->long_idx);
  d->hdr = nullptr; d->vec_off = nullptr; d->cplx = nullptr; d->cigar = nullptr; d->planes = nullptr; d->long_idx = nullptr;
}

int upload_plain(mb_ctx* c, const mb_plain_reads* h, mb_plain_reads* d) {
  const uint64_t n = h->n_reads;
  *d = *h;
  if (n == 0) return MB_OK;
  const uint64_t n_cig = h->cigar_off[n], n_el = h->seq_off[n];
  const size_t seq_bytes = h->seq_nibbles ? (size_t)((n_el + 1) / 2) : (size_t)n_el;
  struct Part { const void** dst; const void* src; size_t bytes; } parts[10] = {
      {(const void**)&d->ref_id, h->ref_id, 4 * (size_t)n},       {(const void**)&d->pos, h->pos, 4 * (size_t)n},
      {(const void**)&d->mapq, h->mapq, (size_t)n},               {(const void**)&d->nm, h->nm, 4 * (size_t)n},
      {(const void**)&d->l_qseq, h->l_qseq, 4 * (size_t)n},       {(const void**)&d->cigar_off, h->cigar_off, 8 * (size_t)(n + 1)},
      {(const void**)&d->cigar, h->cigar, 4 * (size_t)n_cig},     {(const void**)&d->seq_off, h->seq_off, 8 * (size_t)(n + 1)},
      {(const void**)&d->seq, h->seq, seq_bytes},                 {(const void**)&d->qual, h->qual, (size_t)n_el}};
  int slot = 16;
  for (auto& p : parts) {
    void* buf = nullptr;
    int rc = ctx_reserve(c, slot++, p.bytes + 256, &buf);
    if (rc) return rc;
    MB_CUDA_TRY(cudaMemsetAsync((char*)buf + p.bytes, 0, 256, c->stream));
    if (p.bytes) MB_CUDA_TRY(cudaMemcpyAsync(buf, p.src, p.bytes, cudaMemcpyHostToDevice, c->stream));
    *p.dst = buf;
  }
  return MB_OK;
}

int pack_on_device(mb_ctx* c, const mb_plain_reads* in, const mb_layout* L, int mode, bool staged, mb_dev_reads* d,
                   const uint64_t* qual_off_dev) {
  memset(d, 0, sizeof(*d));
  d->owned = !staged;
  const uint64_t n = in->n_reads;
  if (n >= 0xFFFFFFFFull) return mb::fail(MB_EOVERFLOW, "mb_reads_pack: more than 2^32-2 reads in one batch");
  if (mode == 0 && L->genes_only) return mb::fail(MB_EINVAL, "mb_reads_pack: snps mode needs a layout with coordinates");
  cudaStream_t st = c->stream;
  PackMeta* meta = (PackMeta*)c->pack_meta;
  PackMeta* hm = (PackMeta*)c->pack_meta_host;
  MB_CUDA_TRY(cudaEventRecord(c->pack_ev[0], st));
  c->pack_timed = true;
  const uint64_t n_groups = (n + MB_GROUP - 1) / MB_GROUP;
  const uint32_t n_sblocks = (uint32_t)((n_groups + SCAN_CHUNK - 1) / SCAN_CHUNK);
  void *src_of = nullptr, *nblk = nullptr, *need = nullptr, *blk_src = nullptr, *blk_dst = nullptr, *g_units = nullptr,
       *bsum = nullptr;
  int rc = ctx_reserve(c, 26, 4 * (size_t)n, &src_of);
  if (!rc) rc = ctx_reserve(c, 27, 2 * (size_t)n, &nblk);
  if (!rc) rc = ctx_reserve(c, 28, (size_t)n, &need);
  if (!rc) rc = ctx_reserve(c, 29, 4 * (size_t)L->n_contigs, &blk_src);
  if (!rc) rc = ctx_reserve(c, 30, 4 * (size_t)L->n_contigs, &blk_dst);
  if (!rc) rc = ctx_reserve(c, 31, 4 * (size_t)n_groups, &g_units);
  if (!rc) rc = ctx_reserve(c, 32, 8 * (size_t)(n_sblocks + 1), &bsum);
  if (!rc) rc = pack_alloc(c, staged, 33, sizeof(mb_read_hdr) * (size_t)n, (void**)&d->hdr);
  if (!rc) rc = pack_alloc(c, staged, 34, 4 * (size_t)(n + 1), (void**)&d->vec_off);
  if (!rc) rc = pack_alloc(c, staged, 35, 4 * (size_t)n, (void**)&d->cplx);
  if (rc) return rc;
  k_pack_meta_init<<<1, 256, 0, st>>>(meta);
  Plain P;
  P.n = n;
  P.ref_id = in->ref_id; P.pos = in->pos; P.nm = in->nm; P.l_qseq = in->l_qseq; P.mapq = in->mapq;
  P.cigar_off = in->cigar_off; P.seq_off = in->seq_off; P.cigar = in->cigar; P.seq = in->seq; P.qual = in->qual;
  P.nib = in->seq_nibbles ? 1u : 0u;
  P.qual_later = 0;
  P.qual_off = qual_off_dev ? qual_off_dev : P.seq_off;
  mb_dev_layout dl{};
  if (mode == 0) { rc = get_dev_layout(c, L, &dl); if (rc) return rc; }
  const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 8);
  const uint32_t wgrid = (uint32_t)std::min<uint64_t>((n + 7) / 8, (uint64_t)c->sm_count * 8);
  if (n) {
    if (mode == 0) {
      k_pack_blocks<<<1, 1024, 0, st>>>(P.ref_id, n, L->n_contigs, dl.order, (uint32_t*)blk_src, (uint32_t*)blk_dst);
      k_pack_analyse<0><<<grid, 256, 0, st>>>(P, L->n_contigs, dl.contig_len, dl.contig_goff, (uint32_t*)blk_src,
                                              (uint32_t*)blk_dst, (uint4*)d->hdr, d->cplx, (uint32_t*)src_of,
                                              (uint16_t*)nblk, (uint8_t*)need, meta);
      c->launches += 2;
    } else {
      k_pack_analyse<1><<<grid, 256, 0, st>>>(P, L->n_contigs, nullptr, nullptr, nullptr, nullptr, (uint4*)d->hdr, d->cplx,
                                              (uint32_t*)src_of, (uint16_t*)nblk, (uint8_t*)need, meta);
      c->launches += 1;
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    if (n) {
      if (pass == 0 && !qual_off_dev) {
        k_pack_qor_flat<<<(uint32_t)c->sm_count * 8, 256, 0, st>>>(P, meta, 0, n);
      } else {   // Phred > 63 somewhere (or garbage in padding): the exact per-record pass decides
        MB_CUDA_TRY(cudaMemsetAsync(&meta->q_or, 0, sizeof(uint32_t), st));
        MB_CUDA_TRY(cudaMemsetAsync(&meta->max_units, 0, sizeof(uint32_t), st));
        k_pack_qor<<<wgrid, 256, 0, st>>>(P, meta);
      }
      c->launches += 1;
    }
    k_pack_finish<<<1, 32, 0, st>>>(meta, n, mode, P.cigar_off);
    c->launches += 1;
    if (n) {
      k_pack_units<<<(uint32_t)std::min<uint64_t>((n_groups + 7) / 8, (uint64_t)c->sm_count * 8), 256, 0, st>>>(
          (const uint16_t*)nblk, n, n_groups, meta, (uint32_t*)g_units);
      k_pack_scan1<<<n_sblocks, SCAN_T, 0, st>>>((const uint32_t*)g_units, n_groups, (unsigned long long*)bsum);
      k_pack_scan2<<<1, SCAN_T, 0, st>>>((unsigned long long*)bsum, n_sblocks, meta);
      c->launches += 3;
    }
    MB_CUDA_TRY(cudaMemcpyAsync(hm, meta, sizeof(PackMeta), cudaMemcpyDeviceToHost, st));
    MB_CUDA_TRY(cudaStreamSynchronize(st));
    MB_CUDA_TRY(cudaGetLastError());
    if (hm->err != ~0ull || !(hm->q_or & 0xC0u) || n == 0) break;
  }
  if (hm->err != ~0ull) {
    int status = MB_EINVAL;
    const char* text = pack_err_text((uint32_t)(hm->err & 0xffu), &status);
    return mb::fail(status, "read %llu: %s", (unsigned long long)(hm->err >> 8), text);
  }
  if (hm->q_or & 0x80u) return mb::fail(MB_EINVAL, "mb_reads_pack: a base quality above 127 is not representable");
  if (hm->unsorted)
    return mb::fail(MB_EUNSORTED, "mb_reads_pack: snps records are not in BAM coordinate order (ref_id, pos); "
                                  "use mb_pack_reads with sort");
  if (hm->n_vec > 0xFFFFFFF0ull) return mb::fail(MB_EOVERFLOW, "mb_reads_pack: batch too large for 32-bit unit offsets");
  d->n_reads = n;
  d->n_vec = hm->n_vec;
  d->total_bases = hm->total_bases;
  d->max_ref_span = hm->max_span;
  d->sorted = mode == 0 ? 1u : 0u;
  d->row_words = hm->row_words;
  d->n_long = hm->n_long;
  d->q_planes = hm->q_planes;
  d->n_planes = hm->n_planes;
  d->max_units = hm->max_units;
  rc = pack_alloc(c, staged, 36, (size_t)d->n_vec * 32, (void**)&d->planes);
  if (!rc) rc = pack_alloc(c, staged, 37, 4 * (size_t)d->n_long, (void**)&d->long_idx);
  if (rc) return rc;
  // the raw CIGAR words are part of the packed batch (cplx[] indexes them): a resident batch owns a copy
  const uint64_t n_cig = hm->n_cig;
  d->n_cigar_words = n_cig;
  if (staged) {
    d->cigar = const_cast<uint32_t*>(in->cigar);
  } else {
    rc = pack_alloc(c, false, 38, 4 * (size_t)n_cig, (void**)&d->cigar);
    if (rc) return rc;
    if (n_cig) MB_CUDA_TRY(cudaMemcpyAsync(d->cigar, in->cigar, 4 * (size_t)n_cig, cudaMemcpyDeviceToDevice, st));
  }
  if (n) {
    k_pack_scan3<<<n_sblocks, SCAN_T, 0, st>>>((const uint32_t*)g_units, n_groups, (const unsigned long long*)bsum, n, meta,
                                               d->vec_off);
    c->launches += 1;
    if (d->n_long) {
      k_pack_long<<<grid, 256, 0, st>>>((const uint8_t*)need, n, meta, (uint4*)d->hdr, d->long_idx);
      c->launches += 1;
    }
    {
      const uint32_t fgrid = (uint32_t)std::min<uint64_t>((n_groups + 7) / 8, (uint64_t)c->sm_count * 8);
      FillFn fill = mode == 0 ? (d->q_planes == 6 ? k_pack_fill<0, 6> : k_pack_fill<0, 7>)
                              : (d->q_planes == 6 ? k_pack_fill<1, 6> : k_pack_fill<1, 7>);
      fill<<<fgrid, 256, 0, st>>>(P, (const uint4*)d->hdr, (const uint32_t*)src_of, d->vec_off, n, meta, d->planes, 0,
                                  n_groups);
    }
    c->launches += 1;
  } else {
    MB_CUDA_TRY(cudaMemsetAsync(d->vec_off, 0, 4, st));
  }
  MB_CUDA_TRY(cudaEventRecord(c->pack_ev[1], st));
  MB_CUDA_TRY(cudaGetLastError());
  d->bytes = sizeof(mb_read_hdr) * n + 4 * (n + 1) + 4 * n + 4 * n_cig + (uint64_t)d->n_vec * 32 + 4 * (uint64_t)d->n_long;
  return MB_OK;
}

}  // namespace mb

namespace mb {

// The host-buffer path of a large batch: the small per-read arrays go first, the analysis and the unit scan run on
// them while seq / qual cross PCIe on a second stream, a chunk of reads at a time; every chunk is turned into
// planes as soon as it has landed. Six Phred planes are assumed (Phred <= 63); the OR of the quality bytes is
// collected on the way and, in the rare batch that needs seven, the caller re-packs (pack_overlapped_verdict).
int pack_from_host_overlapped(mb_ctx* c, const mb_plain_reads* h, const mb_layout* L, int mode, mb_plain_reads* dev,
                              mb_dev_reads* d, bool* done) {
  *done = false;
  const uint64_t n = h->n_reads;
  uint64_t min_reads = 1u << 18;
  if (const char* e = getenv("MB_OVERLAP_MIN_READS")) min_reads = (uint64_t)atoll(e);
  if (getenv("MB_NO_OVERLAP") || n < min_reads || n < 64 || n >= 0xFFFFFFFFull) return MB_OK;
  if (mode == 0 && (L->genes_only || !L->identity_order)) return MB_OK;   // contigs would move: sources are not chunk-local
  {  // only page-locked sources are copied asynchronously; pageable ones would serialise on the host anyway
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h->qual) != cudaSuccess || at.type != cudaMemoryTypeHost) { cudaGetLastError(); return MB_OK; }
    if (cudaPointerGetAttributes(&at, h->seq) != cudaSuccess || at.type != cudaMemoryTypeHost) { cudaGetLastError(); return MB_OK; }
  }
  memset(d, 0, sizeof(*d));
  d->owned = false;
  // all transfers stay on the context stream (one queue, one copy engine, the order the data is needed in); the
  // re-layout kernels run on the second stream behind events and rejoin the context stream at the end
  const bool swap = getenv("MB_OVERLAP_KERNELS_ON_MAIN") != nullptr;     // experiment: the other assignment
  cudaStream_t cs = swap ? c->copy_stream : c->stream;                   // transfers
  cudaStream_t st = swap ? c->stream : c->copy_stream;                   // kernels
  PackMeta* meta = (PackMeta*)c->pack_meta;
  PackMeta* hm = (PackMeta*)c->pack_meta_host;
  MB_CUDA_TRY(cudaEventRecord(c->pack_ev[0], c->stream));
  c->pack_timed = true;
  *dev = *h;
  const uint64_t n_cig = h->cigar_off[n], n_el = h->seq_off[n];
  const size_t seq_bytes = h->seq_nibbles ? (size_t)((n_el + 1) / 2) : (size_t)n_el;
  struct Part { const void** dst; const void* src; size_t bytes; bool bulk; } parts[10] = {
      {(const void**)&dev->ref_id, h->ref_id, 4 * (size_t)n, false},       {(const void**)&dev->pos, h->pos, 4 * (size_t)n, false},
      {(const void**)&dev->mapq, h->mapq, (size_t)n, false},               {(const void**)&dev->nm, h->nm, 4 * (size_t)n, false},
      {(const void**)&dev->l_qseq, h->l_qseq, 4 * (size_t)n, false},       {(const void**)&dev->cigar_off, h->cigar_off, 8 * (size_t)(n + 1), false},
      {(const void**)&dev->cigar, h->cigar, 4 * (size_t)n_cig, false},     {(const void**)&dev->seq_off, h->seq_off, 8 * (size_t)(n + 1), false},
      {(const void**)&dev->seq, h->seq, seq_bytes, true},                  {(const void**)&dev->qual, h->qual, (size_t)n_el, true}};
  int slot = 16;
  for (auto& p : parts) {
    void* buf = nullptr;
    int rc = ctx_reserve(c, slot++, p.bytes + 256, &buf);
    if (rc) return rc;
    MB_CUDA_TRY(cudaMemsetAsync((char*)buf + p.bytes, 0, 256, cs));
    if (!p.bulk && p.bytes) MB_CUDA_TRY(cudaMemcpyAsync(buf, p.src, p.bytes, cudaMemcpyHostToDevice, cs));
    *p.dst = buf;
  }
  // the bulk arrays: behind the small ones on the wire, chunk after chunk on the copy stream
  const int C = mb_ctx::kChunks;
  const bool trace = getenv("MB_OVERLAP_TRACE") != nullptr;              // timeline of one call on stderr
  cudaEvent_t tr[2 * mb_ctx::kChunks + 3];
  if (trace) { for (auto& e : tr) cudaEventCreate(&e); cudaEventRecord(tr[0], cs); }
  uint64_t rb[mb_ctx::kChunks + 1];
  for (int k = 0; k <= C; ++k) rb[k] = k == C ? n : (n * (uint64_t)k / C) & ~(uint64_t)(MB_GROUP - 1);
  if (trace) cudaEventRecord(tr[1], cs);
  MB_CUDA_TRY(cudaEventRecord(c->chunk_ev[C], cs));                        // the small arrays are on the device
  MB_CUDA_TRY(cudaStreamWaitEvent(st, c->chunk_ev[C], 0));
  for (int k = 0; k < C; ++k) {
    const uint64_t e0 = h->seq_off[rb[k]], e1 = h->seq_off[rb[k + 1]];
    if (e1 > e0) {
      MB_CUDA_TRY(cudaMemcpyAsync((uint8_t*)dev->qual + e0, h->qual + e0, (size_t)(e1 - e0), cudaMemcpyHostToDevice, cs));
      const uint64_t b0 = h->seq_nibbles ? e0 / 2 : e0, b1 = h->seq_nibbles ? (e1 + 1) / 2 : e1;
      MB_CUDA_TRY(cudaMemcpyAsync((uint8_t*)dev->seq + b0, h->seq + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, cs));
    }
    MB_CUDA_TRY(cudaEventRecord(c->chunk_ev[k], cs));
    if (trace) cudaEventRecord(tr[2 + k], cs);
  }
  // analysis + unit scan on the small arrays (six Phred planes assumed)
  const uint64_t n_groups = (n + MB_GROUP - 1) / MB_GROUP;
  const uint32_t n_sblocks = (uint32_t)((n_groups + SCAN_CHUNK - 1) / SCAN_CHUNK);
  void *src_of = nullptr, *nblk = nullptr, *need = nullptr, *blk_src = nullptr, *blk_dst = nullptr, *g_units = nullptr,
       *bsum = nullptr;
  int rc = ctx_reserve(c, 26, 4 * (size_t)n, &src_of);
  if (!rc) rc = ctx_reserve(c, 27, 2 * (size_t)n, &nblk);
  if (!rc) rc = ctx_reserve(c, 28, (size_t)n, &need);
  if (!rc) rc = ctx_reserve(c, 29, 4 * (size_t)std::max<uint32_t>(1, L->n_contigs), &blk_src);
  if (!rc) rc = ctx_reserve(c, 30, 4 * (size_t)std::max<uint32_t>(1, L->n_contigs), &blk_dst);
  if (!rc) rc = ctx_reserve(c, 31, 4 * (size_t)n_groups, &g_units);
  if (!rc) rc = ctx_reserve(c, 32, 8 * (size_t)(n_sblocks + 1), &bsum);
  if (!rc) rc = pack_alloc(c, true, 33, sizeof(mb_read_hdr) * (size_t)n, (void**)&d->hdr);
  if (!rc) rc = pack_alloc(c, true, 34, 4 * (size_t)(n + 1), (void**)&d->vec_off);
  if (!rc) rc = pack_alloc(c, true, 35, 4 * (size_t)n, (void**)&d->cplx);
  if (rc) return rc;
  k_pack_meta_init<<<1, 256, 0, st>>>(meta);
  Plain P;
  P.n = n;
  P.ref_id = dev->ref_id; P.pos = dev->pos; P.nm = dev->nm; P.l_qseq = dev->l_qseq; P.mapq = dev->mapq;
  P.cigar_off = dev->cigar_off; P.seq_off = dev->seq_off; P.cigar = dev->cigar; P.seq = dev->seq; P.qual = dev->qual;
  P.nib = h->seq_nibbles ? 1u : 0u;
  P.qual_later = 1;
  P.qual_off = P.seq_off;
  mb_dev_layout dl{};
  if (mode == 0) { rc = get_dev_layout(c, L, &dl); if (rc) return rc; }
  const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 8);
  if (mode == 0) {
    k_pack_blocks<<<1, 1024, 0, st>>>(P.ref_id, n, L->n_contigs, dl.order, (uint32_t*)blk_src, (uint32_t*)blk_dst);
    k_pack_analyse<0><<<grid, 256, 0, st>>>(P, L->n_contigs, dl.contig_len, dl.contig_goff, (uint32_t*)blk_src,
                                            (uint32_t*)blk_dst, (uint4*)d->hdr, d->cplx, (uint32_t*)src_of, (uint16_t*)nblk,
                                            (uint8_t*)need, meta);
    c->launches += 2;
  } else {
    k_pack_analyse<1><<<grid, 256, 0, st>>>(P, L->n_contigs, nullptr, nullptr, nullptr, nullptr, (uint4*)d->hdr, d->cplx,
                                            (uint32_t*)src_of, (uint16_t*)nblk, (uint8_t*)need, meta);
    c->launches += 1;
  }
  k_pack_finish<<<1, 32, 0, st>>>(meta, n, mode, P.cigar_off);
  k_pack_units<<<(uint32_t)std::min<uint64_t>((n_groups + 7) / 8, (uint64_t)c->sm_count * 8), 256, 0, st>>>(
      (const uint16_t*)nblk, n, n_groups, meta, (uint32_t*)g_units);
  k_pack_scan1<<<n_sblocks, SCAN_T, 0, st>>>((const uint32_t*)g_units, n_groups, (unsigned long long*)bsum);
  k_pack_scan2<<<1, SCAN_T, 0, st>>>((unsigned long long*)bsum, n_sblocks, meta);
  c->launches += 4;
  MB_CUDA_TRY(cudaMemcpyAsync(hm, meta, sizeof(PackMeta), cudaMemcpyDeviceToHost, st));
  MB_CUDA_TRY(cudaStreamSynchronize(st));          // the small arrays have landed and are analysed; the bulk is on its way
  MB_CUDA_TRY(cudaGetLastError());
  auto drain = [&]() { cudaStreamSynchronize(cs); cudaStreamSynchronize(st); };   // never leave work in flight behind an error
  if (hm->err != ~0ull) {
    drain();
    int status = MB_EINVAL;
    const char* text = pack_err_text((uint32_t)(hm->err & 0xffu), &status);
    return mb::fail(status, "read %llu: %s", (unsigned long long)(hm->err >> 8), text);
  }
  if (hm->unsorted) {
    drain();
    return mb::fail(MB_EUNSORTED, "mb_reads_pack: snps records are not in BAM coordinate order (ref_id, pos); "
                                  "use mb_pack_reads with sort");
  }
  if (hm->n_vec > 0xFFFFFFF0ull) { drain(); return mb::fail(MB_EOVERFLOW, "mb_reads_pack: batch too large for 32-bit unit offsets"); }
  d->n_reads = n;
  d->n_vec = hm->n_vec;
  d->total_bases = hm->total_bases;
  d->max_ref_span = hm->max_span;
  d->sorted = mode == 0 ? 1u : 0u;
  d->row_words = hm->row_words;
  d->n_long = hm->n_long;
  d->q_planes = hm->q_planes;           // 6 (assumed)
  d->n_planes = hm->n_planes;
  d->max_units = hm->max_units;
  d->n_cigar_words = hm->n_cig;
  d->cigar = const_cast<uint32_t*>(dev->cigar);
  rc = pack_alloc(c, true, 36, (size_t)d->n_vec * 32, (void**)&d->planes);
  if (!rc) rc = pack_alloc(c, true, 37, 4 * (size_t)d->n_long, (void**)&d->long_idx);
  if (rc) { drain(); return rc; }
  k_pack_scan3<<<n_sblocks, SCAN_T, 0, st>>>((const uint32_t*)g_units, n_groups, (const unsigned long long*)bsum, n, meta,
                                             d->vec_off);
  c->launches += 1;
  if (d->n_long) {
    k_pack_long<<<grid, 256, 0, st>>>((const uint8_t*)need, n, meta, (uint4*)d->hdr, d->long_idx);
    c->launches += 1;
  }
  FillFn fill = mode == 0 ? k_pack_fill<0, 6> : k_pack_fill<1, 6>;
  for (int k = 0; k < C; ++k) {
    MB_CUDA_TRY(cudaStreamWaitEvent(st, c->chunk_ev[k], 0));                 // this chunk's seq / qual have landed
    const uint64_t g_lo = rb[k] / MB_GROUP, g_hi = k == C - 1 ? n_groups : rb[k + 1] / MB_GROUP;
    if (g_hi <= g_lo) continue;
    k_pack_qor_flat<<<(uint32_t)c->sm_count * 2, 256, 0, st>>>(P, meta, rb[k], rb[k + 1]);
    const uint32_t fgrid = (uint32_t)std::min<uint64_t>((g_hi - g_lo + 7) / 8, (uint64_t)c->sm_count * 8);
    fill<<<fgrid, 256, 0, st>>>(P, (const uint4*)d->hdr, (const uint32_t*)src_of, d->vec_off, n, meta, d->planes, g_lo, g_hi);
    c->launches += 2;
    if (trace) cudaEventRecord(tr[2 + C + k], st);
  }
  if (trace) {
    cudaStreamSynchronize(st);
    cudaStreamSynchronize(cs);
    float ms = 0;
    cudaEventElapsedTime(&ms, tr[0], tr[1]);
    fprintf(stderr, "overlap trace: small arrays landed %.2f ms;", ms);
    for (int k = 0; k < C; ++k) { cudaEventElapsedTime(&ms, tr[0], tr[2 + k]); fprintf(stderr, " copy%d %.2f", k, ms); }
    for (int k = 0; k < C; ++k) { cudaEventElapsedTime(&ms, tr[0], tr[2 + C + k]); fprintf(stderr, " fill%d %.2f", k, ms); }
    fprintf(stderr, "\n");
    for (auto& e : tr) cudaEventDestroy(e);
  }
  MB_CUDA_TRY(cudaMemcpyAsync(hm, meta, sizeof(PackMeta), cudaMemcpyDeviceToHost, st));   // q_or / late errors: read after the caller's sync
  if (st != c->stream) {                                                   // the context stream continues behind the kernels
    MB_CUDA_TRY(cudaEventRecord(c->chunk_ev[C], st));
    MB_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->chunk_ev[C], 0));
  }
  MB_CUDA_TRY(cudaEventRecord(c->pack_ev[1], c->stream));
  MB_CUDA_TRY(cudaGetLastError());
  d->bytes = sizeof(mb_read_hdr) * n + 4 * (n + 1) + 4 * n + 4 * d->n_cigar_words + (uint64_t)d->n_vec * 32 + 4 * (uint64_t)d->n_long;
  *done = true;
  return MB_OK;
}

int pack_overlapped_verdict(mb_ctx* c, bool* respec) {
  const PackMeta* hm = (const PackMeta*)c->pack_meta_host;
  *respec = false;
  if (hm->err != ~0ull) {
    int status = MB_EINVAL;
    const char* text = pack_err_text((uint32_t)(hm->err & 0xffu), &status);
    return mb::fail(status, "read %llu: %s", (unsigned long long)(hm->err >> 8), text);
  }
  if (hm->q_or & 0xC0u) *respec = true;   // Phred > 63 (seven planes) or > 127 (error): the exact path decides
  return MB_OK;
}

}  // namespace mb

extern "C" int mb_reads_pack_dev(mb_ctx* c, const mb_plain_reads* plain_dev, const mb_layout* L, int mode,
                                 mb_dev_reads** out) {
  if (!c || !plain_dev || !L || !out) return mb::fail(MB_EINVAL, "mb_reads_pack_dev: null argument");
  if (mode != 0 && mode != 1) return mb::fail(MB_EINVAL, "mb_reads_pack_dev: mode must be 0 (snps) or 1 (genes)");
  MB_CUDA_TRY(cudaSetDevice(c->device));
  mb_dev_reads* d = new mb_dev_reads();
  int rc = mb::pack_on_device(c, plain_dev, L, mode, false, d);
  if (rc) { cudaStreamSynchronize(c->stream); mb::free_dev_reads_arrays(d); delete d; return rc; }
  *out = d;
  return MB_OK;
}

extern "C" int mb_reads_pack(mb_ctx* c, const mb_plain_reads* plain_host, const mb_layout* L, int mode,
                             mb_dev_reads** out) {
  if (!c || !plain_host || !L || !out) return mb::fail(MB_EINVAL, "mb_reads_pack: null argument");
  MB_CUDA_TRY(cudaSetDevice(c->device));
  mb_plain_reads dev;
  int rc = mb::upload_plain(c, plain_host, &dev);
  if (rc) return rc;
  return mb_reads_pack_dev(c, &dev, L, mode, out);
}

// ------------------------------------------------------------------------------------------------------------
// BAM record stream on the device: the inflated bytes as they lie in the file; one thread parses one record
// (SAM/BAM specification v1 section 4.2: refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, l_seq, next_refID,
// next_pos, tlen, read_name, cigar, seq, qual, aux). Nothing is copied except the CIGAR words (to an aligned
// array): seq and qual stay in the stream and the packer reads them where they are.
namespace {

__device__ __forceinline__ uint32_t ld_u16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
__device__ __forceinline__ uint32_t ld_u32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// integer value of the NM aux tag, or -1 when absent / malformed (as host_bam.cpp::find_nm)
__device__ int32_t bam_find_nm(const uint8_t* p, const uint8_t* end) {
  while (p + 3 <= end) {
    const uint8_t t0 = p[0], t1 = p[1], ty = p[2];
    p += 3;
    uint32_t sz = 0;
    switch (ty) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'd': sz = 8; break;
      case 'Z': case 'H': {
        while (p < end && *p) ++p;
        if (p >= end) return -1;
        ++p;
        continue;
      }
      case 'B': {
        if (p + 5 > end) return -1;
        const uint8_t sub = p[0];
        const uint32_t cnt = ld_u32(p + 1);
        const uint32_t es = (sub == 'c' || sub == 'C') ? 1u : (sub == 's' || sub == 'S') ? 2u : 4u;
        const size_t remaining = (size_t)(end - p);
        if ((size_t)cnt > (remaining - 5) / es) return -1;
        p += 5 + (size_t)cnt * es;
        continue;
      }
      default: return -1;
    }
    if (p + sz > end) return -1;
    if (t0 == 'N' && t1 == 'M') {
      switch (ty) {
        case 'c': return (int32_t)(int8_t)p[0];
        case 'C': return (int32_t)p[0];
        case 's': return (int32_t)(int16_t)ld_u16(p);
        case 'S': return (int32_t)ld_u16(p);
        case 'i': return (int32_t)ld_u32(p);
        case 'I': { const uint32_t v = ld_u32(p); return v > 0x7fffffffu ? 0x7fffffff : (int32_t)v; }
        default: return -1;
      }
    }
    p += sz;
  }
  return -1;
}

__global__ void __launch_bounds__(256) k_bam_parse(const uint8_t* __restrict__ data, uint64_t total,
                                                   const uint64_t* __restrict__ rec_off, uint64_t n,
                                                   const int32_t* __restrict__ tid_map, const uint64_t* __restrict__ cigar_off,
                                                   int32_t* ref_id, int32_t* pos, uint8_t* mapq, int32_t* nm, int32_t* l_qseq,
                                                   uint32_t* cigar, uint64_t* seq_off, uint64_t* qual_off) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t ro = rec_off[i];
    const uint8_t* __restrict__ r = data + ro;
    const uint32_t bs = ld_u32(r - 4);
    const uint32_t l_name = r[8], n_cig = ld_u16(r + 12), l_seq = ld_u32(r + 16);
    ref_id[i] = tid_map[(int32_t)ld_u32(r)];
    pos[i] = (int32_t)ld_u32(r + 4);
    mapq[i] = r[9];
    l_qseq[i] = (int32_t)l_seq;
    const uint8_t* __restrict__ cg = r + 32 + l_name;
    uint32_t* __restrict__ dst = cigar + cigar_off[i];
    for (uint32_t k = 0; k < n_cig; ++k) dst[k] = ld_u32(cg + 4 * k);
    const uint64_t so = ro + 32 + l_name + 4ull * n_cig;            // seq: two bases per byte, first base in the high nibble
    seq_off[i] = 2 * so;
    qual_off[i] = so + (l_seq + 1) / 2;
    nm[i] = bam_find_nm(r + 32 + l_name + 4ull * n_cig + (l_seq + 1) / 2 + l_seq, r + bs);
    if (i == n - 1) seq_off[n] = 2 * total;
  }
}

}  // namespace

namespace mb {
// BAM record stream -> packed batch in context staging buffers (H2D of the stream, k_bam_parse, k_pack_*).
int pack_from_bam_stream(mb_ctx* c, const mb_bam_stream* s, const int32_t* tid_map, uint32_t n_targets, const mb_layout* L,
                         int mode, mb_dev_reads* dr, const char* who) {
  const uint32_t n_ref = mb_bam_stream_n_ref(s);
  for (uint32_t t = 0; t < n_ref; ++t)
    if (tid_map[t] >= (int32_t)n_targets) return mb::fail(MB_EINVAL, "%s: tid_map[%u] outside the layout", who, t);
  size_t total = 0;
  const uint8_t* data = mb::bam_stream_data(s, &total);
  cudaStream_t st = c->stream;
  const bool timing = getenv("MB_BAM_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_phase = now();
  auto phase = [&](const char* name) {
    if (timing) { cudaStreamSynchronize(st); const double t = now(); fprintf(stderr, "%s: %-10s %8.1f ms\n", who, name, t - t_phase); t_phase = t; }
  };
  // the record walk (host, sequential: 4 bytes per record decide where the next one starts) runs while the stream
  // itself crosses PCIe
  std::vector<uint64_t> rec_off, cig_off;
  int walk_rc = MB_OK;
  std::string walk_msg;
  double walk_ms = 0;
  std::thread walker([&] {
    const double t0 = now();
    walk_rc = mb::bam_stream_walk(s, tid_map, &rec_off, &cig_off);
    if (walk_rc) walk_msg = mb_last_error();
    walk_ms = now() - t0;
  });
  void* d_data = nullptr;
  const bool resident = c->bam_resident == s;          // mb_bam_stream_open_dev sent it while the host inflated
  int rc = mb::ctx_reserve(c, 24, total + 512, &d_data);
  cudaError_t e = cudaSuccess;
  if (!rc && !resident) {
    c->bam_resident = nullptr;
    e = cudaMemsetAsync((char*)d_data + total, 0, 512, st);
    if (e == cudaSuccess && total) e = cudaMemcpyAsync(d_data, data, total, cudaMemcpyHostToDevice, st);
  }
  phase("h2d stream");
  walker.join();
  if (timing) fprintf(stderr, "%s: walk (thread) %6.1f ms\n", who, walk_ms);
  phase("walk join");
  if (rc) return rc;
  if (e != cudaSuccess) return mb::fail(MB_ECUDA, "%s: %s", who, cudaGetErrorString(e));
  if (walk_rc) return mb::fail(walk_rc, "%s", walk_msg.c_str());
  const uint64_t n = rec_off.size();
  if (n >= 0xFFFFFFFFull) return mb::fail(MB_EOVERFLOW, "%s: more than 2^32-2 records", who);
  const uint64_t n_cig = cig_off.empty() ? 0 : cig_off.back();
  void *d_rec = nullptr, *d_tid = nullptr, *d_qoff = nullptr;
  mb_plain_reads dev;
  memset(&dev, 0, sizeof(dev));
  dev.n_reads = n;
  dev.seq_nibbles = 1;
  struct A { const void** p; int slot; size_t bytes; } arrs[8] = {
      {(const void**)&dev.ref_id, 16, 4 * (size_t)n}, {(const void**)&dev.pos, 17, 4 * (size_t)n},
      {(const void**)&dev.mapq, 18, (size_t)n},       {(const void**)&dev.nm, 19, 4 * (size_t)n},
      {(const void**)&dev.l_qseq, 20, 4 * (size_t)n}, {(const void**)&dev.cigar_off, 21, 8 * (size_t)(n + 1)},
      {(const void**)&dev.cigar, 22, 4 * (size_t)n_cig}, {(const void**)&dev.seq_off, 23, 8 * (size_t)(n + 1)}};
  for (auto& a : arrs) {
    void* buf = nullptr;
    rc = mb::ctx_reserve(c, a.slot, a.bytes + 256, &buf);
    if (rc) return rc;
    MB_CUDA_TRY(cudaMemsetAsync((char*)buf + a.bytes, 0, 256, st));
    *a.p = buf;
  }
  rc = mb::ctx_reserve(c, 40, 8 * (size_t)n + 256, &d_rec);
  if (!rc) rc = mb::ctx_reserve(c, 41, 4 * (size_t)n_ref + 256, &d_tid);
  if (!rc) rc = mb::ctx_reserve(c, 42, 8 * (size_t)n + 256, &d_qoff);
  if (rc) return rc;
  dev.seq = (const uint8_t*)d_data;
  dev.qual = (const uint8_t*)d_data;
  phase("reserve");
  if (n) {
    MB_CUDA_TRY(cudaMemcpyAsync(d_rec, rec_off.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    MB_CUDA_TRY(cudaMemcpyAsync((void*)dev.cigar_off, cig_off.data(), 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    MB_CUDA_TRY(cudaMemcpyAsync(d_tid, tid_map, 4 * (size_t)n_ref, cudaMemcpyHostToDevice, st));
    phase("h2d index");
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 8);
    k_bam_parse<<<grid, 256, 0, st>>>((const uint8_t*)d_data, (uint64_t)total, (const uint64_t*)d_rec, n, (const int32_t*)d_tid,
                                      dev.cigar_off, (int32_t*)dev.ref_id, (int32_t*)dev.pos, (uint8_t*)dev.mapq,
                                      (int32_t*)dev.nm, (int32_t*)dev.l_qseq, (uint32_t*)dev.cigar, (uint64_t*)dev.seq_off,
                                      (uint64_t*)d_qoff);
    c->launches += 1;
    MB_CUDA_TRY(cudaStreamSynchronize(st));      // the host vectors above are read by the copies: keep them alive until here
  }
  phase("parse");
  rc = mb::pack_on_device(c, &dev, L, mode, /*staged=*/true, dr, (const uint64_t*)d_qoff);
  if (rc) return rc;
  phase("pack");
  return MB_OK;
}
}  // namespace mb

extern "C" int mb_snps_pileup_bam(mb_ctx* c, const mb_bam_stream* s, const int32_t* tid_map, const mb_layout* L,
                                  const mb_filters* f, uint32_t* counts_host, mb_species_stats* stats_host,
                                  uint64_t* n_reads_out, uint64_t* n_bases_out) {
  if (!c || !s || !tid_map || !L || !f || !counts_host || !stats_host)
    return mb::fail(MB_EINVAL, "mb_snps_pileup_bam: null argument");
  MB_CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  mb_dev_reads dr;
  int rc = mb::pack_from_bam_stream(c, s, tid_map, L->n_contigs, L, 0, &dr, "mb_snps_pileup_bam");
  if (rc) return rc;
  if (n_reads_out) *n_reads_out = dr.n_reads;
  if (n_bases_out) *n_bases_out = dr.total_bases;
  void *counts = nullptr, *stats = nullptr;
  const size_t cbytes = sizeof(uint32_t) * 4 * (size_t)L->gpad, sbytes = sizeof(mb_species_stats) * (size_t)L->n_species;
  rc = mb::ctx_reserve(c, 0, cbytes, &counts);
  if (!rc) rc = mb::ctx_reserve(c, 1, sbytes, &stats);
  if (rc) return rc;
  c->rng_for = nullptr;
  rc = mb_snps_pileup_dev(c, &dr, L, f, (uint32_t*)counts, (mb_species_stats*)stats);
  c->rng_for = nullptr;
  if (rc) return rc;
  if (cbytes) {
    cudaError_t e = cudaMemcpyAsync(counts_host, counts, cbytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && sbytes) e = cudaMemcpyAsync(stats_host, stats, sbytes, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return mb::fail(MB_ECUDA, "mb_snps_pileup_bam: %s", cudaGetErrorString(e));
  }
  return mb::check_overflow(c, "mb_snps_pileup_bam");
}

extern "C" int mb_genes_coverage_bam(mb_ctx* c, const mb_bam_stream* s, const int32_t* tid_map, const uint32_t* gene_len_host,
                                     uint32_t n_genes, const mb_filters* f, int exact_order, mb_gene_out* out_host,
                                     uint64_t* n_reads_out) {
  if (!c || !s || !tid_map || !gene_len_host || !f || !out_host)
    return mb::fail(MB_EINVAL, "mb_genes_coverage_bam: null argument");
  MB_CUDA_TRY(cudaSetDevice(c->device));
  mb_layout* L = nullptr;
  int rc = mb_layout_create_genes(n_genes, &L);
  if (rc) return rc;
  mb_dev_reads dr;
  rc = mb::pack_from_bam_stream(c, s, tid_map, n_genes, L, 1, &dr, "mb_genes_coverage_bam");
  mb_layout_destroy(L);
  if (rc) return rc;
  if (n_reads_out) *n_reads_out = dr.n_reads;
  void *gl = nullptr, *out = nullptr;
  rc = mb::ctx_reserve(c, 2, sizeof(uint32_t) * (size_t)n_genes, &gl);
  if (!rc) rc = mb::ctx_reserve(c, 3, sizeof(mb_gene_out) * (size_t)n_genes, &out);
  if (rc) return rc;
  if (n_genes) MB_CUDA_TRY(cudaMemcpyAsync(gl, gene_len_host, sizeof(uint32_t) * (size_t)n_genes, cudaMemcpyHostToDevice, c->stream));
  rc = mb_genes_coverage_dev(c, &dr, (const uint32_t*)gl, n_genes, f, exact_order, (mb_gene_out*)out);
  if (rc) return rc;
  if (n_genes) MB_CUDA_TRY(cudaMemcpyAsync(out_host, out, sizeof(mb_gene_out) * (size_t)n_genes, cudaMemcpyDeviceToHost, c->stream));
  MB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return MB_OK;
}

extern "C" int mb_ctx_last_pack_ms(mb_ctx* c, float* ms) {
  if (!c || !ms) return mb::fail(MB_EINVAL, "mb_ctx_last_pack_ms: null argument");
  if (!c->pack_timed) return mb::fail(MB_EINVAL, "mb_ctx_last_pack_ms: the device packer has not run yet");
  MB_CUDA_TRY(cudaEventSynchronize(c->pack_ev[1]));
  MB_CUDA_TRY(cudaEventElapsedTime(ms, c->pack_ev[0], c->pack_ev[1]));
  return MB_OK;
}

extern "C" int mb_reads_download(mb_ctx* c, const mb_dev_reads* r, mb_packed_reads** out) {
  if (!c || !r || !out) return mb::fail(MB_EINVAL, "mb_reads_download: null argument");
  MB_CUDA_TRY(cudaSetDevice(c->device));
  MB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  mb_packed_owner* o = new mb_packed_owner();
  memset(o, 0, sizeof(*o));
  o->pinned = false;
  const size_t sizes[7] = {sizeof(mb_read_hdr) * (size_t)r->n_reads, 4 * (size_t)(r->n_reads + 1), 4 * (size_t)r->n_reads,
                           4 * (size_t)r->n_cigar_words, (size_t)r->n_vec * 32, 0, 4 * (size_t)r->n_long};
  const void* src[7] = {r->hdr, r->vec_off, r->cplx, r->cigar, r->planes, nullptr, r->long_idx};
  for (int k = 0; k < 7; ++k) {
    void* b = nullptr;
    if (posix_memalign(&b, 256, sizes[k] ? sizes[k] : 16) != 0) { mb_packed_free(&o->pub); return mb::fail(MB_ENOMEM, "mb_reads_download: out of host memory"); }
    o->buf[k] = b;
    if (sizes[k] && src[k]) {
      cudaError_t e = cudaMemcpy(b, src[k], sizes[k], cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { mb_packed_free(&o->pub); return mb::fail(MB_ECUDA, "mb_reads_download: %s", cudaGetErrorString(e)); }
    }
  }
  o->pub.n_reads = r->n_reads;
  o->pub.n_vec = r->n_vec;
  o->pub.n_cigar_words = r->n_cigar_words;
  o->pub.total_bases = r->total_bases;
  o->pub.max_ref_span = r->max_ref_span;
  o->pub.sorted = r->sorted;
  o->pub.row_words = r->row_words;
  o->pub.n_long = r->n_long;
  o->pub.q_planes = r->q_planes;
  o->pub.n_planes = r->n_planes;
  o->pub.max_units = r->max_units;
  o->pub.hdr = (mb_read_hdr*)o->buf[0];
  o->pub.vec_off = (uint32_t*)o->buf[1];
  o->pub.cplx = (uint32_t*)o->buf[2];
  o->pub.cigar = (uint32_t*)o->buf[3];
  o->pub.planes = (uint32_t*)o->buf[4];
  o->pub.long_idx = (uint32_t*)o->buf[6];
  *out = &o->pub;
  return MB_OK;
}
