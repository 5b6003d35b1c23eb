// SNP pileup on sm_100a: per-reference-position A/C/G/T counts + per-species stats.
//
// Replaces, for all species of a sample at once (reference paths):
//   midas/run/snps.py:141-162  keep_read        -> header filters + mean-quality test, fused
//   midas/run/snps.py:194-199  count_coverage   -> CIGAR walk + per-base quality test + count
//   midas/run/snps.py:201-213  per-position loop -> depth / covered_bases / total_depth, fused
//                                                  into the tile write-out (no second pass)
//
// The count is integer and HBM-bound (about 1.4 B per base in, 16 B per position out), so the
// design removes every per-base atomic and every per-base scalar operation that it can:
//
//   * the padded genome is cut into KT-position tiles; reads are sorted by coordinate, so the reads
//     overlapping a tile are one contiguous index range (k_tile_ranges) and a CTA owns a tile;
//   * STAGE: one thread turns one read into a *bit-plane row in reference coordinates*, entirely in
//     registers: its qualities arrive as up to 2W independent 128-bit loads, every 16 of them become
//     a 16-bit "countable" mask (byte-SIMD compare + multiply gather, no per-base branch), the two
//     2-bit sequence planes are taken as they lie in HBM, and each CIGAR run is applied with one
//     funnel shift per plane per 32-position word. The row is stored in shared memory as
//     W x {A,C,G,T} 32-bit masks aligned to the global 32-position grid;
//   * COUNT: a lane owns a 32-position column and adds the rows covering it with carry-save adders
//     into bit-sliced ("vertical") counters held in registers: one LOP3 works on 32 positions at
//     once, ~0.6 instructions per aligned base instead of one shared-memory atomic per base;
//     KSUB lanes share a column (rows strided) and are merged by a shuffle butterfly of bit-sliced
//     adders into shared-memory vertical counters;
//   * EXPAND: the vertical counters are transposed to u32 counts (4 positions per thread, multiply
//     spread), depth / covered_bases / total_depth are reduced on the way, and the four planes are
//     written with coalesced 128-bit stores. Every output element is written exactly once, no
//     pre-zeroing, no global atomics on counters.
//
// Reads that straddle a tile boundary are staged by both tiles (each counts only its own columns);
// the tile containing a read's start owns it for aligned_reads / mapped_reads. Reads that do not
// fit a W-word row (MB_RF_LONG, at most 0.1 % by construction of W) are counted by k_long_reads
// with global atomics and the position stats are then recomputed by k_restat.
#include <cstdio>
#include <cstdlib>
#include <cmath>

#include "common.h"

namespace {

constexpr int KT = 2048;                 // positions per kernel tile (divides MB_TILE)
constexpr int NCOL = KT / 32;            // 32-position columns per tile
constexpr int KSUB = 4;                  // lanes sharing a column
constexpr int NT = NCOL * KSUB;          // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int NLEV = 8;                  // register vertical-counter levels (per-lane count <= 255 per round)
constexpr int XLEV = NLEV + 2;           // levels after merging KSUB = 4 lanes
constexpr int VCL = 16;                  // shared-memory vertical-counter levels (<= 65535 rows per pass)
constexpr int WMAX = MB_ROW_WORDS_MAX;
constexpr uint32_t PASS_ROWS = 65528;    // rows folded into the shared counters before they are expanded
constexpr int LANE_ROWS_MAX = 248;       // rows one lane may take in one round (keeps its count <= 255)
constexpr int FIRST_N = NCOL + WMAX + 1;

static_assert(MB_TILE % KT == 0, "kernel tile must divide the layout tile");
static_assert(KSUB == 4, "the merge butterfly is written for 4 lanes per column");

enum : uint32_t { F_NEEDQ = 1u, F_OVL = 2u, F_OWNER = 4u, F_SIMPLE = 8u };

struct PileParams {
  const uint4* hdr;
  const uint32_t* blk_off;
  const uint32_t* cplx;
  const uint32_t* cigar;
  const uint32_t* seq2;
  const uint4* qual;          // 16 qualities per element
  const uint32_t* tile_rng;   // [2 * n_ktiles]
  const int32_t* tile_species;  // per MB_TILE tile
  uint32_t* counts;
  unsigned long long* stats;
  uint32_t* sched;
  uint64_t gpad;
  uint32_t n_ktiles;
  uint32_t n_blocks;          // 16-base blocks in seq2 / qual
  uint32_t avg_blocks;        // blocks per read, rounded up (prefetch size estimate)
  uint32_t l2_prefetch;       // 1: bulk-prefetch the next round into L2
  uint32_t row_prefetch;      // 1: each thread prefetches its next read's lines into L2
  uint32_t W;                 // row words
  uint32_t rmax;              // rows staged per round
  const int32_t* t_pid;       // [65536] smallest 100*(alen-NM) that passes mapid, per alen
  const int32_t* t_cov;       // [65536] smallest alen that passes aln_cov, per l_qseq
  uint32_t pid_always;        // mapid is -inf or NaN (reads with a saturated NM pass only then)
  uint32_t readq_eff;         // readq clamped to [0,128]
  int mapq;
  uint32_t qadd;              // (0x80 - baseq) * 0x01010101
  uint32_t defer_restat;      // 1: long reads follow, position stats are recomputed afterwards
};

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// carry-save adder on 32 bit-sliced positions: (sum, carry) of three one-bit addends
#define CSA(s, c, a, b, d)                          \
  do {                                              \
    const uint32_t _a = (a), _b = (b), _d = (d);    \
    (s) = _a ^ _b ^ _d;                             \
    (c) = (_a & _b) | ((_a ^ _b) & _d);             \
  } while (0)

// 16 quality bytes (bit 7 = literal A/C/G/T, bits 0-6 = Phred) -> 2 bytes whose low bytes are the
// 8 + 8 "countable" bits (Phred >= baseq and A/C/G/T); the Phred sum is accumulated into qsum.
// Per 4 bases: mask, dot-product-accumulate, add, and, multiply-high (bit gather).
__device__ __forceinline__ void conv16(const uint4 q, const uint32_t qadd, uint32_t& qsum, uint32_t& b01,
                                       uint32_t& b23) {
  constexpr uint32_t MUL0 = (1u << 25) | (1u << 18) | (1u << 11) | (1u << 4);  // bits 7,15,23,31 -> 32..35
  constexpr uint32_t MUL4 = MUL0 << 4;                                         //                 -> 36..39
  const uint32_t m0 = q.x & 0x7f7f7f7fu, m1 = q.y & 0x7f7f7f7fu, m2 = q.z & 0x7f7f7f7fu, m3 = q.w & 0x7f7f7f7fu;
  qsum = __dp4a(m0, 0x01010101u, qsum);
  qsum = __dp4a(m1, 0x01010101u, qsum);
  qsum = __dp4a(m2, 0x01010101u, qsum);
  qsum = __dp4a(m3, 0x01010101u, qsum);
  const uint32_t t0 = (m0 + qadd) & q.x & 0x80808080u;
  const uint32_t t1 = (m1 + qadd) & q.y & 0x80808080u;
  const uint32_t t2 = (m2 + qadd) & q.z & 0x80808080u;
  const uint32_t t3 = (m3 + qadd) & q.w & 0x80808080u;
  const uint32_t n0 = __umulhi(t0, MUL0), n1 = __umulhi(t1, MUL4);
  const uint32_t n2 = __umulhi(t2, MUL0), n3 = __umulhi(t3, MUL4);
  b01 = (n0 & 0x0fu) | (n1 & ~0x0fu);  // low byte: bases 0-7 (higher bytes hold gather by-products)
  b23 = (n2 & 0x0fu) | (n3 & ~0x0fu);  // low byte: bases 8-15
}

// Ask the copy engine to pull [ptr, ptr + bytes) into L2 (no registers, no shared memory, no wait):
// the loads that follow a round later find their lines on chip instead of in HBM.
__device__ __forceinline__ void l2_prefetch(const void* ptr, uint32_t bytes) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(ptr) & ~(uintptr_t)15;
  const uint32_t n = (bytes + 31u) & ~15u;
  if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(n) : "memory");
}

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n) {   // v << n, 0 for n >= 32
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
  return r;
}
__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, uint32_t n) {   // v >> n, 0 for n >= 32
  uint32_t r;
  asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
  return r;
}

__device__ __forceinline__ uint32_t bit_range_mask(int lo, int hi) {  // bits [lo, hi) of a word, 0 <= lo < hi <= 32
  const uint32_t up = 0xffffffffu << lo;
  const uint32_t dn = 0xffffffffu >> (32 - hi);
  return up & dn;
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

// For every kernel tile the contiguous range of reads that can overlap it: a read that fits a W-word
// row and overlaps [ts, ts+KT) starts after ts - 32 W; the pileup kernel skips the rest by header.
__global__ void k_tile_ranges(const uint4* __restrict__ hdr, uint32_t n_reads, uint32_t back, uint32_t n_tiles,
                              uint32_t* __restrict__ rng) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const uint32_t ts = t * KT, te = ts + KT;
  const uint32_t lo_key = ts > back ? ts - back : 0u;
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

template <int WT>
__global__ void __launch_bounds__(NT, 2) k_pileup_v2(const PileParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr uint32_t W = (uint32_t)WT;          // row words (compile time: every per-lane loop below unrolls)
  uint4* s_vc = reinterpret_cast<uint4*>(smem);                       // [VCL][NCOL] x {A,C,G,T}
  uint4* s_rows = s_vc + VCL * NCOL;                                  // [rmax][W] x {A,C,G,T}
  __shared__ uint32_t s_first[FIRST_N];
  __shared__ unsigned long long s_red[4];                             // covered, depth, aligned, mapped
  __shared__ uint32_t s_tile, s_nmax;
  __shared__ uint8_t s_mod[NCOL + 2 * WMAX + 1];                      // x % W for small x

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int col = tid >> 2, ksub = tid & 3;                           // COUNT role

  for (int i = tid; i < VCL * NCOL; i += NT) s_vc[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < NCOL + 2 * WMAX + 1; i += NT) s_mod[i] = (uint8_t)((uint32_t)i % W);

  uint32_t next_tile = 0;                                             // thread 0: the tile after the current one
  if (tid == 0) next_tile = atomicAdd(p.sched, 1u);
  for (;;) {
    __syncthreads();
    if (tid == 0) { s_tile = next_tile; s_nmax = 0; }
    if (tid < 4) s_red[tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.n_ktiles) break;
    if (tid == 0) next_tile = atomicAdd(p.sched, 1u);                 // its latency hides behind this tile's work
    const uint32_t ts = tile * KT;
    const uint32_t r0 = p.tile_rng[2 * tile], r1 = p.tile_rng[2 * tile + 1];
    const uint32_t tcol = ts >> 5, tb = tcol % W;
    const uint32_t slot_c = s_mod[tb + (uint32_t)col];                // where my column's word sits in a row
    const uint32_t loc_base = ts - 32u * WMAX;                        // row coordinates are stored relative to it (wraps for ts = 0: no read starts before 0)

    uint32_t col_rows = 0;      // rows of my column folded since the last expansion
    uint32_t my_aligned = 0, my_mapped = 0;
    uint32_t covered = 0;
    unsigned long long depth_sum = 0;

    // transpose the shared vertical counters to u32 counts and write the tile (or add to it)
    auto expand = [&](const bool accumulate, const bool last) {
      __syncthreads();
      if (ksub == 0 && col_rows) atomicMax(&s_nmax, col_rows);
      __syncthreads();
      const int nl = 32 - __clz(s_nmax);
      const int nlo = nl < 8 ? nl : 8;
      for (int q = tid; q < KT / 4; q += NT) {
        const int c4 = q >> 3, sh = (q & 7) * 4;
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
        for (int l = 0; l < nlo; ++l) {
          const uint4 w = s_vc[l * NCOL + c4];
          lo[0] += ((((w.x >> sh) & 15u) * 0x00204081u) & 0x01010101u) << l;
          lo[1] += ((((w.y >> sh) & 15u) * 0x00204081u) & 0x01010101u) << l;
          lo[2] += ((((w.z >> sh) & 15u) * 0x00204081u) & 0x01010101u) << l;
          lo[3] += ((((w.w >> sh) & 15u) * 0x00204081u) & 0x01010101u) << l;
        }
        for (int l = 8; l < nl; ++l) {
          const uint4 w = s_vc[l * NCOL + c4];
          hi[0] += ((((w.x >> sh) & 15u) * 0x00204081u) & 0x01010101u) << (l - 8);
          hi[1] += ((((w.y >> sh) & 15u) * 0x00204081u) & 0x01010101u) << (l - 8);
          hi[2] += ((((w.z >> sh) & 15u) * 0x00204081u) & 0x01010101u) << (l - 8);
          hi[3] += ((((w.w >> sh) & 15u) * 0x00204081u) & 0x01010101u) << (l - 8);
        }
        uint32_t d[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v;
          v.x = (lo[c] & 0xffu) | ((hi[c] & 0xffu) << 8);
          v.y = ((lo[c] >> 8) & 0xffu) | (((hi[c] >> 8) & 0xffu) << 8);
          v.z = ((lo[c] >> 16) & 0xffu) | (((hi[c] >> 16) & 0xffu) << 8);
          v.w = (lo[c] >> 24) | ((hi[c] >> 24) << 8);
          uint4* dst = reinterpret_cast<uint4*>(p.counts + (size_t)c * p.gpad + ts + 4 * q);
          if (accumulate) {
            const uint4 o = *dst;
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          *dst = v;
          d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
        }
        if (last) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { covered += d[k] > 0; depth_sum += d[k]; }  // snps.py:204,212-213
        }
      }
      __syncthreads();
      for (int i = tid; i < (nl < VCL ? nl + 1 : VCL) * NCOL; i += NT) s_vc[i] = make_uint4(0, 0, 0, 0);
      if (tid == 0) s_nmax = 0;
      col_rows = 0;
    };

    uint32_t pass_rows = 0;
    bool accumulate = false;
    bool pf_valid = false;                                            // thread 0: pf_b0 = first block of the next round
    uint32_t pf_b0 = 0;
    if (tid == 0 && r0 + p.rmax < r1) { pf_valid = true; pf_b0 = __ldg(p.blk_off + r0 + p.rmax); }
    for (uint32_t rb = r0; rb < r1; rb += p.rmax) {
      const int n = (int)min(p.rmax, r1 - rb);
      if (pass_rows + (uint32_t)n > PASS_ROWS) {  // keep the shared counters below 2^VCL
        expand(accumulate, false);
        accumulate = true;
        pass_rows = 0;
      }
      pass_rows += (uint32_t)n;
      __syncthreads();
      if (tid == 0) {   // L2 prefetch (copy engine) of the rows after this round; their first block was fetched a round ago
        const uint32_t pr0 = rb + (uint32_t)n, pr1 = min(r1, pr0 + p.rmax);
        if (pr1 > pr0 && pf_valid && p.l2_prefetch) {
          const uint32_t nrow = pr1 - pr0;
          uint32_t nblk = nrow * p.avg_blocks + 16u;
          nblk = pf_b0 < p.n_blocks ? min(nblk, p.n_blocks - pf_b0) : 0u;
          l2_prefetch(p.hdr + pr0, nrow * 16u);
          l2_prefetch(p.blk_off + pr0, nrow * 4u);
          l2_prefetch(p.cplx + pr0, nrow * 4u);
          l2_prefetch(p.qual + pf_b0, nblk * 16u);
          l2_prefetch(p.seq2 + pf_b0, nblk * 4u);
        }
        const uint32_t nn0 = pr0 + p.rmax;                           // start of the round after the next one
        pf_valid = nn0 < r1;
        if (pf_valid) pf_b0 = __ldg(p.blk_off + nn0);
      }
      // ---- STAGE: one thread per read, everything in registers: header filters (snps.py:148,154,157),
      // row index table, 2W independent 128-bit quality loads (L1 absorbs the stride between lanes),
      // conversion to bit planes, mean-quality filter, CIGAR placement, row store
      uint4 nh = make_uint4(0, 0, 0, 0);                       // header, block offset, run word and predecessor
      uint32_t nb0 = 0, ncw = 0, nprev = 0;                     // of this thread's next read: one step ahead
      if (tid < n) {
        nh = __ldg(&p.hdr[rb + tid]);
        nb0 = __ldg(&p.blk_off[rb + tid]);
        ncw = __ldg(&p.cplx[rb + tid]);
        if (tid > 0) nprev = __ldg(&p.hdr[rb + tid - 1].x);
      }
      for (int i = tid; i < n; i += NT) {
        const uint32_t r = rb + (uint32_t)i;
        const uint4 h = nh;
        const uint32_t b0 = nb0, cw = ncw, prev_loc = nprev;
        if (p.row_prefetch && i + NT < n) {                    // my next read's lines -> L2 while this one is converted
          const char* q = reinterpret_cast<const char*>(p.qual + nb0);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 128));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.seq2 + nb0));
        }
        if (i + NT < n) {
          nh = __ldg(&p.hdr[r + NT]);
          nb0 = __ldg(&p.blk_off[r + NT]);
          ncw = __ldg(&p.cplx[r + NT]);
          nprev = __ldg(&p.hdr[r + NT - 1].x);
        }
        const uint32_t loc = h.x, lq = h.y & 0xffffu, alen = h.y >> 16, qstart = h.z & 0xffffu, nm = h.z >> 16;
        const uint32_t span = h.w & 0xffffu, mq = (h.w >> 16) & 0xffu, fl = h.w >> 24;
        bool needq = false, ovl = false, owner = false;
        if (!(fl & MB_RF_LONG)) {
          owner = loc >= ts;
          ovl = span > 0 && loc + span > ts;
          // exact integer forms of the two double-precision tests (tables built in ctx.cu)
          bool hk = (fl & MB_RF_NMSAT) ? (p.pid_always != 0)
                                       : (100 * ((int)alen - (int)nm) >= __ldg(p.t_pid + alen));   // snps.py:148
          hk = hk && (int)mq >= p.mapq;                                                             // snps.py:154
          hk = hk && (int)alen >= __ldg(p.t_cov + lq);                                              // snps.py:157
          if (owner) ++my_aligned;                                                                  // snps.py:143
          needq = hk && (ovl || owner);
        }
        const int rel = (int)(loc >> 5) - (int)tcol;                      // start column, in [-W, NCOL)
        {  // s_first[x] = first row of this round whose start column + (W-1) >= x
          const int kmax = NCOL + (int)W - 1;
          int xi = rel + (int)W - 1;
          xi = xi < 0 ? 0 : (xi > kmax ? kmax : xi);
          int xlo = 0;
          if (i > 0) {
            int xp = (int)(prev_loc >> 5) - (int)tcol + (int)W - 1;
            xp = xp < 0 ? 0 : (xp > kmax ? kmax : xp);
            xlo = xp + 1;
          }
          for (int x = xlo; x <= xi; ++x) s_first[x] = (uint32_t)i;
          if (i == n - 1)
            for (int x = xi + 1; x <= kmax + 1; ++x) s_first[x] = (uint32_t)n;
        }
        // query bit planes, 32 bases per word: low / high bit of the base code, countable mask
        // (query word w lives at index w + 1; index 0 is "word -1", zero until the planes are slid)
        uint32_t ql[WT + 2], qh[WT + 2], qv[WT + 2];
#pragma unroll
        for (int w = 0; w <= WT + 1; ++w) { ql[w] = 0; qh[w] = 0; qv[w] = 0; }
        bool keep = false;
        if (needq) {
          const uint32_t nblk = (lq + 15u) >> 4;
          const uint4* __restrict__ qp = p.qual + b0;
          const uint32_t* __restrict__ sp = p.seq2 + b0;
          uint32_t qsum = 0;
#pragma unroll
          for (int w0 = 0; w0 < WT; w0 += 4) {              // 8 blocks = 8 independent 128-bit loads per step
            uint4 qq[8];
            uint32_t ss[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t bidx = (uint32_t)(2 * w0 + u);
              qq[u] = make_uint4(0, 0, 0, 0);
              ss[u] = 0;
              if (2 * w0 + u < 2 * WT && bidx < nblk) { qq[u] = __ldg(qp + bidx); ss[u] = __ldg(sp + bidx); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (w0 + u < WT) {
                uint32_t a01, a23, c01, c23;
                conv16(qq[2 * u], p.qadd, qsum, a01, a23);
                conv16(qq[2 * u + 1], p.qadd, qsum, c01, c23);
                qv[w0 + u + 1] = __byte_perm(__byte_perm(a01, a23, 0x0040), __byte_perm(c01, c23, 0x0040), 0x5410);
                ql[w0 + u + 1] = __byte_perm(ss[2 * u], ss[2 * u + 1], 0x5410);
                qh[w0 + u + 1] = __byte_perm(ss[2 * u], ss[2 * u + 1], 0x7632);
              }
            }
          }
          // mean-quality filter (snps.py:151): mean(q) < readq  <=>  sum(q) < readq * l_qseq (exact)
          keep = qsum >= p.readq_eff * lq;
          if (keep && owner) ++my_mapped;                                    // snps.py:161
        }
        // ---- CIGAR: every M/=/X run is one funnel-shifted, masked copy of the query planes
        uint32_t rl[WT], rh[WT], rv[WT];
#pragma unroll
        for (int j = 0; j < WT; ++j) { rl[j] = 0; rh[j] = 0; rv[j] = 0; }
        if (keep && ovl) {
          const int o = (int)(loc & 31u);
          // run (q0, x0, len): query bases [q0, q0+len) lie at row bits [o + x0, o + x0 + len); row word j
          // takes query bits [32 j + dq, +32) with dq = q0 - o - x0 = 32 d + sh. d is -1 for an
          // unclipped read that does not start on a word boundary; any other d first slides the
          // planes by whole words (static register moves) so that the per-word code is branch-free.
          const bool simple = (fl & MB_RF_SIMPLE) != 0;
          const uint32_t* __restrict__ rec = p.cigar;
          uint32_t n_runs = 1;
          int wshift = 0;                                           // words the planes have been slid so far
          const bool runs2 = (fl & MB_RF_RUNS2) != 0;              // both runs are in the cplx word: no record fetch
          if (runs2) n_runs = 2;
          else if (!simple) { rec = p.cigar + cw; n_runs = __ldg(rec); }
          for (uint32_t k = 0; k < n_runs; ++k) {
            int q0 = (int)qstart, x0 = 0, len = (int)span;
            if (runs2) {
              const int len0 = (int)(cw & 0xffffu);
              if (k == 0) len = len0;
              else { x0 = len0 + (int)(cw >> 24); q0 = (int)qstart + len0 + (int)((cw >> 16) & 0xffu); len = (int)span - x0; }
            } else if (!simple) {
              const uint32_t w0 = __ldg(rec + 1 + 2 * k);
              q0 = (int)(w0 & 0xffffu); len = (int)(w0 >> 16); x0 = (int)__ldg(rec + 2 + 2 * k);
            }
            const int p0 = o + x0, dq = q0 - p0;
            int d = (dq >> 5) - wshift;
            int sh = dq & 31;
            if (d == 0 && sh == 0) { d = -1; sh = 32; }   // a read starting on a word boundary: shift by a whole word
            // heavily clipped reads / long indels: slide the planes by whole words (static register
            // moves in a short loop) until the word offset is -1 or 0 again
            while (d > -1) {
#pragma unroll
              for (int w = 0; w <= WT; ++w) { ql[w] = ql[w + 1]; qh[w] = qh[w + 1]; qv[w] = qv[w + 1]; }
              ql[WT + 1] = 0; qh[WT + 1] = 0; qv[WT + 1] = 0;
              --d; ++wshift;
            }
            while (d < -1) {
#pragma unroll
              for (int w = WT + 1; w > 0; --w) { ql[w] = ql[w - 1]; qh[w] = qh[w - 1]; qv[w] = qv[w - 1]; }
              ql[0] = 0; qh[0] = 0; qv[0] = 0;
              ++d; --wshift;
            }
            // now d == -1: row word j = query words (j-1 : j) funnel-shifted by sh
            // mask of the run's bits inside word j = [p0 - 32 j, p0 + len - 32 j) clipped to [0, 32): built
            // branch-free from clamped shifts (a PTX shift by >= 32 gives 0)
#pragma unroll
            for (int j = 0; j < WT; ++j) {
              const int r0b = p0 - 32 * j, r1b = r0b + len;
              const uint32_t up = shl_clamp(0xffffffffu, (uint32_t)max(r0b, 0));
              const uint32_t dn = shr_clamp(0xffffffffu, (uint32_t)(32 - min(max(r1b, 0), 32)));
              const uint32_t msk = up & dn;
              rl[j] |= __funnelshift_rc(ql[j], ql[j + 1], sh) & msk;   // clamped: sh = 32 selects the upper word
              rh[j] |= __funnelshift_rc(qh[j], qh[j + 1], sh) & msk;
              rv[j] |= __funnelshift_rc(qv[j], qv[j + 1], sh) & msk;
            }
          }
        }
        const uint32_t slot0 = s_mod[(int)tb + rel + (int)W];
        uint4* __restrict__ dst = s_rows + (uint32_t)i * W;
#pragma unroll
        for (int j = 0; j < WT; ++j) {
          uint32_t slot = slot0 + (uint32_t)j;
          if (slot >= W) slot -= W;
          dst[slot] = make_uint4(rv[j] & ~rh[j] & ~rl[j], rv[j] & ~rh[j] & rl[j], rv[j] & rh[j] & ~rl[j],
                                 rv[j] & rh[j] & rl[j]);
        }
      }
      __syncthreads();
      // ---- COUNT: my column's rows of this round, KSUB-strided, 8 rows per carry-save tree into
      // bit-sliced register counters; then the KSUB lanes of a column are merged by a shuffle
      // butterfly of bit-sliced adders and lane ksub folds class ksub into the shared counters
      {
        const uint32_t lo = s_first[col], hi = s_first[col + (int)W];
        if (__any_sync(0xffffffffu, hi > lo)) {
          if (ksub == 0) col_rows += hi - lo;
          uint32_t cnt[4][XLEV];
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int l = 0; l < NLEV; ++l) cnt[c][l] = 0;
          const uint4* __restrict__ base = s_rows + slot_c;
          for (uint32_t r = lo + (uint32_t)ksub; r < hi; r += 8 * KSUB) {
            uint4 mrow[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t rr = r + (uint32_t)(u * KSUB);
              mrow[u] = rr < hi ? base[rr * W] : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t in[8];
#pragma unroll
              for (int u = 0; u < 8; ++u)
                in[u] = c == 0 ? mrow[u].x : (c == 1 ? mrow[u].y : (c == 2 ? mrow[u].z : mrow[u].w));
              uint32_t t0, t1, t2, t3, f0, f1, e;
              CSA(cnt[c][0], t0, cnt[c][0], in[0], in[1]);
              CSA(cnt[c][0], t1, cnt[c][0], in[2], in[3]);
              CSA(cnt[c][1], f0, cnt[c][1], t0, t1);
              CSA(cnt[c][0], t2, cnt[c][0], in[4], in[5]);
              CSA(cnt[c][0], t3, cnt[c][0], in[6], in[7]);
              CSA(cnt[c][1], f1, cnt[c][1], t2, t3);
              CSA(cnt[c][2], e, cnt[c][2], f0, f1);
#pragma unroll
              for (int l = 3; l < NLEV; ++l) {
                const uint32_t t = cnt[c][l] & e;
                cnt[c][l] ^= e;
                e = t;
              }
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int step = 0; step < 2; ++step) {
              uint32_t carry = 0;
#pragma unroll
              for (int l = 0; l < NLEV + step; ++l) {
                const uint32_t y = __shfl_xor_sync(0xffffffffu, cnt[c][l], 1 << step);
                CSA(cnt[c][l], carry, cnt[c][l], y, carry);
              }
              cnt[c][NLEV + step] = carry;
            }
          }
          uint32_t* vc = reinterpret_cast<uint32_t*>(s_vc) + col * 4 + ksub;
          uint32_t carry = 0;
#pragma unroll
          for (int l = 0; l < VCL; ++l) {
            if (l >= XLEV && !__any_sync(0xffffffffu, carry != 0)) break;
            const uint32_t mine = l >= XLEV ? 0u : (ksub == 0 ? cnt[0][l < XLEV ? l : 0]
                                  : (ksub == 1 ? cnt[1][l < XLEV ? l : 0]
                                  : (ksub == 2 ? cnt[2][l < XLEV ? l : 0] : cnt[3][l < XLEV ? l : 0])));
            const uint32_t a = vc[l * NCOL * 4];
            uint32_t sum;
            CSA(sum, carry, a, mine, carry);
            vc[l * NCOL * 4] = sum;
          }
        }
      }
    }
    expand(accumulate, true);

    // ---- per-tile stats -> per-species stats
    covered = warp_sum(covered);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) depth_sum += __shfl_xor_sync(0xffffffffu, depth_sum, o);
    my_aligned = warp_sum(my_aligned);
    my_mapped = warp_sum(my_mapped);
    if (lane == 0) {
      if (covered) atomicAdd(&s_red[0], (unsigned long long)covered);
      if (depth_sum) atomicAdd(&s_red[1], depth_sum);
      if (my_aligned) atomicAdd(&s_red[2], (unsigned long long)my_aligned);
      if (my_mapped) atomicAdd(&s_red[3], (unsigned long long)my_mapped);
    }
    __syncthreads();
    if (tid < 4) {
      const int sp = p.tile_species[tile / (MB_TILE / KT)];
      const unsigned long long v = s_red[tid];
      if (sp >= 0 && v && !(p.defer_restat && tid < 2)) atomicAdd(&p.stats[5 * sp + 1 + tid], v);
    }
  }
}

// ---- reads that do not fit a W-word row: one warp per read, global atomics (rare by construction)
struct LongParams {
  const uint4* hdr;
  const uint32_t* blk_off;
  const uint32_t* cplx;
  const uint32_t* cigar;
  const uint32_t* seq2;
  const uint8_t* qual;
  const uint32_t* long_idx;
  const int32_t* tile_species;
  uint32_t* counts;
  unsigned long long* stats;
  uint64_t gpad;
  uint32_t n_long;
  double mapid, aln_cov;
  int readq, mapq;
  uint32_t qthr;  // 0x80 + baseq
};

__global__ void __launch_bounds__(256) k_long_reads(const LongParams p) {
  const int lane = threadIdx.x & 31;
  const uint32_t wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t k = wg; k < p.n_long; k += nw) {
    const uint32_t r = __ldg(p.long_idx + k);
    const uint4 h = __ldg(&p.hdr[r]);
    const uint32_t loc = h.x, lq = h.y & 0xffffu, alen = h.y >> 16, qstart = h.z & 0xffffu, nm = h.z >> 16;
    const uint32_t span = h.w & 0xffffu, mq = (h.w >> 16) & 0xffu, fl = h.w >> 24;
    const int sp = p.tile_species[loc / MB_TILE];
    double pid = __ddiv_rn((double)(100 * ((int)alen - (int)nm)), (double)alen);
    if (fl & MB_RF_NMSAT) pid = -__longlong_as_double(0x7ff0000000000000ll);
    bool keep = !(pid < p.mapid) && !((int)mq < p.mapq) && !(__ddiv_rn((double)alen, (double)lq) < p.aln_cov);
    const uint32_t b0 = __ldg(&p.blk_off[r]);
    const uint8_t* __restrict__ q = p.qual + (size_t)b0 * 16;
    if (keep) {
      uint32_t s = 0;
      for (uint32_t i = lane; i < lq; i += 32) s += __ldg(q + i) & 0x7fu;
      s = warp_sum(s);
      keep = (long long)s >= (long long)p.readq * (long long)lq;
    }
    if (lane == 0 && sp >= 0) {
      atomicAdd(&p.stats[5 * sp + 3], 1ull);
      if (keep) atomicAdd(&p.stats[5 * sp + 4], 1ull);
    }
    if (!keep) continue;
    const bool simple = (fl & MB_RF_SIMPLE) != 0;
    const uint32_t* __restrict__ rec = (simple || (fl & MB_RF_RUNS2)) ? p.cigar : p.cigar + __ldg(p.cplx + r);
    const bool runs2 = (fl & MB_RF_RUNS2) != 0;
    const uint32_t cw = simple ? 0u : __ldg(p.cplx + r);
    const uint32_t n_runs = simple ? 1u : (runs2 ? 2u : __ldg(rec));
    for (uint32_t o = 0; o < n_runs; ++o) {
      uint32_t w0, x0;
      if (simple) { w0 = qstart | (span << 16); x0 = 0; }
      else if (runs2) {
        const uint32_t len0 = cw & 0xffffu, ins = (cw >> 16) & 0xffu, del = cw >> 24;
        if (o == 0) { w0 = qstart | (len0 << 16); x0 = 0; }
        else { x0 = len0 + del; w0 = (qstart + len0 + ins) | ((span - x0) << 16); }
      } else { w0 = __ldg(rec + 1 + 2 * o); x0 = __ldg(rec + 2 + 2 * o); }
      const uint32_t q0 = w0 & 0xffffu, len = w0 >> 16;
      for (uint32_t i = lane; i < len; i += 32) {
        const uint32_t qi = q0 + i;
        if (__ldg(q + qi) >= p.qthr) {
          const uint32_t sw = __ldg(p.seq2 + b0 + (qi >> 4)), bb = qi & 15u;
          const uint32_t cd = ((sw >> bb) & 1u) | (((sw >> (16 + bb)) & 1u) << 1);
          atomicAdd(p.counts + (size_t)cd * p.gpad + loc + x0 + i, 1u);
        }
      }
    }
  }
}

// covered_bases / total_depth from finished counts (snps.py:204,212-213), one block per MB_TILE tile
__global__ void __launch_bounds__(256) k_restat(const uint32_t* __restrict__ counts, uint64_t gpad, uint32_t n_tiles,
                                                const int32_t* __restrict__ tile_species,
                                                unsigned long long* stats) {
  __shared__ unsigned long long s_red[2];
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    if (threadIdx.x < 2) s_red[threadIdx.x] = 0;
    __syncthreads();
    uint32_t covered = 0;
    unsigned long long depth = 0;
    for (uint32_t x4 = threadIdx.x * 4; x4 < MB_TILE; x4 += blockDim.x * 4) {
      const size_t o = (size_t)t * MB_TILE + x4;
      const uint4 a = *reinterpret_cast<const uint4*>(counts + o);
      const uint4 c = *reinterpret_cast<const uint4*>(counts + gpad + o);
      const uint4 g = *reinterpret_cast<const uint4*>(counts + 2 * gpad + o);
      const uint4 tt = *reinterpret_cast<const uint4*>(counts + 3 * gpad + o);
      const uint32_t d0 = a.x + c.x + g.x + tt.x, d1 = a.y + c.y + g.y + tt.y;
      const uint32_t d2 = a.z + c.z + g.z + tt.z, d3 = a.w + c.w + g.w + tt.w;
      covered += (d0 > 0) + (d1 > 0) + (d2 > 0) + (d3 > 0);
      depth += (unsigned long long)d0 + d1 + d2 + d3;
    }
    covered = warp_sum(covered);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) depth += __shfl_xor_sync(0xffffffffu, depth, o);
    if ((threadIdx.x & 31) == 0) {
      if (covered) atomicAdd(&s_red[0], (unsigned long long)covered);
      if (depth) atomicAdd(&s_red[1], depth);
    }
    __syncthreads();
    const int sp = tile_species[t];
    if (threadIdx.x < 2 && sp >= 0 && s_red[threadIdx.x]) atomicAdd(&stats[5 * sp + 1 + threadIdx.x], s_red[threadIdx.x]);
    __syncthreads();
  }
}

// the row widths the tile kernel is instantiated for; a batch runs at the smallest one >= its row_words
struct PileKernel { uint32_t W; void (*fn)(const PileParams); };
PileKernel pick_kernel(uint32_t row_words) {
  static const PileKernel k[] = {{3, k_pileup_v2<3>},   {4, k_pileup_v2<4>},   {5, k_pileup_v2<5>},  {6, k_pileup_v2<6>},
                                 {7, k_pileup_v2<7>},   {8, k_pileup_v2<8>},   {10, k_pileup_v2<10>}, {12, k_pileup_v2<12>},
                                 {16, k_pileup_v2<16>}};
  for (const PileKernel& e : k)
    if (e.W >= row_words) return e;
  return k[sizeof(k) / sizeof(k[0]) - 1];
}

size_t pile_smem_bytes(uint32_t W, uint32_t rmax) {
  return sizeof(uint4) * ((size_t)VCL * NCOL + (size_t)rmax * W);
}

}  // namespace

extern "C" int mb_snps_pileup_dev(mb_ctx* ctx, const mb_dev_reads* reads, const mb_layout* L, const mb_filters* f,
                                  uint32_t* counts_dev, mb_species_stats* stats_dev) {
  if (!ctx || !reads || !L || !f || !counts_dev || !stats_dev)
    return mb::fail(MB_EINVAL, "mb_snps_pileup_dev: null argument");
  if (!reads->sorted) return mb::fail(MB_EUNSORTED, "mb_snps_pileup_dev: reads were not packed in snps (sorted) mode");
  if (f->baseq < 0 || f->baseq > 127) return mb::fail(MB_EINVAL, "mb_snps_pileup_dev: baseq %d outside [0,127]", f->baseq);
  if (reads->row_words < 2 || reads->row_words > (uint32_t)WMAX)
    return mb::fail(MB_EINVAL, "mb_snps_pileup_dev: row_words %u outside [2,%d]", reads->row_words, WMAX);
  MB_CUDA_TRY(cudaSetDevice(ctx->device));
  mb_dev_layout dl;
  int rc = mb::get_dev_layout(ctx, L, &dl);
  if (rc) return rc;
  const uint32_t n_ktiles = L->n_tiles * (MB_TILE / KT);
  if (ctx->tile_rng_cap < (size_t)2 * n_ktiles + 2) {
    if (ctx->tile_rng) { MB_CUDA_TRY(cudaStreamSynchronize(ctx->stream)); MB_CUDA_TRY(cudaFree(ctx->tile_rng)); }
    ctx->tile_rng = nullptr;
    ctx->tile_rng_cap = 0;
    ctx->rng_for = nullptr;
    MB_CUDA_TRY(cudaMalloc(&ctx->tile_rng, sizeof(uint32_t) * ((size_t)2 * n_ktiles + 2)));
    ctx->tile_rng_cap = (size_t)2 * n_ktiles + 2;
  }
  // shared memory per CTA: two CTAs per SM (override with MB_PILE_SMEM_KB for experiments)
  const PileKernel kern = pick_kernel(reads->row_words);
  const uint32_t W = kern.W;                                          // instantiated row width >= row_words
  if (!ctx->pile_smem) {
    size_t kb = 100;
    if (const char* e = getenv("MB_PILE_SMEM_KB")) { long v = atol(e); if (v >= 32 && v <= 220) kb = (size_t)v; }
    ctx->pile_smem = kb * 1024;
  }
  if (!(ctx->pile_attr_mask & (1u << W))) {
    MB_CUDA_TRY(cudaFuncSetAttribute(kern.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->pile_smem));
    int occ = 0;
    MB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern.fn, NT, ctx->pile_smem));
    ctx->pile_occ = occ > 0 ? occ : 1;
    ctx->pile_attr_mask |= 1u << W;
  }
  const size_t fixed = sizeof(uint4) * (size_t)VCL * NCOL;
  uint32_t rmax = (uint32_t)((ctx->pile_smem - fixed) / (sizeof(uint4) * W));
  if (rmax > (uint32_t)LANE_ROWS_MAX * KSUB) rmax = LANE_ROWS_MAX * KSUB;
  if (rmax >= 2u * NT && !getenv("MB_PILE_RAGGED")) rmax = rmax / NT * NT;   // the same number of reads for every thread
  rmax &= ~7u;
  cudaStream_t st = ctx->stream;
  const uint32_t ns = L->n_species;
  k_init_stats<<<ns ? (ns + 255) / 256 : 1, 256, 0, st>>>(
      reinterpret_cast<unsigned long long*>(stats_dev), dl.species_len, ns, ctx->sched);
  ctx->launches++;
  if (n_ktiles == 0) { MB_CUDA_TRY(cudaGetLastError()); return MB_OK; }
  // the tile -> read-range index depends only on the resident batch: built once per (batch, row width)
  if (ctx->rng_for != reads || ctx->rng_W != W || ctx->rng_ntiles != n_ktiles) {
    k_tile_ranges<<<(n_ktiles + 255) / 256, 256, 0, st>>>(reinterpret_cast<const uint4*>(reads->hdr),
                                                          (uint32_t)reads->n_reads, 32u * W, n_ktiles, ctx->tile_rng);
    ctx->launches++;
    ctx->rng_for = reads;
    ctx->rng_W = W;
    ctx->rng_ntiles = n_ktiles;
  }
  PileParams p;
  p.hdr = reinterpret_cast<const uint4*>(reads->hdr);
  p.blk_off = reads->blk_off;
  p.cplx = reads->cplx;
  p.cigar = reads->cigar;
  p.seq2 = reads->seq2;
  p.qual = reinterpret_cast<const uint4*>(reads->qual);
  p.tile_rng = ctx->tile_rng;
  p.tile_species = dl.tile_species;
  p.counts = counts_dev;
  p.stats = reinterpret_cast<unsigned long long*>(stats_dev);
  p.sched = ctx->sched;
  p.gpad = L->gpad;
  p.n_ktiles = n_ktiles;
  p.n_blocks = (uint32_t)reads->n_blocks;
  p.row_prefetch = getenv("MB_PILE_ROWPF") ? 1u : 0u;
  p.l2_prefetch = getenv("MB_PILE_L2PF") ? 1u : 0u;   // measured slower on C2 (1.4x DRAM traffic): off unless asked for
  p.avg_blocks = (uint32_t)((reads->n_blocks + reads->n_reads - 1) / (reads->n_reads ? reads->n_reads : 1));
  p.W = W;
  p.rmax = rmax;
  rc = mb::get_filter_tables(ctx, f, &p.t_pid, &p.t_cov);
  if (rc) return rc;
  p.pid_always = (f->mapid != f->mapid || f->mapid == -HUGE_VAL) ? 1u : 0u;
  p.readq_eff = f->readq <= 0 ? 0u : (f->readq > 128 ? 128u : (uint32_t)f->readq);
  p.mapq = f->mapq;
  p.qadd = (0x80u - (uint32_t)f->baseq) * 0x01010101u;
  p.defer_restat = reads->n_long ? 1u : 0u;
  const uint32_t grid = min(n_ktiles, (uint32_t)(ctx->sm_count * ctx->pile_occ));
  { int trc = mb::timing_begin(ctx); if (trc) return trc; }
  kern.fn<<<grid, NT, pile_smem_bytes(W, rmax), st>>>(p);
  { int trc = mb::timing_end(ctx); if (trc) return trc; }
  ctx->launches++;
  if (reads->n_long) {
    LongParams lp;
    lp.hdr = p.hdr; lp.blk_off = p.blk_off; lp.cplx = p.cplx; lp.cigar = p.cigar; lp.seq2 = p.seq2;
    lp.qual = reads->qual;
    lp.long_idx = reads->long_idx;
    lp.tile_species = dl.tile_species;
    lp.counts = counts_dev;
    lp.stats = p.stats;
    lp.gpad = L->gpad;
    lp.n_long = reads->n_long;
    lp.mapid = f->mapid; lp.aln_cov = f->aln_cov; lp.readq = f->readq; lp.mapq = f->mapq;
    lp.qthr = 0x80u + (uint32_t)f->baseq;
    const uint32_t lgrid = min((reads->n_long + 7u) / 8u, (uint32_t)ctx->sm_count * 8u);
    k_long_reads<<<lgrid, 256, 0, st>>>(lp);
    ctx->launches++;
    k_restat<<<min(L->n_tiles, (uint32_t)ctx->sm_count * 8u), 256, 0, st>>>(counts_dev, L->gpad, L->n_tiles,
                                                                            dl.tile_species, p.stats);
    ctx->launches++;
  }
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
  ctx->rng_for = nullptr;   // `dr` is a fresh batch in recycled staging buffers: rebuild the tile index
  if (!rc) rc = mb_snps_pileup_dev(ctx, d, L, f, (uint32_t*)counts, (mb_species_stats*)stats);
  ctx->rng_for = nullptr;
  if (!rc && cbytes) {
    cudaError_t e = cudaMemcpyAsync(counts_host, counts, cbytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && sbytes) e = cudaMemcpyAsync(stats_host, stats, sbytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = mb::fail(MB_ECUDA, "mb_snps_pileup: %s", cudaGetErrorString(e));
  }
  return rc;
}
