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
//   * the padded genome is cut into work units of at most KT positions (shorter towards the end of the genome, so that
//     the last claims of the persistent grid are small); reads are sorted by coordinate, so the reads overlapping a
//     unit are one contiguous index range (k_tile_ranges) and a CTA owns a unit;
//   * STAGE: one thread turns one read into a *bit-plane row in reference coordinates*, entirely in
//     registers: the read's bit planes (mb_packed_reads: 32 bases per word) arrive as independent
//     256-bit loads (the 32-byte units of 32 consecutive reads are interleaved, so each load of a warp
//     is 1 KB of consecutive memory), issued as soon as the read's header is there and before its filters are
//     evaluated; Phred >= baseq is a bit-sliced comparator (one LOP3 per Phred bit per 32 bases, two per 32 bases in
//     the instance for baseq = 30; no per-base work at all), the Phred sum is one POPC per plane word (and the two top
//     planes alone settle the mean-quality test for all but poor reads), the two sequence planes are
//     taken as they lie in HBM, and each CIGAR run is one funnel shift per plane per 32-position
//     word (the packer has already cleared the candidate bit of every base outside an M/=/X run, so
//     single-run reads need no masks). The row is stored in shared memory as W x {A,C,G,T} 32-bit
//     masks aligned to the global 32-position grid;
//   * COUNT: the 32-position columns a round of rows touches are dealt over all threads of the CTA (4 NS lanes per
//     column: a lane adds one class word of every NS-th row of its column with a 16-row carry-save tree into
//     bit-sliced ("vertical") counters held in registers: one LOP3 works on 32 positions at once); the NS partial
//     counters are merged by a shuffle butterfly of bit-sliced adders into shared-memory vertical counters.
//     (The round-1 form - a fixed lane -> column map, KSUB lanes per column - is kept behind MB_PILE_VAR=8.)
//   * EXPAND: the vertical counters are transposed to u32 counts (4 positions per thread, multiply
//     spread), depth / covered_bases / total_depth are reduced on the way, and the four planes are
//     written with coalesced 128-bit stores. Every output element is written exactly once, no
//     pre-zeroing, no global atomics on counters. The last CTA to finish hands the per-species stats rows to the
//     caller and clears the running rows and the scheduler counters: a call is one launch.
//
// Reads that straddle a tile boundary are staged by both tiles (each counts only its own columns);
// the tile containing a read's start owns it for aligned_reads / mapped_reads. Reads that do not
// fit a W-word row (MB_RF_LONG, at most 0.1 % by construction of W) are counted by k_long_reads
// with global atomics and the position stats are then recomputed by k_restat.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <type_traits>

#include "common.h"

namespace {

// The positions per kernel tile (a divisor of MB_TILE) are a template choice of the kernel: 2048 (256
// threads, 2 CTAs per SM; the one in use) or 1024 (128 threads, 4 CTAs per SM: measured 3 % slower).
constexpr int KSUB = 4;                  // lanes sharing a column
constexpr int NLEV = 8;                  // register vertical-counter levels (per-lane count <= 255 per round)
constexpr int XLEV = NLEV + 2;           // levels after merging KSUB = 4 lanes
constexpr int VCL = 15;                  // shared-memory vertical-counter levels (<= 32767 rows per pass)
constexpr int WMAX = MB_ROW_WORDS_MAX;
constexpr uint32_t PASS_ROWS = 32760;    // rows folded into the shared counters before they are expanded
constexpr int LANE_ROWS_MAX = 248;       // rows one lane may take in one round (keeps its count <= 255)

static_assert(MB_TILE % 2048 == 0, "kernel tile must divide the layout tile");
static_assert(KSUB == 4, "the merge butterfly is written for 4 lanes per column");

struct PileParams {
  const uint4* hdr;
  const uint32_t* vec_off;
  const uint32_t* cplx;
  const uint32_t* cigar;
  const uint32_t* planes;     // bit-planar read data (mb_packed_reads), regions in 32-byte units
  const uint32_t* zero_unit;  // 32 bytes of zeros (the slack behind planes[]): what a load past a read's end fetches
  const uint32_t* tile_rng;   // [2 * n_ktiles]
  const uint2* units;         // [n_ktiles] (first position, positions) of the work units ("tiles" of at most KT positions)
  const int32_t* tile_species;  // per MB_TILE tile
  uint32_t* counts;
  unsigned long long* stats;
  uint32_t* sched;
  uint64_t gpad;
  uint32_t n_ktiles;
  uint32_t W;                 // row words
  uint32_t rmax;              // rows staged per round
  const int32_t* t_pid;       // [65536] smallest 100*(alen-NM) that passes mapid, per alen
  const int32_t* t_cov;       // [65536] smallest alen that passes aln_cov, per l_qseq
  uint32_t pid_always;        // mapid is -inf or NaN (reads with a saturated NM pass only then)
  uint32_t readq_eff;         // readq clamped to [0,128]
  int mapq;
  uint32_t baseq;             // per-base threshold (0: every candidate base counts)
  uint32_t defer_restat;      // 1: long reads follow, position stats are recomputed afterwards
  uint32_t slot_units;        // TMA staging: 32-byte units per read a warp's shared-memory slot holds
  uint32_t n_reads;
  uint32_t* ovf;              // set to 1 when a u32 counter saturates (MB_EOVERFLOW)
  unsigned long long* mirror; // peer-memory copy of the stats rows (multi-GPU), written by the last CTA; or null
  unsigned long long* acc;    // [5 * n_species] the launch's running stats (zero before and after the launch)
  const uint64_t* species_len;
  uint32_t n_species;
};

// Phase clocks of the tile kernel (a separate build, `make prof`: -DMB_PROF): per CTA and warp, cycles spent in
// STAGE / COUNT / EXPAND and at the barriers behind them, read back by mb_debug_prof (scripts/phase_prof.py)
#ifdef MB_PROF
constexpr int PROF_K = 8;                // 0 other, 1 stage, 2 wait after stage, 3 count, 4 wait after count, 5 expand, 6 start ns, 7 end ns
constexpr int PROF_CTAS = 1024, PROF_WARPS = 8;
__device__ unsigned long long g_prof[PROF_CTAS * PROF_WARPS * PROF_K];
#define PROF_DECL __shared__ uint32_t s_prof[PROF_WARPS][PROF_K]; __shared__ uint32_t s_prof_t[PROF_WARPS];
#define PROF_INIT do { if ((threadIdx.x & 31) == 0) { for (int k_ = 0; k_ < PROF_K; ++k_) s_prof[threadIdx.x >> 5][k_] = 0; \
    s_prof_t[threadIdx.x >> 5] = (uint32_t)clock(); unsigned long long ns_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_)); \
    if (blockIdx.x < PROF_CTAS) g_prof[(blockIdx.x * PROF_WARPS + (threadIdx.x >> 5)) * PROF_K + 6] = ns_; } } while (0)
#define PROF_MARK(k) do { if ((threadIdx.x & 31) == 0) { const uint32_t t_ = (uint32_t)clock(); \
    s_prof[threadIdx.x >> 5][k] += t_ - s_prof_t[threadIdx.x >> 5]; s_prof_t[threadIdx.x >> 5] = t_; } } while (0)
#define PROF_FINI do { if ((threadIdx.x & 31) == 0 && blockIdx.x < PROF_CTAS) { unsigned long long ns_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_)); \
    unsigned long long* o_ = g_prof + (blockIdx.x * PROF_WARPS + (threadIdx.x >> 5)) * PROF_K; \
    for (int k_ = 0; k_ < 6; ++k_) o_[k_] = s_prof[threadIdx.x >> 5][k_]; o_[7] = ns_; } } while (0)
#else
#define PROF_DECL
#define PROF_INIT do {} while (0)
#define PROF_MARK(k) do {} while (0)
#define PROF_FINI do {} while (0)
#endif

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

// one 256-bit load: a whole 32-byte sector per lane (the regions of the packed reads are 32-byte aligned)
__device__ __forceinline__ void ldg256(const uint32_t* __restrict__ p, uint32_t* r) {
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "l"(p));
}

// ---- asynchronous staging (VAR & 4): one bulk copy (TMA, UBLKCP) per warp and group, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MB_DONE_%=;\n"
      "bra MB_WAIT_%=;\n"
      "MB_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void lds256(const uint32_t* p, uint32_t* r) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 4);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
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

__global__ void k_init_stats(unsigned long long* stats, const uint64_t* species_len, uint32_t n_species,
                             uint32_t* sched) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) { sched[0] = 0; sched[1] = 0; }
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
// One warp per tile, 32-ary search: every level is one coalesced-ish probe per lane and a ballot (5 dependent
// levels for 5 M reads instead of the 22 of a binary search).
__device__ __forceinline__ uint32_t warp_lower_bound(const uint4* __restrict__ hdr, uint32_t lo, uint32_t hi, uint32_t key,
                                                     const uint32_t lane) {
  // first index in [lo, hi) whose loc is >= key
  while (hi - lo > 32u) {
    const uint32_t step = (hi - lo + 31u) / 32u;
    const uint32_t probe = lo + (lane + 1u) * step - 1u;             // last element of my sub-range
    const bool below = probe < hi ? __ldg(&hdr[probe].x) < key : false;
    const uint32_t k = (uint32_t)__popc(__ballot_sync(0xffffffffu, below));   // sub-ranges entirely below the key
    lo += k * step;
    hi = min(hi, lo + step);
  }
  const bool below = lo + lane < hi ? __ldg(&hdr[lo + lane].x) < key : false;
  return lo + (uint32_t)__popc(__ballot_sync(0xffffffffu, below));
}

__global__ void __launch_bounds__(256) k_tile_ranges(const uint4* __restrict__ hdr, uint32_t n_reads, uint32_t back,
                                                     uint32_t n_tiles, const uint2* __restrict__ units,
                                                     uint32_t* __restrict__ rng) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= n_tiles) return;
  const uint2 un = units[t];
  const uint32_t ts = un.x, te = ts + un.y;
  const uint32_t r0 = warp_lower_bound(hdr, 0u, n_reads, ts > back ? ts - back : 0u, lane);
  const uint32_t r1 = warp_lower_bound(hdr, r0, n_reads, te, lane);
  if (lane == 0) { rng[2 * t] = r0; rng[2 * t + 1] = r1; }
}

// slide three register arrays of WT+2 words until the word offset d is -1 (static register moves in a
// short loop; taken only by heavily clipped reads and by indels that push a run across a word boundary)
#define MB_SLIDE(d, wshift, A, B, C)                                                        \
  do {                                                                                      \
    if ((d) == -1) break;                                                                   \
    while ((d) > -1) {                                                                      \
      _Pragma("unroll") for (int w = 0; w <= WT; ++w) { A[w] = A[w + 1]; B[w] = B[w + 1]; C[w] = C[w + 1]; } \
      A[WT + 1] = 0; B[WT + 1] = 0; C[WT + 1] = 0;                                          \
      --(d); ++(wshift);                                                                    \
    }                                                                                       \
    while ((d) < -1) {                                                                      \
      _Pragma("unroll") for (int w = WT + 1; w > 0; --w) { A[w] = A[w - 1]; B[w] = B[w - 1]; C[w] = C[w - 1]; } \
      A[0] = 0; B[0] = 0; C[0] = 0;                                                         \
      ++(d); --(wshift);                                                                    \
    }                                                                                       \
  } while (0)

// VAR: bit 0: 128-bit plane loads in steps of four blocks instead of 256-bit loads (experiment switch);
//      bit 1: 1024-position tiles, 128 threads, 4 CTAs per SM instead of 2048 / 256 / 2
template <int WT, int NQ, int VAR>
__global__ void __launch_bounds__((VAR & 2) ? 128 : 256, (VAR & 2) ? 4 : ((VAR & 128) ? 3 : 2)) k_pileup_v2(const PileParams p) {
  constexpr int KT = (VAR & 2) ? 1024 : 2048;
  constexpr int NCOL = KT / 32;            // 32-position columns per tile
  constexpr int NT = NCOL * KSUB;          // threads per CTA
  constexpr int FIRST_N = NCOL + WMAX + 1;
  constexpr bool TMA = (VAR & 4) != 0;          // bit 2: the planes arrive by bulk copy into a per-warp slot
  constexpr bool PFL2 = (VAR & 8) != 0;         // bit 3: the next round's arrays are prefetched into L2 during COUNT
  constexpr bool SORTED = (VAR & 16) != 0;      // bit 4: the rows of a round are dealt to the threads by class (below)
  constexpr int KMAX = 3;                       // SORTED: rows per thread and round (rmax <= KMAX * NT)
  constexpr bool DEFER = (VAR & 32) != 0;       // bit 5: reads with an indel are placed after the warp's simple ones (below)
  constexpr int DCAP = 16, DWORDS = 3 * WT + 6; // DEFER: entries per warp and round; words per entry
  constexpr bool BAL = (VAR & 64) != 0;         // bit 6: COUNT deals the round's active columns over all threads (below)
  constexpr bool BQ30 = (VAR & 512) != 0;       // bit 9: instance for baseq = 30, the CLI default (comparator below)
  // bit 12: no read of the batch has more than WT - 1 blocks of 32 bases (150-base reads: 5 blocks, while a row needs
  // 6 words for any alignment of the start): one block less to load, convert and carry in registers
  constexpr int QB = ((VAR & 4096) != 0 && WT <= 8 && WT > 2) ? WT - 1 : WT;
  // bit 10: a row's plane loads are issued as soon as its header is there, before the header filters are evaluated
  constexpr bool SPEC = (VAR & 1024) != 0 && WT <= 8 && !(VAR & 1) && !TMA && !SORTED;
  // the next unit's table entries and my first header of it are requested during the current unit's last round
  constexpr bool XU = PFL2 && !TMA && !SORTED;
  constexpr int NLB = 10;                       // BAL: register counter levels of a lane (a round has < 1024 rows)
  extern __shared__ __align__(128) uint32_t smem[];
  constexpr uint32_t W = (uint32_t)WT;          // row words (compile time: every per-lane loop below unrolls)
  constexpr int NP = 3 + NQ;                    // words per 32-base block: seq lo, seq hi, candidate, Phred bits
  uint32_t tmk[NQ];                             // bit k of baseq, spread over a word (warp-uniform)
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    if (BQ30) { tmk[k] = 0; continue; }
    tmk[k] = ((p.baseq >> k) & 1u) ? 0xffffffffu : 0u;
    // kept as a plain register operand: seen as a known-uniform value the compiler turns every comparator
    // step into an AND plus a predicated OR (two instructions) instead of one three-input LOP3
    asm volatile("mov.b32 %0, %0;" : "+r"(tmk[k]));
  }
  const uint32_t q_allow = (p.baseq >> NQ) ? 0u : 0xffffffffu;   // a threshold above every representable Phred
  uint4* s_vc = reinterpret_cast<uint4*>(smem);                       // [VCL][NCOL] x {A,C,G,T}
  uint4* s_rows = s_vc + VCL * NCOL;                                  // [rmax][W] x {A,C,G,T}
  __shared__ uint32_t s_first[FIRST_N];
  __shared__ unsigned long long s_red[4];                             // covered, depth, aligned, mapped
  __shared__ uint32_t s_tile, s_nmax;
  __shared__ uint8_t s_mod[NCOL + 2 * WMAX + 1];                      // x % W for small x
  __shared__ __align__(8) unsigned long long s_mbar[NT / 32];         // TMA: one per warp
  __shared__ uint32_t s_pf[4];                                         // PFL2: first read, last read, first unit, units
  __shared__ uint32_t s_nx[5];                                         // XU: the next unit (start, positions, reads r0, r1, its index)
  __shared__ uint32_t s_cr[2];                                         // BAL: first / past-the-last column the round's rows touch
  __shared__ uint32_t s_colrows[NCOL];                                 // BAL: rows folded into a column since the last expansion
  // the two filter tables for the lengths a row of W words can hold (a read that is staged has l_qseq <= 32 W)
  constexpr int NTAB = 32 * WT + 1;
  __shared__ int32_t s_tpid[NTAB], s_tcov[NTAB];
  for (int i = threadIdx.x; i < NTAB; i += NT) { s_tpid[i] = p.t_pid[i]; s_tcov[i] = p.t_cov[i]; }
  // TMA: slot of warp w = slot_units x {32 lanes x 32 bytes}, as the group lies in HBM; then one all-zero unit
  uint32_t* const s_slots = reinterpret_cast<uint32_t*>(s_rows + ((size_t)p.rmax + 1) * W);
  uint32_t* const my_slot = s_slots + (size_t)(threadIdx.x >> 5) * p.slot_units * (8 * MB_GROUP);
  const uint32_t* const s_zero = s_slots + (size_t)(NT / 32) * p.slot_units * (8 * MB_GROUP);
  uint32_t mb_phase = 0;                                               // parity of my warp's next wait
  // SORTED: the round's headers (bit 31 of .w = the read passed the header filters and touches the tile), the
  // rows in processing order, the unit offsets of the round's groups
  // DEFER: per warp, the converted planes + header of the reads whose placement is put off: the indel code then runs
  // once per warp and round over those reads only, instead of in every iteration that holds one of them
  uint32_t* const s_defer = const_cast<uint32_t*>(s_zero) + 8;
  uint32_t wdefer = 0;                                                 // DEFER: entries of my warp's list (warp-uniform)
  uint4* const s_hdr = reinterpret_cast<uint4*>(s_defer + (DEFER ? (NT / 32) * DCAP * DWORDS : 0));
  uint16_t* const s_perm = reinterpret_cast<uint16_t*>(s_hdr + p.rmax);
  uint32_t* const s_goff = reinterpret_cast<uint32_t*>(s_perm + p.rmax);
  __shared__ uint32_t s_cnt[2];                                        // SORTED: simple / complex rows of the round
  uint4 hq[KMAX];                                                      // SORTED: my rows' headers of the coming round
  uint32_t hq_base = 0xffffffffu;                                      //   ... and the round (first read) they belong to
  if (SORTED && threadIdx.x < 2) s_cnt[threadIdx.x] = 0;

  const int tid = threadIdx.x, lane = tid & 31;
  const int col = tid >> 2, ksub = tid & 3;                           // COUNT role
  PROF_DECL
  PROF_INIT;

  for (int i = tid; i < VCL * NCOL; i += NT) s_vc[i] = make_uint4(0, 0, 0, 0);
  if (tid < (int)W) s_rows[p.rmax * W + tid] = make_uint4(0, 0, 0, 0);   // row rmax stays empty: COUNT reads it for padding
  for (int i = tid; i < NCOL + 2 * WMAX + 1; i += NT) s_mod[i] = (uint8_t)((uint32_t)i % W);
  if (BAL && tid < NCOL) s_colrows[tid] = 0;
  if (TMA) {
    if (tid < NT / 32) mbar_init(&s_mbar[tid], 1);
    if (tid < 8) const_cast<uint32_t*>(s_zero)[tid] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  uint32_t next_tile = 0;                                             // thread 0: the tile after the current one
  if (tid == 0) { next_tile = atomicAdd(p.sched, 1u); s_nx[4] = 0xffffffffu; }
  uint4 nh = make_uint4(0, 0, 0, 0);                                  // header, block offset, run word and predecessor
  uint32_t nb0 = 0, ncw = 0, nprev = 0;                                // of this thread's next read: one step ahead
  uint32_t pre_tile = 0xffffffffu;                                     // XU: the unit they were requested for, if it is the next one
  for (;;) {
    __syncthreads();
    if (tid == 0) { s_tile = next_tile; s_nmax = 0; }
    if (tid < 4) s_red[tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.n_ktiles) break;
    if (tid == 0) next_tile = atomicAdd(p.sched, 1u);                 // its latency hides behind this tile's work
    uint2 unit;
    uint32_t r0, r1;
    if (XU && s_nx[4] == tile) {                                        // published during the previous unit's last round
      unit = make_uint2(s_nx[0], s_nx[1]);
      r0 = s_nx[2];
      r1 = s_nx[3];
    } else {
      unit = p.units[tile];
      r0 = p.tile_rng[2 * tile];
      r1 = p.tile_rng[2 * tile + 1];
    }
    const uint32_t ts = unit.x;
    const int ucol = (int)(unit.y >> 5);                                // columns of this unit (<= NCOL)
    const uint32_t r0a = TMA ? (r0 & ~31u) : r0;                       // TMA: rounds start on a group boundary
    const uint32_t n_groups = r1 > r0a ? (r1 - r0a + 31u) >> 5 : 0u;   // (the leading reads end before the tile: empty rows)
    const uint32_t GR = p.rmax >> 5;                                   // groups per round
    uint32_t nv0 = 0, nv1 = 0;                                         // TMA, lane 0: unit range of my warp's next group
    if (TMA && lane == 0 && (uint32_t)(tid >> 5) < n_groups) {         // my warp's first group of this tile
      const uint32_t ra = r0a + 32u * (uint32_t)(tid >> 5);
      const uint32_t v0 = __ldg(&p.vec_off[ra]), v1 = __ldg(&p.vec_off[min(ra + 32u, p.n_reads)]);
      const uint32_t units = min((v1 - v0) >> 5, p.slot_units);
      mbar_expect_tx(&s_mbar[tid >> 5], units * 1024u);
      bulk_g2s(my_slot, p.planes + (size_t)v0 * 8, units * 1024u, &s_mbar[tid >> 5]);
      if ((uint32_t)(tid >> 5) + NT / 32 < n_groups) {
        nv0 = __ldg(&p.vec_off[ra + 256u]);
        nv1 = __ldg(&p.vec_off[min(ra + 288u, p.n_reads)]);
      }
    }
    const uint32_t tcol = ts >> 5, tb = tcol % W;
    const uint32_t slot_c = s_mod[tb + (uint32_t)col];                // where my column's word sits in a row

    uint32_t col_rows = 0;      // rows of my column folded since the last expansion
    uint32_t my_aligned = 0, my_mapped = 0;
    uint32_t covered = 0;
    unsigned long long depth_sum = 0;

    // transpose the shared vertical counters to u32 counts and write the tile (or add to it)
    auto expand = [&](const bool accumulate, const bool last) {
      PROF_MARK(0);
      __syncthreads();
      PROF_MARK(4);
      if (BAL) {
        if (tid < NCOL && s_colrows[tid]) atomicMax(&s_nmax, s_colrows[tid]);
      } else if (ksub == 0 && col_rows) atomicMax(&s_nmax, col_rows);
      __syncthreads();
      const int nl = 32 - __clz(s_nmax);
      const int nlo = nl < 8 ? nl : 8;
      for (int q = tid; q < ucol * 8; q += NT) {
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
          if (accumulate) {   // a column deeper than PASS_ROWS rows: u32 sums saturate and raise MB_EOVERFLOW
            const uint4 o = *dst;
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            if (v.x < o.x || v.y < o.y || v.z < o.z || v.w < o.w) {
              if (v.x < o.x) v.x = 0xffffffffu;
              if (v.y < o.y) v.y = 0xffffffffu;
              if (v.z < o.z) v.z = 0xffffffffu;
              if (v.w < o.w) v.w = 0xffffffffu;
              *p.ovf = 1u;
            }
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
      if (BAL && tid < NCOL) s_colrows[tid] = 0;
      col_rows = 0;
      PROF_MARK(5);
    };

    // ---- placement + row store of one read whose query planes are converted (ql / qh / qv: query word w at index
    // w + 1, indices 0 and WT + 1 zero); every class of read
    auto place_store = [&](uint32_t (&ql)[WT + 2], uint32_t (&qh)[WT + 2], uint32_t (&qv)[WT + 2], const uint32_t loc,
                           const uint32_t qstart, const uint32_t lq, const uint32_t span, const uint32_t cw,
                           const uint32_t fl, const int i) {
          const int rel = (int)(loc >> 5) - (int)tcol;
          const uint32_t slot0 = s_mod[(int)tb + rel + (int)W];
          uint4* __restrict__ dst = s_rows + (uint32_t)i * W;
          uint32_t rl[WT], rh[WT], rv[WT];
          const int o = (int)(loc & 31u);
          const bool simple = (fl & MB_RF_SIMPLE) != 0;
          const bool runs2 = (fl & MB_RF_RUNS2) != 0;              // both runs are in the cplx word: no record fetch
          int wshift = 0;                                           // words the planes have been slid so far
          if (simple || runs2) {
            // run 1 starts at the read's position (x0 = 0). Bases outside the runs carry no candidate bit,
            // so the run needs no mask of its own.
            const int dq = (int)qstart - o;
            int d = (dq - 1) >> 5;
            const int sh = ((dq - 1) & 31) + 1;
            MB_SLIDE(d, wshift, ql, qh, qv);
            if (runs2) {
              // one indel: the candidate plane is split at query bit q_start + len0; the part above it is
              // placed with the second run's shift (inserted bases carry no candidate bit)
              const int len0 = (int)(cw & 0xffffu), ins = (int)((cw >> 16) & 0xffu), del = (int)(cw >> 24);
              const int b1 = (int)qstart + len0 - 32 * (wshift - 1);        // split, relative to array word 0
              uint32_t c2[WT + 2];
#pragma unroll
              for (int w = 0; w <= WT + 1; ++w) {
                const uint32_t above = shl_clamp(0xffffffffu, (uint32_t)max(b1 - 32 * w, 0));
                c2[w] = qv[w] & above;
                qv[w] &= ~above;
              }
#pragma unroll
              for (int j = 0; j < WT; ++j) {
                rv[j] = __funnelshift_rc(qv[j], qv[j + 1], sh);
                rl[j] = __funnelshift_rc(ql[j], ql[j + 1], sh);
                rh[j] = __funnelshift_rc(qh[j], qh[j + 1], sh);
              }
              const int dq2 = dq + ins - del;
              int d2 = ((dq2 - 1) >> 5) - wshift;
              const int sh2 = ((dq2 - 1) & 31) + 1;
              MB_SLIDE(d2, wshift, ql, qh, c2);
#pragma unroll
              for (int j = 0; j < WT; ++j) {
                const uint32_t v2 = __funnelshift_rc(c2[j], c2[j + 1], sh2);
                const uint32_t l2 = __funnelshift_rc(ql[j], ql[j + 1], sh2);
                const uint32_t h2 = __funnelshift_rc(qh[j], qh[j + 1], sh2);
                rl[j] = (rl[j] & rv[j]) | (l2 & ~rv[j]);             // bit select by the first run's support
                rh[j] = (rh[j] & rv[j]) | (h2 & ~rv[j]);
                rv[j] |= v2;
              }
            } else {
#pragma unroll
              for (int j = 0; j < WT; ++j) {
                rv[j] = __funnelshift_rc(qv[j], qv[j + 1], sh);
                rl[j] = __funnelshift_rc(ql[j], ql[j + 1], sh);
                rh[j] = __funnelshift_rc(qh[j], qh[j + 1], sh);
              }
            }
          } else {
            // three or more runs: the read's raw CIGAR words (len << 4 | op) are walked until l_qseq query bases
            // are consumed; every M/=/X run is clipped at ref_span (the contig end) and masked to its own row
            // bits [o + x0, o + x0 + len) (clamped shifts: a PTX shift by >= 32 gives 0)
            const uint32_t* __restrict__ ops = p.cigar + cw;
#pragma unroll
            for (int j = 0; j < WT; ++j) { rl[j] = 0; rh[j] = 0; rv[j] = 0; }
            for (int q0 = 0, x0 = 0; q0 < (int)lq;) {
              const uint32_t opw = __ldg(ops++);
              const uint32_t op = opw & 15u;
              const int oplen = (int)(opw >> 4);
              if (op == mb::OP_I || op == mb::OP_S) { q0 += oplen; continue; }
              if (op == mb::OP_D || op == mb::OP_N) { x0 += oplen; continue; }
              if (op != mb::OP_M && op != mb::OP_EQ && op != mb::OP_X) continue;       // H
              const int len = min(oplen, (int)span - x0);
              if (len > 0) {
                const int p0 = o + x0, dq = q0 - p0;
                int d = ((dq - 1) >> 5) - wshift;
                const int sh = ((dq - 1) & 31) + 1;
                MB_SLIDE(d, wshift, ql, qh, qv);
#pragma unroll
                for (int j = 0; j < WT; ++j) {
                  const int r0b = p0 - 32 * j, r1b = r0b + len;
                  const uint32_t up = shl_clamp(0xffffffffu, (uint32_t)max(r0b, 0));
                  const uint32_t dn = shr_clamp(0xffffffffu, (uint32_t)(32 - min(max(r1b, 0), 32)));
                  const uint32_t msk = up & dn;
                  rl[j] |= __funnelshift_rc(ql[j], ql[j + 1], sh) & msk;
                  rh[j] |= __funnelshift_rc(qh[j], qh[j + 1], sh) & msk;
                  rv[j] |= __funnelshift_rc(qv[j], qv[j + 1], sh) & msk;
                }
              }
              q0 += oplen;
              x0 += oplen;
            }
          }
          // row word j lives in slot (slot0 + j) mod W: two bases, one compare per word
          uint4* __restrict__ const dA = dst + slot0;
          uint4* __restrict__ const dB = dA - W;
          const int wrap = (int)W - (int)slot0;                      // the first j that wraps
#pragma unroll
          for (int j = 0; j < WT; ++j) {
            uint4* __restrict__ const d = j < wrap ? dA : dB;
            d[j] = make_uint4(rv[j] & ~rh[j] & ~rl[j], rv[j] & ~rh[j] & rl[j], rv[j] & rh[j] & ~rl[j],
                              rv[j] & rh[j] & rl[j]);
          }
    };

    // ---- STAGE of one read (row i of a round of n rows): header filters (snps.py:148,154,157), row index
    // table, the read's bit planes, bit-sliced base-quality test, mean-quality filter, CIGAR placement, row store
    auto stage_one = [&](const int i, const int n, const uint4 h, const uint32_t b0, const uint32_t cw,
                         const uint32_t prev_loc, const uint32_t* __restrict__ raw_in, const bool defer_ok = false,
                         const uint32_t row_mask = 0xffffffffu) {
        const uint32_t loc = h.x, lq = h.y & 0xffffu, alen = h.y >> 16, qstart = h.z & 0xffffu, nm = h.z >> 16;
        const uint32_t span = h.w & 0xffffu, mq = (h.w >> 16) & 0xffu, fl = (h.w >> 24) & 0x7fu;
        constexpr int NV8S = (QB * NP + 7) / 8;
        uint32_t raw_sp[SPEC ? 8 * NV8S : 1];
        if (SPEC) {
          // the loads do not wait for the filters (a filtered read's planes are fetched and dropped: the byte model of
          // the kernel counts every read's planes anyway); the filter code and the index table run under their latency
          const uint32_t nv8 = (((lq + 31u) >> 5) * (uint32_t)NP + 7u) >> 3;
          const uint32_t* __restrict__ vp = p.planes + (size_t)b0 * 8;
#pragma unroll
          for (int u = 0; u < NV8S; ++u)
            ldg256((u < 2 || (uint32_t)u < nv8) ? vp + 8 * MB_GROUP * u : p.zero_unit, raw_sp + 8 * u);
        }
        bool needq = false, ovl = false, owner = false;
        if (SORTED) {                              // classified before (bit 31 of the stashed header)
          if (!(fl & MB_RF_LONG)) {
            owner = loc >= ts;
            ovl = span > 0 && loc + span > ts;
            if (owner) ++my_aligned;                                                                // snps.py:143
            needq = (h.w >> 31) != 0;
          }
        } else if (!(fl & MB_RF_LONG)) {
          owner = loc >= ts;
          ovl = span > 0 && loc + span > ts;
          // exact integer forms of the two double-precision tests (tables built in ctx.cu)
          bool hk = (fl & MB_RF_NMSAT) ? (p.pid_always != 0)
                                       : (100 * ((int)alen - (int)nm) >= s_tpid[alen]);   // snps.py:148
          hk = hk && (int)mq >= p.mapq;                                                             // snps.py:154
          hk = hk && (int)alen >= s_tcov[lq];                                              // snps.py:157
          if (owner) ++my_aligned;                                                                  // snps.py:143
          needq = hk && (ovl || owner);
        }
        const int rel = (int)(loc >> 5) - (int)tcol;                      // start column, in [-W, NCOL)
        // s_first[x] = first row of this round whose start column + (W-1) >= x. Rows are sorted by start: only a row
        // that starts in a later 32-position column than its predecessor (or the first / last row) has entries to write
        if (i == 0 || i == n - 1 || (loc >> 5) != (prev_loc >> 5)) {
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
          if (BAL) {   // rows are sorted by start: the first and the last row bound the columns of the round
            if (i == 0) s_cr[0] = (uint32_t)max(rel, 0);
            if (i == n - 1) s_cr[1] = (uint32_t)min(rel + (int)W, ucol);
          }
        }
        // query bit planes, 32 bases per word: low / high bit of the base code, countable mask
        // (query word w lives at index w + 1; index 0 is "word -1", zero until the planes are slid)
        // (words 1..WT are all written by the conversion below; the planes are only read under `keep`)
        uint32_t ql[WT + 2], qh[WT + 2], qv[WT + 2];
        ql[0] = 0; qh[0] = 0; qv[0] = 0;
        ql[WT + 1] = 0; qh[WT + 1] = 0; qv[WT + 1] = 0;
        const uint32_t slot0 = s_mod[(int)tb + rel + (int)W];
        uint4* __restrict__ dst = s_rows + (uint32_t)i * W;
        bool keep = false, placed = false;
        if (needq) {
          uint32_t pc[NQ];
#pragma unroll
          for (int k = 0; k < NQ; ++k) pc[k] = 0;
          constexpr bool LD256 = WT <= 8 && !(VAR & 1);
          // mean-quality filter (snps.py:151): mean(q) < readq  <=>  sum(q) < readq * l_qseq (exact). QUICK: the sum of
          // the two top Phred planes alone is a lower bound of sum(q); for all but poor reads it already reaches the
          // threshold and the four (five) lower planes need not be counted at all
          constexpr bool QUICK = LD256;
          const uint32_t q_thr = p.readq_eff * lq;
          // one 32-base block -> query word w of the three planes
          auto convert = [&](const int w, const uint32_t* __restrict__ blk) {
            // Phred >= baseq, bit-sliced over 32 bases: from the low bit up, "greater or equal so far"
            uint32_t ge = 0xffffffffu;
            if (BQ30) {
              // 30 = 0b011110: plane 0 cannot change an all-ones start, planes 1-4 are ANDed, the planes above ORed - two
              // LOP3 per block instead of one per plane, no threshold registers
              ge = blk[3 + 1] & blk[3 + 2] & blk[3 + 3] & blk[3 + 4];
#pragma unroll
              for (int k = 5; k < NQ; ++k) ge |= blk[3 + k];
#pragma unroll
              for (int k = 0; k < NQ; ++k)
                if (!QUICK || k >= NQ - 2) pc[k] += __popc(blk[3 + k]);
            } else {
#pragma unroll
              for (int k = 0; k < NQ; ++k) {
                const uint32_t pk = blk[3 + k];
                // threshold bit set: ge &= plane, clear: ge |= plane - one LOP3 (lut 0xD4 = c ? a & b : a | b)
                asm("lop3.b32 %0, %1, %0, %2, 0xD4;" : "+r"(ge) : "r"(pk), "r"(tmk[k]));
                if (!QUICK || k >= NQ - 2) pc[k] += __popc(pk);          // QUICK: the two top planes first (below)
              }
            }
            ql[w + 1] = blk[0];
            qh[w + 1] = blk[1];
            qv[w + 1] = BQ30 ? (blk[2] & ge) : (blk[2] & ge & q_allow);
          };
          if (LD256) {
            // the whole read at once: ceil(9 W / 8) independent 256-bit loads; the units of the 32 reads of a
            // group are interleaved, so the lanes of a warp (consecutive reads) read consecutive sectors
            constexpr int NV8 = (QB * NP + 7) / 8;
            const uint32_t nv8 = (((lq + 31u) >> 5) * (uint32_t)NP + 7u) >> 3;
            const uint32_t* __restrict__ vp = p.planes + (size_t)b0 * 8;
            uint32_t raw_ld[(TMA || SPEC) ? 1 : 8 * NV8];
            if (!TMA && !SPEC) {
#pragma unroll
              for (int u = 0; u < NV8; ++u) {
                // every read has a block, i.e. at least NP > 8 words: its first two units always exist; units
                // past the read's end come from one shared all-zero unit (no predicated loads, no zeroed defaults)
                ldg256((u < 2 || (uint32_t)u < nv8) ? vp + 8 * MB_GROUP * u : p.zero_unit, raw_ld + 8 * u);
              }
            }
            const uint32_t* __restrict__ raw = TMA ? raw_in : (SPEC ? raw_sp : raw_ld);   // TMA: already copied out of the warp's slot
#pragma unroll
            for (int bl = 0; bl < QB; ++bl) convert(bl, raw + bl * NP);
#pragma unroll
            for (int bl = QB; bl < WT; ++bl) { ql[bl + 1] = 0; qh[bl + 1] = 0; qv[bl + 1] = 0; }
            if (QUICK) {
              keep = (pc[NQ - 1] << (NQ - 1)) + (pc[NQ - 2] << (NQ - 2)) >= q_thr;
              if (!keep) {                                         // a poor read: the exact sum
#pragma unroll
                for (int bl = 0; bl < QB; ++bl)
#pragma unroll
                  for (int k = 0; k < NQ - 2; ++k) pc[k] += __popc(raw[bl * NP + 3 + k]);
              }
            }
          } else {
            // wide rows: steps of four blocks = NP independent 128-bit loads
            const uint32_t nvec = (((lq + 31u) >> 5) * (uint32_t)NP + 3u) >> 2;
            const uint32_t* __restrict__ vp = p.planes + (size_t)b0 * 8;
#pragma unroll
            for (int s4 = 0; s4 < WT; s4 += 4) {
              const int nb = WT - s4 < 4 ? WT - s4 : 4;
              const int nv = (nb * NP + 3) / 4;
              uint32_t raw[4 * NP];
#pragma unroll
              for (int u = 0; u < NP; ++u) {
                uint4 t = make_uint4(0, 0, 0, 0);
                const int v = s4 / 4 * NP + u;                    // 16-byte half of unit v / 2
                if (u < nv && (uint32_t)v < nvec)
                  t = __ldg(reinterpret_cast<const uint4*>(vp + (size_t)(v >> 1) * (8 * MB_GROUP)) + (v & 1));
                raw[4 * u] = t.x; raw[4 * u + 1] = t.y; raw[4 * u + 2] = t.z; raw[4 * u + 3] = t.w;
              }
#pragma unroll
              for (int bl = 0; bl < 4; ++bl)
                if (bl < nb) convert(s4 + bl, raw + bl * NP);
            }
          }
          if (!keep) {
            uint32_t qsum = 0;
#pragma unroll
            for (int k = 0; k < NQ; ++k) qsum += pc[k] << k;
            keep = qsum >= q_thr;
          }
          if (keep && owner) ++my_mapped;                                    // snps.py:161
        }
        // ---- CIGAR: every M/=/X run is one funnel-shifted copy of the query planes. Query bit q of a run
        // (q0, x0) lies at row bit q - dq with dq = q0 - o - x0 = 32 d + sh, sh in [1, 32]: row word j is
        // query words (j + d, j + d + 1) funnel-shifted by sh; d is -1 for every read that is not heavily
        // clipped, other d first slide the planes by whole words so that the per-word code is static.
        bool now = keep && ovl;
        if (DEFER && defer_ok) {
          // the lanes of this warp that hold a read with an indel take consecutive entries of the warp's list (all
          // lanes of the iteration meet here: row_mask names them)
          const bool want = now && !(fl & MB_RF_SIMPLE);
          const uint32_t dm = __ballot_sync(row_mask, want);
          const uint32_t pos = wdefer + (uint32_t)__popc(dm & ((1u << lane) - 1u));
          wdefer += (uint32_t)__popc(dm);
          if (want && pos < (uint32_t)DCAP) {
            uint32_t* e = s_defer + ((uint32_t)(tid >> 5) * DCAP + pos) * DWORDS;
#pragma unroll
            for (int w = 0; w < WT; ++w) { e[w] = ql[w + 1]; e[WT + w] = qh[w + 1]; e[2 * WT + w] = qv[w + 1]; }
            e[3 * WT] = h.x; e[3 * WT + 1] = h.y; e[3 * WT + 2] = h.z; e[3 * WT + 3] = h.w;
            e[3 * WT + 4] = cw; e[3 * WT + 5] = (uint32_t)i;
            now = false;
            placed = true;
          }
        }
        if (now) {
          place_store(ql, qh, qv, loc, qstart, lq, span, cw, fl, i);
          placed = true;
        }
        if (!placed) {                                          // filtered out, or not on this tile: an empty row
#pragma unroll
          for (int j = 0; j < WT; ++j) dst[j] = make_uint4(0, 0, 0, 0);
        }
    };

    uint32_t pass_rows = 0;
    bool accumulate = false;
    if (TMA) {                                                  // my read of the warp's first group of this tile
      const uint32_t rf = r0a + 32u * (uint32_t)(tid >> 5) + (uint32_t)lane;
      if ((uint32_t)(tid >> 5) < n_groups && rf < r1) {
        nh = __ldg(&p.hdr[rf]);
        ncw = __ldg(&p.cplx[rf]);
        if (lane == 0 && rf > r0a) nprev = __ldg(&p.hdr[rf - 1].x);
      }
    }
    for (uint32_t rb = r0a; rb < r1; rb += p.rmax) {
      const int n = (int)min(p.rmax, r1 - rb);
      if (pass_rows + (uint32_t)n > PASS_ROWS) {  // keep the shared counters below 2^VCL
        expand(accumulate, false);
        accumulate = true;
        pass_rows = 0;
      }
      pass_rows += (uint32_t)n;
      const bool hdr_here = rb == r0 && !(XU && pre_tile == tile);   // else my first header was requested a round ago
      PROF_MARK(0);
      __syncthreads();
      PROF_MARK(4);
      // ---- STAGE: one thread per read, everything in registers: header filters (snps.py:148,154,157),
      // row index table, the read's bit planes as independent 256-bit loads, bit-sliced base-quality test,
      // mean-quality filter, CIGAR placement, row store
      if (TMA) {
        // one aligned group of 32 reads per warp and step; the group's units are one contiguous block in HBM and
        // arrive by one bulk copy issued a whole group (>= 500 instructions per lane) before they are needed
        const int wid = tid >> 5;
        const uint32_t gq = (rb - r0a) >> 5;                                  // first group of this round
        for (uint32_t g = gq + (uint32_t)wid; g < gq + GR && g < n_groups; g += NT / 32) {
          const uint32_t r = r0a + 32u * g + (uint32_t)lane;
          const int i = (int)(r - rb);
          const uint4 h = nh;                                                  // requested one group ago
          const uint32_t cw = ncw;
          uint32_t prev_loc = nprev;
          {
            const uint32_t up = __shfl_up_sync(0xffffffffu, h.x, 1);
            if (lane > 0) prev_loc = up;
          }
          mbar_wait(&s_mbar[wid], mb_phase);                                   // this group's units have landed
          mb_phase ^= 1u;
          // my units out of the slot (unit u of lane l at (32 u + l) x 32 bytes), so that the slot is free for the
          // next group while this one is converted and placed
          constexpr int NV8T = (WT * NP + 7) / 8;
          uint32_t raw[8 * NV8T];
          {
            const uint32_t nv8 = i < n ? min(((((h.y & 0xffffu) + 31u) >> 5) * (uint32_t)NP + 7u) >> 3, p.slot_units) : 0u;
#pragma unroll
            for (int u = 0; u < NV8T; ++u)
              lds256((uint32_t)u < nv8 ? my_slot + (8 * MB_GROUP) * u + 8 * lane : s_zero, raw + 8 * u);
          }
          __syncwarp();
          if (lane == 0 && g + NT / 32 < n_groups) {                           // my next group, into the same slot
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t units = min((nv1 - nv0) >> 5, p.slot_units);
            mbar_expect_tx(&s_mbar[wid], units * 1024u);
            bulk_g2s(my_slot, p.planes + (size_t)nv0 * 8, units * 1024u, &s_mbar[wid]);
            if (g + 2 * (NT / 32) < n_groups) {                                // and the unit range of the one after it
              const uint32_t ra = r0a + 32u * (g + 2 * (NT / 32));
              nv0 = __ldg(&p.vec_off[ra]);
              nv1 = __ldg(&p.vec_off[min(ra + 32u, p.n_reads)]);
            }
          }
          // header and run word of my read of the warp's next group: in flight during this group's work
          const uint4 h_cur = h;
          const uint32_t cw_cur = cw, prev_cur = prev_loc;
          if (g + NT / 32 < n_groups) {
            const uint32_t rn = r + 8u * 32u;
            if (rn < r1) {
              nh = __ldg(&p.hdr[rn]);
              ncw = __ldg(&p.cplx[rn]);
              if (lane == 0) nprev = __ldg(&p.hdr[rn - 1].x);
            }
          }
          if (i < n) stage_one(i, n, h_cur, 0u, cw_cur, prev_cur, raw);
        }
      } else {
      if ((PFL2 || SORTED) && tid == 0) {
        uint32_t pf_r0 = 0, pf_r1 = 0, pf_u0 = 0, pf_u1 = 0;
        // the reads of the next round (or of the next tile's first round) and their unit range: asked for now,
        // needed when this round's STAGE ends
        pf_r0 = rb + p.rmax;
        pf_r1 = min(pf_r0 + p.rmax, r1);
        if (pf_r0 >= r1) {
          pf_r0 = pf_r1 = 0;
          if (next_tile < p.n_ktiles) {
            const uint32_t a = __ldg(&p.tile_rng[2 * next_tile]), b = __ldg(&p.tile_rng[2 * next_tile + 1]);
            pf_r0 = a;
            pf_r1 = min(a + p.rmax, b);
            if (XU) {
              const uint2 nu = __ldg(&p.units[next_tile]);
              s_nx[0] = nu.x; s_nx[1] = nu.y; s_nx[2] = a; s_nx[3] = b; s_nx[4] = next_tile;
            }
          } else if (XU) {
            s_nx[4] = 0xffffffffu;
          }
        }
        if (pf_r1 > pf_r0) {
          pf_u0 = __ldg(&p.vec_off[pf_r0 & ~31u]);
          pf_u1 = __ldg(&p.vec_off[min((pf_r1 + 31u) & ~31u, p.n_reads)]);
        }
        // (the readers of the previous round's values are past the barrier above)
        s_pf[0] = pf_r0; s_pf[1] = pf_r1; s_pf[2] = pf_u0; s_pf[3] = pf_u1 > pf_u0 ? pf_u1 - pf_u0 : 0u;
      }
      if (SORTED) {
        // ---- phase A: my rows' headers (requested before the previous COUNT) are classified - simple (one run),
        // complex (an indel or more: the long placement code), or nothing to stage - and stashed; rows to stage
        // go into s_perm, simple ones from the front, complex ones from the back, so that phase B's warps each
        // run one kind of read: no lane idles through a conversion it does not need, and the few warps that hold
        // the complex reads are the only ones that execute the indel code
        if (hq_base != rb) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            const int i = tid + k * NT;
            hq[k] = i < n ? __ldg(&p.hdr[rb + (uint32_t)i]) : make_uint4(0, 0, 0, 0);
          }
        }
        if ((uint32_t)tid <= (p.rmax >> 5) + 1u)
          s_goff[tid] = __ldg(&p.vec_off[min((rb & ~31u) + 32u * (uint32_t)tid, p.n_reads)]);
        uint32_t skip_mask = 0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const int i = tid + k * NT;
          uint4 h = hq[k];
          int cls = 2;                                                       // 0 simple, 1 complex, 2 nothing to stage
          if (i < n) {
            const uint32_t loc = h.x, lq = h.y & 0xffffu, alen = h.y >> 16, nm = h.z >> 16;
            const uint32_t span = h.w & 0xffffu, mq = (h.w >> 16) & 0xffu, fl = h.w >> 24;
            if (!(fl & MB_RF_LONG)) {
              const bool owner = loc >= ts, ovl = span > 0 && loc + span > ts;
              bool hk = (fl & MB_RF_NMSAT) ? (p.pid_always != 0)
                                           : (100 * ((int)alen - (int)nm) >= s_tpid[alen]);   // snps.py:148
              hk = hk && (int)mq >= p.mapq;                                                             // snps.py:154
              hk = hk && (int)alen >= s_tcov[lq];                                              // snps.py:157
              if (hk && (ovl || owner)) {
                cls = (fl & MB_RF_SIMPLE) ? 0 : 1;
                h.w |= 0x80000000u;
              }
            }
            s_hdr[i] = h;
            if (cls == 2) skip_mask |= 1u << k;
          }
          const uint32_t m0 = __ballot_sync(0xffffffffu, cls == 0), m1 = __ballot_sync(0xffffffffu, cls == 1);
          uint32_t b0 = 0, b1 = 0;
          if (lane == 0) {
            if (m0) b0 = atomicAdd(&s_cnt[0], (uint32_t)__popc(m0));
            if (m1) b1 = atomicAdd(&s_cnt[1], (uint32_t)__popc(m1));
          }
          b0 = __shfl_sync(0xffffffffu, b0, 0);
          b1 = __shfl_sync(0xffffffffu, b1, 0);
          const uint32_t below = (1u << lane) - 1u;
          if (cls == 0) s_perm[b0 + (uint32_t)__popc(m0 & below)] = (uint16_t)i;
          if (cls == 1) s_perm[p.rmax - 1u - (b1 + (uint32_t)__popc(m1 & below))] = (uint16_t)i;
        }
        __syncthreads();
        // ---- phase B
        const uint32_t n_s = s_cnt[0], n_c = s_cnt[1];
        const uint32_t g_first = rb >> 5;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {                                     // my rows with nothing to stage: index table, empty row
          if (skip_mask & (1u << k)) {
            const int i = tid + k * NT;
            stage_one(i, n, s_hdr[i], 0u, 0u, i > 0 ? s_hdr[i - 1].x : 0u, nullptr);
          }
        }
        for (uint32_t slot = (uint32_t)tid; slot < p.rmax; slot += NT) {
          if (slot >= n_s && slot < p.rmax - n_c) continue;
          const int i = (int)s_perm[slot];
          const uint4 h = s_hdr[i];
          const uint32_t r = rb + (uint32_t)i;
          const uint32_t b0 = s_goff[(r >> 5) - g_first] + (r & 31u);
          const uint32_t cw = ((h.w >> 24) & MB_RF_SIMPLE) ? 0xffffffffu : __ldg(&p.cplx[r]);
          stage_one(i, n, h, b0, cw, i > 0 ? s_hdr[i - 1].x : 0u, nullptr);
        }
      } else {
      if (hdr_here) {
        nh = make_uint4(0, 0, 0, 0); nb0 = 0; ncw = 0; nprev = 0;
      }
      if (hdr_here && tid < n) {
        nh = __ldg(&p.hdr[rb + tid]);
        nb0 = __ldg(&p.vec_off[rb + tid]);
        ncw = __ldg(&p.cplx[rb + tid]);
        if (tid > 0) nprev = __ldg(&p.hdr[rb + tid - 1].x);
      }
      for (int i = tid; i < n; i += NT) {
        const uint32_t r = rb + (uint32_t)i;
        const uint4 h = nh;
        const uint32_t b0 = nb0, cw = ncw, prev_loc = nprev;
        if (i + NT < n) {
          nh = __ldg(&p.hdr[r + NT]);
          nb0 = __ldg(&p.vec_off[r + NT]);
          ncw = __ldg(&p.cplx[r + NT]);
          nprev = __ldg(&p.hdr[r + NT - 1].x);
        }
        // the lanes of my warp in this iteration: rows i - lane .. i - lane + 31 below n
        const int left = n - (i - lane);
        stage_one(i, n, h, b0, cw, prev_loc, nullptr, true, left >= 32 ? 0xffffffffu : ((1u << left) - 1u));
      }
      if (DEFER) {
        // the reads this warp put off: one pass (their number rarely exceeds the lanes), all lanes in the indel code
        __syncwarp();
        const uint32_t nd = min(__shfl_sync(0xffffffffu, wdefer, 0), (uint32_t)DCAP);   // lane 0 was in every iteration
        wdefer = 0;
        for (uint32_t k = (uint32_t)lane; k < nd; k += 32) {
          const uint32_t* e = s_defer + ((uint32_t)(tid >> 5) * DCAP + k) * DWORDS;
          uint32_t ql[WT + 2], qh[WT + 2], qv[WT + 2];
          ql[0] = 0; qh[0] = 0; qv[0] = 0;
          ql[WT + 1] = 0; qh[WT + 1] = 0; qv[WT + 1] = 0;
#pragma unroll
          for (int w = 0; w < WT; ++w) { ql[w + 1] = e[w]; qh[w + 1] = e[WT + w]; qv[w + 1] = e[2 * WT + w]; }
          const uint32_t hx = e[3 * WT], hy = e[3 * WT + 1], hz = e[3 * WT + 2], hw = e[3 * WT + 3];
          place_store(ql, qh, qv, hx, hz & 0xffffu, hy & 0xffffu, hw & 0xffffu, e[3 * WT + 4], (hw >> 24) & 0x7fu,
                      (int)e[3 * WT + 5]);
        }
        __syncwarp();
      }
      if (rb + p.rmax < r1) {   // the first read of my next round: requested now, in flight during COUNT
        const uint32_t nr = rb + p.rmax + (uint32_t)tid;
        nh = make_uint4(0, 0, 0, 0); nb0 = 0; ncw = 0; nprev = 0;
        if (nr < r1) {
          nh = __ldg(&p.hdr[nr]);
          nb0 = __ldg(&p.vec_off[nr]);
          ncw = __ldg(&p.cplx[nr]);
          if (tid > 0) nprev = __ldg(&p.hdr[nr - 1].x);
        }
      }
      }
      }
      PROF_MARK(1);
      __syncthreads();
      PROF_MARK(2);
      if (SORTED) {
        // my rows' headers of the coming round (this tile's next round or the next tile's first one): in flight
        // during COUNT; the class counters are free again
        const uint32_t a = s_pf[0], b = s_pf[1];
        hq_base = b > a ? a : 0xffffffffu;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const uint32_t rr = a + (uint32_t)(tid + k * NT);
          hq[k] = rr < b ? __ldg(&p.hdr[rr]) : make_uint4(0, 0, 0, 0);
        }
        if (tid < 2) s_cnt[tid] = 0;
      }
      if (XU && rb + p.rmax >= r1) {
        // the unit's last round: my first read of the next unit, in flight during COUNT, EXPAND and the hand-over
        pre_tile = s_nx[4];
        nh = make_uint4(0, 0, 0, 0); nb0 = 0; ncw = 0; nprev = 0;
        const uint32_t nr = s_nx[2] + (uint32_t)tid;
        if (pre_tile != 0xffffffffu && nr < s_nx[3]) {
          nh = __ldg(&p.hdr[nr]);
          nb0 = __ldg(&p.vec_off[nr]);
          ncw = __ldg(&p.cplx[nr]);
          if (tid > 0) nprev = __ldg(&p.hdr[nr - 1].x);
        }
      }
      if (PFL2 && !TMA) {
        // HBM -> L2 traffic of the next round runs under this round's COUNT: 8 KB pieces of the plane range, then the
        // headers, unit offsets and run words
        const uint32_t a = s_pf[0], b = s_pf[1], u0 = s_pf[2], nu = s_pf[3];
        const uint32_t n_pl = (nu + 255u) >> 8;                               // 256 units = 8 KB
        if ((uint32_t)tid < n_pl) {
          const uint32_t units = min(256u, nu - 256u * (uint32_t)tid);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.planes + ((size_t)u0 + 256u * (uint32_t)tid) * 8),
                       "r"(units * 32u) : "memory");
        } else if (b > a && (uint32_t)tid < n_pl + 3u) {
          const uint32_t k = (uint32_t)tid - n_pl;
          const uint32_t a16 = a & ~3u, nb = ((b - a16 + 3u) & ~3u);          // 16-byte aligned element ranges
          if (k == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.hdr + a16), "r"(nb * 16u) : "memory");
          else if (k == 1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.vec_off + a16), "r"(nb * 4u) : "memory");
          else asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.cplx + a16), "r"(nb * 4u) : "memory");
        }
      }
      // ---- COUNT: my column's rows of this round, KSUB-strided, 8 rows per carry-save tree into
      // bit-sliced register counters; then the KSUB lanes of a column are merged by a shuffle
      // butterfly of bit-sliced adders and lane ksub folds class ksub into the shared counters
      if (BAL) {
        // A round of rmax coordinate-sorted rows touches only part of the tile's columns (C2: ~27 of 64), so with a
        // fixed thread -> column map most warps would idle through COUNT. Here the columns [c0, c1) the round touches
        // are dealt over all threads: L = 4 NS lanes per column, lane = (class A/C/G/T, sub): a lane adds one class
        // word of every NS-th row of its column with a 16-row carry-save tree; the NS partial counters are merged by
        // a shuffle butterfly and sub 0 folds its class into the shared counters.
        const uint32_t c0 = s_cr[0], c1 = s_cr[1];
        const uint32_t nc = c1 > c0 ? c1 - c0 : 0u;
        auto count_bal = [&](auto lns_tag) {
          constexpr int LNS = decltype(lns_tag)::value, NS = 1 << LNS, L = 4 * NS;
          const uint32_t cb = c0 + (uint32_t)(tid / L);
          const bool act = cb < c1;
          const uint32_t colb = act ? cb : c0;
          const int cls = tid & 3, sub = (tid >> 2) & (NS - 1);
          const uint32_t lo = act ? s_first[colb] : 0u, hi = act ? s_first[colb + W] : 0u;
          if (!__any_sync(0xffffffffu, hi > lo)) return;
          uint32_t cnt[NLB];
#pragma unroll
          for (int l = 0; l < NLB; ++l) cnt[l] = 0;
          const uint32_t* __restrict__ base = reinterpret_cast<const uint32_t*>(s_rows + s_mod[tb + colb]) + cls;
          auto add16 = [&](const uint32_t (&in)[16]) {
            uint32_t t[8], f[4], g0, g1, e;
#pragma unroll
            for (int k = 0; k < 8; ++k) CSA(cnt[0], t[k], cnt[0], in[2 * k], in[2 * k + 1]);
#pragma unroll
            for (int k = 0; k < 4; ++k) CSA(cnt[1], f[k], cnt[1], t[2 * k], t[2 * k + 1]);
            CSA(cnt[2], g0, cnt[2], f[0], f[1]);
            CSA(cnt[2], g1, cnt[2], f[2], f[3]);
            CSA(cnt[3], e, cnt[3], g0, g1);
#pragma unroll
            for (int l = 4; l < NLB; ++l) {
              const uint32_t tt = cnt[l] & e;
              cnt[l] ^= e;
              e = tt;
            }
          };
          uint32_t r = lo + (uint32_t)sub;
          const uint32_t* __restrict__ rp = base + r * (W * 4);
          for (; r + 15u * NS < hi; r += 16 * NS, rp += 16 * NS * (W * 4)) {      // whole batches: constant offsets
            uint32_t in[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) in[u] = rp[u * NS * (int)(W * 4)];
            add16(in);
          }
          if (r < hi) {                                                            // the last, partial batch
            uint32_t in[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) in[u] = r + (uint32_t)(u * NS) < hi ? rp[u * NS * (int)(W * 4)] : 0u;
            add16(in);
          }
#pragma unroll
          for (int step = 0; step < LNS; ++step) {       // the sum over the subs is a round's row count at most: NLB levels hold it
            uint32_t carry = 0;
#pragma unroll
            for (int l = 0; l < NLB; ++l) {
              const uint32_t y = __shfl_xor_sync(0xffffffffu, cnt[l], 4 << step);
              CSA(cnt[l], carry, cnt[l], y, carry);
            }
          }
          const bool folder = act && sub == 0 && hi > lo;
          if (folder && cls == 0) s_colrows[colb] += hi - lo;
          uint32_t* vc = reinterpret_cast<uint32_t*>(s_vc) + colb * 4 + cls;
          uint32_t carry = 0;
#pragma unroll
          for (int l = 0; l < VCL; ++l) {
            if (l >= NLB && !__any_sync(0xffffffffu, folder && carry != 0)) break;
            if (folder) {
              const uint32_t a = vc[l * NCOL * 4];
              uint32_t sum;
              CSA(sum, carry, a, l < NLB ? cnt[l < NLB ? l : 0] : 0u, carry);
              vc[l * NCOL * 4] = sum;
            }
          }
        };
        if (nc <= (uint32_t)(NT / 32)) count_bal(std::integral_constant<int, 3>());
        else if (nc <= (uint32_t)(NT / 16)) count_bal(std::integral_constant<int, 2>());
        else if (nc <= (uint32_t)(NT / 8)) count_bal(std::integral_constant<int, 1>());
        else count_bal(std::integral_constant<int, 0>());
      } else {
        const uint32_t lo = s_first[col], hi = col < ucol ? s_first[col + (int)W] : lo;
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
              mrow[u] = base[(rr < hi ? rr : p.rmax) * W];          // past the end: the all-zero row behind the last one
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
      PROF_MARK(3);
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
      const int sp = p.tile_species[ts / MB_TILE];
      const unsigned long long v = s_red[tid];
      if (sp >= 0 && v && !(p.defer_restat && tid < 2)) atomicAdd(&p.acc[5 * sp + 1 + tid], v);
    }
  }
  PROF_MARK(0);
  PROF_FINI;
  // ---- the last CTA to finish moves the stats from the running rows to the caller's (genome_length: snps.py:211),
  // clears the running rows and the scheduler counters for the next launch - no separate initialisation launch - and,
  // multi-GPU, copies the finished rows into the peer buffer (e.g. rank 0's table over NVLink), so that gathering the
  // per-species rows needs no collective
  {
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(p.sched + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (uint32_t i = tid; i < 5u * p.n_species; i += NT) {
        const unsigned long long v = i % 5u == 0 ? (unsigned long long)p.species_len[i / 5u] : __ldcg(p.acc + i);
        p.stats[i] = v;
        if (p.mirror) p.mirror[i] = v;
        p.acc[i] = 0;
      }
      if (tid == 0) { p.sched[0] = 0; p.sched[1] = 0; }
      if (p.mirror) __threadfence_system();
    }
  }
}

__global__ void k_mirror_stats(const unsigned long long* __restrict__ stats, unsigned long long* mirror, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mirror[i] = stats[i];
  __threadfence_system();
}

// ---- reads that do not fit a W-word row: one warp per read, global atomics (rare by construction)
struct LongParams {
  const uint4* hdr;
  const uint32_t* vec_off;
  const uint32_t* cplx;
  const uint32_t* cigar;
  const uint32_t* planes;
  const uint32_t* long_idx;
  const int32_t* tile_species;
  uint32_t* counts;
  unsigned long long* stats;
  uint64_t gpad;
  uint32_t n_long;
  uint32_t* ovf;
  double mapid, aln_cov;
  int readq, mapq;
  uint32_t baseq;
  uint32_t NQ, NP;  // Phred planes / words per 32-base block
};

// word w of a read's planes: 32-byte unit w / 8 of the read, units of a group interleaved (mb_packed_reads)
__device__ __forceinline__ size_t widx(uint32_t w) { return (size_t)(w >> 3) * (8 * MB_GROUP) + (w & 7u); }

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
    const uint32_t* __restrict__ q = p.planes + (size_t)__ldg(&p.vec_off[r]) * 8;
    if (keep) {
      uint32_t s = 0;
      const uint32_t nw = ((lq + 31u) >> 5) * p.NQ;
      for (uint32_t i = lane; i < nw; i += 32) {
        const uint32_t b = i / p.NQ, k = i - b * p.NQ;
        s += (uint32_t)__popc(__ldg(q + widx(b * p.NP + 3 + k))) << k;
      }
      s = warp_sum(s);
      keep = (long long)s >= (long long)p.readq * (long long)lq;
    }
    if (lane == 0 && sp >= 0) {
      atomicAdd(&p.stats[5 * sp + 3], 1ull);
      if (keep) atomicAdd(&p.stats[5 * sp + 4], 1ull);
    }
    if (!keep) continue;
    const bool simple = (fl & MB_RF_SIMPLE) != 0;
    const bool runs2 = (fl & MB_RF_RUNS2) != 0;
    const uint32_t cw = simple ? 0u : __ldg(p.cplx + r);
    // one run: query bases [q0, q0 + len) on reference offsets [x0, x0 + len)
    auto count_run = [&](const uint32_t q0, const uint32_t len, const uint32_t x0) {
      for (uint32_t i = lane; i < len; i += 32) {
        const uint32_t qi = q0 + i, bb = qi & 31u;
        const uint32_t blk = (qi >> 5) * p.NP;                     // first word of the base's block
        uint32_t ph = 0;
        for (uint32_t b = 0; b < p.NQ; ++b) ph |= ((__ldg(q + widx(blk + 3 + b)) >> bb) & 1u) << b;
        if (((__ldg(q + widx(blk + 2)) >> bb) & 1u) && ph >= p.baseq) {     // candidate base with Phred >= baseq
          const uint32_t cd = ((__ldg(q + widx(blk)) >> bb) & 1u) | (((__ldg(q + widx(blk + 1)) >> bb) & 1u) << 1);
          const uint32_t old = atomicAdd(p.counts + (size_t)cd * p.gpad + loc + x0 + i, 1u);
          if (old == 0xffffffffu) {                                 // wrapped: saturate and flag (MB_EOVERFLOW)
            atomicExch(p.counts + (size_t)cd * p.gpad + loc + x0 + i, 0xffffffffu);
            *p.ovf = 1u;
          }
        }
      }
    };
    if (simple) {
      count_run(qstart, span, 0u);
    } else if (runs2) {
      const uint32_t len0 = cw & 0xffffu, ins = (cw >> 16) & 0xffu, del = cw >> 24;
      count_run(qstart, len0, 0u);
      count_run(qstart + len0 + ins, span - (len0 + del), len0 + del);
    } else {
      // raw CIGAR words, walked until l_qseq query bases are consumed; runs clipped at ref_span
      const uint32_t* __restrict__ ops = p.cigar + cw;
      for (uint32_t q0 = 0, x0 = 0; q0 < lq;) {
        const uint32_t opw = __ldg(ops++), op = opw & 15u, oplen = opw >> 4;
        if (op == mb::OP_I || op == mb::OP_S) { q0 += oplen; continue; }
        if (op == mb::OP_D || op == mb::OP_N) { x0 += oplen; continue; }
        if (op != mb::OP_M && op != mb::OP_EQ && op != mb::OP_X) continue;
        if (x0 < span) count_run(q0, min(oplen, span - x0), x0);
        q0 += oplen;
        x0 += oplen;
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

// the row widths (and Phred plane counts) the tile kernel is instantiated for; a batch runs at the smallest
// width >= its row_words
struct PileKernel { uint32_t W; void (*fn)(const PileParams); uint32_t kt; bool tma; bool sorted; bool defer = false; bool cta3 = false; bool bal = false; };   // kt: positions per tile
PileKernel pick_kernel(uint32_t row_words, uint32_t q_planes, bool want_tma, int baseq, uint32_t max_units) {
  // the bulk-copy staging experiment (MB_PILE_TMA=1): the W = 6, six-plane instance only
  static const PileKernel ktma = {6, k_pileup_v2<6, 6, 4>, 2048, true, false};
  if (want_tma && q_planes == 6 && row_words > 5 && row_words <= 6) return ktma;
#define MB_PK(W, Q) {W, k_pileup_v2<W, Q, 1096>, 2048, false, false, false, false, true}
  static const PileKernel k6[] = {MB_PK(3, 6), MB_PK(4, 6), MB_PK(5, 6),  MB_PK(6, 6),  MB_PK(7, 6),
                                  MB_PK(8, 6), MB_PK(10, 6), MB_PK(12, 6), MB_PK(16, 6)};
  static const PileKernel k7[] = {MB_PK(3, 7), MB_PK(4, 7), MB_PK(5, 7),  MB_PK(6, 7),  MB_PK(7, 7),
                                  MB_PK(8, 7), MB_PK(10, 7), MB_PK(12, 7), MB_PK(16, 7)};
#undef MB_PK
  // baseq = 30 (the default of run_midas.py snps) with six Phred planes: the comparator is specialised
#define MB_PB(W) {W, k_pileup_v2<W, 6, 1608>, 2048, false, false, false, false, true}
  static const PileKernel k6b[] = {MB_PB(3), MB_PB(4), MB_PB(5), MB_PB(6), MB_PB(7), MB_PB(8), MB_PB(10), MB_PB(12), MB_PB(16)};
#undef MB_PB
  // experiment switches (MB_PILE_VAR, W = 6 with 6 Phred planes only)
  static const PileKernel kx[4] = {{6, k_pileup_v2<6, 6, 0>, 2048, false, false}, {6, k_pileup_v2<6, 6, 1>, 2048, false, false},
                                   {6, k_pileup_v2<6, 6, 2>, 1024, false, false}, {6, k_pileup_v2<6, 6, 3>, 1024, false, false}};
  static const PileKernel kpf = {6, k_pileup_v2<6, 6, 8>, 2048, false, false};
  static const PileKernel kso = {6, k_pileup_v2<6, 6, 24>, 2048, false, true};
  static const PileKernel kdf = {6, k_pileup_v2<6, 6, 40>, 2048, false, false, true};
  static const PileKernel kbl = {6, k_pileup_v2<6, 6, 72>, 2048, false, false, false, false, true};
  static const PileKernel kb3 = {6, k_pileup_v2<6, 6, 200>, 2048, false, false, false, true, true};
  static const PileKernel kb1 = {6, k_pileup_v2<6, 6, 74>, 1024, false, false, false, false, true};
  static const PileKernel knsp = {6, k_pileup_v2<6, 6, 584>, 2048, false, false, false, false, true};
  if (const char* e = getenv("MB_PILE_VAR")) {
    const long v = atol(e);
    if (v >= 0 && v < 4 && q_planes == 6 && row_words > 5 && row_words <= 6) return kx[v];
    if (v == 8 && q_planes == 6 && row_words > 5 && row_words <= 6) return kpf;
    if (v == 24 && q_planes == 6 && row_words > 5 && row_words <= 6) return kso;
    if (v == 40 && q_planes == 6 && row_words > 5 && row_words <= 6) return kdf;
    if (v == 72 && q_planes == 6 && row_words > 5 && row_words <= 6) return kbl;
    if (v == 200 && q_planes == 6 && row_words > 5 && row_words <= 6) return kb3;
    if (v == 74 && q_planes == 6 && row_words > 5 && row_words <= 6) return kb1;
    if (v == 584 && q_planes == 6 && row_words > 5 && row_words <= 6) return knsp;
  }
  const bool b30 = baseq == 30 && !getenv("MB_PILE_GENERIC_BASEQ");
  if (q_planes == 6 && max_units && !getenv("MB_PILE_FULL_BLOCKS")) {
    // instances for batches whose reads all have at most W - 1 blocks of 32 bases
#define MB_PQ(W, V) {W, k_pileup_v2<W, 6, V>, 2048, false, false, false, false, true}
    static const PileKernel k6q[] = {MB_PQ(3, 5192), MB_PQ(4, 5192), MB_PQ(5, 5192), MB_PQ(6, 5192), MB_PQ(7, 5192), MB_PQ(8, 5192)};
    static const PileKernel k6bq[] = {MB_PQ(3, 5704), MB_PQ(4, 5704), MB_PQ(5, 5704), MB_PQ(6, 5704), MB_PQ(7, 5704), MB_PQ(8, 5704)};
#undef MB_PQ
    const PileKernel* kq = b30 ? k6bq : k6q;
    for (int i = 0; i < 6; ++i)
      if (kq[i].W >= row_words) {
        if (max_units <= ((kq[i].W - 1) * 9 + 7) / 8) return kq[i];
        break;
      }
  }
  const PileKernel* k = q_planes == 6 ? (b30 ? k6b : k6) : k7;
  for (int i = 0; i < 9; ++i)
    if (k[i].W >= row_words) return k[i];
  return k[8];
}

size_t sorted_smem_bytes(uint32_t rmax) { return (size_t)rmax * 18 + 4 * ((size_t)(rmax >> 5) + 4) + 32; }
size_t defer_smem_bytes(uint32_t W, uint32_t n_warps) { return (size_t)n_warps * 16 * (3 * W + 6) * 4; }
size_t pile_smem_bytes(uint32_t W, uint32_t rmax, uint32_t ncol, uint32_t slot_units, uint32_t n_warps, bool sorted,
                       bool defer) {
  return sizeof(uint4) * ((size_t)VCL * ncol + ((size_t)rmax + 1) * W)     // + the empty padding row
         + (slot_units ? (size_t)n_warps * slot_units * 1024 : 0) + 32     // TMA: one slot per warp; a zero unit
         + (defer ? defer_smem_bytes(W, n_warps) : 0)                      // DEFER: the warps' lists of put-off reads
         + (sorted ? sorted_smem_bytes(rmax) : 0);                         // SORTED: header stash, order, group offsets
}

}  // namespace

#ifdef MB_PROF
extern "C" int mb_debug_prof(unsigned long long* out, size_t n_words) {
  const size_t all = sizeof(g_prof) / sizeof(g_prof[0]);
  if (!out || n_words > all) return MB_EINVAL;
  MB_CUDA_TRY(cudaDeviceSynchronize());
  MB_CUDA_TRY(cudaMemcpyFromSymbol(out, g_prof, n_words * sizeof(unsigned long long)));
  return MB_OK;
}
#endif

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
  // shared memory per CTA: two CTAs per SM (override with MB_PILE_SMEM_KB for experiments)
  if (reads->n_planes != 3 + reads->q_planes || (reads->q_planes != 6 && reads->q_planes != 7))
    return mb::fail(MB_EINVAL, "mb_snps_pileup_dev: reads were not packed in snps mode");
  // staging variant. Default: register path (256-bit loads) with the next round's arrays prefetched into L2 under
  // COUNT. MB_PILE_TMA=1 selects the bulk-copy variant (per-warp shared-memory slots, mbarrier completion): measured
  // 17 % slower on C2 (profiles/r02_k_pileup_variants.md), kept as a switch
  bool want_tma = false;
  if (const char* e = getenv("MB_PILE_TMA")) want_tma = atol(e) != 0;
  if (getenv("MB_PILE_VAR")) want_tma = false;
  const PileKernel kern = pick_kernel(reads->row_words, reads->q_planes, want_tma, f->baseq, reads->max_units);
  const uint32_t W = kern.W;                                          // instantiated row width >= row_words
  const uint32_t KT = kern.kt, NCOL = KT / 32, NT = NCOL * KSUB;
  {
    size_t kb = (kern.tma || kern.sorted) ? 112 : (kern.cta3 ? 72 : (KT == 2048 ? 100 : 50));   // 2 / 3 / 4 CTAs per SM
    if (const char* e = getenv("MB_PILE_SMEM_KB")) { long v = atol(e); if (v >= 24 && v <= 220) kb = (size_t)v; }
    ctx->pile_smem = kb * 1024;
    if (const char* e = getenv("MB_PILE_SMEM_B")) { long v = atol(e); if (v >= 24 * 1024 && v <= 220 * 1024) ctx->pile_smem = (size_t)v; }
  }
  // the shared-memory attribute is per kernel instance: set once per (instance, size)
  {
    const void* fn = reinterpret_cast<const void*>(kern.fn);
    auto it = ctx->pile_attr.find(fn);
    if (it == ctx->pile_attr.end() || it->second != ctx->pile_smem) {
      MB_CUDA_TRY(cudaFuncSetAttribute(kern.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->pile_smem));
      ctx->pile_attr[fn] = ctx->pile_smem;
    }
  }
  {
    int occ = 0;
    MB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern.fn, NT, ctx->pile_smem));
    ctx->pile_occ = occ > 0 ? occ : 1;
  }
  // ---- work units. A unit is at most KT positions of one species region; the persistent grid takes them in genome
  // order from an atomic counter. Phase clocks (profiles/r02r_phase.log) show the CTAs of a C2 launch finishing up to
  // one unit time apart (6.6 units of 2048 positions per CTA, SMs differ in speed by a few per cent): 7-10 % of the
  // CTA time idle at the tail. So the units get shorter towards the end of the genome - half length over the last
  // 10 % of the positions, quarter length over the last 3 % - and the last claims are small. Shorter units stage more
  // reads twice (those that straddle a boundary), which is why only the tail gets them. MB_PILE_UNIT overrides the
  // (then uniform) length (tests, sweeps).
  const uint32_t P = (uint32_t)(ctx->sm_count * ctx->pile_occ);
  uint32_t unit_len = KT;
  bool guided = L->gpad / KT > 2ull * P;
  if (const char* e = getenv("MB_PILE_UNIT")) {
    const long v = atol(e);
    if (v >= 32 && v <= (long)KT && v % 32 == 0) { unit_len = (uint32_t)v; guided = false; }
    if (v == 0) guided = false;
  }
  if (guided) unit_len |= 1u;                                    // cache key: the guided table of this tile length
  if (ctx->units_layout != L->id || ctx->units_len != unit_len || getenv("MB_PILE_GUIDE")) {
    std::vector<uint2> tab;
    const uint32_t base_len = unit_len & ~31u;
    uint64_t pct_half = 90, pct_quarter = 97;                       // MB_PILE_GUIDE="half,quarter" (per cent): sweeps
    if (const char* e = getenv("MB_PILE_GUIDE")) {
      unsigned a = 0, b = 0;
      if (sscanf(e, "%u,%u", &a, &b) == 2 && a <= b && b <= 100) { pct_half = a; pct_quarter = b; }
    }
    const uint64_t half_from = guided ? L->gpad / 100 * pct_half : L->gpad;
    const uint64_t quarter_from = guided ? L->gpad / 100 * pct_quarter : L->gpad;
    tab.reserve((size_t)(L->gpad / base_len + L->n_species + 16));
    for (uint32_t sp = 0; sp < L->n_species; ++sp) {
      const uint32_t end = L->species_goff[sp + 1];
      for (uint32_t x = L->species_goff[sp]; x < end;) {
        uint32_t len = x >= quarter_from ? base_len / 4 : (x >= half_from ? base_len / 2 : base_len);
        len = std::max(32u, len & ~31u);
        len = std::min(len, end - x);
        tab.push_back(make_uint2(x, len));
        x += len;
      }
    }
    MB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));            // a launch in flight may still read the old table
    if (ctx->units_cap < tab.size() + 1) {
      if (ctx->units) MB_CUDA_TRY(cudaFree(ctx->units));
      ctx->units = nullptr;
      ctx->units_cap = 0;
      MB_CUDA_TRY(cudaMalloc(&ctx->units, sizeof(uint2) * (tab.size() + 1)));
      ctx->units_cap = tab.size() + 1;
    }
    if (!tab.empty())
      MB_CUDA_TRY(cudaMemcpy(ctx->units, tab.data(), sizeof(uint2) * tab.size(), cudaMemcpyHostToDevice));
    ctx->units_layout = L->id;
    ctx->units_len = unit_len;
    ctx->n_units = (uint32_t)tab.size();
    ctx->rng_for = nullptr;
  }
  const uint32_t n_ktiles = ctx->n_units;
  if (ctx->tile_rng_cap < (size_t)2 * n_ktiles + 2) {
    if (ctx->tile_rng) { MB_CUDA_TRY(cudaStreamSynchronize(ctx->stream)); MB_CUDA_TRY(cudaFree(ctx->tile_rng)); }
    ctx->tile_rng = nullptr;
    ctx->tile_rng_cap = 0;
    ctx->rng_for = nullptr;
    MB_CUDA_TRY(cudaMalloc(&ctx->tile_rng, sizeof(uint32_t) * ((size_t)2 * n_ktiles + 2)));
    ctx->tile_rng_cap = (size_t)2 * n_ktiles + 2;
  }
  const size_t fixed = sizeof(uint4) * (size_t)VCL * NCOL;
  uint32_t rmax, slot_units = 0;
  if (kern.tma) {
    // per warp one slot of slot_units x 1 KB (the units of a group of 32 reads, as they lie in HBM); what is left
    // of the shared memory holds the rows of a round, a whole number of groups per warp
    const uint32_t nv8 = (W * (3u + reads->q_planes) + 7u) / 8u;
    slot_units = reads->max_units ? std::min(nv8, reads->max_units) : nv8;
    if (slot_units < 2) slot_units = 2;
    const size_t slots = (size_t)(NT / 32) * slot_units * 1024 + 32;
    const size_t left = ctx->pile_smem > fixed + slots ? ctx->pile_smem - fixed - slots : 0;
    rmax = (uint32_t)(left / (sizeof(uint4) * W));
    rmax = rmax > 0 ? rmax - 1u : 0u;
    if (rmax > (uint32_t)LANE_ROWS_MAX * KSUB) rmax = LANE_ROWS_MAX * KSUB;
    rmax = rmax / NT * NT;
    if (rmax == 0) return mb::fail(MB_EINVAL, "mb_snps_pileup_dev: no shared memory left for rows (W=%u)", W);
  } else if (kern.sorted) {
    rmax = 3u * NT;                                                              // KMAX rows per thread
    while (rmax > NT && pile_smem_bytes(W, rmax, NCOL, 0, NT / 32, true, false) > ctx->pile_smem) rmax -= NT;
  } else {
    const size_t extra = 32 + (kern.defer ? defer_smem_bytes(W, NT / 32) : 0);
    rmax = (uint32_t)((ctx->pile_smem - fixed - extra) / (sizeof(uint4) * W)) - 1u;   // one row is the empty padding row
    const uint32_t cap = kern.bal ? 1016u : (uint32_t)LANE_ROWS_MAX * KSUB;             // what a lane's register counters hold
    if (rmax > cap) rmax = cap;
    if (rmax >= 2u * NT && !getenv("MB_PILE_RAGGED")) rmax = rmax / NT * NT;   // the same number of reads for every thread
    rmax &= ~7u;
  }
  cudaStream_t st = ctx->stream;
  const uint32_t ns = L->n_species;
  if (ctx->stats_acc_cap < 5 * (size_t)ns + 8) {
    if (ctx->stats_acc) { MB_CUDA_TRY(cudaStreamSynchronize(st)); MB_CUDA_TRY(cudaFree(ctx->stats_acc)); }
    ctx->stats_acc = nullptr;
    ctx->stats_acc_cap = 0;
    const size_t cap = 5 * (size_t)ns + 64;
    MB_CUDA_TRY(cudaMalloc(&ctx->stats_acc, sizeof(unsigned long long) * cap));
    MB_CUDA_TRY(cudaMemsetAsync(ctx->stats_acc, 0, sizeof(unsigned long long) * cap, st));
    ctx->stats_acc_cap = cap;
  }
  if (n_ktiles == 0) {
    // nothing to count: the rows are initialised by a launch of their own (genome_length, zeros)
    k_init_stats<<<ns ? (ns + 255) / 256 : 1, 256, 0, st>>>(
        reinterpret_cast<unsigned long long*>(stats_dev), dl.species_len, ns, ctx->sched);
    ctx->launches++;
    if (ctx->stats_mirror && ns) {
      k_mirror_stats<<<1, 256, 0, st>>>(reinterpret_cast<unsigned long long*>(stats_dev),
                                        reinterpret_cast<unsigned long long*>(ctx->stats_mirror), 5u * ns);
      ctx->launches++;
    }
    MB_CUDA_TRY(cudaGetLastError());
    return MB_OK;
  }
  // the tile -> read-range index depends only on the resident batch: built once per (batch, row width)
  if (ctx->rng_for != reads || ctx->rng_W != W || ctx->rng_ntiles != n_ktiles) {
    k_tile_ranges<<<(n_ktiles + 7) / 8, 256, 0, st>>>(reinterpret_cast<const uint4*>(reads->hdr),
                                                          (uint32_t)reads->n_reads, 32u * W, n_ktiles, ctx->units, ctx->tile_rng);
    ctx->launches++;
    ctx->rng_for = reads;
    ctx->rng_W = W;
    ctx->rng_ntiles = n_ktiles;
  }
  PileParams p;
  p.hdr = reinterpret_cast<const uint4*>(reads->hdr);
  p.vec_off = reads->vec_off;
  p.cplx = reads->cplx;
  p.cigar = reads->cigar;
  p.planes = reads->planes;
  p.zero_unit = reads->planes + (size_t)reads->n_vec * 8;       // upload_reads keeps 256 zero bytes behind every array
  p.tile_rng = ctx->tile_rng;
  p.units = ctx->units;
  p.tile_species = dl.tile_species;
  p.counts = counts_dev;
  p.stats = reinterpret_cast<unsigned long long*>(stats_dev);
  p.sched = ctx->sched;
  p.gpad = L->gpad;
  p.n_ktiles = n_ktiles;
  p.W = W;
  p.rmax = rmax;
  rc = mb::get_filter_tables(ctx, f, &p.t_pid, &p.t_cov);
  if (rc) return rc;
  p.pid_always = (f->mapid != f->mapid || f->mapid == -HUGE_VAL) ? 1u : 0u;
  p.readq_eff = f->readq <= 0 ? 0u : (f->readq > 128 ? 128u : (uint32_t)f->readq);
  p.mapq = f->mapq;
  p.baseq = (uint32_t)f->baseq;
  p.defer_restat = reads->n_long ? 1u : 0u;
  p.slot_units = slot_units;
  p.n_reads = (uint32_t)reads->n_reads;
  p.ovf = ctx->ovf;
  p.mirror = reads->n_long ? nullptr : reinterpret_cast<unsigned long long*>(ctx->stats_mirror);
  p.n_species = ns;
  p.acc = ctx->stats_acc;
  p.species_len = dl.species_len;
  const uint32_t grid = min(n_ktiles, P);
  { int trc = mb::timing_begin(ctx); if (trc) return trc; }
  kern.fn<<<grid, NT, pile_smem_bytes(W, rmax, NCOL, slot_units, NT / 32, kern.sorted, kern.defer), st>>>(p);
  { int trc = mb::timing_end(ctx); if (trc) return trc; }
  ctx->launches++;
  if (reads->n_long) {
    LongParams lp;
    lp.hdr = p.hdr; lp.vec_off = p.vec_off; lp.cplx = p.cplx; lp.cigar = p.cigar; lp.planes = reads->planes;
    lp.long_idx = reads->long_idx;
    lp.tile_species = dl.tile_species;
    lp.counts = counts_dev;
    lp.stats = p.stats;
    lp.gpad = L->gpad;
    lp.n_long = reads->n_long;
    lp.ovf = ctx->ovf;
    lp.mapid = f->mapid; lp.aln_cov = f->aln_cov; lp.readq = f->readq; lp.mapq = f->mapq;
    lp.baseq = (uint32_t)f->baseq;
    lp.NQ = reads->q_planes; lp.NP = reads->n_planes;
    const uint32_t lgrid = min((reads->n_long + 7u) / 8u, (uint32_t)ctx->sm_count * 8u);
    k_long_reads<<<lgrid, 256, 0, st>>>(lp);
    ctx->launches++;
    k_restat<<<min(L->n_tiles, (uint32_t)ctx->sm_count * 8u), 256, 0, st>>>(counts_dev, L->gpad, L->n_tiles,
                                                                            dl.tile_species, p.stats);
    ctx->launches++;
    if (ctx->stats_mirror && ns) {
      k_mirror_stats<<<1, 256, 0, st>>>(p.stats, reinterpret_cast<unsigned long long*>(ctx->stats_mirror), 5u * ns);
      ctx->launches++;
    }
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
    if (e != cudaSuccess) rc = mb::fail(MB_ECUDA, "mb_snps_pileup: %s", cudaGetErrorString(e));
  }
  if (!rc) rc = mb::check_overflow(ctx, "mb_snps_pileup");   // synchronises the stream
  return rc;
}

extern "C" int mb_snps_pileup_plain(mb_ctx* ctx, const mb_plain_reads* plain_host, const mb_layout* L,
                                    const mb_filters* f, uint32_t* counts_host, mb_species_stats* stats_host) {
  if (!ctx || !plain_host || !L || !f || !counts_host || !stats_host)
    return mb::fail(MB_EINVAL, "mb_snps_pileup_plain: null argument");
  MB_CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t cbytes = sizeof(uint32_t) * 4 * (size_t)L->gpad, sbytes = sizeof(mb_species_stats) * (size_t)L->n_species;
  void *counts = nullptr, *stats = nullptr;
  int rc = mb::ctx_reserve(ctx, 0, cbytes, &counts);
  if (!rc) rc = mb::ctx_reserve(ctx, 1, sbytes, &stats);
  if (rc) return rc;
  mb_plain_reads dev;
  mb_dev_reads dr;
  // large batches: seq / qual cross PCIe in chunks on a second stream while the chunks already there are re-laid
  bool overlapped = false;
  rc = mb::pack_from_host_overlapped(ctx, plain_host, L, 0, &dev, &dr, &overlapped);
  if (rc) return rc;
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (!overlapped) {
      if (attempt == 0) {
        rc = mb::upload_plain(ctx, plain_host, &dev);                   // H2D of the decoded records
        if (rc) return rc;
      }
      rc = mb::pack_on_device(ctx, &dev, L, 0, /*staged=*/true, &dr);   // k_pack_*: the re-layout, on the GPU
      if (rc) return rc;
    }
    ctx->rng_for = nullptr;
    rc = mb_snps_pileup_dev(ctx, &dr, L, f, (uint32_t*)counts, (mb_species_stats*)stats);
    ctx->rng_for = nullptr;
    if (rc) { cudaStreamSynchronize(ctx->copy_stream); return rc; }
    if (cbytes) {
      cudaError_t e = cudaMemcpyAsync(counts_host, counts, cbytes, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess && sbytes) e = cudaMemcpyAsync(stats_host, stats, sbytes, cudaMemcpyDeviceToHost, ctx->stream);
      if (e != cudaSuccess) return mb::fail(MB_ECUDA, "mb_snps_pileup_plain: %s", cudaGetErrorString(e));
    }
    if (!overlapped) break;
    // the overlapped pack assumed six Phred planes: its verdict is known once the stream has drained
    MB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    bool respec = false;
    rc = mb::pack_overlapped_verdict(ctx, &respec);
    if (rc) return rc;
    if (!respec) break;
    overlapped = false;                                                  // everything is on the device: re-pack exactly
  }
  if (!rc) rc = mb::check_overflow(ctx, "mb_snps_pileup_plain");
  return rc;
}
