// Internal declarations shared by the translation units of libmidas_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <atomic>
#include <mutex>
#include <unordered_map>

#include "../../include/midas_b200.h"

namespace mb {

void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);

#define MB_CUDA_TRY(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return mb::fail(MB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                      __FILE__, __LINE__);                                                  \
  } while (0)

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

}  // namespace mb

// --------------------------------------------------------------------------------------
struct mb_layout {
  uint64_t id;  // unique, for device-side caches
  bool genes_only;  // mb_layout_create_genes: only n_contigs is meaningful
  uint32_t n_contigs, n_species;
  uint64_t gpad;
  uint32_t n_tiles;
  std::vector<uint32_t> contig_len, contig_goff;
  std::vector<int32_t> contig_species;
  std::vector<uint32_t> species_goff;  // [n_species+1]
  std::vector<uint64_t> species_len;   // [n_species]
  std::vector<int32_t> tile_species;   // [n_tiles], -1 for none
  std::vector<uint32_t> order;         // [n_contigs] contig indices in layout (global coordinate) order
  bool identity_order;                 // order[k] == k
};

struct mb_packed_owner {  // allocation record behind a library-owned mb_packed_reads
  mb_packed_reads pub;
  bool pinned;
  void* buf[7];
};

struct mb_dev_reads {
  uint64_t n_reads, n_vec, n_cigar_words, total_bases;
  uint32_t max_ref_span, sorted, row_words, n_long, q_planes, n_planes, max_units;
  uint32_t* long_idx;
  mb_read_hdr* hdr;
  uint32_t* vec_off;
  uint32_t* cplx;
  uint32_t* cigar;
  uint32_t* planes;
  uint64_t bytes;
  bool owned;  // the arrays are this object's own allocations (false: context staging buffers)
};

struct mb_dev_layout {  // device mirror of the parts of a layout the kernels read
  int32_t* tile_species;
  uint64_t* species_len;
  uint32_t* contig_len;   // [n_contigs]
  uint32_t* contig_goff;  // [n_contigs]
  uint32_t* order;        // [n_contigs]
};

struct mb_ctx {
  int device;
  int sm_count;
  cudaStream_t own_stream;
  cudaStream_t stream;
  // ring of event pairs bracketing the dominant kernel of each *_dev call
  static constexpr int kRing = 64;
  cudaEvent_t ring[kRing][2];
  uint64_t ring_n;        // pairs recorded since the last timing reset
  double ring_acc_ms;     // elapsed time of pairs already folded out of the ring
  uint64_t ring_acc_n;
  uint64_t launches;
  size_t pile_smem;   // dynamic shared memory per CTA of the pileup kernel (0 = not configured yet)
  int pile_occ;       // resident CTAs per SM at that size
  std::unordered_map<const void*, size_t> pile_attr;  // kernel instances whose shared-memory attribute is set, and to what
  // scratch (grown on demand)
  uint32_t* tile_rng;  // [2*n_tiles]
  size_t tile_rng_cap;
  const mb_dev_reads* rng_for;  // resident batch the tile index in tile_rng was built for (null: none)
  uint32_t rng_W, rng_ntiles;
  const mb_bam_stream* bam_resident;   // the BAM stream whose inflated bytes lie in staging slot 24 (mb_bam_stream_open_dev)
  // work units of the pileup kernel: (first position, positions) of every unit, per species region; rebuilt when the
  // layout, the unit length or the kernel tile changes
  uint2* units;
  size_t units_cap;
  uint64_t units_layout;
  uint32_t units_len, n_units;
  uint32_t* sched;     // dynamic tile scheduler counters
  // per-species stats of the launch in flight: zero between launches (the kernel's last CTA moves them to the
  // caller's rows and clears them), so that no launch is spent on initialising the rows
  unsigned long long* stats_acc;
  size_t stats_acc_cap;
  // exact integer thresholds of the two floating-point read filters, per alen / l_qseq (see ctx.cu)
  int32_t* ftab;       // device: [2][65536]
  double ftab_mapid, ftab_cov;
  bool ftab_valid;
  std::unordered_map<uint64_t, mb_dev_layout> layouts;
  // staging for the host-buffer convenience calls
  static constexpr int kStage = 48;
  void* stage[kStage];
  size_t stage_cap[kStage];
  // device packer (pack.cu)
  void* pack_meta;        // device PackMeta
  void* pack_meta_host;   // pinned mirror
  cudaEvent_t pack_ev[2];
  bool pack_timed;
  // u32 counter overflow flag of the pileup kernels (device word + pinned mirror)
  uint32_t* ovf;
  uint32_t* ovf_host;
  // peer-memory mirror of the stats rows (multi-GPU), or null
  mb_species_stats* stats_mirror;
  // overlapped host-buffer path: the bulk arrays cross PCIe on a second stream, a chunk of reads at a time
  static constexpr int kChunks = 8;
  cudaStream_t copy_stream;
  cudaEvent_t chunk_ev[kChunks + 1];
};

namespace mb {
int timing_begin(mb_ctx* ctx);  // records the start event of the next pair on ctx->stream
int timing_end(mb_ctx* ctx);
int upload_reads(mb_ctx* ctx, const mb_packed_reads* h, mb_dev_reads* d, bool staged);
int ctx_reserve(mb_ctx* ctx, int slot, size_t bytes, void** out);
int get_dev_layout(mb_ctx* ctx, const mb_layout* L, mb_dev_layout* out);
int get_filter_tables(mb_ctx* ctx, const mb_filters* f, const int32_t** t_pid, const int32_t** t_cov);
// device packer: plain arrays already on the device -> packed batch `d`; staged: into context staging buffers
int pack_on_device(mb_ctx* ctx, const mb_plain_reads* dev, const mb_layout* L, int mode, bool staged, mb_dev_reads* d,
                   const uint64_t* qual_off_dev = nullptr);   // qual_off_dev: qualities not at seq_off (BAM stream)
// copy the plain arrays of `host` into context staging buffers (slots from `slot0`), `dev` gets the device pointers
int upload_plain(mb_ctx* ctx, const mb_plain_reads* host, mb_plain_reads* dev);
// host records -> packed batch in context staging buffers with the transfer of seq / qual overlapped with the
// re-layout (chunks of reads). *done = false: the batch does not qualify (small, or contigs must move), nothing was done.
// *respec = true after the final stream sync means: Phred > 63 seen, the caller must re-pack with pack_on_device.
int pack_from_host_overlapped(mb_ctx* ctx, const mb_plain_reads* host, const mb_layout* L, int mode, mb_plain_reads* dev,
                              mb_dev_reads* d, bool* done);
int pack_overlapped_verdict(mb_ctx* ctx, bool* respec);   // after a stream sync: errors / re-pack needed
// streaming BAM (host_bam.cpp): the inflated stream and the start of every kept record
const uint8_t* bam_stream_data(const mb_bam_stream* S, size_t* total);
int bam_stream_walk(const mb_bam_stream* S, const int32_t* tid_map, std::vector<uint64_t>* rec_off,
                    std::vector<uint64_t>* cigar_off);
void free_dev_reads_arrays(mb_dev_reads* d);
bool fast_inflate(const uint8_t* in, size_t in_len, size_t in_slack, uint8_t* out, size_t out_len);
int check_overflow(mb_ctx* ctx, const char* who);   // after a stream sync: MB_EOVERFLOW if the flag is set (and clears it)
}  // namespace mb
