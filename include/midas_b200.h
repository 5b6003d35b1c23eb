/*
 * midas_b200.h — C-ABI of libmidas_b200.so, the B200 (sm_100a) drop-in for the
 * per-sample pileup hot path of snayfach/MIDAS v1.3.0.
 *
 * The reference has no FFI layer of its own; the seam this ABI replaces is two
 * Python call sites plus the third-party method beneath them (all paths are
 * relative to the reference tree):
 *
 *   mb_snps_pileup*      <- midas/run/snps.py:219-244  pysam_pileup()
 *                           midas/run/snps.py:164-216  species_pileup()   (counting + stats)
 *                           midas/run/snps.py:141-162  keep_read()        (per-read filter)
 *                           midas/run/snps.py:194-199  pysam count_coverage() call site
 *   mb_genes_coverage*   <- midas/run/genes.py:171-203 count_mapped_bp()
 *                           midas/run/genes.py:153-169 keep_read()
 *   mb_bam_open          <- pysam.AlignmentFile(bam,'rb') at midas/run/snps.py:186, genes.py:175 (BAM decode)
 *   mb_format_snps_rows  <- midas/run/snps.py:179-182,201-210 (row text of <species>.snps.gz)
 *   mb_merge_sites_dev   <- midas/merge/snps.py:42-104 GenomicSite (per-site arithmetic of merge_midas.py snps)
 *   mb_format_merge_*    <- midas/merge/snps.py:196-201 (row text of snps_freq.txt / snps_depth.txt)
 *   mb_reads_pack*       <- what pysam/htslib hands to those loops (decoded BAM records),
 *   mb_*_plain              re-laid ON THE GPU (k_pack_* kernels) as the struct-of-arrays the
 *                           pileup kernels read; mb_pack_reads is the host version of the same
 *                           re-layout (kept as its checker and for unsorted input)
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns MB_OK (0) or a negative
 *     mb_status; mb_last_error() gives a thread-local message for the last failure.
 *     The library never calls exit()/abort() and never throws across the ABI
 *     (contrast sys.exit in midas/utility.py:227-232).
 *   - "host" pointers are caller-owned and only read/written during the call.
 *     "dev" pointers are CUDA device pointers on the context's device (e.g. the
 *     data_ptr() of a torch tensor); work on them is enqueued on the context's
 *     stream (mb_ctx_set_stream) and is asynchronous unless stated otherwise.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with MB_ECUDA.
 */
#ifndef MIDAS_B200_H
#define MIDAS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
typedef enum {
  MB_OK = 0,
  MB_EINVAL = -1,     /* bad argument; also: input on which the reference itself raises
                         (missing NM tag -> KeyError snps.py:148; alen==0 or l_qseq==0 ->
                         ZeroDivisionError snps.py:148/157; missing qualities -> np.mean(None)) */
  MB_ECUDA = -2,      /* CUDA runtime error / no device */
  MB_ENOMEM = -3,
  MB_EOVERFLOW = -4,  /* a u32 counter or coordinate would overflow */
  MB_EMALFORMED = -5, /* CIGAR/seq length mismatch, P op, read beyond contig start */
  MB_EUNSORTED = -6   /* snps reads must be sorted by (layout order, pos) */
} mb_status;

const char* mb_version(void);
const char* mb_last_error(void);

/* species regions of the padded coordinate space start on multiples of MB_TILE positions
 * (a multiple of the kernel's tile, so no tile ever holds two species) */
#define MB_TILE 4096u
/* bases per packed block: every per-base bit of 32 consecutive query bases is one 32-bit word */
#define MB_BLOCK 32u
/* consecutive reads whose 32-byte units are interleaved in planes[] (one warp, one read per lane) */
#define MB_GROUP 32u

/* ----------------------------------------------------------------- filters */
/* args['mapid','aln_cov','readq','mapq','baseq'] of the reference
 * (scripts/run_midas.py:410-419 snps defaults 94.0/0.75/20/20/30,
 *  scripts/run_midas.py:277-284 genes defaults 94.0/0.75/20/0, baseq unused). */
typedef struct {
  double mapid;    /* drop if 100*(alen-NM)/float(alen) < mapid      (snps.py:148) */
  double aln_cov;  /* drop if alen/float(l_qseq)       < aln_cov    (snps.py:157) */
  int32_t readq;   /* drop if mean(qual)               < readq      (snps.py:151) */
  int32_t mapq;    /* drop if mapping_quality          < mapq       (snps.py:154) */
  int32_t baseq;   /* per base: count only qual >= baseq; 0 disables the test (pysam) */
  int32_t reserved;
} mb_filters;

/* ------------------------------------------------ plain (BAM-like) records */
/* One decoded alignment per index, exactly the fields the reference loops read
 * from pysam.AlignedSegment (snps.py:145-157, genes.py:154-166,180). */
typedef struct {
  uint64_t n_reads;
  const int32_t* ref_id;      /* [n] index into mb_ref contigs (snps) / genes (genes) */
  const int32_t* pos;         /* [n] 0-based leftmost reference position */
  const uint8_t* mapq;        /* [n] */
  const int32_t* nm;          /* [n] NM tag; negative = tag missing -> MB_EINVAL */
  const int32_t* l_qseq;      /* [n] stored query length (soft clips included) */
  const uint64_t* cigar_off;  /* [n+1] prefix offsets into cigar[] */
  const uint32_t* cigar;      /* BAM encoding: len<<4 | op, op in MIDNSHP=X (0..8) */
  const uint64_t* seq_off;    /* [n+1] offsets (in bases) into seq[] and qual[]: base k of read i is element
                                 seq_off[i] + k; seq_off[i+1] - seq_off[i] >= l_qseq[i] (padding allowed) */
  const uint8_t* seq;         /* BAM 4-bit codes =ACMGRSVTWYHKDBN (A=1,C=2,G=4,T=8): one per byte
                                 (seq_nibbles = 0), or as they lie in a BAM record, two per byte with the
                                 even element in the high nibble (seq_nibbles = 1; every seq_off[i] even) */
  const uint8_t* qual;        /* raw Phred, one per base; 0xFF in the first byte = missing */
  uint32_t seq_nibbles;
  uint32_t reserved;
} mb_plain_reads;

/* -------------------------------------------------------- reference layout */
/* Contigs of all selected species (snps.py:55-67 initialize_contigs). The
 * library lays them out species-major in one padded global coordinate space:
 * each species starts at a multiple of MB_TILE, contigs of a species are
 * consecutive in the order given. */
typedef struct {
  uint32_t n_contigs;
  uint32_t n_species;
  const uint32_t* contig_len;      /* [n_contigs] */
  const int32_t* contig_species;   /* [n_contigs] in [0, n_species) */
} mb_ref;

typedef struct mb_layout mb_layout;
int mb_layout_create(const mb_ref* ref, mb_layout** out);
/* genes mode needs no coordinates (loc = gene index): a layout that only knows how many genes there
 * are, without the 32-bit padded-genome limit of mb_layout_create */
int mb_layout_create_genes(uint32_t n_genes, mb_layout** out);
void mb_layout_destroy(mb_layout* L);
uint64_t mb_layout_gpad(const mb_layout* L);             /* padded genome size, multiple of MB_TILE */
uint32_t mb_layout_n_tiles(const mb_layout* L);
const uint32_t* mb_layout_contig_goff(const mb_layout* L);   /* [n_contigs] global offset of contig base 0 */
const uint32_t* mb_layout_species_goff(const mb_layout* L);  /* [n_species+1] tile-aligned region starts */
const uint64_t* mb_layout_species_len(const mb_layout* L);   /* [n_species] sum of contig lengths */

/* --------------------------------------------------------- packed reads SoA */
/* 16-byte per-read header. For snps `loc` is the global padded coordinate of the
 * first aligned reference base; for genes it is the gene index. */
typedef struct {
  uint32_t loc;
  uint16_t l_qseq;    /* query_length (snps.py:146) */
  uint16_t alen;      /* len(query_alignment_sequence) = l_qseq - leading S - trailing S (snps.py:145) */
  uint16_t q_start;   /* sum of leading soft clips = query index of the first aligned base */
  uint16_t nm;        /* NM tag, saturated at 65535 */
  uint16_t ref_span;  /* reference bases spanned, clipped at the contig end */
  uint8_t mapq;
  uint8_t flags;      /* MB_RF_* */
} mb_read_hdr;

#define MB_RF_SIMPLE 0x01u /* one aligned segment [q_start, q_start+ref_span) -> [loc, loc+ref_span): no CIGAR walk */
#define MB_RF_NMSAT 0x02u  /* NM exceeded 65535 and was saturated */
#define MB_RF_LONG 0x04u   /* snps: the read does not fit a row of `row_words` 32-position words
                              (reference span or query length too long); it is listed in long_idx[]
                              and counted by the long-read kernel instead of the tile kernel */
#define MB_RF_RUNS2 0x08u  /* snps, not SIMPLE: exactly two M/=/X runs, the first starting at q_start and at
                              the read's position; cplx[i] = len0 | inserted << 16 | deleted << 24
                              (second run: q_start+len0+inserted, ref offset len0+deleted, to ref_span) */
/* widest bit-plane row the tile kernel stages (32-position words per read) */
#define MB_ROW_WORDS_MAX 16u

/* Struct of arrays, identical on host and in HBM (upload is a plain copy).
 * The per-base data of a read is bit-planar: the read is cut into MB_BLOCK-base blocks and every
 * per-base bit of a block is one 32-bit word (bit k = base 32*b + k), so the kernels test and place
 * 32 bases with one logic instruction. The words of a read, block after block (n_planes words per
 * block, zero padded to a multiple of 8), are cut into 32-byte units, and the units of MB_GROUP = 32
 * consecutive reads are interleaved: unit u of read i lies at 32-byte unit vec_off[i] + 32*u of
 * planes[], i.e. word w of read i is
 *   planes[8*(vec_off[i] + 32*(w/8)) + w%8],   w = b*n_planes + p  (block b, plane p)
 * so that the 32 lanes of a warp, one read each, fetch 1 KB of consecutive memory with every
 * 256-bit load. A group holds max over its reads of ceil(words/8) units per read.
 * snps mode (n_planes = 3 + q_planes):
 *   p = 0, 1  low / high bit of the base code (A=0 C=1 G=2 T=3; anything else 0)
 *   p = 2     "candidate": the base is a literal A/C/G/T, lies in an M/=/X run of the CIGAR and its
 *             reference position is inside the contig: the only bases count_coverage can ever
 *             count, whatever the filters are (soft clips, insertions, N never count)
 *   p = 3+k   bit k of the base's Phred quality, k < q_planes
 * genes mode (n_planes = q_planes): only the Phred bit planes (the path reads nothing else).
 * q_planes is 6 when every quality of the batch is <= 63, else 7 (Phred <= 127 is required).
 *   cplx[i]   0xFFFFFFFF for MB_RF_SIMPLE reads; the two runs themselves for MB_RF_RUNS2 reads; for
 *             every other read the index into cigar[] of the read's first raw BAM CIGAR word
 *             (len << 4 | op): the kernels walk the ops until l_qseq query bases are consumed and
 *             clip every M/=/X run at ref_span (= the contig end)
 */
typedef struct {
  uint64_t n_reads;
  uint64_t n_vec;            /* 32-byte units in planes[] */
  uint64_t n_cigar_words;
  uint64_t total_bases;      /* sum of l_qseq (the Gbases numerator) */
  uint32_t max_ref_span;
  uint32_t sorted;           /* 1: hdr[].loc ascending (required by the snps path) */
  uint32_t row_words;        /* snps: 32-position words per staged read row, chosen by the packer so
                                that at most 0.1 % of the reads need more (those carry MB_RF_LONG);
                                2..MB_ROW_WORDS_MAX; 0 in genes mode */
  uint32_t n_long;           /* number of MB_RF_LONG reads */
  uint32_t q_planes;         /* Phred bit planes per block: 6 or 7 */
  uint32_t n_planes;         /* words per block: 3 + q_planes (snps) or q_planes (genes) */
  uint32_t max_units;        /* largest number of 32-byte units per read of any group */
  uint32_t reserved;
  const uint32_t* long_idx;  /* [n_long] indices of the MB_RF_LONG reads (ascending from mb_pack_reads, any order
                                from the device packer) */
  const mb_read_hdr* hdr;    /* [n_reads] */
  const uint32_t* vec_off;   /* [n_reads+1] first 32-byte unit of each read in planes[] (its further units
                                follow at a stride of MB_GROUP units); vec_off[n_reads] = n_vec */
  const uint32_t* cplx;      /* [n_reads] */
  const uint32_t* cigar;     /* [n_cigar_words] */
  const uint32_t* planes;    /* [8*n_vec], 32-byte aligned */
} mb_packed_reads;

/* Host-side re-layout of decoded BAM records into the packed SoA (the step
 * pysam/htslib performs for the reference before its loops see a record). Output
 * buffers are library-owned (page-locked when `pinned` != 0 and a CUDA device is
 * present) and released with mb_packed_free. mode 0 = snps (loc = global
 * coordinate from the layout; `sort` != 0 stable-sorts by loc, otherwise unsorted
 * input is MB_EUNSORTED), mode 1 = genes (loc = ref_id, file order preserved). */
int mb_pack_reads(const mb_plain_reads* in, const mb_layout* L, int mode, int sort,
                  int pinned, int n_threads, mb_packed_reads** out);
void mb_packed_free(mb_packed_reads* p);

/* page-locked host memory for buffers that cross PCIe (cudaHostAlloc); MB_ECUDA without a device */
int mb_host_alloc(size_t bytes, void** out);
void mb_host_free(void* p);

/* ------------------------------------------------------------------ context */
typedef struct mb_ctx mb_ctx;
int mb_ctx_create(int device, mb_ctx** out);
void mb_ctx_destroy(mb_ctx* ctx);
int mb_ctx_set_stream(mb_ctx* ctx, void* cuda_stream); /* cudaStream_t; NULL = context-owned stream
                                                          (pass cudaStreamLegacy, (void*)1, for the default stream) */
int mb_ctx_sync(mb_ctx* ctx);
int mb_ctx_device(const mb_ctx* ctx);
/* kernels launched by this context since creation (for bench.py gpu_launches) */
uint64_t mb_ctx_launch_count(const mb_ctx* ctx);
/* device time (ms, CUDA events on the context stream) of the dominant pileup /
 * coverage kernel of the most recent *_dev call; synchronises the stream. */
int mb_ctx_last_kernel_ms(mb_ctx* ctx, float* ms);
/* accumulated device time of that kernel over every *_dev call since the last reset
 * (event pairs recorded on the context stream; reading synchronises the stream) */
int mb_ctx_timing_reset(mb_ctx* ctx);
int mb_ctx_timing_read(mb_ctx* ctx, double* total_ms, uint64_t* n_calls);

/* ------------------------------------------------- device-resident reads */
typedef struct mb_dev_reads mb_dev_reads;
int mb_reads_upload(mb_ctx* ctx, const mb_packed_reads* host, mb_dev_reads** out);
void mb_reads_free(mb_ctx* ctx, mb_dev_reads* r);
uint64_t mb_reads_bytes(const mb_dev_reads* r); /* bytes resident in HBM */
uint64_t mb_reads_total_bases(const mb_dev_reads* r);
uint64_t mb_reads_n(const mb_dev_reads* r);

/* The re-layout of decoded records into the packed SoA, on the GPU (csrc/pack.cu): CIGAR geometry,
 * validation, candidate mask and the bit transposition of sequence / Phred bytes into planes run as
 * kernels over the plain arrays. `plain` holds HOST pointers for mb_reads_pack (copied to the device
 * first) and DEVICE pointers for mb_reads_pack_dev. mode 0 = snps: the records must be in BAM
 * coordinate order (ref_id, then pos, ascending - what `samtools sort` writes, reference
 * midas/run/snps.py:118); contigs are moved as whole blocks into the layout's order, no sort runs.
 * Otherwise MB_EUNSORTED (use mb_pack_reads with sort). mode 1 = genes (file order kept).
 * The result is resident in HBM and byte-identical to mb_reads_upload(mb_pack_reads(...)) up to the
 * indexing of cigar[] / order of long_idx[]. Synchronises the context stream once. */
int mb_reads_pack(mb_ctx* ctx, const mb_plain_reads* plain_host, const mb_layout* L, int mode, mb_dev_reads** out);
int mb_reads_pack_dev(mb_ctx* ctx, const mb_plain_reads* plain_dev, const mb_layout* L, int mode, mb_dev_reads** out);
/* host copy of a resident batch (for checking the device packer against mb_pack_reads); free with mb_packed_free */
int mb_reads_download(mb_ctx* ctx, const mb_dev_reads* r, mb_packed_reads** out);
/* device time (ms) of the k_pack_* kernels of the most recent mb_reads_pack* / mb_*_plain call */
int mb_ctx_last_pack_ms(mb_ctx* ctx, float* ms);

/* -------------------------------------------------------------- snps path */
/* per species: the aln_stats dict of species_pileup (snps.py:172-176, 211-213) */
typedef struct {
  uint64_t genome_length;
  uint64_t covered_bases;
  uint64_t total_depth;
  uint64_t aligned_reads;
  uint64_t mapped_reads;
} mb_species_stats;

/* Device-resident pileup. counts_dev: uint32 [4][gpad] (A,C,G,T planes, padded
 * global coordinates; every element is written, no pre-zeroing needed).
 * stats_dev: mb_species_stats [n_species] (overwritten). Asynchronous. */
int mb_snps_pileup_dev(mb_ctx* ctx, const mb_dev_reads* reads, const mb_layout* L,
                       const mb_filters* f, uint32_t* counts_dev, mb_species_stats* stats_dev);

/* Host-buffer convenience = upload + mb_snps_pileup_dev + download; synchronous.
 * counts_host: uint32 [4][gpad]; stats_host: [n_species]. */
int mb_snps_pileup(mb_ctx* ctx, const mb_packed_reads* reads, const mb_layout* L,
                   const mb_filters* f, uint32_t* counts_host, mb_species_stats* stats_host);

/* From decoded records in (pinned or pageable) host memory: H2D of the plain arrays + device re-layout
 * + mb_snps_pileup_dev + D2H of counts and stats; staging buffers are owned by the context and reused.
 * This is the host-buffer call of the path (what bench.py's e2e times and run/snps.py calls). */
int mb_snps_pileup_plain(mb_ctx* ctx, const mb_plain_reads* plain_host, const mb_layout* L,
                         const mb_filters* f, uint32_t* counts_host, mb_species_stats* stats_host);

/* A batch in which a u32 counter would overflow (more than 2^32-1 reads counted on one position) makes
 * mb_snps_pileup / mb_snps_pileup_plain return MB_EOVERFLOW; the counter saturates at 0xFFFFFFFF. After
 * an asynchronous mb_snps_pileup_dev, ask with mb_ctx_overflowed (synchronises the stream). */
int mb_ctx_overflowed(mb_ctx* ctx, int* flag);

/* Multi-GPU (species sharded by rank, midas/run/snps.py:225-228): a device pointer in ANOTHER GPU's
 * memory (peer / symmetric memory, e.g. rank 0's stats table) to which the kernel epilogue of every
 * following mb_snps_pileup_dev also writes its mb_species_stats rows, so that no collective is needed
 * to gather them. NULL switches the mirror off. */
int mb_ctx_set_stats_mirror(mb_ctx* ctx, mb_species_stats* peer_rows);

/* ------------------------------------------------------------- genes path */
/* per gene: Gene.{aligned_reads,mapped_reads,depth} of count_mapped_bp
 * (genes.py:181-189); mapped_bp = sum of alen over kept reads, depth =
 * mapped_bp / gene_len in fp64 (exact_order = 0) or the reference's running
 * sum of alen/float(gene_len) in file order (exact_order = 1, bit-exact). */
typedef struct {
  uint32_t aligned_reads;
  uint32_t mapped_reads;
  uint64_t mapped_bp;
  double depth;
} mb_gene_out;

int mb_genes_coverage_dev(mb_ctx* ctx, const mb_dev_reads* reads, const uint32_t* gene_len_dev,
                          uint32_t n_genes, const mb_filters* f, int exact_order,
                          mb_gene_out* out_dev);
int mb_genes_coverage(mb_ctx* ctx, const mb_packed_reads* reads, const uint32_t* gene_len_host,
                      uint32_t n_genes, const mb_filters* f, int exact_order,
                      mb_gene_out* out_host);
/* from decoded records in host memory (H2D + device re-layout + coverage + D2H), see mb_snps_pileup_plain */
int mb_genes_coverage_plain(mb_ctx* ctx, const mb_plain_reads* plain_host, const uint32_t* gene_len_host,
                            uint32_t n_genes, const mb_filters* f, int exact_order, mb_gene_out* out_host);

/* ------------------------------------------------------------ merge (snps) */
/* Pooled allele call across samples for every site of one species: the per-site arithmetic of
 * GenomicSite in the reference's merge_midas.py snps (midas/merge/snps.py:42-104, driven from
 * build_sharded_tables :324-360). counts_dev: uint32 [n_samples][4][n_sites] = the per-sample
 * outputs of mb_snps_pileup stacked; mean_depth_dev: double [n_samples] (Sample mean_coverage,
 * midas/merge/merge.py:66-74; must be > 0: the reference divides by it, snps.py:99).
 * Outputs: pooled [4][n_sites] (saturating u32), call [n_sites] (bit field below), count_samples
 * [n_sites] (samples passing site_depth / site_ratio), sample_depth [n_samples][n_sites]
 * (major+minor count), sample_maf double [n_samples][n_sites]. Asynchronous on the context stream. */
typedef struct {
  double snp_freq;    /* --allele_freq (scripts/merge_midas.py:237) */
  double site_ratio;  /* --site_ratio  (:245) */
  int32_t site_depth; /* --site_depth  (:242) */
  int32_t reserved;
} mb_merge_params;
/* call byte of a site: bits 0-1 major allele (index into ACGT), 2-3 minor allele, 4-6 snp_type
 * (0 none, 1 mono, 2 bi, 3 tri, 4 quad), bit 7 = the site has a minor allele. A site has a major
 * allele iff its pooled depth is > 0 (snps.py:52-63); otherwise the byte is 0. */
#define MB_CALL_MAJOR(c) ((c) & 3u)
#define MB_CALL_MINOR(c) (((c) >> 2) & 3u)
#define MB_CALL_SNP_TYPE(c) (((c) >> 4) & 7u)
#define MB_CALL_HAS_MINOR 0x80u
int mb_merge_sites_dev(mb_ctx* ctx, const uint32_t* counts_dev, uint32_t n_samples, uint64_t n_sites,
                       const double* mean_depth_dev, const mb_merge_params* mp, uint32_t* pooled_dev,
                       uint8_t* call_dev, uint32_t* count_samples_dev, uint32_t* sample_depth_dev,
                       double* sample_maf_dev);

/* ---------------------------------------------------------------- BAM ingest */
/* Decode a BAM file (BGZF inflate + record parse on n_threads host threads, 0 = all) into the
 * plain records above, in file order; records with refID < 0 are dropped. Replaces what
 * pysam.AlignmentFile(path, 'rb') + fetch / count_coverage decode for the reference
 * (midas/run/snps.py:186,194; midas/run/genes.py:175-180); no index (snps.py:130-139) is needed.
 * The arrays behind mb_bam_reads() belong to the handle. Host only. */
typedef struct mb_bam mb_bam;
#define MB_BAM_PINNED 1     /* the record arrays are page-locked (they are about to cross PCIe) */
#define MB_BAM_SEQ_BYTES 2  /* seq expanded to one code per byte (default: as in the file, seq_nibbles = 1) */
int mb_bam_open(const char* path, int n_threads, int flags, mb_bam** out);
void mb_bam_close(mb_bam* b);
uint32_t mb_bam_n_ref(const mb_bam* b);
const char* mb_bam_ref_name(const mb_bam* b, uint32_t i);
uint32_t mb_bam_ref_len(const mb_bam* b, uint32_t i);
const mb_plain_reads* mb_bam_reads(const mb_bam* b);
const uint16_t* mb_bam_flags(const mb_bam* b);   /* [n_reads] SAM flags (informational) */

/* Streaming form of the ingest: the host only inflates the BGZF blocks (n_threads host threads), reads the header
 * and finds where the records start; the inflated record stream goes to the GPU as it is and the device parses it
 * (refID, pos, mapq, CIGAR, the NM tag, where seq and qual lie), re-lays it and counts - no host-side record decode.
 * tid_map[i] = index into the layout's contigs of BAM reference i, or -1 to drop its records (reference
 * midas/run/snps.py:190-199 only ever fetches the contigs of the selected species). The records must be in BAM
 * coordinate order and the kept references must keep their relative order in the layout's numbering
 * (MB_EUNSORTED otherwise: use mb_bam_open + mb_pack_reads with sort). */
typedef struct mb_bam_stream mb_bam_stream;
int mb_bam_stream_open(const char* path, int n_threads, mb_bam_stream** out);
/* The same, with the transfer under the inflate: the blocks are inflated front to back and every 16 MB of the stream
 * that is complete crosses PCIe (on ctx's stream, into ctx's staging memory) while the host threads inflate the
 * rest; a following mb_snps_pileup_bam / mb_genes_coverage_bam on the same ctx finds the stream on the device. One
 * resident stream per context: opening another one, or passing a different stream to those calls, replaces it. */
int mb_bam_stream_open_dev(mb_ctx* ctx, const char* path, int n_threads, mb_bam_stream** out);
void mb_bam_stream_close(mb_bam_stream* s);
uint32_t mb_bam_stream_n_ref(const mb_bam_stream* s);
const char* mb_bam_stream_ref_name(const mb_bam_stream* s, uint32_t i);
uint32_t mb_bam_stream_ref_len(const mb_bam_stream* s, uint32_t i);
uint64_t mb_bam_stream_bytes(const mb_bam_stream* s);   /* inflated bytes (what crosses PCIe) */
/* <- midas/run/snps.py:186-199 + :141-216 from the BAM bytes: H2D of the inflated stream, k_bam_* (parse), k_pack_*
 * (re-layout), pileup, D2H. n_reads_out / n_bases_out (may be NULL): records kept and their bases. */
int mb_snps_pileup_bam(mb_ctx* ctx, const mb_bam_stream* s, const int32_t* tid_map, const mb_layout* L,
                       const mb_filters* f, uint32_t* counts_host, mb_species_stats* stats_host,
                       uint64_t* n_reads_out, uint64_t* n_bases_out);
/* <- midas/run/genes.py:175-189 from the BAM bytes (file order kept; tid_map[i] = gene index of BAM reference i or -1) */
int mb_genes_coverage_bam(mb_ctx* ctx, const mb_bam_stream* s, const int32_t* tid_map, const uint32_t* gene_len_host,
                          uint32_t n_genes, const mb_filters* f, int exact_order, mb_gene_out* out_host,
                          uint64_t* n_reads_out);

/* ------------------------------------------------------------- text output */
/* Rows of <species>.snps.gz for n_rows consecutive positions of one contig, starting at 0-based
 * position first_pos0 (reference midas/run/snps.py:201-210):
 *   ref_id \t ref_pos \t ref_allele \t depth \t count_a \t count_c \t count_g \t count_t \n
 * count_x point at the first of the n_rows counters of each plane, ref_bases at the upper-cased
 * reference characters. Returns the number of bytes written (< 0: mb_status). Host only. */
uint64_t mb_format_snps_bound(uint64_t name_len, uint64_t n_rows);
int64_t mb_format_snps_rows(const char* contig_name, const uint8_t* ref_bases, uint64_t first_pos0,
                            uint64_t n_rows, const uint32_t* count_a, const uint32_t* count_c,
                            const uint32_t* count_g, const uint32_t* count_t, char* out, uint64_t cap);

/* Rows of snps_freq.txt / snps_depth.txt of merge_midas.py snps (reference midas/merge/snps.py:196-201) for
 * n_rows kept sites: "site_id \t v(sample 0) \t v(sample 1) ... \n" with '{0:.3g}'.format(maf) resp.
 * str(depth). maf / depth are the sample-major outputs of mb_merge_sites_dev copied to the host (`stride`
 * sites per sample); row r is site site_index[r] (0-based within the arrays), printed as site id
 * id_base + site_index[r] + 1 (:352; id_base = first site of the chunk the arrays hold).
 * Return the number of bytes written (< 0: mb_status). Host only. */
uint64_t mb_format_merge_bound(uint64_t n_rows, uint32_t n_samples);
int64_t mb_format_merge_freq_rows(const uint64_t* site_index, uint64_t n_rows, uint32_t n_samples, uint64_t stride,
                                  const double* maf, uint64_t id_base, char* out, uint64_t cap);
int64_t mb_format_merge_depth_rows(const uint64_t* site_index, uint64_t n_rows, uint32_t n_samples, uint64_t stride,
                                   const uint32_t* depth, uint64_t id_base, char* out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* MIDAS_B200_H */
