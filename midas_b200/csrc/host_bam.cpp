// BAM -> decoded records (mb_plain_reads), natively: BGZF inflate on all host threads, then a
// parallel record parse. This is the ingest step pysam/htslib performs for the reference when it
// opens <outdir>/snps/temp/genomes.bam (midas/run/snps.py:186) or genes/temp/pangenomes.bam
// (midas/run/genes.py:175-178); records come out in file order, unplaced records (refID < 0) are
// dropped, no index is needed. Format: SAM/BAM specification v1, sections 4.1 (BGZF) and 4.2.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <chrono>
#include <vector>

#include <functional>
#include <thread>
#include <atomic>
#include <chrono>

#include "common.h"

// a buffer whose pages are first touched by the threads that fill it (std::vector would zero it on one thread)
template <class T>
struct RawBuf {
  T* p = nullptr;
  size_t n = 0;
  bool pinned = false;   // page-locked (cudaHostAlloc): the array is about to cross PCIe
  RawBuf() = default;
  RawBuf(const RawBuf&) = delete;
  RawBuf& operator=(const RawBuf&) = delete;
  ~RawBuf() { release(); }
  void release() { if (p) { if (pinned) cudaFreeHost(p); else free(p); } p = nullptr; }
  bool resize(size_t k, bool pin = false) {
    release();
    n = k;
    const size_t bytes = (k ? k : 1) * sizeof(T) + 16;
    pinned = false;
    if (pin) {
      void* q = nullptr;
      if (cudaHostAlloc(&q, bytes, cudaHostAllocDefault) == cudaSuccess) { p = (T*)q; pinned = true; return true; }
      cudaGetLastError();   // no device / pin failure: pageable memory
    }
    if (bytes >= (64u << 20)) {
      // a buffer of gigabytes first touched by many threads at once: 2 MB pages where the system grants them on request
      // (transparent_hugepage = madvise) - 512 times fewer page faults under the address-space lock
      void* q = nullptr;
      if (posix_memalign(&q, 2u << 20, (bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20)) == 0) {
        madvise(q, bytes, MADV_HUGEPAGE);
        p = (T*)q;
        return true;
      }
    }
    p = (T*)malloc(bytes);
    return p != nullptr;
  }
  T* data() { return p; }
  const T* data() const { return p; }
};

struct mb_bam {
  std::vector<std::string> ref_name;
  std::vector<uint32_t> ref_len;
  RawBuf<int32_t> ref_id, pos, nm, l_qseq;
  RawBuf<uint8_t> mapq, seq, qual;
  RawBuf<uint16_t> flag;
  RawBuf<uint64_t> cigar_off, seq_off;
  RawBuf<uint32_t> cigar;
  mb_plain_reads pub;
};

namespace {

inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

template <class F>
void pfor(size_t n, int n_threads, F f) {
  if (n_threads <= 1 || n < 2) { f(0, n); return; }
  std::vector<std::thread> th;
  const size_t per = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const size_t lo = std::min(n, per * t), hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=] { f(lo, hi); });
  }
  for (auto& x : th) x.join();
}

// integer value of the NM aux tag, or -1 when absent / malformed
int32_t find_nm(const uint8_t* p, const uint8_t* end) {
  while (p + 3 <= end) {
    const uint8_t t0 = p[0], t1 = p[1], ty = p[2];
    p += 3;
    size_t sz = 0;
    switch (ty) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'd': sz = 8; break;
      case 'Z': case 'H': {
        const uint8_t* z = (const uint8_t*)memchr(p, 0, (size_t)(end - p));
        if (!z) return -1;
        p = z + 1;
        continue;
      }
      case 'B': {
        if (p + 5 > end) return -1;
        const uint8_t sub = p[0];
        const uint32_t cnt = rd32(p + 1);
        const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        const size_t remaining = (size_t)(end - p);
        if ((size_t)cnt > (remaining - 5) / es) return -1;   // the array would run past the record
        p += 5 + (size_t)cnt * es;
        continue;
      }
      default: return -1;
    }
    if (p + sz > end) return -1;
    if (t0 == 'N' && t1 == 'M') {
      switch (ty) {
        case 'c': return (int8_t)p[0];
        case 'C': return p[0];
        case 's': return (int16_t)rd16(p);
        case 'S': return rd16(p);
        case 'i': return (int32_t)rd32(p);
        case 'I': { const uint32_t v = rd32(p); return v > 0x7fffffffu ? 0x7fffffff : (int32_t)v; }
        default: return -1;
      }
    }
    p += sz;
  }
  return -1;
}

}  // namespace

// The inflated byte stream of a BAM file and what its header says (steps 1-3 of every reader below).
struct BamData {
  RawBuf<uint8_t> data;
  size_t total = 0, first_record = 0;
  std::vector<std::string> ref_name;
  std::vector<uint32_t> ref_len;
  // mb_bam_stream_open_dev: every record's start (of its refID field), refID and CIGAR length, found by a walker
  // thread that trails the inflate; bam_stream_walk then only applies the tid_map
  bool walked = false, header_ok = false;
  int walk_rc = 0;
  std::string walk_msg;
  std::vector<uint64_t> all_off;
  std::vector<int32_t> all_rid;
  std::vector<uint32_t> all_ncig;
};

// BAM header of the first `avail` inflated bytes: 0 = parsed (ref_name / ref_len / first_record set), 1 = the header
// runs past `avail` (call again with more), < 0 = mb_status
static int parse_bam_header(const uint8_t* d, size_t avail, BamData* B, const char* who, const char* path) {
  B->ref_name.clear();
  B->ref_len.clear();
  if (avail < 12) return 1;
  if (memcmp(d, "BAM\1", 4) != 0) return mb::fail(MB_EMALFORMED, "%s: %s: not a BAM file", who, path);
  size_t o = 8 + (size_t)rd32(d + 4);
  if (o + 4 > avail) return 1;
  const uint32_t n_ref = rd32(d + o);
  o += 4;
  for (uint32_t i = 0; i < n_ref; ++i) {
    if (o + 4 > avail) return 1;
    const uint32_t ln = rd32(d + o);
    if (ln == 0) return mb::fail(MB_EMALFORMED, "%s: %s: truncated reference list", who, path);
    if (o + 8 + ln > avail) return 1;
    B->ref_name.emplace_back((const char*)d + o + 4, ln - 1);
    B->ref_len.push_back(rd32(d + o + 4 + ln));
    o += 8 + ln;
  }
  B->first_record = o;
  return 0;
}

// One step of the record walk: records that lie entirely below `avail`, from *o on. < 0 = mb_status.
static int walk_records(const uint8_t* d, size_t avail, size_t total, uint32_t n_ref, size_t* o_io, BamData* B) {
  size_t o = *o_io;
  // the chain is a dependent walk (the next offset is in the record) through lines other cores wrote a moment ago; the
  // records are a few hundred bytes apart at irregular strides, so the lines 4 KB ahead are requested by hand
  size_t pf = o;
  while (o + 4 <= avail) {
    for (const size_t want = std::min(o + 4096, avail); pf < want; pf += 64) __builtin_prefetch(d + pf);
    const uint32_t bs = rd32(d + o);
    if (bs < 32 || o + 4 + bs > total) return mb::fail(MB_EMALFORMED, "mb_bam_stream: truncated record at %zu", o);
    if (o + 4 + 32 > avail) break;                               // its fixed fields are not all there yet
    const uint8_t* r = d + o + 4;
    const int32_t rid = (int32_t)rd32(r);
    const uint32_t l_name = r[8], n_cig = rd16(r + 12), l_seq = rd32(r + 16);
    if (32 + (size_t)l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + (size_t)l_seq > bs)
      return mb::fail(MB_EMALFORMED, "mb_bam_stream: record fields exceed block_size at %zu", o);
    if (rid >= 0 && (uint32_t)rid >= n_ref) return mb::fail(MB_EMALFORMED, "mb_bam_stream: refID %d out of range", rid);
    B->all_off.push_back((uint64_t)(o + 4));
    B->all_rid.push_back(rid);
    B->all_ncig.push_back(n_cig);
    o += 4 + bs;
  }
  *o_io = o;
  return MB_OK;
}

// Optional consumer of the inflated stream while it is being produced (mb_bam_stream_open_dev): `total` is called once
// the block table says how long the stream is, `ready` for consecutive chunks, in order, as soon as every block of a
// chunk is inflated - on the calling thread, while the worker threads go on inflating.
struct ChunkSink {
  size_t chunk_bytes = 16u << 20;
  std::function<int(size_t)> total;
  std::function<int(size_t, size_t, const uint8_t*)> ready;
};

static int inflate_bam(const char* path, int n_threads, BamData* B, const char* who, ChunkSink* sink = nullptr) {
  const bool timing = getenv("MB_BAM_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_phase = now();
  auto phase = [&](const char* name) { if (timing) { const double t = now(); fprintf(stderr, "%s: %-8s %8.1f ms\n", who, name, t - t_phase); t_phase = t; } };
  // the file is mapped, not copied: the inflate threads read the compressed blocks straight from the page cache
  struct RawFile {
    const uint8_t* p = nullptr;
    size_t n = 0;
    ~RawFile() { release(); }
    void release() { if (p && n) munmap(const_cast<uint8_t*>(p), n); p = nullptr; n = 0; }
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
    const uint8_t& operator[](size_t i) const { return p[i]; }
  } raw;
  {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return mb::fail(MB_EINVAL, "%s: cannot open %s", who, path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return mb::fail(MB_EINVAL, "%s: cannot stat %s", who, path); }
    if (st.st_size > 0) {
      void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) { close(fd); return mb::fail(MB_EINVAL, "%s: cannot map %s", who, path); }
      madvise(m, (size_t)st.st_size, MADV_WILLNEED);
      raw.p = (const uint8_t*)m;
      raw.n = (size_t)st.st_size;
    }
    close(fd);
  }
  phase("read");
  // 1. BGZF block table
  struct Blk { size_t off, cbeg, clen; uint32_t isize; size_t dst; };
  std::vector<Blk> blks;
  size_t off = 0, total = 0;
  while (off + 18 <= raw.size()) {
    const uint8_t* b = raw.data() + off;
    if (b[0] != 0x1f || b[1] != 0x8b || b[2] != 8 || !(b[3] & 4))
      return mb::fail(MB_EMALFORMED, "%s: %s: not a BGZF block at offset %zu", who, path, off);
    const uint16_t xlen = rd16(b + 10);
    size_t x = off + 12, xend = x + xlen, bsize = 0;
    while (x + 4 <= xend && xend <= raw.size()) {
      const uint16_t slen = rd16(raw.data() + x + 2);
      if (raw[x] == 66 && raw[x + 1] == 67 && slen == 2 && x + 6 <= xend) bsize = (size_t)rd16(raw.data() + x + 4) + 1;
      x += 4 + slen;
    }
    if (!bsize || off + bsize > raw.size() || bsize < (size_t)xlen + 20)
      return mb::fail(MB_EMALFORMED, "%s: %s: bad BGZF block size at offset %zu", who, path, off);
    Blk k;
    k.off = off;
    k.cbeg = xend;
    k.clen = off + bsize - 8 - xend;
    k.isize = rd32(raw.data() + off + bsize - 4);
    if (k.isize > 65536u)
      return mb::fail(MB_EMALFORMED, "%s: %s: BGZF block at offset %zu claims %u inflated bytes (> 64 KiB)", who, path, off, k.isize);
    k.dst = total;
    total += k.isize;
    blks.push_back(k);
    off += bsize;
  }
  if (off != raw.size()) return mb::fail(MB_EMALFORMED, "%s: %s: trailing bytes after the last BGZF block", who, path);
  phase("blocks");
  // 2. inflate in parallel
  RawBuf<uint8_t>& data = B->data;
  if (!data.resize(total + 64)) return mb::fail(MB_ENOMEM, "%s: out of memory", who);
  std::atomic<int> bad_flag{0};
  std::atomic<int>* bad = &bad_flag;
  static const bool use_zlib_only = [] { const char* e = getenv("MB_INFLATE"); return e && !strcmp(e, "zlib"); }();
  std::atomic<uint64_t> n_fallback{0};
  auto inflate_block = [&](size_t i) {
    const Blk& k = blks[i];
    if (!k.isize) return true;
    // the block decoder of host_inflate.cpp first (the 8-byte CRC / ISIZE footer behind the data is its read slack);
    // whatever it refuses, or whatever fails the CRC after it, goes through zlib
    if (!use_zlib_only && mb::fast_inflate(raw.data() + k.cbeg, k.clen, 8, data.data() + k.dst, k.isize) &&
        crc32(0L, data.data() + k.dst, k.isize) == rd32(raw.data() + k.cbeg + k.clen))
      return true;
    n_fallback.fetch_add(1, std::memory_order_relaxed);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<uint8_t*>(raw.data()) + k.cbeg;
    zs.avail_in = (uInt)k.clen;
    zs.next_out = data.data() + k.dst;
    zs.avail_out = k.isize;
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && zs.total_out == k.isize &&
                    (crc32(0L, data.data() + k.dst, k.isize) == rd32(raw.data() + k.off + (k.cbeg - k.off) + k.clen));
    inflateEnd(&zs);
    return ok;
  };
  if (!sink) {
    pfor(blks.size(), n_threads, [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i)
        if (!inflate_block(i)) { bad->store(1); return; }
    });
  } else {
    // blocks are handed out in file order (an atomic cursor), so the stream completes front to back; a chunk is the
    // blocks that END inside [k C, (k + 1) C): when their count reaches zero the bytes up to the last one's end are
    // final, and the calling thread passes them on while the workers carry on
    int src = sink->total(total);
    if (src) return src;
    const size_t C = sink->chunk_bytes;
    const size_t n_chunks = total ? (total + C - 1) / C : 0;
    std::vector<std::atomic<uint32_t>> left(n_chunks ? n_chunks : 1);
    for (auto& a : left) a.store(0);
    std::vector<size_t> chunk_end(n_chunks, 0);
    for (const Blk& k : blks) {
      if (!k.isize) continue;
      const size_t c = (k.dst + k.isize - 1) / C;
      left[c].fetch_add(1);
      chunk_end[c] = std::max(chunk_end[c], k.dst + (size_t)k.isize);
    }
    std::atomic<size_t> cursor{0};
    std::atomic<size_t> avail{0};          // the stream is final up to here
    std::atomic<int> done{0};              // 1: every chunk is through (or the inflate failed)
    B->all_off.reserve(total / 256 + 16);
    B->all_rid.reserve(total / 256 + 16);
    B->all_ncig.reserve(total / 256 + 16);
    // the walker: header as soon as it is complete, then the block_size chain behind the inflate front
    std::thread walker([&] {
      bool have_header = false;
      size_t o = 0;
      int rc = MB_OK;
      for (;;) {
        const bool last = done.load(std::memory_order_acquire) != 0;
        const size_t a = avail.load(std::memory_order_acquire);
        if (!have_header) {
          const int h = parse_bam_header(data.data(), a, B, who, path);
          if (h < 0) { rc = h; break; }
          if (h == 0) { have_header = true; B->header_ok = true; o = B->first_record; }
        }
        if (have_header) {
          rc = walk_records(data.data(), a, total, (uint32_t)B->ref_name.size(), &o, B);
          if (rc) break;
        }
        if (last && a == avail.load(std::memory_order_acquire)) {
          if (!have_header && !bad->load()) rc = mb::fail(MB_EMALFORMED, "%s: %s: truncated header", who, path);
          break;
        }
        if (a == avail.load(std::memory_order_acquire)) std::this_thread::sleep_for(std::chrono::microseconds(100));
      }
      B->walk_rc = rc;
      if (rc) B->walk_msg = mb_last_error();
      B->walked = true;
      if (timing) fprintf(stderr, "%s: walker done at %8.1f ms after the block table\n", who, now() - t_phase);
    });
    std::vector<std::thread> workers;
    // one hardware thread is left to the walker (5 M dependent steps through lines other cores have just written: it is
    // as long as the inflate itself and must not be time-sliced with it); MB_BAM_WORKERS overrides
    int nw = std::max(1, n_threads - 1);
    if (const char* e = getenv("MB_BAM_WORKERS")) nw = std::max(1, atoi(e));
    for (int t = 0; t < nw; ++t)
      workers.emplace_back([&] {
        for (;;) {
          const size_t i = cursor.fetch_add(1);
          if (i >= blks.size() || bad->load()) return;
          if (!inflate_block(i)) { bad->store(1); return; }
          if (blks[i].isize) left[(blks[i].dst + blks[i].isize - 1) / C].fetch_sub(1, std::memory_order_release);
        }
      });
    size_t sent = 0;
    for (size_t c = 0; c < n_chunks && !src; ++c) {
      while (left[c].load(std::memory_order_acquire) != 0 && !bad->load()) std::this_thread::sleep_for(std::chrono::microseconds(50));
      if (bad->load()) break;
      const size_t end = std::max(chunk_end[c], sent);      // a chunk no block ends in adds nothing
      avail.store(end, std::memory_order_release);
      if (end > sent) src = sink->ready(sent, end - sent, data.data() + sent);
      sent = end;
    }
    if (src) bad->store(2);
    for (auto& w : workers) w.join();
    done.store(1, std::memory_order_release);
    walker.join();
    if (src) return src;
  }
  if (bad->load()) return mb::fail(MB_EMALFORMED, "%s: %s: BGZF inflate / CRC failure", who, path);
  raw.release();
  if (timing) fprintf(stderr, "%s: %llu of %zu blocks through zlib\n", who, (unsigned long long)n_fallback.load(), blks.size());
  phase("inflate");
  // 3. header
  memset(data.data() + total, 0, 64);
  B->total = total;
  if (B->walked) {                          // the walker has read it (and the records) already
    if (!B->header_ok) return mb::fail(B->walk_rc ? B->walk_rc : MB_EMALFORMED, "%s", B->walk_msg.c_str());
    return MB_OK;                            // (a malformed record is reported by the call that uses the stream)
  }
  {
    const int h = parse_bam_header(data.data(), total, B, who, path);
    if (h < 0) return h;
    if (h > 0) return mb::fail(MB_EMALFORMED, "%s: %s: truncated header", who, path);
  }
  return MB_OK;
}

extern "C" int mb_bam_open(const char* path, int n_threads, int flags, mb_bam** out) {
  if (!path || !out) return mb::fail(MB_EINVAL, "mb_bam_open: null argument");
  const bool pin = (flags & MB_BAM_PINNED) != 0, nib = !(flags & MB_BAM_SEQ_BYTES);
  if (n_threads <= 0) n_threads = (int)std::min<unsigned>(32, std::max(1u, std::thread::hardware_concurrency()));
  const bool timing = getenv("MB_BAM_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_phase = now();
  auto phase = [&](const char* name) { if (timing) { const double t = now(); fprintf(stderr, "mb_bam_open: %-8s %8.1f ms\n", name, t - t_phase); t_phase = t; } };
  BamData bd;
  { int rc = inflate_bam(path, n_threads, &bd, "mb_bam_open"); if (rc) return rc; }
  t_phase = now();
  const uint8_t* d = bd.data.data();
  const size_t total = bd.total;
  size_t o = bd.first_record;
  const uint32_t n_ref = (uint32_t)bd.ref_name.size();
  mb_bam* B = new mb_bam();
  B->ref_name = bd.ref_name;
  B->ref_len = bd.ref_len;
  // 4. record table (sequential: 4 bytes per record decide the next offset)
  std::vector<size_t> rec;
  std::vector<uint64_t> coff(1, 0), soff(1, 0);
  while (o + 4 <= total) {
    const uint32_t bs = rd32(d + o);
    if (bs < 32 || o + 4 + bs > total) { delete B; return mb::fail(MB_EMALFORMED, "mb_bam_open: %s: truncated record at %zu", path, o); }
    const uint8_t* r = d + o + 4;
    const int32_t rid = (int32_t)rd32(r);
    const uint32_t l_name = r[8], n_cig = rd16(r + 12), l_seq = rd32(r + 16);
    if (32 + (size_t)l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + (size_t)l_seq > bs) {
      delete B;
      return mb::fail(MB_EMALFORMED, "mb_bam_open: %s: record fields exceed block_size at %zu", path, o);
    }
    if (rid >= 0) {
      if ((uint32_t)rid >= n_ref) { delete B; return mb::fail(MB_EMALFORMED, "mb_bam_open: %s: refID %d out of range", path, rid); }
      rec.push_back(o);
      coff.push_back(coff.back() + n_cig);
      soff.push_back(soff.back() + (nib ? ((uint64_t)l_seq + 1) & ~1ull : (uint64_t)l_seq));   // nibbles: whole bytes per read
    }
    o += 4 + bs;
  }
  phase("records");
  const size_t n = rec.size();
  if (!(B->ref_id.resize(n, pin) && B->pos.resize(n, pin) && B->nm.resize(n, pin) && B->l_qseq.resize(n, pin) &&
        B->mapq.resize(n, pin) && B->flag.resize(n) && B->cigar.resize(coff.back(), pin) &&
        B->seq.resize(nib ? (soff.back() + 1) / 2 : soff.back(), pin) && B->qual.resize(soff.back(), pin) &&
        B->cigar_off.resize(n + 1, pin) && B->seq_off.resize(n + 1, pin))) {
    delete B;
    return mb::fail(MB_ENOMEM, "mb_bam_open: out of memory");
  }
  memcpy(B->cigar_off.data(), coff.data(), sizeof(uint64_t) * (n + 1));
  memcpy(B->seq_off.data(), soff.data(), sizeof(uint64_t) * (n + 1));
  // 5. parallel decode
  pfor(n, n_threads, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
      const uint8_t* r = d + rec[i] + 4;
      const uint32_t bs = rd32(d + rec[i]);
      const uint32_t l_name = r[8], n_cig = rd16(r + 12), l_seq = rd32(r + 16);
      B->ref_id.p[i] = (int32_t)rd32(r);
      B->pos.p[i] = (int32_t)rd32(r + 4);
      B->mapq.p[i] = r[9];
      B->flag.p[i] = rd16(r + 14);
      B->l_qseq.p[i] = (int32_t)l_seq;
      const uint8_t* q = r + 32 + l_name;
      memcpy(B->cigar.data() + B->cigar_off.p[i], q, 4 * (size_t)n_cig);
      q += 4 * (size_t)n_cig;
      if (nib) {
        memcpy(B->seq.data() + (B->seq_off.p[i] >> 1), q, (l_seq + 1) / 2);     // as in the file: two codes per byte
      } else {
        uint8_t* s = B->seq.data() + B->seq_off.p[i];
        for (uint32_t k = 0; k < l_seq; ++k) s[k] = (k & 1) ? (q[k >> 1] & 15) : (q[k >> 1] >> 4);
      }
      q += (l_seq + 1) / 2;
      memcpy(B->qual.data() + B->seq_off.p[i], q, l_seq);
      if (nib && (l_seq & 1)) B->qual.p[B->seq_off.p[i] + l_seq] = 0;            // the pad element of an odd read
      q += l_seq;
      B->nm.p[i] = find_nm(q, r + bs);
    }
  });
  phase("decode");
  B->pub.n_reads = n;
  B->pub.ref_id = B->ref_id.data();
  B->pub.pos = B->pos.data();
  B->pub.mapq = B->mapq.data();
  B->pub.nm = B->nm.data();
  B->pub.l_qseq = B->l_qseq.data();
  B->pub.cigar_off = B->cigar_off.data();
  B->pub.cigar = B->cigar.data();
  B->pub.seq_off = B->seq_off.data();
  B->pub.seq = B->seq.data();
  B->pub.qual = B->qual.data();
  B->pub.seq_nibbles = nib ? 1u : 0u;
  B->pub.reserved = 0;
  *out = B;
  return MB_OK;
}

extern "C" void mb_bam_close(mb_bam* B) { delete B; }
extern "C" uint32_t mb_bam_n_ref(const mb_bam* B) { return B ? (uint32_t)B->ref_name.size() : 0; }
extern "C" const char* mb_bam_ref_name(const mb_bam* B, uint32_t i) { return B && i < B->ref_name.size() ? B->ref_name[i].c_str() : ""; }
extern "C" uint32_t mb_bam_ref_len(const mb_bam* B, uint32_t i) { return B && i < B->ref_len.size() ? B->ref_len[i] : 0; }
extern "C" const mb_plain_reads* mb_bam_reads(const mb_bam* B) { return B ? &B->pub : nullptr; }
extern "C" const uint16_t* mb_bam_flags(const mb_bam* B) { return B ? B->flag.data() : nullptr; }

// ------------------------------------------------------------------------------------------- streaming form
// The inflated record stream goes to the GPU as it is; the device parses the records (csrc/pack.cu k_bam_*). The
// host only inflates, reads the header and walks the 4-byte block_size chain to find where the records start.
struct mb_bam_stream {
  BamData bd;
  mb_ctx* dev_ctx = nullptr;     // mb_bam_stream_open_dev: the context whose staging buffer holds a copy of the stream
};

extern "C" int mb_bam_stream_open(const char* path, int n_threads, mb_bam_stream** out) {
  if (!path || !out) return mb::fail(MB_EINVAL, "mb_bam_stream_open: null argument");
  if (n_threads <= 0) n_threads = (int)std::min<unsigned>(32, std::max(1u, std::thread::hardware_concurrency()));
  mb_bam_stream* S = new mb_bam_stream();
  const int rc = inflate_bam(path, n_threads, &S->bd, "mb_bam_stream_open");
  if (rc) { delete S; return rc; }
  *out = S;
  return MB_OK;
}
extern "C" int mb_bam_stream_open_dev(mb_ctx* c, const char* path, int n_threads, mb_bam_stream** out) {
  if (!c || !path || !out) return mb::fail(MB_EINVAL, "mb_bam_stream_open_dev: null argument");
  if (n_threads <= 0) n_threads = (int)std::min<unsigned>(32, std::max(1u, std::thread::hardware_concurrency()));
  MB_CUDA_TRY(cudaSetDevice(c->device));
  mb_bam_stream* S = new mb_bam_stream();
  uint8_t* dev = nullptr;
  ChunkSink sink;
  sink.total = [&](size_t total) {
    void* d = nullptr;
    c->bam_resident = nullptr;
    const int rc = mb::ctx_reserve(c, 24, total + 512, &d);
    if (rc) return rc;
    dev = (uint8_t*)d;
    MB_CUDA_TRY(cudaMemsetAsync(dev + total, 0, 512, c->stream));
    return (int)MB_OK;
  };
  sink.ready = [&](size_t off, size_t len, const uint8_t* src) {
    // pageable memory: the call returns when the driver has staged the chunk; the DMA itself runs under the next one
    MB_CUDA_TRY(cudaMemcpyAsync(dev + off, src, len, cudaMemcpyHostToDevice, c->stream));
    return (int)MB_OK;
  };
  const int rc = inflate_bam(path, n_threads, &S->bd, "mb_bam_stream_open_dev", &sink);
  if (rc) { delete S; return rc; }
  S->dev_ctx = c;
  c->bam_resident = S;
  *out = S;
  return MB_OK;
}
extern "C" void mb_bam_stream_close(mb_bam_stream* S) {
  if (S && S->dev_ctx && S->dev_ctx->bam_resident == S) S->dev_ctx->bam_resident = nullptr;
  delete S;
}
extern "C" uint32_t mb_bam_stream_n_ref(const mb_bam_stream* S) { return S ? (uint32_t)S->bd.ref_name.size() : 0; }
extern "C" const char* mb_bam_stream_ref_name(const mb_bam_stream* S, uint32_t i) {
  return S && i < S->bd.ref_name.size() ? S->bd.ref_name[i].c_str() : "";
}
extern "C" uint32_t mb_bam_stream_ref_len(const mb_bam_stream* S, uint32_t i) {
  return S && i < S->bd.ref_len.size() ? S->bd.ref_len[i] : 0;
}
extern "C" uint64_t mb_bam_stream_bytes(const mb_bam_stream* S) { return S ? (uint64_t)S->bd.total : 0; }

namespace mb {
const uint8_t* bam_stream_data(const mb_bam_stream* S, size_t* total) {
  *total = S->bd.total;
  return S->bd.data.data();
}

// Offsets (of the refID field, i.e. 4 bytes behind block_size) of the records on references that tid_map keeps
// (tid_map[refID] >= 0; records with refID < 0 - unplaced - are dropped like in mb_bam_open), in file order.
int bam_stream_walk(const mb_bam_stream* S, const int32_t* tid_map, std::vector<uint64_t>* rec_off,
                    std::vector<uint64_t>* cigar_off) {
  const uint8_t* d = S->bd.data.data();
  const size_t total = S->bd.total;
  const uint32_t n_ref = (uint32_t)S->bd.ref_name.size();
  size_t o = S->bd.first_record;
  uint64_t ncig = 0;
  rec_off->clear();
  cigar_off->clear();
  if (S->bd.walked) {                       // mb_bam_stream_open_dev walked the chain behind the inflate
    if (S->bd.walk_rc) return mb::fail(S->bd.walk_rc, "%s", S->bd.walk_msg.c_str());
    // keep the records of the references tid_map keeps; their CIGAR offsets are the running sum of n_cigar. Two
    // passes over slices of the record list on the host threads (count, then fill behind the slices' prefix sums).
    const size_t n = S->bd.all_off.size();
    const int T = (int)std::min<size_t>(std::max(1u, std::min(16u, std::thread::hardware_concurrency())), n / 65536 + 1);
    const size_t per = (n + (size_t)T - 1) / (size_t)T;
    std::vector<uint64_t> kept(T + 1, 0), cig(T + 1, 0);
    const int32_t* rid = S->bd.all_rid.data();
    const uint32_t* nc = S->bd.all_ncig.data();
    const uint64_t* off = S->bd.all_off.data();
    pfor((size_t)T, T, [&](size_t lo, size_t hi) {
      for (size_t t = lo; t < hi; ++t) {
        uint64_t k = 0, c = 0;
        for (size_t i = t * per, e = std::min(n, i + per); i < e; ++i)
          if (rid[i] >= 0 && tid_map[rid[i]] >= 0) { ++k; c += nc[i]; }
        kept[t + 1] = k;
        cig[t + 1] = c;
      }
    });
    for (int t = 0; t < T; ++t) { kept[t + 1] += kept[t]; cig[t + 1] += cig[t]; }
    rec_off->resize(kept[T]);
    cigar_off->resize(kept[T] + 1);
    uint64_t* ro = rec_off->data();
    uint64_t* co = cigar_off->data();
    pfor((size_t)T, T, [&](size_t lo, size_t hi) {
      for (size_t t = lo; t < hi; ++t) {
        uint64_t k = kept[t], c = cig[t];
        for (size_t i = t * per, e = std::min(n, i + per); i < e; ++i)
          if (rid[i] >= 0 && tid_map[rid[i]] >= 0) { ro[k] = off[i]; co[k] = c; ++k; c += nc[i]; }
      }
    });
    co[kept[T]] = cig[T];
    (void)ncig;
    return MB_OK;
  }
  rec_off->reserve((total - o) / 256 + 16);
  cigar_off->reserve((total - o) / 256 + 17);
  while (o + 4 <= total) {
    const uint32_t bs = rd32(d + o);
    if (bs < 32 || o + 4 + bs > total) return mb::fail(MB_EMALFORMED, "mb_bam_stream: truncated record at %zu", o);
    const uint8_t* r = d + o + 4;
    const int32_t rid = (int32_t)rd32(r);
    const uint32_t l_name = r[8], n_cig = rd16(r + 12), l_seq = rd32(r + 16);
    if (32 + (size_t)l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + (size_t)l_seq > bs)
      return mb::fail(MB_EMALFORMED, "mb_bam_stream: record fields exceed block_size at %zu", o);
    if (rid >= 0) {
      if ((uint32_t)rid >= n_ref) return mb::fail(MB_EMALFORMED, "mb_bam_stream: refID %d out of range", rid);
      if (tid_map[rid] >= 0) { rec_off->push_back((uint64_t)(o + 4)); cigar_off->push_back(ncig); ncig += n_cig; }
    }
    o += 4 + bs;
  }
  cigar_off->push_back(ncig);
  return MB_OK;
}
}  // namespace mb
