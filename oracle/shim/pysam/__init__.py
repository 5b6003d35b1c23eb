"""Minimal pysam stand-in (TEST INFRASTRUCTURE; see oracle/shim/README.md).

Surface = what the reference's hot path touches:
  midas/run/snps.py:186      pysam.AlignmentFile(bampath, 'rb') as context manager
  midas/run/snps.py:194-199  .count_coverage(contig, start=0, end=len, quality_threshold=, read_callback=)
  midas/run/genes.py:175-180 pysam.AlignmentFile(path, "rb"), .fetch(until_eof=True), .getrname(tid)
  snps.py:145-157 / genes.py:154-166  AlignedSegment.query_alignment_sequence, .query_length, .tags,
                             .query_qualities, .mapping_quality, .reference_id

Semantics follow upstream pysam (libcalignmentfile.pyx::count_coverage,
libcalignedsegment.pyx::get_aligned_pairs/getQueryStart/getQueryEnd); it is a restatement of
that published behaviour, not pysam itself. Own BGZF/BAM reader (SAM spec v1 §4.1-4.2).
"""
import array
import struct
import zlib

__version__ = "0.15.3-shim"

BAM_CMATCH, BAM_CINS, BAM_CDEL, BAM_CREF_SKIP, BAM_CSOFT_CLIP, BAM_CHARD_CLIP, BAM_CPAD, BAM_CEQUAL, BAM_CDIFF = range(9)
_NT16 = "=ACMGRSVTWYHKDBN"


def _inflate_bgzf(raw):
    chunks, off = [], 0
    while off < len(raw):
        if raw[off] != 0x1F or raw[off + 1] != 0x8B:
            raise IOError("bad BGZF magic")
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        p, bsize = off + 12, None
        while p < off + 12 + xlen:
            s1, s2, sl = struct.unpack_from("<BBH", raw, p)
            if (s1, s2) == (66, 67):
                bsize = struct.unpack_from("<H", raw, p + 4)[0] + 1
            p += 4 + sl
        d = zlib.decompressobj(-15)
        chunks.append(d.decompress(raw[off + 12 + xlen:off + bsize - 8]))
        off += bsize
    return b"".join(chunks)


class AlignedSegment(object):
    __slots__ = ("reference_id", "reference_start", "mapping_quality", "flag", "cigartuples", "_seq",
                 "_qual", "tags", "query_name")

    # --- query geometry (pysam getQueryStart / getQueryEnd) ---
    @property
    def query_length(self):
        return len(self._seq)

    @property
    def query_alignment_start(self):
        start = 0
        for op, ln in self.cigartuples or ():
            if op == BAM_CHARD_CLIP:
                continue
            elif op == BAM_CSOFT_CLIP:
                start += ln
            else:
                break
        return start

    @property
    def query_alignment_end(self):
        end = len(self._seq)
        cig = self.cigartuples or ()
        for k in range(len(cig) - 1, 0, -1):      # upstream loop never visits op 0
            op, ln = cig[k]
            if op == BAM_CHARD_CLIP:
                continue
            elif op == BAM_CSOFT_CLIP:
                end -= ln
            else:
                break
        return end

    @property
    def query_alignment_sequence(self):
        return self._seq[self.query_alignment_start:self.query_alignment_end]

    @property
    def query_sequence(self):
        return self._seq

    seq = query_sequence

    @property
    def query_qualities(self):
        if self._qual is None:
            return None
        return array.array("B", self._qual)

    def get_tag(self, name):
        return dict(self.tags)[name]

    def get_aligned_pairs(self, matches_only=False):
        out = []
        qpos, pos = 0, self.reference_start
        for op, ln in self.cigartuples or ():
            if op in (BAM_CMATCH, BAM_CEQUAL, BAM_CDIFF):
                for i in range(pos, pos + ln):
                    out.append((qpos, i))
                    qpos += 1
                pos += ln
            elif op in (BAM_CINS, BAM_CSOFT_CLIP, BAM_CPAD):
                if not matches_only:
                    for i in range(ln):
                        out.append((qpos + i, None))
                qpos += ln
            elif op == BAM_CDEL:
                if not matches_only:
                    for i in range(pos, pos + ln):
                        out.append((None, i))
                pos += ln
            elif op == BAM_CHARD_CLIP:
                pass
            elif op == BAM_CREF_SKIP:
                pos += ln
        return out

    @property
    def reference_end(self):
        end = self.reference_start
        for op, ln in self.cigartuples or ():
            if op in (BAM_CMATCH, BAM_CDEL, BAM_CREF_SKIP, BAM_CEQUAL, BAM_CDIFF):
                end += ln
        return end


_AUX_FMT = {"A": "<c", "c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f", "d": "<d"}


def _parse_aux(data, off, end):
    tags = []
    while off + 3 <= end:
        tag = data[off:off + 2].decode()
        ty = chr(data[off + 2])
        off += 3
        if ty in _AUX_FMT:
            fmt = _AUX_FMT[ty]
            val = struct.unpack_from(fmt, data, off)[0]
            off += struct.calcsize(fmt)
            if ty == "A":
                val = val.decode()
        elif ty in "ZH":
            z = data.index(b"\0", off)
            val = data[off:z].decode()
            off = z + 1
        elif ty == "B":
            sub = chr(data[off])
            cnt = struct.unpack_from("<I", data, off + 1)[0]
            fmt = _AUX_FMT[sub]
            sz = struct.calcsize(fmt)
            val = [struct.unpack_from(fmt, data, off + 5 + i * sz)[0] for i in range(cnt)]
            off += 5 + cnt * sz
        else:
            raise ValueError("unknown aux type " + ty)
        tags.append((tag, val))
    return tags


class AlignmentFile(object):
    def __init__(self, path, mode="rb", **kw):
        if "b" not in mode or "r" not in mode:
            raise ValueError("shim reads BAM only")
        with open(path, "rb") as fh:
            data = _inflate_bgzf(fh.read())
        if data[:4] != b"BAM\x01":
            raise ValueError("not BAM")
        l_text = struct.unpack_from("<i", data, 4)[0]
        self.text = data[8:8 + l_text].decode(errors="replace")
        off = 8 + l_text
        n_ref = struct.unpack_from("<i", data, off)[0]
        off += 4
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            ln = struct.unpack_from("<i", data, off)[0]
            self.references.append(data[off + 4:off + 4 + ln - 1].decode())
            self.lengths.append(struct.unpack_from("<i", data, off + 4 + ln)[0])
            off += 8 + ln
        self._data = data
        self._first = off
        self._tid = dict((n, i) for i, n in enumerate(self.references))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    def close(self):
        self._data = None

    def getrname(self, tid):
        return self.references[tid]

    get_reference_name = getrname

    def get_tid(self, name):
        return self._tid.get(name, -1)

    def get_reference_length(self, name):
        return self.lengths[self._tid[name]]

    def _records(self):
        data, off = self._data, self._first
        while off + 4 <= len(data):
            bs = struct.unpack_from("<i", data, off)[0]
            rec = off + 4
            tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, _a, _b, _c = struct.unpack_from("<iiBBHHHIiii", data, rec)
            o = rec + 32
            a = AlignedSegment()
            a.query_name = data[o:o + l_name - 1].decode()
            o += l_name
            a.cigartuples = [(w & 0xF, w >> 4) for w in struct.unpack_from("<%dI" % n_cig, data, o)] if n_cig else None
            o += 4 * n_cig
            pk = data[o:o + (l_seq + 1) // 2]
            o += (l_seq + 1) // 2
            s = []
            for byte in pk:
                s.append(_NT16[byte >> 4])
                s.append(_NT16[byte & 15])
            a._seq = "".join(s[:l_seq])
            q = data[o:o + l_seq]
            o += l_seq
            a._qual = None if (l_seq and q[0] == 0xFF) else bytes(q)
            a.tags = _parse_aux(data, o, rec + bs)
            a.reference_id, a.reference_start, a.mapping_quality, a.flag = tid, pos, mapq, flag
            off = rec + bs
            yield a

    def fetch(self, contig=None, start=None, stop=None, region=None, until_eof=False, end=None, reference=None):
        contig = contig if contig is not None else reference
        stop = stop if stop is not None else end
        if contig is None:
            for a in self._records():
                yield a
            return
        tid = self._tid[contig]
        lo = 0 if start is None else start
        hi = self.lengths[tid] if stop is None else stop
        for a in self._records():
            if a.reference_id != tid:
                continue
            e = a.reference_end
            if e == a.reference_start:
                e += 1
            if a.reference_start < hi and e > lo:
                yield a

    def count_coverage(self, contig, start=None, stop=None, region=None, quality_threshold=15,
                       read_callback="all", reference=None, end=None):
        contig_length = self.get_reference_length(contig)
        stop = stop if stop is not None else end
        _start = start if start is not None else 0
        _stop = stop if stop is not None else contig_length
        _stop = _stop if _stop < contig_length else contig_length
        if _stop == _start:
            raise ValueError("interval of size 0")
        if _stop < _start:
            raise ValueError("interval of size less than 0")
        length = _stop - _start
        counts = [array.array("L", [0]) * length for _ in range(4)]
        filter_method = 1 if read_callback == "all" else (2 if read_callback == "nofilter" else 0)
        _threshold = quality_threshold or 0
        for read in self.fetch(contig=contig, start=start, stop=stop):
            if filter_method == 1:
                if read.flag & (0x4 | 0x100 | 0x200 | 0x400):
                    continue
            elif filter_method == 2:
                pass
            else:
                if not read_callback(read):
                    continue
            seq = read.seq
            if seq is None:
                continue
            quality = read.query_qualities
            for qpos, refpos in read.get_aligned_pairs(True):
                if qpos is not None and refpos is not None and _start <= refpos < _stop:
                    if (_threshold and quality and quality[qpos] >= _threshold) or not _threshold:
                        ch = seq[qpos]
                        if ch == "A":
                            counts[0][refpos - _start] += 1
                        if ch == "C":
                            counts[1][refpos - _start] += 1
                        if ch == "G":
                            counts[2][refpos - _start] += 1
                        if ch == "T":
                            counts[3][refpos - _start] += 1
        return tuple(counts)
