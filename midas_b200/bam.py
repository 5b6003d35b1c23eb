"""BAM / BGZF codec (host ingest for the pileup path).

Replaces what the reference obtains through pysam/htslib when it opens
`<outdir>/snps/temp/genomes.bam` (reference midas/run/snps.py:186) and
`<outdir>/genes/temp/pangenomes.bam` (midas/run/genes.py:175): BGZF inflate, BAM
record parse, NM aux lookup. Format per the SAM/BAM specification v1 (sections 4.1,
4.2, 4.2.4). No index is needed: the GPU path buckets by ref_id itself, which is why
`samtools index` (snps.py:130-139) has no counterpart here.

This pure-Python reader is the portable implementation (small inputs, tests, CLI);
the bulk path is the native reader in csrc/ (mb_bam_*), which must agree with it.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

from .soa import PlainReads, RefSet

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_MAX_BLOCK = 0xFF00


def bgzf_compress(data: bytes, level: int = 6) -> bytes:
    out = bytearray()
    for i in range(0, len(data), _MAX_BLOCK):
        chunk = data[i:i + _MAX_BLOCK]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        bsize = len(comp) + 25
        out += struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 66, 67, 2, bsize)
        out += comp
        out += struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    out += _BGZF_EOF
    return bytes(out)


def bgzf_decompress(raw: bytes) -> bytes:
    out = []
    off, n = 0, len(raw)
    while off < n:
        if raw[off:off + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block at offset %d" % off)
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        x, xend, bsize = off + 12, off + 12 + xlen, None
        while x < xend:
            si1, si2, slen = struct.unpack_from("<BBH", raw, x)
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", raw, x + 4)[0] + 1
            x += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC field")
        cdata = raw[xend:off + bsize - 8]
        crc, isize = struct.unpack_from("<II", raw, off + bsize - 8)
        data = zlib.decompress(cdata, -15) if isize else b""
        if len(data) != isize or (zlib.crc32(data) & 0xFFFFFFFF) != crc:
            raise ValueError("BGZF block CRC/size mismatch")
        out.append(data)
        off += bsize
    return b"".join(out)


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def write_bam(path: str, ref_names, ref_lens, reads: PlainReads, extra_tags: bool = True,
              header_text: str | None = None, level: int = 6) -> None:
    """Write `reads` as a BAM file (records in the given order)."""
    reads = reads.unpacked()
    if header_text is None:
        header_text = "@HD\tVN:1.0\tSO:unsorted\n" + "".join(
            f"@SQ\tSN:{n}\tLN:{int(l)}\n" for n, l in zip(ref_names, ref_lens))
    buf = bytearray()
    buf += b"BAM\x01"
    ht = header_text.encode()
    buf += struct.pack("<i", len(ht)) + ht
    buf += struct.pack("<i", len(ref_names))
    for n, l in zip(ref_names, ref_lens):
        nb = n.encode() + b"\0"
        buf += struct.pack("<i", len(nb)) + nb + struct.pack("<i", int(l))
    flag = reads.flag if reads.flag is not None else np.zeros(reads.n_reads, np.uint16)
    for i in range(reads.n_reads):
        c0, c1 = int(reads.cigar_off[i]), int(reads.cigar_off[i + 1])
        s0, s1 = int(reads.seq_off[i]), int(reads.seq_off[i + 1])
        cig = reads.cigar[c0:c1]
        lq = s1 - s0
        span = int(sum(int(w) >> 4 for w in cig if (int(w) & 0xF) in (0, 2, 3, 7, 8)))
        pos = int(reads.pos[i])
        name = b"r%d\0" % i
        codes = reads.seq[s0:s1]
        if lq & 1:
            codes = np.concatenate((codes, np.zeros(1, np.uint8)))
        packed = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).tobytes()
        aux = b""
        if extra_tags:
            aux += b"ASi" + struct.pack("<i", -3 * int(max(0, reads.nm[i])))
            aux += b"XNC" + struct.pack("<B", 0)
        nm = int(reads.nm[i])
        if nm >= 0:
            if nm < 256:
                aux += b"NMC" + struct.pack("<B", nm)
            elif nm < 65536:
                aux += b"NMS" + struct.pack("<H", nm)
            else:
                aux += b"NMI" + struct.pack("<I", nm)
        if extra_tags:
            aux += b"MDZ" + b"%d\0" % lq
            aux += b"YTZUU\0"
        body = struct.pack("<iiBBHHHIiii", int(reads.ref_id[i]), pos, len(name), int(reads.mapq[i]),
                           _reg2bin(pos, pos + max(1, span)), len(cig), int(flag[i]), lq, -1, -1, 0)
        body += name + cig.astype("<u4").tobytes() + packed + reads.qual[s0:s1].tobytes() + aux
        buf += struct.pack("<i", len(body)) + body
    with open(path, "wb") as fh:
        fh.write(bgzf_compress(bytes(buf), level))


_AUX_SIZE = {ord("A"): 1, ord("c"): 1, ord("C"): 1, ord("s"): 2, ord("S"): 2, ord("i"): 4, ord("I"): 4,
             ord("f"): 4, ord("d"): 8}
_AUX_FMT = {ord("c"): "<b", ord("C"): "<B", ord("s"): "<h", ord("S"): "<H", ord("i"): "<i", ord("I"): "<I"}


def _find_nm(data: bytes, off: int, end: int) -> int:
    """Walk the aux block and return the integer value of NM, or -1 if absent."""
    while off + 3 <= end:
        t0, t1, ty = data[off], data[off + 1], data[off + 2]
        off += 3
        if ty in _AUX_SIZE:
            if t0 == 78 and t1 == 77 and ty in _AUX_FMT:      # 'N','M'
                return int(struct.unpack_from(_AUX_FMT[ty], data, off)[0])
            off += _AUX_SIZE[ty]
        elif ty in (ord("Z"), ord("H")):
            off = data.index(b"\0", off) + 1
        elif ty == ord("B"):
            sub = data[off]
            cnt = struct.unpack_from("<I", data, off + 1)[0]
            off += 5 + cnt * _AUX_SIZE[sub]
        else:
            raise ValueError("bad aux type %r" % chr(ty))
    return -1


def read_bam_header(data: bytes):
    if data[:4] != b"BAM\x01":
        raise ValueError("not a BAM file")
    l_text = struct.unpack_from("<i", data, 4)[0]
    text = data[8:8 + l_text].decode(errors="replace")
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", data, off)[0]
    off += 4
    names, lens = [], []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, off)[0]
        names.append(data[off + 4:off + 4 + l_name - 1].decode())
        lens.append(struct.unpack_from("<i", data, off + 4 + l_name)[0])
        off += 8 + l_name
    return text, names, lens, off


def read_bam(path: str):
    """Decode a BAM file into (ref_names, ref_lens, PlainReads) in file order.

    Records with refID < 0 (unplaced) are skipped; bowtie2 --no-unal (reference snps.py:101)
    never writes them and neither pysam.fetch(region) nor getrname could use them.
    """
    with open(path, "rb") as fh:
        data = bgzf_decompress(fh.read())
    _, names, lens, off = read_bam_header(data)
    ref_id, pos, mapq, nm, lq, flag = [], [], [], [], [], []
    coff, soff = [0], [0]
    cig_parts, seq_parts, qual_parts = [], [], []
    n = len(data)
    hi_lut = np.arange(256, dtype=np.uint8) >> 4
    lo_lut = np.arange(256, dtype=np.uint8) & 15
    while off + 4 <= n:
        bs = struct.unpack_from("<i", data, off)[0]
        rec = off + 4
        (rid, p, l_name, mq, _bin, n_cig, fl, l_seq, _nr, _np, _tl) = struct.unpack_from("<iiBBHHHIiii", data, rec)
        o = rec + 32 + l_name
        cig = np.frombuffer(data, "<u4", n_cig, o)
        o += 4 * n_cig
        pk = np.frombuffer(data, np.uint8, (l_seq + 1) // 2, o)
        o += (l_seq + 1) // 2
        q = np.frombuffer(data, np.uint8, l_seq, o)
        o += l_seq
        off = rec + bs
        if rid < 0:
            continue
        codes = np.empty(2 * len(pk), np.uint8)
        codes[0::2] = hi_lut[pk]
        codes[1::2] = lo_lut[pk]
        ref_id.append(rid); pos.append(p); mapq.append(mq); flag.append(fl); lq.append(l_seq)
        nm.append(_find_nm(data, o, off))
        cig_parts.append(cig); seq_parts.append(codes[:l_seq]); qual_parts.append(q)
        coff.append(coff[-1] + n_cig); soff.append(soff[-1] + l_seq)
    cat = lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dt)
    reads = PlainReads(np.array(ref_id, np.int32), np.array(pos, np.int32), np.array(mapq, np.uint8),
                       np.array(nm, np.int32), np.array(lq, np.int32), np.array(coff, np.uint64),
                       cat(cig_parts, np.uint32), np.array(soff, np.uint64), cat(seq_parts, np.uint8),
                       cat(qual_parts, np.uint8), np.array(flag, np.uint16))
    return names, lens, reads
