"""Host-side containers for decoded alignments and reference layout.

`PlainReads` holds exactly the per-record fields the reference's loops read from
pysam.AlignedSegment (reference midas/run/snps.py:145-157, midas/run/genes.py:154-166,180)
as a numpy struct-of-arrays; it is the input of `mb_pack_reads` (include/midas_b200.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

# BAM CIGAR op codes (MIDNSHP=X)
OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X = range(9)
CIGAR_CHARS = "MIDNSHP=X"
# BAM 4-bit base codes "=ACMGRSVTWYHKDBN"
BAM_NIBBLE_CHARS = "=ACMGRSVTWYHKDBN"
NIB_A, NIB_C, NIB_G, NIB_T, NIB_N = 1, 2, 4, 8, 15


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


@dataclass
class PlainReads:
    """Decoded BAM records, struct of arrays (mirrors C struct mb_plain_reads)."""
    ref_id: np.ndarray      # int32 [n]
    pos: np.ndarray         # int32 [n] 0-based leftmost
    mapq: np.ndarray        # uint8 [n]
    nm: np.ndarray          # int32 [n]  (<0 = NM tag missing)
    l_qseq: np.ndarray      # int32 [n]
    cigar_off: np.ndarray   # uint64 [n+1]
    cigar: np.ndarray       # uint32 [sum n_cigar]  len<<4|op
    seq_off: np.ndarray     # uint64 [n+1] element offsets into seq / qual (span >= l_qseq: padding allowed)
    seq: np.ndarray         # uint8: BAM 4-bit codes, one per byte, or two per byte (seq_nibbles)
    qual: np.ndarray        # uint8 [total elements] raw Phred
    flag: np.ndarray | None = None   # uint16 [n] SAM flag (carried for BAM round trips only)
    seq_nibbles: bool = False        # seq as in a BAM record: element e in byte e >> 1, even e in the high nibble

    def __post_init__(self):
        self.ref_id = _c(self.ref_id, np.int32)
        self.pos = _c(self.pos, np.int32)
        self.mapq = _c(self.mapq, np.uint8)
        self.nm = _c(self.nm, np.int32)
        self.l_qseq = _c(self.l_qseq, np.int32)
        self.cigar_off = _c(self.cigar_off, np.uint64)
        self.cigar = _c(self.cigar, np.uint32)
        self.seq_off = _c(self.seq_off, np.uint64)
        self.seq = _c(self.seq, np.uint8)
        self.qual = _c(self.qual, np.uint8)
        n = self.n_reads
        assert self.pos.shape == (n,) and self.mapq.shape == (n,) and self.nm.shape == (n,)
        assert self.l_qseq.shape == (n,) and self.cigar_off.shape == (n + 1,) and self.seq_off.shape == (n + 1,)
        if self.seq_nibbles:
            assert self.seq.shape[0] == (self.qual.shape[0] + 1) // 2
        else:
            assert self.seq.shape == self.qual.shape

    def unpacked(self) -> "PlainReads":
        """One code per byte, spans exactly l_qseq (the form the oracle and the BAM writer take)."""
        if not self.seq_nibbles:
            if getattr(self, "_exact_spans", None) is None:     # checked once: the arrays are not resized in place
                self._exact_spans = bool(np.array_equal((self.seq_off[1:] - self.seq_off[:-1]).astype(np.int64),
                                                        self.l_qseq.astype(np.int64)))
            if self._exact_spans:
                return self
        n = self.n_reads
        lq = self.l_qseq.astype(np.int64)
        soff = np.zeros(n + 1, np.uint64)
        np.cumsum(lq, out=soff[1:])
        total = int(soff[-1])
        if total:
            src = np.repeat(self.seq_off[:-1].astype(np.int64) - soff[:-1].astype(np.int64), lq) + np.arange(total, dtype=np.int64)
            if self.seq_nibbles:
                b = self.seq[src >> 1]
                seq = np.where(src & 1, b & 15, b >> 4).astype(np.uint8)
            else:
                seq = self.seq[src]
            qual = self.qual[src]
        else:
            seq, qual = np.zeros(0, np.uint8), np.zeros(0, np.uint8)
        return PlainReads(self.ref_id, self.pos, self.mapq, self.nm, self.l_qseq, self.cigar_off, self.cigar, soff,
                          seq, qual, self.flag)

    def nibble_packed(self) -> "PlainReads":
        """seq as in a BAM record (two codes per byte, every read on a byte boundary; qual padded alike)."""
        if self.seq_nibbles:
            return self
        u = self.unpacked()
        n = u.n_reads
        lq = u.l_qseq.astype(np.int64)
        span = (lq + 1) & ~1
        soff = np.zeros(n + 1, np.uint64)
        np.cumsum(span, out=soff[1:])
        total = int(soff[-1])
        codes = np.zeros(total, np.uint8)
        qual = np.zeros(total, np.uint8)
        if total:
            dst = np.repeat(soff[:-1].astype(np.int64) - u.seq_off[:-1].astype(np.int64), lq) + np.arange(int(lq.sum()), dtype=np.int64)
            codes[dst] = u.seq & 15
            qual[dst] = u.qual
        seq = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)
        return PlainReads(u.ref_id, u.pos, u.mapq, u.nm, u.l_qseq, u.cigar_off, u.cigar, soff, seq, qual, u.flag, True)

    @property
    def n_reads(self) -> int:
        return int(self.ref_id.shape[0])

    @property
    def total_bases(self) -> int:
        return int(self.l_qseq.sum(dtype=np.int64))

    def select(self, idx) -> "PlainReads":
        """Sub-select / reorder reads (idx: index array or slice)."""
        if self.seq_nibbles:
            return self.unpacked().select(idx).nibble_packed()
        idx = np.arange(self.n_reads)[idx] if isinstance(idx, slice) else np.asarray(idx, dtype=np.int64)
        ncig = (self.cigar_off[1:] - self.cigar_off[:-1]).astype(np.int64)[idx]
        nseq = (self.seq_off[1:] - self.seq_off[:-1]).astype(np.int64)[idx]
        coff = np.zeros(len(idx) + 1, np.uint64)
        soff = np.zeros(len(idx) + 1, np.uint64)
        np.cumsum(ncig, out=coff[1:])
        np.cumsum(nseq, out=soff[1:])

        def gather(src, starts, lens, total):
            if total == 0:
                return src[:0].copy()
            rep = np.repeat(starts.astype(np.int64) - np.concatenate(([0], np.cumsum(lens)[:-1])), lens)
            return src[rep + np.arange(total, dtype=np.int64)]

        cig = gather(self.cigar, self.cigar_off[:-1][idx], ncig, int(coff[-1]))
        seq = gather(self.seq, self.seq_off[:-1][idx], nseq, int(soff[-1]))
        qual = gather(self.qual, self.seq_off[:-1][idx], nseq, int(soff[-1]))
        return PlainReads(self.ref_id[idx], self.pos[idx], self.mapq[idx], self.nm[idx], self.l_qseq[idx],
                          coff, cig, soff, seq, qual, None if self.flag is None else self.flag[idx])

    @staticmethod
    def concat(parts: list["PlainReads"]) -> "PlainReads":
        if any(p.seq_nibbles for p in parts):
            parts = [p.unpacked() for p in parts]
        coffs, soffs = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)]
        cb = sb = 0
        for p in parts:
            coffs.append(p.cigar_off[1:] + np.uint64(cb))
            soffs.append(p.seq_off[1:] + np.uint64(sb))
            cb += int(p.cigar_off[-1])
            sb += int(p.seq_off[-1])
        cat = np.concatenate
        flag = None if any(p.flag is None for p in parts) else cat([p.flag for p in parts])
        return PlainReads(cat([p.ref_id for p in parts]), cat([p.pos for p in parts]),
                          cat([p.mapq for p in parts]), cat([p.nm for p in parts]),
                          cat([p.l_qseq for p in parts]), cat(coffs), cat([p.cigar for p in parts]),
                          cat(soffs), cat([p.seq for p in parts]), cat([p.qual for p in parts]), flag)

    @staticmethod
    def from_records(records) -> "PlainReads":
        """records: iterable of dicts with ref_id,pos,mapq,nm,cigar [(op,len)...],seq (str or codes),qual (list/bytes)."""
        ref_id, pos, mapq, nm, lq, flag = [], [], [], [], [], []
        coff, soff, cig, seq, qual = [0], [0], [], [], []
        lut = {c: i for i, c in enumerate(BAM_NIBBLE_CHARS)}
        for r in records:
            ref_id.append(r["ref_id"]); pos.append(r["pos"]); mapq.append(r["mapq"]); nm.append(r.get("nm", 0))
            flag.append(r.get("flag", 0))
            s = r["seq"]
            codes = [lut[ch] for ch in s] if isinstance(s, str) else list(s)
            q = list(r["qual"])
            assert len(q) == len(codes)
            lq.append(len(codes))
            seq += codes; qual += q
            for op, ln in r["cigar"]:
                cig.append((ln << 4) | op)
            coff.append(len(cig)); soff.append(len(seq))
        return PlainReads(np.array(ref_id, np.int32), np.array(pos, np.int32), np.array(mapq, np.uint8),
                          np.array(nm, np.int32), np.array(lq, np.int32), np.array(coff, np.uint64),
                          np.array(cig, np.uint32), np.array(soff, np.uint64), np.array(seq, np.uint8),
                          np.array(qual, np.uint8), np.array(flag, np.uint16))


@dataclass
class RefSet:
    """Reference contigs of the selected species (reference snps.py:55-67 initialize_contigs),
    or the genes of the selected pangenomes (genes.py:61-82) when used for the genes path."""
    contig_names: list[str]
    contig_len: np.ndarray          # uint32 [n_contigs]
    contig_species: np.ndarray      # int32 [n_contigs] -> index into species_ids
    species_ids: list[str]
    contig_seq: list[bytes] | None = None   # upper-cased reference bases per contig (for ref_allele)

    def __post_init__(self):
        self.contig_len = _c(self.contig_len, np.uint32)
        self.contig_species = _c(self.contig_species, np.int32)

    @property
    def n_contigs(self) -> int:
        return len(self.contig_names)

    @property
    def n_species(self) -> int:
        return len(self.species_ids)


@dataclass
class Filters:
    """args['mapid','aln_cov','readq','mapq','baseq'] (reference scripts/run_midas.py:410-419)."""
    mapid: float = 94.0
    aln_cov: float = 0.75
    readq: int = 20
    mapq: int = 20
    baseq: int = 30

    @staticmethod
    def genes_defaults() -> "Filters":
        # scripts/run_midas.py:277-284 (mapq default 0; baseq unused by the genes path)
        return Filters(94.0, 0.75, 20, 0, 0)


class CFilters(C.Structure):
    _fields_ = [("mapid", C.c_double), ("aln_cov", C.c_double), ("readq", C.c_int32),
                ("mapq", C.c_int32), ("baseq", C.c_int32), ("reserved", C.c_int32)]

    @staticmethod
    def of(f: Filters) -> "CFilters":
        return CFilters(float(f.mapid), float(f.aln_cov), int(f.readq), int(f.mapq), int(f.baseq), 0)


def cigar_tuples(reads: PlainReads, i: int):
    a, b = int(reads.cigar_off[i]), int(reads.cigar_off[i + 1])
    return [(int(w) & 0xF, int(w) >> 4) for w in reads.cigar[a:b]]


def cigar_string(reads: PlainReads, i: int) -> str:
    return "".join(f"{ln}{CIGAR_CHARS[op]}" for op, ln in cigar_tuples(reads, i)) or "*"
