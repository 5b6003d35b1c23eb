"""ctypes wrapper of oracle/pileup_oracle.c plus a second, independent pure-Python
restatement for small cases. TEST INFRASTRUCTURE — see the C file's header.

Reference lines restated (relative to the reference tree):
  snps:  midas/run/snps.py:141-162 keep_read, :164-216 species_pileup, :219-244 pysam_pileup
  genes: midas/run/genes.py:153-169 keep_read, :171-203 count_mapped_bp
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OFilters(C.Structure):
    _fields_ = [("mapid", C.c_double), ("aln_cov", C.c_double), ("readq", C.c_int32),
                ("mapq", C.c_int32), ("baseq", C.c_int32), ("reserved", C.c_int32)]


class OracleError(Exception):
    """Input on which the reference itself raises (rc -1) or that is malformed (rc -5)."""
    def __init__(self, rc):
        super().__init__(f"oracle rc={rc}")
        self.rc = rc


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pileup_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_snps_pileup.restype = C.c_int
        _LIB.oracle_genes_coverage.restype = C.c_int
        _LIB.oracle_snps_finalize.restype = None
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _of(f):
    return OFilters(float(f.mapid), float(f.aln_cov), int(f.readq), int(f.mapq), int(f.baseq), 0)


def snps_pileup(reads, ref, filters, threads=1, species_ranges=None):
    """Returns (counts uint64 [4][g_total] in the given contig order, stats uint64 [n_species][5]).

    threads > 1 shards reads over a thread pool in contiguous ranges (the C call releases the
    GIL); with `species_ranges` [(lo, hi), ...] the shards are the reference's own unit of
    parallelism, one task per species (midas/run/snps.py:225-228).
    """
    L = lib()
    reads = reads.unpacked()       # the C restatement reads one base code per byte
    g_total = int(ref.contig_len.sum(dtype=np.int64))
    n_sp = ref.n_species
    f = _of(filters)
    n = reads.n_reads

    def call(lo, hi, counts, stats, zero):
        rc = L.oracle_snps_pileup(C.c_uint64(lo), C.c_uint64(hi), _p(reads.ref_id), _p(reads.pos), _p(reads.mapq),
                                  _p(reads.nm), _p(reads.l_qseq), _p(reads.cigar_off), _p(reads.cigar),
                                  _p(reads.seq_off), _p(reads.seq), _p(reads.qual),
                                  C.c_uint32(ref.n_contigs), _p(ref.contig_len), _p(ref.contig_species),
                                  C.c_uint32(n_sp), C.byref(f), _p(counts), _p(stats), C.c_int(zero))
        if rc != 0:
            raise OracleError(rc)

    counts = np.zeros((4, g_total), np.uint64)
    stats = np.zeros((n_sp, 5), np.uint64)
    if species_ranges is None and threads <= 1:
        call(0, n, counts, stats, 0)
    else:
        if species_ranges is None:
            b = np.linspace(0, n, threads + 1).astype(np.int64)
            species_ranges = [(int(b[i]), int(b[i + 1])) for i in range(threads)]
            private = True      # position ranges overlap between shards: private planes, then sum
        else:
            private = False     # species own disjoint counter ranges (reads sorted by species)
        def task(rg):
            if private:
                c = np.zeros((4, g_total), np.uint64); s = np.zeros((n_sp, 5), np.uint64)
                call(rg[0], rg[1], c, s, 0)
                return c, s
            s = np.zeros((n_sp, 5), np.uint64)
            call(rg[0], rg[1], counts, s, 0)
            return None, s
        with ThreadPoolExecutor(max(1, threads)) as ex:
            for c, s in ex.map(task, species_ranges):
                if c is not None:
                    counts += c
                stats += s
    L.oracle_snps_finalize(C.c_uint32(ref.n_contigs), _p(ref.contig_len), _p(ref.contig_species),
                           C.c_uint32(n_sp), _p(counts), _p(stats))
    return counts, stats


def genes_coverage(reads, gene_len, filters):
    """Returns dict(aligned u64[n_genes], mapped u64[n_genes], mapped_bp u64[n_genes], depth f64[n_genes])."""
    L = lib()
    reads = reads.unpacked()
    gene_len = np.ascontiguousarray(gene_len, np.uint32)
    ng = len(gene_len)
    al = np.zeros(ng, np.uint64); mp = np.zeros(ng, np.uint64); bp = np.zeros(ng, np.uint64)
    dp = np.zeros(ng, np.float64)
    f = _of(filters)
    rc = L.oracle_genes_coverage(C.c_uint64(reads.n_reads), _p(reads.ref_id), _p(reads.mapq), _p(reads.nm),
                                 _p(reads.l_qseq), _p(reads.cigar_off), _p(reads.cigar), _p(reads.seq_off),
                                 _p(reads.qual), C.c_uint32(ng), _p(gene_len), C.byref(f),
                                 _p(al), _p(mp), _p(bp), _p(dp))
    if rc != 0:
        raise OracleError(rc)
    return dict(aligned=al, mapped=mp, mapped_bp=bp, depth=dp)


# ---------------------------------------------------------------------------
# Independent pure-Python restatement (small cases only). Written against the
# pysam semantics list in SURVEY.md §8c, not against the C file above.
# ---------------------------------------------------------------------------
_NIB2IDX = {1: 0, 2: 1, 4: 2, 8: 3}


def py_keep_read(cig, l_qseq, nm, mapq, qual, f):
    """cig: [(op, len)]. Mirrors keep_read (snps.py:141-162) incl. exceptions as OracleError(-1)."""
    if nm < 0 or l_qseq <= 0 or (len(qual) and qual[0] == 0xFF):
        raise OracleError(-1)
    start, end = 0, l_qseq
    for op, ln in cig:
        if op == 5:
            continue
        if op == 4:
            start += ln
        else:
            break
    for k in range(len(cig) - 1, 0, -1):
        op, ln = cig[k]
        if op == 5:
            continue
        if op == 4:
            end -= ln
        else:
            break
    alen = max(0, end - start)
    if alen == 0:
        raise OracleError(-1)
    if 100 * (alen - nm) / float(alen) < f.mapid:
        return False, alen
    if np.mean(np.asarray(qual, dtype=np.uint8)) < f.readq:
        return False, alen
    if mapq < f.mapq:
        return False, alen
    if alen / float(l_qseq) < f.aln_cov:
        return False, alen
    return True, alen


def py_snps_pileup(reads, ref, filters):
    reads = reads.unpacked()
    g_total = int(ref.contig_len.sum(dtype=np.int64))
    coff = np.concatenate(([0], np.cumsum(ref.contig_len.astype(np.int64))))
    counts = np.zeros((4, g_total), np.uint64)
    stats = np.zeros((ref.n_species, 5), np.uint64)
    for r in range(reads.n_reads):
        c = int(reads.ref_id[r]); sp = int(ref.contig_species[c])
        a, b = int(reads.cigar_off[r]), int(reads.cigar_off[r + 1])
        cig = [(int(w) & 0xF, int(w) >> 4) for w in reads.cigar[a:b]]
        if any(op == 6 or op > 8 for op, _ in cig):
            raise OracleError(-5)
        s0, s1 = int(reads.seq_off[r]), int(reads.seq_off[r + 1])
        lq = int(reads.l_qseq[r])
        if cig and sum(ln for op, ln in cig if op in (0, 1, 4, 7, 8)) != lq:
            raise OracleError(-5)
        q = reads.qual[s0:s1]; s = reads.seq[s0:s1]
        stats[sp, 3] += 1
        keep, _ = py_keep_read(cig, lq, int(reads.nm[r]), int(reads.mapq[r]), q, filters)
        if not keep:
            continue
        stats[sp, 4] += 1
        qpos, rpos = 0, int(reads.pos[r])
        clen = int(ref.contig_len[c])
        for op, ln in cig:
            if op in (0, 7, 8):
                for _ in range(ln):
                    if 0 <= rpos < clen and (filters.baseq == 0 or q[qpos] >= filters.baseq):
                        k = _NIB2IDX.get(int(s[qpos]))
                        if k is not None:
                            counts[k, coff[c] + rpos] += 1
                    qpos += 1; rpos += 1
            elif op in (1, 4):
                qpos += ln
            elif op in (2, 3):
                rpos += ln
    depth = counts.sum(0)
    for c in range(ref.n_contigs):
        sp = int(ref.contig_species[c])
        d = depth[coff[c]:coff[c + 1]]
        stats[sp, 0] += len(d); stats[sp, 1] += int((d > 0).sum()); stats[sp, 2] += int(d.sum())
    return counts, stats
