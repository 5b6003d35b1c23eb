"""Seeded synthetic aligned reads of the shapes BASELINE.json names.

Everything is drawn from numpy.random.Generator(PCG64) streams derived from one
seed with SeedSequence, chunk by chunk, so the result does not depend on the
number of worker threads. The output is `PlainReads` (decoded-BAM-like records),
i.e. what pysam would hand to the reference loops; packing for the GPU is a
separate, tested step (mb_pack_reads).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .soa import PlainReads, RefSet, OP_M, OP_I, OP_D, OP_S

_NIB = np.array([1, 2, 4, 8], np.uint8)          # code 0..3 -> BAM nibble
_ASCII = np.frombuffer(b"ACGT", np.uint8)

SEED_C2 = 20260102
SEED_C3 = 20260103
SEED_C4 = 20260104


def _rng(seed, *path):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence([int(seed), *map(int, path)])))


def _weights(kind: str, n: int, rng) -> np.ndarray:
    if kind == "uniform":
        w = np.ones(n)
    elif kind == "lognormal":
        w = rng.lognormal(0.0, 1.0, n)
    else:
        raise ValueError(kind)
    return w / w.sum()


def _mapq(rng, n):
    u = rng.random(n)
    hi = rng.integers(24, 41, n)
    lo = rng.integers(0, 20, n)
    return np.where(u < 0.85, 42, np.where(u < 0.92, hi, lo)).astype(np.uint8)


def _qual(rng, n, L):
    q = rng.standard_normal((n, L), dtype=np.float32)
    q *= 5.0
    q += 36.0
    np.rint(q, out=q)
    q[:, max(0, L - 20):] -= 6.0
    np.clip(q, 2, 41, out=q)
    return q.astype(np.uint8)


def _snps_chunk(seed, ci, ref_codes, gbase, clen, ref_id, pos, L):
    """Generate one chunk of reads given their (sorted) placement."""
    rng = _rng(seed, 2, ci)
    n = len(pos)
    kind = rng.choice(4, size=n, p=[0.90, 0.04, 0.04, 0.02]).astype(np.int8)
    k = rng.integers(1, 4, n).astype(np.int32)                 # indel size 1..3
    a = rng.integers(10, L - 13, n).astype(np.int32)           # indel breakpoint
    s = rng.integers(1, 21, n).astype(np.int32)                # soft clips 1..20
    e = rng.integers(1, 21, n).astype(np.int32)
    side = rng.integers(0, 3, n)                               # 0 left, 1 right, 2 both
    s = np.where(side == 1, 0, s)
    e = np.where(side == 0, 0, e)
    q = np.arange(L, dtype=np.int32)[None, :]
    kind_ = kind[:, None]
    a_, k_, s_, e_ = a[:, None], k[:, None], s[:, None], e[:, None]
    is_ins, is_del, is_sc = kind_ == 1, kind_ == 2, kind_ == 3
    shift = np.where(is_ins & (q >= a_ + k_), -k_, 0) + np.where(is_del & (q >= a_), k_, 0) \
        + np.where(is_sc, -s_, 0)
    unaligned = (is_ins & (q >= a_) & (q < a_ + k_)) | (is_sc & ((q < s_) | (q >= L - e_)))
    ridx = gbase[ref_id][:, None] + pos[:, None].astype(np.int64) + q + shift
    lo = gbase[ref_id][:, None]
    hi = lo + clen[ref_id][:, None] - 1
    np.clip(ridx, lo, hi, out=ridx)
    b = ref_codes[ridx]
    del ridx
    nun = int(unaligned.sum())
    b[unaligned] = rng.integers(0, 4, nun, dtype=np.uint8)
    v = rng.integers(0, 65536, (n, L), dtype=np.uint16)
    sub = (v < 655) & ~unaligned                               # ~1 % substitutions
    isn = (v >= 655) & (v < 786)                               # ~0.2 % N
    nsub = int(sub.sum())
    b[sub] = (b[sub] + rng.integers(1, 4, nsub, dtype=np.uint8)) & 3
    nm = sub.sum(1).astype(np.int32) + (isn & ~unaligned).sum(1).astype(np.int32) \
        + np.where((kind == 1) | (kind == 2), k, 0)
    seq = _NIB[b]
    seq[isn] = 15
    qual = _qual(rng, n, L)
    mapq = _mapq(rng, n)
    # CIGARs
    nops = np.where(kind == 0, 1, np.where(kind == 3, 1 + (s > 0) + (e > 0), 3)).astype(np.int64)
    coff = np.zeros(n + 1, np.int64)
    np.cumsum(nops, out=coff[1:])
    cig = np.zeros(int(coff[-1]), np.uint32)
    c0 = coff[:-1]
    m0 = kind == 0
    cig[c0[m0]] = (L << 4) | OP_M
    m1 = kind == 1
    cig[c0[m1]] = (a[m1] << 4) | OP_M
    cig[c0[m1] + 1] = (k[m1] << 4) | OP_I
    cig[c0[m1] + 2] = ((L - a[m1] - k[m1]) << 4) | OP_M
    m2 = kind == 2
    cig[c0[m2]] = (a[m2] << 4) | OP_M
    cig[c0[m2] + 1] = (k[m2] << 4) | OP_D
    cig[c0[m2] + 2] = ((L - a[m2]) << 4) | OP_M
    m3 = kind == 3
    has_s = m3 & (s > 0)
    cig[c0[has_s]] = (s[has_s] << 4) | OP_S
    midx = c0 + (s > 0)
    cig[midx[m3]] = ((L - s[m3] - e[m3]) << 4) | OP_M
    has_e = m3 & (e > 0)
    cig[midx[has_e] + 1] = (e[has_e] << 4) | OP_S
    return dict(mapq=mapq, nm=nm, nops=nops, cigar=cig, seq=seq.reshape(-1), qual=qual.reshape(-1))


def synth_snps(seed=SEED_C2, n_species=1, contigs_per_species=1, contig_len=4_000_000,
               n_reads=5_000_000, read_len=150, abundance="uniform", len_jitter=0.0,
               chunk=200_000, threads=None, keep_ref_seq=True):
    """SURVEY §8d C2/C4 recipe: coordinate-sorted reads over `n_species` reference genomes.

    Returns (RefSet, PlainReads). CIGAR mix 90 % `LM`, 4 % one 1-3 bp insertion, 4 % one
    1-3 bp deletion, 2 % soft-clipped ends (1-20 bp); 1 % substitutions, 0.2 % N;
    qual = clip(round(N(36,5)) [-6 on the last 20 cycles], 2, 41); mapq 42 (85 %),
    24-40 (7 %), 0-19 (8 %); NM = mismatches + inserted + deleted bases.
    """
    L = int(read_len)
    rng = _rng(seed, 0)
    n_contigs = n_species * contigs_per_species
    if len_jitter > 0:
        clen = np.maximum(L + 8, (contig_len * (1.0 + len_jitter * (rng.random(n_contigs) * 2 - 1))).astype(np.int64))
    else:
        clen = np.full(n_contigs, int(contig_len), np.int64)
    gbase = np.zeros(n_contigs + 1, np.int64)
    np.cumsum(clen, out=gbase[1:])
    ref_codes = rng.integers(0, 4, int(gbase[-1]), dtype=np.uint8)
    contig_species = np.repeat(np.arange(n_species, dtype=np.int32), contigs_per_species)
    sp_w = _weights(abundance, n_species, rng)
    cw = sp_w[contig_species] * clen / np.bincount(contig_species, weights=clen, minlength=n_species)[contig_species]
    per_contig = rng.multinomial(n_reads, cw / cw.sum())
    ref_id = np.repeat(np.arange(n_contigs, dtype=np.int32), per_contig)
    u = rng.random(n_reads)
    # sort within contig: contigs are already grouped, so sort (ref_id, u)
    order = np.lexsort((u, ref_id))
    u = u[order]
    pos = np.floor(u * (clen[ref_id] - (L + 3))).astype(np.int32)
    gb = gbase[:-1]

    bounds = list(range(0, n_reads, chunk)) + [n_reads]
    jobs = [(ci, bounds[ci], bounds[ci + 1]) for ci in range(len(bounds) - 1)]
    threads = threads or min(16, os.cpu_count() or 1)

    def run(job):
        ci, lo, hi = job
        return _snps_chunk(seed, ci, ref_codes, gb, clen, ref_id[lo:hi], pos[lo:hi], L)

    if threads > 1 and len(jobs) > 1:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(run, jobs))
    else:
        parts = [run(j) for j in jobs]
    nops = np.concatenate([p["nops"] for p in parts]) if parts else np.zeros(0, np.int64)
    coff = np.zeros(n_reads + 1, np.uint64)
    np.cumsum(nops, out=coff[1:])
    soff = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(L))
    cat = lambda k, dt: (np.concatenate([p[k] for p in parts]) if parts else np.zeros(0, dt))
    reads = PlainReads(ref_id, pos, cat("mapq", np.uint8), cat("nm", np.int32),
                       np.full(n_reads, L, np.int32), coff, cat("cigar", np.uint32), soff,
                       cat("seq", np.uint8), cat("qual", np.uint8),
                       np.zeros(n_reads, np.uint16))
    names, seqs = [], []
    species_ids = [f"Species_{i:03d}_{60000 + i}" for i in range(n_species)]
    for c in range(n_contigs):
        names.append(f"{species_ids[contig_species[c]]}.c{c % contigs_per_species:04d}")
        if keep_ref_seq:
            seqs.append(_ASCII[ref_codes[gbase[c]:gbase[c + 1]]].tobytes())
    ref = RefSet(names, clen.astype(np.uint32), contig_species, species_ids, seqs if keep_ref_seq else None)
    return ref, reads


def genes_model(seed, n_species, genes_per_species, read_len, abundance="lognormal", n_markers=15):
    """The pangenome of the C3 recipe: (gene_len int64 [n_genes], read weight per gene, gene_species, marker index
    or -1). Gene length = clip(lognormal(ln 900, 0.6), 150, 20000); species abundance lognormal(sigma=1);
    per-gene copy-number variation inside a species, 15 % of the genes absent."""
    L = int(read_len)
    rng = _rng(seed, 0)
    n_genes = n_species * genes_per_species
    glen = np.clip(rng.lognormal(np.log(900.0), 0.6, n_genes), max(150, L + 1), 20000).astype(np.int64)
    gene_species = np.repeat(np.arange(n_species, dtype=np.int32), genes_per_species)
    sp_w = _weights(abundance, n_species, rng)
    gw = rng.lognormal(0, 0.5, n_genes) * (rng.random(n_genes) > 0.15) * glen
    gw = sp_w[gene_species] * gw / np.bincount(gene_species, weights=gw, minlength=n_species)[gene_species]
    marker = np.full(n_genes, -1, np.int32)
    for s in range(n_species):
        pick = rng.choice(genes_per_species, size=min(n_markers, genes_per_species), replace=False)
        marker[s * genes_per_species + pick] = np.arange(len(pick), dtype=np.int32)
    return glen, gw, gene_species, marker


def synth_genes(seed=SEED_C3, n_species=100, genes_per_species=20_000, n_reads=50_000_000,
                read_len=150, abundance="lognormal", chunk=500_000, threads=None, n_markers=15):
    """SURVEY §8d C3 recipe: reads in random (file) order over a multi-species pangenome.

    Gene length = clip(lognormal(ln 900, 0.6), 150, 20000); 10 % of reads soft-clipped
    (local mode); NM ~ Binomial(L, 0.012) with a 3 % tail of divergent reads; species
    abundance lognormal(sigma=1). Returns (RefSet with genes as "contigs", PlainReads,
    marker_ids: int32 [n_genes] marker index or -1).
    """
    L = int(read_len)
    n_genes = n_species * genes_per_species
    glen, gw, gene_species, marker = genes_model(seed, n_species, genes_per_species, L, abundance, n_markers)
    cdf = np.cumsum(gw / gw.sum())

    bounds = list(range(0, n_reads, chunk)) + [n_reads]
    jobs = [(ci, bounds[ci], bounds[ci + 1]) for ci in range(len(bounds) - 1)]
    threads = threads or min(16, os.cpu_count() or 1)

    def run(job):
        ci, lo, hi = job
        r = _rng(seed, 3, ci)
        n = hi - lo
        g = np.minimum(np.searchsorted(cdf, r.random(n)), n_genes - 1).astype(np.int32)
        sc = r.random(n) < 0.10
        s = np.where(sc, r.integers(0, 41, n), 0).astype(np.int32)
        e = np.where(sc, r.integers(0, 41, n), 0).astype(np.int32)
        m = L - s - e
        pos = np.floor(r.random(n) * (glen[g] - m + 1)).astype(np.int32)
        nm = r.binomial(L, 0.012, n).astype(np.int32)
        div = r.random(n) < 0.03
        nm = np.where(div, nm + r.integers(6, 20, n), nm).astype(np.int32)
        nops = (1 + (s > 0) + (e > 0)).astype(np.int64)
        coff = np.zeros(n + 1, np.int64)
        np.cumsum(nops, out=coff[1:])
        cig = np.zeros(int(coff[-1]), np.uint32)
        c0 = coff[:-1]
        hs, he = s > 0, e > 0
        cig[c0[hs]] = (s[hs] << 4) | OP_S
        mi = c0 + hs
        cig[mi] = (m << 4) | OP_M
        cig[mi[he] + 1] = (e[he] << 4) | OP_S
        seq = _NIB[r.integers(0, 4, (n, L), dtype=np.uint8)]
        seq[r.integers(0, 65536, (n, L), dtype=np.uint16) < 131] = 15
        return dict(ref_id=g, pos=pos, nm=nm, nops=nops, cigar=cig, mapq=_mapq(r, n),
                    seq=seq.reshape(-1), qual=_qual(r, n, L).reshape(-1))

    if threads > 1 and len(jobs) > 1:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(run, jobs))
    else:
        parts = [run(j) for j in jobs]
    cat = lambda k, dt: (np.concatenate([p[k] for p in parts]) if parts else np.zeros(0, dt))
    nops = cat("nops", np.int64)
    coff = np.zeros(n_reads + 1, np.uint64)
    np.cumsum(nops, out=coff[1:])
    soff = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(L))
    reads = PlainReads(cat("ref_id", np.int32), cat("pos", np.int32), cat("mapq", np.uint8), cat("nm", np.int32),
                       np.full(n_reads, L, np.int32), coff, cat("cigar", np.uint32), soff,
                       cat("seq", np.uint8), cat("qual", np.uint8), np.zeros(n_reads, np.uint16))
    species_ids = [f"Species_{i:03d}_{60000 + i}" for i in range(n_species)]
    names = [f"{species_ids[gene_species[g]]}.g{g % genes_per_species:06d}" for g in range(n_genes)]
    ref = RefSet(names, glen.astype(np.uint32), gene_species, species_ids, None)
    return ref, reads, marker
