"""CPU oracle for the pooled allele call of merge_midas.py snps (TEST INFRASTRUCTURE).

numpy restatement of GenomicSite's arithmetic in the reference (midas/merge/snps.py):
  compute_pooled_counts :42-47, call_alleles :49-76, compute_per_sample_mafs :78-91,
  compute_prevalence :93-104. `run_reference_sites` executes the reference's unmodified class
(imported from /root/reference over oracle/shim) and is what tests/golden/merge_small.npz was made with.
"""
from __future__ import annotations

import numpy as np

SNP_TYPES = [None, "mono", "bi", "tri", "quad"]


def merge_sites(counts, mean_depth, snp_freq=0.01, site_depth=1, site_ratio=2.0):
    """counts: uint [S][4][G]. Returns dict(pooled u64 [4][G], major/minor int8 [G] (-1 = None),
    snp_type int8 [G] (index into SNP_TYPES), count_samples, sample_depth u64 [S][G], sample_maf f64 [S][G])."""
    counts = np.asarray(counts).astype(np.uint64)
    S, _, G = counts.shape
    pooled = counts.sum(0)                                   # snps.py:42-47
    depth = pooled.sum(0)
    # stable descending sort by frequency == by count (same positive denominator), ties keep ACGT order
    order = np.argsort(-pooled.astype(np.int64), axis=0, kind="stable")
    sorted_counts = np.take_along_axis(pooled, order, 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        freqs = sorted_counts.astype(np.float64) / depth.astype(np.float64)     # float(count)/pooled_depth
    has = depth > 0
    major = np.where(has & (freqs[0] > 0), order[0], -1).astype(np.int8)
    minor = np.where(has & (freqs[1] > 0), order[1], -1).astype(np.int8)
    snp = np.zeros(G, np.int8)
    for t, k in ((1, 0), (2, 1), (3, 2), (4, 3)):            # mono, bi, tri, quad: the rarest allele that reaches snp_freq wins
        snp = np.where(has & (freqs[k] >= snp_freq), t, snp)
    sample_depth = np.zeros((S, G), np.uint64)
    sample_maf = np.zeros((S, G), np.float64)
    g = np.arange(G)
    maj_c = counts[:, np.maximum(major, 0), g]
    min_c = counts[:, np.maximum(minor, 0), g]
    only_major = (major >= 0) & (minor < 0)
    both = (major >= 0) & (minor >= 0)
    sample_depth[:, only_major] = maj_c[:, only_major]
    sd_both = maj_c + min_c
    sample_depth[:, both] = sd_both[:, both]
    with np.errstate(divide="ignore", invalid="ignore"):
        maf = np.where(sd_both > 0, min_c.astype(np.float64) / sd_both.astype(np.float64), 0.0)
    sample_maf[:, both] = maf[:, both]
    md = np.asarray(mean_depth, np.float64)[:, None]
    fail = (sample_depth.astype(np.int64) < site_depth) | (sample_depth.astype(np.float64) / md > site_ratio)
    count_samples = (~fail).sum(0)
    return dict(pooled=pooled, major=major, minor=minor, snp_type=snp, count_samples=count_samples.astype(np.int64),
                sample_depth=sample_depth, sample_maf=sample_maf)


def run_reference_sites(counts, mean_depth, snp_freq=0.01, site_depth=1, site_ratio=2.0):
    """Execute the reference's unmodified GenomicSite per site (build container only)."""
    from oracle.run_reference import _import
    snps = _import("midas.merge.snps")
    counts = np.asarray(counts)
    S, _, G = counts.shape
    out = dict(pooled=np.zeros((4, G), np.uint64), major=np.full(G, -1, np.int8), minor=np.full(G, -1, np.int8),
               snp_type=np.zeros(G, np.int8), count_samples=np.zeros(G, np.int64),
               sample_depth=np.zeros((S, G), np.uint64), sample_maf=np.zeros((S, G), np.float64), freq_text=[])
    for g in range(G):
        values = ["ctg|%d|A" % (g + 1)] + [",".join(str(int(x)) for x in counts[s, :, g]) for s in range(S)]
        site = snps.GenomicSite(g + 1, values)
        site.call_alleles(snp_freq)
        site.compute_per_sample_mafs()
        site.compute_prevalence(list(mean_depth), site_depth, site_ratio)
        out["pooled"][:, g] = site.pooled_counts
        out["major"][g] = "ACGT".index(site.major_allele) if site.major_allele else -1
        out["minor"][g] = "ACGT".index(site.minor_allele) if site.minor_allele else -1
        out["snp_type"][g] = SNP_TYPES.index(site.snp_type)
        out["count_samples"][g] = site.count_samples
        out["sample_depth"][:, g] = site.sample_depths
        out["sample_maf"][:, g] = site.sample_mafs
        out["freq_text"].append("\t".join("{0:.3g}".format(f) for f in site.sample_mafs))
    return out


def synth_counts(seed, S, G, depth_mean=20.0, snp_frac=0.05):
    """SURVEY 8d C5 recipe: Poisson depths x multinomial(major 0.97 / minor 0.03 at a fraction of sites)."""
    rng = np.random.default_rng(seed)
    major = rng.integers(0, 4, G)
    minor = (major + rng.integers(1, 4, G)) % 4
    sample_depth = rng.gamma(2.0, depth_mean / 2.0, S)
    counts = np.zeros((S, 4, G), np.uint32)
    is_snp = rng.random(G) < snp_frac
    for s in range(S):
        d = rng.poisson(sample_depth[s], G)
        d[rng.random(G) < 0.1] = 0                     # uncovered sites
        nminor = np.where(is_snp, rng.binomial(d, rng.choice([0.03, 0.3, 0.5], G)), 0)
        err = rng.binomial(d - nminor, 0.002)          # sequencing errors onto a third allele
        third = (major + 2) % 4
        third = np.where(third == minor, (major + 3) % 4, third)
        g = np.arange(G)
        counts[s, major, g] += (d - nminor - err).astype(np.uint32)
        counts[s, minor, g] += nminor.astype(np.uint32)
        counts[s, third, g] += err.astype(np.uint32)
    mean_depth = np.maximum(counts.sum(1).mean(1), 0.5)
    return counts, mean_depth
