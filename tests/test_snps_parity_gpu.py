"""GPU parity: CUDA pileup (through the C-ABI) vs the CPU oracle on the same seeded inputs.
Integer work -> bit-exact (reference midas/run/snps.py:141-216)."""
import numpy as np
import pytest

from midas_b200 import lib, synth
from midas_b200.soa import Filters
from oracle import oracle

pytestmark = pytest.mark.gpu


def _check(ctx, ref, reads, filters):
    lay = lib.Layout.of(ref)
    pk = lib.PackedReads(reads, lay, "snps", sort=True)
    counts, stats = ctx.snps_pileup(pk, lay, filters)
    exp_counts, exp_stats = oracle.snps_pileup(reads, ref, filters)
    got = lay.unpad(counts).astype(np.uint64)
    assert got.shape == exp_counts.shape
    assert np.array_equal(got, exp_counts), f"{int((got != exp_counts).sum())} counters differ"
    for k, name in enumerate(lib.STATS_DTYPE.names):
        assert np.array_equal(stats[name], exp_stats[:, k]), name
    # padding between species must be written as zeros
    mask = np.ones(lay.gpad, bool)
    for c in range(lay.n_contigs):
        o = int(lay.contig_goff[c]); mask[o:o + int(lay.contig_len[c])] = False
    assert not counts[:, mask].any()


@pytest.mark.parametrize("seed,n_species,cps,clen,n_reads,L", [
    (11, 1, 1, 20_000, 30_000, 150),
    (12, 3, 2, 9_000, 40_000, 150),
    (13, 2, 5, 3_000, 20_000, 100),
    (14, 1, 1, 300_000, 400_000, 150),
    (15, 4, 1, 5_000, 1_000, 250),
])
def test_synthetic_default_filters(ctx, seed, n_species, cps, clen, n_reads, L):
    ref, reads = synth.synth_snps(seed=seed, n_species=n_species, contigs_per_species=cps, contig_len=clen,
                                  n_reads=n_reads, read_len=L, len_jitter=0.3, abundance="lognormal")
    _check(ctx, ref, reads, Filters())


@pytest.mark.parametrize("filters", [
    Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=0),
    Filters(mapid=99.0, aln_cov=0.95, readq=36, mapq=40, baseq=41),
    Filters(mapid=94.0, aln_cov=0.75, readq=20, mapq=20, baseq=1),
    Filters(mapid=98.66666666666667, aln_cov=0.9866666666666667, readq=35, mapq=0, baseq=30),
])
def test_filter_settings(ctx, filters):
    ref, reads = synth.synth_snps(seed=21, n_species=2, contigs_per_species=2, contig_len=8_000, n_reads=25_000,
                                  len_jitter=0.2)
    _check(ctx, ref, reads, filters)
