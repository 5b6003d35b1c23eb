"""GPU parity: gene coverage kernel (through the C-ABI) vs the CPU oracle (reference
midas/run/genes.py:153-189). Counters bit-exact; depth bit-exact with exact_order=True, within
1e-12 relative otherwise (one division of the integer sum instead of the running float sum)."""
import numpy as np
import pytest

from midas_b200 import lib, synth
from midas_b200.soa import Filters
from oracle import oracle

pytestmark = pytest.mark.gpu


def _check(ctx, gref, greads, filters):
    lay = lib.Layout.of(gref)
    pk = lib.PackedReads(greads, lay, "genes")
    exp = oracle.genes_coverage(greads, gref.contig_len, filters)
    for exact in (False, True):
        out = ctx.genes_coverage(pk, gref.contig_len, filters, exact_order=exact)
        assert np.array_equal(out["aligned_reads"], exp["aligned"])
        assert np.array_equal(out["mapped_reads"], exp["mapped"])
        assert np.array_equal(out["mapped_bp"], exp["mapped_bp"])
        if exact:
            assert np.array_equal(out["depth"].view(np.uint64), exp["depth"].view(np.uint64))
        else:
            assert np.allclose(out["depth"], exp["depth"], rtol=1e-12, atol=0)
        # decoded records in, re-layout on the GPU: the same bits
        for r in (greads, greads.nibble_packed()):
            out2 = ctx.genes_coverage_plain(r, gref.contig_len, filters, exact_order=exact)
            assert np.array_equal(out2.view(np.uint8), out.view(np.uint8))


@pytest.mark.parametrize("seed,n_species,gps,n_reads,L", [
    (31, 3, 200, 30_000, 150),
    (32, 10, 50, 60_000, 100),
    (33, 1, 5, 20_000, 250),       # few hot genes: long same-gene runs for the exact-order sum
    (34, 2, 1000, 1_000, 150),     # most genes empty
])
def test_genes_defaults(ctx, seed, n_species, gps, n_reads, L):
    gref, greads, _ = synth.synth_genes(seed=seed, n_species=n_species, genes_per_species=gps, n_reads=n_reads,
                                        read_len=L)
    _check(ctx, gref, greads, Filters.genes_defaults())


def test_genes_filters_and_empty(ctx):
    gref, greads, _ = synth.synth_genes(seed=35, n_species=2, genes_per_species=40, n_reads=8_000)
    _check(ctx, gref, greads, Filters(mapid=97.0, aln_cov=0.9, readq=30, mapq=30, baseq=0))
    _check(ctx, gref, greads, Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=0))
    _check(ctx, gref, greads.select(slice(0, 0)), Filters.genes_defaults())
