"""GPU parity: CUDA pileup (through the C-ABI) vs the CPU oracle on the same seeded inputs.
Integer work -> bit-exact (reference midas/run/snps.py:141-216)."""
import numpy as np
import pytest

from midas_b200 import lib, synth
from midas_b200.soa import Filters
from oracle import oracle

import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

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
    # the host-buffer call of the product path: decoded records in, re-layout on the GPU (k_pack_*)
    key = (reads.ref_id.astype(np.int64) << 32) | reads.pos.astype(np.int64)
    if bool(np.all(np.diff(key) >= 0)):                    # BAM coordinate order
        for r in (reads, reads.nibble_packed()):
            c2, s2 = ctx.snps_pileup_plain(r, lay, filters)
            assert np.array_equal(c2, counts), "plain-record path differs from the host-packed path"
            assert np.array_equal(s2, stats)
    else:
        with pytest.raises(lib.MidasB200Error) as e:
            ctx.snps_pileup_plain(reads, lay, filters)
        assert e.value.status == -6


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
    Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=64),   # above every Phred a 6-plane batch can hold
    Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=37),
])
def test_filter_settings(ctx, filters):
    ref, reads = synth.synth_snps(seed=21, n_species=2, contigs_per_species=2, contig_len=8_000, n_reads=25_000,
                                  len_jitter=0.2)
    _check(ctx, ref, reads, filters)


# ---- arbitrary CIGARs (clips, =/X, N skips, long deletions -> long-read kernel, contig-end overhang)
@pytest.mark.parametrize("seed,n_species,cps,lo,hi,n_reads,max_len", [
    (101, 3, 2, 500, 6000, 3000, 300),
    (102, 1, 1, 5000, 5000, 4000, 150),
    (103, 2, 3, 40, 900, 2500, 120),        # contigs shorter than reads
    (104, 1, 1, 70000, 70000, 20000, 100),  # many tiles
    (105, 4, 1, 3000, 9000, 6000, 520),     # query longer than a 16-word row
])
def test_random_cigars(ctx, seed, n_species, cps, lo, hi, n_reads, max_len):
    import gen
    rng = np.random.default_rng(seed)
    ref = gen.make_ref(rng, n_species, cps, lo, hi)
    reads = gen.random_reads(seed, ref, n_reads, max_len=max_len)
    for filters in (Filters(), Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=0),
                    Filters(mapid=90.0, aln_cov=0.5, readq=25, mapq=10, baseq=20),
                    Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=70)):   # 7 Phred planes, top bit set
        _check(ctx, ref, reads, filters)


@pytest.mark.parametrize("unit", [32, 96, 480, 1056, 1952])
def test_work_unit_lengths(ctx, monkeypatch, unit):
    """The kernel's work units are any multiple of 32 positions up to its tile (the library shortens them so that
    the units divide evenly over the persistent grid): every length gives the oracle's counts, reads that straddle a
    unit boundary included, on single- and multi-species layouts with W = 6 and wider rows."""
    import gen
    monkeypatch.setenv("MB_PILE_UNIT", str(unit))
    ref, reads = synth.synth_snps(seed=21, n_species=3, contigs_per_species=2, contig_len=7_000, n_reads=30_000,
                                  read_len=150, len_jitter=0.3, abundance="lognormal")
    _check(ctx, ref, reads, Filters())
    rng = np.random.default_rng(300 + unit)
    ref = gen.make_ref(rng, 2, 2, 3000, 9000)
    reads = gen.random_reads(301 + unit, ref, 5000, max_len=300)
    _check(ctx, ref, reads, Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=0))


@pytest.mark.parametrize("var", ["0", "8", "24", "40", "72", "74", "200", "584", "tma"])
def test_count_variants(ctx, monkeypatch, var):
    """The experiment switches of the W = 6 instance (round-1 kernel, L2 prefetch, class-sorted dealing, deferred indel
    reads, dealt COUNT, 1024-position tiles, three CTAs per SM, late plane loads, bulk-copy staging) against the
    oracle, at two unit lengths: every variant the DESIGN table cites still gives the reference's counts."""
    if var == "tma":
        monkeypatch.setenv("MB_PILE_TMA", "1")
    else:
        monkeypatch.setenv("MB_PILE_VAR", var)
    ref, reads = synth.synth_snps(seed=22, n_species=2, contigs_per_species=1, contig_len=60_000, n_reads=120_000,
                                  read_len=150, len_jitter=0.0, abundance="lognormal")
    _check(ctx, ref, reads, Filters())
    monkeypatch.setenv("MB_PILE_UNIT", "736")
    _check(ctx, ref, reads, Filters())


def test_unsorted_input_is_sorted_by_the_packer(ctx):
    import gen
    rng = np.random.default_rng(7)
    ref = gen.make_ref(rng, 2, 2, 2000, 4000)
    reads = gen.random_reads(8, ref, 3000, max_len=160, sort=False)
    _check(ctx, ref, reads, Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=10))


@pytest.mark.parametrize("n_reads,frac", [(9000, 0.9), (80000, 0.97)])
def test_deep_pile(ctx, n_reads, frac):
    """Thousands of reads on one locus: register counters are merged early (> 511 rows per lane) and,
    past 65 528 rows per tile, the shared counters are expanded in several passes."""
    import gen
    rng = np.random.default_rng(5)
    ref = gen.make_ref(rng, 1, 2, 3000, 5000)
    reads = gen.random_reads(9, ref, n_reads, max_len=60, min_len=20, pos_hot=(1, 2040, frac), long_del_frac=0.0)
    _check(ctx, ref, reads, Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=0))
    _check(ctx, ref, reads, Filters())


def test_empty_and_tiny(ctx):
    import gen
    rng = np.random.default_rng(3)
    ref = gen.make_ref(rng, 2, 1, 100, 5000)
    empty = gen.random_reads(1, ref, 0)
    _check(ctx, ref, empty, Filters())
    one = gen.random_reads(2, ref, 1, max_len=50)
    _check(ctx, ref, one, Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=0))


def test_full_size_c2(ctx):
    """BASELINE.json configs[1] at full size (5 M x 150 bp on 4 Mbp, the bench workload): every counter
    and every stat against the oracle (run on all host threads), plus the size-independent identities
    total_depth = sum of counters and covered_bases = #positions with depth > 0."""
    import os
    ref, reads = synth.synth_snps(seed=synth.SEED_C2, n_species=1, contigs_per_species=1, contig_len=4_000_000,
                                  n_reads=5_000_000, read_len=150, keep_ref_seq=False)
    lay = lib.Layout.of(ref)
    counts, stats = ctx.snps_pileup_plain(reads, lay, Filters())
    got = lay.unpad(counts)
    depth = got.astype(np.uint64).sum(0)
    assert int(stats["total_depth"][0]) == int(depth.sum())
    assert int(stats["covered_bases"][0]) == int((depth > 0).sum())
    assert int(stats["genome_length"][0]) == 4_000_000 and int(stats["aligned_reads"][0]) == 5_000_000
    exp_counts, exp_stats = oracle.snps_pileup(reads, ref, Filters(), threads=os.cpu_count() or 1)
    assert np.array_equal(got, exp_counts.astype(np.uint32))
    for k, name in enumerate(lib.STATS_DTYPE.names):
        assert int(stats[name][0]) == int(exp_stats[0, k]), name


def test_overlapped_host_path(ctx, monkeypatch):
    """Page-locked records, batch above the overlap threshold: seq / qual cross PCIe in chunks on a second stream
    while earlier chunks are re-laid (six Phred planes assumed). Same counts as the oracle; a batch that needs seven
    planes is re-packed exactly; missing qualities in a late chunk are still reported."""
    import gen
    monkeypatch.setenv("MB_OVERLAP_MIN_READS", "2000")
    ref, reads = synth.synth_snps(seed=31, n_species=1, contigs_per_species=3, contig_len=20_000, n_reads=60_000,
                                  len_jitter=0.2)
    lay = lib.Layout.of(ref)
    exp_counts, exp_stats = oracle.snps_pileup(reads, ref, Filters())
    for r in (lib.pinned_reads(reads), lib.pinned_reads(reads.nibble_packed())):
        launches = ctx.launch_count
        counts, stats = ctx.snps_pileup_plain(r, lay, Filters())
        assert np.array_equal(lay.unpad(counts).astype(np.uint64), exp_counts)
        for k, name in enumerate(lib.STATS_DTYPE.names):
            assert np.array_equal(stats[name], exp_stats[:, k]), name
        assert ctx.launch_count - launches >= 20          # the chunked path ran (one qor + fill pair per chunk)
    # Phred 64 in the last read: the speculation fails, the exact re-pack must give the seven-plane answer
    hi = lib.pinned_reads(reads)
    hi.qual[-5:] = 70
    e2, s2 = oracle.snps_pileup(hi, ref, Filters(baseq=65))
    c2, st2 = ctx.snps_pileup_plain(hi, lay, Filters(baseq=65))
    assert np.array_equal(lay.unpad(c2).astype(np.uint64), e2) and int(st2["mapped_reads"][0]) == int(s2[0, 4])
    # qualities missing in a read of the last chunk
    bad = lib.pinned_reads(reads)
    bad.qual[int(bad.seq_off[bad.n_reads - 3])] = 0xFF
    with pytest.raises(lib.MidasB200Error) as e:
        ctx.snps_pileup_plain(bad, lay, Filters())
    assert e.value.status == -1
    # random CIGARs (non-simple reads, long reads) through the chunked path
    rng = np.random.default_rng(3)
    ref2 = gen.make_ref(rng, 1, 4, 3000, 9000)
    rr = gen.random_reads(41, ref2, 5000, max_len=260)
    rr.qual[rr.qual > 63] = 63
    lay2 = lib.Layout.of(ref2)
    e3, s3 = oracle.snps_pileup(rr, ref2, Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=10))
    c3, st3 = ctx.snps_pileup_plain(lib.pinned_reads(rr), lay2, Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=10))
    assert np.array_equal(lay2.unpad(c3).astype(np.uint64), e3)
