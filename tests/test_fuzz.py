"""Property-based CIGAR fuzz: hypothesis draws small batches of alignments with arbitrary
CIGAR structure (hard / soft clips on either end, M / = / X runs, insertions, deletions, N skips, reads running
off the contig end, N and ambiguity codes, qualities around every threshold); the three CPU restatements of the
pysam path - C, pure Python and the shim under the reference's UNMODIFIED midas/run/snps.py - must agree on every
counter and stat. The GPU leg replays a fixed set of drawn batches through the CUDA path (marked gpu)."""
import gzip
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, seed, settings
from hypothesis import strategies as st

from midas_b200.soa import Filters, PlainReads, RefSet, OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_EQ, OP_X
from oracle import oracle

CONTIGS = [("fa1", 90, 0), ("fb1", 40, 1), ("fa2", 33, 0)]          # species 0 owns two non-adjacent contigs


def make_ref():
    rng = np.random.default_rng(1)
    seqs = [np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, c[1])].tobytes() for c in CONTIGS]
    return RefSet([c[0] for c in CONTIGS], np.array([c[1] for c in CONTIGS], np.uint32),
                  np.array([c[2] for c in CONTIGS], np.int32), ["spA", "spB"], seqs)


@st.composite
def cigars(draw):
    """[(op, len)] with at least one query-consuming aligned op; returns (ops, query_length)."""
    ops = []
    if draw(st.booleans()):
        ops.append((OP_H, draw(st.integers(1, 9))))
    if draw(st.booleans()):
        ops.append((OP_S, draw(st.integers(1, 12))))
    n_runs = draw(st.integers(1, 4))
    for k in range(n_runs):
        ops.append((draw(st.sampled_from([OP_M, OP_M, OP_EQ, OP_X])), draw(st.integers(1, 40))))
        if k < n_runs - 1:
            gap = draw(st.sampled_from([OP_I, OP_D, OP_N, OP_D]))
            ops.append((gap, draw(st.integers(1, 45 if gap != OP_I else 6))))
    if draw(st.booleans()):
        ops.append((OP_S, draw(st.integers(1, 12))))
    if draw(st.booleans()):
        ops.append((OP_H, draw(st.integers(1, 9))))
    lq = sum(l for o, l in ops if o in (OP_M, OP_I, OP_S, OP_EQ, OP_X))
    return ops, lq


@st.composite
def records(draw):
    c = draw(st.integers(0, len(CONTIGS) - 1))
    ops, lq = draw(cigars())
    pos = draw(st.integers(0, CONTIGS[c][1] - 1))
    seq = draw(st.lists(st.sampled_from([1, 2, 4, 8, 1, 2, 4, 8, 15, 5, 0]), min_size=lq, max_size=lq))
    qmode = draw(st.sampled_from(["high", "edge", "any"]))
    lo, hi = {"high": (28, 42), "edge": (17, 23), "any": (0, 93)}[qmode]
    qual = draw(st.lists(st.integers(lo, hi), min_size=lq, max_size=lq))
    return dict(ref_id=c, pos=pos, mapq=draw(st.sampled_from([0, 19, 20, 21, 60])),
                nm=draw(st.integers(0, max(1, lq // 8))), cigar=ops, seq=seq, qual=qual,
                flag=draw(st.sampled_from([0, 16, 256, 1024])))


FILTERS = st.sampled_from([Filters(), Filters(0.0, 0.0, 0, 0, 0), Filters(90.0, 0.5, 20, 20, 20),
                           Filters(97.0, 0.9, 30, 21, 41), Filters(94.0, 0.75, 20, 0, 64)])


def batch(recs):
    recs = sorted(recs, key=lambda r: (r["ref_id"], r["pos"]))       # BAM coordinate order
    return PlainReads.from_records(recs)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(recs=st.lists(records(), min_size=1, max_size=12), filters=FILTERS)
def test_c_and_python_restatements_agree(recs, filters):
    ref, reads = make_ref(), batch(recs)
    c1, s1 = oracle.snps_pileup(reads, ref, filters)
    c2, s2 = oracle.py_snps_pileup(reads, ref, filters)
    assert np.array_equal(c1, c2) and np.array_equal(s1, s2)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(recs=st.lists(records(), min_size=1, max_size=10), filters=FILTERS)
def test_unmodified_reference_over_the_shim_agrees_with_the_c_restatement(tmp_path_factory, recs, filters):
    from oracle import fixtures, run_reference
    if not run_reference.available():
        pytest.skip("needs /root/reference (build container)")
    ref, reads = make_ref(), batch(recs)
    exp_c, exp_s = oracle.snps_pileup(reads, ref, filters)
    tmp = tmp_path_factory.mktemp("fuzz")
    db, out = fixtures.write_snps_fixture(str(tmp), ref, reads)
    species = run_reference.run_snps(out, db, filters)
    off = np.concatenate(([0], np.cumsum(ref.contig_len.astype(np.int64))))
    got = np.zeros_like(exp_c)
    for sid in ref.species_ids:
        with gzip.open(os.path.join(out, "snps", "output", sid + ".snps.gz"), "rt") as fh:
            for ln in list(fh)[1:]:
                rid, pos, _a, _d, a, c, g, t = ln.rstrip("\n").split("\t")
                got[:, off[ref.contig_names.index(rid)] + int(pos) - 1] = [int(a), int(c), int(g), int(t)]
    assert np.array_equal(got, exp_c)
    for k, sid in enumerate(ref.species_ids):
        sp = species[sid]
        assert [sp.genome_length, sp.covered_bases, sp.total_depth, sp.aligned_reads, sp.mapped_reads] == \
            [int(x) for x in exp_s[k]]


@pytest.mark.gpu
def test_cuda_path_agrees_on_drawn_batches(ctx):
    """300 batches drawn by the same strategies (fixed seed) through both host entry points of the C-ABI."""
    from midas_b200 import lib
    ref = make_ref()
    lay = lib.Layout.of(ref)
    seen = []

    @seed(20260102)
    @settings(max_examples=300, deadline=None, database=None, suppress_health_check=list(HealthCheck))
    @given(recs=st.lists(records(), min_size=1, max_size=40), filters=FILTERS)
    def run(recs, filters):
        reads = batch(recs)
        exp_c, exp_s = oracle.snps_pileup(reads, ref, filters)
        for counts, stats in (ctx.snps_pileup_plain(reads, lay, filters),
                              ctx.snps_pileup(lib.PackedReads(reads, lay, "snps", sort=True), lay, filters)):
            assert np.array_equal(lay.unpad(counts).astype(np.uint64), exp_c)
            for k, name in enumerate(lib.STATS_DTYPE.names):
                assert np.array_equal(stats[name], exp_s[:, k]), name
        seen.append(reads.n_reads)

    run()
    assert len(seen) >= 100
