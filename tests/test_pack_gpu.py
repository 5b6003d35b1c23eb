"""GPU: the device packer (csrc/pack.cu, mb_reads_pack) against the host packer (mb_pack_reads) - the same
bytes in every array of the packed batch - and the inputs on which the reference itself raises."""
import numpy as np
import pytest

import gen
from midas_b200 import lib, synth
from midas_b200.soa import PlainReads, RefSet, OP_M, OP_I, OP_D, OP_S, OP_P

pytestmark = pytest.mark.gpu


def _same(ctx, reads, lay, mode):
    host = lib.PackedReads(reads, lay, mode, sort=(mode == "snps"))
    dev = ctx.pack(reads, lay, mode)
    got = dev.download()
    for f in ("n_reads", "n_vec", "total_bases", "max_ref_span", "row_words", "n_long", "q_planes", "n_planes"):
        assert getattr(got, f) == getattr(host, f), f
    assert np.array_equal(got.hdr, host.hdr)
    assert np.array_equal(got.vec_off, host.vec_off)
    assert np.array_equal(got.planes, host.planes), f"{int((got.planes != host.planes).sum())} plane words differ"
    assert np.array_equal(np.sort(got.long_idx), host.long_idx)
    # cplx: identical for simple / single-indel reads; otherwise an index of the read's raw CIGAR words
    fl = host.hdr["flags"]
    inline = (fl & 1).astype(bool) | (fl & 8).astype(bool)
    assert np.array_equal(got.cplx[inline], host.cplx[inline])
    gc, hc = got.cigar, host.cigar
    for i in np.nonzero(~inline)[0][:2000]:
        a, b, q, lq = int(got.cplx[i]), int(host.cplx[i]), 0, int(host.hdr["l_qseq"][i])
        while q < lq:
            assert gc[a] == hc[b]
            if int(hc[b]) & 15 in (0, 1, 4, 7, 8):
                q += int(hc[b]) >> 4
            a += 1; b += 1
    return dev


@pytest.mark.parametrize("seed,n_species,cps,lo,hi,n_reads,max_len", [
    (201, 3, 2, 500, 6000, 3000, 300),
    (202, 1, 1, 5000, 5000, 4000, 150),
    (203, 2, 3, 40, 900, 2500, 120),        # contigs shorter than reads: clipping at the contig end
    (204, 4, 1, 3000, 9000, 6000, 520),     # long reads: MB_RF_LONG, 7 Phred planes
])
def test_random_cigars_pack_identically(ctx, seed, n_species, cps, lo, hi, n_reads, max_len):
    rng = np.random.default_rng(seed)
    ref = gen.make_ref(rng, n_species, cps, lo, hi)
    reads = gen.random_reads(seed, ref, n_reads, max_len=max_len)
    lay = lib.Layout.of(ref)
    _same(ctx, reads, lay, "snps")
    _same(ctx, reads.nibble_packed(), lay, "snps")


def test_bam_order_differs_from_layout_order(ctx):
    """A BAM is sorted by (tid, pos); the layout is species-major. Contigs of two species interleaved in the
    BAM header move as whole blocks."""
    rng = np.random.default_rng(5)
    ref = gen.make_ref(rng, 3, 3, 800, 3000)
    perm = rng.permutation(ref.n_contigs)
    ref2 = RefSet([ref.contig_names[c] for c in perm], ref.contig_len[perm], ref.contig_species[perm], ref.species_ids,
                  [ref.contig_seq[c] for c in perm])
    assert (np.diff(ref2.contig_species) < 0).any()
    reads = gen.random_reads(6, ref2, 5000, max_len=150)
    _same(ctx, reads, lib.Layout.of(ref2), "snps")


def test_synthetic_shapes_and_genes_mode(ctx):
    ref, reads = synth.synth_snps(seed=8, n_species=3, contigs_per_species=2, contig_len=30_000, n_reads=200_000,
                                  len_jitter=0.3, abundance="lognormal")
    _same(ctx, reads, lib.Layout.of(ref), "snps")
    gref, greads, _ = synth.synth_genes(seed=9, n_species=4, genes_per_species=300, n_reads=100_000, read_len=125)
    _same(ctx, greads, lib.Layout.of(gref), "genes")
    _same(ctx, greads.nibble_packed(), lib.Layout.genes(gref.n_contigs), "genes")


def test_empty_batch(ctx):
    rng = np.random.default_rng(3)
    ref = gen.make_ref(rng, 2, 1, 100, 5000)
    _same(ctx, gen.random_reads(1, ref, 0), lib.Layout.of(ref), "snps")


def _one(cigar, seq=None, qual=None, pos=10, nm=0, mapq=30, lq=None):
    lq = lq if lq is not None else sum(l for o, l in cigar if o in (OP_M, OP_I, OP_S, 7, 8))
    seq = seq if seq is not None else [1, 2, 4, 8] * (lq // 4) + [1] * (lq % 4)
    qual = qual if qual is not None else [30] * lq
    return PlainReads.from_records([dict(ref_id=0, pos=pos, mapq=mapq, nm=nm, cigar=cigar, seq=seq, qual=qual)])


@pytest.mark.parametrize("kw,status", [
    (dict(nm=-1), -1),                        # KeyError 'NM' (snps.py:148)
    (dict(cigar=[(OP_S, 10)]), -1),           # aligned length 0 -> ZeroDivisionError (snps.py:148)
    (dict(cigar=[(OP_M, 5), (OP_P, 1), (OP_M, 5)]), -5),
    (dict(cigar=[(OP_M, 9)], lq=10), -5),     # CIGAR does not consume the query
    (dict(qual=[0xFF] + [30] * 9, cigar=[(OP_M, 10)]), -1),   # qualities missing -> np.mean(None)
    (dict(qual=[200] + [30] * 9, cigar=[(OP_M, 10)]), -1),    # a quality above 127
    (dict(pos=1000, cigar=[(OP_M, 10)]), -5),
])
def test_inputs_on_which_the_reference_raises(ctx, kw, status):
    kw.setdefault("cigar", [(OP_M, 10)])
    r = _one(**kw)
    lay = lib.Layout([1000], [0], 1)
    with pytest.raises(lib.MidasB200Error) as e:
        ctx.pack(r, lay, "snps")
    assert e.value.status == status
    # and the context is still usable afterwards
    ok = ctx.pack(_one([(OP_M, 10)]), lay, "snps")
    assert ok.n_reads == 1


def test_unsorted_records_are_refused(ctx):
    rng = np.random.default_rng(1)
    ref = gen.make_ref(rng, 2, 2, 500, 900)
    reads = gen.random_reads(3, ref, 500, max_len=80, sort=False)
    with pytest.raises(lib.MidasB200Error) as e:
        ctx.pack(reads, lib.Layout.of(ref), "snps")
    assert e.value.status == -6
