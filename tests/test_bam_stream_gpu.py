"""GPU: the streaming ingest (mb_bam_stream_open + mb_snps_pileup_bam: the inflated BAM record stream parsed on the
device) gives the counts of the host-decoded path and of the oracle, on BAM files with arbitrary CIGARs, unplaced
records, references that are dropped, missing NM / qualities."""
import numpy as np
import pytest

import gen
from midas_b200 import bam, lib, synth
from midas_b200.soa import Filters, RefSet
from oracle import oracle

pytestmark = pytest.mark.gpu


def _via_stream(ctx, path, ref, filters, tid_map=None, threads=4):
    """Both forms of the open: the stream uploaded by the pileup call, and uploaded chunk by chunk while the host
    threads inflate (mb_bam_stream_open_dev); a second call on the resident stream, and a call with another stream
    in between, must not disturb the result."""
    lay = lib.Layout.of(ref)
    out = []
    for dev in (False, True, True):
        st = lib.BamStream(path, threads, ctx=ctx if dev else None)
        index = {n: k for k, n in enumerate(ref.contig_names)}
        tm = np.array([index.get(n, -1) for n in st.names], np.int32) if tid_map is None else tid_map
        if dev and len(out) == 2:
            other = lib.BamStream(path, threads)              # replaces the resident copy: st is uploaded again
            ctx.snps_pileup_bam(other, tm, lay, filters)
            other.close()
        out.append(ctx.snps_pileup_bam(st, tm, lay, filters))
        if dev:
            again = ctx.snps_pileup_bam(st, tm, lay, filters)
            assert np.array_equal(again[0], out[-1][0]) and np.array_equal(again[1], out[-1][1])
        st.close()
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and np.array_equal(o[1], out[0][1]) and o[2:] == out[0][2:]
    counts, stats, n_reads, n_bases = out[0]
    return lay, counts, stats, n_reads, n_bases


@pytest.mark.parametrize("seed,n_species,cps,lo,hi,n_reads,max_len", [
    (301, 3, 2, 500, 6000, 3000, 300), (302, 1, 1, 5000, 5000, 4000, 150), (303, 2, 3, 40, 900, 2500, 120)])
def test_random_cigars_from_a_bam_file(ctx, tmp_path, seed, n_species, cps, lo, hi, n_reads, max_len):
    rng = np.random.default_rng(seed)
    ref = gen.make_ref(rng, n_species, cps, lo, hi)
    reads = gen.random_reads(seed, ref, n_reads, max_len=max_len)
    reads.ref_id[::97] = -1                                     # unplaced records: dropped
    p = str(tmp_path / "x.bam")
    bam.write_bam(p, ref.contig_names, ref.contig_len, reads)
    kept = reads.select(np.nonzero(reads.ref_id >= 0)[0])
    for filters in (Filters(), Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=0)):
        exp_counts, exp_stats = oracle.snps_pileup(kept, ref, filters)
        lay, counts, stats, n, nb = _via_stream(ctx, p, ref, filters)
        assert n == kept.n_reads and nb == kept.total_bases
        assert np.array_equal(lay.unpad(counts).astype(np.uint64), exp_counts)
        for k, name in enumerate(lib.STATS_DTYPE.names):
            assert np.array_equal(stats[name], exp_stats[:, k]), name


def test_dropped_references_and_c2_shape(ctx, tmp_path):
    ref, reads = synth.synth_snps(seed=5, n_species=3, contigs_per_species=2, contig_len=15_000, n_reads=80_000,
                                  len_jitter=0.2, abundance="lognormal")
    p = str(tmp_path / "y.bam")
    bam.write_bam(p, ref.contig_names, ref.contig_len, reads)
    # only species 0 and 2 are selected: the records of species 1's contigs are never fetched by the reference
    keep_c = np.nonzero(ref.contig_species != 1)[0]
    sub = RefSet([ref.contig_names[c] for c in keep_c], ref.contig_len[keep_c],
                 np.searchsorted([0, 2], ref.contig_species[keep_c]).astype(np.int32), [ref.species_ids[0], ref.species_ids[2]], None)
    cmap = -np.ones(ref.n_contigs, np.int32)
    cmap[keep_c] = np.arange(len(keep_c), dtype=np.int32)
    kept = reads.select(np.nonzero(cmap[reads.ref_id] >= 0)[0])
    kept.ref_id = np.ascontiguousarray(cmap[kept.ref_id], np.int32)
    exp_counts, exp_stats = oracle.snps_pileup(kept, sub, Filters())
    lay, counts, stats, n, nb = _via_stream(ctx, p, sub, Filters())
    assert n == kept.n_reads
    assert np.array_equal(lay.unpad(counts).astype(np.uint64), exp_counts)
    for k, name in enumerate(lib.STATS_DTYPE.names):
        assert np.array_equal(stats[name], exp_stats[:, k]), name


def test_errors_and_unsorted(ctx, tmp_path):
    rng = np.random.default_rng(1)
    ref = gen.make_ref(rng, 1, 2, 800, 900)
    reads = gen.random_reads(3, ref, 300, max_len=80)
    bad = reads.select(slice(0, 300))
    bad.nm[200] = -1                                            # no NM tag: KeyError in the reference
    p = str(tmp_path / "nm.bam")
    bam.write_bam(p, ref.contig_names, ref.contig_len, bad, extra_tags=False)
    with pytest.raises(lib.MidasB200Error) as e:
        _via_stream(ctx, p, ref, Filters())
    assert e.value.status == -1
    uns = gen.random_reads(4, ref, 300, max_len=80, sort=False)
    p2 = str(tmp_path / "uns.bam")
    bam.write_bam(p2, ref.contig_names, ref.contig_len, uns)
    with pytest.raises(lib.MidasB200Error) as e:
        _via_stream(ctx, p2, ref, Filters())
    assert e.value.status == -6
    with pytest.raises(lib.MidasB200Error):
        lib.BamStream(str(tmp_path / "missing.bam"))


def test_genes_coverage_from_a_bam_file(ctx, tmp_path):
    gref, greads, _ = synth.synth_genes(seed=12, n_species=3, genes_per_species=150, n_reads=40_000, read_len=110)
    p = str(tmp_path / "g.bam")
    bam.write_bam(p, gref.contig_names, gref.contig_len, greads)
    st = lib.BamStream(p, 4)
    assert st.names == gref.contig_names
    tid_map = np.arange(len(st.names), dtype=np.int32)
    for filters in (Filters.genes_defaults(), Filters(mapid=97.0, aln_cov=0.9, readq=30, mapq=30, baseq=0)):
        exp = oracle.genes_coverage(greads, gref.contig_len, filters)
        for exact in (False, True):
            out, n = ctx.genes_coverage_bam(st, tid_map, gref.contig_len, filters, exact_order=exact)
            assert n == greads.n_reads
            assert np.array_equal(out["aligned_reads"], exp["aligned"]) and np.array_equal(out["mapped_reads"], exp["mapped"])
            assert np.array_equal(out["mapped_bp"], exp["mapped_bp"])
            if exact:
                assert np.array_equal(out["depth"].view(np.uint64), exp["depth"].view(np.uint64))
            else:
                assert np.allclose(out["depth"], exp["depth"], rtol=1e-12, atol=0)
    st.close()


@pytest.mark.parametrize("dev", [False, True])
def test_malformed_streams(ctx, tmp_path, dev):
    """Both forms of the open refuse what is not a BAM; a record whose block_size runs past the end of the stream or
    whose fields exceed it is MB_EMALFORMED from the call that walks the records - never a crash."""
    import struct
    ref, reads = synth.synth_snps(seed=31, n_species=1, contigs_per_species=2, contig_len=4_000, n_reads=3_000)
    good = str(tmp_path / "good.bam")
    bam.write_bam(good, ref.contig_names, ref.contig_len, reads)
    data = bytearray(bam.bgzf_decompress(open(good, "rb").read()))
    lay = lib.Layout.of(ref)
    tm = np.arange(ref.n_contigs, dtype=np.int32)

    def run(payload, name):
        p = str(tmp_path / name)
        with open(p, "wb") as fh:
            fh.write(bam.bgzf_compress(bytes(payload)))
        st = lib.BamStream(p, 4, ctx=ctx if dev else None)
        try:
            return ctx.snps_pileup_bam(st, tm, lay, Filters())
        finally:
            st.close()

    assert run(data, "ok.bam")[2] == reads.n_reads
    with pytest.raises(lib.MidasB200Error) as e:
        run(b"NOT\1" + bytes(data[4:]), "magic.bam")
    assert e.value.status == -5
    with pytest.raises(lib.MidasB200Error) as e:
        run(data[:40], "header.bam")                      # the reference list is cut
    assert e.value.status == -5
    # first record: block_size far past the end
    off = bam.read_bam_header(bytes(data))[3]
    bad = bytearray(data)
    struct.pack_into("<I", bad, off, 0x7FFFFFFF)
    with pytest.raises(lib.MidasB200Error) as e:
        run(bad, "trunc.bam")
    assert e.value.status == -5
    bad = bytearray(data)
    struct.pack_into("<I", bad, off + 4 + 16, 0x00FFFFFF)   # l_seq larger than the record
    with pytest.raises(lib.MidasB200Error) as e:
        run(bad, "fields.bam")
    assert e.value.status == -5
