"""CPU: BAM/BGZF codec round trip, and agreement with the independent reader inside the pysam shim."""
import os
import sys

import numpy as np
import pytest

import gen
from midas_b200 import bam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    ref = gen.make_ref(rng, 2, 2, 300, 2000)
    reads = gen.random_reads(4, ref, 400, max_len=120, sort=False)
    reads.nm[::7] = 300          # 16-bit NM
    p = str(tmp_path / "x.bam")
    bam.write_bam(p, ref.contig_names, ref.contig_len, reads)
    names, lens, back = bam.read_bam(p)
    assert names == ref.contig_names and [int(x) for x in lens] == [int(x) for x in ref.contig_len]
    for f in ("ref_id", "pos", "mapq", "nm", "l_qseq", "cigar_off", "cigar", "seq_off", "seq", "qual"):
        assert np.array_equal(getattr(back, f), getattr(reads, f)), f


def test_shim_reader_sees_the_same_records(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
    import pysam
    rng = np.random.default_rng(3)
    ref = gen.make_ref(rng, 1, 2, 300, 900)
    reads = gen.random_reads(5, ref, 150, max_len=90)
    p = str(tmp_path / "y.bam")
    bam.write_bam(p, ref.contig_names, ref.contig_len, reads)
    with pysam.AlignmentFile(p, "rb") as fh:
        recs = list(fh.fetch(until_eof=True))
        assert fh.getrname(0) == ref.contig_names[0]
    assert len(recs) == reads.n_reads
    for i in (0, 33, 149):
        a = recs[i]
        s0, s1 = int(reads.seq_off[i]), int(reads.seq_off[i + 1])
        assert a.reference_id == reads.ref_id[i] and a.mapping_quality == reads.mapq[i]
        assert dict(a.tags)["NM"] == reads.nm[i] and a.query_length == s1 - s0
        assert list(a.query_qualities) == list(reads.qual[s0:s1])


def test_empty_bam(tmp_path):
    from midas_b200.soa import PlainReads
    p = str(tmp_path / "e.bam")
    empty = PlainReads.from_records([])
    bam.write_bam(p, ["c1"], [100], empty)
    names, lens, back = bam.read_bam(p)
    assert names == ["c1"] and back.n_reads == 0


def test_native_reader_agrees_with_python_reader(tmp_path):
    from midas_b200 import lib
    rng = np.random.default_rng(9)
    ref = gen.make_ref(rng, 3, 2, 300, 3000)
    reads = gen.random_reads(10, ref, 3000, max_len=200, sort=False)
    reads.nm[::5] = 70000           # 32-bit NM
    reads.ref_id[::97] = -1         # unplaced records are dropped by both readers
    p = str(tmp_path / "n.bam")
    bam.write_bam(p, ref.contig_names, ref.contig_len, reads)
    n1, l1, r1 = bam.read_bam(p)
    fields = ("ref_id", "pos", "mapq", "nm", "l_qseq", "cigar_off", "cigar", "seq_off", "seq", "qual", "flag")
    for threads in (1, 4):
        n2, l2, r2 = lib.read_bam(p, threads, seq_bytes=True)
        assert n1 == n2 and [int(x) for x in l1] == l2 and not r2.seq_nibbles
        for f in fields:
            assert np.array_equal(getattr(r1, f), getattr(r2, f)), f
        # default: seq as it lies in the file (two codes per byte, every read on a byte boundary)
        _, _, r3 = lib.read_bam(p, threads)
        assert r3.seq_nibbles and not np.any(r3.seq_off & np.uint64(1))
        u = r3.unpacked()
        for f in fields:
            assert np.array_equal(getattr(r1, f), getattr(u, f)), f
        assert np.array_equal(r1.nibble_packed().seq, r3.seq) and np.array_equal(r1.nibble_packed().qual, r3.qual)
    assert r1.n_reads == reads.n_reads - len(range(0, reads.n_reads, 97))


def test_native_reader_rejects_garbage(tmp_path):
    import pytest
    from midas_b200 import lib
    p = tmp_path / "bad.bam"
    p.write_bytes(b"this is not a bam file at all, not even gzip")
    with pytest.raises(lib.MidasB200Error) as e:
        lib.read_bam(str(p))
    assert e.value.status == -5
    with pytest.raises(lib.MidasB200Error):
        lib.read_bam(str(tmp_path / "missing.bam"))


@pytest.mark.parametrize("level", [0, 1, 4, 9])
def test_native_reader_inflates_every_kind_of_block(tmp_path, level, capfd, monkeypatch):
    """The library's own DEFLATE decoder (csrc/host_inflate.cpp; CRC-checked, zlib behind it for whatever it refuses)
    on BGZF blocks written at several zlib levels - stored blocks at level 0, fixed-Huffman blocks for tiny members,
    dynamic ones otherwise - must hand the readers the same bytes as Python's zlib, and must be the path taken."""
    from midas_b200 import lib
    monkeypatch.setenv("MB_BAM_TIMING", "1")
    for seed, n_reads, max_len in ((1, 2500, 150), (2, 12, 40), (3, 9000, 300)):
        rng = np.random.default_rng(100 + seed)
        ref = gen.make_ref(rng, 2, 2, 3000, 9000)
        reads = gen.random_reads(seed, ref, n_reads, max_len=max_len)
        p = str(tmp_path / f"l{level}_{seed}.bam")
        bam.write_bam(p, ref.contig_names, ref.contig_len, reads, level=level)
        n1, l1, r1 = bam.read_bam(p)
        capfd.readouterr()
        n2, l2, r2 = lib.read_bam(p, 3, seq_bytes=True)
        err = capfd.readouterr().err
        assert n1 == n2
        for f in ("ref_id", "pos", "mapq", "nm", "l_qseq", "cigar_off", "cigar", "seq_off", "seq", "qual"):
            assert np.array_equal(getattr(r1, f), getattr(r2, f)), f
        assert " 0 of " in err and "blocks through zlib" in err, err          # no block needed the fallback
