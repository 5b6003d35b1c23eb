"""CPU: the host packer (mb_pack_reads) - geometry fields, block encoding, run records, row width,
ordering, and the inputs on which the reference itself raises."""
import numpy as np
import pytest

import gen
from midas_b200 import lib, synth
from midas_b200.soa import PlainReads, cigar_tuples, OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_EQ, OP_X, OP_P


def _one(cigar, seq=None, qual=None, pos=10, nm=0, mapq=30, lq=None):
    lq = lq if lq is not None else sum(l for o, l in cigar if o in (OP_M, OP_I, OP_S, OP_EQ, OP_X))
    seq = seq if seq is not None else [1, 2, 4, 8] * (lq // 4) + [1] * (lq % 4)
    qual = qual if qual is not None else [30] * lq
    return PlainReads.from_records([dict(ref_id=0, pos=pos, mapq=mapq, nm=nm, cigar=cigar, seq=seq, qual=qual)])


LAY = None


def lay():
    global LAY
    if LAY is None:
        LAY = lib.Layout([1000], [0], 1)
    return LAY


def test_geometry_fields_follow_pysam_query_alignment():
    r = _one([(OP_H, 3), (OP_S, 5), (OP_M, 20), (OP_I, 2), (OP_M, 10), (OP_D, 4), (OP_M, 8), (OP_S, 7), (OP_H, 2)])
    pk = lib.PackedReads(r, lay(), "snps")
    h = pk.hdr[0]
    assert h["l_qseq"] == 52 and h["q_start"] == 5 and h["alen"] == 40       # snps.py:145-146
    assert h["ref_span"] == 42 and h["loc"] == 10 and not (h["flags"] & 1)
    # three runs: cplx points at the read's raw CIGAR words, which the kernels walk
    rec = pk.cigar[int(pk.cplx[0]):]
    assert [(int(w) & 15, int(w) >> 4) for w in rec[:9]] == cigar_tuples(r, 0)


def test_simple_reads_and_soft_clips():
    pk = lib.PackedReads(_one([(OP_S, 4), (OP_EQ, 30), (OP_S, 2)]), lay(), "snps")
    h = pk.hdr[0]
    assert h["flags"] & 1 and h["q_start"] == 4 and h["alen"] == 30 and h["ref_span"] == 30
    assert pk.cplx[0] == 0xFFFFFFFF and pk.row_words >= 2


def test_contig_end_overhang_is_clipped():
    pk = lib.PackedReads(_one([(OP_M, 50)], pos=980), lay(), "snps")
    assert pk.hdr[0]["ref_span"] == 20 and pk.hdr[0]["alen"] == 50
    pk = lib.PackedReads(_one([(OP_M, 30), (OP_D, 5), (OP_M, 20)], pos=960), lay(), "snps")
    # two runs -> carried in the cplx word (MB_RF_RUNS2): len0 | inserted << 16 | deleted << 24; the second
    # run is clipped at the contig end through ref_span (40 = 30 + 5 deleted + 5 of the 20 matched)
    assert pk.hdr[0]["flags"] & 8 and pk.hdr[0]["ref_span"] == 40
    cw = int(pk.cplx[0])
    assert (cw & 0xFFFF, (cw >> 16) & 0xFF, cw >> 24) == (30, 0, 5)


def test_two_run_reads_that_do_not_fit_the_run_word_keep_a_record():
    pk = lib.PackedReads(_one([(OP_M, 30), (OP_D, 300), (OP_M, 20)], pos=100), lay(), "snps")
    assert not (pk.hdr[0]["flags"] & 8)
    rec = pk.cigar[int(pk.cplx[0]):]
    assert [(int(w) & 15, int(w) >> 4) for w in rec[:3]] == [(OP_M, 30), (OP_D, 300), (OP_M, 20)]
    pk = lib.PackedReads(_one([(OP_D, 2), (OP_M, 30), (OP_I, 3), (OP_M, 20)], pos=100), lay(), "snps")
    assert not (pk.hdr[0]["flags"] & 8)            # first run not at the read's position


def _decode(pk, i):
    """(base code, candidate bit, Phred) per query base of read i from the bit-planar blocks."""
    hdr, off, planes = pk.hdr, pk.vec_off, pk.planes
    lq, NP, NQ = int(hdr[i]["l_qseq"]), pk.n_planes, pk.q_planes
    p0 = NP - NQ
    nw = -(-lq // 32) * NP
    # word w of a read = unit w // 8 of the read; the units of 32 consecutive reads are interleaved
    reg = np.array([planes[8 * (int(off[i]) + 32 * (w // 8)) + w % 8] for w in range(-(-nw // 8) * 8)])
    code, cand, ph = [], [], []
    for q in range(lq):
        w = reg[(q // 32) * NP:(q // 32 + 1) * NP]
        bit = q % 32
        if p0:
            code.append(((int(w[0]) >> bit) & 1) | (((int(w[1]) >> bit) & 1) << 1))
            cand.append((int(w[2]) >> bit) & 1)
        ph.append(sum(((int(w[p0 + k]) >> bit) & 1) << k for k in range(NQ)))
    # everything after the last base of the read is zero padding
    for q in range(lq, 32 * (-(-lq // 32))):
        w = reg[(q // 32) * NP:(q // 32 + 1) * NP]
        assert not any((int(x) >> (q % 32)) & 1 for x in w)
    assert not reg[-(-lq // 32) * NP:].any()
    return code, cand, ph


def test_block_encoding_round_trips():
    ref, reads = synth.synth_snps(seed=5, n_species=1, contigs_per_species=1, contig_len=5000, n_reads=300, read_len=77)
    pk = lib.PackedReads(reads, lib.Layout.of(ref), "snps")
    assert pk.n_reads == 300 and pk.total_bases == 300 * 77
    assert pk.q_planes == 6 and pk.n_planes == 9                    # every synthetic Phred is <= 41
    nib = np.array([1, 2, 4, 8])
    for i in (0, 17, 299):
        s0 = int(reads.seq_off[i])
        code, cand, ph = _decode(pk, i)
        # which query bases sit in an M/=/X run (and inside the contig): only those may be candidates
        in_run, q, r = [0] * 77, 0, 0
        room = int(ref.contig_len[reads.ref_id[i]]) - int(reads.pos[i])
        for c in reads.cigar[int(reads.cigar_off[i]):int(reads.cigar_off[i + 1])]:
            op, ln = int(c) & 15, int(c) >> 4
            if op in (OP_M, 7, 8):
                for x in range(ln):
                    in_run[q + x] = int(r + x < room)
                q += ln; r += ln
            elif op in (OP_I, 4):
                q += ln
            elif op in (OP_D, 3):
                r += ln
        for q in range(77):
            src = int(reads.seq[s0 + q])
            assert ph[q] == int(reads.qual[s0 + q])
            assert cand[q] == int(src in (1, 2, 4, 8) and in_run[q])
            if src in (1, 2, 4, 8):
                assert nib[code[q]] == src


def test_ragged_reads_share_interleaved_groups():
    """Reads of different lengths in one group of 32: every read decodes to its own bases, the group holds
    max-over-its-reads units per read, and vec_off follows the interleaving rule."""
    rng = np.random.default_rng(9)
    ref = gen.make_ref(rng, 1, 1, 4000, 4000)
    reads = gen.random_reads(10, ref, 100, max_len=200, min_len=5, long_del_frac=0.0)
    pk = lib.PackedReads(reads, lib.Layout.of(ref), "snps")
    lq = pk.hdr["l_qseq"].astype(np.int64)
    assert lq.min() < 64 < 130 < lq.max()                            # genuinely ragged
    units = (-(-lq // 32) * pk.n_planes + 7) // 8
    off, base = pk.vec_off, 0
    for g0 in range(0, 100, 32):
        g1 = min(100, g0 + 32)
        assert [int(x) for x in off[g0:g1]] == [base + k for k in range(g1 - g0)]
        base += 32 * int(units[g0:g1].max())
    assert int(off[100]) == base == pk.n_vec
    assert np.array_equal(pk.hdr["loc"], reads.pos.astype(np.uint32))   # generator output is already sorted
    for i in (0, 31, 32, 63, 99):
        s0 = int(reads.seq_off[i])
        _, _, ph = _decode(pk, i)
        assert ph == [int(x) for x in reads.qual[s0:s0 + int(lq[i])]]


def test_candidate_plane_excludes_clips_insertions_and_bases_past_the_contig_end():
    seq = "ACGTN" * 8                                                # 40 bases, every fifth is N
    pk = lib.PackedReads(_one([(4, 3), (OP_M, 10), (OP_I, 2), (OP_M, 20), (4, 5)], seq=seq, pos=985), lay(), "snps")
    _, cand, _ = _decode(pk, 0)                                      # contig length 1000: 15 reference bases left
    exp = [0] * 40
    for q, r in [(3 + k, k) for k in range(10)] + [(15 + k, 10 + k) for k in range(20)]:
        exp[q] = int(seq[q] != "N" and 985 + r < 1000)
    assert cand == exp


def test_seven_phred_planes_only_when_needed_and_genes_mode_holds_phred_planes_only():
    pk = lib.PackedReads(_one([(OP_M, 40)], qual=[63] * 40), lay(), "snps")
    assert (pk.q_planes, pk.n_planes) == (6, 9)
    pk = lib.PackedReads(_one([(OP_M, 40)], qual=[64] + [3] * 39), lay(), "snps")
    assert (pk.q_planes, pk.n_planes) == (7, 10)
    assert _decode(pk, 0)[2] == [64] + [3] * 39
    pk = lib.PackedReads(_one([(OP_M, 40)], qual=list(range(40))), lay(), "genes")
    assert (pk.q_planes, pk.n_planes) == (6, 6) and pk.n_vec == 64   # 2 blocks x 6 words = 2 units, group of 32
    assert _decode(pk, 0)[2] == list(range(40))


def test_row_width_and_long_reads():
    ref, reads = synth.synth_snps(seed=6, n_species=1, contigs_per_species=1, contig_len=30000, n_reads=4000)
    pk = lib.PackedReads(reads, lib.Layout.of(ref), "snps")
    assert pk.row_words == 6 and pk.n_long == 0                              # 150 bp + <= 3 bp deletion
    extra = _one([(OP_M, 40), (OP_D, 900), (OP_M, 40)], pos=100)
    extra.ref_id[:] = 0
    both = PlainReads.concat([reads, extra])
    pk = lib.PackedReads(both, lib.Layout.of(ref), "snps", sort=True)
    assert pk.row_words == 6 and pk.n_long == 1
    assert pk.hdr[int(pk.long_idx[0])]["flags"] & 4


def test_sorting_and_unsorted_rejection():
    rng = np.random.default_rng(1)
    ref = gen.make_ref(rng, 2, 2, 500, 900)
    reads = gen.random_reads(3, ref, 500, max_len=80, sort=False)
    L = lib.Layout.of(ref)
    with pytest.raises(lib.MidasB200Error) as e:
        lib.PackedReads(reads, L, "snps", sort=False)
    assert e.value.status == -6
    pk = lib.PackedReads(reads, L, "snps", sort=True)
    loc = pk.hdr["loc"]
    assert (np.diff(loc.astype(np.int64)) >= 0).all()
    # stable: equal keys keep input order -> same multiset of (loc, l_qseq) as a stable numpy sort
    key = L.contig_goff[reads.ref_id].astype(np.int64) + reads.pos
    order = np.argsort(key, kind="stable")
    assert np.array_equal(pk.hdr["l_qseq"], reads.l_qseq[order].astype(np.uint16))


@pytest.mark.parametrize("kw,status", [
    (dict(nm=-1), -1),                        # KeyError 'NM' (snps.py:148)
    (dict(cigar=[(OP_S, 10)]), -1),           # aligned length 0 -> ZeroDivisionError (snps.py:148)
    (dict(cigar=[(OP_M, 5), (OP_P, 1), (OP_M, 5)]), -5),
    (dict(cigar=[(OP_M, 9)], lq=10), -5),     # CIGAR does not consume the query
    (dict(qual=[0xFF] + [30] * 9, cigar=[(OP_M, 10)]), -1),   # qualities missing -> np.mean(None)
    (dict(pos=1000, cigar=[(OP_M, 10)]), -5),
])
def test_inputs_on_which_the_reference_raises(kw, status):
    kw.setdefault("cigar", [(OP_M, 10)])
    r = _one(**kw)
    with pytest.raises(lib.MidasB200Error) as e:
        lib.PackedReads(r, lay(), "snps")
    assert e.value.status == status


def test_nibble_packed_input_packs_identically():
    """seq as it lies in a BAM record (two codes per byte, odd reads padded) gives the same packed batch."""
    rng = np.random.default_rng(4)
    ref = gen.make_ref(rng, 2, 2, 500, 900)
    reads = gen.random_reads(5, ref, 400, max_len=90)
    L = lib.Layout.of(ref)
    a = lib.PackedReads(reads, L, "snps", sort=True)
    nb = reads.nibble_packed()
    assert nb.seq_nibbles and np.array_equal(nb.unpacked().seq, reads.seq)
    b = lib.PackedReads(nb, L, "snps", sort=True)
    for f in ("hdr", "vec_off", "cplx", "cigar", "planes", "long_idx"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_genes_layout_has_no_coordinate_limit():
    L = lib.Layout.genes(3_000_000_000)          # more "positions" than a 32-bit padded genome could hold
    r = _one([(OP_M, 10)])
    r.ref_id[:] = 5
    pk = lib.PackedReads(r, L, "genes")
    assert pk.hdr[0]["loc"] == 5
    with pytest.raises(lib.MidasB200Error):
        lib.PackedReads(r, L, "snps")


def test_genes_mode_keeps_file_order_and_gene_index():
    gref, greads, _ = synth.synth_genes(seed=9, n_species=2, genes_per_species=30, n_reads=500, read_len=100)
    pk = lib.PackedReads(greads, lib.Layout.of(gref), "genes")
    assert np.array_equal(pk.hdr["loc"], greads.ref_id.astype(np.uint32))
    assert pk.row_words == 0 and pk.n_long == 0
    assert np.array_equal(pk.hdr["mapq"], greads.mapq)
