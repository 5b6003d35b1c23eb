"""Known-answer tests for the pysam half of the snps path (tests/golden/kat_pileup.json: expected counts
written by hand from the SAM specification and pysam's documented behaviour, nobody's code produced them).
Every implementation must reproduce them: the two CPU restatements (C and pure Python), the pysam shim under
the reference's UNMODIFIED midas/run/snps.py (build container only), and - marked gpu - the CUDA path through
both host entry points of the C-ABI."""
import gzip
import json
import os
import re

import numpy as np
import pytest

from midas_b200.soa import BAM_NIBBLE_CHARS, CIGAR_CHARS, Filters, PlainReads, RefSet
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "kat_pileup.json")) as fh:
    KAT = json.load(fh)["cases"]
OK_CASES = [c for c in KAT if "raises" not in c]
BAD_CASES = [c for c in KAT if "raises" in c]


def _expand(v, as_int=False):
    """"A*50" / "20*49,19" / list -> list."""
    if isinstance(v, list):
        return v
    out = []
    for part in v.split(","):
        m = re.fullmatch(r"(\w+)\*(\d+)", part)
        item, n = (m.group(1), int(m.group(2))) if m else (part, 1)
        out += [int(item) if as_int else item] * n
    return out


def build(case):
    names = [c[0] for c in case["contigs"]]
    species = []
    for c in case["contigs"]:
        if c[2] not in species:
            species.append(c[2])
    rng = np.random.default_rng(0)
    seqs = [np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, c[1])].tobytes() for c in case["contigs"]]
    ref = RefSet(names, np.array([c[1] for c in case["contigs"]], np.uint32),
                 np.array([species.index(c[2]) for c in case["contigs"]], np.int32), species, seqs)
    recs = []
    for r in case["reads"]:
        seq = "".join(_expand(r["seq"])) if not isinstance(r["seq"], list) else r["seq"]
        cig = [] if r["cigar"] == "*" else [(CIGAR_CHARS.index(op), int(n)) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", r["cigar"])]
        qual = [0xFF] * len(seq) if r["qual"] == "*" else _expand(r["qual"], as_int=True)
        assert all(ch in BAM_NIBBLE_CHARS for ch in seq) and len(qual) == len(seq)
        recs.append(dict(ref_id=names.index(r["contig"]), pos=r["pos"], mapq=r["mapq"],
                         nm=-1 if r["nm"] is None else r["nm"], cigar=cig, seq=seq, qual=qual, flag=r["flag"]))
    reads = PlainReads.from_records(recs)
    f = case["filters"]
    return ref, reads, Filters(f["mapid"], f["aln_cov"], f["readq"], f["mapq"], f["baseq"])


def expected(case, ref):
    """(counts u64 [4][sum contig_len] in contig order, stats u64 [n_species][5])."""
    off = np.concatenate(([0], np.cumsum(ref.contig_len.astype(np.int64))))
    counts = np.zeros((4, int(off[-1])), np.uint64)
    for name, d in case.get("counts", {}).items():
        c = ref.contig_names.index(name)
        for pos, v in d.items():
            counts[:, off[c] + int(pos)] = v
    for name, runs in case.get("counts_runs", {}).items():
        c = ref.contig_names.index(name)
        for lo, hi, v in runs:
            counts[:, off[c] + lo:off[c] + hi] = np.array(v, np.uint64)[:, None]
    stats = np.array([case["stats"][s] for s in ref.species_ids], np.uint64)
    # the hand-written stats must agree with the hand-written counts
    depth = counts.sum(0)
    for k in range(ref.n_species):
        sel = np.concatenate([np.arange(off[c], off[c + 1]) for c in range(ref.n_contigs) if ref.contig_species[c] == k])
        assert int(stats[k, 0]) == len(sel) and int(stats[k, 1]) == int((depth[sel] > 0).sum())
        assert int(stats[k, 2]) == int(depth[sel].sum())
    return counts, stats


@pytest.mark.parametrize("case", OK_CASES, ids=lambda c: c["name"])
def test_cpu_restatements_reproduce_the_known_answers(case):
    ref, reads, filters = build(case)
    exp_c, exp_s = expected(case, ref)
    for fn in (oracle.snps_pileup, oracle.py_snps_pileup):
        c, s = fn(reads, ref, filters)
        assert np.array_equal(c, exp_c), fn.__name__
        assert np.array_equal(s, exp_s), fn.__name__


@pytest.mark.parametrize("case", BAD_CASES, ids=lambda c: c["name"])
def test_cpu_restatements_raise_where_the_reference_raises(case):
    ref, reads, filters = build(case)
    for fn in (oracle.snps_pileup, oracle.py_snps_pileup):
        with pytest.raises(oracle.OracleError) as e:
            fn(reads, ref, filters)
        assert e.value.rc == -1


def _run_reference(case, tmp_path):
    from oracle import fixtures, run_reference
    ref, reads, filters = build(case)
    db, out = fixtures.write_snps_fixture(str(tmp_path), ref, reads)
    species = run_reference.run_snps(out, db, filters)
    return ref, out, species


@pytest.mark.parametrize("case", OK_CASES, ids=lambda c: c["name"])
def test_unmodified_reference_over_the_shim_reproduces_the_known_answers(case, tmp_path):
    from oracle import run_reference
    if not run_reference.available():
        pytest.skip("needs /root/reference (build container)")
    ref, out, species = _run_reference(case, tmp_path)
    exp_c, exp_s = expected(case, ref)
    off = np.concatenate(([0], np.cumsum(ref.contig_len.astype(np.int64))))
    got = np.zeros_like(exp_c)
    for sid in ref.species_ids:
        with gzip.open(os.path.join(out, "snps", "output", sid + ".snps.gz"), "rt") as fh:
            rows = [ln.rstrip("\n").split("\t") for ln in fh][1:]
        for rid, pos, _allele, depth, a, c, g, t in rows:
            k = ref.contig_names.index(rid)
            got[:, off[k] + int(pos) - 1] = [int(a), int(c), int(g), int(t)]
            assert int(depth) == int(a) + int(c) + int(g) + int(t)
    assert np.array_equal(got, exp_c)
    for k, sid in enumerate(ref.species_ids):
        sp = species[sid]
        assert [sp.genome_length, sp.covered_bases, sp.total_depth, sp.aligned_reads, sp.mapped_reads] == \
            [int(x) for x in exp_s[k]]


@pytest.mark.parametrize("case", BAD_CASES, ids=lambda c: c["name"])
def test_unmodified_reference_over_the_shim_raises(case, tmp_path):
    from oracle import run_reference
    if not run_reference.available():
        pytest.skip("needs /root/reference (build container)")
    with pytest.raises((KeyError, TypeError)):
        _run_reference(case, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("case", OK_CASES, ids=lambda c: c["name"])
def test_cuda_path_reproduces_the_known_answers(ctx, case):
    from midas_b200 import lib
    ref, reads, filters = build(case)
    exp_c, exp_s = expected(case, ref)
    lay = lib.Layout.of(ref)
    results = [ctx.snps_pileup_plain(reads, lay, filters), ctx.snps_pileup_plain(reads.nibble_packed(), lay, filters),
               ctx.snps_pileup(lib.PackedReads(reads, lay, "snps", sort=True), lay, filters)]
    for counts, stats in results:
        assert np.array_equal(lay.unpad(counts).astype(np.uint64), exp_c)
        for k, name in enumerate(lib.STATS_DTYPE.names):
            assert np.array_equal(stats[name], exp_s[:, k]), name


@pytest.mark.gpu
@pytest.mark.parametrize("case", BAD_CASES, ids=lambda c: c["name"])
def test_cuda_path_raises_where_the_reference_raises(ctx, case):
    from midas_b200 import lib
    ref, reads, filters = build(case)
    lay = lib.Layout.of(ref)
    with pytest.raises(lib.MidasB200Error) as e:
        ctx.snps_pileup_plain(reads, lay, filters)
    assert e.value.status == -1
    with pytest.raises(lib.MidasB200Error) as e:
        lib.PackedReads(reads, lay, "snps", sort=True)
    assert e.value.status == -1
