"""`merge_midas.py snps`: the mirror in midas_b200/merge writes, byte for byte, the files
the reference's UNMODIFIED pipeline wrote for the same sample directories and database
(tests/golden/merge_pipeline, made by tests/golden/make_merge_golden.py from /root/reference).
CPU: species / sample selection, table parsing, site filter, annotation and text, with the numpy oracle
standing in for the GPU call. GPU: the same through mb_merge_sites_dev."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden", "merge_pipeline")
CASES = sorted(os.listdir(os.path.join(GOLD, "cases")))
FILES = ["snps_info.txt", "snps_freq.txt", "snps_depth.txt", "snps_summary.txt"]

from midas_b200.merge import snps as msnps          # noqa: E402
from oracle import merge_oracle as mo               # noqa: E402
from oracle.run_reference import merge_args         # noqa: E402


def _inputs():
    db = os.path.join(GOLD, "input", "db")
    sroot = os.path.join(GOLD, "input", "samples")
    return db, [os.path.join(sroot, d) for d in sorted(os.listdir(sroot))]


def _run_and_compare(case, tmp_path, engine):
    with open(os.path.join(GOLD, "cases", case, "args.json")) as fh:
        over = json.load(fh)
    db, dirs = _inputs()
    args = merge_args(str(tmp_path / "out"), db, dirs, **over)
    os.makedirs(args["outdir"])
    msnps.run_pipeline(args, engine=engine)
    expected_species = sorted(d for d in os.listdir(os.path.join(GOLD, "cases", case)) if d != "args.json")
    assert sorted(os.listdir(args["outdir"])) == expected_species
    for sp in expected_species:
        for fn in FILES:
            with open(os.path.join(GOLD, "cases", case, sp, fn), "rb") as a, \
                    open(os.path.join(args["outdir"], sp, fn), "rb") as b:
                assert a.read() == b.read(), (case, sp, fn)


@pytest.mark.parametrize("case", CASES)
def test_pipeline_files_match_the_reference_with_the_oracle_engine(case, tmp_path):
    _run_and_compare(case, tmp_path, mo.merge_sites)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_pipeline_files_match_the_reference_on_the_gpu(case, tmp_path):
    _run_and_compare(case, tmp_path, None)


@pytest.mark.parametrize("chunk_rows", [1, 37, 500])
def test_site_chunks_of_any_size_give_the_same_files(chunk_rows, tmp_path):
    """The sites are processed a chunk at a time (memory O(samples x chunk), not O(samples x genome)): whatever the
    chunk size - one row, a size that leaves a ragged last chunk, larger than a table - the files are the reference's."""
    case = "all_sites"
    with open(os.path.join(GOLD, "cases", case, "args.json")) as fh:
        over = json.load(fh)
    db, dirs = _inputs()
    args = merge_args(str(tmp_path / "out"), db, dirs, chunk_rows=chunk_rows, **over)
    os.makedirs(args["outdir"])
    msnps.run_pipeline(args, engine=mo.merge_sites)
    for sp in sorted(d for d in os.listdir(os.path.join(GOLD, "cases", case)) if d != "args.json"):
        for fn in FILES:
            with open(os.path.join(GOLD, "cases", case, sp, fn), "rb") as a, open(os.path.join(args["outdir"], sp, fn), "rb") as b:
                assert a.read() == b.read(), (sp, fn)


def test_gene_lookup_equals_the_reference_pointer_walk():
    """GeneIndex.find against a direct restatement of the pointer walk of annotate (snps.py:121-139)."""
    db, _ = _inputs()
    genes = msnps.read_genes("Sp_alpha_1", db)
    index = msnps.GeneIndex(genes)
    ptr = 0
    for rid in ("alpha.c1", "alpha.c2", "alpha.c3"):
        for pos in range(1, 1600):
            while ptr < len(genes) and (rid > genes[ptr]["scaffold_id"] or
                                        (rid == genes[ptr]["scaffold_id"] and pos > genes[ptr]["end"])):
                ptr += 1
            exp = None
            if ptr < len(genes) and not (rid < genes[ptr]["scaffold_id"] or
                                         (rid == genes[ptr]["scaffold_id"] and pos < genes[ptr]["start"])):
                exp = genes[ptr]
            assert index.find(rid, pos) is exp, (rid, pos)


def test_cli_presets_follow_the_reference():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import merge_midas
    a = merge_midas.snps_arguments(["snps", "out", "-i", "x", "-t", "list"])
    assert (a["snp_type"], a["site_prev"], a["allele_freq"], a["site_depth"], a["site_ratio"], a["sample_depth"],
            a["fract_cov"], a["min_samples"], a["max_sites"]) == (["bi"], 0.95, 0.01, 1, 2.0, 5.0, 0.4, 1, float("inf"))
    a = merge_midas.snps_arguments(["snps", "out", "-i", "x", "-t", "list", "--all_sites", "--all_samples"])
    assert (a["snp_type"], a["site_prev"], a["sample_depth"], a["fract_cov"]) == (["any"], 0.0, 0.0, 0.0)
    a = merge_midas.snps_arguments(["snps", "out", "-i", "x", "-t", "list", "--all_sites", "--core_snps"])
    assert (a["snp_type"], a["site_prev"]) == (["bi"], 0.95)      # later presets win, as in the reference


def test_native_row_formatter_equals_python_format():
    """mb_format_merge_{freq,depth}_rows against '{0:.3g}'.format / str (snps.py:196-201) on ratios of small
    integers (what a MAF is), values that round across a digit boundary, tiny and exact values."""
    import numpy as np
    from midas_b200 import lib
    rng = np.random.default_rng(11)
    S, G = 7, 4000
    num = rng.integers(0, 400, (S, G)).astype(np.float64)
    den = num + rng.integers(0, 400, (S, G))
    maf = np.where(den > 0, num / np.maximum(den, 1), 0.0)
    maf[0, :12] = [0.0, 1.0, 0.5, 0.9995, 0.99949999, 0.0009995, 1e-5, 1.2345e-7, 0.1, 0.125, 0.3335, 0.33349999]
    depth = rng.integers(0, 2 ** 32, (S, G), dtype=np.uint64).astype(np.uint32)
    keep = np.sort(rng.choice(G, 1500, replace=False)).astype(np.uint64)
    freq, dep = lib.format_merge_rows(keep, maf, depth)
    exp_f = "".join(str(int(g) + 1) + "\t" + "\t".join("{0:.3g}".format(x) for x in maf[:, g].tolist()) + "\n" for g in keep)
    exp_d = "".join(str(int(g) + 1) + "\t" + "\t".join(str(x) for x in depth[:, g].tolist()) + "\n" for g in keep)
    assert freq.decode() == exp_f and dep.decode() == exp_d
    # a chunk that starts at site 123456: the printed ids are offset, the data indices are not
    freq2, dep2 = lib.format_merge_rows(keep[:5], maf, depth, first_site=123456)
    assert freq2.decode().split("\n")[0].split("\t")[0] == str(123456 + int(keep[0]) + 1)
    assert freq2.decode().split("\t", 1)[1].split("\n")[0] == exp_f.split("\t", 1)[1].split("\n")[0]


def test_species_shard_over_ranks_without_a_collective(tmp_path):
    """Two ranks, each merging every second species into the same outdir: the union is the single-process result."""
    case = "all_sites"
    with open(os.path.join(GOLD, "cases", case, "args.json")) as fh:
        over = json.load(fh)
    db, dirs = _inputs()
    args = merge_args(str(tmp_path / "out"), db, dirs, **over)
    os.makedirs(args["outdir"])
    done = []
    for rank in (0, 1):
        done += [sp.id for sp in msnps.run_pipeline(dict(args), engine=mo.merge_sites, shard=(rank, 2))]
    assert sorted(done) == ["Sp_alpha_1", "Sp_beta_2"] and len(set(done)) == 2
    for sp in done:
        for fn in FILES:
            with open(os.path.join(GOLD, "cases", case, sp, fn), "rb") as a, open(os.path.join(args["outdir"], sp, fn), "rb") as b:
                assert a.read() == b.read(), (sp, fn)
