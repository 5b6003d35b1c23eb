"""GPU: the drop-in (midas_b200.run.snps / .genes run_pipeline -> C-ABI -> CUDA kernels) writes, byte
for byte, the files the reference's unmodified run_pipeline wrote for the committed golden cases."""
import gzip
import io
import os

import numpy as np
import pytest

from golden_util import GENES_CASES, SNPS_CASES, expected, load_case
from oracle import fixtures

pytestmark = pytest.mark.gpu


def _args(outdir, db, filters, **kw):
    a = {"outdir": outdir, "db": db, "threads": 2, "mapid": filters.mapid, "readq": filters.readq,
         "mapq": filters.mapq, "aln_cov": filters.aln_cov, "baseq": filters.baseq, "build_db": False,
         "align": False, "call": True, "cov": True, "remove_temp": False, "log": io.StringIO()}
    a.update(kw)
    return a


@pytest.mark.parametrize("case", SNPS_CASES)
def test_snps_pipeline_writes_the_reference_files(case, tmp_path):
    from midas_b200.run import snps
    ref, reads, filters, extra = load_case(case)
    db, out = fixtures.write_snps_fixture(str(tmp_path), ref, reads, lowercase_some=extra.get("lowercase_some", False))
    snps.run_pipeline(_args(out, db, filters))
    exp = expected(case)
    for sid in ref.species_ids:
        got = gzip.open(os.path.join(out, "snps", "output", sid + ".snps.gz"), "rb").read()
        assert got == exp[sid + ".snps.tsv"], sid
    assert open(os.path.join(out, "snps", "summary.txt"), "rb").read() == exp["summary.txt"]


@pytest.mark.parametrize("case", GENES_CASES)
def test_genes_pipeline_writes_the_reference_files(case, tmp_path):
    from midas_b200.run import genes
    ref, reads, filters, extra = load_case(case)
    db, out = fixtures.write_genes_fixture(str(tmp_path), ref, reads, np.array(extra["marker"], np.int32))
    with pytest.warns(RuntimeWarning):
        genes.run_pipeline(_args(out, db, filters, exact_order=True))    # --exact_order: bit-exact float text
    exp = expected(case)
    for sid in ref.species_ids:
        got = gzip.open(os.path.join(out, "genes", "output", sid + ".genes.gz"), "rb").read()
        assert got == exp[sid + ".genes.tsv"], sid
    assert open(os.path.join(out, "genes", "summary.txt"), "rb").read() == exp["summary.txt"]
    # the default (one division of the integer sum, no device sort): same counts, floats within 1e-12
    with pytest.warns(RuntimeWarning):
        genes.run_pipeline(_args(out, db, filters))
    for sid in ref.species_ids:
        got = gzip.open(os.path.join(out, "genes", "output", sid + ".genes.gz"), "rt").read().splitlines()
        want = exp[sid + ".genes.tsv"].decode().splitlines()
        assert got[0] == want[0] and len(got) == len(want)
        for g, w in zip(got[1:], want[1:]):
            g, w = g.split("\t"), w.split("\t")
            assert g[:2] == w[:2]
            assert np.allclose([float(g[2]), float(g[3])], [float(w[2]), float(w[3])], rtol=1e-12, atol=0)


def test_cli_snps_and_genes(tmp_path):
    """scripts/run_midas.py {snps --pileup, genes --call_genes} as a user runs them (subprocess)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "scripts", "run_midas.py")
    ref, reads, filters, extra = load_case("snps_c2_small")
    db, out = fixtures.write_snps_fixture(str(tmp_path / "s"), ref, reads, lowercase_some=True)
    r = subprocess.run([sys.executable, cli, "snps", out, "-d", db, "--pileup", "-t", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exp = expected("snps_c2_small")
    for sid in ref.species_ids:
        assert gzip.open(os.path.join(out, "snps", "output", sid + ".snps.gz"), "rb").read() == exp[sid + ".snps.tsv"]
    assert open(os.path.join(out, "snps", "summary.txt"), "rb").read() == exp["summary.txt"]
    log = open(os.path.join(out, "snps", "log.txt")).read()
    assert "===========Parameters===========" in log and "minimum base quality score: 30" in log
    assert "Counting alleles" in log
    readme = open(os.path.join(out, "snps", "readme.txt")).read()
    assert "output/{SPECIES_ID}.snps.gz" in readme and readme.rstrip().endswith("/rep_genomes")
    # a missing BAM is a clean error, not a crash
    os.remove(os.path.join(out, "snps", "temp", "genomes.bam"))
    r = subprocess.run([sys.executable, cli, "snps", out, "-d", db, "--pileup"], capture_output=True, text=True)
    assert r.returncode != 0 and "genomes.bam not found" in r.stderr
    gref, greads, gfilters, gextra = load_case("genes_c3_small")
    db2, out2 = fixtures.write_genes_fixture(str(tmp_path / "g"), gref, greads, np.array(gextra["marker"], np.int32))
    r = subprocess.run([sys.executable, cli, "genes", out2, "-d", db2, "--exact_order"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    gexp = expected("genes_c3_small")
    for sid in gref.species_ids:
        assert gzip.open(os.path.join(out2, "genes", "output", sid + ".genes.gz"), "rb").read() == gexp[sid + ".genes.tsv"]
    assert open(os.path.join(out2, "genes", "summary.txt"), "rb").read() == gexp["summary.txt"]


def test_cli_snps_under_torchrun_two_gpus(tmp_path):
    """torchrun --nproc-per-node 2 scripts/run_midas.py snps --pileup: species dealt to two ranks (one GPU each),
    every rank writes its species, rank 0 the summary - byte for byte the reference's files. Needs two GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "scripts", "run_midas.py")
    for case in SNPS_CASES:
        ref, reads, filters, extra = load_case(case)
        if len(ref.species_ids) < 2:
            continue
        db, out = fixtures.write_snps_fixture(str(tmp_path / case), ref, reads,
                                              lowercase_some=extra.get("lowercase_some", False))
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29731", cli, "snps", out, "-d", db, "--pileup", "-t", "2",
               "--mapid", str(filters.mapid), "--mapq", str(filters.mapq), "--baseq", str(filters.baseq),
               "--readq", str(filters.readq), "--aln_cov", str(filters.aln_cov)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        exp = expected(case)
        for sid in ref.species_ids:
            got = gzip.open(os.path.join(out, "snps", "output", sid + ".snps.gz"), "rb").read()
            assert got == exp[sid + ".snps.tsv"], (case, sid)
        assert open(os.path.join(out, "snps", "summary.txt"), "rb").read() == exp["summary.txt"]
