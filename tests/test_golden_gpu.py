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
        genes.run_pipeline(_args(out, db, filters))          # exact_order=True: bit-exact float text
    exp = expected(case)
    for sid in ref.species_ids:
        got = gzip.open(os.path.join(out, "genes", "output", sid + ".genes.gz"), "rb").read()
        assert got == exp[sid + ".genes.tsv"], sid
    assert open(os.path.join(out, "genes", "summary.txt"), "rb").read() == exp["summary.txt"]
