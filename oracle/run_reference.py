"""Execute the reference's UNMODIFIED hot-path functions (TEST INFRASTRUCTURE).

Imports `midas.run.snps` / `midas.run.genes` straight from the read-only reference tree
(default /root/reference; never copied) with oracle/shim on sys.path to stand in for the
two dependencies that cannot be installed here (pysam, Biopython). Only usable in the
build container — /root/reference does not exist on the GPU box — so its outputs are
committed as tests/golden/ by tests/golden/make_golden.py.

Executed verbatim from the reference:
  snps:  initialize_species, initialize_contigs, pysam_pileup -> species_pileup -> keep_read,
         snps_summary                                    (midas/run/snps.py:38-67,141-262)
  genes: initialize_species, initialize_genes, count_mapped_bp -> keep_read, normalize,
         write_results                                   (midas/run/genes.py:32-82,153-245)
Restated (shim): pysam.AlignmentFile.{fetch,count_coverage}, AlignedSegment, Bio.SeqIO.parse.
"""
from __future__ import annotations

import importlib
import io
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MIDAS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "midas", "run", "snps.py"))


def _import(modname):
    shim = os.path.join(_HERE, "shim")
    for p in (REFERENCE_ROOT, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module(modname)


def _args(outdir, db, filters, threads):
    return {"outdir": outdir, "db": db, "threads": threads, "mapid": float(filters.mapid),
            "readq": int(filters.readq), "mapq": int(filters.mapq), "aln_cov": float(filters.aln_cov),
            "baseq": int(filters.baseq), "build_db": False, "align": False, "call": True, "cov": True,
            "remove_temp": False, "log": io.StringIO(), "species_id": None}


def run_snps(outdir, db, filters, threads=1):
    """Reference `--pileup` stage minus index_bam (a samtools subprocess; the shim needs no index).
    Writes <outdir>/snps/output/<sp>.snps.gz and <outdir>/snps/summary.txt. Returns the Species dict."""
    snps = _import("midas.run.snps")
    args = _args(outdir, db, filters, threads)
    species = snps.initialize_species(args)
    contigs = snps.initialize_contigs(species)
    snps.pysam_pileup(args, species, contigs)
    snps.snps_summary(args, species)
    return species


def run_genes(outdir, db, filters):
    """Reference `--call_genes` stage. Writes <outdir>/genes/output/<sp>.genes.gz and genes/summary.txt."""
    genes_mod = _import("midas.run.genes")
    args = _args(outdir, db, filters, 1)
    species = genes_mod.initialize_species(args)
    genes = genes_mod.initialize_genes(args, species)
    genes_mod.pangenome_coverage(args, species, genes)
    return species, genes


def merge_args(outdir, db, indirs, **over):
    """args dict of scripts/merge_midas.py snps with the reference's defaults (:150-262)."""
    args = {"outdir": outdir, "db": db, "indirs": list(indirs), "threads": 1, "species_id": None, "min_samples": 1,
            "max_species": None, "sample_depth": 5.0, "fract_cov": 0.4, "max_samples": None, "all_samples": False,
            "snp_type": ["bi"], "allele_freq": 0.01, "site_depth": 1, "site_ratio": 2.0, "site_prev": 0.95,
            "max_sites": float("Inf")}
    args.update(over)
    return args


def run_merge_snps(args):
    """The reference's unmodified `merge_midas.py snps` pipeline (midas/merge/snps.py run_pipeline :469-508:
    select_species, the temporary count matrices, the forked site shards, GenomicSite per site, annotation,
    the three tables, snps_summary.txt). No pysam involved: only Bio.SeqIO comes from the shim."""
    import contextlib
    snps = _import("midas.merge.snps")
    os.makedirs(args["outdir"], exist_ok=True)
    with contextlib.redirect_stdout(io.StringIO()):
        snps.run_pipeline(dict(args))
