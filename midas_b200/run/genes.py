"""`run_midas.py genes` coverage stage on the B200 — mirror of the reference's midas/run/genes.py.

Same names and meaning as the reference module: Species, Gene, initialize_species (genes.py:32-48),
initialize_genes (:61-82), pangenome_coverage (:147-151), count_mapped_bp (:171-203), normalize
(:205-218), write_results (:220-245), run_pipeline (:252-294). The per-record loop of
count_mapped_bp runs on the GPU (mb_genes_coverage); the per-species / per-marker epilogues are
the reference's own few lines of numpy on the per-gene results.
"""
from __future__ import annotations

import csv
import os
import shutil
import sys
from collections import defaultdict
from time import time

import numpy as np

from .. import bam, lib
from ..soa import Filters, RefSet
from . import io as mio


_PAN_FILES = ("centroids.ffn", "cluster_info.txt", "gene_info.txt")


class Species:
    """Per-species record; attribute names are the reference's (genes.py:12-30), so its own epilogue
    and merge code can consume these objects."""

    def __init__(self, id):
        self.id = id
        self.paths = {}
        self.genes = []                      # per-gene depths, filled by count_mapped_bp
        self.pangenome_size = 0
        self.aligned_reads = 0
        self.mapped_reads = 0
        self.markers = defaultdict(float)    # marker_id -> summed depth

    def init_ref_db(self, ref_db):
        """Locate <db>/pan_genomes/<species>/{centroids.ffn,cluster_info.txt,gene_info.txt}[.gz]; a .gz
        file wins over a plain one, as in the reference (the later extension overwrites)."""
        self.dir = os.path.join(ref_db, "pan_genomes", self.id).replace(os.sep, "/")
        for name in _PAN_FILES:
            for candidate in (self.dir + "/" + name, self.dir + "/" + name + ".gz"):
                if os.path.isfile(candidate):
                    self.paths[name] = candidate


class Gene:
    """Per-gene record (reference genes.py:50-59 attribute names)."""

    def __init__(self, id):
        self.id = id
        self.aligned_reads = 0
        self.mapped_reads = 0
        self.depth = 0.0
        self.length = 0
        self.copies = 0.0
        self.marker_id = None


def initialize_species(args):
    """Species listed in <outdir>/genes/species.txt (reference genes.py:32-48). Selecting species from
    a species profile (--build_db) is upstream of the hot path and not part of this package."""
    if args.get("build_db"):
        sys.exit("\nError: the --build_db stage is not part of midas_b200; run it with the reference and "
                 "re-enter with --call_genes")
    species = {}
    listing = "%s/genes/species.txt" % args["outdir"]
    if os.path.isfile(listing):
        with open(listing) as fh:
            for line in fh:
                sid = line.rstrip()
                species[sid] = Species(sid)
    for sp in species.values():
        sp.init_ref_db(args["db"])
    return species


def initialize_genes(args, species):
    """Gene ids, owning species and lengths from each pangenome's centroids.ffn; marker ids from
    <db>/marker_genes/phyeco.map for genes that exist (reference genes.py:61-82)."""
    genes = {}
    for sp in species.values():
        for gene_id, seq in mio.read_fasta(sp.paths["centroids.ffn"]):
            g = Gene(gene_id)
            g.species_id, g.length = sp.id, len(seq)
            genes[gene_id] = g
            sp.pangenome_size += 1
    with mio.iopen("%s/marker_genes/phyeco.map" % args["db"]) as fh:
        for row in csv.DictReader(fh, delimiter="\t"):
            g = genes.get(row["gene_id"])
            if g is not None:
                g.marker_id = row["marker_id"]
    return genes


def pangenome_coverage(args, species, genes, ctx=None):
    count_mapped_bp(args, species, genes, ctx=ctx)
    normalize(args, species, genes)
    write_results(args, species, genes)


def count_mapped_bp(args, species, genes, ctx=None):
    """reference genes.py:171-203: per gene aligned/mapped reads and depth, per species summaries."""
    bam_path = "/".join([args["outdir"], "genes/temp/pangenomes.bam"])
    threads = int(args.get("threads", 0) or 0)
    gene_ids = list(genes.keys())
    gindex = {g: k for k, g in enumerate(gene_ids)}
    gene_len = np.array([genes[g].length for g in gene_ids], np.uint32)
    f = Filters(float(args["mapid"]), float(args["aln_cov"]), int(args["readq"]), int(args["mapq"]), 0)
    exact = bool(args.get("exact_order", False))
    own_ctx = ctx is None
    if own_ctx:
        ctx = lib.Context(int(args.get("device", 0)), use_torch_stream=False)
    try:
        # depth is one division of the integer sum (within 1e-12 of the reference's running float sum);
        # --exact_order reproduces the running sum bit for bit at the price of a device sort
        stream = lib.BamStream(bam_path, threads, ctx=None if args.get("host_decode") else ctx)
        tid_map = np.array([gindex.get(n, -1) for n in stream.names], np.int32)
        if len(tid_map) and (tid_map >= 0).all() and not (gene_len[tid_map] == 0).any() and not args.get("host_decode"):
            # the streaming ingest: the host only inflates; the record stream is parsed, re-laid and counted on the GPU
            out, _ = ctx.genes_coverage_bam(stream, tid_map, gene_len, f, exact_order=exact)
            stream.close()
        else:
            # a BAM reference that is not a gene of the selected pangenomes, or a zero-length gene: the reference raises
            # only if a record actually hits it (genes.py:180,189) - decode on the host and look
            stream.close()
            names, lens, reads = lib.read_bam(bam_path, threads)
            missing = [n for n in set(names[t] for t in np.unique(reads.ref_id))
                       if n not in gindex] if reads.n_reads else []
            if missing:
                raise KeyError(missing[0])                                   # genes[...] KeyError, genes.py:180
            tid_map = np.array([gindex.get(n, -1) for n in names], np.int32) if names else np.zeros(0, np.int32)
            if reads.n_reads and not np.array_equal(tid_map, np.arange(len(names), dtype=np.int32)):
                reads.ref_id[...] = tid_map[reads.ref_id]
            if reads.n_reads:
                hit = np.unique(reads.ref_id)
                if (gene_len[hit] == 0).any():
                    raise ZeroDivisionError("float division by zero")    # alen / float(gene.length), genes.py:189
            out = ctx.genes_coverage_plain(reads, gene_len, f, exact_order=exact)
    finally:
        if own_ctx:
            ctx.close()
    for k, g in enumerate(gene_ids):
        gene = genes[g]
        gene.aligned_reads = int(out["aligned_reads"][k])
        gene.mapped_reads = int(out["mapped_reads"][k])
        gene.depth = float(out["depth"][k])
        sp = species[gene.species_id]
        sp.aligned_reads += gene.aligned_reads
        sp.mapped_reads += gene.mapped_reads
    print("  total aligned reads: %s" % sum([sp.aligned_reads for sp in species.values()]))
    print("  total mapped reads: %s" % sum([sp.mapped_reads for sp in species.values()]))
    # per-species summaries (genes.py:191-203): mean over covered genes in dict order (np.mean of a list)
    for g in genes.values():
        species[g.species_id].genes.append(g.depth)
    for sp in species.values():
        covered = [d for d in sp.genes if d > 0]
        sp.covered_genes = len(covered)
        sp.mean_coverage = np.mean(covered) if covered else 0
        sp.fraction_covered = sp.covered_genes / float(sp.pangenome_size)


def normalize(args, species, genes):
    """Copy number = depth / median depth of the species' marker genes (reference genes.py:205-218).
    A species without marker genes gets np.median([]) = nan (with numpy's warning), as there, and its
    genes keep copies = 0.0 because nan > 0 is false."""
    for g in genes.values():
        if g.marker_id is not None:
            species[g.species_id].markers[g.marker_id] += g.depth
    for sp in species.values():
        sp.marker_coverage = np.median(list(sp.markers.values()))
    for g in genes.values():
        cov = species[g.species_id].marker_coverage
        if cov > 0:
            g.copies = g.depth / cov


GENES_HEADER = ("gene_id", "count_reads", "coverage", "copy_number")
SUMMARY_HEADER = ("species_id", "pangenome_size", "covered_genes", "fraction_covered", "mean_coverage",
                  "marker_coverage", "aligned_reads", "mapped_reads")


def _tsv(values):
    return "\t".join(str(v) for v in values) + "\n"          # numbers print with str(), as in the reference


def write_results(args, species, genes):
    """<outdir>/genes/output/<species>.genes.gz (genes in sorted id order, genes.py:229) and
    <outdir>/genes/summary.txt (genes.py:238-245)."""
    rows = {sid: [_tsv(GENES_HEADER)] for sid in species}
    for gene_id in sorted(genes):
        g = genes[gene_id]
        rows[g.species_id].append(_tsv((g.id, g.mapped_reads, g.depth, g.copies)))
    threads = int(args.get("threads", 1) or 1)
    for sid, lines in rows.items():
        # independent gzip members compressed on the host threads (gzip.open() at level 9 on one thread would be
        # the slowest step of the stage at pangenome scale)
        out = mio.GzipMembersWriter("%s/genes/output/%s.genes.gz" % (args["outdir"], sid),
                                    level=args.get("gzip_level"), threads=threads)
        for lo in range(0, len(lines), 1 << 15):
            out.write("".join(lines[lo:lo + (1 << 15)]).encode())
        out.close()
    with open("%s/genes/summary.txt" % args["outdir"], "w") as fh:
        fh.write(_tsv(SUMMARY_HEADER))
        for sp in species.values():
            fh.write(_tsv((sp.id, sp.pangenome_size, sp.covered_genes, sp.fraction_covered, sp.mean_coverage,
                           sp.marker_coverage, sp.aligned_reads, sp.mapped_reads)))


def remove_tmp(args):
    shutil.rmtree("/".join([args["outdir"], "genes/temp"]))


def run_pipeline(args):
    """reference genes.py:252-294; only the coverage stage (args['cov']) runs here."""
    print("\nReading reference data")
    start = time()
    species = initialize_species(args)
    genes = initialize_genes(args, species)
    print("  %s minutes" % round((time() - start) / 60, 2))
    if args.get("align"):
        sys.exit("\nError: the --align stage (bowtie2 | samtools) is not part of midas_b200; "
                 "run it with the reference and re-enter with --call_genes")
    if args.get("cov"):
        start = time()
        print("\nComputing coverage of pangenomes")
        args["log"].write("\nComputing coverage of pangenomes\n")
        pangenome_coverage(args, species, genes)
        print("  %s minutes" % round((time() - start) / 60, 2))
    if args.get("remove_temp"):
        remove_tmp(args)
    return species, genes
