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


class Species:
    """reference genes.py:12-30"""

    def __init__(self, id):
        self.id = id
        self.paths = {}
        self.genes = []
        self.pangenome_size = 0
        self.aligned_reads = 0
        self.mapped_reads = 0
        self.markers = defaultdict(float)

    def init_ref_db(self, ref_db):
        self.dir = "%s/pan_genomes/%s" % (ref_db, self.id)
        for ext in ["", ".gz"]:
            for file in ["centroids.ffn", "cluster_info.txt", "gene_info.txt"]:
                inpath = "%s/%s%s" % (self.dir, file, ext)
                if os.path.isfile(inpath):
                    self.paths[file] = inpath


class Gene:
    """reference genes.py:50-59"""

    def __init__(self, id):
        self.id = id
        self.aligned_reads = 0
        self.mapped_reads = 0
        self.depth = 0.0
        self.length = 0
        self.copies = 0.0
        self.marker_id = None


def initialize_species(args):
    species = {}
    splist = "%s/genes/species.txt" % args["outdir"]
    if args.get("build_db"):
        sys.exit("\nError: the --build_db stage is not part of midas_b200; run it with the reference and "
                 "re-enter with --call_genes")
    elif os.path.isfile(splist):
        for line in open(splist):
            id = line.rstrip()
            species[id] = Species(id)
    for sp in species.values():
        sp.init_ref_db(args["db"])
    return species


def initialize_genes(args, species):
    """reference genes.py:61-82"""
    genes = {}
    for sp in species.values():
        for seq_id, seq in mio.read_fasta(sp.paths["centroids.ffn"]):
            genes[seq_id] = Gene(seq_id)
            genes[seq_id].species_id = sp.id
            genes[seq_id].length = len(seq)
            sp.pangenome_size += 1
    path = "%s/marker_genes/phyeco.map" % args["db"]
    file = mio.iopen(path)
    reader = csv.DictReader(file, delimiter="\t")
    for r in reader:
        if r["gene_id"] in genes:
            genes[r["gene_id"]].marker_id = r["marker_id"]
    file.close()
    return genes


def pangenome_coverage(args, species, genes, ctx=None):
    count_mapped_bp(args, species, genes, ctx=ctx)
    normalize(args, species, genes)
    write_results(args, species, genes)


def count_mapped_bp(args, species, genes, ctx=None):
    """reference genes.py:171-203: per gene aligned/mapped reads and depth, per species summaries."""
    bam_path = "/".join([args["outdir"], "genes/temp/pangenomes.bam"])
    names, lens, reads = lib.read_bam(bam_path)
    gene_ids = list(genes.keys())
    gindex = {g: k for k, g in enumerate(gene_ids)}
    missing = [n for n in set(names[t] for t in np.unique(reads.ref_id))
               if n not in gindex] if reads.n_reads else []
    if missing:
        raise KeyError(missing[0])                                   # genes[...] KeyError, genes.py:180
    tid_map = np.array([gindex.get(n, -1) for n in names], np.int32) if names else np.zeros(0, np.int32)
    if reads.n_reads:
        reads.ref_id = np.ascontiguousarray(tid_map[reads.ref_id], np.int32)
    sp_ids = list(species.keys())
    sp_index = {s: k for k, s in enumerate(sp_ids)}
    gene_len = np.array([genes[g].length for g in gene_ids], np.uint32)
    ref = RefSet(gene_ids, np.maximum(gene_len, 1), np.array([sp_index[genes[g].species_id] for g in gene_ids],
                                                             np.int32), sp_ids, None)
    own_ctx = ctx is None
    if own_ctx:
        ctx = lib.Context(int(args.get("device", 0)), use_torch_stream=False)
    try:
        layout = lib.Layout.of(ref)
        packed = lib.PackedReads(reads, layout, "genes")
        f = Filters(float(args["mapid"]), float(args["aln_cov"]), int(args["readq"]), int(args["mapq"]), 0)
        out = ctx.genes_coverage(packed, gene_len, f, exact_order=bool(args.get("exact_order", True)))
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
    for gene in genes.values():
        species[gene.species_id].genes.append(gene.depth)
    for sp in species.values():
        non_zero_genes = [_ for _ in sp.genes if _ > 0]
        sp.covered_genes = len(non_zero_genes)
        sp.mean_coverage = np.mean(non_zero_genes) if len(non_zero_genes) > 0 else 0
        sp.fraction_covered = sp.covered_genes / float(sp.pangenome_size)


def normalize(args, species, genes):
    """reference genes.py:205-218 (np.median([]) -> nan for a species without marker genes, as there)."""
    for gene in genes.values():
        if gene.marker_id is not None:
            species[gene.species_id].markers[gene.marker_id] += gene.depth
    for sp in species.values():
        sp.marker_coverage = np.median(list(sp.markers.values()))
    for gene in genes.values():
        sp = species[gene.species_id]
        if sp.marker_coverage > 0:
            gene.copies = gene.depth / sp.marker_coverage


def write_results(args, species, genes):
    """reference genes.py:220-245"""
    header = ["gene_id", "count_reads", "coverage", "copy_number"]
    for sp in species.values():
        path = "/".join([args["outdir"], "genes/output/%s.genes.gz" % sp.id])
        sp.out = mio.iopen(path, "w")
        sp.out.write("\t".join(header) + "\n")
    for gene_id in sorted(genes):
        gene = genes[gene_id]
        sp = species[gene.species_id]
        values = [gene.id, gene.mapped_reads, gene.depth, gene.copies]
        sp.out.write("\t".join([str(_) for _ in values]) + "\n")
    for sp in species.values():
        sp.out.close()
    path = "/".join([args["outdir"], "genes/summary.txt"])
    file = open(path, "w")
    header = ["species_id", "pangenome_size", "covered_genes", "fraction_covered", "mean_coverage",
              "marker_coverage", "aligned_reads", "mapped_reads"]
    file.write("\t".join(header) + "\n")
    for sp in species.values():
        values = [sp.id, sp.pangenome_size, sp.covered_genes, sp.fraction_covered, sp.mean_coverage,
                  sp.marker_coverage, sp.aligned_reads, sp.mapped_reads]
        file.write("\t".join([str(_) for _ in values]) + "\n")
    file.close()


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
