"""Species / sample selection for the merge stage — mirror of the reference's midas/merge/merge.py.

Same names and meaning: Sample (:59-76), Species (:10-57), init_samples (:78-88), read_species_info (:90-96),
read_genome_info (:98-104), filter_sample_species (:106-122), init_species (:129-142), filter_species
(:144-159), select_species (:161-166). Pure host bookkeeping over the per-sample `summary.txt` files
that run_midas.py wrote (Appendix A of SURVEY.md); nothing here touches the GPU.
"""
from __future__ import annotations

import csv
import os

SNPS_FIELDS = ["genome_length", "covered_bases", "fraction_covered", "mean_coverage", "aligned_reads", "mapped_reads"]
GENES_FIELDS = ["pangenome_size", "covered_genes", "fraction_covered", "mean_coverage", "marker_coverage",
                "aligned_reads", "mapped_reads"]


def _table(path, key):
    with open(path) as fh:
        return {row[key]: row for row in csv.DictReader(fh, delimiter="\t")}


class Sample:
    """One run_midas.py output directory; `info` = its <dtype>/summary.txt keyed by species id, or None."""

    def __init__(self, dir, data_type):
        self.dir = dir
        self.id = os.path.basename(dir)
        path = "%s/%s/summary.txt" % (dir, data_type)
        self.info = _table(path, "species_id") if os.path.isfile(path) else None


class Species:
    def __init__(self, id, species_info, genome_info):
        self.id = id
        self.samples = []
        self.info = species_info[id]
        self.genome_info = genome_info[self.info["rep_genome"]]

    def fetch_sample_depth(self):
        self.sample_depth = [float(s.info[self.id]["mean_coverage"]) for s in self.samples]

    def write_sample_info(self, dtype, outdir):
        """<outdir>/<species>/<dtype>_summary.txt: the samples' own summary rows, verbatim."""
        fields = SNPS_FIELDS if dtype == "snps" else GENES_FIELDS
        with open("%s/%s/%s_summary.txt" % (outdir, self.id, dtype), "w") as out:
            out.write("\t".join(["sample_id"] + fields) + "\n")
            for s in self.samples:
                out.write("\t".join([s.id] + [str(s.info[self.id][f]) for f in fields]) + "\n")


def init_samples(indirs, data_type):
    return [s for s in (Sample(d, data_type) for d in indirs) if s.info is not None]


def read_species_info(db):
    return _table(os.path.join(db, "species_info.txt"), "species_id")


def read_genome_info(db):
    return _table(os.path.join(db, "genome_info.txt"), "genome_id")


def filter_sample_species(sample, species, species_id, args, dtype):
    """True = skip this (sample, species) pair; tests in the reference's order."""
    info = sample.info[species_id]
    if args["species_id"] and species_id not in args["species_id"].split(","):
        return True
    if args["max_samples"] and species_id in species and len(species[species_id].samples) >= args["max_samples"]:
        return True
    if float(info["mean_coverage"]) < args["sample_depth"]:
        return True
    if dtype == "snps" and float(info["fraction_covered"]) < args["fract_cov"]:
        return True
    return False


def init_species(samples, args, dtype):
    species = {}
    species_info, genome_info = read_species_info(args["db"]), read_genome_info(args["db"])
    for sample in samples:
        for species_id in sample.info:
            if species_id not in species:
                species[species_id] = Species(species_id, species_info, genome_info)
            if not filter_sample_species(sample, species, species_id, args, dtype):
                species[species_id].samples.append(sample)
    return list(species.values())


def filter_species(species, args):
    """Most samples first (stable), at least min_samples, at most max_species; creates <outdir>/<species>."""
    keep = []
    for sp in sorted(species, key=lambda s: len(s.samples), reverse=True):
        sp.nsamples = len(sp.samples)
        if sp.nsamples < int(args["min_samples"]) or (args["max_species"] and len(keep) >= args["max_species"]):
            continue
        sp.fetch_sample_depth()
        sp.outdir = args["outdir"] + "/" + sp.id
        keep.append(sp)
        os.makedirs(sp.outdir, exist_ok=True)
    return keep


def select_species(args, dtype):
    return filter_species(init_species(init_samples(args["indirs"], dtype), args, dtype), args)
