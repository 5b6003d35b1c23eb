"""On-disk fixtures in the reference's own layout (TEST INFRASTRUCTURE).

Writes a miniature MIDAS database and sample directory so that both the reference's
unmodified run_pipeline stages and this repo's drop-in can be pointed at the same files:
  <db>/rep_genomes/<sp>/genome.fna.gz              (reference midas/run/snps.py:25-31)
  <db>/pan_genomes/<sp>/centroids.ffn.gz           (midas/run/genes.py:23-30)
  <db>/marker_genes/phyeco.map                     (midas/run/genes.py:75-81)
  <db>/species_info.txt
  <out>/snps/{species.txt,temp/genomes.fa,temp/genomes.bam,output/}
  <out>/genes/{species.txt,temp/pangenomes.bam,output/}
"""
from __future__ import annotations

import gzip
import os

import numpy as np

from midas_b200 import bam


def _wrap(seq: str, width=70):
    return "\n".join(seq[i:i + width] for i in range(0, len(seq), width))


def write_snps_fixture(root, ref, reads, lowercase_some=False):
    """ref.contig_seq must be present. Returns (db, outdir)."""
    db, out = os.path.join(root, "db"), os.path.join(root, "sample")
    for d in ("marker_genes", "pan_genomes", "rep_genomes"):
        os.makedirs(os.path.join(db, d), exist_ok=True)
    with open(os.path.join(db, "species_info.txt"), "w") as fh:
        fh.write("species_id\trep_genome\tcount_genomes\n")
        for sp in ref.species_ids:
            fh.write(f"{sp}\tg1\t1\n")
    for si, sp in enumerate(ref.species_ids):
        d = os.path.join(db, "rep_genomes", sp)
        os.makedirs(d, exist_ok=True)
        with gzip.open(os.path.join(d, "genome.fna.gz"), "wt") as fh:
            for c in range(ref.n_contigs):
                if ref.contig_species[c] != si:
                    continue
                s = ref.contig_seq[c].decode()
                if lowercase_some:          # reference upper-cases (snps.py:62)
                    s = s[: len(s) // 3].lower() + s[len(s) // 3:]
                fh.write(f">{ref.contig_names[c]} some description\n{_wrap(s)}\n")
    for sub in ("snps/output", "snps/temp"):
        os.makedirs(os.path.join(out, sub), exist_ok=True)
    with open(os.path.join(out, "snps", "species.txt"), "w") as fh:
        for sp in ref.species_ids:
            fh.write(sp + "\n")
    with open(os.path.join(out, "snps", "temp", "genomes.fa"), "w") as fh:
        for c in range(ref.n_contigs):
            fh.write(f">{ref.contig_names[c]}\n{ref.contig_seq[c].decode()}\n")
    bam.write_bam(os.path.join(out, "snps", "temp", "genomes.bam"), ref.contig_names, ref.contig_len, reads)
    return db, out


def write_genes_fixture(root, ref, reads, marker, seed=7):
    """ref: genes as contigs; marker: int32 [n_genes] marker index or -1. Returns (db, outdir)."""
    db, out = os.path.join(root, "db"), os.path.join(root, "sample")
    for d in ("marker_genes", "pan_genomes", "rep_genomes"):
        os.makedirs(os.path.join(db, d), exist_ok=True)
    with open(os.path.join(db, "species_info.txt"), "w") as fh:
        fh.write("species_id\trep_genome\tcount_genomes\n")
        for sp in ref.species_ids:
            fh.write(f"{sp}\tg1\t1\n")
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for si, sp in enumerate(ref.species_ids):
        d = os.path.join(db, "pan_genomes", sp)
        os.makedirs(d, exist_ok=True)
        with gzip.open(os.path.join(d, "centroids.ffn.gz"), "wt") as fh:
            for g in range(ref.n_contigs):
                if ref.contig_species[g] != si:
                    continue
                s = acgt[rng.integers(0, 4, int(ref.contig_len[g]))].tobytes().decode()
                fh.write(f">{ref.contig_names[g]}\n{_wrap(s)}\n")
    with open(os.path.join(db, "marker_genes", "phyeco.map"), "w") as fh:
        fh.write("species_id\tgenome_id\tgene_id\tgene_length\tmarker_id\n")
        for g in range(ref.n_contigs):
            if marker[g] >= 0:
                sp = ref.species_ids[ref.contig_species[g]]
                fh.write(f"{sp}\tg1\t{ref.contig_names[g]}\t{int(ref.contig_len[g])}\tB{int(marker[g]):06d}\n")
        # a marker row for a gene that is not in any selected pangenome (must be ignored, genes.py:79)
        fh.write("Other_species_1\tg9\tOther_species_1.g000001\t900\tB000001\n")
    for sub in ("genes/output", "genes/temp"):
        os.makedirs(os.path.join(out, sub), exist_ok=True)
    with open(os.path.join(out, "genes", "species.txt"), "w") as fh:
        for sp in ref.species_ids:
            fh.write(sp + "\n")
    bam.write_bam(os.path.join(out, "genes", "temp", "pangenomes.bam"), ref.contig_names, ref.contig_len, reads)
    return db, out
