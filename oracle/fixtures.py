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


def write_merge_fixture(root, seed=20260105, n_samples=6):
    """A database and run_midas.py sample directories for `merge_midas.py snps` (reference
    midas/merge/merge.py:59-104, midas/merge/snps.py:225-259, midas/utility.py:244-294):
      <db>/{species_info.txt,genome_info.txt}, <db>/rep_genomes/<sp>/{genome.fna.gz,genome.features}
      <samples>/<id>/snps/{summary.txt,output/<sp>.snps.gz}
    Two species; genes on both strands, overlapping and nested genes, a gene whose length is not a
    multiple of three, a gene over an N, a non-coding gene; one sample below the depth cut for one
    species, one sample directory without snps output. Returns (db, [sample dirs])."""
    rng = np.random.default_rng(seed)
    db, sroot = os.path.join(root, "db"), os.path.join(root, "samples")
    genomes = {"Sp_alpha_1": [("alpha.c1", 1500), ("alpha.c2", 900)], "Sp_beta_2": [("beta.c1", 1200)]}
    features = {
        "Sp_alpha_1": [("a1", "alpha.c1", 31, 330, "+", "CDS"), ("a2", "alpha.c1", 301, 600, "-", "CDS"),
                       ("a3", "alpha.c1", 350, 409, "+", "CDS"), ("a4", "alpha.c1", 700, 800, "+", "CDS"),
                       ("a5", "alpha.c1", 900, 1019, "-", "CDS"), ("a6", "alpha.c1", 1100, 1180, "+", "tRNA"),
                       ("a7", "alpha.c2", 10, 309, "+", "CDS"), ("a8", "alpha.c2", 600, 899, "-", "CDS")],
        "Sp_beta_2": [("b1", "beta.c1", 100, 399, "-", "CDS"), ("b2", "beta.c1", 500, 799, "+", "CDS")],
    }
    os.makedirs(db, exist_ok=True)
    with open(os.path.join(db, "species_info.txt"), "w") as fh:
        fh.write("species_id\trep_genome\tcount_genomes\n")
        for k, sp in enumerate(genomes):
            fh.write(f"{sp}\tG{k}\t1\n")
    with open(os.path.join(db, "genome_info.txt"), "w") as fh:
        fh.write("genome_id\tgenome_name\tlength\tcontigs\n")
        for k, (sp, ctgs) in enumerate(genomes.items()):
            fh.write(f"G{k}\t{sp} strain\t{sum(l for _, l in ctgs)}\t{len(ctgs)}\n")
    seqs = {}
    for sp, ctgs in genomes.items():
        d = os.path.join(db, "rep_genomes", sp)
        os.makedirs(d, exist_ok=True)
        with gzip.open(os.path.join(d, "genome.fna.gz"), "wt") as fh:
            for cid, ln in ctgs:
                s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, ln)].tobytes().decode()
                if cid == "alpha.c1":
                    s = s[:949] + "N" + s[950:]                  # gene a5 runs over an N
                seqs[cid] = s
                fh.write(f">{cid} synthetic\n{_wrap(s.lower() if cid == 'beta.c1' else s)}\n")
        with open(os.path.join(d, "genome.features"), "w") as fh:
            fh.write("gene_id\tscaffold_id\tstart\tend\tstrand\tgene_type\tfunctions\n")
            for g in features[sp]:
                fh.write("\t".join(str(x) for x in g) + "\t\n")
    dirs = []
    for s in range(n_samples):
        sdir = os.path.join(sroot, f"sample_{s:02d}")
        dirs.append(sdir)
        if s == n_samples - 1:                                   # a directory without snps output is skipped
            os.makedirs(os.path.join(sdir, "genes"), exist_ok=True)
            continue
        os.makedirs(os.path.join(sdir, "snps", "output"), exist_ok=True)
        rows = []
        for sp, ctgs in genomes.items():
            mean = float(rng.uniform(8, 30)) if not (sp == "Sp_beta_2" and s == 1) else 1.5
            tot, cov, glen = 0, 0, 0
            with gzip.open(os.path.join(sdir, "snps", "output", sp + ".snps.gz"), "wt") as fh:
                fh.write("ref_id\tref_pos\tref_allele\tdepth\tcount_a\tcount_c\tcount_g\tcount_t\n")
                for cid, ln in sorted(ctgs):
                    ref = seqs[cid]
                    # site-specific minor alleles shared by the samples (seeded per contig), plus noise
                    srng = np.random.default_rng(sum(map(ord, cid)))
                    is_snp = srng.random(ln) < 0.08
                    minor = srng.integers(0, 4, ln)
                    hot = srng.random(ln) < 0.01                 # abnormally deep sites (site_ratio filter)
                    d = rng.poisson(mean, ln)
                    d[rng.random(ln) < 0.05] = 0
                    d = np.where(hot, d * 4, d)
                    for i in range(ln):
                        c = [0, 0, 0, 0]
                        r = "ACGT".find(ref[i])
                        maj = r if r >= 0 else int(minor[i])
                        nmin = int(rng.binomial(d[i], 0.3)) if is_snp[i] else int(rng.binomial(d[i], 0.003))
                        c[maj] += int(d[i]) - nmin
                        c[int(minor[i])] += nmin
                        dep = sum(c)
                        fh.write(f"{cid}\t{i + 1}\t{ref[i]}\t{dep}\t{c[0]}\t{c[1]}\t{c[2]}\t{c[3]}\n")
                        tot += dep; cov += dep > 0; glen += 1
            rows.append((sp, glen, cov, cov / float(glen), tot / float(cov) if cov else 0.0, 1000 + s, 900 + s))
        with open(os.path.join(sdir, "snps", "summary.txt"), "w") as fh:
            fh.write("species_id\tgenome_length\tcovered_bases\tfraction_covered\tmean_coverage\taligned_reads\tmapped_reads\n")
            for r in rows:
                fh.write("\t".join(str(x) for x in r) + "\n")
    return db, dirs
