#!/usr/bin/env python
"""Generate tests/golden/: outputs of the reference's UNMODIFIED hot-path functions on small seeded inputs.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference's midas/run/snps.py and midas/run/genes.py are imported from /root/reference and
executed verbatim over oracle/shim (pysam / Bio.SeqIO stand-ins, see oracle/shim/README.md) on
BAM + database fixtures written by oracle/fixtures.py. What is committed per case:
  input.npz      the decoded alignments and reference (numpy arrays; names / sequences as JSON)
  expected/      every file the reference wrote: <species>.snps.gz / <species>.genes.gz (stored
                 decompressed as .tsv) and summary.txt
The pysam half of the semantics is a restatement (the shim), so these vectors pin the MIDAS-owned
half (filters, loop order, stats, text) and the shim <-> C-oracle agreement; see DESIGN.md "Oracle".
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from midas_b200 import synth                     # noqa: E402
from midas_b200.soa import Filters               # noqa: E402
from oracle import fixtures, run_reference       # noqa: E402
import gen                                       # noqa: E402

READ_FIELDS = ["ref_id", "pos", "mapq", "nm", "l_qseq", "cigar_off", "cigar", "seq_off", "seq", "qual", "flag"]


def save_input(path, ref, reads, filters, extra=None):
    meta = dict(contig_names=ref.contig_names, species_ids=ref.species_ids,
                contig_seq=[s.decode() for s in ref.contig_seq] if ref.contig_seq else None,
                filters=dict(mapid=filters.mapid, aln_cov=filters.aln_cov, readq=filters.readq, mapq=filters.mapq,
                             baseq=filters.baseq), extra=extra or {})
    arrays = {k: getattr(reads, k) for k in READ_FIELDS}
    arrays["contig_len"] = ref.contig_len
    arrays["contig_species"] = ref.contig_species
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(path, **arrays)


def collect(outdir, sub, dst):
    os.makedirs(dst, exist_ok=True)
    shutil.copy(os.path.join(outdir, sub, "summary.txt"), os.path.join(dst, "summary.txt"))
    od = os.path.join(outdir, sub, "output")
    for fn in sorted(os.listdir(od)):
        with gzip.open(os.path.join(od, fn), "rb") as fh:
            data = fh.read()
        with open(os.path.join(dst, fn[:-3] + ".tsv"), "wb") as out:
            out.write(data)


def snps_case(name, ref, reads, filters, lowercase_some=False):
    dst = os.path.join(HERE, name)
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    tmp = tempfile.mkdtemp()
    db, out = fixtures.write_snps_fixture(tmp, ref, reads, lowercase_some=lowercase_some)
    run_reference.run_snps(out, db, filters, threads=1)
    save_input(os.path.join(dst, "input.npz"), ref, reads, filters, dict(kind="snps", lowercase_some=lowercase_some))
    collect(out, "snps", os.path.join(dst, "expected"))
    shutil.rmtree(tmp)


def genes_case(name, ref, reads, marker, filters):
    dst = os.path.join(HERE, name)
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    tmp = tempfile.mkdtemp()
    db, out = fixtures.write_genes_fixture(tmp, ref, reads, marker)
    run_reference.run_genes(out, db, filters)
    save_input(os.path.join(dst, "input.npz"), ref, reads, filters, dict(kind="genes", marker=[int(x) for x in marker]))
    collect(out, "genes", os.path.join(dst, "expected"))
    shutil.rmtree(tmp)


def merge_case(name, snp_freq, site_depth, site_ratio):
    """Pooled allele call (merge_midas.py snps): the reference's unmodified GenomicSite per site."""
    from oracle import merge_oracle as mo
    counts, md = mo.synth_counts(5, 7, 600)
    counts[:, :, 0] = 0                                   # no reads at all
    counts[:, :, 1] = 3                                   # four-way tie
    counts[:, :, 2] = 0
    counts[0, 2, 2] = 5                                   # one allele in one sample
    counts[:, :, 3] = [[4, 4, 0, 0]] * 7                  # A/C tie
    ref = mo.run_reference_sites(counts, md, snp_freq, site_depth, site_ratio)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), counts=counts, mean_depth=md, snp_freq=snp_freq,
                        site_depth=site_depth, site_ratio=site_ratio,
                        **{("exp_" + k): v for k, v in ref.items() if k != "freq_text"},
                        freq_text=np.array(ref["freq_text"]))


def main():
    # 1. C2-shaped reads (90 % full matches, indels, clips), two species, default snps filters
    ref, reads = synth.synth_snps(seed=20260102, n_species=2, contigs_per_species=2, contig_len=2500, n_reads=2500,
                                  read_len=100, len_jitter=0.2, abundance="lognormal")
    snps_case("snps_c2_small", ref, reads, Filters(), lowercase_some=True)
    # 2. arbitrary CIGARs (hard/soft clips, =/X, N skips, long deletions, overhang), non-default filters
    rng = np.random.default_rng(77)
    ref2 = gen.make_ref(rng, 2, 2, 300, 2500)
    reads2 = gen.random_reads(78, ref2, 1200, max_len=160)
    snps_case("snps_random_cigars", ref2, reads2, Filters(mapid=90.0, aln_cov=0.6, readq=22, mapq=10, baseq=25))
    # 3. same reads, every filter disabled (baseq 0 switches the per-base test off in pysam)
    snps_case("snps_no_filters", ref2, reads2, Filters(mapid=0.0, aln_cov=0.0, readq=0, mapq=0, baseq=0))
    # 4. genes: three species, unsorted reads, one species without marker genes (np.median([]) -> nan)
    gref, greads, marker = synth.synth_genes(seed=20260103, n_species=3, genes_per_species=40, n_reads=3000,
                                             read_len=100, n_markers=5)
    marker = marker.copy()
    marker[np.asarray(gref.contig_species) == 2] = -1
    genes_case("genes_c3_small", gref, greads, marker, Filters.genes_defaults())
    # 5. merge_midas.py snps site arithmetic, default and strict parameters
    merge_case("merge_small", 0.01, 1, 2.0)
    merge_case("merge_small_strict", 0.2, 5, 1.2)
    print("golden vectors written under", HERE)


if __name__ == "__main__":
    main()
