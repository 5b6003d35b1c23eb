#!/usr/bin/env python
"""Time the reference's OWN Python pileup loop (build container only: it imports /root/reference).

midas/run/snps.py pysam_pileup -> species_pileup -> keep_read is executed verbatim
(oracle/run_reference.py) over the pysam shim on a 1/100 sample of the C2 workload (50 000 reads x 150 bp
over 40 kbp, same generator and seed as bench.py), single process, as the reference runs one species.
What is timed is pysam_pileup alone (the BAM is already written and 'indexed'; FASTA parsing, summary
writing are outside). The figure is labelled context for bench.py's cpu_baseline
(`python_reference_value`): it is the shim's restatement of pysam underneath, not htslib, so it says how
slow the per-position Python loop of snps.py:201-213 plus a Python count_coverage is, nothing more.
Writes profiles/python_reference.json.
"""
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from midas_b200 import synth                       # noqa: E402
from midas_b200.soa import Filters                 # noqa: E402
from oracle import fixtures, run_reference         # noqa: E402


def main():
    assert run_reference.available(), "needs /root/reference"
    ref, reads = synth.synth_snps(seed=synth.SEED_C2, n_species=1, contigs_per_species=1, contig_len=40_000,
                                  n_reads=50_000, read_len=150)
    with tempfile.TemporaryDirectory() as tmp:
        db, out = fixtures.write_snps_fixture(tmp, ref, reads)
        snps = run_reference._import("midas.run.snps")
        args = run_reference._args(out, db, Filters(), 1)
        species = snps.initialize_species(args)
        contigs = snps.initialize_contigs(species)
        times = []
        for _ in range(3):
            t0 = time.perf_counter()
            import contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                snps.pysam_pileup(args, species, contigs)
            times.append(time.perf_counter() - t0)
    t = min(times)
    rec = {"value": reads.total_bases / t / 1e9, "unit": "Gbases/s", "seconds": t, "cores": 1,
           "reads": reads.n_reads, "bases": reads.total_bases,
           "note": "the reference's unmodified midas/run/snps.py pysam_pileup (species_pileup + keep_read) over the "
                   "pysam shim (oracle/shim, a Python restatement of count_coverage - pysam itself is not installable "
                   "here), 50 000 x 150 bp reads on 40 kbp (C2 / 100), one process, best of 3, measured in the BUILD "
                   "container (8 cores), not on the GPU box; context only",
           "measured_by": "scripts/time_python_reference.py", "cpu_count": os.cpu_count()}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "python_reference.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
