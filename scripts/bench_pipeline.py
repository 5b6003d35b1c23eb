#!/usr/bin/env python
"""Stage times of the whole `run_midas.py snps --pileup` stage, from a coordinate-sorted BAM on disk to
<species>.snps.gz + summary.txt, on a C2-shaped workload (scaled by --scale; the BAM writer used to make the fixture
is pure Python). The stage replaces reference midas/run/snps.py:299-302 (index_bam, pysam_pileup, snps_summary); its
steps here are: BGZF inflate on the host threads with the inflated record stream crossing PCIe behind the inflate
front and the record walk trailing it (mb_bam_stream_open_dev), one mb_snps_pileup_bam call (records parsed and
re-laid on the GPU, pileup, D2H), native row formatting + threaded multi-member gzip. --host-decode times the
round-1 path (records decoded on the host, mb_snps_pileup_plain) instead.

  python scripts/bench_pipeline.py [--scale 0.1] [--threads 16]

One JSON line: seconds per stage (of the last of PIPE_REPS passes, default 2), bases per second from BAM to the
finished files."""
import argparse
import gzip
import io
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_fixture(root, ref, reads):
    from midas_b200 import bam
    db, out = os.path.join(root, "db"), os.path.join(root, "sample")
    for si, sp in enumerate(ref.species_ids):
        d = os.path.join(db, "rep_genomes", sp)
        os.makedirs(d, exist_ok=True)
        with gzip.open(os.path.join(d, "genome.fna.gz"), "wt", compresslevel=1) as fh:
            for c in range(ref.n_contigs):
                if ref.contig_species[c] == si:
                    fh.write(">%s\n%s\n" % (ref.contig_names[c], ref.contig_seq[c].decode()))
    for sub in ("snps/output", "snps/temp"):
        os.makedirs(os.path.join(out, sub), exist_ok=True)
    with open(os.path.join(out, "snps", "species.txt"), "w") as fh:
        fh.write("".join(sp + "\n" for sp in ref.species_ids))
    bam.write_bam(os.path.join(out, "snps", "temp", "genomes.bam"), ref.contig_names, ref.contig_len, reads, level=1)
    return db, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--host-decode", action="store_true", help="decode the BAM records on the host (the round-1 path)")
    a = ap.parse_args()
    from midas_b200 import synth
    from midas_b200.run import snps
    n_reads, clen = int(5_000_000 * a.scale), int(4_000_000 * a.scale)
    ref, reads = synth.synth_snps(seed=synth.SEED_C2, n_species=1, contigs_per_species=1, contig_len=clen,
                                  n_reads=n_reads, read_len=150, keep_ref_seq=True)
    tmp = tempfile.mkdtemp()
    try:
        db, out = write_fixture(tmp, ref, reads)
        bam_bytes = os.path.getsize(os.path.join(out, "snps", "temp", "genomes.bam"))
        args = {"outdir": out, "db": db, "threads": a.threads, "mapid": 94.0, "readq": 20, "mapq": 20, "aln_cov": 0.75,
                "baseq": 30, "build_db": False, "align": False, "call": True, "remove_temp": False,
                "log": io.StringIO(), "species_id": None, "device": 0, "host_decode": a.host_decode}
        best = None
        for rep in range(int(os.environ.get("PIPE_REPS", "2"))):   # first pass pays CUDA context creation and page-cache misses
            timing = {}
            args["timing"] = timing
            with open(os.devnull, "w") as devnull:
                saved, sys.stdout = sys.stdout, devnull
                try:
                    t0 = time.time()
                    snps.run_pipeline(args)
                    timing["total"] = time.time() - t0
                finally:
                    sys.stdout = saved
            best = timing
        gz = sum(os.path.getsize(os.path.join(out, "snps", "output", f)) for f in os.listdir(os.path.join(out, "snps", "output")))
        stages = {k: round(v, 4) for k, v in best.items() if k not in ("n_reads", "n_bases")}
        print(json.dumps({"workload": f"run_midas.py snps --pileup: {n_reads} x 150 bp reads on {clen} bp (C2 x {a.scale:g}), "
                                      f"BAM {bam_bytes / 1e6:.0f} MB -> snps.gz {gz / 1e6:.0f} MB",
                          "host_threads": a.threads, "seconds": stages,
                          "gbases_per_s_bam_to_files": best["n_bases"] / best["total"] / 1e9,
                          "gbases_per_s_gpu_call": best["n_bases"] / best.get("gpu_h2d_parse_pack_kernels_d2h", best.get("gpu_h2d_pack_kernels_d2h")) / 1e9}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
