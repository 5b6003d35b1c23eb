#!/usr/bin/env python
"""Gene-coverage kernel timing (BASELINE.json configs[2] shape, scaled to fit host generation time):
reads resident in HBM, CUDA-event time of mb_genes_coverage_dev, algorithmic bytes of SURVEY 8d."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=5_000_000)
    ap.add_argument("--species", type=int, default=100)
    ap.add_argument("--genes", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    import torch
    from midas_b200 import lib, synth
    from midas_b200.soa import Filters
    gref, greads, _ = synth.synth_genes(n_species=a.species, genes_per_species=a.genes, n_reads=a.reads)
    lay = lib.Layout.of(gref)
    pk = lib.PackedReads(greads, lay, "genes", pinned=True)
    ctx = lib.Context(0)
    dr = ctx.upload(pk)
    n_genes = gref.n_contigs
    glen = torch.as_tensor(gref.contig_len.astype(np.int32), device="cuda")
    out = torch.zeros((n_genes, 3), dtype=torch.int64, device="cuda")
    f = Filters.genes_defaults()
    res = {}
    for exact in (False, True):
        for _ in range(3):
            ctx.genes_coverage_dev(dr, glen, n_genes, f, out, exact_order=exact)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.timing_reset()
        e0.record()
        for _ in range(a.steps):
            ctx.genes_coverage_dev(dr, glen, n_genes, f, out, exact_order=exact)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        kms, kn = ctx.timing_read()
        res["exact_order" if exact else "fast"] = dict(ms_per_step=ms, kernel_ms=kms / max(1, kn),
                                                       gbases_per_s=pk.total_bases / ms / 1e6)
    n_cig = (greads.cigar_off[1:] - greads.cigar_off[:-1]).astype(np.int64)
    alg = int((16 + 4 * n_cig + greads.l_qseq.astype(np.int64)).sum() + 24 * n_genes)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    k = res["fast"]["kernel_ms"]
    print(json.dumps(dict(workload=f"genes: {a.species} species x {a.genes} genes, {a.reads} x 150 bp reads (file order)",
                          algorithmic_bytes=alg, results=res,
                          roofline=dict(kernel="k_genes_v2", achieved_gbs=alg / k / 1e6, peak_gbs=peak,
                                        frac=alg / k / 1e6 / peak))))


if __name__ == "__main__":
    main()
