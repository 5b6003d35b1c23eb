#!/usr/bin/env python
"""Pooled allele call timing (BASELINE.json configs[4] shape, one species): S samples x G sites of
count planes resident in HBM, CUDA-event time of mb_merge_sites_dev, algorithmic bytes
16 S (counts in) + 12 S (depth + maf out) + 25 (pooled, call, count_samples) per site."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=32)
    ap.add_argument("--sites", type=int, default=4_000_000)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    import torch
    from midas_b200 import lib
    S, G = a.samples, a.sites
    g = torch.Generator(device="cuda").manual_seed(20260105)
    counts = torch.zeros((S, 4, G), dtype=torch.int32, device="cuda")
    major = torch.randint(0, 4, (G,), device="cuda", generator=g)
    depth = torch.poisson(torch.full((S, G), 20.0, device="cuda"), generator=g).to(torch.int32)
    minor_n = (torch.rand((S, G), device="cuda", generator=g) < 0.01).to(torch.int32) * (depth // 3)
    counts.scatter_(1, major.view(1, 1, G).expand(S, 1, G), (depth - minor_n).unsqueeze(1))
    counts.scatter_add_(1, ((major + 1) % 4).view(1, 1, G).expand(S, 1, G), minor_n.unsqueeze(1))
    md = counts.sum(1).double().mean(1).clamp_min(0.5)
    ctx = lib.Context(0)
    for _ in range(3):
        out = ctx.merge_sites(counts, md)
    torch.cuda.synchronize()
    ctx.timing_reset()
    for _ in range(a.steps):
        out = ctx.merge_sites(counts, md)
    torch.cuda.synchronize()
    kms, kn = ctx.timing_read()
    k = kms / max(1, kn)
    alg = G * (16 * S + 12 * S + 25)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(json.dumps(dict(workload=f"merge snps: {S} samples x {G} sites", kernel="k_merge_sites", kernel_ms=k,
                          sites_per_s=G / k * 1e3, algorithmic_bytes=alg, achieved_gbs=alg / k / 1e6, peak_gbs=peak,
                          frac=alg / k / 1e6 / peak, snp_sites=int(((out["call"] >> 4) & 7 >= 2).sum()))))


if __name__ == "__main__":
    main()
