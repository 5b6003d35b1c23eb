"""Pooled allele call across samples (reference midas/merge/snps.py GenomicSite).
CPU: the numpy oracle reproduces the outputs of the reference's unmodified class (tests/golden/merge_*.npz).
GPU: the CUDA kernel (through the C-ABI) equals the oracle bit for bit (integers, and the doubles too:
the same correctly-rounded divisions), including the '{0:.3g}' text of snps_freq.txt."""
import os

import numpy as np
import pytest

from oracle import merge_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["merge_small", "merge_small_strict"]


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, float(z["snp_freq"]), int(z["site_depth"]), float(z["site_ratio"])


@pytest.mark.parametrize("case", CASES)
def test_numpy_oracle_matches_reference_genomic_site(case):
    z, f, d, r = _load(case)
    got = mo.merge_sites(z["counts"], z["mean_depth"], f, d, r)
    for k in ("pooled", "major", "minor", "snp_type", "count_samples", "sample_depth"):
        assert np.array_equal(got[k], z["exp_" + k]), k
    assert np.array_equal(got["sample_maf"].view(np.uint64), z["exp_sample_maf"].view(np.uint64))


def _gpu_check(ctx, counts, mean_depth, f, d, r, exp):
    import torch
    from midas_b200 import lib  # noqa: F401
    t = torch.as_tensor(counts.astype(np.int64).astype(np.uint32).view(np.int32), device="cuda").contiguous()
    out = ctx.merge_sites(t, mean_depth, snp_freq=f, site_depth=d, site_ratio=r)
    torch.cuda.synchronize()
    call = out["call"].cpu().numpy().astype(np.int16)
    pooled = out["pooled"].cpu().numpy().view(np.uint32).astype(np.uint64)
    assert np.array_equal(pooled, exp["pooled"])
    has_major = pooled.sum(0) > 0
    major = np.where(has_major, call & 3, -1)
    minor = np.where((call & 0x80) != 0, (call >> 2) & 3, -1)
    assert np.array_equal(major, exp["major"]) and np.array_equal(minor, exp["minor"])
    assert np.array_equal((call >> 4) & 7, exp["snp_type"])
    assert np.array_equal(out["count_samples"].cpu().numpy(), exp["count_samples"])
    assert np.array_equal(out["sample_depth"].cpu().numpy().view(np.uint32).astype(np.uint64), exp["sample_depth"])
    maf = out["sample_maf"].cpu().numpy()
    assert np.array_equal(maf.view(np.uint64), exp["sample_maf"].view(np.uint64))
    return maf


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_kernel_matches_golden(ctx, case):
    z, f, d, r = _load(case)
    exp = {k[4:]: z[k] for k in z.files if k.startswith("exp_")}
    maf = _gpu_check(ctx, z["counts"], z["mean_depth"], f, d, r, exp)
    text = ["\t".join("{0:.3g}".format(x) for x in maf[:, g]) for g in range(maf.shape[1])]   # snps.py:197
    assert text == [str(t) for t in z["freq_text"]]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,S,G,depth", [(1, 3, 10_000, 30.0), (2, 20, 50_001, 30.0), (3, 1, 777, 30.0),
                                             (4, 64, 4096, 30.0), (5, 40, 3000, 180.0), (6, 100, 2000, 60.0),
                                             (7, 33, 1500, 400.0), (8, 129, 1500, 20.0), (9, 200, 1200, 25.0),
                                             (10, 129, 900, 300.0), (11, 200, 700, 350.0)])
def test_kernel_matches_oracle(ctx, seed, S, G, depth):
    """depth 180-400: many sites hold counts above 255 and take the re-read path instead of the packed (8-bit)
    registers. S = 33..128: samples past the 32 register-resident ones live in the shared-memory column;
    S = 129, 200: the samples past the 128th take the separate re-read branch of both passes (low depth = packed
    sites, high depth = wide sites)."""
    counts, md = mo.synth_counts(seed, S, G, depth_mean=depth, snp_frac=0.2)
    exp = mo.merge_sites(counts, md, 0.05, 3, 1.5)
    _gpu_check(ctx, counts, md, 0.05, 3, 1.5, exp)


@pytest.mark.gpu
def test_zero_mean_depth_is_rejected(ctx):
    import torch
    from midas_b200 import lib
    t = torch.zeros((2, 4, 8), dtype=torch.int32, device="cuda")
    with pytest.raises(lib.MidasB200Error):
        ctx.merge_sites(t, [1.0, 0.0])
