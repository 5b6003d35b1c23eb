#!/usr/bin/env python
"""Time the pileup kernel on the C2 workload under a list of environment settings (experiment switches
of csrc/pileup.cu) and check every setting's counts bit for bit against the first one.

  python scripts/sweep_pileup.py "MB_PILE_VAR=0" "MB_PILE_VAR=1" "MB_PILE_VAR=3 MB_PILE_SMEM_KB=96"

One JSON line per setting: kernel ms (CUDA events recorded by the library around the launch), agreement."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from midas_b200 import lib  # noqa: E402
from midas_b200.soa import Filters  # noqa: E402

KNOBS = ("MB_PILE_VAR", "MB_PILE_SMEM_KB", "MB_PILE_RAGGED", "MB_PILE_TMA", "MB_PILE_UNIT", "MB_PILE_GUIDE", "MB_PILE_GENERIC_BASEQ", "MB_PILE_SMEM_B", "MB_PILE_FULL_BLOCKS")


def main():
    settings = sys.argv[1:] or ["MB_PILE_VAR=0"]
    steps = int(os.environ.get("SWEEP_STEPS", "20"))
    ref, reads = bench.make_workload(0, float(os.environ.get("SWEEP_SCALE", "1.0")))
    lay = lib.Layout.of(ref)
    packed = lib.PackedReads(reads, lay, "snps", sort=False, pinned=False)
    alg = bench.algorithmic_bytes(reads, ref)
    dev = torch.device("cuda", 0)
    base = None
    for setting in settings:
        for k in KNOBS:
            os.environ.pop(k, None)
        for kv in setting.split():
            k, v = kv.split("=")
            os.environ[k] = v
        ctx = lib.Context(0)               # a fresh context: shared-memory size and occupancy are per context
        dreads = ctx.upload(packed)
        counts = torch.empty((4, lay.gpad), dtype=torch.int32, device=dev)
        stats = torch.zeros((1, 5), dtype=torch.int64, device=dev)
        for _ in range(3):
            ctx.snps_pileup_dev(dreads, lay, Filters(), counts, stats)
        torch.cuda.synchronize()
        ctx.timing_reset()
        for _ in range(steps):
            ctx.snps_pileup_dev(dreads, lay, Filters(), counts, stats)
        torch.cuda.synchronize()
        ms, n = ctx.timing_read()
        got = (counts.cpu().numpy().copy(), stats.cpu().numpy().copy())
        if base is None:
            base = got
        same = bool(np.array_equal(got[0], base[0]) and np.array_equal(got[1], base[1]))
        k_ms = ms / max(1, n)
        print(json.dumps({"setting": setting, "kernel_ms": round(k_ms, 4), "alg_GBps": round(alg / k_ms / 1e6, 1),
                          "same_as_first": same}), flush=True)
        dreads.free()
        ctx.close()


if __name__ == "__main__":
    main()
