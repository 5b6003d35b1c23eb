#!/usr/bin/env python
"""Phase clocks of the pileup tile kernel on the C2 workload: cycles per warp in STAGE / COUNT / EXPAND and at the
barriers behind them, and when each CTA finished (the tail of the persistent grid). Needs the phase-clock build:

  make -C midas_b200/csrc prof
  MB_LIB_PATH=midas_b200/libmidas_b200_prof.so python scripts/phase_prof.py "MB_PILE_VAR=8" "MB_PILE_VAR=72"

The clocks perturb the kernel a little (one shared-memory update per warp and phase): the split is what is read
here, kernel times come from scripts/sweep_pileup.py on the normal build."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from midas_b200 import lib  # noqa: E402
from midas_b200.soa import Filters  # noqa: E402

NAMES = ("other", "stage", "wait_after_stage", "count", "wait_after_count", "expand")


def main():
    settings = sys.argv[1:] or ["MB_PILE_VAR=8"]
    ref, reads = bench.make_workload(0, float(os.environ.get("SWEEP_SCALE", "1.0")))
    lay = lib.Layout.of(ref)
    packed = lib.PackedReads(reads, lay, "snps", sort=False, pinned=False)
    L = lib.load()
    L.mb_debug_prof.argtypes = [C.c_void_p, C.c_size_t]
    dev = torch.device("cuda", 0)
    for setting in settings:
        for k in ("MB_PILE_VAR", "MB_PILE_SMEM_KB", "MB_PILE_TMA", "MB_PILE_UNIT"):
            os.environ.pop(k, None)
        for kv in setting.split():
            k, v = kv.split("=")
            os.environ[k] = v
        ctx = lib.Context(0)
        dreads = ctx.upload(packed)
        counts = torch.empty((4, lay.gpad), dtype=torch.int32, device=dev)
        stats = torch.zeros((1, 5), dtype=torch.int64, device=dev)
        for _ in range(3):
            ctx.snps_pileup_dev(dreads, lay, Filters(), counts, stats)
        torch.cuda.synchronize()
        ctx.timing_reset()
        for _ in range(5):
            ctx.snps_pileup_dev(dreads, lay, Filters(), counts, stats)
        torch.cuda.synchronize()
        ms, n = ctx.timing_read()
        n_cta = int(os.environ.get("PROF_CTAS", "296"))
        buf = np.zeros(n_cta * 8 * 8, dtype=np.uint64)
        rc = L.mb_debug_prof(buf.ctypes.data, buf.size)
        assert rc == 0, rc
        a = buf.reshape(n_cta, 8, 8).astype(np.float64)
        cyc = a[:, :, :6]
        tot = cyc.sum(axis=2)
        share = cyc.sum(axis=(0, 1)) / tot.sum()
        t0 = a[:, :, 6].min()
        end = (a[:, 0, 7] - t0) / 1e3                       # us, per CTA
        out = {"setting": setting, "kernel_ms_with_clocks": round(ms / max(1, n), 4),
               "share": {k: round(float(v), 4) for k, v in zip(NAMES, share)},
               "cycles_per_warp": {k: int(v) for k, v in zip(NAMES, cyc.mean(axis=(0, 1)))},
               "per_warp_stage_share": [round(float(x), 3) for x in (cyc[:, :, 1].sum(axis=0) / tot.sum(axis=0))],
               "per_warp_count_share": [round(float(x), 3) for x in (cyc[:, :, 3].sum(axis=0) / tot.sum(axis=0))],
               "per_warp_wait_after_stage": [round(float(x), 3) for x in (cyc[:, :, 2].sum(axis=0) / tot.sum(axis=0))],
               "cta_end_us": {"min": round(float(end.min()), 1), "p10": round(float(np.percentile(end, 10)), 1),
                              "median": round(float(np.median(end)), 1), "p90": round(float(np.percentile(end, 90)), 1),
                              "max": round(float(end.max()), 1)},
               "idle_tail_share": round(float((end.max() - end).mean() / end.max()), 4)}
        print(json.dumps(out), flush=True)
        dreads.free()
        ctx.close()


if __name__ == "__main__":
    main()
