#!/usr/bin/env python
"""bench.py — aligned-base pileup throughput (Gbases/s) on N B200s, next to the CPU path.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
N > 1 is launched under torchrun, one rank per GPU. `--impl reference` times the CPU port of the
reference path (oracle/) instead.

Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): per GPU one species with a 4 Mbp
reference and 5 M coordinate-sorted synthetic 150-bp reads (seed 20260102 + rank); default snps
filters (mapid 94, mapq 20, baseq 30, readq 20, aln_cov 0.75). N > 1 = weak scaling: reads shard
by species, one species per rank, no data-path collective; the per-species stats rows
(5 x u64 per species) are all-reduced inside the timed step so rank 0 can write summary.txt.

A step = one pass of the hot path over the resident batch: mb_snps_pileup_dev (stats init,
tile index, pileup kernel). `value` = total bases of all ranks / max-over-ranks device time.
`e2e` = the same through mb_snps_pileup with pinned HOST buffers (H2D + kernels + D2H per step).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C2 = dict(contig_len=4_000_000, n_reads=5_000_000, read_len=150)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only; prints it)")
    ap.add_argument("--reduce", default="stats", choices=["stats", "counts"],
                    help="N>1: what is all-reduced inside the timed step (counts = north_star's full allreduce)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel: str):
    """dram__bytes_read+write per launch of `kernel` from the committed ncu capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as fh:
            t = json.load(fh).get(kernel)
        if t:
            return t["dram_bytes_per_launch"], t["source"]
    return None, None


def algorithmic_bytes(reads, ref) -> int:
    """sum_reads [16 + 4*n_cigar + ceil(l/4) + l] + 16*G."""
    n_cig = (reads.cigar_off[1:] - reads.cigar_off[:-1]).astype(np.int64)
    l = reads.l_qseq.astype(np.int64)
    return int((16 + 4 * n_cig + (l + 3) // 4 + l).sum() + 16 * int(ref.contig_len.sum(dtype=np.int64)))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.active = index, [], set(), False, False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                 "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                        nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.005)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_workload(rank: int, scale: float):
    from midas_b200 import synth
    n_reads = max(1000, int(C2["n_reads"] * scale))
    clen = max(10_000, int(C2["contig_len"] * scale))
    ref, reads = synth.synth_snps(seed=synth.SEED_C2 + rank, n_species=1, contigs_per_species=1, contig_len=clen,
                                  n_reads=n_reads, read_len=C2["read_len"], keep_ref_seq=False)
    return ref, reads


def cpu_port(reads, ref, filters, threads):
    from oracle import oracle
    oracle.lib()
    t0 = time.perf_counter()
    oracle.snps_pileup(reads, ref, filters, threads=threads)
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the CPU port of the reference path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from midas_b200.soa import Filters
    cores = os.cpu_count() or 1
    # bounded sample: the first 1/4 of the C2 reads (same generator, same seed) per step
    frac = 0.25 * args.scale
    ref, reads = make_workload(0, frac)
    filters = Filters()
    n_warm, n_steps = min(args.warmup, 3), max(1, min(args.steps, 50))       # ~0.13 s per pass on 16 threads
    for _ in range(n_warm):
        cpu_port(reads, ref, filters, cores)
    times = [cpu_port(reads, ref, filters, cores) for _ in range(n_steps)]
    t = float(np.mean(times))
    v = reads.total_bases / t / 1e9
    sample = (f"{reads.n_reads} reads x {C2['read_len']} bp over {int(ref.contig_len[0])} bp (C2 generator at "
              f"{frac:g} scale), {len(times)} timed passes")
    line = {"impl": "reference", "metric": "aligned-base pileup throughput", "value": v, "unit": "Gbases/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": n_warm, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "C2 snps pileup: 1 species x 4 Mbp, 5M x 150bp sorted reads (bounded sample)",
                       "filters": "mapid94 mapq20 baseq30 readq20 aln_cov0.75"},
            "cpu_baseline": {"value": v, "unit": "Gbases/s", "cores": cores, "kind": "port",
                             "sample": sample + "; C port of the pysam_pileup path (oracle/pileup_oracle.c), reads "
                             "sharded over all host threads with private count planes (the reference itself runs one "
                             "process per species, midas/utility.py:81-107)"},
            "e2e": {"value": v, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from midas_b200 import lib
    from midas_b200.soa import Filters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        # NCCL writes its version banner to stdout at the first collective: send it to stderr so that
        # stdout carries the one JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local if world > 1 else 0)
    filters = Filters()

    ref, reads = make_workload(rank, args.scale)
    lay = lib.Layout.of(ref)
    t_pack = time.perf_counter()
    packed = lib.PackedReads(reads, lay, "snps", sort=False, pinned=True)
    t_pack = time.perf_counter() - t_pack
    alg_bytes = algorithmic_bytes(reads, ref)
    total_bases = packed.total_bases

    ctx = lib.Context(dev.index)           # launches on torch's current stream
    dreads = ctx.upload(packed)
    counts = torch.empty((4, lay.gpad), dtype=torch.int32, device=dev)
    stats = torch.zeros((world, 5), dtype=torch.int64, device=dev)   # row per species (= per rank)
    my_stats = stats[rank:rank + 1]

    def step():
        if world > 1:
            stats.zero_()
        ctx.snps_pileup_dev(dreads, lay, filters, counts, my_stats)
        if world > 1:
            dist.all_reduce(counts if args.reduce == "counts" else stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(dev.index)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ctx.timing_reset()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active = True
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.active = False
    ms = e0.elapsed_time(e1)
    kern_ms, kern_n = ctx.timing_read()
    launches = ctx.launch_count - launches0   # this library's kernels only (NCCL / torch fills not counted)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tb = torch.tensor([float(total_bases)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    ms_max, bases_all = float(t.item()), float(tb.item())
    ms_per_step = ms_max / args.steps
    value = bases_all / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host-buffer C-ABI call (pinned host in, pinned host out)
    counts_host = torch.empty((4, lay.gpad), dtype=torch.int32).pin_memory()
    counts_np = counts_host.numpy().view(np.uint32)
    ctx.snps_pileup(packed, lay, filters, counts_np)          # warm-up: grows the staging buffers
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        _, st_host = ctx.snps_pileup(packed, lay, filters, counts_np)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    sampler.active = False
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = bases_all / float(te.item()) / 1e9
    h2d = packed.nbytes * world               # whole job: every rank copies its own species in and out
    d2h = (counts_np.nbytes + st_host.nbytes) * world
    sampler.stop_flag = True

    # sanity: the device-resident and host-buffer paths agree, and something was counted
    dev_counts = counts.cpu().numpy().view(np.uint32)
    assert np.array_equal(dev_counts, counts_np), "device-resident and host-buffer results differ"
    assert int(st_host["mapped_reads"][0]) > 0

    if rank == 0:
        peak, peak_src = peaks()
        kern_avg_ms = kern_ms / max(1, kern_n)
        achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic("k_pileup_v2")
        line = {
            "metric": "aligned-base pileup throughput", "value": value, "unit": "Gbases/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "C2 snps pileup: 1 species x 4 Mbp, 5M x 150bp sorted reads per GPU"
                       + ("" if args.scale == 1.0 else f" (scale {args.scale:g})"),
                       "filters": "mapid94 mapq20 baseq30 readq20 aln_cov0.75",
                       "sharding": f"species x{world}, allreduce={args.reduce}" if world > 1 else "single GPU",
                       "l2": "inputs (1.1 GB resident SoA per GPU) larger than the 126 MB L2",
                       "bases_per_step": bases_all,
                       "host_pack_ms": round(t_pack * 1e3, 1),
                       "note": "value: reads resident in HBM, one mb_snps_pileup_dev per step; e2e: mb_snps_pileup "
                               "with pinned host buffers (packed SoA in, counts + stats out), H2D + kernels + D2H per "
                               "step; host_pack_ms: one-off mb_pack_reads of the decoded records (not in either)"},
            "e2e": {"value": e2e_value, "unit": "Gbases/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te.item()) * 1e3},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_pileup_v2", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_avg_ms,
                         "kernel_share_of_step": kern_avg_ms / ms_per_step},
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle  # noqa: F401  (the checker, timed as the CPU baseline)
            frac = 0.25 * args.scale
            cref, creads = make_workload(0, frac)
            cores = os.cpu_count() or 1
            cpu_port(creads, cref, filters, cores)                                   # warm-up: page faults, thread start
            t_all = float(np.mean([cpu_port(creads, cref, filters, cores) for _ in range(5)]))
            t_one = cpu_port(creads, cref, filters, 1)
            line["cpu_baseline"] = {
                "value": creads.total_bases / t_all / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "port",
                "sample": f"{creads.n_reads} reads x {C2['read_len']} bp (C2 generator at {frac:g} scale), mean of 5 "
                          "passes after one warm-up; C port of the pysam_pileup path, reads sharded over all host threads",
                "single_thread_value": creads.total_bases / t_one / 1e9}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
