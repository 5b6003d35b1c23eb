#!/usr/bin/env python
"""bench.py — aligned-base pileup throughput (Gbases/s) on N B200s, next to the CPU path.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
N > 1 is launched under torchrun, one rank per GPU. `--impl reference` times the CPU port of the
reference path (oracle/) on the same configuration instead.

Headline workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): per GPU one species with a 4 Mbp
reference and 5 M coordinate-sorted synthetic 150-bp reads (seed 20260102 + rank); default snps filters
(mapid 94, mapq 20, baseq 30, readq 20, aln_cov 0.75). N > 1 = weak scaling: reads shard by species, one
species per rank, no data-path collective and no NCCL call inside the timed step: every rank's kernel
epilogue writes its per-species stats row (5 x u64) straight into rank 0's table over NVLink (symmetric
memory), so rank 0 can write summary.txt.

  value  one mb_snps_pileup_dev per step over the batch resident in HBM (packed there by the device packer).
  e2e    mb_snps_pileup_plain per step: decoded records (what the BAM reader hands out: tid, pos, mapq, NM,
         CIGAR, 4-bit seq, qual) in PINNED HOST memory -> H2D -> k_pack_* (the re-layout, on the GPU) -> pileup
         kernels -> D2H of the counts and stats. Nothing of the path runs on the host.

Sub-records (config.c3 / c4 / c5) measure the other BASELINE.json configurations at their stated sizes:
  c4  100 species x 4 Mbp, 200 M reads, lognormal abundance, sharded with midas_b200.shard.plan over the
      ranks (N >= 2; reads drawn on the device), with and without north_star's all-reduce of the count planes
  c3  genes coverage, 100 species x 20 000 genes, 50 M reads (N = 1)
  c5  merge_midas.py snps site arithmetic, 100 samples x 100 species x 4 Mbp, species sharded over the ranks
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
C3 = dict(n_species=100, genes_per_species=20_000, n_reads=50_000_000, read_len=150)
C4 = dict(n_species=100, contig_len=4_000_000, n_reads=200_000_000, read_len=150)
C5 = dict(n_samples=100, n_species=100, n_sites=4_000_000)
FILTERS_TEXT = "mapid94 mapq20 baseq30 readq20 aln_cov0.75"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every workload (debugging only; printed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--only", default="c2,c3,c4,c5", help="comma list of the records to run (c2 is always run)")
    ap.add_argument("--c4-single", action="store_true", help="run the C4 record at N = 1 too (104 GB of HBM)")
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


def python_reference():
    """The reference's own Python loop (species_pileup over the pysam shim), timed in the build container
    (it cannot run on the GPU box: /root/reference is not there); scripts/time_python_reference.py wrote it."""
    p = os.path.join(ROOT, "profiles", "python_reference.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh)
    return None


def algorithmic_bytes(reads, ref) -> int:
    """sum_reads [16 + 4*n_cigar + ceil(l/4) + l] + 16*G."""
    n_cig = (reads.cigar_off[1:] - reads.cigar_off[:-1]).astype(np.int64)
    l = reads.l_qseq.astype(np.int64)
    return int((16 + 4 * n_cig + (l + 3) // 4 + l).sum() + 16 * int(ref.contig_len.sum(dtype=np.int64)))


def plain_nbytes(r) -> int:
    return int(sum(a.nbytes for a in (r.ref_id, r.pos, r.mapq, r.nm, r.l_qseq, r.cigar_off, r.cigar, r.seq_off, r.seq, r.qual)))


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
    """--impl reference: the CPU port of the reference path on this box's host cores, on the SAME workload as
    the b200 arm (the whole C2 batch per step), honouring --steps / --warmup."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from midas_b200.soa import Filters
    cores = os.cpu_count() or 1
    ref, reads = make_workload(0, args.scale)
    filters = Filters()
    n_warm, n_steps = max(0, args.warmup), max(1, args.steps)                # ~0.5 s per pass on 32 threads
    for _ in range(n_warm):
        cpu_port(reads, ref, filters, cores)
    times = [cpu_port(reads, ref, filters, cores) for _ in range(n_steps)]
    t = float(np.mean(times))
    v = reads.total_bases / t / 1e9
    sample = (f"{reads.n_reads} reads x {C2['read_len']} bp over {int(ref.contig_len[0])} bp (the whole C2 batch"
              + ("" if args.scale == 1.0 else f" at scale {args.scale:g}") + f"), {len(times)} timed passes")
    line = {"impl": "reference", "metric": "aligned-base pileup throughput", "value": v, "unit": "Gbases/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": n_warm, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "C2 snps pileup: 1 species x 4 Mbp, 5M x 150bp sorted reads per GPU"
                       + ("" if args.scale == 1.0 else f" (scale {args.scale:g})"),
                       "filters": FILTERS_TEXT},
            "cpu_baseline": {"value": v, "unit": "Gbases/s", "cores": cores, "kind": "port",
                             "sample": sample + "; C port of the pysam_pileup path (oracle/pileup_oracle.c), reads "
                             "sharded over all host threads with private count planes (the reference itself runs one "
                             "process per species, midas/utility.py:81-107)"},
            "e2e": {"value": v, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    pr = python_reference()
    if pr:
        line["cpu_baseline"]["python_reference_value"] = pr["value"]
        line["cpu_baseline"]["python_reference_note"] = pr["note"]
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def timed(fn, steps, warmup, barrier, torch):
    """W warm-up steps, then K steps between CUDA events on the current stream; returns ms for the K steps."""
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def setup_stats_mirror(ctx, world, rank, dev, torch, dist):
    """Rank 0's [world][5] stats table in symmetric memory; every rank's kernels write their row into it."""
    try:
        import torch.distributed._symmetric_memory as symm_mem
        table = symm_mem.empty((world, 5), dtype=torch.int64, device=dev)
        table.zero_()
        hdl = symm_mem.rendezvous(table, dist.group.WORLD)
        ptr0 = int(hdl.buffer_ptrs[0])
        ctx.set_stats_mirror(ptr0 + 40 * rank)
        torch.cuda.synchronize()
        dist.barrier()
        return table, hdl, "kernel epilogue writes the row into rank 0's table over NVLink (symmetric memory)"
    except Exception as e:  # noqa: BLE001 - any failure to set up peer memory: keep the collective
        sys.stderr.write(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); stats rows go through NCCL\n")
        ctx.set_stats_mirror(None)
        return None, None, "nccl all_reduce of the stats rows inside the step (symmetric memory unavailable)"


def bench_c4(args, ctx, world, rank, dev, torch, dist, barrier, peak):
    """BASELINE.json configs[3]: 100 species x 4 Mbp, 200 M reads, lognormal abundance, sharded with shard.plan."""
    from midas_b200 import lib, shard, synth, synth_dev
    from midas_b200.soa import Filters, RefSet
    S, G, L = C4["n_species"], C4["contig_len"], C4["read_len"]
    n_total = max(10_000, int(C4["n_reads"] * args.scale))
    G = max(20_000, int(G * min(1.0, args.scale * 4)))
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([synth.SEED_C4, 0])))
    w = rng.lognormal(0.0, 1.0, S)
    n_s = rng.multinomial(n_total, w / w.sum()).astype(np.int64)
    owner, dom = shard.plan(n_s, world)
    parts, split_mask, mine = [], [], []
    for s in range(S):
        if owner[s] == rank:
            parts.append((s, int(n_s[s]), 0, int(n_s[s])))
            split_mask.append(False)
            mine.append(s)
        elif s in dom:
            b = shard.split_bounds(int(n_s[s]), world)
            parts.append((s, int(n_s[s]), int(b[rank]), int(b[rank + 1])))
            split_mask.append(True)
            mine.append(s)
    filters = Filters()
    prep = {}

    def local_prep():
        t0 = time.perf_counter()
        t = synth_dev.synth_snps_dev(synth.SEED_C4, parts, G, L, device=dev)
        torch.cuda.synchronize()
        prep["gen_s"] = time.perf_counter() - t0
        prep["t"] = t
        return _c4_prep(ctx, t, parts, split_mask, mine, G, L, rank, dev, torch, filters, prep)

    ok, err = True, None
    try:
        local_prep()
    except Exception as e:  # noqa: BLE001
        import traceback
        traceback.print_exc(file=sys.stderr)
        ok, err = False, f"{type(e).__name__}: {e}"
    prep.pop("t", None)
    torch.cuda.empty_cache()
    if not all_ok(ok, world, dev, torch, dist):      # every rank takes the same branch: no collective is left hanging
        return {"error": err or "another rank failed while preparing C4"} if rank == 0 else None
    lay, dreads, counts, stats, check = prep["lay"], prep["dreads"], prep["counts"], prep["stats"], prep["check"]
    alg, pack_ms, gen_s = prep["alg"], prep["pack_ms"], prep["gen_s"]
    return _c4_timed(args, ctx, world, rank, dev, torch, dist, barrier, peak, lay, dreads, counts, stats, check, alg,
                     pack_ms, gen_s, split_mask, mine, dom, owner, S, G, L, filters)


def all_ok(ok, world, dev, torch, dist):
    if world == 1:
        return ok
    f = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(f, op=dist.ReduceOp.MIN)
    return bool(f.item())


def _c4_prep(ctx, t, parts, split_mask, mine, G, L, rank, dev, torch, filters, prep):
    from midas_b200 import lib, synth_dev
    from midas_b200.soa import RefSet
    ref = RefSet([f"sp{s:03d}.c0" for s in mine], np.full(len(mine), G, np.uint32), np.arange(len(mine), dtype=np.int32),
                 [f"sp{s:03d}" for s in mine], None)
    lay = lib.Layout.of(ref)
    n_cig = int(t["cigar_off"][-1].item())
    alg = t["n_reads"] * (16 + (L + 3) // 4 + L) + 4 * n_cig + 16 * G * len(mine)
    dreads = ctx.pack_dev(t, lay, "snps", seq_nibbles=True)
    pack_ms = ctx.last_pack_ms()
    # bit-exact check of one owned species against the CPU oracle (rank 0: its species with the fewest reads)
    check = None
    counts = torch.empty((4, lay.gpad), dtype=torch.int32, device=dev)
    stats = torch.zeros((len(mine), 5), dtype=torch.int64, device=dev)
    ctx.snps_pileup_dev(dreads, lay, filters, counts, stats)
    torch.cuda.synchronize()
    if rank == 0:
        from oracle import oracle
        own = [k for k in range(len(mine)) if not split_mask[k]]
        k = min(own, key=lambda q: parts[q][3] - parts[q][2])
        lo = sum(p[3] - p[2] for p in parts[:k])
        hr = synth_dev.to_host(t, lo, lo + parts[k][3] - parts[k][2])
        hr.ref_id[:] = 0
        one = RefSet([ref.contig_names[k]], np.array([G], np.uint32), np.array([0], np.int32), [ref.species_ids[k]], None)
        exp_c, exp_s = oracle.snps_pileup(hr, one, filters, threads=os.cpu_count() or 1)
        o = int(lay.contig_goff[k])
        got = counts[:, o:o + G].cpu().numpy().view(np.uint32)
        ok = bool(np.array_equal(got, exp_c.astype(np.uint32))) and \
            bool(np.array_equal(stats[k].cpu().numpy().astype(np.uint64), exp_s[0]))
        check = {"bit_exact": ok, "species": mine[k], "reads": int(hr.n_reads),
                 "against": "oracle/pileup_oracle.c on the same records copied back from the device"}
    prep.update(lay=lay, dreads=dreads, counts=counts, stats=stats, check=check, alg=alg, pack_ms=pack_ms)


def _c4_timed(args, ctx, world, rank, dev, torch, dist, barrier, peak, lay, dreads, counts, stats, check, alg, pack_ms,
              gen_s, split_mask, mine, dom, owner, S, G, L, filters):
    from midas_b200 import shard
    any_split = bool(len(dom))

    def step():
        ctx.snps_pileup_dev(dreads, lay, filters, counts, stats)
        if any_split:
            shard.allreduce_split_counts(counts, stats, lay, np.array(split_mask))

    steps = max(1, min(args.steps, 20))
    ctx.timing_reset()
    ms = timed(step, steps, max(args.warmup, 3), barrier, torch)
    kms, kn = ctx.timing_read()
    mine_ms = ms / steps
    # north_star's literal final step: all-reduce of the per-position count arrays of the whole sample
    allred_ms = None
    if world > 1:
        full = torch.zeros((4, S * G), dtype=torch.int32, device=dev)
        for k, s in enumerate(mine):
            if not split_mask[k] or rank == 0:
                o = int(lay.contig_goff[k])
                full[:, s * G:(s + 1) * G] = counts[:, o:o + G]
        dist.all_reduce(full)                       # warm-up (communicator, buffers)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        dist.all_reduce(full)
        a1.record()
        barrier()
        allred_ms = a0.elapsed_time(a1)
        del full
    v = torch.tensor([mine_ms, float(dreads.total_bases), float(alg), kms / max(1, kn), float(dreads.n_reads),
                      allred_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        allv = torch.stack(allv).cpu().numpy()
    else:
        allv = v.cpu().numpy()[None, :]
    del dreads, counts
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    ms_max, ms_mean = float(allv[:, 0].max()), float(allv[:, 0].mean())
    bases = float(allv[:, 1].sum())
    kern = allv[:, 3]
    frac = (allv[:, 2] / (kern * 1e-3) / 1e9 / peak)
    ar = float(allv[:, 5].max())
    return {"workload": f"C4 snps pileup: {S} species x {G} bp, {int(allv[:, 4].sum())} x {L}bp reads, lognormal abundance"
                        + ("" if args.scale == 1.0 else f" (scale {args.scale:g})"),
            "sharding": f"shard.plan over {world} ranks: {int((owner >= 0).sum())} species owned whole (LPT on read "
                        f"count), {len(dom)} split by read index" + (" + NCCL all-reduce of their planes inside the step" if any_split else ""),
            "value": bases / (ms_max * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": ms_max,
            "rank_ms": {"max": ms_max, "mean": ms_mean, "imbalance": ms_max / ms_mean},
            "reads_per_rank": {"max": int(allv[:, 4].max()), "min": int(allv[:, 4].min())},
            "roofline_frac_per_rank": {"min": float(frac.min()), "max": float(frac.max())},
            "allreduce_counts_ms": ar if world > 1 else None,
            "value_with_allreduce_counts": bases / ((ms_max + ar) * 1e-3) / 1e9 if world > 1 else None,
            "allreduce_counts_note": "north_star's final NCCL all-reduce of the [4][100 x 4 Mbp] u32 planes (6.4 GB "
                                     "per rank, disjoint supports); not needed by the path (each rank writes its own "
                                     "species), reported beside it",
            "device_pack_ms": pack_ms, "synth_on_device_s": round(gen_s, 2), "steps": steps, "check": check}


def bench_c3(args, ctx, dev, torch, peak):
    """BASELINE.json configs[2]: genes coverage, 100 species x 20 000 genes, 50 M x 150 bp reads, one GPU."""
    from midas_b200 import lib, synth, synth_dev
    from midas_b200.soa import Filters
    from oracle import oracle
    L = C3["read_len"]
    n_reads = max(10_000, int(C3["n_reads"] * args.scale))
    gps = max(50, int(C3["genes_per_species"] * min(1.0, args.scale * 10)))
    glen, gw, _, _ = synth.genes_model(synth.SEED_C3, C3["n_species"], gps, L)
    n_genes = len(glen)
    filters = Filters.genes_defaults()
    t0 = time.perf_counter()
    t = synth_dev.synth_genes_dev(synth.SEED_C3, glen, gw, n_reads, L, device=dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    n_cig = int(t["cigar_off"][-1].item())
    alg = n_reads * (16 + L) + 4 * n_cig + 24 * n_genes
    lay = lib.Layout.genes(n_genes)
    dreads = ctx.pack_dev(t, lay, "genes", seq_nibbles=True)
    pack_ms = ctx.last_pack_ms()
    gl = torch.as_tensor(glen.astype(np.int32), device=dev)
    out = torch.zeros((n_genes, 3), dtype=torch.int64, device=dev)
    # check + end-to-end on the first 2 M records: decoded records on the host -> mb_genes_coverage_plain vs the oracle
    n_chk = min(n_reads, 2_000_000)
    hr = lib.pinned_reads(synth_dev.to_host(t, 0, n_chk))
    del t
    torch.cuda.empty_cache()
    glen32 = glen.astype(np.uint32)
    got = ctx.genes_coverage_plain(hr, glen32, filters)
    t1 = time.perf_counter()
    for _ in range(3):
        got = ctx.genes_coverage_plain(hr, glen32, filters)
    e2e_s = (time.perf_counter() - t1) / 3
    exp = oracle.genes_coverage(hr, glen32, filters)
    ok = bool(np.array_equal(got["aligned_reads"], exp["aligned"]) and np.array_equal(got["mapped_reads"], exp["mapped"])
              and np.array_equal(got["mapped_bp"], exp["mapped_bp"]) and np.allclose(got["depth"], exp["depth"], rtol=1e-12, atol=0))
    steps = max(1, min(args.steps, 20))

    def step():
        ctx.genes_coverage_dev(dreads, gl, n_genes, filters, out)

    ctx.timing_reset()
    ms = timed(step, steps, max(args.warmup, 3), torch.cuda.synchronize, torch)
    kms, kn = ctx.timing_read()
    k = kms / max(1, kn)
    bases = float(dreads.total_bases)
    rec = {"workload": f"C3 genes coverage: {C3['n_species']} species x {gps} genes, {n_reads} x {L}bp reads in file order"
                       + ("" if args.scale == 1.0 else f" (scale {args.scale:g})"),
           "value": bases / (ms / steps * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": ms / steps,
           "roofline": {"bound": "hbm", "kernel": "k_genes_v2", "achieved": alg / (k * 1e-3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": alg / (k * 1e-3) / 1e9 / peak, "kernel_ms": k,
                        "algorithmic_bytes_per_launch": alg},
           "e2e": {"value": n_chk * L / e2e_s / 1e9, "unit": "Gbases/s", "sample": f"first {n_chk} records, pinned host "
                   "buffers -> mb_genes_coverage_plain (H2D + device re-layout + kernel + D2H)",
                   "h2d_bytes_per_step": plain_nbytes(hr) + glen32.nbytes, "d2h_bytes_per_step": int(got.nbytes)},
           "device_pack_ms": pack_ms, "synth_on_device_s": round(gen_s, 2), "steps": steps,
           "check": {"bit_exact_counters": ok, "reads": n_chk, "against": "oracle/pileup_oracle.c (depth within 1e-12)"}}
    del dreads
    torch.cuda.empty_cache()
    return rec


def bench_c5(args, ctx, world, rank, dev, torch, dist, barrier, peak):
    """BASELINE.json configs[4]: merge_midas.py snps site arithmetic, 100 samples x 100 species, species sharded."""
    from oracle import merge_oracle as mo
    S, NS = C5["n_samples"], C5["n_species"]
    G = max(50_000, int(C5["n_sites"] * args.scale))
    g = torch.Generator(device=dev).manual_seed(20260105 + rank)
    pool = []
    for _ in range(2):                       # two distinct tensors, alternated: no species sees a warm L2
        major = torch.randint(0, 4, (G,), device=dev, generator=g)
        depth = torch.poisson(torch.full((S, G), 20.0, device=dev), generator=g).to(torch.int32)
        minor_n = (torch.rand((S, G), device=dev, generator=g) < 0.01).to(torch.int32) * (depth // 3)
        counts = torch.zeros((S, 4, G), dtype=torch.int32, device=dev)
        counts.scatter_(1, major.view(1, 1, G).expand(S, 1, G), (depth - minor_n).unsqueeze(1))
        counts.scatter_add_(1, ((major + 1) % 4).view(1, 1, G).expand(S, 1, G), minor_n.unsqueeze(1))
        md = counts.sum(1).double().mean(1).clamp_min(0.5)
        pool.append((counts, md))
        del depth, minor_n, major
    mine = [s for s in range(NS) if s % world == rank]

    def step():
        for s in mine:
            c, md = pool[s & 1]
            ctx.merge_sites(c, md)

    out = ctx.merge_sites(*pool[0])
    torch.cuda.synchronize()
    n_chk = min(G, 4000)
    exp = mo.merge_sites(pool[0][0][:, :, :n_chk].cpu().numpy().view(np.uint32), pool[0][1].cpu().numpy())
    call = out["call"][:n_chk].cpu().numpy().astype(np.int16)
    ok = bool(np.array_equal(out["pooled"][:, :n_chk].cpu().numpy().view(np.uint32).astype(np.uint64), exp["pooled"])
              and np.array_equal(out["count_samples"][:n_chk].cpu().numpy(), exp["count_samples"])
              and np.array_equal((call >> 4) & 7, exp["snp_type"])
              and np.array_equal(out["sample_maf"][:, :n_chk].cpu().numpy().view(np.uint64), exp["sample_maf"].view(np.uint64)))
    del out
    steps = max(1, min(args.steps, 3))
    ctx.timing_reset()
    ms = timed(step, steps, 1, barrier, torch)
    kms, kn = ctx.timing_read()
    k = kms / max(1, kn)
    v = torch.tensor([ms / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    del pool
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    alg = G * (16 * S + 12 * S + 25)
    return {"workload": f"C5 merge_midas.py snps: {S} samples x {NS} species x {G} sites"
                        + ("" if args.scale == 1.0 else f" (scale {args.scale:g})"),
            "sharding": f"species s -> rank s % {world}, no collective",
            "value": NS * G / (float(v.item()) * 1e-3) / 1e9, "unit": "Gsites/s", "ms_per_step": float(v.item()),
            "roofline": {"bound": "hbm", "kernel": "k_merge_sites", "achieved": alg / (k * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (k * 1e-3) / 1e9 / peak, "kernel_ms": k,
                         "algorithmic_bytes_per_launch": alg},
            "steps": steps, "check": {"bit_exact": ok, "sites": n_chk, "against": "oracle/merge_oracle.py"}}


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
    only = set(args.only.split(","))
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
    peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- C2: the headline line
    ref, reads = make_workload(rank, args.scale)
    lay = lib.Layout.of(ref)
    alg_bytes = algorithmic_bytes(reads, ref)
    # what the BAM reader hands out: decoded records, seq as in the file, in page-locked memory
    plain = lib.pinned_reads(reads.nibble_packed())
    del reads
    ctx = lib.Context(dev.index)           # launches on torch's current stream
    dreads = ctx.pack(plain, lay, "snps")  # resident batch, packed on the GPU
    pack_ms = ctx.last_pack_ms()
    total_bases = dreads.total_bases
    counts = torch.empty((4, lay.gpad), dtype=torch.int32, device=dev)
    my_stats = torch.zeros((1, 5), dtype=torch.int64, device=dev)   # this rank's species
    table, hdl, stats_mode = (None, None, "single GPU")
    if world > 1:
        table, hdl, stats_mode = setup_stats_mirror(ctx, world, rank, dev, torch, dist)
    gathered = torch.zeros((world, 5), dtype=torch.int64, device=dev) if (world > 1 and table is None) else None

    def step():
        ctx.snps_pileup_dev(dreads, lay, filters, counts, my_stats)
        if gathered is not None:
            gathered.zero_()
            gathered[rank] = my_stats[0]
            dist.all_reduce(gathered)

    sampler = ClockSampler(dev.index)
    sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
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
    launches = ctx.launch_count - launches0   # this library's kernels only
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tb = torch.tensor([float(total_bases)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    ms_max, bases_all = float(t.item()), float(tb.item())
    ms_per_step = ms_max / args.steps
    value = bases_all / (ms_per_step * 1e-3) / 1e9
    rows_ok = None
    if world > 1 and rank == 0:
        rows = (table if table is not None else gathered).cpu().numpy()
        rows_ok = bool((rows[:, 4] > 0).all() and (rows[:, 0] == int(ref.contig_len[0])).all())
    ctx.set_stats_mirror(None)                # the table has one row per rank: only the headline step writes it

    # ---- end to end through the host-buffer C-ABI call: decoded records in pinned host memory in,
    # counts + stats in pinned host memory out; H2D, the re-layout (k_pack_*), the pileup and D2H are all inside
    counts_host = torch.empty((4, lay.gpad), dtype=torch.int32).pin_memory()
    counts_np = counts_host.numpy().view(np.uint32)
    ctx.snps_pileup_plain(plain, lay, filters, counts_np)          # warm-up: grows the staging buffers
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        _, st_host = ctx.snps_pileup_plain(plain, lay, filters, counts_np)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    sampler.active = False
    e2e_pack_ms = ctx.last_pack_ms()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = bases_all / float(te.item()) / 1e9
    h2d_rank = plain_nbytes(plain)
    h2d = h2d_rank * world                    # whole job: every rank copies its own species in and out
    d2h = (counts_np.nbytes + st_host.nbytes) * world

    # sanity: the device-resident and host-buffer paths agree, and something was counted
    dev_counts = counts.cpu().numpy().view(np.uint32)
    assert np.array_equal(dev_counts, counts_np), "device-resident and host-buffer results differ"
    assert int(st_host["mapped_reads"][0]) > 0
    assert np.array_equal(my_stats.cpu().numpy().astype(np.uint64).ravel(),
                          np.array([st_host[n][0] for n in lib.STATS_DTYPE.names], np.uint64))

    line = None
    if rank == 0:
        kern_avg_ms = kern_ms / max(1, kern_n)
        achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic("k_pileup_v2")
        line = {
            "metric": "aligned-base pileup throughput", "value": value, "unit": "Gbases/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "C2 snps pileup: 1 species x 4 Mbp, 5M x 150bp sorted reads per GPU"
                       + ("" if args.scale == 1.0 else f" (scale {args.scale:g})"),
                       "filters": FILTERS_TEXT,
                       "sharding": f"species x{world}, no collective in the step; stats rows: {stats_mode}"
                       if world > 1 else "single GPU",
                       "stats_rows_complete_on_rank0": rows_ok,
                       "l2": "inputs (1.1 GB resident SoA per GPU) larger than the 126 MB L2",
                       "bases_per_step": bases_all,
                       "device_pack_ms": pack_ms,
                       "note": "value: batch resident in HBM, one mb_snps_pileup_dev per step. e2e: "
                               "mb_snps_pileup_plain per step - decoded records (tid, pos, mapq, NM, CIGAR, 4-bit "
                               "seq, qual: what the BAM reader hands out) in pinned host memory -> H2D -> k_pack_* "
                               "(re-layout on the GPU) -> pileup -> D2H of counts + stats; no host-side packing "
                               "anywhere on the path"},
            "e2e": {"value": e2e_value, "unit": "Gbases/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te.item()) * 1e3, "device_pack_ms": e2e_pack_ms,
                    "h2d_gbs_aggregate": h2d / float(te.item()) / 1e9,
                    "note": "h2d_gbs_aggregate = all ranks' input bytes / e2e step time: the host-side copy ceiling "
                            "the e2e curve runs into at N = 4, 8 (one NUMA node feeds every GPU)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_pileup_v2", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_avg_ms,
                         "kernel_share_of_step": kern_avg_ms / ms_per_step},
        }
    del dreads, counts, counts_host
    torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N = 1): the C port on the whole C2 batch, all host threads
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle  # noqa: F401  (the checker, timed as the CPU baseline)
        cref, creads = make_workload(0, args.scale)
        cores = os.cpu_count() or 1
        cpu_port(creads, cref, filters, cores)                                   # warm-up: page faults, thread start
        t_all = float(np.mean([cpu_port(creads, cref, filters, cores) for _ in range(5)]))
        sub = creads.select(slice(0, max(1000, creads.n_reads // 8)))
        t_one = cpu_port(sub, cref, filters, 1)
        line["cpu_baseline"] = {
            "value": creads.total_bases / t_all / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "port",
            "sample": f"{creads.n_reads} reads x {C2['read_len']} bp (the whole C2 batch), mean of 5 passes after one "
                      "warm-up; C port of the pysam_pileup path, reads sharded over all host threads",
            "single_thread_value": sub.total_bases / t_one / 1e9}
        pr = python_reference()
        if pr:
            line["cpu_baseline"]["python_reference_value"] = pr["value"]
            line["cpu_baseline"]["python_reference_note"] = pr["note"]
        del creads, sub
    del plain

    # ---- the other BASELINE.json configurations, as sub-records
    def guarded(name, fn):
        try:
            return fn()
        except Exception as e:  # noqa: BLE001 - a sub-record must never take the headline line down
            import traceback
            traceback.print_exc(file=sys.stderr)
            return {"error": f"{type(e).__name__}: {e}"} if rank == 0 else None

    if "c4" in only and (world > 1 or args.c4_single):
        rec = guarded("c4", lambda: bench_c4(args, ctx, world, rank, dev, torch, dist, barrier, peak))
        if rank == 0:
            line["config"]["c4"] = rec
    if "c3" in only and world == 1:
        rec = guarded("c3", lambda: bench_c3(args, ctx, dev, torch, peak))
        line["config"]["c3"] = rec
    if "c5" in only:
        rec = guarded("c5", lambda: bench_c5(args, ctx, world, rank, dev, torch, dist, barrier, peak))
        if rank == 0:
            line["config"]["c5"] = rec
    sampler.stop_flag = True
    if rank == 0:
        line["clocks"] = sampler.summary()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
