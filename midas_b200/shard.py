"""Multi-GPU sharding of the pileup path: reads shard by species, one process per GPU.

The reference's only parallelism is one forked worker per species (midas/run/snps.py:225-228 via
midas/utility.py:81-107); species are independent (a read belongs to one contig, a contig to one
species, snps.py:64,191), so ranks own disjoint species and the data path needs no collective.
What is exchanged is small: the per-species stats rows (5 x u64) so that rank 0 can write
summary.txt, and - only if a consumer wants every rank to hold the whole sample - the count planes
(disjoint supports, so an all-reduce(sum) of zero-initialised planes is an all-gather).
"""
from __future__ import annotations

import numpy as np

from .soa import PlainReads, RefSet


def assign_species(read_counts, n_ranks: int) -> np.ndarray:
    """Greedy longest-processing-time bin packing of species onto ranks by read count.
    Returns int32 [n_species] rank of each species (deterministic: ties -> lower species index first)."""
    read_counts = np.asarray(read_counts, np.int64)
    order = np.lexsort((np.arange(len(read_counts)), -read_counts))
    load = np.zeros(n_ranks, np.int64)
    owner = np.zeros(len(read_counts), np.int32)
    for s in order:
        r = int(np.argmin(load))          # first minimum -> deterministic
        owner[s] = r
        load[r] += read_counts[s]
    return owner


def species_read_counts(reads: PlainReads, ref: RefSet) -> np.ndarray:
    sp = ref.contig_species[reads.ref_id] if reads.n_reads else np.zeros(0, np.int32)
    return np.bincount(sp, minlength=ref.n_species).astype(np.int64)


def shard(reads: PlainReads, ref: RefSet, rank: int, n_ranks: int):
    """The part of (ref, reads) that `rank` owns. Returns (sub RefSet, sub PlainReads, species_index)
    where species_index[k] is the global index of the rank's k-th species."""
    owner = assign_species(species_read_counts(reads, ref), n_ranks)
    mine = np.nonzero(owner == rank)[0].astype(np.int32)
    sp_new = -np.ones(ref.n_species, np.int32)
    sp_new[mine] = np.arange(len(mine), dtype=np.int32)
    keep_contig = np.nonzero(sp_new[ref.contig_species] >= 0)[0]
    c_new = -np.ones(ref.n_contigs, np.int32)
    c_new[keep_contig] = np.arange(len(keep_contig), dtype=np.int32)
    sub_ref = RefSet([ref.contig_names[c] for c in keep_contig], ref.contig_len[keep_contig],
                     sp_new[ref.contig_species[keep_contig]], [ref.species_ids[s] for s in mine],
                     [ref.contig_seq[c] for c in keep_contig] if ref.contig_seq is not None else None)
    if reads.n_reads:
        sel = np.nonzero(c_new[reads.ref_id] >= 0)[0]
        sub_reads = reads.select(sel)
        sub_reads.ref_id = np.ascontiguousarray(c_new[sub_reads.ref_id], np.int32)
    else:
        sub_reads = reads
    return sub_ref, sub_reads, mine


def allreduce_stats(local_stats: np.ndarray, species_index: np.ndarray, n_species: int, device=None,
                    split_mask=None):
    """Sum-combine per-species stats rows across ranks with torch.distributed (NCCL on GPUs, gloo on
    CPU). local_stats: int64/uint64 [n_local][5]; returns int64 [n_species][5] on every rank.
    Rows of split species (split_mask, already complete on every rank after
    allreduce_split_counts) are contributed by rank 0 only."""
    import torch
    import torch.distributed as dist
    local_stats = np.asarray(local_stats).astype(np.int64).reshape(-1, 5)
    if split_mask is not None and dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
        local_stats = np.where(np.asarray(split_mask, bool)[:, None], 0, local_stats)
    full = torch.zeros((n_species, 5), dtype=torch.int64, device=device)
    if len(species_index):
        full[torch.as_tensor(np.asarray(species_index, np.int64), device=device)] = torch.as_tensor(
            np.asarray(local_stats).astype(np.int64).reshape(-1, 5), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(full)
    return full.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# A species holding more than 1/n_ranks of the reads cannot be balanced by ownership alone
# (SURVEY 8e "fallback"): its reads are split by index over all ranks, every rank counts its part
# into a full replica of that species' planes, and the replicas are summed with one all-reduce
# (integer sums: bit-exact in any reduction order). aligned_reads / mapped_reads add up across
# ranks; covered_bases / total_depth are recomputed from the reduced planes.

def dominant_species(read_counts, n_ranks: int):
    """Indices of species with more than 1/n_ranks of all reads (n_ranks > 1)."""
    read_counts = np.asarray(read_counts, np.int64)
    total = int(read_counts.sum())
    if n_ranks <= 1 or total == 0:
        return np.zeros(0, np.int32)
    return np.nonzero(read_counts * n_ranks > total)[0].astype(np.int32)


def plan(read_counts, n_ranks: int):
    """The sharding decision for a sample, from the per-species read counts alone (every rank computes the
    same plan): dominant species (> 1/n_ranks of the reads) are split by read index over all ranks, the
    others are owned by one rank each (LPT bin packing on what is left).
    Returns (owner int32 [n_species], -1 for split species; dom int32 [n_dom] indices of the split species)."""
    read_counts = np.asarray(read_counts, np.int64)
    dom = dominant_species(read_counts, n_ranks)
    rest = read_counts.copy()
    rest[dom] = 0
    owner = assign_species(rest, n_ranks)
    owner[dom] = -1
    return owner, dom


def split_bounds(n: int, n_ranks: int) -> np.ndarray:
    """Index ranges [b[r], b[r+1]) of a split species' reads (contiguous: they keep the sort order)."""
    return np.linspace(0, int(n), n_ranks + 1).astype(np.int64)


def shard_split(reads: PlainReads, ref: RefSet, rank: int, n_ranks: int):
    """Like shard(), but dominant species are split by read index over all ranks.
    Returns (sub RefSet, sub PlainReads, species_index, split_mask) where split_mask[k] is True for the
    rank's k-th species if it is a split (replicated) one that needs the count all-reduce."""
    counts = species_read_counts(reads, ref)
    owner, dom = plan(counts, n_ranks)
    sp_of_read = ref.contig_species[reads.ref_id] if reads.n_reads else np.zeros(0, np.int32)
    is_dom_sp = np.zeros(ref.n_species, bool)
    is_dom_sp[dom] = True
    mine_sp = np.nonzero(is_dom_sp | (owner == rank))[0].astype(np.int32)
    sp_new = -np.ones(ref.n_species, np.int32)
    sp_new[mine_sp] = np.arange(len(mine_sp), dtype=np.int32)
    keep_contig = np.nonzero(sp_new[ref.contig_species] >= 0)[0]
    c_new = -np.ones(ref.n_contigs, np.int32)
    c_new[keep_contig] = np.arange(len(keep_contig), dtype=np.int32)
    sub_ref = RefSet([ref.contig_names[c] for c in keep_contig], ref.contig_len[keep_contig],
                     sp_new[ref.contig_species[keep_contig]], [ref.species_ids[s] for s in mine_sp],
                     [ref.contig_seq[c] for c in keep_contig] if ref.contig_seq is not None else None)
    if reads.n_reads:
        take = (~is_dom_sp[sp_of_read]) & (owner[sp_of_read] == rank)
        for s in dom:                                   # contiguous index ranges keep the sort order
            idx = np.nonzero(sp_of_read == s)[0]
            b = split_bounds(len(idx), n_ranks)
            take[idx[b[rank]:b[rank + 1]]] = True
        sel = np.nonzero(take)[0]
        sub_reads = reads.select(sel)
        sub_reads.ref_id = np.ascontiguousarray(c_new[sub_reads.ref_id], np.int32)
    else:
        sub_reads = reads
    return sub_ref, sub_reads, mine_sp, is_dom_sp[mine_sp]


def allreduce_split_counts(counts, stats, layout, split_mask):
    """Sum the replicated planes of split species across ranks and recompute their position stats.
    counts: torch tensor [4][gpad] (int32 bits of u32) on the rank's device (or CPU for gloo);
    stats: torch int64 [n_local][5] (genome_length, covered_bases, total_depth, aligned, mapped).
    Species that are not split are left untouched. Returns (counts, stats) updated in place."""
    import torch
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    for k in np.nonzero(np.asarray(split_mask))[0]:
        lo, hi = int(layout.species_goff[k]), int(layout.species_goff[k + 1])
        region = counts[:, lo:hi].contiguous()
        am = stats[k, 3:5].clone()
        if multi:
            dist.all_reduce(region)
            dist.all_reduce(am)
        counts[:, lo:hi] = region
        depth = (region.to(torch.int64) & 0xFFFFFFFF).sum(0)          # the planes hold u32 bit patterns
        stats[k, 1] = (depth > 0).sum()                               # tensor assignments: no host synchronisation
        stats[k, 2] = depth.sum()
        stats[k, 3:5] = am
    return counts, stats
