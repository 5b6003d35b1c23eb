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


def allreduce_stats(local_stats: np.ndarray, species_index: np.ndarray, n_species: int, device=None):
    """Sum-combine per-species stats rows across ranks with torch.distributed (NCCL on GPUs, gloo on
    CPU). local_stats: int64/uint64 [n_local][5]; returns int64 [n_species][5] on every rank."""
    import torch
    import torch.distributed as dist
    full = torch.zeros((n_species, 5), dtype=torch.int64, device=device)
    if len(species_index):
        full[torch.as_tensor(np.asarray(species_index, np.int64), device=device)] = torch.as_tensor(
            np.asarray(local_stats).astype(np.int64).reshape(-1, 5), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(full)
    return full.cpu().numpy()
