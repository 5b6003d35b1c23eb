"""GPU: the multi-GPU sharding path on CUDA tensors. shard.plan / shard_split deal the species (and split a
dominant one by read index); every "rank" runs the real CUDA pileup on its part; allreduce_split_counts (here with
the replicas summed by hand, as the NCCL all-reduce would) must reproduce the oracle's whole-sample result."""
import numpy as np
import pytest

from midas_b200 import lib, shard, synth
from midas_b200.soa import Filters
from oracle import oracle

pytestmark = pytest.mark.gpu


def _rank_result(ctx, reads, ref, rank, world, filters):
    import torch
    sub_ref, sub_reads, mine, split = shard.shard_split(reads, ref, rank, world)
    lay = lib.Layout.of(sub_ref)
    dreads = ctx.pack(sub_reads, lay, "snps")
    counts = torch.empty((4, lay.gpad), dtype=torch.int32, device="cuda")
    stats = torch.zeros((len(mine), 5), dtype=torch.int64, device="cuda")
    ctx.snps_pileup_dev(dreads, lay, filters, counts, stats)
    torch.cuda.synchronize()
    return sub_ref, lay, mine, split, counts, stats


@pytest.mark.parametrize("world", [2, 4])
def test_species_sharding_with_a_split_species_matches_the_oracle(ctx, world):
    import torch
    # one species holds most of the reads: it must be split over the ranks
    ref, reads = synth.synth_snps(seed=77, n_species=5, contigs_per_species=2, contig_len=12_000, n_reads=60_000,
                                  len_jitter=0.2, abundance="lognormal")
    counts_s = shard.species_read_counts(reads, ref)
    owner, dom = shard.plan(counts_s, world)
    if len(dom) == 0:                                   # force the situation the fallback exists for
        big = int(np.argmax(counts_s))
        sel = np.nonzero((ref.contig_species[reads.ref_id] == big) | (np.arange(reads.n_reads) % 7 == 0))[0]
        reads = reads.select(sel)
        counts_s = shard.species_read_counts(reads, ref)
        owner, dom = shard.plan(counts_s, world)
    assert len(dom) >= 1 and (owner[dom] == -1).all()
    filters = Filters()
    exp_counts, exp_stats = oracle.snps_pileup(reads, ref, filters)
    off = np.concatenate(([0], np.cumsum(ref.contig_len.astype(np.int64))))
    results = [_rank_result(ctx, reads, ref, r, world, filters) for r in range(world)]
    # what NCCL does for the split species: sum of the replicas (integer, any order)
    for sp in dom:
        regions, am = [], []
        for sub_ref, lay, mine, split, counts, stats in results:
            k = int(np.nonzero(mine == sp)[0][0])
            assert split[k]
            lo, hi = int(lay.species_goff[k]), int(lay.species_goff[k + 1])
            regions.append(counts[:, lo:hi].clone())
            am.append(stats[k, 3:5].clone())
        total, am_total = torch.stack(regions).sum(0), torch.stack(am).sum(0)
        for (sub_ref, lay, mine, split, counts, stats) in results:
            k = int(np.nonzero(mine == sp)[0][0])
            lo, hi = int(lay.species_goff[k]), int(lay.species_goff[k + 1])
            counts[:, lo:hi] = total
            stats[k, 3:5] = am_total
    # every rank: recompute the split species' position stats from the reduced planes (world 1: no collective)
    seen = np.zeros(ref.n_species, bool)
    for sub_ref, lay, mine, split, counts, stats in results:
        shard.allreduce_split_counts(counts, stats, lay, split)
        c = counts.cpu().numpy().view(np.uint32)
        st = stats.cpu().numpy()
        for k, g in enumerate(mine):
            for cl in np.nonzero(sub_ref.contig_species == k)[0]:
                cg = ref.contig_names.index(sub_ref.contig_names[cl])
                got = lay.contig_counts(c, int(cl)).astype(np.uint64)
                assert np.array_equal(got, exp_counts[:, off[cg]:off[cg + 1]]), (g, cl)
            assert np.array_equal(st[k].astype(np.uint64), exp_stats[g]), g
            seen[g] = True
    assert seen.all()
