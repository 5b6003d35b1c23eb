"""CPU: the N>1 path - species sharding and the stats exchange - with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import gen
from midas_b200 import shard
from midas_b200.soa import Filters


def test_assign_species_balances_and_is_deterministic():
    counts = [100, 5, 5, 90, 40, 40, 1, 0]
    own = shard.assign_species(counts, 3)
    load = np.bincount(own, weights=counts, minlength=3)
    assert load.max() <= 100 and load.sum() == sum(counts)
    assert np.array_equal(own, shard.assign_species(counts, 3))
    assert (shard.assign_species(counts, 1) == 0).all()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    ref = gen.make_ref(rng, 5, 2, 300, 1500)
    reads = gen.random_reads(12, ref, 1500, max_len=100)
    sub_ref, sub_reads, mine = shard.shard(reads, ref, rank, world)
    f = Filters(mapid=80.0, aln_cov=0.3, readq=10, mapq=0, baseq=10)
    # the oracle stands in for the GPU call: what is under test is the sharding + exchange
    counts, stats = oracle.snps_pileup(sub_reads, sub_ref, f)
    full = shard.allreduce_stats(stats, mine, ref.n_species)
    exp_counts, exp_stats = oracle.snps_pileup(reads, ref, f)
    ok = np.array_equal(full.astype(np.uint64), exp_stats)
    # every species' counts on its owner equal the unsharded result
    off = np.concatenate(([0], np.cumsum(ref.contig_len.astype(np.int64))))
    soff = np.concatenate(([0], np.cumsum(sub_ref.contig_len.astype(np.int64))))
    for k, name in enumerate(sub_ref.contig_names):
        c = ref.contig_names.index(name)
        ok = ok and np.array_equal(counts[:, soff[k]:soff[k + 1]], exp_counts[:, off[c]:off[c + 1]])
    q.put((rank, bool(ok), [int(x) for x in mine]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_unsharded_result():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    owned = sorted(x for _, _, mine in res for x in mine)
    assert owned == [0, 1, 2, 3, 4]


def _worker_split(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(21)
    ref = gen.make_ref(rng, 3, 2, 400, 1200)
    reads = gen.random_reads(22, ref, 2000, max_len=90, pos_hot=(0, 100, 0.7), long_del_frac=0.0)  # species 0 dominates
    f = Filters(mapid=80.0, aln_cov=0.3, readq=10, mapq=0, baseq=10)
    sub_ref, sub_reads, mine, split = shard.shard_split(reads, ref, rank, world)

    class Lay:        # unpadded stand-in for lib.Layout: species regions back to back
        pass
    lay = Lay()
    lens = [int(sub_ref.contig_len[sub_ref.contig_species == k].sum()) for k in range(sub_ref.n_species)]
    lay.species_goff = np.concatenate(([0], np.cumsum(lens)))
    counts, stats = oracle.snps_pileup(sub_reads, sub_ref, f)          # stands in for the GPU call
    # oracle counts are in contig order; contigs of a species are consecutive here
    tc = torch.as_tensor(counts.astype(np.int64))
    ts = torch.as_tensor(stats.astype(np.int64))
    shard.allreduce_split_counts(tc, ts, lay, split)
    full = shard.allreduce_stats(ts.numpy(), mine, ref.n_species, split_mask=split)
    exp_counts, exp_stats = oracle.snps_pileup(reads, ref, f)
    off = np.concatenate(([0], np.cumsum(ref.contig_len.astype(np.int64))))
    soff = np.concatenate(([0], np.cumsum(sub_ref.contig_len.astype(np.int64))))
    ok = bool(split.any()) and np.array_equal(full.astype(np.uint64), exp_stats)
    for k, name in enumerate(sub_ref.contig_names):
        c = ref.contig_names.index(name)
        ok = ok and np.array_equal(tc.numpy()[:, soff[k]:soff[k + 1]].astype(np.uint64), exp_counts[:, off[c]:off[c + 1]])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_dominant_species_is_split_and_all_reduced():
    assert list(shard.dominant_species([90, 5, 5], 2)) == [0] and len(shard.dominant_species([50, 50], 2)) == 0
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_split, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
