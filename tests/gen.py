"""Seeded random alignments with arbitrary CIGARs for parity tests (edge cases the C2 generator
never draws: hard/soft clips on either end, =/X runs, N skips, long deletions, reads running off
the contig end, 1-base reads, N bases, variable read length)."""
from __future__ import annotations

import numpy as np

from midas_b200.soa import PlainReads, RefSet, OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_EQ, OP_X


def make_ref(rng, n_species, contigs_per_species, lo, hi):
    n = n_species * contigs_per_species
    clen = rng.integers(lo, hi + 1, n).astype(np.uint32)
    sp = np.repeat(np.arange(n_species, dtype=np.int32), contigs_per_species)
    ids = [f"Sp_{i:02d}_{1000 + i}" for i in range(n_species)]
    names = [f"{ids[sp[c]]}.c{c}" for c in range(n)]
    seqs = [np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(l))].tobytes() for l in clen]
    return RefSet(names, clen, sp, ids, seqs)


def random_reads(seed, ref: RefSet, n_reads, max_len=300, long_del_frac=0.02, sort=True, pos_hot=None,
                 min_len=1):
    """Returns PlainReads. `pos_hot`: (contig, pos, frac) puts a fraction of reads at one locus (deep pile)."""
    rng = np.random.default_rng(seed)
    recs = []
    for _ in range(n_reads):
        c = int(rng.integers(0, ref.n_contigs))
        clen = int(ref.contig_len[c])
        pos = int(rng.integers(0, clen))
        if pos_hot is not None and rng.random() < pos_hot[2]:
            c, pos = pos_hot[0], min(int(ref.contig_len[pos_hot[0]]) - 1, pos_hot[1] + int(rng.integers(0, 3)))
        lq = int(rng.integers(min_len, max_len + 1))
        # split the query into runs
        ops = []
        style = rng.random()
        remaining = lq
        if style < 0.35:
            ops = [(OP_M, lq)]
        else:
            if rng.random() < 0.2:
                ops.append((OP_H, int(rng.integers(1, 30))))
            if rng.random() < 0.4 and remaining > 1:
                s = int(rng.integers(1, max(2, min(remaining, 40))))
                s = min(s, remaining - 1)
                if s > 0:
                    ops.append((OP_S, s)); remaining -= s
            n_runs = int(rng.integers(1, 5))
            tail_s = 0
            if rng.random() < 0.4 and remaining > 1:
                tail_s = min(int(rng.integers(1, 40)), remaining - 1)
                remaining -= tail_s
            for k in range(n_runs):
                last = k == n_runs - 1
                take = remaining if last else int(rng.integers(1, max(2, remaining - (n_runs - 1 - k))))
                take = max(1, min(take, remaining - (0 if last else (n_runs - 1 - k))))
                ops.append((int(rng.choice([OP_M, OP_M, OP_M, OP_EQ, OP_X])), take))
                remaining -= take
                if last or remaining <= 0:
                    break
                g = rng.random()
                if g < 0.4 and remaining > 1:
                    ins = min(int(rng.integers(1, 6)), remaining - 1)
                    ops.append((OP_I, ins)); remaining -= ins
                elif g < 0.8:
                    big = rng.random() < long_del_frac * 5
                    ops.append((OP_D, int(rng.integers(200, 900)) if big else int(rng.integers(1, 6))))
                elif g < 0.9:
                    ops.append((OP_N, int(rng.integers(1, 50))))
            if remaining > 0:   # leftovers go to the last match run
                for k in range(len(ops) - 1, -1, -1):
                    if ops[k][0] in (OP_M, OP_EQ, OP_X):
                        ops[k] = (ops[k][0], ops[k][1] + remaining); break
                remaining = 0
            if tail_s:
                ops.append((OP_S, tail_s))
            if rng.random() < 0.1:
                ops.append((OP_H, int(rng.integers(1, 30))))
        qconsumed = sum(l for o, l in ops if o in (OP_M, OP_I, OP_S, OP_EQ, OP_X))
        assert qconsumed == lq, (ops, lq)
        seq = rng.choice(np.array([1, 2, 4, 8, 15, 3], np.uint8), lq, p=[0.24, 0.24, 0.24, 0.24, 0.03, 0.01])
        mode = rng.random()
        if mode < 0.7:
            qual = np.clip(np.rint(rng.normal(33, 7, lq)), 0, 60).astype(np.uint8)
        elif mode < 0.85:
            qual = np.clip(np.rint(rng.normal(20, 3, lq)), 0, 60).astype(np.uint8)   # near the readq threshold
        else:
            qual = rng.integers(0, 94, lq).astype(np.uint8)
        alen_hint = max(1, lq)
        nm = int(rng.choice([0, 1, 2, 3, int(0.06 * alen_hint), int(0.06 * alen_hint) + 1, int(rng.integers(0, 30))]))
        recs.append(dict(ref_id=c, pos=pos, mapq=int(rng.choice([0, 10, 19, 20, 21, 42, 60])), nm=nm,
                         cigar=ops, seq=seq, qual=qual, flag=int(rng.choice([0, 16, 256, 1024, 2048]))))
    reads = PlainReads.from_records(recs)
    if sort:
        order = np.lexsort((reads.pos, reads.ref_id))
        reads = reads.select(order)
    return reads
