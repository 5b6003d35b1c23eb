"""Seeded synthetic aligned reads generated ON THE DEVICE (torch CUDA ops), for the BASELINE.json
configurations that are too large to draw with numpy inside a benchmark run (C3: 50 M reads, C4: 200 M
reads over 100 species). Same recipes as midas_b200.synth: CIGAR mix 90 % `LM`,
4 % one 1-3 bp insertion, 4 % one 1-3 bp deletion, 2 % soft-clipped ends; 1 % substitutions, 0.2 % N;
qual = clip(round(N(36,5)) [-6 on the last 20 cycles], 2, 41); mapq 42 (85 %), 24-40 (7 %), 0-19 (8 %);
NM = mismatches + inserted + deleted bases. The reference genome is never stored: the base at position p
of species s is a hash of (seed, s, p).

The output is a dict of torch CUDA tensors named like the fields of mb_plain_reads (seq in BAM nibble
packing), i.e. decoded records - what the device packer (Context.pack_dev) takes. This is benchmark /
test plumbing, not part of the product path; `to_host` gives the same records as a PlainReads for the
CPU oracle.
"""
from __future__ import annotations

import numpy as np
import torch

from .soa import PlainReads, OP_M, OP_I, OP_D, OP_S

_MASK = (1 << 62) - 1
_K1, _K2, _K3 = 0x2545F4914F6CDD1D & _MASK, 0x1B873593CC9E2D51 & _MASK, 0x27D4EB2F165667C5 & _MASK


def _ref_code(seed: int, species, pos):
    """Reference base code (0..3) at `pos` (int64 tensor) of `species` (int64 tensor or int)."""
    x = (pos + species * 4_000_037 + (int(seed) * 7_919 + 1)) & _MASK
    x = (x * _K1) & _MASK
    x = x ^ (x >> 29)
    x = (x * _K2) & _MASK
    x = x ^ (x >> 31)
    x = (x * _K3) & _MASK
    return ((x >> 33) & 3).to(torch.uint8)


def ref_codes_host(seed: int, species: int, length: int) -> np.ndarray:
    """The same hash on the host (numpy), for tests that need the reference sequence."""
    pos = np.arange(length, dtype=np.uint64)
    x = (pos + np.uint64(species * 4_000_037 + (int(seed) * 7_919 + 1))) & np.uint64(_MASK)
    with np.errstate(over="ignore"):
        x = (x * np.uint64(_K1)) & np.uint64(_MASK)
        x ^= x >> np.uint64(29)
        x = (x * np.uint64(_K2)) & np.uint64(_MASK)
        x ^= x >> np.uint64(31)
        x = (x * np.uint64(_K3)) & np.uint64(_MASK)
    return ((x >> np.uint64(33)) & np.uint64(3)).astype(np.uint8)


def _mapq(g, n, dev):
    u = torch.rand(n, device=dev, generator=g)
    hi = torch.randint(24, 41, (n,), device=dev, generator=g)
    lo = torch.randint(0, 20, (n,), device=dev, generator=g)
    return torch.where(u < 0.85, torch.full_like(hi, 42), torch.where(u < 0.92, hi, lo)).to(torch.uint8)


def _qual(g, n, L, dev):
    q = torch.randn((n, L), device=dev, generator=g, dtype=torch.float32) * 5.0 + 36.0
    q = torch.round(q)
    q[:, max(0, L - 20):] -= 6.0
    return q.clamp_(2, 41).to(torch.uint8)


def _pack_nibbles(nib):
    """[n, L] uint8 BAM codes -> [n, ceil(L/2)] bytes, even element in the high nibble."""
    n, L = nib.shape
    if L & 1:
        nib = torch.cat([nib, torch.zeros((n, 1), dtype=torch.uint8, device=nib.device)], 1)
    return (nib[:, 0::2] << 4) | nib[:, 1::2]


def synth_snps_dev(seed: int, parts, contig_len: int, read_len: int = 150, device="cuda", chunk: int = 1_000_000):
    """Coordinate-sorted reads for a list of species parts. parts: [(species_global_index, n_total, lo, hi)] -
    the reads lo..hi (by sorted position) of the species' n_total reads; local contig k = the k-th part.
    Returns a dict of CUDA tensors (mb_plain_reads fields; seq nibble-packed) + 'n_reads', 'total_bases'."""
    dev = torch.device(device)
    L = int(read_len)
    Lp = (L + 1) & ~1
    n_all = int(sum(hi - lo for _, _, lo, hi in parts))
    out = dict(ref_id=torch.empty(n_all, dtype=torch.int32, device=dev), pos=torch.empty(n_all, dtype=torch.int32, device=dev),
               mapq=torch.empty(n_all, dtype=torch.uint8, device=dev), nm=torch.empty(n_all, dtype=torch.int32, device=dev),
               l_qseq=torch.full((n_all,), L, dtype=torch.int32, device=dev),
               seq=torch.empty((n_all, Lp // 2), dtype=torch.uint8, device=dev),
               qual=torch.zeros((n_all, Lp), dtype=torch.uint8, device=dev))
    nops_all = torch.empty(n_all, dtype=torch.int64, device=dev)
    cig_parts = []
    base = 0
    qidx = torch.arange(L, device=dev, dtype=torch.int32)[None, :]
    for k, (sp, n_total, lo, hi) in enumerate(parts):
        g = torch.Generator(device=dev).manual_seed(int(seed) * 1_000_003 + int(sp))
        u = torch.rand(int(n_total), device=dev, generator=g, dtype=torch.float64)
        pos_all = torch.sort(torch.floor(u * (contig_len - (L + 3))).to(torch.int32)).values
        del u
        pos_sp = pos_all[lo:hi]
        g = torch.Generator(device=dev).manual_seed(int(seed) * 1_000_003 + int(sp) + 500_009 * (1 + int(lo)))
        for c0 in range(0, hi - lo, chunk):
            c1 = min(hi - lo, c0 + chunk)
            n = c1 - c0
            pos = pos_sp[c0:c1]
            uk = torch.rand(n, device=dev, generator=g)
            kind = (uk >= 0.90).to(torch.int32) + (uk >= 0.94).to(torch.int32) + (uk >= 0.98).to(torch.int32)
            kk = torch.randint(1, 4, (n,), device=dev, generator=g, dtype=torch.int32)
            a = torch.randint(10, L - 13, (n,), device=dev, generator=g, dtype=torch.int32)
            s = torch.randint(1, 21, (n,), device=dev, generator=g, dtype=torch.int32)
            e = torch.randint(1, 21, (n,), device=dev, generator=g, dtype=torch.int32)
            side = torch.randint(0, 3, (n,), device=dev, generator=g)
            s = torch.where(side == 1, torch.zeros_like(s), s)
            e = torch.where(side == 0, torch.zeros_like(e), e)
            k_, a_, s_, e_ = kk[:, None], a[:, None], s[:, None], e[:, None]
            is_ins, is_del, is_sc = (kind == 1)[:, None], (kind == 2)[:, None], (kind == 3)[:, None]
            zero = torch.zeros((), dtype=torch.int32, device=dev)
            shift = torch.where(is_ins & (qidx >= a_ + k_), -k_, zero) + torch.where(is_del & (qidx >= a_), k_, zero) \
                + torch.where(is_sc, -s_, zero)
            unaligned = (is_ins & (qidx >= a_) & (qidx < a_ + k_)) | (is_sc & ((qidx < s_) | (qidx >= L - e_)))
            rpos = (pos[:, None] + qidx + shift).clamp_(0, contig_len - 1).to(torch.int64)
            b = _ref_code(seed, int(sp), rpos)
            del rpos, shift
            b = torch.where(unaligned, torch.randint(0, 4, (n, L), device=dev, generator=g, dtype=torch.uint8), b)
            v = torch.randint(0, 65536, (n, L), device=dev, generator=g, dtype=torch.int32)
            sub = (v < 655) & ~unaligned                                  # ~1 % substitutions
            isn = (v >= 655) & (v < 786)                                  # ~0.2 % N
            del v
            b = torch.where(sub, (b + torch.randint(1, 4, (n, L), device=dev, generator=g, dtype=torch.uint8)) & 3, b)
            nm = sub.sum(1, dtype=torch.int32) + (isn & ~unaligned).sum(1, dtype=torch.int32) \
                + torch.where((kind == 1) | (kind == 2), kk, torch.zeros_like(kk))
            nib = (torch.ones_like(b) << b)
            nib[isn] = 15
            sl = slice(base + c0, base + c1)
            out["seq"][sl] = _pack_nibbles(nib)
            out["qual"][sl, :L] = _qual(g, n, L, dev)
            out["mapq"][sl] = _mapq(g, n, dev)
            out["nm"][sl] = nm
            out["pos"][sl] = pos
            out["ref_id"][sl] = k
            # CIGAR: three op slots per read, the valid ones flattened in order
            Lt = torch.full_like(a, L)
            op = torch.zeros((n, 3), dtype=torch.int32, device=dev)
            ln = torch.zeros((n, 3), dtype=torch.int32, device=dev)
            ok = torch.zeros((n, 3), dtype=torch.bool, device=dev)
            m0, m1, m2, m3 = kind == 0, kind == 1, kind == 2, kind == 3
            op[:, 0] = torch.where(m3, OP_S, OP_M)
            ln[:, 0] = torch.where(m0, Lt, torch.where(m3, s, a))
            ok[:, 0] = ~m3 | (s > 0)
            op[:, 1] = torch.where(m1, OP_I, torch.where(m2, OP_D, OP_M))
            ln[:, 1] = torch.where(m3, Lt - s - e, kk)
            ok[:, 1] = ~m0
            op[:, 2] = torch.where(m3, OP_S, OP_M)
            ln[:, 2] = torch.where(m1, Lt - a - kk, torch.where(m2, Lt - a, e))
            ok[:, 2] = m1 | m2 | (m3 & (e > 0))
            cig_parts.append(((ln << 4) | op)[ok])
            nops_all[sl] = ok.sum(1)
        base += hi - lo
        del pos_all
    out["cigar"] = torch.cat(cig_parts) if cig_parts else torch.zeros(0, dtype=torch.int32, device=dev)
    coff = torch.zeros(n_all + 1, dtype=torch.int64, device=dev)
    torch.cumsum(nops_all, 0, out=coff[1:])
    out["cigar_off"] = coff
    out["seq_off"] = torch.arange(n_all + 1, dtype=torch.int64, device=dev) * Lp
    out["seq"] = out["seq"].reshape(-1)
    out["qual"] = out["qual"].reshape(-1)
    out["n_reads"] = n_all
    out["total_bases"] = n_all * L
    return out


def synth_genes_dev(seed: int, gene_len: np.ndarray, gene_weight: np.ndarray, n_reads: int, read_len: int = 150,
                    device="cuda", chunk: int = 2_000_000):
    """Reads in random (file) order over a pangenome (C3 recipe): gene ~ weights, 10 % of the reads soft-clipped
    0-40 bp on either end, NM ~ Binomial(L, 0.012) with a 3 % tail of divergent reads."""
    dev = torch.device(device)
    L = int(read_len)
    Lp = (L + 1) & ~1
    n_genes = len(gene_len)
    cdf = torch.as_tensor(np.cumsum(gene_weight / gene_weight.sum()), device=dev, dtype=torch.float64)
    glen = torch.as_tensor(np.asarray(gene_len, np.int64), device=dev)
    out = dict(ref_id=torch.empty(n_reads, dtype=torch.int32, device=dev), pos=torch.empty(n_reads, dtype=torch.int32, device=dev),
               mapq=torch.empty(n_reads, dtype=torch.uint8, device=dev), nm=torch.empty(n_reads, dtype=torch.int32, device=dev),
               l_qseq=torch.full((n_reads,), L, dtype=torch.int32, device=dev),
               seq=torch.empty((n_reads, Lp // 2), dtype=torch.uint8, device=dev),
               qual=torch.zeros((n_reads, Lp), dtype=torch.uint8, device=dev))
    nops_all = torch.empty(n_reads, dtype=torch.int64, device=dev)
    cig_parts = []
    g = torch.Generator(device=dev).manual_seed(int(seed))
    for c0 in range(0, n_reads, chunk):
        c1 = min(n_reads, c0 + chunk)
        n = c1 - c0
        gi = torch.searchsorted(cdf, torch.rand(n, device=dev, generator=g, dtype=torch.float64)).clamp_(max=n_genes - 1)
        sc = torch.rand(n, device=dev, generator=g) < 0.10
        s = torch.where(sc, torch.randint(0, 41, (n,), device=dev, generator=g, dtype=torch.int32), 0)
        e = torch.where(sc, torch.randint(0, 41, (n,), device=dev, generator=g, dtype=torch.int32), 0)
        m = L - s - e
        pos = torch.floor(torch.rand(n, device=dev, generator=g, dtype=torch.float64) * (glen[gi] - m + 1).clamp_(min=1)).to(torch.int32)
        nm = torch.binomial(torch.full((n,), float(L), device=dev), torch.full((n,), 0.012, device=dev), generator=g).to(torch.int32)
        div = torch.rand(n, device=dev, generator=g) < 0.03
        nm = torch.where(div, nm + torch.randint(6, 20, (n,), device=dev, generator=g, dtype=torch.int32), nm)
        nib = torch.ones((n, L), dtype=torch.uint8, device=dev) << torch.randint(0, 4, (n, L), device=dev, generator=g, dtype=torch.uint8)
        nib[torch.randint(0, 65536, (n, L), device=dev, generator=g, dtype=torch.int32) < 131] = 15
        sl = slice(c0, c1)
        out["seq"][sl] = _pack_nibbles(nib)
        out["qual"][sl, :L] = _qual(g, n, L, dev)
        out["mapq"][sl] = _mapq(g, n, dev)
        out["nm"][sl] = nm
        out["pos"][sl] = pos
        out["ref_id"][sl] = gi.to(torch.int32)
        op = torch.zeros((n, 3), dtype=torch.int32, device=dev)
        ln = torch.zeros((n, 3), dtype=torch.int32, device=dev)
        ok = torch.zeros((n, 3), dtype=torch.bool, device=dev)
        op[:, 0] = OP_S; ln[:, 0] = s; ok[:, 0] = s > 0
        op[:, 1] = OP_M; ln[:, 1] = m; ok[:, 1] = True
        op[:, 2] = OP_S; ln[:, 2] = e; ok[:, 2] = e > 0
        cig_parts.append(((ln << 4) | op)[ok])
        nops_all[sl] = ok.sum(1)
    out["cigar"] = torch.cat(cig_parts) if cig_parts else torch.zeros(0, dtype=torch.int32, device=dev)
    coff = torch.zeros(n_reads + 1, dtype=torch.int64, device=dev)
    torch.cumsum(nops_all, 0, out=coff[1:])
    out["cigar_off"] = coff
    out["seq_off"] = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * Lp
    out["seq"] = out["seq"].reshape(-1)
    out["qual"] = out["qual"].reshape(-1)
    out["n_reads"] = n_reads
    out["total_bases"] = n_reads * L
    return out


def to_host(t: dict, lo: int = 0, hi: int | None = None) -> PlainReads:
    """Records lo..hi of a device batch as a host PlainReads (nibble-packed seq), e.g. for the CPU oracle."""
    n = t["n_reads"]
    hi = n if hi is None else hi
    coff = t["cigar_off"][lo:hi + 1].cpu().numpy().astype(np.uint64)
    soff = t["seq_off"][lo:hi + 1].cpu().numpy().astype(np.uint64)
    c0, c1, s0, s1 = int(coff[0]), int(coff[-1]), int(soff[0]), int(soff[-1])
    f = lambda k: t[k][lo:hi].cpu().numpy()
    return PlainReads(f("ref_id"), f("pos"), f("mapq"), f("nm"), f("l_qseq"), coff - np.uint64(c0),
                      t["cigar"][c0:c1].cpu().numpy().view(np.uint32), soff - np.uint64(s0),
                      t["seq"][s0 // 2:(s1 + 1) // 2].cpu().numpy(), t["qual"][s0:s1].cpu().numpy(), None, True)
