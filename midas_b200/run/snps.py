"""`run_midas.py snps` pileup stage on the B200 — mirror of the reference's midas/run/snps.py.

Same names and meaning as the reference module:
  Species, Contig, initialize_species (snps.py:38-53), initialize_contigs (:55-67),
  index_bam (:130-139, a no-op here), pysam_pileup (:219-244), snps_summary (:247-262),
  run_pipeline (:268-306).
`pysam_pileup(args, species, contigs)` keeps its signature and side effects (per-species
`<outdir>/snps/output/<species>.snps.gz`, Species.{genome_length, covered_bases, total_depth,
aligned_reads, mapped_reads, fraction_covered, mean_coverage}), but instead of one forked worker per
species walking pysam (utility.parallel, midas/utility.py:81-107) it decodes the BAM once into
page-locked arrays and makes ONE call into libmidas_b200.so (mb_snps_pileup_plain: H2D, re-layout of the
records on the GPU, pileup of all species, D2H) from the parent process. Under torchrun every rank takes
the species shard.plan gives it on its own GPU. There is no CPU fallback.
"""
from __future__ import annotations

import os
import shutil
import sys
from time import time

import numpy as np

from .. import bam, lib
from ..soa import Filters, RefSet
from . import io as mio

CHUNK_ROWS = 1 << 15          # rows per gzip member: ~1 MB of text, enough members to keep every host thread busy


class Species:
    """Base class for species (reference snps.py:12-31)."""

    def __init__(self, id):
        self.id = id
        self.paths = {}
        self.aligned_reads = 0
        self.mapped_reads = 0
        self.genome_length = 0
        self.covered_bases = 0
        self.total_depth = 0
        self.fraction_covered = 0
        self.mean_coverage = 0

    def fetch_paths(self, ref_db):
        indir = "%s/rep_genomes/%s" % (ref_db, self.id)
        for ext in ["", ".gz"]:
            for type in ["fna", "features"]:
                path = "%s/genome.%s%s" % (indir, type, ext)
                if os.path.isfile(path):
                    self.paths[type] = path


class Contig:
    """Base class for contig (reference snps.py:33-36)."""

    def __init__(self, id):
        self.id = id


def initialize_species(args):
    """reference snps.py:38-53; species selection from a species profile (--build_db) is upstream of
    the hot path and not part of this package."""
    species = {}
    splist = "%s/snps/species.txt" % args["outdir"]
    if args.get("build_db"):
        sys.exit("\nError: the --build_db stage (species selection, bowtie2-build) is not part of midas_b200; "
                 "run it with the reference and re-enter with --pileup")
    elif os.path.isfile(splist):
        for line in open(splist):
            id = line.rstrip()
            species[id] = Species(id)
    for sp in species.values():
        sp.fetch_paths(ref_db=args["db"])
    return species


def initialize_contigs(species):
    """reference snps.py:55-67 (dict keyed by contig id across all species; later ids overwrite)."""
    contigs = {}
    for sp in species.values():
        for rec_id, seq in mio.read_fasta(sp.paths["fna"]):
            contig = Contig(rec_id)
            contig.seq = seq.upper()
            contig.length = len(contig.seq)
            contig.species_id = sp.id
            contigs[contig.id] = contig
    return contigs


def index_bam(args):
    """reference snps.py:130-139 runs `samtools index`; the GPU path buckets reads by reference id
    itself and needs no index."""
    return None


def _filters(args) -> Filters:
    return Filters(float(args["mapid"]), float(args["aln_cov"]), int(args["readq"]), int(args["mapq"]),
                   int(args["baseq"]))


def build_refset(species, contigs, bam_names=None):
    """The contigs of the selected species as a RefSet. With the BAM header's reference names, contigs are
    numbered in BAM tid order (so that the file's coordinate order is the batch's order and the device packer
    only moves whole contigs); contigs the BAM does not name follow. species_pileup's own visiting order
    (sorted ids, snps.py:187) is applied when the rows are written, not here."""
    sp_ids = list(species.keys())
    sp_index = {s: k for k, s in enumerate(sp_ids)}
    known = {cid for cid in contigs if contigs[cid].species_id in sp_index}
    ids, seen = [], set()
    for n in (bam_names or []):
        if n in known and n not in seen:
            ids.append(n)
            seen.add(n)
    ids += [cid for cid in sorted(known) if cid not in seen]
    if any(contigs[c].length == 0 for c in ids):
        raise lib.MidasB200Error(-1, "a reference contig has length 0 (pysam: ValueError, interval of size 0)")
    ref = RefSet(ids, np.array([contigs[c].length for c in ids], np.uint32),
                 np.array([sp_index[contigs[c].species_id] for c in ids], np.int32), sp_ids, None)
    return ref


def load_alignments(bampath, species, contigs, threads=0, pinned=False):
    """Decode the BAM (native reader, seq left as in the file; page-locked arrays only on request: pinning ~0.3 GB per
    million reads costs more than it saves for a sample that crosses PCIe once) and number its records
    on a RefSet in tid order. Records on references that are not contigs of a selected species are never
    fetched by the reference (snps.py:190-199) and are dropped. Returns (RefSet, PlainReads)."""
    names, lens, reads = lib.read_bam(bampath, threads, pinned=pinned)
    ref = build_refset(species, contigs, names)
    index = {n: k for k, n in enumerate(ref.contig_names)}
    tid_map = np.array([index.get(n, -1) for n in names], np.int32) if names else np.zeros(0, np.int32)
    for n, l in zip(names, lens):
        k = index.get(n)
        if k is not None and int(ref.contig_len[k]) != int(l):
            raise lib.MidasB200Error(-5, f"BAM header length of {n} ({l}) != FASTA length ({int(ref.contig_len[k])})")
    if reads.n_reads and not np.array_equal(tid_map, np.arange(len(names), dtype=np.int32)):
        new_id = tid_map[reads.ref_id]
        keep = np.nonzero(new_id >= 0)[0]
        if len(keep) != reads.n_reads:
            reads = reads.select(keep)
            new_id = new_id[keep]
        if reads.ref_id.flags.writeable:
            reads.ref_id[...] = new_id                       # in place: the array may be page-locked
        else:
            reads.ref_id = np.ascontiguousarray(new_id, np.int32)
    return ref, reads


def write_species_tsv(path, ref: RefSet, contigs, layout, counts, species_index, threads=8, level=None):
    """<species>.snps.gz (reference snps.py:179-182 header, :201-210 rows)."""
    out = mio.GzipMembersWriter(path, level=level, threads=threads)
    out.write(b"ref_id\tref_pos\tref_allele\tdepth\tcount_a\tcount_c\tcount_g\tcount_t\n")
    index = {cid: c for c, cid in enumerate(ref.contig_names)}
    for cid in sorted(index):                                      # species_pileup's order, snps.py:187
        c = index[cid]
        if int(ref.contig_species[c]) != species_index:
            continue
        seq = np.frombuffer(contigs[cid].seq.encode("latin-1"), np.uint8)
        view = layout.contig_counts(counts, c)
        n = int(ref.contig_len[c])
        for lo in range(0, n, CHUNK_ROWS):
            hi = min(n, lo + CHUNK_ROWS)
            out.write_deferred(lambda cid=cid, lo=lo, hi=hi, seq=seq, view=view:
                               lib.format_snps_rows(cid, seq[lo:hi], lo, view[:, lo:hi]))
    out.close()


def _world():
    """(rank, world_size) of an initialised torch.distributed job (torchrun), else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def _run_batch(ctx, reads, layout, filters, threads):
    """Decoded records -> counts + stats through the C-ABI (H2D, re-layout on the GPU, pileup, D2H)."""
    try:
        return ctx.snps_pileup_plain(reads, layout, filters)
    except lib.MidasB200Error as e:
        if e.status != -6:
            raise
        # a BAM that is not coordinate-sorted (the reference sorts it, snps.py:118): sort on the host
        packed = lib.PackedReads(reads, layout, "snps", sort=True, pinned=False, threads=threads)
        return ctx.snps_pileup(packed, layout, filters)


def pysam_pileup(args, species, contigs, ctx=None):
    """reference snps.py:219-244. One GPU call for all species of this rank; fills the Species stats and writes
    one <species>.snps.gz per species. Under torchrun (one process per GPU) the species are dealt to the ranks
    by shard.plan - the reference's own unit of parallelism, one worker per species (snps.py:225-228) - a species
    holding more than 1/world of the reads is split by read index and its planes are all-reduced; the stats rows
    are gathered so that every rank (rank 0 writes summary.txt) sees all species."""
    start = time()
    print("\nCounting alleles")
    args["log"].write("\nCounting alleles\n")
    rank, world = _world()
    own_ctx = ctx is None
    timing = args.get("timing")                     # optional dict: stage -> seconds (scripts/bench_pipeline.py)

    def lap(stage, t0):
        if timing is not None:
            timing[stage] = timing.get(stage, 0.0) + (time() - t0)
        return time()

    if own_ctx:
        t = time()
        ctx = lib.Context(int(args.get("device", 0)), use_torch_stream=False)
        lap("context_create", t)

    try:
        t = time()
        threads = int(args.get("threads", 1) or 1)
        bampath = "%s/snps/temp/genomes.bam" % args["outdir"]
        filters = _filters(args)
        streamed = None
        if world == 1 and not args.get("host_decode"):
            # the streaming ingest: host threads only inflate the BGZF blocks; the record stream goes to the GPU as it
            # is and the device parses, re-lays and counts it (mb_snps_pileup_bam)
            # (the stream crosses PCIe in chunks while the host threads inflate the rest: mb_bam_stream_open_dev)
            stream = lib.BamStream(bampath, threads, ctx=ctx)
            ref = build_refset(species, contigs, stream.names)
            index = {n: k for k, n in enumerate(ref.contig_names)}
            tid_map = np.array([index.get(n, -1) for n in stream.names], np.int32)
            for n, l in zip(stream.names, stream.lens):
                k = index.get(n)
                if k is not None and int(ref.contig_len[k]) != int(l):
                    raise lib.MidasB200Error(-5, f"BAM header length of {n} ({l}) != FASTA length ({int(ref.contig_len[k])})")
            t = lap("bam_inflate", t)
            layout = lib.Layout.of(ref)
            try:
                counts, stats, n_reads, n_bases = ctx.snps_pileup_bam(stream, tid_map, layout, filters)
                streamed = True
                if timing is not None:
                    timing["n_reads"], timing["n_bases"] = n_reads, n_bases
                t = lap("gpu_h2d_parse_pack_kernels_d2h", t)
            except lib.MidasB200Error as e:
                if e.status != -6:          # not coordinate-sorted: decode and sort on the host below
                    raise
            finally:
                stream.close()
        if not streamed:
            ref, reads = load_alignments(bampath, species, contigs, threads=threads)
            if timing is not None:
                timing["n_reads"], timing["n_bases"] = int(reads.n_reads), int(reads.l_qseq.sum(dtype=np.int64))
            t = lap("bam_decode", t)
        if world == 1:
            layout = lib.Layout.of(ref)
            if not streamed:
                counts, stats = _run_batch(ctx, reads, layout, filters, threads)
                t = lap("gpu_h2d_pack_kernels_d2h", t)
            mine, writes = np.arange(ref.n_species), np.ones(ref.n_species, bool)
            sub_ref = ref
            rows = np.stack([stats[name].astype(np.int64) for name in lib.STATS_DTYPE.names], 1)
        else:
            import torch
            from .. import shard
            sub_ref, sub_reads, mine, split = shard.shard_split(reads, ref, rank, world)
            layout = lib.Layout.of(sub_ref)
            counts, stats = _run_batch(ctx, sub_reads, layout, filters, threads)
            local = np.stack([stats[name].astype(np.int64) for name in lib.STATS_DTYPE.names], 1)
            dev = torch.device("cuda", ctx.device)
            if split.any():     # the one real exchange step of the path: sum the replicas of the split species
                ct = torch.from_numpy(counts.view(np.int32)).to(dev)
                st = torch.from_numpy(local).to(dev)
                shard.allreduce_split_counts(ct, st, layout, split)
                counts = ct.cpu().numpy().view(np.uint32)
                local = st.cpu().numpy()
            rows = shard.allreduce_stats(local, mine, ref.n_species, device=dev, split_mask=split)
            writes = ~split | (rank == 0)                       # a split species is written once, by rank 0
            t = lap("gpu_h2d_pack_kernels_d2h", t)
        for k, g in enumerate(mine):
            if writes[k]:
                out_path = "%s/snps/output/%s.snps.gz" % (args["outdir"], ref.species_ids[g])
                write_species_tsv(out_path, sub_ref, contigs, layout, counts, k, threads=threads,
                                  level=args.get("gzip_level"))
        for g, species_id in enumerate(ref.species_ids):
            sp = species[species_id]
            sp.genome_length, sp.covered_bases, sp.total_depth, sp.aligned_reads, sp.mapped_reads = (int(x) for x in rows[g])
            if sp.genome_length > 0:
                sp.fraction_covered = sp.covered_bases / float(sp.genome_length)
            if sp.covered_bases > 0:
                sp.mean_coverage = sp.total_depth / float(sp.covered_bases)
        lap("tsv_gzip", t)
    finally:
        if own_ctx:
            t = time()
            ctx.close()
            lap("context_close", t)
    print("  %s minutes" % round((time() - start) / 60, 2))


SUMMARY_FIELDS = ("species_id", "genome_length", "covered_bases", "fraction_covered", "mean_coverage",
                  "aligned_reads", "mapped_reads")


def snps_summary(args, species):
    """<outdir>/snps/summary.txt (reference snps.py:247-262): one row per species, numbers printed
    with str() (floats as shortest round-trip repr), species in dict order."""
    with open("%s/snps/summary.txt" % args["outdir"], "w") as fh:
        fh.write("\t".join(SUMMARY_FIELDS) + "\n")
        for sp in species.values():
            fh.write("\t".join(str(getattr(sp, name) if name != "species_id" else sp.id)
                               for name in SUMMARY_FIELDS) + "\n")


def remove_tmp(args):
    shutil.rmtree("/".join([args["outdir"], "snps/temp"]))


def run_pipeline(args):
    """reference snps.py:268-306. Only the pileup stage (args['call']) runs here; bowtie2-build /
    bowtie2 alignment stay with the reference (host-side, out of scope)."""
    print("\nReading reference data")
    start = time()
    species = initialize_species(args)
    contigs = initialize_contigs(species)
    if args.get("timing") is not None:
        args["timing"]["read_reference"] = time() - start
    print("  %s minutes" % round((time() - start) / 60, 2))
    if args.get("align"):
        sys.exit("\nError: the --align stage (bowtie2 | samtools) is not part of midas_b200; "
                 "run it with the reference and re-enter with --pileup")
    if args.get("call"):
        index_bam(args)
        pysam_pileup(args, species, contigs)
        if _world()[0] == 0:
            snps_summary(args, species)
    if args.get("remove_temp") and _world()[0] == 0:
        remove_tmp(args)
    return species
