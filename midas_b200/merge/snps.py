"""`merge_midas.py snps` on the B200 — mirror of the reference's midas/merge/snps.py.

What the reference does per species (run_pipeline :469-508): paste the samples' `<species>.snps.gz` rows
into temporary text matrices (build_temp_count_matrix :233-259), then, sharded over forked workers, build a
Python `GenomicSite` per genomic site (build_sharded_tables :324-360): pooled counts, major / minor allele,
snp_type, per-sample depth and minor-allele frequency, prevalence (:42-104), the site filter (`flag`
:106-114), the gene annotation of the sites that are kept (`annotate` :116-172) and three text rows
(`write` :174-201); finally the shards are concatenated (merge_sharded_tables :388-407).

Here the sites of a species are processed in fixed-size chunks (a few hundred thousand rows, sized so that a chunk
of all samples stays near 1 GB: the reference streams one row at a time at O(samples) memory and is run on thousands
of samples, so nothing here may scale with samples x genome): the count columns of all samples are parsed straight
into one `uint32 [samples][4][chunk]` array, the per-site arithmetic of the chunk is ONE call into
libmidas_b200.so (`mb_merge_sites_dev`, csrc/merge.cu), and only the sites that survive the filter are annotated
and formatted on the host, appended to the three tables. Same outputs, byte for byte:
`<outdir>/<species>/snps_{info,freq,depth,summary}.txt` (readme.txt is this package's own wording).
Site ids are the 1-based row numbers of the samples' tables, as in the reference (:352).

The site arithmetic can be replaced by any callable with the signature of `gpu_site_engine` (the CPU test
suite passes the numpy oracle); the default, and the only one this package ships, is the GPU.
"""
from __future__ import annotations

import os

import numpy as np

from ..run import io as mio
from . import merge

ALLELES = "ACGT"
SNP_TYPES = [None, "mono", "bi", "tri", "quad"]                      # index = snp_type code of the kernel
INFO_FIELDS = ["site_id", "ref_id", "ref_pos", "ref_allele", "major_allele", "minor_allele", "count_samples",
               "count_a", "count_c", "count_g", "count_t", "locus_type", "gene_id", "snp_type", "site_type",
               "amino_acids"]
CODON_TABLE = {a + b + c: aa for (a, b, c), aa in zip(
    ((x, y, z) for x in "TCAG" for y in "TCAG" for z in "TCAG"),
    "FFLLSSSSYY__CC_WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG")}   # standard code, stop = '_' (utility.py:306-326)
COMPLEMENT = {"A": "T", "T": "A", "G": "C", "C": "G"}
ROWS_PER_CHUNK = 1 << 14                                             # kept sites per formatting task


# ------------------------------------------------------------------------------------------ site engine
def gpu_site_engine(device=0):
    """Pooled allele call of every site on the GPU. Returns f(counts u32 [S][4][G], mean_depth, allele_freq,
    site_depth, site_ratio) -> dict(pooled [4][G], major / minor int8 [G] (-1 = none), snp_type int8 [G],
    count_samples [G], sample_depth [S][G], sample_maf f64 [S][G]) as numpy arrays."""
    import torch
    from .. import lib
    ctx = lib.Context(device)

    def engine(counts, mean_depth, allele_freq, site_depth, site_ratio):
        t = torch.from_numpy(np.ascontiguousarray(counts, np.uint32).view(np.int32)).to(f"cuda:{device}")
        out = ctx.merge_sites(t, mean_depth, snp_freq=allele_freq, site_depth=site_depth, site_ratio=site_ratio)
        torch.cuda.synchronize()
        call = out["call"].cpu().numpy()
        pooled = out["pooled"].cpu().numpy().view(np.uint32)
        has_major = pooled.astype(np.uint64).sum(0) > 0
        return dict(pooled=pooled,
                    major=np.where(has_major, call & 3, -1).astype(np.int8),
                    minor=np.where((call & 0x80) != 0, (call >> 2) & 3, -1).astype(np.int8),
                    snp_type=((call >> 4) & 7).astype(np.int8),
                    count_samples=out["count_samples"].cpu().numpy(),
                    sample_depth=out["sample_depth"].cpu().numpy().view(np.uint32),
                    sample_maf=out["sample_maf"].cpu().numpy())
    return engine


# ------------------------------------------------------------------------------------------ inputs
CHUNK_TARGET_BYTES = 1 << 30          # counts in + depth / maf out of one chunk of all samples (28 B per sample and site)


def chunk_rows_for(n_samples: int) -> int:
    return int(min(1 << 20, max(1 << 14, CHUNK_TARGET_BYTES // (28 * max(1, n_samples)))))


def iter_sample_counts(species_id, samples, max_sites=float("inf"), threads=1, chunk_rows=None):
    """The samples' `<dir>/snps/output/<species>.snps.gz` tables (reference read_run_midas_snps :225-231 +
    build_temp_count_matrix :233-259), `chunk_rows` sites at a time: yields (first_site, ref_id [g], ref_pos [g],
    ref_allele [g], counts u32 [S][4][g]). Site identity comes from the first sample; the table ends with the
    shortest sample and at max_sites. Memory is O(samples x chunk_rows), whatever the genome size."""
    import pandas as pd
    from concurrent.futures import ThreadPoolExecutor
    cols = ["ref_id", "ref_pos", "ref_allele", "depth", "count_a", "count_c", "count_g", "count_t"]
    nrows = None if max_sites == float("inf") else int(max_sites)
    chunk_rows = int(chunk_rows or chunk_rows_for(len(samples)))

    def reader(k):
        path = "%s/snps/output/%s.snps.gz" % (samples[k].dir, species_id)
        return pd.read_csv(path, sep="\t", header=0, names=cols, usecols=cols if k == 0 else cols[4:], nrows=nrows,
                           keep_default_na=False, compression="gzip", engine="c", chunksize=chunk_rows,
                           dtype={"ref_id": str, "ref_pos": np.int64, "ref_allele": str, "count_a": np.uint32,
                                  "count_c": np.uint32, "count_g": np.uint32, "count_t": np.uint32})

    readers = [reader(k) for k in range(len(samples))]
    if not readers:
        return

    def pull(r):
        try:
            return next(r)
        except StopIteration:
            return None

    first_site = 0
    # one chunk per sample, parsed concurrently (inflate and the C parser release the GIL)
    with ThreadPoolExecutor(max(1, min(int(threads or 1), len(samples)))) as pool:
        while True:
            frames = list(pool.map(pull, readers))
            if any(f is None for f in frames):
                break
            g = min(len(f) for f in frames)
            if g == 0:
                break
            counts = np.empty((len(frames), 4, g), np.uint32)
            for s_, f in enumerate(frames):
                for k, c in enumerate(cols[4:]):
                    counts[s_, k] = f[c].to_numpy()[:g]
            first = frames[0]
            yield (first_site, first["ref_id"].to_numpy()[:g], first["ref_pos"].to_numpy()[:g],
                   first["ref_allele"].to_numpy()[:g], counts)
            first_site += g
            if any(len(f) != g for f in frames) or g < chunk_rows:      # the shortest sample ended inside this chunk
                break
    for r in readers:
        r.close()


def read_sample_counts(species_id, samples, max_sites=float("inf"), threads=1):
    """Whole-table form of iter_sample_counts (small inputs / tests): (ref_id, ref_pos, ref_allele, counts [S][4][G])."""
    parts = list(iter_sample_counts(species_id, samples, max_sites, threads))
    if not parts:
        return np.zeros(0, object), np.zeros(0, np.int64), np.zeros(0, object), np.zeros((len(samples), 4, 0), np.uint32)
    return (np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]),
            np.concatenate([p[3] for p in parts]), np.concatenate([p[4] for p in parts], axis=2))


def read_genes(species_id, db):
    """Coding genes of the representative genome, sorted by (scaffold_id, start, -end), each with its
    sequence oriented start -> stop (reference utility.read_genes :244-269, get_gene_seq :288-294)."""
    base = "%s/rep_genomes/%s/genome" % (db, species_id)
    fna = next((base + ".fna" + e for e in ("", ".gz") if os.path.exists(base + ".fna" + e)), None)
    feat = next((base + ".features" + e for e in ("", ".gz") if os.path.exists(base + ".features" + e)), None)
    if fna is None or feat is None:
        raise SystemExit("\nError: rep genome for %s not found\n" % species_id)
    genome = {cid: seq.upper() for cid, seq in mio.read_fasta(fna)}
    genes = []
    fh = mio.iopen(feat)
    fields = next(fh).rstrip("\n").split("\t")
    for line in fh:
        values = line.rstrip("\n").split("\t")
        if len(values) != len(fields):
            continue
        g = dict(zip(fields, values))
        if "gene_type" in g and g["gene_type"] != "CDS":
            continue
        g["start"], g["end"] = int(g["start"]), int(g["end"])
        seq = genome[g["scaffold_id"]][g["start"] - 1:g["end"]]
        g["seq"] = "".join(COMPLEMENT.get(b, b) for b in reversed(seq)) if g["strand"] == "-" else seq
        genes.append(g)
    fh.close()
    genes.sort(key=lambda g: (g["scaffold_id"], g["start"], -g["end"]))
    return genes


class GeneIndex:
    """The gene a site is annotated with. The reference walks one pointer over the sorted gene list
    (annotate :116-172): genes the site lies beyond are dropped, the first remaining gene either starts after
    the site (intergenic) or contains it. Sites arrive in ascending (ref_id, ref_pos) order, so that gene is
    the first one, in list order, whose scaffold is the site's with end >= ref_pos — found per scaffold with a
    running maximum of the ends."""

    def __init__(self, genes):
        self.by_scaffold = {}
        for g in genes:
            self.by_scaffold.setdefault(g["scaffold_id"], []).append(g)
        self.prefix_max = {sid: np.maximum.accumulate([g["end"] for g in gl]) for sid, gl in self.by_scaffold.items()}

    def find(self, ref_id, ref_pos):
        gl = self.by_scaffold.get(ref_id)
        if gl is None:
            return None
        i = int(np.searchsorted(self.prefix_max[ref_id], ref_pos, side="left"))
        if i >= len(gl) or ref_pos < gl[i]["start"]:
            return None
        return gl[i]


def annotate(ref_id, ref_pos, gene):
    """(locus_type, gene_id, site_type, amino_acids) of a kept site (reference annotate :116-172,
    fetch_ref_codon :174-182)."""
    if gene is None:
        return "IGR", None, None, None
    if gene["gene_type"] != "CDS":
        return gene["gene_type"], gene["gene_id"], None, None
    seq = gene["seq"]
    if len(seq) % 3 != 0:                                           # codons cannot be identified
        return "CDS", gene["gene_id"], None, None
    gene_pos = ref_pos - gene["start"] if gene["strand"] == "+" else gene["end"] - ref_pos
    codon_pos = gene_pos % 3
    codon = seq[gene_pos - codon_pos:gene_pos - codon_pos + 3]
    if not all(b in "ATCG" for b in codon):
        return "CDS", gene["gene_id"], None, None
    aas = []
    for allele in ALLELES:                                          # alleles are on the + strand
        b = allele if gene["strand"] == "+" else COMPLEMENT[allele]
        aas.append(CODON_TABLE[codon[:codon_pos] + b + codon[codon_pos + 1:]])
    return "CDS", gene["gene_id"], "%sD" % (4 - len(set(aas)) + 1), ",".join(aas)


# ------------------------------------------------------------------------------------------ outputs
def _na(x):
    return "NA" if x is None else x


class TableWriter:
    """snps_info.txt / snps_freq.txt / snps_depth.txt (reference write_merge_midas :276-322, write :174-201),
    appended chunk by chunk."""

    def __init__(self, species, args, genes):
        from concurrent.futures import ThreadPoolExecutor
        outdir = "%s/%s" % (args["outdir"], species.id)
        header = "\t".join(["site_id"] + [s.id for s in species.samples]) + "\n"
        self.index = GeneIndex(genes)
        self.f_info = open(outdir + "/snps_info.txt", "w")
        self.f_freq = open(outdir + "/snps_freq.txt", "wb")
        self.f_depth = open(outdir + "/snps_depth.txt", "wb")
        self.pool = ThreadPoolExecutor(max(1, int(args.get("threads", 1) or 1)))
        self.f_info.write("\t".join(INFO_FIELDS) + "\n")
        self.f_freq.write(header.encode())
        self.f_depth.write(header.encode())

    def write_chunk(self, first_site, ref_id, ref_pos, ref_allele, res, keep):
        """`keep`: indices (within the chunk) of the sites that passed the filter; site ids are 1-based row numbers
        of the samples' tables (:352), i.e. first_site + index + 1."""
        from .. import lib
        pooled, major, minor = res["pooled"], res["major"], res["minor"]
        maf = np.ascontiguousarray(res["sample_maf"], np.float64)
        dep = np.ascontiguousarray(res["sample_depth"]).astype(np.uint32, copy=False)
        keep = np.ascontiguousarray(keep, np.uint64)
        chunks = [keep[lo:lo + ROWS_PER_CHUNK] for lo in range(0, len(keep), ROWS_PER_CHUNK)]
        # the two matrices: native row formatter (mb_format_merge_*_rows), pieces formatted concurrently
        texts = self.pool.map(lambda c: lib.format_merge_rows(c, maf, dep, first_site), chunks)
        for g in keep.tolist():                                      # the annotated info rows, kept sites only
            sid, rid, pos = str(first_site + g + 1), str(ref_id[g]), int(ref_pos[g])
            locus, gene_id, site_type, aas = annotate(rid, pos, self.index.find(rid, pos))
            info = [sid, rid, str(pos), str(ref_allele[g]),
                    ALLELES[major[g]] if major[g] >= 0 else None, ALLELES[minor[g]] if minor[g] >= 0 else None,
                    str(int(res["count_samples"][g]))] + [str(int(pooled[k][g])) for k in range(4)] + \
                   [locus, gene_id, SNP_TYPES[int(res["snp_type"][g])], site_type, aas]
            self.f_info.write("\t".join(_na(x) for x in info) + "\n")
        for freq, depth in texts:
            self.f_freq.write(freq)
            self.f_depth.write(depth)

    def close(self):
        self.pool.shutdown()
        for f in (self.f_info, self.f_freq, self.f_depth):
            f.close()


def write_snps_readme(args, sp):
    with open("%s/%s/readme.txt" % (args["outdir"], sp.id), "w") as out:
        out.write("merge_midas.py snps (midas_b200): snps_freq.txt = minor allele frequency per site and sample, "
                  "snps_depth.txt = major + minor allele reads per site and sample, snps_info.txt = per-site "
                  "alleles, pooled counts and annotation, snps_summary.txt = per-sample alignment statistics.\n"
                  "Formats are those of MIDAS v1.3 (docs/merge_snvs.md). Reference database: %s/rep_genomes/%s\n"
                  % (args["db"], sp.id))


# ------------------------------------------------------------------------------------------ pipeline
def merge_species(species, args, engine, chunk_rows=None):
    """All sites of one species, a chunk of sites at a time: parse, pooled call (GPU), filter (reference flag
    :106-114), annotate, append. Returns (n_sites, n_kept)."""
    S = len(species.samples)
    writer = TableWriter(species, args, read_genes(species.id, args["db"]))
    n_sites = n_kept = 0
    try:
        for first_site, ref_id, ref_pos, ref_allele, counts in iter_sample_counts(
                species.id, species.samples, args["max_sites"], threads=args.get("threads", 1), chunk_rows=chunk_rows):
            res = engine(counts, species.sample_depth, args["allele_freq"], args["site_depth"], args["site_ratio"])
            del counts
            prevalence = np.asarray(res["count_samples"], np.int64) / float(S)           # compute_prevalence :104
            ok = ~(prevalence < args["site_prev"])
            if "any" not in args["snp_type"]:
                wanted = [t for t in range(1, 5) if SNP_TYPES[t] in args["snp_type"]]
                ok &= np.isin(res["snp_type"], wanted)
            keep = np.flatnonzero(ok)
            writer.write_chunk(first_site, ref_id, ref_pos, ref_allele, res, keep)
            n_sites += len(ref_pos)
            n_kept += len(keep)
    finally:
        writer.close()
    return n_sites, n_kept


def run_pipeline(args, engine=None, shard=None):
    """reference run_pipeline :469-508. Species are independent, so under `torchrun` (or with an explicit
    shard = (rank, world)) every process selects the same species list and merges every world-th species on
    its own GPU (LOCAL_RANK); no collective is needed, each species has its own output directory."""
    print("Identifying species and samples")
    species_list = merge.select_species(args, dtype="snps")
    for species in species_list:
        print("  %s" % species.id)
        print("    count samples: %s" % len(species.samples))
    rank, world = shard if shard is not None else (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
    if world > 1:
        species_list = species_list[rank::world]
        print("\nRank %s of %s: %s species" % (rank, world, len(species_list)))
    if engine is None and species_list:
        device = args.get("device")
        engine = gpu_site_engine(int(os.environ.get("LOCAL_RANK", "0")) if device is None else int(device))
    print("\nMerging snps")
    for species in species_list:
        print("  %s" % species.id)
        n_sites, n_kept = merge_species(species, args, engine, chunk_rows=args.get("chunk_rows"))
        print("    %s sites, %s written" % (n_sites, n_kept))
        write_snps_readme(args, species)
        species.write_sample_info(dtype="snps", outdir=args["outdir"])
    return species_list
