"""Load a committed golden case (tests/golden/<name>) back into RefSet / PlainReads / Filters."""
import json
import os

import numpy as np

from midas_b200.soa import Filters, PlainReads, RefSet

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SNPS_CASES = ["snps_c2_small", "snps_random_cigars", "snps_no_filters"]
GENES_CASES = ["genes_c3_small"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name, "input.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    reads = PlainReads(z["ref_id"], z["pos"], z["mapq"], z["nm"], z["l_qseq"], z["cigar_off"], z["cigar"],
                       z["seq_off"], z["seq"], z["qual"], z["flag"])
    seqs = [s.encode() for s in meta["contig_seq"]] if meta["contig_seq"] else None
    ref = RefSet(meta["contig_names"], z["contig_len"], z["contig_species"], meta["species_ids"], seqs)
    f = meta["filters"]
    filters = Filters(f["mapid"], f["aln_cov"], f["readq"], f["mapq"], f["baseq"])
    return ref, reads, filters, meta["extra"]


def expected(name):
    d = os.path.join(GOLDEN, name, "expected")
    return {fn: open(os.path.join(d, fn), "rb").read() for fn in sorted(os.listdir(d))}
