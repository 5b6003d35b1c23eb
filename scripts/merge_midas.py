#!/usr/bin/env python
"""merge_midas.py snps - the reference CLI (scripts/merge_midas.py of snayfach/MIDAS v1.3.0, snps flags
:150-262, presets :264-286, input handling :309-331) for the multi-sample SNP calling stage, with the
per-site arithmetic on the B200 (midas_b200/merge/snps.py -> mb_merge_sites_dev).

Flags, defaults, presets and the files written under <outdir>/<species_id>/ follow the reference. What differs:
  * `species` and `genes` are not provided (outside the pileup path and its consumers);
  * --threads sets the host threads that parse the samples' tables and format the output rows; the site
    arithmetic itself is one GPU call per species instead of forked site shards;
  * --device selects the CUDA device; under torchrun every rank merges every WORLD_SIZE-th species on its own
    GPU (species are independent: no collective); readme.txt is this package's own wording.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def get_program():
    if len(sys.argv) == 1 or sys.argv[1] in ["-h", "--help"]:
        print("usage: merge_midas.py snps outdir -i INPUT -t {list,file,dir} [options]\n\n"
              "  snps    multi-sample core-genome SNP calling (pooled allele call on the B200)")
        sys.exit(0)
    if sys.argv[1] != "snps":
        sys.exit("\nError: unrecognized command: '%s' (this package provides snps)\n" % sys.argv[1])
    return sys.argv[1]


def snps_arguments(argv=None):
    p = argparse.ArgumentParser(usage="merge_midas.py snps <outdir> [options]")
    p.add_argument("program", help=argparse.SUPPRESS)
    p.add_argument("outdir", type=str, help="directory for output files; one subdirectory per species_id")
    p.add_argument("--threads", type=int, default=1, help="host threads for parsing and formatting (1)")
    p.add_argument("--device", type=int, default=None,
                   help="CUDA device (this package only; default: LOCAL_RANK, so that `torchrun --nproc-per-node N "
                        "scripts/merge_midas.py snps ...` merges every N-th species on each of N GPUs)")
    p.add_argument("-i", type=str, dest="input", required=True, help="sample directories written by run_midas.py; see -t")
    p.add_argument("-t", choices=["list", "file", "dir"], dest="intype", required=True,
                   help="list: comma-separated paths; dir: directory of sample directories; file: one path per line")
    p.add_argument("-d", type=str, dest="db", default=os.environ.get("MIDAS_DB"), help="path to reference database")
    for preset in ("core_snps", "core_sites", "all_snps", "all_sites"):
        p.add_argument("--" + preset, action="store_true")
    p.add_argument("--min_samples", type=int, default=1)
    p.add_argument("--species_id", type=str)
    p.add_argument("--max_species", type=int)
    p.add_argument("--sample_depth", type=float, default=5.0)
    p.add_argument("--fract_cov", type=float, default=0.4)
    p.add_argument("--max_samples", type=int)
    p.add_argument("--all_samples", default=False, action="store_true")
    p.add_argument("--snp_type", choices=["any", "mono", "bi", "tri", "quad"], nargs="+", default=["bi"])
    p.add_argument("--allele_freq", type=float, default=0.01)
    p.add_argument("--site_depth", type=int, default=1)
    p.add_argument("--site_ratio", type=float, default=2.0)
    p.add_argument("--site_prev", type=float, default=0.95)
    p.add_argument("--max_sites", type=int, default=float("Inf"))
    return add_snp_presets(vars(p.parse_args(argv)))


def add_snp_presets(args):
    """Presets overwrite the individual flags, in the reference's order (scripts/merge_midas.py:264-286)."""
    if args["all_samples"]:
        args.update(sample_depth=0.0, fract_cov=0.0)
    if args["all_sites"]:
        args.update(site_prev=0.0, snp_type=["any"])
    if args["all_snps"]:
        args.update(site_prev=0.0, snp_type=["bi"])
    if args["core_sites"]:
        args.update(site_depth=1, site_ratio=2.0, site_prev=0.95, snp_type=["any"])
    if args["core_snps"]:
        args.update(site_depth=1, site_ratio=2.0, site_prev=0.95, snp_type=["bi"])
    return args


def check_arguments(args):
    for arg in ["allele_freq", "fract_cov", "site_prev"]:
        if args[arg] and (args[arg] < 0 or args[arg] > 1):
            sys.exit("\nError: --%s must be between 0.0 and 1.0\n" % arg)
    for arg in ["max_samples", "min_samples", "max_species", "threads", "site_depth", "max_sites", "sample_depth",
                "site_ratio"]:
        if args[arg] and args[arg] < 0:
            sys.exit("\nError: --%s cannot be a negative value\n" % arg)
    if not args["db"] or not os.path.isdir(args["db"]):
        sys.exit("\nError: no database specified or found: use -d or set MIDAS_DB\n")
    args["indirs"] = input_directories(args)
    os.makedirs(args["outdir"], exist_ok=True)


def input_directories(args):
    error = "\nError: specified input %s does not exist: %s\n"
    kind, src = args["intype"], args["input"]
    if kind == "dir":
        if not os.path.isdir(src):
            sys.exit(error % (kind, os.path.abspath(src)))
        return [os.path.join(src, d) for d in os.listdir(src)]
    if kind == "file":
        if not os.path.isfile(src):
            sys.exit(error % (kind, os.path.abspath(src)))
        dirs = [line.rstrip().rstrip("/") for line in open(src)]
    else:
        dirs = src.split(",")
    for d in dirs:
        if not os.path.isdir(d):
            sys.exit(error % ("dir", d))
    return dirs


def main():
    get_program()
    args = snps_arguments()
    check_arguments(args)
    from midas_b200.merge import snps
    snps.run_pipeline(args)


if __name__ == "__main__":
    main()
