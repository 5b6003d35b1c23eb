#!/usr/bin/env python
"""run_midas.py {genes,snps} - the reference CLI (scripts/run_midas.py of snayfach/MIDAS v1.3.0) for
the two stages this package accelerates: `snps --pileup` and `genes --call_genes`.

Flags, defaults, directory layout, log.txt and the on-disk outputs follow the reference
(snps flags scripts/run_midas.py:370-430, genes flags :236-289, directories :195-202). What differs:
  * --build_db / --align (bowtie2-build, bowtie2 | samtools) stay with the reference: run them
    there, then re-enter here with --pileup / --call_genes (the reference's own stage flags);
    giving no stage flag means "the stages this package owns", not all three.
  * -1 is not required for those stages (reference :391 requires it even for --pileup).
  * utility.add_executables (midas/utility.py:109-150) is not called: no samtools / bowtie2 needed.
  * `species` is not provided (alignment-bound, upstream of the hot path).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def get_program():
    if len(sys.argv) == 1 or sys.argv[1] in ["-h", "--help"]:
        print("usage: run_midas.py {genes,snps} outdir [options]\n\n"
              "  genes   map reads to pangenomes: quantify gene coverage (--call_genes stage on the B200)\n"
              "  snps    map reads to representative genomes: per-position A/C/G/T counts (--pileup stage on the B200)")
        sys.exit(0)
    if sys.argv[1] not in ("genes", "snps"):
        sys.exit("\nError: unrecognized command: '%s' (this package provides genes and snps)\n" % sys.argv[1])
    return sys.argv[1]


def _common(parser):
    parser.add_argument("program", help=argparse.SUPPRESS)
    parser.add_argument("outdir", type=str, help="path to directory to store results")
    parser.add_argument("--remove_temp", default=False, action="store_true")
    parser.add_argument("-d", type=str, dest="db", default=os.environ.get("MIDAS_DB"), help="path to reference database")
    parser.add_argument("--species_cov", type=float)
    parser.add_argument("--species_topn", type=int)
    parser.add_argument("--species_id", type=str)
    parser.add_argument("-1", type=str, dest="m1")
    parser.add_argument("-2", type=str, dest="m2")
    parser.add_argument("--interleaved", action="store_true", default=False)
    parser.add_argument("-s", type=str, dest="speed", default="very-sensitive",
                        choices=["very-fast", "fast", "sensitive", "very-sensitive"])
    parser.add_argument("-m", type=str, dest="mode", default="global", choices=["local", "global"])
    parser.add_argument("-n", type=int, dest="max_reads")
    parser.add_argument("-t", dest="threads", default=1)
    parser.add_argument("--device", type=int, default=0, help="CUDA device (this package only)")
    parser.add_argument("--build_db", action="store_true", default=False)
    parser.add_argument("--align", action="store_true", default=False)


def snp_arguments():
    p = argparse.ArgumentParser(usage="run_midas.py snps <outdir> [options]")
    _common(p)
    p.add_argument("--pileup", action="store_true", dest="call", default=False)
    p.add_argument("--mapid", type=float, default=94.0)
    p.add_argument("--mapq", type=int, default=20)
    p.add_argument("--baseq", type=int, default=30)
    p.add_argument("--readq", type=int, default=20)
    p.add_argument("--aln_cov", type=float, default=0.75)
    p.add_argument("--trim", type=int, default=0)
    # parsed and printed but never read by the reference's pileup
    p.add_argument("--discard", default=False, action="store_true")
    p.add_argument("--baq", default=False, action="store_true")
    p.add_argument("--adjust_mq", default=False, action="store_true")
    return vars(p.parse_args())


def gene_arguments():
    p = argparse.ArgumentParser(usage="run_midas.py genes <outdir> [options]")
    _common(p)
    p.add_argument("--call_genes", action="store_true", dest="cov", default=False)
    p.add_argument("--readq", type=int, default=20)
    p.add_argument("--mapid", type=float, default=94.0)
    p.add_argument("--mapq", type=int, default=0)
    p.add_argument("--aln_cov", type=float, default=0.75)
    p.add_argument("--trim", type=int, default=0)
    p.add_argument("--exact_order", dest="exact_order", action="store_true", default=False,
                   help="reproduce the reference's running float sum of alen/gene_length in BAM order bit for bit "
                        "(adds a device sort); default: sum mapped bp as an integer and divide once, coverage "
                        "within 1e-12 of the reference")
    return vars(p.parse_args())


def main():
    program = get_program()
    args = snp_arguments() if program == "snps" else gene_arguments()
    stage = "call" if program == "snps" else "cov"
    if not (args["build_db"] or args["align"] or args[stage]):
        args[stage] = True
    if args["db"] is None or not os.path.isdir(args["db"]):
        sys.exit("\nError: No reference database specified (-d / MIDAS_DB)\n")
    for d in ("", program, program + "/output", program + "/temp"):
        os.makedirs(os.path.join(args["outdir"], d), exist_ok=True)
    bam = os.path.join(args["outdir"], program, "temp", "genomes.bam" if program == "snps" else "pangenomes.bam")
    if args[stage] and not (args["build_db"] or args["align"]) and not os.path.isfile(bam):
        sys.exit("\nError: %s not found; run the --build_db and --align stages with the reference first\n" % bam)
    args["log"] = open(os.path.join(args["outdir"], program, "log.txt"), "a")
    args["log"].write("\nMIDAS (midas_b200): %s stage on CUDA device %d\n" % (program, args["device"]))
    if program == "snps":
        from midas_b200.run import snps
        snps.run_pipeline(args)
    else:
        from midas_b200.run import genes
        genes.run_pipeline(args)
    args["log"].close()


if __name__ == "__main__":
    main()
