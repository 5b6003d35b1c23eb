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
  * multi-GPU: `torchrun --nproc-per-node N scripts/run_midas.py snps OUT -d DB --pileup` runs one process per GPU;
    the species are dealt to the ranks (midas_b200.shard.plan: the reference's own unit of parallelism, one worker
    per species, midas/run/snps.py:225-228), every rank writes its species' .snps.gz, rank 0 writes summary.txt,
    log.txt and readme.txt. Without torchrun, --device picks the GPU.
  * --gzip_level N: zlib level of the .gz outputs (default 1; the reference's gzip.open() uses 9). The files hold the
    same bytes once decompressed.
log.txt (opened 'w', parameter block) and readme.txt follow the reference (:432-481, :290-336, :630-746).
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
    parser.add_argument("--gzip_level", type=int, default=None, choices=range(1, 10), metavar="1-9",
                        help="zlib level of the .gz outputs (this package only; default 1: the stage is zlib-bound)")
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


def init_distributed(args):
    """Under torchrun: one process per GPU, NCCL for the small exchanges of the path. Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args["device"] = local
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_rank(), world


def main():
    program = get_program()
    args = snp_arguments() if program == "snps" else gene_arguments()
    stage = "call" if program == "snps" else "cov"
    if args["species_id"]:
        args["species_id"] = args["species_id"].split(",")
    if not (args["build_db"] or args["align"] or args[stage]):
        args[stage] = True
        if args["m1"]:
            sys.stderr.write("Note: this package runs the %s stage only; reads given with -1 are not aligned here - run "
                             "--build_db / --align with the reference first\n" % ("--pileup" if program == "snps" else "--call_genes"))
    if args["db"] is None or not os.path.isdir(args["db"]):
        sys.exit("\nError: No reference database specified (-d / MIDAS_DB)\n")
    rank, world = init_distributed(args)
    if world > 1 and program != "snps":
        sys.exit("\nError: multi-GPU runs are provided for `snps --pileup` (species shard by rank)\n")
    for d in ("", program, program + "/output", program + "/temp"):
        os.makedirs(os.path.join(args["outdir"], d), exist_ok=True)
    bam = os.path.join(args["outdir"], program, "temp", "genomes.bam" if program == "snps" else "pangenomes.bam")
    if args[stage] and not (args["build_db"] or args["align"]) and not os.path.isfile(bam):
        sys.exit("\nError: %s not found; run the --build_db and --align stages with the reference first\n" % bam)
    from midas_b200.run import cli_text
    if rank == 0:
        args["log"] = open(os.path.join(args["outdir"], program, "log.txt"), "w")     # reference :30-33
        args["log"].write("\nMIDAS hot path on B200 (midas_b200): %s stage, %d GPU%s\n" % (program, world, "s" if world > 1 else ""))
        cli_text.print_arguments(program, args)
    else:
        args["log"] = open(os.devnull, "w")
        sys.stdout = open(os.devnull, "w")
    if program == "snps":
        from midas_b200.run import snps
        snps.run_pipeline(args)
    else:
        from midas_b200.run import genes
        genes.run_pipeline(args)
    if rank == 0:
        cli_text.write_readme(program, args)
    args["log"].close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
