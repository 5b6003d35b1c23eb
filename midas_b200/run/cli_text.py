"""The text the reference's run_midas.py writes around the two stages this package runs: the parameter
block of <outdir>/<program>/log.txt (reference scripts/run_midas.py print_snp_arguments :432-481,
print_gene_arguments :290-336) and <outdir>/<program>/readme.txt (write_readme :630-746). Both are part of
the on-disk surface a user of the reference finds after a run; tests/test_cli_text.py holds them identical to
what the reference's own functions write (tests/golden/cli/, produced by tests/golden/make_golden_cli.py).
The text is assembled from the tables below: one row per file / column, as the formats are documented."""
from __future__ import annotations

import sys

# ---------------------------------------------------------------------------------------------- log.txt
_STAGE_LINE = {"snps": "use samtools to generate pileups and count variants",
               "genes": "quantify coverage of pangenomes genes"}
_DB_LINE = {"snps": "build bowtie2 database of genomes", "genes": "build bowtie2 database of pangenomes"}
_ALIGN_LINE = {"snps": "align reads to bowtie2 genome database", "genes": "align reads to bowtie2 pangenome database"}
# option label -> args key, in print order
_SNP_OPTIONS = (("minimum alignment percent identity", "mapid"), ("minimum mapping quality score", "mapq"),
                ("minimum base quality score", "baseq"), ("minimum read quality score", "readq"),
                ("minimum alignment coverage of reads", "aln_cov"))
_GENE_OPTIONS = (("minimum alignment percent identity", "mapid"), ("minimum alignment coverage of reads", "aln_cov"),
                 ("minimum read quality score", "readq"), ("minimum mapping quality score", "mapq"))
_SNP_SWITCHES = (("discard", "discard discordant read-pairs"), ("baq", "enable BAQ (per-base alignment quality)"),
                 ("adjust_mq", "adjust MAPQ"))


def parameter_lines(program: str, args: dict) -> list[str]:
    stage = "call" if program == "snps" else "cov"
    out = ["===========Parameters===========", "Command: %s" % " ".join(sys.argv),
           "Script: run_midas.py %s" % program, "Database: %s" % args["db"], "Output directory: %s" % args["outdir"],
           "Remove temporary files: %s" % args["remove_temp"], "Pipeline options:"]
    if args.get("build_db"):
        out.append("  " + _DB_LINE[program])
    if args.get("align"):
        out.append("  " + _ALIGN_LINE[program])
    if args.get(stage):
        out.append("  " + _STAGE_LINE[program])
    if args.get("build_db"):
        out.append("Database options:")
        if args.get("species_topn"):
            out.append("  include top %s most abundant species" % args["species_topn"])
        if args.get("species_cov"):
            out.append("  include all species with >=%sX genome coverage" % args["species_cov"])
        if args.get("species_id"):
            out.append("  include specified species id(s): %s" % args["species_id"])
    if args.get("align"):
        out.append("Read alignment options:")
        if args.get("interleaved"):
            out.append("  input reads (1st + 2nd mate): %s" % args["m1"])
        elif args.get("m2"):
            out += ["  input reads (1st mate): %s" % args["m1"], "  input reads (2nd mate): %s" % args["m2"]]
        else:
            out.append("  input reads (unpaired): %s" % args["m1"])
        out += ["  alignment speed/sensitivity: %s" % args["speed"], "  alignment mode: %s" % args["mode"],
                "  number of reads to use from input: %s" % (args["max_reads"] if args.get("max_reads") else "use all"),
                "  number of threads for database search: %s" % args["threads"]]
    if args.get(stage):
        out.append("SNP calling options:" if program == "snps" else "Gene coverage options:")
        for label, key in (_SNP_OPTIONS if program == "snps" else _GENE_OPTIONS):
            out.append("  %s: %s" % (label, args[key]))
        out.append("  trim %s base-pairs from 3'/right end of read" % args["trim"])
        if program == "snps":
            out += ["  " + text for key, text in _SNP_SWITCHES if args.get(key)]
    out.append("================================")
    return out


def print_arguments(program: str, args: dict) -> None:
    """Parameter block to log.txt and stdout, as the reference does before running a program."""
    text = "\n".join(parameter_lines(program, args)) + "\n"
    args["log"].write(text)
    sys.stdout.write(text)


# ------------------------------------------------------------------------------------------- readme.txt
_COMMON_FILES = (("species.txt", ("list of species_ids included in local database",)),
                 ("summary.txt", ("tab-delimited with header", "summarizes alignment results per-species")),
                 ("log.txt", ("log file containing parameters used",)),
                 ("temp", ("directory of intermediate files", "run with `--remove_temp` to remove these files")))
_TAIL_READS = (("aligned_reads", "number of aligned reads BEFORE quality filtering"),
               ("mapped_reads", "number of aligned reads AFTER quality filtering"))
_README = {
    "snps": dict(
        ext="snps", db_dir="rep_genomes",
        columns=(("ref_id", "id of reference scaffold/contig/genome"), ("ref_pos", "position in ref_id (1-indexed)"),
                 ("ref_allele", "reference nucleotide"), ("depth", "number of mapped reads"),
                 ("count_a", "count of A allele"), ("count_c", "count of C allele"), ("count_g", "count of G allele"),
                 ("count_t", "count of T allele")),
        summary=(("species_id", "species id"), ("genome_length", "number of base pairs in representative genome"),
                 ("covered_bases", "number of reference sites with at least 1 mapped read"),
                 ("fraction_covered", "proportion of reference sites with at least 1 mapped read"),
                 ("mean_coverage", "average read-depth across reference sites with at least 1 mapped read")) + _TAIL_READS,
        after_summary="  \n"),
    "genes": dict(
        ext="genes", db_dir="pan_genomes",
        columns=(("gene_id", "id of non-redundant gene used for read mapping; 'peg' and 'rna' indicate coding & RNA "
                             "genes respectively"),
                 ("count_reads", "number of aligned reads to gene_id after quality filtering"),
                 ("coverage", "average read-depth of gene_id based on aligned reads (# aligned bp / gene length in bp)"),
                 ("copy_number", "estimated copy-number of gene_id based on aligned reads (coverage of gene_id / median "
                                 "coverage of 15 universal single copy genes)")),
        summary=(("species_id", "species id"), ("pangenome_size", "number of non-redundant genes in reference pan-genome"),
                 ("covered_genes", "number of genes with at least 1 mapped read"),
                 ("fraction_covered", "proportion of genes with at least 1 mapped read"),
                 ("mean_coverage", "average read-depth across genes with at least 1 mapped read"),
                 ("marker_coverage", "median read-depth across 15 universal single copy genes")) + _TAIL_READS,
        after_summary="\n"),
}


def readme_text(program: str, args: dict) -> str:
    r = _README[program]
    out = ["", "Description of output files and file formats from 'run_midas.py %s'" % program, "", "Output files",
           "############", "output", "  directory of per-species output files",
           "  files are tab-delimited, gzip-compressed, with header",
           "  naming convention of each file is: {SPECIES_ID}.%s.gz" % r["ext"]]
    for name, lines in _COMMON_FILES:
        out.append(name)
        out += ["  " + ln for ln in lines]
    out += ["", "Output formats", "############", "output/{SPECIES_ID}.%s.gz" % r["ext"]]
    out += ["  %s: %s" % kv for kv in r["columns"]]
    out += ["", "summary.txt"]
    out += ["  %s: %s" % kv for kv in r["summary"]]
    text = "\n".join(out) + "\n" + r["after_summary"]
    text += "Additional information for each species can be found in the reference database:\n %s/%s\n" % (args["db"], r["db_dir"])
    return text


def write_readme(program: str, args: dict) -> None:
    with open("%s/%s/readme.txt" % (args["outdir"], program), "w") as fh:
        fh.write(readme_text(program, args))
