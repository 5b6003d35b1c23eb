#!/usr/bin/env python
"""Generate tests/golden/cli/: what the reference's UNMODIFIED scripts/run_midas.py writes around the hot path -
readme.txt (write_readme, :630-746) and the parameter block of log.txt (print_snp_arguments :432-481,
print_gene_arguments :290-336) - for fixed argument sets. Build container only (imports /root/reference over
oracle/shim). The cases (argument dicts) are stored beside the outputs so the test replays them."""
import importlib.util
import io
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("MIDAS_REFERENCE", "/root/reference")
for p in (REF, os.path.join(ROOT, "oracle", "shim")):
    if p not in sys.path:
        sys.path.insert(0, p)

BASE = dict(remove_temp=False, build_db=False, align=False, species_topn=None, species_cov=None, species_id=None,
            m1=None, m2=None, interleaved=False, speed="very-sensitive", mode="global", max_reads=None, threads=1,
            trim=0)
CASES = {
    "snps_pileup_defaults": ("snps", dict(BASE, call=True, mapid=94.0, mapq=20, baseq=30, readq=20, aln_cov=0.75,
                                          discard=False, baq=False, adjust_mq=False)),
    "snps_pileup_flags": ("snps", dict(BASE, call=True, remove_temp=True, mapid=97.5, mapq=30, baseq=35, readq=25,
                                       aln_cov=0.9, trim=5, discard=True, baq=True, adjust_mq=True, threads=8)),
    "genes_call_defaults": ("genes", dict(BASE, cov=True, mapid=94.0, mapq=0, readq=20, aln_cov=0.75, mode="local")),
    "genes_call_strict": ("genes", dict(BASE, cov=True, remove_temp=True, mapid=99.0, mapq=10, readq=30, aln_cov=0.95,
                                        trim=3, mode="local")),
}
ARGV = ["run_midas.py", "snps", "OUT", "--pileup"]


def main():
    spec = importlib.util.spec_from_file_location("ref_run_midas", os.path.join(REF, "scripts", "run_midas.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    dst = os.path.join(HERE, "cli")
    os.makedirs(dst, exist_ok=True)
    json.dump({"argv": ARGV, "cases": CASES}, open(os.path.join(dst, "cases.json"), "w"), indent=1, sort_keys=True)
    for name, (program, a) in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(os.path.join(tmp, program))
            args = dict(a, outdir=tmp, db="/db/midas_db_v1.2", log=io.StringIO())
            old_argv, old_out = sys.argv, sys.stdout
            sys.argv, sys.stdout = list(ARGV), io.StringIO()
            try:
                (mod.print_snp_arguments if program == "snps" else mod.print_gene_arguments)(args)
                mod.write_readme(program, args)
            finally:
                sys.argv, sys.stdout = old_argv, old_out
            text = args["log"].getvalue().replace(tmp, "OUT")
            open(os.path.join(dst, name + ".params.txt"), "w").write(text)
            open(os.path.join(dst, name + ".readme.txt"), "w").write(open(os.path.join(tmp, program, "readme.txt")).read())
    print("wrote", dst)


if __name__ == "__main__":
    main()
