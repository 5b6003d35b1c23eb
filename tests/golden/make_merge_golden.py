#!/usr/bin/env python
"""Generate tests/golden/merge_pipeline/: the files the reference's UNMODIFIED `merge_midas.py snps` pipeline
(midas/merge/snps.py run_pipeline, imported from /root/reference) writes for a small database and five
sample directories, under several argument sets. Build container only:
    python tests/golden/make_merge_golden.py
Nothing on this path is a restatement except Bio.SeqIO.parse (oracle/shim): the merge stage does not use
pysam, so these vectors pin SURVEY §8 row f.3 against the reference itself. Committed:
  input/db, input/samples    the fixture (oracle/fixtures.py::write_merge_fixture)
  cases/<name>/args.json     the argument overrides of the case
  cases/<name>/<species>/snps_{info,freq,depth,summary}.txt   what the reference wrote
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import fixtures, run_reference       # noqa: E402

CASES = {
    "default": {},
    "all_sites": dict(site_prev=0.0, snp_type=["any"], threads=3),
    "core_sites_max": dict(snp_type=["any"], max_sites=700, threads=2),
    "strict": dict(allele_freq=0.2, site_depth=5, site_ratio=1.2, site_prev=0.6, snp_type=["bi", "tri"]),
    "all_samples_one_species": dict(sample_depth=0.0, fract_cov=0.0, species_id="Sp_beta_2", site_prev=0.5,
                                    snp_type=["mono", "bi"]),
    "min_samples": dict(min_samples=5, site_prev=0.0),
}
KEEP = ["snps_info.txt", "snps_freq.txt", "snps_depth.txt", "snps_summary.txt"]


def main():
    dst = os.path.join(HERE, "merge_pipeline")
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    db, dirs = fixtures.write_merge_fixture(os.path.join(dst, "input"))
    for name, over in CASES.items():
        tmp = tempfile.mkdtemp()
        run_reference.run_merge_snps(run_reference.merge_args(os.path.join(tmp, "out"), db, dirs, **over))
        cdir = os.path.join(dst, "cases", name)
        os.makedirs(cdir)
        with open(os.path.join(cdir, "args.json"), "w") as fh:
            json.dump(over, fh, indent=1, sort_keys=True)
        for sp in sorted(os.listdir(os.path.join(tmp, "out"))):
            os.makedirs(os.path.join(cdir, sp))
            for fn in KEEP:
                shutil.copy(os.path.join(tmp, "out", sp, fn), os.path.join(cdir, sp, fn))
        shutil.rmtree(tmp)
        print(name, sorted(os.listdir(cdir)))
    print("written", dst)


if __name__ == "__main__":
    main()
