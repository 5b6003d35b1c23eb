"""Drop-in mirrors of the reference's per-sample hot-path modules (midas/run/snps.py, midas/run/genes.py):
same function names, `args` dict keys, Species/Contig/Gene attributes and on-disk outputs; the record
loops run on the B200 through libmidas_b200.so."""
