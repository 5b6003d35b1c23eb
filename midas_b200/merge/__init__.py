"""Mirrors of the reference's midas/merge/{merge,snps}.py for `merge_midas.py snps`."""
