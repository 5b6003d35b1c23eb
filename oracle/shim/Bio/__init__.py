"""Minimal Biopython stand-in (test infrastructure): only Bio.SeqIO.parse(handle, 'fasta')."""
__version__ = "1.79-shim"
from . import SeqIO  # noqa: F401
