"""midas_b200 — B200-native pileup hot path of snayfach/MIDAS behind a C-ABI.

Package layout (only what the path needs):
  csrc/      CUDA kernels (sm_100a) + the C-ABI (libmidas_b200.so)
  lib.py     ctypes binding of include/midas_b200.h (fails loudly without the .so / a GPU)
  soa.py     host-side struct-of-arrays containers
  synth.py   seeded synthetic workloads of BASELINE.json's shapes
  bam.py     BAM/BGZF codec (host ingest)
  run/       mirrors of the reference's midas/run/{snps,genes}.py call sites
"""
__version__ = "0.1.0"
