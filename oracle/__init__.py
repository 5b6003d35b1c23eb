"""CPU oracle — test infrastructure only (see oracle/pileup_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package. The product (midas_b200/) never does.
"""
