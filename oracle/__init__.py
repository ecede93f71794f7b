"""CPU checkers for the stacking hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this package.  The product library (siril-0.9_b200/) never does.

  oracle.refrun  -- ctypes front end of oracle/_ref/libsiril_ref.so: the
                    reference's own stacking.c / statistics.c compiled unmodified.
  oracle.port    -- ctypes front end of oracle/libsiril_oracle.so: the independent
                    C restatement (siril_oracle.c).
"""
