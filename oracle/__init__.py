"""ORACLE: CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package.  The product package `jyutvoice_b200` never does.
Parity status: PINNED by golden vectors generated from the unmodified reference
(oracle/make_golden.py -> tests/golden/); the reference's own tests hold no vector for this path.
"""
