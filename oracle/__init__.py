"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Restates, in numpy + SciPy (SuperLU), what the reference computes on the hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package; nothing under
solverz_b200/ does.  Pinning: see oracle/README.md and tests/golden/.
"""
