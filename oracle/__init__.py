"""CPU oracle for the Smith-Waterman hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product never does.
"""
from .oracle import Oracle, OracleParams, build_oracle  # noqa: F401
