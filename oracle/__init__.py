"""Test-only oracles (tier rule 3).

Nothing in the product package (s3gaussian_b200/) imports this package.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may, and only as the checker.
"""
