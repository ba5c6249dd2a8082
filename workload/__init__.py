"""Synthetic workloads for bench.py and the tests.  Not part of the product library."""
