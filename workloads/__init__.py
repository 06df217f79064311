"""Synthetic workloads of the BASELINE.json configurations (inputs only): shared by bench.py, the tests and smoke()."""
