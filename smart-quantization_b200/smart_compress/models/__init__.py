"""Benchmark drivers only: the networks BASELINE.json's training configs name, with synthetic inputs
and random initialisation.  Not part of the codec; they exist so that the hooks see the tensor shapes
and call counts of the reference's workloads."""
