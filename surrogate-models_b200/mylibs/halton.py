"""Halton sequence sampling plan -- same class / method as the reference's ``mylibs/halton.py``
(generate_sample_plan :92-112; bases default to the first primes, index i+1 so that 0 is excluded).
Host-side input generator; bit-identical to the reference (tests/test_host_logic.py)."""
from sm_b200 import workloads


class Halton:

    def __init__(self):
        pass

    def prime(self, n):
        """the n-th prime, n = 0 -> 2 (reference: halton.py:22-38)"""
        return workloads._primes(n + 1)[-1]

    def halton(self, i, prim):
        """radical inverse of i+1 in base `prim` (reference: halton.py:40-50)"""
        return workloads.halton_point(i, prim)

    def generate_sample_plan(self, point_count, dimension, bounds, base=None):
        return workloads.halton_plan(point_count, dimension, bounds, base).tolist()
