"""CPU oracle for the brute-force likelihood path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Product code (brutus_b200/) must never import it.
"""
from .oracle import (RefOptions, build, loglike, loglike_batch, select, num_threads)  # noqa: F401
