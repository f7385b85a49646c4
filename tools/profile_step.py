# -*- coding: utf-8 -*-
"""Profiling driver: STEPS x (compute + dot_solve) of the bench workload through the C ABI, nothing else.
Used under ncu (see profiles/README.md); numbers printed under a profiler are never bench values."""
import argparse, sys
import numpy as np
sys.path.insert(0, ".")
import bench
from george_b200.solvers._hodlr import HODLRSolver

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--exhaust", default="lowrank")
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
n = a.n or wl["n"]
x, yerr, y = bench.make_data(n)
k = bench.make_kernel(a.workload)
s = HODLRSolver()
for i in range(a.steps):
    s.compute(k, x[:, None], yerr, min_size=wl["min_size"], tol=wl["tol"], seed=42, exhaust=a.exhaust)
    d = s.dot_solve(y)
    print(i, s.log_determinant, d, s.timing())
