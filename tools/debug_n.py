# development probe: one compute + solve at a given N / workload on a FRESH handle
import sys, numpy as np
sys.path.insert(0, ".")
import bench
from george_b200.solvers._hodlr import HODLRSolver
name, n, ms = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
x, yerr, y = bench.make_data(n)
k = bench.make_kernel(name)
s = HODLRSolver()
for r in range(reps):
    s.compute(k, x[:, None], yerr, min_size=ms, tol=1e-10, seed=42, rng_mode="pernode", exhaust="lowrank")
    print(name, n, r, s.log_determinant, s.dot_solve(y), flush=True)
