# -*- coding: utf-8 -*-
"""Development probe: BasicSolver (dense Cholesky) timing, config 4 family (Matern52 3-D)."""
import argparse, ctypes as C, json, sys, time
import numpy as np
sys.path.insert(0, ".")
import george_b200 as george
from george_b200 import kernels, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
rng = np.random.default_rng(1)
n = a.n
x = rng.uniform(0, 1, (n, 3)); x = x[np.argsort(x[:, 0])]
yerr = 0.1 * np.ones(n)
y = np.sin(x.sum(axis=1))
s = george.BasicSolver(1.0 * kernels.Matern52Kernel(0.5, ndim=3))
for rep in range(a.reps):
    t0 = time.perf_counter(); s.compute(x, yerr); t1 = time.perf_counter()
    d = s.dot_solve(y); t2 = time.perf_counter()
    tm = (C.c_double * 2)(); _lib.load().bgp_dense_last_timing(s._handle.ptr, tm)
    print(json.dumps({"n": n, "compute_wall_ms": (t1 - t0) * 1e3, "dot_solve_ms": (t2 - t1) * 1e3, "build_ms": tm[0],
                      "potrf_ms": tm[1], "potrf_tflops": n ** 3 / 3 / (tm[1] * 1e-3) * 1e-12,
                      "build_gbs": 8.0 * n * n / (tm[0] * 1e-3) * 1e-9, "logdet": s.log_determinant}))
if n <= 8192:
    import scipy.linalg
    K = s.kernel.get_value(x); K[np.diag_indices_from(K)] += yerr ** 2
    c = scipy.linalg.cholesky(K, lower=True)
    print("logdet ref", 2 * np.sum(np.log(np.diag(c))), "dot ref", y @ scipy.linalg.cho_solve((c, True), y), "dot", d)
