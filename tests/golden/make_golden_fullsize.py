# -*- coding: utf-8 -*-
"""
Golden vectors of the HEADLINE configuration at full size, from the CPU oracle in the CUDA path's own mode.

    python tests/golden/make_golden_fullsize.py [N]          (N = 262144: ~10-15 min of one core)

BASELINE.json configs[2]: Matern32Kernel 1-D, N = 262144, HODLRSolver(min_size=256, tol=1e-10, seed=42), inputs of
bench.py (x = sort(U(0, 10 N / 1000)), rng 1234; yerr = 0.1; y = sin x + 0.1 N(0,1)).  Mode: rng_mode = per-node streams,
exhaust = lowrank (the two documented deviations; the reference's own mode stores the exhausted blocks
densely and needs O(N^3) work at this size).  Stored: log-determinant, y^T K^-1 y, log-likelihood and, for every
internal node, (rank, rng draws, exhausted flag) plus the pivot lists — the CUDA path must reproduce the integers
exactly and the scalars to 1e-9 (tests/test_gpu_zz_fullsize.py).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from george_b200 import kernels  # noqa: E402
from george_b200._spec import flatten  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
rng = np.random.default_rng(1234)
x = np.sort(rng.uniform(0, 10 * n / 1000, n))
yerr = 0.1 * np.ones(n)
y = np.sin(x) + 0.1 * rng.normal(size=n)
t0 = time.time()
h = oracle.HODLR(flatten(1.0 * kernels.Matern32Kernel(1.0)), x, yerr, min_size=256, tol=1e-10, seed=42, rng_mode=0,
                 exhaust=1)
logdet = h.log_determinant
quad = h.dot_solve(y)
secs = time.time() - t0
ll = -0.5 * (n * np.log(2 * np.pi) + logdet) - 0.5 * quad
nodes = h.nodes()
info = np.array([[nd["rank"], nd["rng_draws"], nd["dense_fallback"], nd["is_leaf"]] for nd in nodes], dtype=np.int32)
piv_r, piv_c, piv_off = [], [], [0]
for i, nd in enumerate(nodes):
    if not nd["is_leaf"]:
        r, c = h.pivots(i, nd["rank"])
        piv_r.extend(r.tolist()); piv_c.extend(c.tolist())
    piv_off.append(len(piv_r))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg3_fullsize_n{0}.npz".format(n))
np.savez_compressed(out, n=n, log_determinant=logdet, quad=quad, log_likelihood=ll, node_info=info,
                    piv_rows=np.array(piv_r, dtype=np.int32), piv_cols=np.array(piv_c, dtype=np.int32),
                    piv_off=np.array(piv_off, dtype=np.int64), oracle_seconds=secs, oracle_evals=h.num_evals)
print("N", n, "seconds", secs, "evals", h.num_evals, "logdet", logdet, "ll", ll, "->", out)
