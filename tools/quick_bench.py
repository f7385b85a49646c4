# -*- coding: utf-8 -*-
"""Development probe: time compute()+dot_solve() for a BASELINE.json config at a given N and print the per-level
rank / fallback picture.  Not the judged benchmark (that is bench.py)."""
import argparse
import json
import sys
import time
from collections import defaultdict

import numpy as np

sys.path.insert(0, ".")
from george_b200 import kernels  # noqa: E402
from george_b200.solvers._hodlr import HODLRSolver  # noqa: E402


def make(cfg, n):
    rng = np.random.default_rng(1234)
    x = np.sort(rng.uniform(0, 10 * n / 1000, n))[:, None]
    yerr = 0.1 * np.ones(n)
    y = np.sin(x[:, 0]) + 0.1 * rng.normal(size=n)
    if cfg == 2:
        return 1.0 * kernels.ExpSquaredKernel(1.0), x, yerr, y, dict(min_size=100, tol=1e-10)
    if cfg == 3:
        return 1.0 * kernels.Matern32Kernel(1.0), x, yerr, y, dict(min_size=256, tol=1e-10)
    if cfg == 5:
        k = 1.0 * kernels.ExpSquaredKernel(1.0) + 0.5 * kernels.ExpSine2Kernel(gamma=1.0, log_period=np.log(3.0))
        return k, x, yerr, y, dict(min_size=100, tol=1e-10)
    raise ValueError(cfg)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tol", type=float, default=None)
    ap.add_argument("--exhaust", default="dense")
    a = ap.parse_args()
    k, x, yerr, y, kw = make(a.cfg, a.n)
    if a.tol is not None:
        kw["tol"] = a.tol
    s = HODLRSolver()
    for rep in range(a.reps):
        t0 = time.perf_counter()
        s.compute(k, x, yerr, seed=42, exhaust=a.exhaust, **kw)
        t1 = time.perf_counter()
        d = s.dot_solve(y)
        t2 = time.perf_counter()
        tm = s.timing()
        print(json.dumps({"cfg": a.cfg, "n": a.n, "rep": rep, "compute_wall_ms": (t1 - t0) * 1e3,
                          "dot_solve_wall_ms": (t2 - t1) * 1e3, **{k_: round(v, 3) for k_, v in tm.items()}}))
    by = defaultdict(list)
    for nd in s.nodes():
        if not nd["is_leaf"]:
            by[nd["depth"]].append(nd)
    for dpt in sorted(by):
        r = by[dpt]
        print(" depth", dpt, "nodes", len(r), "half", r[0]["half"], "ranks", sorted(set(n_["rank"] for n_ in r))[:8],
              "fallbacks", sum(n_["dense_fallback"] for n_ in r), "max draws", max(n_["rng_draws"] for n_ in r))
    print("logdet", s.log_determinant, "dot", d, "work", s.work())
