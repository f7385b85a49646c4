# -*- coding: utf-8 -*-
"""N-sweep of gp.compute() + gp.log_likelihood() (N in {2^16 .. 2^20}, 1-D and a 3-D variant) on one GPU.

    python tools/sweep_n.py [--kernels expsq m32 cfg5 m52_3d] [--log2n 16 17 18 19 20] [--reps 3] > gpurun_out/sweep.jsonl

One JSON line per (kernel, N): best-of-reps wall time of compute + dot_solve through the C ABI with host inputs, the
device-event phase times, ranks and the algorithmic work counters of `bgp_hodlr_last_work`.  Not the judged benchmark
(that is bench.py); meant for `profiles/sweep_*.jsonl`.  Each case is bounded by --budget-s seconds.
"""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from george_b200 import kernels  # noqa: E402
from george_b200.solvers._hodlr import HODLRSolver  # noqa: E402

CASES = {
    # name: (kernel factory, ndim, min_size, exhaust)
    "expsq": (lambda: 1.0 * kernels.ExpSquaredKernel(1.0), 1, 100, "lowrank"),
    "m32": (lambda: 1.0 * kernels.Matern32Kernel(1.0), 1, 256, "lowrank"),
    "cfg5": (lambda: 1.0 * kernels.ExpSquaredKernel(1.0)
             + 0.5 * kernels.ExpSine2Kernel(gamma=1.0, log_period=np.log(3.0)), 1, 100, "lowrank"),
    "m52_3d": (lambda: 1.0 * kernels.Matern52Kernel(0.5, ndim=3), 3, 100, "lowrank"),
}


def data(n, ndim):
    rng = np.random.default_rng(1234)
    if ndim == 1:
        x = np.sort(rng.uniform(0, 10 * n / 1000, n))[:, None]
        y = np.sin(x[:, 0]) + 0.1 * rng.normal(size=n)
    else:
        x = rng.uniform(0, (n / 4096.0) ** (1.0 / 3.0), (n, ndim))  # constant density in 3-D
        x = x[np.argsort(x[:, 0])]
        y = np.sin(x.sum(axis=1)) + 0.1 * rng.normal(size=n)
    return x, 0.1 * np.ones(n), y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernels", nargs="+", default=["expsq", "m32"])
    ap.add_argument("--log2n", nargs="+", type=int, default=[16, 17, 18, 19, 20])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--budget-s", type=float, default=60.0)
    a = ap.parse_args()
    for name in a.kernels:
        mk, ndim, min_size, exhaust = CASES[name]
        for l2 in a.log2n:
            n = 1 << l2
            x, yerr, y = data(n, ndim)
            s = HODLRSolver()
            best, t_case, rec = None, time.perf_counter(), None
            for rep in range(a.reps):
                t0 = time.perf_counter()
                try:
                    s.compute(mk(), x, yerr, min_size=min_size, tol=a.tol, seed=42, exhaust=exhaust)
                    q = s.dot_solve(y)
                except Exception as exc:  # capacity / memory: report and move on
                    rec = {"kernel": name, "n": n, "error": repr(exc)}
                    break
                dt = time.perf_counter() - t0
                if best is None or dt < best:
                    best = dt
                    nodes = s.nodes()
                    rec = {"kernel": name, "ndim": ndim, "n": n, "tol": a.tol, "min_size": min_size, "exhaust": exhaust,
                           "seconds": dt, "points_per_s": n / dt,
                           "log_likelihood": -0.5 * (n * np.log(2 * np.pi) + s.log_determinant) - 0.5 * q,
                           "phases_ms": s.timing(), "work": s.work(),
                           "max_rank": max(nd["rank"] for nd in nodes),
                           "exhausted_nodes": sum(nd["dense_fallback"] for nd in nodes)}
                if time.perf_counter() - t_case > a.budget_s:
                    break
            print(json.dumps(rec), flush=True)
            del s


if __name__ == "__main__":
    main()
