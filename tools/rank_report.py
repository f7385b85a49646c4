# -*- coding: utf-8 -*-
"""Per-level ranks / draws and the per-kernel times of the lock-step ACA loop for one workload (development probe).
usage: python tools/rank_report.py --workload cfg5 --n 262144"""
import argparse, sys, json, collections
import numpy as np
sys.path.insert(0, ".")
import bench
from george_b200.solvers._hodlr import HODLRSolver

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg5")
ap.add_argument("--n", type=int, default=262144)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
x, yerr, y = bench.make_data(a.n)
k = bench.make_kernel(a.workload)
s = HODLRSolver()
for rep in range(a.reps):
    s.set_profiling(rep == a.reps - 1)
    s.compute(k, x[:, None], yerr, min_size=wl["min_size"], tol=wl["tol"], seed=42, exhaust="lowrank")
    d = s.dot_solve(y)
    print(rep, s.log_determinant, d, s.timing())
lv = collections.defaultdict(list)
for nd in s.nodes():
    if not nd["is_leaf"]:
        lv[nd["depth"]].append((nd["rank"], nd["rng_draws"], nd["size"]))
for dpt in sorted(lv):
    r = [v[0] for v in lv[dpt]]
    print("level", dpt, "nodes", len(r), "size", lv[dpt][0][2], "rank min/median/max", min(r), int(np.median(r)), max(r),
          "draws max", max(v[1] for v in lv[dpt]))
print(json.dumps(s.aca_profile()))
print(json.dumps(s.work()))
