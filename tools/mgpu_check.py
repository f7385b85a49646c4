# -*- coding: utf-8 -*-
"""Multi-GPU parity probe (run under torchrun): sharded HODLR (sub-tree per rank + one all-gather) must reproduce the
single-GPU factorisation — same per-node RNG streams, so log-det / solve agree to rounding."""
import os, sys, json
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from george_b200 import kernels, _lib
from george_b200.parallel import ShardedHODLRSolver
from george_b200.solvers._hodlr import HODLRSolver

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
_lib.check(_lib.load().bgp_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, kernel, n, ms, exhaust in [
    ("expsq", 1.0 * kernels.ExpSquaredKernel(1.0), 20000, 100, "dense"),
    ("m32", 1.0 * kernels.Matern32Kernel(1.0), 16384, 256, "lowrank"),
    ("odd", 1.0 * kernels.ExpSquaredKernel(1.0), 8191, 100, "dense"),
    ("m32dense", 1.0 * kernels.Matern32Kernel(1.0), 6000, 100, "dense"),   # big-rank (blocked LU) top levels
    ("cfg5", 1.0 * kernels.ExpSquaredKernel(1.0) + 0.5 * kernels.ExpSine2Kernel(gamma=1.0, log_period=np.log(3.0)), 32768, 100, "lowrank"),
    ("m32big", 1.0 * kernels.Matern32Kernel(1.0), 262144, 256, "lowrank"),
]:
    rng = np.random.default_rng(1234)
    x = np.sort(rng.uniform(0, 10 * n / 1000, n)); yerr = 0.1 * np.ones(n); y = np.sin(x) + 0.1 * rng.normal(size=n)
    sh = ShardedHODLRSolver(kernel, min_size=ms, tol=1e-10, seed=42, exhaust=exhaust)
    sh.compute(x[:, None], yerr)
    ld, ds = sh.log_determinant, sh.dot_solve(y)
    a = sh.apply_inverse(y)[:, 0]
    if rank == 0:
        s = HODLRSolver(); s.compute(kernel, x[:, None], yerr, min_size=ms, tol=1e-10, seed=42, exhaust=exhaust)
        ld1, ds1 = s.log_determinant, s.dot_solve(y)
        a1 = s.apply_inverse(y)[:, 0]
        good = abs(ld - ld1) <= 1e-10 * abs(ld1) and abs(ds - ds1) <= 1e-9 * abs(ds1) and np.linalg.norm(a - a1) <= 1e-9 * np.linalg.norm(a1)
        ok = ok and good
        print(json.dumps({"case": name, "world": world, "logdet_sharded": ld, "logdet_single": ld1, "dot_sharded": ds, "dot_single": ds1,
                          "solve_relerr": float(np.linalg.norm(a - a1) / np.linalg.norm(a1)), "ok": bool(good)}))
dist.barrier()
if rank == 0:
    print("MGPU_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()
