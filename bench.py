#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
bench.py — gp.compute() + gp.log_likelihood() throughput (N-points/s) for the HODLR path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg5]

One "step" = one gp.compute(x, yerr) + one gp.log_likelihood(y) on a fixed synthetic data set:
    x = sort(U(0, 10*N/1000)) (rng 1234), yerr = 0.1, y = sin(x) + 0.1*N(0,1).
Default workload = BASELINE.json's metric config: Matern32Kernel 1-D, N = 2^18 per GPU, HODLRSolver(min_size=256,
tol=1e-10, seed=42).  With --gpus N (launched by torchrun, one rank per GPU) the SAME GP (N = 2^18, strong scaling, as
the metric is quoted) is sharded by top-level sub-tree: one all-gather of the top-level factor rows, plus a MAX
all-reduce per ACA iteration while the scans of the nodes above the cut are split across the ranks.

JSON keys follow the driver contract; extra: roofline{}, cpu_baseline{}, clocks{}, e2e{}, gpu_launches.
The "reference" arm times the CPU oracle port (oracle/: Eigen-free restatement of george's hodlr.h; the reference's
own _hodlr extension needs Eigen, which is absent) on a bounded sample of the same workload, single thread.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_TENSOR_PEAK_TFLOPS = 37.2   # DMMA.8x8x4 issue-bound peak measured on this pool (profiles/fp64_peaks_r01.txt)
FP64_DFMA_PEAK_TFLOPS = 34.1     # DFMA issue-bound peak, same measurement
# FP64 operations the specialised evaluator executes per covariance entry, counted in the SASS of
# a2_eval_kernel<shape> (cuobjdump; DFMA = 2 flop, DMUL/DADD = 1; exp and sqrt are software sequences on this pipe)
EVAL_FLOPS = {"cfg3": 56.0, "cfg2": 40.0, "cfg5": 120.0}
HBM_FALLBACK_GBS = 6650.0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


WORKLOADS = {
    # name: (kernel factory, solver kwargs, label)
    "cfg3": dict(label="Matern32Kernel 1D N=262144 HODLRSolver leaf=256", n=262144, min_size=256, tol=1e-10,
                 cpu_sample_n=4096, cpu_sample_n_many_steps=2048),
    "cfg2": dict(label="ExpSquaredKernel 1D N=65536 HODLRSolver tol=1e-10", n=65536, min_size=100, tol=1e-10,
                 cpu_sample_n=16384),
    "cfg5": dict(label="ExpSquared+ExpSine2 sum kernel 1D N=131072/GPU HODLRSolver tol=1e-10", n=131072, min_size=100,
                 tol=1e-10, cpu_sample_n=8192),
}


def make_kernel(name):
    from george_b200 import kernels
    if name == "cfg3":
        return 1.0 * kernels.Matern32Kernel(1.0)
    if name == "cfg2":
        return 1.0 * kernels.ExpSquaredKernel(1.0)
    return 1.0 * kernels.ExpSquaredKernel(1.0) + 0.5 * kernels.ExpSine2Kernel(gamma=1.0, log_period=np.log(3.0))


def make_data(n):
    rng = np.random.default_rng(1234)
    x = np.sort(rng.uniform(0, 10 * n / 1000, n))
    yerr = 0.1 * np.ones(n)
    y = np.sin(x) + 0.1 * rng.normal(size=n)
    return x, yerr, y


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(name, steps=1, n=None):
    """The oracle port on a bounded sample of the workload (single thread, as the reference is).

    The reference algorithm stores a block DENSELY when its ACA runs out of rows (hodlr.h:161-176), which Matern-3/2
    on sorted 1-D inputs triggers on most nodes: its cost grows like N^3 there (this container: N=2048 1.6 s,
    4096 23 s, 8192 199 s), so the sample size decides the points/s this leg reports; it is stated in `sample`."""
    import oracle
    from george_b200._spec import flatten
    wl = WORKLOADS[name]
    n = n or wl["cpu_sample_n"]
    x, yerr, y = make_data(n)
    spec = flatten(make_kernel(name))
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        h = oracle.HODLR(spec, x, yerr, min_size=wl["min_size"], tol=wl["tol"], seed=42, rng_mode=1)
        ll = -0.5 * (n * np.log(2 * np.pi) + h.log_determinant) - 0.5 * h.dot_solve(y)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": n / best, "unit": "points/s", "cores": 1, "kind": "port", "seconds": best, "log_likelihood": ll,
            "sample": "same kernel/min_size/tol/seed at N={0} (reference rng order); the full N is not feasible on the "
                      "CPU path".format(n)}


def run_reference(args):
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded so that the whole --steps K --warmup W run ends within a few minutes (one warm-up pass is enough for a
    # single-threaded CPU code: there is nothing to compile or cache beyond the page faults of the first run)
    n = wl["cpu_sample_n"] if args.steps <= 8 else wl.get("cpu_sample_n_many_steps", wl["cpu_sample_n"])
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline(args.workload, n=n)
    times = []
    cb = None
    for _ in range(args.steps):
        cb = cpu_baseline(args.workload, n=n)
        times.append(cb["seconds"])
    value = n * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "gp.compute+log_likelihood N-points/sec", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "min_size": wl["min_size"], "tol": wl["tol"], "seed": 42,
                   "reference_impl": "oracle/ C++ restatement of george hodlr.h (Eigen absent, _hodlr not buildable)"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": 1, "kind": "port", "sample": cb["sample"]},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def pinned_array(lib, n):
    from george_b200 import _lib
    p = C.c_void_p()
    _lib.check(lib.bgp_host_alloc_pinned(C.byref(p), n * 8))
    buf = (C.c_double * n).from_address(p.value)
    return np.frombuffer(buf, dtype=np.float64), p


def run_ours(args):
    import torch
    from george_b200 import _lib, GP, HODLRSolver
    from george_b200._spec import flatten
    from george_b200.solvers._hodlr import HODLRSolver as Native

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    lib = _lib.load()
    _lib.check(lib.bgp_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = WORKLOADS[args.workload]
    n = wl["n"]  # strong scaling: the metric is quoted at a fixed N = 2^18; more GPUs share the same problem
    kernel = make_kernel(args.workload)
    x, yerr, y = make_data(n)
    spec = flatten(kernel)
    exhaust = args.exhaust
    solver_kw = dict(min_size=wl["min_size"], tol=wl["tol"], seed=42, exhaust=exhaust)

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    # ---------------- leg 1: device-resident inputs (value) ----------------
    ll_value = None
    if world == 1:
        native = Native()
        dx, dyerr, dy = C.c_void_p(), C.c_void_p(), C.c_void_p()
        for p, a in ((dx, x), (dyerr, yerr), (dy, y)):
            _lib.check(lib.bgp_dev_alloc(C.byref(p), a.nbytes))
            _lib.check(lib.bgp_dev_upload(p, _lib.ptr(a), a.nbytes))
        native.set_profiling(True)  # CUDA events around every a2_eval launch (the dominant kernel)
        opts = native._opts(wl["min_size"], wl["tol"], 42, "pernode", 0, 0, 1, exhaust)
        out = C.c_double()

        def step_value():
            _lib.check(lib.bgp_hodlr_compute_dev(native._ptr, C.byref(spec), dx, n, 1, dyerr, C.byref(opts)))
            ld = C.c_double()
            _lib.check(lib.bgp_hodlr_log_determinant(native._ptr, C.byref(ld)))
            _lib.check(lib.bgp_hodlr_dot_solve_dev(native._ptr, dy, C.byref(out)))
            return -0.5 * (n * np.log(2 * np.pi) + ld.value) - 0.5 * out.value
    else:
        from george_b200.parallel import ShardedHODLRSolver
        sharded = ShardedHODLRSolver(kernel, **solver_kw)

        def step_value():
            sharded.compute(x[:, None], yerr)
            return -0.5 * (n * np.log(2 * np.pi) + sharded.log_determinant) - 0.5 * sharded.dot_solve(y)

    for _ in range(max(args.warmup, 3)):
        ll_value = step_value()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.bgp_launch_count()
    t_steps, t_wall, leaf_ms, aca_ms, up_ms, solve_ms, aca_prof = [], [], [], [], [], [], []
    for _ in range(args.steps):
        flush_l2()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        ll_value = step_value()
        torch.cuda.synchronize()
        e1.record()
        e1.synchronize()
        t_wall.append(time.perf_counter() - t0)
        t_steps.append(1e-3 * e0.elapsed_time(e1))
        if world == 1:
            tm = native.timing()
            leaf_ms.append(tm["leaves_ms"]); aca_ms.append(tm["aca_ms"]); up_ms.append(tm["upsweep_ms"]); solve_ms.append(tm["solve_ms"])
            aca_prof.append(native.aca_profile())
    barrier()
    launches = lib.bgp_launch_count() - launches0
    total = sum(t_steps)
    if dist is not None:
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())

    # ---------------- leg 2: end to end through the public GP API with pinned HOST buffers (e2e) ----------------
    hx, px = pinned_array(lib, n)
    hyerr, pyerr = pinned_array(lib, n)
    hy, py = pinned_array(lib, n)
    hx[:] = x; hyerr[:] = yerr; hy[:] = y
    if world == 1:
        gp = GP(kernel, solver=HODLRSolver, **solver_kw)
    else:
        from george_b200.parallel import ShardedHODLRSolver
        gp = GP(kernel, solver=ShardedHODLRSolver, **solver_kw)

    def step_e2e():
        gp.compute(hx, hyerr)
        return gp.log_likelihood(hy)

    for _ in range(max(args.warmup, 3)):
        ll_e2e = step_e2e()
    barrier()
    t_e2e = []
    for _ in range(args.steps):
        flush_l2()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ll_e2e = step_e2e()
        torch.cuda.synchronize()
        e1.record()
        e1.synchronize()
        t_e2e.append(1e-3 * e0.elapsed_time(e1))
    barrier()
    total_e2e = sum(t_e2e)
    if dist is not None:
        t = torch.tensor([total_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_e2e = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peaks_kind = measured_peaks()
    value = n * args.steps / total
    e2e_value = n * args.steps / total_e2e
    line = {
        "metric": "gp.compute+log_likelihood N-points/sec", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * total / args.steps,
        "ms_per_step_wall": 1e3 * sum(t_wall) / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"] + (" sharded by top-level sub-tree over {0} GPUs".format(world) if world > 1 else ""),
                   "N": n, "min_size": wl["min_size"], "tol": wl["tol"], "seed": 42, "rng_mode": "pernode",
                   "exhausted_rows": exhaust, "l2": "flushed between timed iterations (256 MB memset)",
                   "timing": "CUDA events bracketing each step (device idle on both sides: the events also cover the host "
                             "gaps of the lock-step loop), summed over the K steps, max over ranks; ms_per_step_wall = host clock"},
        "log_likelihood": ll_value, "log_likelihood_e2e": ll_e2e,
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": 3 * n * 8, "d2h_bytes_per_step": 16,
                "ms_per_step": 1e3 * total_e2e / args.steps, "api": "george_b200.GP.compute + GP.log_likelihood"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if world == 1:
        # dominant kernel of this workload (profiles/launches_*_summary.txt): a2_eval_kernel — residual rows of the ACA
        # candidates, FP64-pipe bound (software exp/sqrt + FMA updates; it writes only per-chunk maxima, so there is no
        # HBM roofline).  Duration: CUDA events around every launch on its stream, inside the timed region.
        work = native.work()
        ev_ms = statistics.mean(p["eval_ms"] for p in aca_prof)
        launches_eval = statistics.mean(p["eval_launches"] for p in aca_prof)
        evals = statistics.mean(p["evals"] for p in aca_prof)
        fmas = statistics.mean(p["update_fmas"] for p in aca_prof)
        flops = evals * EVAL_FLOPS[args.workload] + 2.0 * fmas
        achieved = flops / (ev_ms * 1e-3) * 1e-12
        traffic = None
        try:
            import csv
            with open(os.path.join(ROOT, "profiles", "prof_a2_eval_r01_v3_raw.csv")) as fh:
                rows = list(csv.reader(fh))
            hdr = rows[0]
            traffic = (float(rows[2][hdr.index("dram__bytes_read.sum")]) + float(rows[2][hdr.index("dram__bytes_write.sum")])) * 1e6
        except Exception:
            pass
        line["roofline"] = {"kernel": "a2_eval_kernel", "bound": "tensor", "pipe": "fp64 (DFMA/DMUL/DADD; no matrix contraction in this kernel)",
                            "achieved": achieved, "peak": FP64_DFMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_DFMA_PEAK_TFLOPS,
                            "traffic": traffic,
                            "peak_source": "FP64 DFMA issue peak measured with tools/fp64_peaks.cu on this pool, profiles/fp64_peaks_r01.txt "
                            "(MEASURED_PEAKS.json has no FP64 entry; its hbm_gbs = {0} [{1}]; DMMA peak {2})".format(
                                peaks.get("hbm_gbs"), peaks_kind, FP64_TENSOR_PEAK_TFLOPS),
                            "algorithmic_flops_per_launch": flops / launches_eval, "launch_ms": ev_ms / launches_eval,
                            "launches_per_step": launches_eval, "kernel_evals_per_step": evals,
                            "flops_per_eval": EVAL_FLOPS[args.workload], "share_of_step": ev_ms / (1e3 * total / args.steps)}
        line["phases_ms"] = {"leaves": statistics.mean(leaf_ms), "aca": statistics.mean(aca_ms),
                             "upsweep": statistics.mean(up_ms), "solve": statistics.mean(solve_ms)}
        line["work"] = work
        if not args.no_cpu:
            cb = cpu_baseline(args.workload)
            line["cpu_baseline"] = {k: v for k, v in cb.items() if k != "log_likelihood"}
            # same run, same inputs: the CUDA path in the reference's own mode (one mt19937 threaded through the
            # pre-order recursion, dense storage of exhausted blocks) against the oracle's number on the CPU sample
            ns = wl["cpu_sample_n"]
            try:
                xs, yerrs, ys = make_data(ns)
                chk = Native()
                t_same = None
                for _ in range(2):  # second pass: buffers and capacities are warm, as in the timed legs
                    t0 = time.perf_counter()
                    chk.compute(kernel, xs[:, None], yerrs, wl["min_size"], wl["tol"], 42, rng_mode="reference", exhaust="dense")
                    ll_gpu = -0.5 * (ns * np.log(2 * np.pi) + chk.log_determinant) - 0.5 * chk.dot_solve(ys)
                    t_same = time.perf_counter() - t0
                line["parity"] = {"n": ns, "mode": "rng_mode=reference, exhaust=dense (the reference algorithm)",
                                  "log_likelihood_gpu": ll_gpu, "log_likelihood_cpu": cb["log_likelihood"],
                                  "rel_err": abs(ll_gpu - cb["log_likelihood"]) / abs(cb["log_likelihood"]), "bar": 1e-6,
                                  # the SAME algorithm on the SAME sample, host inputs, wall clock: the like-for-like ratio
                                  "gpu_seconds_same_algorithm": t_same, "cpu_seconds": cb["seconds"],
                                  "speedup_same_algorithm_same_n": cb["seconds"] / t_same}
            except Exception as exc:  # the check must never cost the bench line
                line["parity"] = {"n": ns, "error": repr(exc)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--exhaust", default="lowrank", choices=["dense", "lowrank"])
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
