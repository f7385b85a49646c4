#!/bin/bash
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -q --maxfail=15 ) > gpurun_out/d_tests.log 2>&1; tail -5 gpurun_out/d_tests.log
( timeout 150 python tools/sweep_n.py --kernels expsq m32 --log2n 16 17 18 19 20 --reps 2 --budget-s 20 > gpurun_out/d_sweep.jsonl ) 2> gpurun_out/d_sweep.err
( timeout 100 python tools/sweep_n.py --kernels cfg5 --log2n 14 15 16 17 18 --reps 1 --budget-s 30 >> gpurun_out/d_sweep.jsonl ) 2>> gpurun_out/d_sweep.err
( timeout 100 python tools/sweep_n.py --kernels m52_3d --log2n 13 14 15 --reps 1 --budget-s 20 --no-cpu >> gpurun_out/d_sweep.jsonl ) 2>> gpurun_out/d_sweep.err
python - <<'PY'
import json
for l in open("gpurun_out/d_sweep.jsonl"):
    d = json.loads(l)
    if "error" in d: print(d["kernel"], d["n"], "ERROR", d["error"][:120]); continue
    c = d.get("cpu", {})
    print(d["kernel"], d["n"], "%.1f ms" % (1e3 * d["seconds"]), "%.3g pts/s" % d["points_per_s"], "maxrank", max(d["max_rank_by_level"]),
          "cpu_s", c.get("seconds", c.get("seconds_EXTRAPOLATED")), "relerr", c.get("rel_err_gpu_vs_cpu"))
PY
tail -2 gpurun_out/d_sweep.err
( time timeout 300 python bench.py --steps 20 --warmup 5 ) > gpurun_out/d_bench.log 2>&1; grep '^{"metric"' gpurun_out/d_bench.log > gpurun_out/d_bench.json; cut -c1-330 gpurun_out/d_bench.json
