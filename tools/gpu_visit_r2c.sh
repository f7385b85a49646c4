#!/bin/bash
# Round-2 GPU visit C: parity suite, bench line, N-sweep, launch list + captures of the final kernels.
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -q --maxfail=15 --durations=5 ) > gpurun_out/c_tests.log 2>&1; tail -12 gpurun_out/c_tests.log
( time timeout 400 python bench.py --steps 20 --warmup 5 ) > gpurun_out/c_bench.log 2>&1; grep '^{"metric"' gpurun_out/c_bench.log > gpurun_out/c_bench.json; cut -c1-330 gpurun_out/c_bench.json; tail -3 gpurun_out/c_bench.log | cut -c1-300
( timeout 150 python tools/sweep_n.py --kernels expsq m32 --log2n 16 17 18 19 20 --reps 2 --budget-s 20 > gpurun_out/c_sweep.jsonl ) 2> gpurun_out/c_sweep.err
( timeout 150 python tools/sweep_n.py --kernels cfg5 --log2n 14 15 16 17 18 --reps 1 --budget-s 30 >> gpurun_out/c_sweep.jsonl ) 2>> gpurun_out/c_sweep.err
( timeout 120 python tools/sweep_n.py --kernels m52_3d --log2n 13 14 15 16 --reps 1 --budget-s 30 >> gpurun_out/c_sweep.jsonl ) 2>> gpurun_out/c_sweep.err
cut -c1-200 gpurun_out/c_sweep.jsonl; tail -3 gpurun_out/c_sweep.err
BGP_NO_GRAPH=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/c_launches_cfg3.csv \
  python tools/profile_step.py --steps 2 > gpurun_out/c_launches_cfg3.log 2>&1
python tools/ncu_summary.py gpurun_out/c_launches_cfg3.csv > gpurun_out/c_launches_cfg3_summary.txt; head -22 gpurun_out/c_launches_cfg3_summary.txt
BGP_NO_GRAPH=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:a2_eval_kernel -s 46 -c 2 -f -o gpurun_out/c_prof_a2_eval \
  python tools/profile_step.py --steps 2 > gpurun_out/c_prof_a2_eval.log 2>&1
ls -la gpurun_out | tail -6
