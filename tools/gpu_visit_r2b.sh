#!/bin/bash
# Round-2 GPU visit B: parity suite, bench line, cfg5, source-level capture of a2_decide.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=15 --durations=5 ) > gpurun_out/b_tests.log 2>&1; tail -12 gpurun_out/b_tests.log
( time timeout 400 python bench.py --steps 10 ) > gpurun_out/b_bench.log 2>&1; grep '^{"metric"' gpurun_out/b_bench.log > gpurun_out/b_bench.json; cut -c1-400 gpurun_out/b_bench.json; tail -3 gpurun_out/b_bench.log | cut -c1-300
( time timeout 120 python tools/profile_step.py --workload cfg5 --n 65536 --steps 3 ) > gpurun_out/b_cfg5_n65536.log 2>&1; tail -3 gpurun_out/b_cfg5_n65536.log
( time timeout 200 python tools/profile_step.py --workload cfg5 --n 131072 --steps 3 ) > gpurun_out/b_cfg5_n131072.log 2>&1; tail -3 gpurun_out/b_cfg5_n131072.log
BGP_NO_GRAPH=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/b_launches_cfg3.csv \
  python tools/profile_step.py --steps 2 > gpurun_out/b_launches_cfg3.log 2>&1
python tools/ncu_summary.py gpurun_out/b_launches_cfg3.csv > gpurun_out/b_launches_cfg3_summary.txt; head -20 gpurun_out/b_launches_cfg3_summary.txt
BGP_NO_GRAPH=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:a2_decide_kernel -s 68 -c 2 -f -o gpurun_out/b_prof_a2_decide \
  python tools/profile_step.py --steps 2 > gpurun_out/b_prof_a2_decide.log 2>&1
BGP_NO_GRAPH=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:a2_eval_kernel -s 46 -c 1 -f -o gpurun_out/b_prof_a2_eval \
  python tools/profile_step.py --steps 2 > gpurun_out/b_prof_a2_eval.log 2>&1
ls -la gpurun_out | tail -8
