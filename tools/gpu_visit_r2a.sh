#!/bin/bash
# Round-2 GPU visit A: parity suite, the bench line, first look at the high-rank workload, launch lists + one full capture.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=15 --durations=8 ) > gpurun_out/a_tests.log 2>&1; tail -25 gpurun_out/a_tests.log
( time timeout 400 python bench.py --steps 10 ) > gpurun_out/a_bench.log 2>&1; grep '^{"metric"' gpurun_out/a_bench.log > gpurun_out/a_bench.json; cut -c1-600 gpurun_out/a_bench.json; tail -3 gpurun_out/a_bench.log | cut -c1-300
( time timeout 120 python tools/profile_step.py --workload cfg5 --n 65536 --steps 2 ) > gpurun_out/a_cfg5_n65536.log 2>&1; tail -4 gpurun_out/a_cfg5_n65536.log
( time timeout 200 python tools/profile_step.py --workload cfg5 --n 131072 --steps 2 ) > gpurun_out/a_cfg5_n131072.log 2>&1; tail -4 gpurun_out/a_cfg5_n131072.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/a_launches_cfg3.csv \
  python tools/profile_step.py --steps 2 > gpurun_out/a_launches_cfg3.log 2>&1
python tools/ncu_summary.py gpurun_out/a_launches_cfg3.csv > gpurun_out/a_launches_cfg3_summary.txt; head -20 gpurun_out/a_launches_cfg3_summary.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:a2_eval_kernel -s 44 -c 3 -f -o gpurun_out/a_prof_a2_eval \
  python tools/profile_step.py --steps 2 > gpurun_out/a_prof_a2_eval.log 2>&1
ls -la gpurun_out | tail -12
