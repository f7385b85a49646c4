#!/bin/bash
# First GPU-box visit of the next round (one call, ~6 minutes of box time): everything the round's plan needs measured.
#   bash tools/gpu_next_visit.sh            (from the repo root, through gpurun --timeout 600)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/n_tests.log 2>&1; tail -6 gpurun_out/n_tests.log
( time timeout 200 python bench.py ) > gpurun_out/n_bench.log 2>&1; grep '^{"metric"' gpurun_out/n_bench.log | cut -c1-200
timeout 200 python tools/sweep_n.py --kernels expsq m32 --log2n 16 17 18 19 20 --reps 2 > gpurun_out/n_sweep.jsonl 2> gpurun_out/n_sweep.err
timeout 200 python tools/sweep_n.py --kernels m52_3d cfg5 --log2n 16 17 --reps 2 --budget-s 40 >> gpurun_out/n_sweep.jsonl 2>> gpurun_out/n_sweep.err
cut -c1-160 gpurun_out/n_sweep.jsonl
timeout 100 python tools/dense_bench.py --n 32768 --reps 3 > gpurun_out/n_dense_cfg4.txt 2>&1; cat gpurun_out/n_dense_cfg4.txt
# ncu: launch list of the bench workload + full captures of the dominant kernels
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/n_launches_cfg3.csv \
  python tools/profile_step.py --steps 2 > gpurun_out/n_launches_cfg3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:a2_eval_kernel -s 20 -c 2 -f -o gpurun_out/n_prof_a2_eval \
  python tools/profile_step.py --steps 1 > gpurun_out/n_prof_a2_eval.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_kernel -s 200 -c 12 -f -o gpurun_out/n_prof_gemm \
  python tools/dense_bench.py --n 16384 --reps 1 > gpurun_out/n_prof_gemm.log 2>&1
ls -la gpurun_out | tail -20
