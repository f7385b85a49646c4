#!/bin/bash
# Second GPU-box visit of the session: full GPU suite on the rewritten dense panel / solve / specialised kmat kernels,
# config-4 timing, launch list + ncu --set full of the new kmat build, default bench line.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/b_tests.log 2>&1
tail -14 gpurun_out/b_tests.log
timeout 120 python tools/dense_bench.py --n 32768 --reps 3 > gpurun_out/b_dense_cfg4.txt 2>&1
BGP_KMAT_GENERIC=1 timeout 120 python tools/dense_bench.py --n 32768 --reps 2 > gpurun_out/b_dense_cfg4_generic_kmat.txt 2>&1
cat gpurun_out/b_dense_cfg4.txt gpurun_out/b_dense_cfg4_generic_kmat.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/b_launches_dense.csv \
  python tools/dense_bench.py --n 8192 --reps 1 > gpurun_out/b_launches_dense.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:kmat_symmetric -c 1 -f -o gpurun_out/b_prof_kmat \
  python tools/dense_bench.py --n 16384 --reps 1 > gpurun_out/b_prof_kmat.log 2>&1
( time timeout 240 python bench.py ) > gpurun_out/b_bench.log 2>&1
tail -2 gpurun_out/b_bench.log | cut -c1-600
