#!/bin/bash
# Third GPU-box visit: full GPU suite on the three-level delayed-update Cholesky / rsqrt potf2 / prefetching forward
# solve, config-4 timing with block-size variants, dense launch list, default bench line.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q --durations=3 ) > gpurun_out/c_tests.log 2>&1
tail -8 gpurun_out/c_tests.log
timeout 100 python tools/dense_bench.py --n 32768 --reps 3 > gpurun_out/c_dense_cfg4.txt 2>&1
echo "BGP_DENSE_OB=2048" >> gpurun_out/c_dense_cfg4.txt; BGP_DENSE_OB=2048 timeout 100 python tools/dense_bench.py --n 32768 --reps 2 >> gpurun_out/c_dense_cfg4.txt 2>&1
echo "BGP_DENSE_MB=512" >> gpurun_out/c_dense_cfg4.txt; BGP_DENSE_MB=512 timeout 100 python tools/dense_bench.py --n 32768 --reps 2 >> gpurun_out/c_dense_cfg4.txt 2>&1
echo "BGP_DENSE_MB=1024 (two levels, as before)" >> gpurun_out/c_dense_cfg4.txt; BGP_DENSE_MB=1024 timeout 100 python tools/dense_bench.py --n 32768 --reps 2 >> gpurun_out/c_dense_cfg4.txt 2>&1
cat gpurun_out/c_dense_cfg4.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c_launches_dense.csv \
  python tools/dense_bench.py --n 8192 --reps 1 > gpurun_out/c_launches_dense.log 2>&1
( time timeout 240 python bench.py ) > gpurun_out/c_bench.log 2>&1
tail -4 gpurun_out/c_bench.log | cut -c1-300
