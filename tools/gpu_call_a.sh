#!/bin/bash
# One GPU-box visit: full GPU test-suite, dense-path tuning sweep, launch list + ncu --set full captures of the dense path.
# usage (from the repo root, through gpurun): bash tools/gpu_call_a.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt 2>&1
( time timeout 420 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/a_tests.log 2>&1
tail -25 gpurun_out/a_tests.log
for ob in 256 512 1024; do
  echo "OB=$ob" >> gpurun_out/a_dense_cfg4.txt
  BGP_DENSE_OB=$ob timeout 120 python tools/dense_bench.py --n 32768 --reps 2 >> gpurun_out/a_dense_cfg4.txt 2>&1
done
cat gpurun_out/a_dense_cfg4.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/a_launches_dense.csv \
  python tools/dense_bench.py --n 8192 --reps 1 > gpurun_out/a_launches_dense.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:kmat_symmetric_kernel -c 1 -f -o gpurun_out/a_prof_kmat \
  python tools/dense_bench.py --n 16384 --reps 1 > gpurun_out/a_prof_kmat.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_kernel -s 80 -c 8 -f -o gpurun_out/a_prof_gemm \
  python tools/dense_bench.py --n 16384 --reps 1 > gpurun_out/a_prof_gemm.log 2>&1
ls -la gpurun_out
( time timeout 240 python bench.py --steps 3 --warmup 3 ) > gpurun_out/a_bench.log 2>&1
tail -3 gpurun_out/a_bench.log
