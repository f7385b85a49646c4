#!/bin/bash
# 8-GPU visit: BASELINE.json configs[4] (ExpSquared+ExpSine2, N = 2^20 over 8 GPUs) and the headline workload at 8 GPUs.
NG=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
( time timeout 500 $TR bench.py --gpus $NG --steps 2 --warmup 3 --workload cfg5 ) > gpurun_out/m${NG}_bench_cfg5.log 2>&1; grep '^{"metric"' gpurun_out/m${NG}_bench_cfg5.log > gpurun_out/m${NG}_bench_cfg5.json; cut -c1-500 gpurun_out/m${NG}_bench_cfg5.json; tail -4 gpurun_out/m${NG}_bench_cfg5.log | cut -c1-300
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8
( time timeout 200 $TR bench.py --gpus $NG --steps 10 --warmup 3 ) > gpurun_out/m${NG}_bench_cfg3.log 2>&1; grep '^{"metric"' gpurun_out/m${NG}_bench_cfg3.log > gpurun_out/m${NG}_bench_cfg3.json; cut -c1-420 gpurun_out/m${NG}_bench_cfg3.json; tail -2 gpurun_out/m${NG}_bench_cfg3.log | cut -c1-200
