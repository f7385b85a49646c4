#!/bin/bash
# 8-GPU visit: BASELINE.json configs[4] (ExpSquared+ExpSine2, N = 2^20 over 8 GPUs).
NG=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
( time timeout 700 $TR bench.py --gpus $NG --steps 2 --warmup 3 --workload cfg5 ) > gpurun_out/m${NG}_bench_cfg5.log 2>&1; grep '^{"metric"' gpurun_out/m${NG}_bench_cfg5.log > gpurun_out/m${NG}_bench_cfg5.json; cut -c1-900 gpurun_out/m${NG}_bench_cfg5.json; grep "\[bench\]" gpurun_out/m${NG}_bench_cfg5.log; tail -4 gpurun_out/m${NG}_bench_cfg5.log | cut -c1-300
