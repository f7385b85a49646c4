#!/bin/bash
# Multi-GPU visit (N = number of GPUs given to gpurun): sharded == single parity, then the bench at N GPUs.
NG=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
( time timeout 300 $TR tools/mgpu_check.py ) > gpurun_out/m${NG}_check.log 2>&1; grep -E "case|MGPU_CHECK|Error|error" gpurun_out/m${NG}_check.log | cut -c1-260 | tail -12
( time timeout 300 $TR bench.py --gpus $NG --steps 10 --warmup 3 ) > gpurun_out/m${NG}_bench_cfg3.log 2>&1; grep '^{"metric"' gpurun_out/m${NG}_bench_cfg3.log | cut -c1-420; tail -2 gpurun_out/m${NG}_bench_cfg3.log | cut -c1-200
( time timeout 400 $TR bench.py --gpus $NG --steps 5 --warmup 3 --workload cfg5 ) > gpurun_out/m${NG}_bench_cfg5.log 2>&1; grep '^{"metric"' gpurun_out/m${NG}_bench_cfg5.log | cut -c1-420; tail -2 gpurun_out/m${NG}_bench_cfg5.log | cut -c1-200
