#!/bin/bash
# Round-2 final visit (2 GPUs): smoke, parity suite, sharded == single probe, bench at 1 and 2 GPUs.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
( timeout 120 python __graft_entry__.py smoke ) 2>&1 | tail -1
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_tests.log 2>&1; tail -4 gpurun_out/g_tests.log
( timeout 200 $TR tools/mgpu_check.py ) > gpurun_out/g_mgpu_check.log 2>&1; grep -E "MGPU_CHECK|\"ok\": false|Error" gpurun_out/g_mgpu_check.log | cut -c1-200
( timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/g_bench_2gpu.log 2>&1; grep '^{"metric"' gpurun_out/g_bench_2gpu.log > gpurun_out/g_bench_2gpu.json; cut -c1-260 gpurun_out/g_bench_2gpu.json; tail -2 gpurun_out/g_bench_2gpu.log | cut -c1-200
( timeout 200 python bench.py --steps 20 --warmup 5 --no-secondary ) > gpurun_out/g_bench_1gpu.json 2> gpurun_out/g_bench_1gpu.err; cut -c1-260 gpurun_out/g_bench_1gpu.json
