#!/bin/bash
# Round-2 GPU visit E: full parity suite (incl. the generated user kernels), a2_eval register-allocation A/B, bench line.
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -q --maxfail=15 ) > gpurun_out/e_tests.log 2>&1; tail -6 gpurun_out/e_tests.log
for m in 2 3; do for w in cfg3 cfg2; do BGP_EVAL_MINB=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --workload $w 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print('minb', $m, '$w', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['kernel_ms']['a2_eval'],3), round(d['roofline']['frac'],3))
"; done; done
