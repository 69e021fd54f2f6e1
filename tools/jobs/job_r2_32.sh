#!/bin/bash
# round 2, job 32: loss terms in one launch; full GPU suite; C3 bench; wgrad x4 ncu capture
set -x
mkdir -p gpurun_out
timeout 2400 python -u -m pytest -x -q --timeout 600 --timeout-method thread tests -m gpu > gpurun_out/r2_32_pytest_all.log 2>&1
tail -4 gpurun_out/r2_32_pytest_all.log | cut -c1-250
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_32_bench_c3.log 2>&1
timeout 300 python bench.py --workload ae_c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_32_bench_ae.log 2>&1
timeout 300 python bench.py --workload fsrgan --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_32_bench_fsrgan.log 2>&1
grep -h '"value"' gpurun_out/r2_32_bench_*.log | cut -c1-200
grep -o '"kernel_nodes_per_step": [0-9]*' gpurun_out/r2_32_bench_*.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_wgrad_kernel|wgrad_reduce" -c 12 -o /tmp/r2_32_wgrad python tools/bench_conv.py --only body_wgrad,body_wgrad_x4 --iters 2 --graph 0 > gpurun_out/r2_32_ncu_wgrad.log 2>&1
ncu -i /tmp/r2_32_wgrad.ncu-rep --page raw --csv > gpurun_out/r2_32_wgrad_raw.csv 2>/dev/null
ls -la gpurun_out/r2_32_wgrad_raw.csv
