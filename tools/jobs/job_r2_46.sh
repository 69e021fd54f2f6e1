#!/bin/bash
# round 2, job 46: pix2pix U-Net concat in place (no copies in either direction)
set -x
mkdir -p gpurun_out
timeout 600 python -u -m pytest -x -q -s --timeout 300 tests/test_models_gpu.py -k "pix2pix" > gpurun_out/r2_46_pytest_pix.log 2>&1; grep -E "worst|passed|failed|Error|assert" gpurun_out/r2_46_pytest_pix.log | tail -8 | cut -c1-400
timeout 300 python bench.py --workload pix2pix_c4 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_46_bench_pix2pix.log 2>&1
DG_INPLACE_CONCAT=0 timeout 300 python bench.py --workload pix2pix_c4 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_46_bench_pix2pix_off.log 2>&1
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r2_46_bench_*.log
tail -3 gpurun_out/r2_46_bench_pix2pix.log | cut -c1-300
