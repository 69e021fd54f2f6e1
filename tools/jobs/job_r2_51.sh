#!/bin/bash
# round 2, job 51: warp-specialised fused block with the A operand double-buffered
set -x
mkdir -p gpurun_out
timeout 120 python -u -m pytest -x -q --timeout 60 tests/test_kernels_gpu.py -k "fsrgan_block" > gpurun_out/r2_51_pytest_new.log 2>&1; tail -4 gpurun_out/r2_51_pytest_new.log | cut -c1-300
timeout 120 python tools/fsrgan_block_timeline.py 2>&1 | head -1
timeout 600 python -u -m pytest -x -q --timeout 600 tests/test_infer_gpu.py -k fsrgan > gpurun_out/r2_51_pytest_infer.log 2>&1; tail -2 gpurun_out/r2_51_pytest_infer.log | cut -c1-200
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_51_bench_infer_fsrgan.log 2>&1
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r2_51_bench_*.log
