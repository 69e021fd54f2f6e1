#!/bin/bash
# round 2, job 48: control and epilogue warps sleep between barrier polls
set -x
mkdir -p gpurun_out
timeout 300 python tools/fsrgan_block_timeline.py > gpurun_out/r2_48_fb_timeline.log 2>&1; head -1 gpurun_out/r2_48_fb_timeline.log; tail -10 gpurun_out/r2_48_fb_timeline.log | cut -c1-250
timeout 300 python -u -m pytest -x -q --timeout 120 tests/test_kernels_gpu.py -k "fsrgan_block" > gpurun_out/r2_48_pytest_new.log 2>&1; tail -2 gpurun_out/r2_48_pytest_new.log | cut -c1-300
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_48_bench_infer_fsrgan.log 2>&1
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r2_48_bench_*.log
