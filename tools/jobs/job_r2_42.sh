#!/bin/bash
# round 2, job 42: fused Fast-SRGAN block v3 (fp16 intermediates / HFMA2 depthwise stage, expand bias on the tensor pipe)
set -x
mkdir -p gpurun_out
timeout 300 python -u -m pytest -x -q --timeout 120 tests/test_kernels_gpu.py -k "fsrgan_block" > gpurun_out/r2_42_pytest_new.log 2>&1; tail -5 gpurun_out/r2_42_pytest_new.log | cut -c1-300
timeout 600 python -u -m pytest -x -q --timeout 600 tests/test_infer_gpu.py -k "fsrgan" > gpurun_out/r2_42_pytest_infer.log 2>&1; tail -3 gpurun_out/r2_42_pytest_infer.log | cut -c1-200
timeout 300 python tools/fsrgan_block_timeline.py > gpurun_out/r2_42_fb_timeline.log 2>&1; head -4 gpurun_out/r2_42_fb_timeline.log | cut -c1-250; tail -1 gpurun_out/r2_42_fb_timeline.log
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_42_bench_infer_fsrgan.log 2>&1
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r2_42_bench_*.log
