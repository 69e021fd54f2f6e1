#!/bin/bash
# round 2, job 43: FMA issue-rate probe; fused block with the depthwise window loaded one step ahead
set -x
mkdir -p gpurun_out
timeout 120 ./probes/fma_rate_probe > gpurun_out/r2_43_fma_rate.log 2>&1; cat gpurun_out/r2_43_fma_rate.log
timeout 300 python -u -m pytest -x -q --timeout 120 tests/test_kernels_gpu.py -k "fsrgan_block" > gpurun_out/r2_43_pytest_new.log 2>&1; tail -3 gpurun_out/r2_43_pytest_new.log | cut -c1-300
timeout 300 python tools/fsrgan_block_timeline.py > gpurun_out/r2_43_fb_timeline.log 2>&1; head -4 gpurun_out/r2_43_fb_timeline.log | cut -c1-250; tail -1 gpurun_out/r2_43_fb_timeline.log
