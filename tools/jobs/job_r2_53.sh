#!/bin/bash
# round 2, job 53: both block kernels through the parametrised test; pix2pix per-call conv timings and per-entry-point step profile at HEAD
set -x
mkdir -p gpurun_out
timeout 120 python -u -m pytest -x -q --timeout 60 tests/test_kernels_gpu.py -k "fsrgan_block" > gpurun_out/r2_53_pytest_new.log 2>&1; tail -3 gpurun_out/r2_53_pytest_new.log | cut -c1-300
timeout 300 python tools/conv_calls.py --model pix2pix --top 130 > gpurun_out/r2_53_calls_pix2pix.log 2>&1; head -4 gpurun_out/r2_53_calls_pix2pix.log
timeout 300 python tools/step_profile.py --model pix2pix --batch 32 --crop 256 > gpurun_out/r2_53_step_pix2pix.log 2>&1; head -24 gpurun_out/r2_53_step_pix2pix.log
