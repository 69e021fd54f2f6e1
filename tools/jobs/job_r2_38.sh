#!/bin/bash
# round 2, job 38: Fast-SRGAN inverted-residual block as one launch (dg_fsrgan_block_infer), first run
set -x
mkdir -p gpurun_out
timeout 300 python -u -m pytest -x -q --timeout 120 tests/test_kernels_gpu.py -k "fsrgan_block" > gpurun_out/r2_38_pytest_new.log 2>&1; tail -15 gpurun_out/r2_38_pytest_new.log | cut -c1-300
