#!/bin/bash
# round 2, job 50: warp-specialised fused block (E under D, 16 x 4 tiles), first run
set -x
mkdir -p gpurun_out
timeout 120 python -u -m pytest -x -q --timeout 60 tests/test_kernels_gpu.py -k "fsrgan_block" > gpurun_out/r2_50_pytest_new.log 2>&1; tail -8 gpurun_out/r2_50_pytest_new.log | cut -c1-300
timeout 120 python tools/fsrgan_block_timeline.py 2>&1 | head -1
DG_FSRGAN_BLOCK_WS=0 timeout 120 python tools/fsrgan_block_timeline.py 2>&1 | head -1
