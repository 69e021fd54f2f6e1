#!/bin/bash
# round 2, job 29 (2 GPUs): which change broke the 2-rank bf16 gradient test
set -x
mkdir -p gpurun_out
R=$PWD



timeout 900 python -u -m pytest -x -v --timeout 600 tests/test_parallel_gpu.py > gpurun_out/r2_29_parallel_fixed.log 2>&1; grep -E "PASSED|FAILED|Error|worst" gpurun_out/r2_29_parallel_fixed.log | cut -c1-200
