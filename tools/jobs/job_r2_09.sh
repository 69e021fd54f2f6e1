#!/bin/bash
# round 2, job 9: BatchNorm-backward sums in the dgrad epilogue, mma-fragment accumulator layout (16x256b)
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 100 --timeout-method thread"
timeout 300 $PYT tests/test_kernels_gpu.py -k "dgrad_fused" > gpurun_out/r2_09_pytest_new.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_09_pytest_new.log | head -30
timeout 200 python tools/dgrad_fused_probe.py > gpurun_out/r2_09_probe.log 2>&1
cat gpurun_out/r2_09_probe.log
DG_DGRAD_BN_BWD=2 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_09_bench_mode2.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_09_bench_mode1.log 2>&1
grep -h '"value"' gpurun_out/r2_09_bench*.log | cut -c1-200
DG_DGRAD_BN_BWD=2 timeout 600 $PYT tests/test_srgan_gpu.py tests/test_models_gpu.py > gpurun_out/r2_09_pytest_models_mode2.log 2>&1
tail -5 gpurun_out/r2_09_pytest_models_mode2.log
