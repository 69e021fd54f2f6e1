#!/bin/bash
# round 2, job 10: yb through the staging buffer (TMA, in place) in the BatchNorm-backward dgrad epilogue
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 100 --timeout-method thread"
timeout 300 $PYT tests/test_kernels_gpu.py -k "dgrad_fused" > gpurun_out/r2_10_pytest_new.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_10_pytest_new.log | head -30
timeout 200 python tools/dgrad_fused_probe.py > gpurun_out/r2_10_probe.log 2>&1
cat gpurun_out/r2_10_probe.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_10_bench.log 2>&1
DG_DGRAD_BN_BWD=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_10_bench_mode1.log 2>&1
grep -h '"value"' gpurun_out/r2_10_bench*.log | cut -c1-200
timeout 900 $PYT tests -m gpu > gpurun_out/r2_10_pytest.log 2>&1
tail -5 gpurun_out/r2_10_pytest.log
