#!/bin/bash
# round 2, job 21: training-pair synthesis on the device against the oracle; throughput of the synthesis
set -x
mkdir -p gpurun_out
timeout 600 python -u -m pytest -x -v --timeout 300 tests/test_pairs_gpu.py > gpurun_out/r2_21_pytest_pairs.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert|differ" gpurun_out/r2_21_pytest_pairs.log | tail -12
timeout 300 python tools/pairs_bench.py > gpurun_out/r2_21_pairs_bench.log 2>&1
cat gpurun_out/r2_21_pairs_bench.log
