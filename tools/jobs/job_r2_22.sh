#!/bin/bash
# round 2, job 22: image summaries of the training loop
set -x
mkdir -p gpurun_out
timeout 900 python -u -m pytest -x -v --timeout 300 tests/test_summaries_gpu.py tests/test_train_loop_gpu.py > gpurun_out/r2_22_pytest.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_22_pytest.log | tail -20
tail -30 gpurun_out/r2_22_pytest.log | cut -c1-220
