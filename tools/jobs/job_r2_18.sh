#!/bin/bash
# round 2, job 18: TMA-fed depthwise 3x3 kernel
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 300 --timeout-method thread"
timeout 600 $PYT tests/test_kernels_gpu.py -k "depthwise" > gpurun_out/r2_18_pytest_new.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_18_pytest_new.log | tail -12
timeout 900 $PYT tests/test_models_gpu.py -k fsrgan tests/test_infer_gpu.py > gpurun_out/r2_18_pytest_models.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_18_pytest_models.log | tail -30
timeout 300 python tools/infer_profile.py --model fsrgan --list 12 > gpurun_out/r2_18_infer_fsrgan.log 2>&1
cat gpurun_out/r2_18_infer_fsrgan.log
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_18_bench_infer_fsrgan.log 2>&1
timeout 300 python bench.py --workload fsrgan --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_18_bench_fsrgan.log 2>&1
DG_DW_TMA=0 timeout 300 python bench.py --workload fsrgan --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_18_bench_fsrgan_off.log 2>&1
grep -h '"value"' gpurun_out/r2_18_bench*.log | cut -c1-200
