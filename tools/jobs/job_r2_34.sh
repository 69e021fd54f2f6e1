#!/bin/bash
# round 2, job 34: depth_to_space + PReLU as the store pattern of the up-sampling convolutions (inference)
set -x
mkdir -p gpurun_out
timeout 600 python -u -m pytest -x -v --timeout 300 tests/test_kernels_gpu.py -k "d2s" > gpurun_out/r2_34_pytest_new.log 2>&1
grep -E "PASSED|FAILED|Error|assert" gpurun_out/r2_34_pytest_new.log | tail -8
timeout 900 python -u -m pytest -x -q --timeout 600 tests/test_infer_gpu.py tests/test_kernels_gpu.py -k "not wgrad" > gpurun_out/r2_34_pytest_infer.log 2>&1
tail -4 gpurun_out/r2_34_pytest_infer.log | cut -c1-220
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_34_bench_infer_fsrgan.log 2>&1
DG_FUSE_D2S=0 timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_34_bench_infer_fsrgan_off.log 2>&1
grep -H '"value"' gpurun_out/r2_34_bench_*.log | cut -c1-220
timeout 300 python tools/infer_profile.py --model fsrgan --list 6 > gpurun_out/r2_34_infer_fsrgan.log 2>&1
head -18 gpurun_out/r2_34_infer_fsrgan.log
