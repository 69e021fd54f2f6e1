#!/bin/bash
# round 2, job 39: fused Fast-SRGAN block in the 1080p frame: parity, A/B, per-call profile
set -x
mkdir -p gpurun_out
timeout 900 python -u -m pytest -x -q --timeout 600 tests/test_infer_gpu.py > gpurun_out/r2_39_pytest_infer.log 2>&1; tail -3 gpurun_out/r2_39_pytest_infer.log | cut -c1-200
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_39_bench_infer_fsrgan.log 2>&1
DG_FSRGAN_BLOCK=0 timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_39_bench_infer_fsrgan_off.log 2>&1
grep -H '"value"' gpurun_out/r2_39_bench_*.log | cut -c1-230
timeout 300 python tools/infer_profile.py --model fsrgan --list 6 > gpurun_out/r2_39_infer_fsrgan.log 2>&1; head -14 gpurun_out/r2_39_infer_fsrgan.log; tail -7 gpurun_out/r2_39_infer_fsrgan.log
