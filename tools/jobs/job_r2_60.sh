#!/bin/bash
# round 2, job 60: staged depth_to_space epilogue with packed fp32 pairs and slopes in registers: tests, per-call profile, bench line
set -x
mkdir -p gpurun_out
timeout 300 python -u -m pytest -x -q --timeout 120 --timeout-method thread tests/test_kernels_gpu.py -k "d2s" > gpurun_out/r2_60_pytest_k.log 2>&1
rc=$?; tail -25 gpurun_out/r2_60_pytest_k.log
if [ $rc -eq 0 ]; then
  timeout 600 python -u -m pytest -x -q --timeout 300 --timeout-method thread tests/test_infer_gpu.py > gpurun_out/r2_60_pytest_infer.log 2>&1; tail -5 gpurun_out/r2_60_pytest_infer.log
  timeout 300 python tools/infer_profile.py --model fsrgan --list 2 > gpurun_out/r2_60_infer_fsrgan.log 2>&1; head -8 gpurun_out/r2_60_infer_fsrgan.log; tail -3 gpurun_out/r2_60_infer_fsrgan.log
  timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_60_bench_infer_fsrgan.log 2>&1
  grep -h '"metric"' gpurun_out/r2_60_bench_*.log | cut -c1-200
fi
