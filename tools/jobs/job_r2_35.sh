#!/bin/bash
# round 2, job 35: 64-channel N blocks for wide layers with a short contraction (Fast-SRGAN expand convolutions)
set -x
mkdir -p gpurun_out
timeout 900 python -u -m pytest -x -q --timeout 600 tests/test_kernels_gpu.py -k "umma or depthwise or bn" > gpurun_out/r2_35_pytest_k.log 2>&1; tail -3 gpurun_out/r2_35_pytest_k.log | cut -c1-200
timeout 900 python -u -m pytest -x -q --timeout 600 tests/test_models_gpu.py tests/test_infer_gpu.py > gpurun_out/r2_35_pytest_m.log 2>&1; tail -3 gpurun_out/r2_35_pytest_m.log | cut -c1-200
for w in infer_fsrgan_1080p fsrgan; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_35_bench_$w.log 2>&1
  DG_DEBUG_NO_SMALLK_NB64=1 timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_35_bench_${w}_off.log 2>&1
done
grep -H '"value"' gpurun_out/r2_35_bench_*.log | cut -c1-230
timeout 300 python tools/infer_profile.py --model fsrgan --list 6 > gpurun_out/r2_35_infer_fsrgan.log 2>&1; head -8 gpurun_out/r2_35_infer_fsrgan.log
timeout 300 python tools/step_profile.py --model fsrgan --batch 16 --crop 384 > gpurun_out/r2_35_step_fsrgan.log 2>&1; head -14 gpurun_out/r2_35_step_fsrgan.log
