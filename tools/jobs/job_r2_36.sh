#!/bin/bash
# round 2, job 36: depth_to_space store epilogue, second version
set -x
mkdir -p gpurun_out
timeout 600 python -u -m pytest -x -q --timeout 300 tests/test_kernels_gpu.py -k "d2s or umma_conv_fwd" > gpurun_out/r2_36_pytest_new.log 2>&1; tail -3 gpurun_out/r2_36_pytest_new.log | cut -c1-200
timeout 900 python -u -m pytest -x -q --timeout 600 tests/test_infer_gpu.py > gpurun_out/r2_36_pytest_infer.log 2>&1; tail -3 gpurun_out/r2_36_pytest_infer.log | cut -c1-200
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_36_bench_infer_fsrgan.log 2>&1
DG_FUSE_D2S=0 timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_36_bench_infer_fsrgan_off.log 2>&1
grep -H '"value"' gpurun_out/r2_36_bench_*.log | cut -c1-230
timeout 300 python tools/infer_profile.py --model fsrgan --list 4 > gpurun_out/r2_36_infer_fsrgan.log 2>&1; head -8 gpurun_out/r2_36_infer_fsrgan.log; tail -5 gpurun_out/r2_36_infer_fsrgan.log
