#!/bin/bash
# round 2, job 59: depth_to_space store of the 32 -> 128 up-convolutions through shared memory + one bulk tensor store per tile: tests, A/B
set -x
mkdir -p gpurun_out
timeout 300 python -u -m pytest -x -q --timeout 120 --timeout-method thread tests/test_kernels_gpu.py -k "d2s or narrow or tapsum" > gpurun_out/r2_59_pytest_k.log 2>&1
rc=$?; tail -25 gpurun_out/r2_59_pytest_k.log
if [ $rc -eq 0 ]; then
  timeout 600 python -u -m pytest -x -q --timeout 300 --timeout-method thread tests/test_infer_gpu.py > gpurun_out/r2_59_pytest_infer.log 2>&1; tail -5 gpurun_out/r2_59_pytest_infer.log
  timeout 300 python tools/infer_profile.py --model fsrgan --list 2 > gpurun_out/r2_59_infer_fsrgan.log 2>&1; head -8 gpurun_out/r2_59_infer_fsrgan.log; tail -3 gpurun_out/r2_59_infer_fsrgan.log
  DG_DEBUG_NO_D2S_TSTORE=1 timeout 300 python tools/infer_profile.py --model fsrgan --list 2 > gpurun_out/r2_59_infer_fsrgan_direct.log 2>&1; grep -h "frame wall\|d2s_prelu" gpurun_out/r2_59_infer_fsrgan_direct.log
  timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_59_bench_infer_fsrgan.log 2>&1
  grep -h '"metric"' gpurun_out/r2_59_bench_*.log | cut -c1-200
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_conv_kernel -s 18 -c 2 -o /tmp/r2_59_uc python tools/infer_profile.py --model fsrgan > gpurun_out/r2_59_ncu_uc.log 2>&1
  ncu -i /tmp/r2_59_uc.ncu-rep --page raw --csv > gpurun_out/r2_59_upconv_raw.csv 2>/dev/null
fi
ls -la gpurun_out/r2_59_*
