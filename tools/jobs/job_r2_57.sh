#!/bin/bash
# round 2, job 57: tap-sum kernel v2 (shuffle-combined filter rows, 4 epilogue groups) + tight padding of video frames: tests, A/B, profile, ncu
set -x
mkdir -p gpurun_out
timeout 300 python -u -m pytest -x -q --timeout 120 --timeout-method thread tests/test_kernels_gpu.py -k "tapsum" > gpurun_out/r2_57_pytest_tapsum.log 2>&1
rc=$?; tail -25 gpurun_out/r2_57_pytest_tapsum.log
if [ $rc -eq 0 ]; then
  timeout 600 python -u -m pytest -x -q --timeout 300 --timeout-method thread tests/test_infer_gpu.py > gpurun_out/r2_57_pytest_infer.log 2>&1; tail -25 gpurun_out/r2_57_pytest_infer.log
  timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_57_bench_infer_fsrgan.log 2>&1
  DG_INFER_TIGHT_PAD=0 timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_57_bench_infer_fsrgan_fullpad.log 2>&1
  grep -h '"metric"' gpurun_out/r2_57_bench_*.log | cut -c1-200
  timeout 300 python tools/infer_profile.py --model fsrgan --list 3 > gpurun_out/r2_57_infer_fsrgan.log 2>&1; head -16 gpurun_out/r2_57_infer_fsrgan.log
  DG_INFER_TIGHT_PAD=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tapsum_kernel -s 2 -c 1 -o /tmp/r2_57_ts python tools/infer_profile.py --model fsrgan > gpurun_out/r2_57_ncu_ts.log 2>&1
  ncu -i /tmp/r2_57_ts.ncu-rep --page raw --csv > gpurun_out/r2_57_tapsum_raw.csv 2>/dev/null
fi
ls -la gpurun_out/r2_57_*
