#!/bin/bash
# round 2, job 44: fused block with early accumulator requests: tests, timeline, frame bench, ncu --set full of the block kernel, launch list of one frame
set -x
mkdir -p gpurun_out
timeout 300 python -u -m pytest -x -q --timeout 120 tests/test_kernels_gpu.py -k "fsrgan_block" > gpurun_out/r2_44_pytest_new.log 2>&1; tail -3 gpurun_out/r2_44_pytest_new.log | cut -c1-300
timeout 600 python -u -m pytest -x -q --timeout 600 tests/test_infer_gpu.py > gpurun_out/r2_44_pytest_infer.log 2>&1; tail -3 gpurun_out/r2_44_pytest_infer.log | cut -c1-200
timeout 300 python tools/fsrgan_block_timeline.py > gpurun_out/r2_44_fb_timeline.log 2>&1; head -3 gpurun_out/r2_44_fb_timeline.log | cut -c1-250; tail -1 gpurun_out/r2_44_fb_timeline.log
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_44_bench_infer_fsrgan.log 2>&1
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r2_44_bench_*.log
timeout 300 python tools/infer_profile.py --model fsrgan --list 3 > gpurun_out/r2_44_infer_fsrgan.log 2>&1; head -12 gpurun_out/r2_44_infer_fsrgan.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsrgan_block_kernel -s 3 -c 2 -o /tmp/r2_44_fb python tools/fsrgan_block_timeline.py > gpurun_out/r2_44_ncu_fb.log 2>&1
ncu -i /tmp/r2_44_fb.ncu-rep --page raw --csv > gpurun_out/r2_44_fb_raw.csv 2>/dev/null
ncu -i /tmp/r2_44_fb.ncu-rep --page details --csv > gpurun_out/r2_44_fb_details.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_44_launches_infer.csv python tools/infer_profile.py --model fsrgan --list 1 > gpurun_out/r2_44_ncu_launches.log 2>&1
ls -la gpurun_out/r2_44_*
