#!/bin/bash
# round 2, job 45: fused block v4: channel groups decoupled (group barriers, per-K-block project products, staggered start)
set -x
mkdir -p gpurun_out
timeout 300 python -u -m pytest -x -q --timeout 120 tests/test_kernels_gpu.py -k "fsrgan_block" > gpurun_out/r2_45_pytest_new.log 2>&1; tail -3 gpurun_out/r2_45_pytest_new.log | cut -c1-300
for sg in 0 800 1200 1800 2600; do
  DG_FSRGAN_BLOCK_STAGGER=$sg timeout 120 python tools/fsrgan_block_timeline.py > gpurun_out/r2_45_fb_timeline_$sg.log 2>&1; echo "stagger $sg: $(head -1 gpurun_out/r2_45_fb_timeline_$sg.log)"
done
sed -n 6,9p gpurun_out/r2_45_fb_timeline_1200.log | cut -c1-330
timeout 600 python -u -m pytest -x -q --timeout 600 tests/test_infer_gpu.py -k fsrgan > gpurun_out/r2_45_pytest_infer.log 2>&1; tail -3 gpurun_out/r2_45_pytest_infer.log | cut -c1-200
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_45_bench_infer_fsrgan.log 2>&1
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r2_45_bench_*.log
