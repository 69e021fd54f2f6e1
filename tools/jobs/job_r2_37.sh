#!/bin/bash
# round 2, job 37: batched weight gradients for every geometry with a single-launch plan (dg_umma_conv2d_wgrad_batch_supported)
set -x
mkdir -p gpurun_out
timeout 900 python -u -m pytest -x -q --timeout 600 tests -m gpu > gpurun_out/r2_37_pytest.log 2>&1; tail -4 gpurun_out/r2_37_pytest.log | cut -c1-220
timeout 300 python tools/debug_batch_wgrad.py 96 4 > gpurun_out/r2_37_debug_wgrad.log 2>&1; tail -3 gpurun_out/r2_37_debug_wgrad.log | cut -c1-400
for w in srgan_c3 pix2pix_c4 fsrgan; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_37_bench_$w.log 2>&1
  DG_WGRAD_BATCH_WIDE=0 timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_37_bench_${w}_off.log 2>&1
done
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r2_37_bench_*.log
