#!/bin/bash
# round 2, job 24: batched weight gradients, group sizes decided from the tape; pix2pix A/B
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -q --timeout 600 --timeout-method thread"
timeout 1800 $PYT tests/test_srgan_gpu.py tests/test_models_gpu.py tests/test_checkpoint_gpu.py tests/test_train_loop_gpu.py tests/test_feed_gpu.py > gpurun_out/r2_24_pytest_models.log 2>&1
tail -8 gpurun_out/r2_24_pytest_models.log | cut -c1-250
for b in 4 1; do
  DG_WGRAD_BATCH=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_24_bench_c3_b$b.log 2>&1
  DG_WGRAD_BATCH=$b timeout 300 python bench.py --workload pix2pix_c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_24_bench_pix2pix_b$b.log 2>&1
  DG_WGRAD_BATCH=$b timeout 300 python bench.py --workload srgan_c3_vgg --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_24_bench_vgg_b$b.log 2>&1
done
grep -H '"value"' gpurun_out/r2_24_bench_*.log | cut -c1-230
