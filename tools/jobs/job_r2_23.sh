#!/bin/bash
# round 2, job 23: weight gradients of identical layers, several per launch
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -q --timeout 600 --timeout-method thread"
timeout 600 python -u -m pytest -x -v --timeout 300 tests/test_kernels_gpu.py -k "wgrad" > gpurun_out/r2_23_pytest_new.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_23_pytest_new.log | tail -25
timeout 1800 $PYT tests/test_srgan_gpu.py tests/test_models_gpu.py tests/test_checkpoint_gpu.py tests/test_train_loop_gpu.py > gpurun_out/r2_23_pytest_models.log 2>&1
tail -12 gpurun_out/r2_23_pytest_models.log | cut -c1-250
for b in 2 1 3 4; do
  DG_WGRAD_BATCH=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_23_bench_b$b.log 2>&1
done
grep -h '"value"' gpurun_out/r2_23_bench_b*.log | cut -c1-200
DG_WGRAD_BATCH=2 timeout 300 python bench.py --workload fsrgan --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_23_bench_fsrgan.log 2>&1
DG_WGRAD_BATCH=2 timeout 300 python bench.py --workload ae_c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_23_bench_ae.log 2>&1
DG_WGRAD_BATCH=2 timeout 300 python bench.py --workload pix2pix_c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_23_bench_pix2pix.log 2>&1
grep -h '"value"' gpurun_out/r2_23_bench_fsrgan.log gpurun_out/r2_23_bench_ae.log gpurun_out/r2_23_bench_pix2pix.log | cut -c1-200
