#!/bin/bash
# round 2, job 31: ReLU backward folded into the consumer's input-gradient launch (conv -> conv chains)
set -x
mkdir -p gpurun_out
timeout 600 python -u -m pytest -x -v --timeout 300 tests/test_kernels_gpu.py -k "relu_mask or dgrad" > gpurun_out/r2_31_pytest_new.log 2>&1
grep -E "FAILED|Error|assert|passed|failed" gpurun_out/r2_31_pytest_new.log | tail -8
timeout 1800 python -u -m pytest -x -q --timeout 600 tests/test_models_gpu.py tests/test_srgan_gpu.py tests/test_checkpoint_gpu.py tests/test_infer_gpu.py > gpurun_out/r2_31_pytest_models.log 2>&1
tail -6 gpurun_out/r2_31_pytest_models.log | cut -c1-250
for w in ae_c2 srgan_c3_vgg; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_31_bench_$w.log 2>&1
  DG_RELU_MASK_DGRAD=0 timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_31_bench_${w}_off.log 2>&1
done
grep -H '"value"' gpurun_out/r2_31_bench_*.log | cut -c1-220
timeout 300 python tools/step_profile.py --model autoencoder --batch 64 --crop 256 > gpurun_out/r2_31_step_profile_ae.log 2>&1
head -12 gpurun_out/r2_31_step_profile_ae.log
