#!/bin/bash
# round 2, job 16: physically padded autoencoder + ReLU backward folded into pooling / up-sampling; VGG path check
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 300 --timeout-method thread"
timeout 600 $PYT tests/test_kernels_gpu.py -k "pool or channel_segments or structural" > gpurun_out/r2_16_pytest_new.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_16_pytest_new.log | tail -12
timeout 1200 $PYT tests/test_models_gpu.py tests/test_infer_gpu.py > gpurun_out/r2_16_pytest_models.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert|relative L2|cosine" gpurun_out/r2_16_pytest_models.log | tail -40
timeout 1200 $PYT tests/test_srgan_gpu.py -k "vgg" > gpurun_out/r2_16_pytest_vgg.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_16_pytest_vgg.log | tail -12
timeout 300 python bench.py --workload ae_c2 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_16_bench_ae.log 2>&1
DG_FOLD_RELU_BWD=0 timeout 300 python bench.py --workload ae_c2 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_16_bench_ae_nofold.log 2>&1
timeout 300 python bench.py --workload srgan_c3_vgg --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_16_bench_vgg.log 2>&1
grep -h '"value"' gpurun_out/r2_16_bench*.log | cut -c1-200
timeout 300 python tools/step_profile.py --model autoencoder --batch 64 --crop 256 > gpurun_out/r2_16_step_profile_ae.log 2>&1
head -12 gpurun_out/r2_16_step_profile_ae.log
timeout 300 python tools/conv_calls.py --model autoencoder --batch 64 --crop 256 --top 30 > gpurun_out/r2_16_calls_ae.log 2>&1
head -40 gpurun_out/r2_16_calls_ae.log
