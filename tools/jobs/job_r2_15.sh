#!/bin/bash
# round 2, job 15: physically padded activations for the autoencoder (segments, vector pool / upsample kernels, in-place concat half)
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 200 --timeout-method thread"
timeout 600 $PYT tests/test_kernels_gpu.py -k "pool_upsample or channel_segments or structural" > gpurun_out/r2_15_pytest_new.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_15_pytest_new.log | tail -12
timeout 900 $PYT tests/test_models_gpu.py tests/test_infer_gpu.py > gpurun_out/r2_15_pytest_models.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_15_pytest_models.log | tail -30
timeout 300 python bench.py --workload ae_c2 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_15_bench_ae.log 2>&1
DG_PHYS_PAD=0 timeout 300 python bench.py --workload ae_c2 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_15_bench_ae_off.log 2>&1
timeout 300 python bench.py --workload infer_ae_1080p --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_15_bench_infer_ae.log 2>&1
grep -h '"value"' gpurun_out/r2_15_bench*.log | cut -c1-200
tail -5 gpurun_out/r2_15_bench_ae.log | cut -c1-600
timeout 300 python tools/step_profile.py --model autoencoder --batch 64 --crop 256 > gpurun_out/r2_15_step_profile_ae.log 2>&1
cat gpurun_out/r2_15_step_profile_ae.log
