#!/bin/bash
# round 2, job 11: bwd epilogue without memory clobbers / hoisted coefficients; pix2pix + autoencoder per-call conv timing
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 100 --timeout-method thread"
timeout 300 $PYT tests/test_kernels_gpu.py -k "dgrad_fused" > gpurun_out/r2_11_pytest_new.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_11_pytest_new.log | head -30
timeout 200 python tools/dgrad_fused_probe.py > gpurun_out/r2_11_probe.log 2>&1
head -5 gpurun_out/r2_11_probe.log; grep -A6 "^epilogue" gpurun_out/r2_11_probe.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_11_bench.log 2>&1
grep -h '"value"' gpurun_out/r2_11_bench*.log | cut -c1-200
timeout 300 python tools/conv_calls.py --model pix2pix --batch 32 --crop 256 > gpurun_out/r2_11_calls_pix2pix.log 2>&1
timeout 300 python tools/conv_calls.py --model autoencoder --batch 64 --crop 256 > gpurun_out/r2_11_calls_ae.log 2>&1
head -50 gpurun_out/r2_11_calls_pix2pix.log
