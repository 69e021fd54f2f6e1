#!/bin/bash
# round 2, job 13: small-map weight gradients as one dense product (pix2pix), chunk blocks along gridDim.z; ncu full captures at HEAD
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 150 --timeout-method thread"
timeout 400 $PYT tests/test_kernels_gpu.py -k "small_map or wgrad" > gpurun_out/r2_13_pytest_new.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_13_pytest_new.log | tail -30
timeout 400 $PYT tests/test_models_gpu.py > gpurun_out/r2_13_pytest_models.log 2>&1
tail -4 gpurun_out/r2_13_pytest_models.log
timeout 300 python bench.py --workload pix2pix_c4 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_13_bench_pix2pix.log 2>&1
DG_SMALL_MAP_GEMM=0 timeout 300 python bench.py --workload pix2pix_c4 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_13_bench_pix2pix_off.log 2>&1
grep -h '"value"' gpurun_out/r2_13_bench*.log | cut -c1-200
timeout 300 python tools/conv_calls.py --model pix2pix --batch 32 --crop 256 --top 25 > gpurun_out/r2_13_calls_pix2pix.log 2>&1
head -32 gpurun_out/r2_13_calls_pix2pix.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel|umma_wgrad_kernel|wgrad_reduce" -c 8 -o /tmp/r2_13_body python tools/bench_conv.py --only body_fwd,body_dgrad,body_wgrad --iters 2 --graph 0 > gpurun_out/r2_13_ncu_body.log 2>&1
ncu -i /tmp/r2_13_body.ncu-rep --page raw --csv > gpurun_out/r2_13_body_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel" -c 24 -o /tmp/r2_13_fused python tools/dgrad_fused_probe.py > gpurun_out/r2_13_ncu_fused.log 2>&1
ncu -i /tmp/r2_13_fused.ncu-rep --page raw --csv > gpurun_out/r2_13_fused_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
