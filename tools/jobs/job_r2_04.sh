#!/bin/bash
# round 2, job 4: trimmed prologue of the conv / wgrad kernels (parallel barrier init, producer ahead of the CTA sync, halo before weights)
set -x
mkdir -p gpurun_out
DG_ALL_ROLES=1 timeout 120 python tools/conv_timeline.py > gpurun_out/r2_04_conv_timeline.log 2>&1
timeout 120 python tools/wgrad_timeline.py 16 96 96 64 64 > gpurun_out/r2_04_wgrad_timeline.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_04_pytest.log 2>&1
tail -5 gpurun_out/r2_04_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_04_bench.log 2>&1
timeout 300 python tools/bench_conv.py > gpurun_out/r2_04_bench_conv.jsonl 2> gpurun_out/r2_04_bench_conv.err
grep -h '"value"' gpurun_out/r2_04_bench.log | cut -c1-300
