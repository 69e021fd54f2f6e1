#!/bin/bash
# round 2, job 26: pix2pix A/B again (per-problem producer instances), C3, wgrad kernel tests
set -x
mkdir -p gpurun_out
R=$PWD
timeout 600 python -u -m pytest -x -q --timeout 300 tests/test_kernels_gpu.py -k "wgrad" > gpurun_out/r2_26_pytest_new.log 2>&1
tail -3 gpurun_out/r2_26_pytest_new.log
(cd _ab/prev && timeout 300 python bench.py --workload pix2pix_c4 --steps 10 --warmup 3 --no-cpu > $R/gpurun_out/r2_26_pix2pix_prev.log 2>&1)
timeout 300 python bench.py --workload pix2pix_c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_26_pix2pix_head.log 2>&1
(cd _ab/prev && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $R/gpurun_out/r2_26_c3_prev.log 2>&1)
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_26_c3_head.log 2>&1
DG_WGRAD_BATCH=3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_26_c3_head_b3.log 2>&1
grep -H '"value"' gpurun_out/r2_26_*.log | cut -c1-200
