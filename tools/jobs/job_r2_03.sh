#!/bin/bash
# round 2, job 3: fused conv+BN+act launch (kernel test, model tests), prologue marks, step benches with/without fusion and PDL
set -x
mkdir -p gpurun_out
DG_ALL_ROLES=1 timeout 120 python tools/conv_timeline.py > gpurun_out/r2_03_conv_timeline.log 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "bn_act_one_launch or fused_bn_partials or bn_act_fwd_bwd" > gpurun_out/r2_03_pytest_kernels.log 2>&1
tail -5 gpurun_out/r2_03_pytest_kernels.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_03_pytest.log 2>&1
tail -5 gpurun_out/r2_03_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_03_bench_fused.log 2>&1
DG_CONV_BN_ACT=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_03_bench_unfused.log 2>&1
DG_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_03_bench_fused_pdl.log 2>&1
timeout 300 python bench.py --workload srgan_c3_vgg --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_03_bench_vgg.log 2>&1
grep -h '"value"' gpurun_out/r2_03_bench_*.log | cut -c1-200
