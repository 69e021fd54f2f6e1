#!/bin/bash
# round 2, job 5: dgrad with fused skip-add + BatchNorm-backward sums, finalize folded into the BN apply pass; hang hunt in the full suite
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 100 --timeout-method thread"
timeout 300 $PYT tests/test_kernels_gpu.py -k "dgrad_fused or bn_act_fwd_from_partials" > gpurun_out/r2_05_pytest_new.log 2>&1
tail -15 gpurun_out/r2_05_pytest_new.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_05_bench.log 2>&1
DG_PDL=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_05_bench_pdl.log 2>&1
DG_DGRAD_BN_BWD=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_05_bench_nobwd.log 2>&1
DG_DGRAD_BN_BWD=0 DG_BN_FINALIZE_APPLY=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_05_bench_base.log 2>&1
grep -h '"value"' gpurun_out/r2_05_bench*.log | cut -c1-200
timeout 900 $PYT tests -m gpu > gpurun_out/r2_05_pytest.log 2>&1
tail -30 gpurun_out/r2_05_pytest.log
