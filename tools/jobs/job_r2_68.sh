#!/bin/bash
# round 2, job 68: last check at HEAD: full GPU suite, the two inference bench lines (executed / reference GFLOP in the line)
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_68_bench_infer_fsrgan.log 2>&1
timeout 300 python bench.py --workload infer_ae_1080p --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_68_bench_infer_ae.log 2>&1
grep -h '"metric"' gpurun_out/r2_68_bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:24], d['config']['compute_size'], 'ms', round(d['ms_per_step'], 3), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'tflops', round(d['step_tflops'], 1), d['step_gflop_executed'], d['roofline']['traffic'])"
tail -2 gpurun_out/r2_68_bench_infer_ae.log | cut -c1-400
timeout 200 python -u -m pytest -x -q --timeout 300 --timeout-method thread tests -m gpu > gpurun_out/r2_68_pytest_all.log 2>&1
tail -4 gpurun_out/r2_68_pytest_all.log
