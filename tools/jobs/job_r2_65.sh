#!/bin/bash
# round 2, job 65: bench `value` of the inference workloads through FrameRunner.device_frame (CUDA-graph replay)
set -x
mkdir -p gpurun_out
timeout 600 python -u -m pytest -x -q --timeout 300 --timeout-method thread tests/test_infer_gpu.py -k "video" > gpurun_out/r2_65_pytest_infer.log 2>&1; tail -3 gpurun_out/r2_65_pytest_infer.log
timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_65_bench_infer_fsrgan.log 2>&1
timeout 300 python bench.py --workload infer_ae_1080p --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_65_bench_infer_ae.log 2>&1
grep -h '"metric"' gpurun_out/r2_65_bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:24], 'ms', round(d['ms_per_step'], 3), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), round(d['e2e']['ms_per_step'], 3), d['gpu_launches'])"
tail -3 gpurun_out/r2_65_bench_infer_ae.log | cut -c1-300
