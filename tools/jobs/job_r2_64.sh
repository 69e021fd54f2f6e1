#!/bin/bash
# round 2, job 64: video frames as CUDA-graph replays in FrameRunner.video(): tests, e2e A/B
set -x
mkdir -p gpurun_out
timeout 600 python -u -m pytest -x -q --timeout 300 --timeout-method thread tests/test_infer_gpu.py > gpurun_out/r2_64_pytest_infer.log 2>&1; tail -12 gpurun_out/r2_64_pytest_infer.log
for v in 1 0; do
  DG_INFER_GRAPH=$v timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_64_bench_infer_fsrgan_graph$v.log 2>&1
  DG_INFER_GRAPH=$v timeout 300 python bench.py --workload infer_ae_1080p --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_64_bench_infer_ae_graph$v.log 2>&1
done
grep -h '"metric"' gpurun_out/r2_64_bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:24], 'ms', round(d['ms_per_step'], 3), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), round(d['e2e']['ms_per_step'], 3))"
