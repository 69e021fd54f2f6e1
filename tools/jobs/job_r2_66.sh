#!/bin/bash
# round 2, job 66 (2 GPUs): data-parallel sanity at HEAD: 2-rank NCCL train-step test, C3 and 1080p inference bench lines at N=2
set -x
mkdir -p gpurun_out
timeout 400 python -u -m pytest -x -q --timeout 300 --timeout-method thread tests/test_parallel_gpu.py > gpurun_out/r2_66_pytest_parallel.log 2>&1; tail -3 gpurun_out/r2_66_pytest_parallel.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_66_bench_n2_srgan_c3.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload infer_fsrgan_1080p --steps 12 --warmup 3 --no-cpu > gpurun_out/r2_66_bench_n2_infer_fsrgan.log 2>&1
grep -h '"metric"' gpurun_out/r2_66_bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:24], 'n', d['n_gpus'], 'ms', round(d['ms_per_step'], 3), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1))"
tail -3 gpurun_out/r2_66_bench_n2_srgan_c3.log | cut -c1-200
