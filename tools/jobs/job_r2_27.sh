#!/bin/bash
# round 2, job 27 (2 GPUs): data-parallel correctness with the batched weight gradients, N=2 lines of C3 and pix2pix
set -x
mkdir -p gpurun_out
timeout 900 python -u -m pytest -x -v --timeout 600 tests/test_parallel_gpu.py > gpurun_out/r2_27_pytest_parallel.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_27_pytest_parallel.log | tail -8
for w in srgan_c3 pix2pix_c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload $w --steps 12 --warmup 3 --no-cpu > gpurun_out/r2_27_bench_n2_$w.log 2>&1
done
timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/r2_27_bench_n1_srgan_c3.log 2>&1
grep -H '"value"' gpurun_out/r2_27_bench_*.log | cut -c1-230
