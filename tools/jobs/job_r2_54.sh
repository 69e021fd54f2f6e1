#!/bin/bash
# round 2, job 54: N=8, 1080p Fast-SRGAN frames sharded across the GPUs at HEAD (one-launch block)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload infer_fsrgan_1080p --steps 12 --warmup 3 --no-cpu > gpurun_out/r2_54_bench_n8_infer_fsrgan_1080p.log 2>&1
grep -h '"value"' gpurun_out/r2_54_bench_n8_*.log | cut -c1-400
