#!/bin/bash
# round 2, job 33 (8 GPUs): final data-parallel lines of C3 (with and without the VGG content loss) at HEAD
set -x
mkdir -p gpurun_out
for w in srgan_c3 srgan_c3_vgg; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_33_bench_n8_$w.log 2>&1
done
grep -h '"value"' gpurun_out/r2_33_bench_n8_*.log | cut -c1-260
