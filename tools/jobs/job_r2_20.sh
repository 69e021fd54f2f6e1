#!/bin/bash
# round 2, job 20 (8 GPUs): data-parallel train step of C3 and C4, frame-sharded inference of C5, one line each
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_20_topo.txt 2>&1
for w in srgan_c3 pix2pix_c4 infer_fsrgan_1080p infer_ae_1080p; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload $w --steps 12 --warmup 3 --no-cpu > gpurun_out/r2_20_bench_n8_$w.log 2>&1
done
grep -h '"value"' gpurun_out/r2_20_bench_n8_*.log | cut -c1-260
