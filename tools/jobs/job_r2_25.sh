#!/bin/bash
# round 2, job 25: pix2pix A/B between HEAD (batched wgrad kernel) and the commit before it, same box
set -x
mkdir -p gpurun_out
R=$PWD
for i in 1 2; do
  (cd _ab/prev && timeout 300 python bench.py --workload pix2pix_c4 --steps 10 --warmup 3 --no-cpu > $R/gpurun_out/r2_25_pix2pix_prev_$i.log 2>&1)
  timeout 300 python bench.py --workload pix2pix_c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_25_pix2pix_head_$i.log 2>&1
done
grep -H '"value"' gpurun_out/r2_25_pix2pix_*.log | cut -c1-200
(cd _ab/prev && timeout 300 python tools/step_profile.py --model pix2pix --batch 32 --crop 256 > $R/gpurun_out/r2_25_step_prev.log 2>&1)
timeout 300 python tools/step_profile.py --model pix2pix --batch 32 --crop 256 > gpurun_out/r2_25_step_head.log 2>&1
head -8 gpurun_out/r2_25_step_prev.log; head -8 gpurun_out/r2_25_step_head.log
