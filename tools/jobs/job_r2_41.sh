#!/bin/bash
# round 2, job 41: phase timeline of the fused Fast-SRGAN block
set -x
mkdir -p gpurun_out
timeout 300 python tools/fsrgan_block_timeline.py > gpurun_out/r2_41_fb_timeline.log 2>&1; cat gpurun_out/r2_41_fb_timeline.log | cut -c1-250
