#!/bin/bash
# round 2, job 52: role timeline of the warp-specialised fused block
set -x
mkdir -p gpurun_out
timeout 120 python tools/fsrgan_block_timeline.py > gpurun_out/r2_52_fb_timeline.log 2>&1; head -1 gpurun_out/r2_52_fb_timeline.log; tail -9 gpurun_out/r2_52_fb_timeline.log | cut -c1-260
