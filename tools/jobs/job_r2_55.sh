#!/bin/bash
# round 2, job 55: ncu --set full of the warp-specialised block kernel at HEAD; launch list of one frame
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsrgan_block_ws_kernel -s 3 -c 2 -o /tmp/r2_55_fb python tools/fsrgan_block_timeline.py > gpurun_out/r2_55_ncu_fb.log 2>&1
ncu -i /tmp/r2_55_fb.ncu-rep --page raw --csv > gpurun_out/r2_55_fb_ws_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_55_launches_infer.csv python tools/infer_profile.py --model fsrgan --list 1 > gpurun_out/r2_55_ncu_launches.log 2>&1
timeout 300 python tools/infer_profile.py --model fsrgan --list 3 > gpurun_out/r2_55_infer_fsrgan.log 2>&1; head -14 gpurun_out/r2_55_infer_fsrgan.log
ls -la gpurun_out/r2_55_*
