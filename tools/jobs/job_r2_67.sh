#!/bin/bash
# round 2, job 67: ncu evidence at HEAD for the headline step: launch list of four eager C3 steps, ncu --set full of the trunk conv / dgrad / wgrad
set -x
mkdir -p gpurun_out
timeout 300 python tools/step_profile.py > gpurun_out/r2_67_step_profile.log 2>&1; head -12 gpurun_out/r2_67_step_profile.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_67_launches.csv python tools/step_profile.py > gpurun_out/r2_67_ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel|umma_wgrad_kernel|wgrad_reduce" -c 8 -o /tmp/r2_67_body python tools/bench_conv.py --only body_fwd,body_dgrad,body_wgrad --iters 2 --graph 0 > gpurun_out/r2_67_ncu_body.log 2>&1
ncu -i /tmp/r2_67_body.ncu-rep --page raw --csv > gpurun_out/r2_67_body_raw.csv 2>/dev/null
ls -la gpurun_out/r2_67_*
