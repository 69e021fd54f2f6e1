#!/bin/bash
# round 2, job 12: ncu evidence at HEAD -- launch list of the step (default switches) and full captures of the trunk kernels
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_12_launches.csv python tools/step_profile.py > gpurun_out/r2_12_ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel|umma_wgrad_kernel|wgrad_reduce" -c 8 -o gpurun_out/r2_12_body python tools/bench_conv.py --only body_fwd,body_dgrad,body_wgrad --iters 2 --graph 0 > gpurun_out/r2_12_ncu_body.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel" -c 6 -o gpurun_out/r2_12_dgrad_fused python tools/dgrad_fused_probe.py > gpurun_out/r2_12_ncu_fused.log 2>&1
ls -la gpurun_out | tail -5
