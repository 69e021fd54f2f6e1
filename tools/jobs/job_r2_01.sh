#!/bin/bash
# round 2, job 1: HEAD baseline on today's box + probes queued in round 1 + ncu at HEAD
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_01_smi.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_01_bench.log 2>&1
timeout 120 ./probes/multicast_probe > gpurun_out/r2_01_multicast.log 2>&1
timeout 200 ./probes/umma_probe > gpurun_out/r2_01_umma_probe.log 2>&1
timeout 300 python tools/bench_conv.py --iters 24 > gpurun_out/r2_01_bench_conv.jsonl 2> gpurun_out/r2_01_bench_conv.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_01_launches.csv python tools/step_profile.py > gpurun_out/r2_01_ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel|umma_wgrad_kernel|wgrad_reduce" -c 8 -o gpurun_out/r2_01_body python tools/bench_conv.py --only body_fwd,body_dgrad,body_wgrad --iters 2 --graph 0 > gpurun_out/r2_01_ncu_body.log 2>&1
ls -la gpurun_out
