#!/bin/bash
# round 2, job 14: BatchNorm folded into the convolutions at inference; launch list at HEAD; AE step profile; ncu of wgrad / fused dgrad
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 200 --timeout-method thread"
timeout 600 $PYT tests/test_infer_gpu.py > gpurun_out/r2_14_pytest_infer.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_14_pytest_infer.log | tail -12
for w in infer_fsrgan_1080p infer_ae_1080p; do
  timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_14_bench_$w.log 2>&1
  DG_FOLD_BN=0 timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_14_bench_${w}_nofold.log 2>&1
done
grep -h '"value"' gpurun_out/r2_14_bench*.log | cut -c1-220
timeout 300 python tools/step_profile.py --model autoencoder --batch 64 --crop 256 > gpurun_out/r2_14_step_profile_ae.log 2>&1
cat gpurun_out/r2_14_step_profile_ae.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_14_launches.csv python tools/step_profile.py > gpurun_out/r2_14_ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_wgrad_kernel|wgrad_reduce" -c 6 -o /tmp/r2_14_wgrad python tools/bench_conv.py --only body_wgrad --iters 2 --graph 0 > gpurun_out/r2_14_ncu_wgrad.log 2>&1
ncu -i /tmp/r2_14_wgrad.ncu-rep --page raw --csv > gpurun_out/r2_14_wgrad_raw.csv 2>/dev/null
DG_PROBE_MODES="res+bn" timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel" -c 6 -o /tmp/r2_14_fused python tools/dgrad_fused_probe.py > gpurun_out/r2_14_ncu_fused.log 2>&1
ncu -i /tmp/r2_14_fused.ncu-rep --page raw --csv > gpurun_out/r2_14_fused_raw.csv 2>/dev/null
ls -la gpurun_out | tail -6
