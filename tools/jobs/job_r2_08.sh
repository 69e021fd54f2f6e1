#!/bin/bash
# round 2, job 8: new whole-model parity tests (C3 full shape bf16, bf16 loss curve, bf16 + VGG), warm per-call step profile, launch list
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v -s --timeout 400 --timeout-method thread"
timeout 900 $PYT tests/test_srgan_gpu.py -k "c3_full_shape or loss_curve_bf16 or with_vgg" > gpurun_out/r2_08_pytest_parity.log 2>&1
grep -E "PASSED|FAILED|C3 bf16|bf16 loss curve|bf16\+VGG|Error" gpurun_out/r2_08_pytest_parity.log | cut -c1-400
timeout 200 python tools/step_profile.py > gpurun_out/r2_08_step_profile.log 2>&1
cat gpurun_out/r2_08_step_profile.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_08_launches.csv python tools/step_profile.py > gpurun_out/r2_08_ncu_step.log 2>&1
