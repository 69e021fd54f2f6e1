#!/bin/bash
# round 2, job 30: end-of-round evidence at HEAD -- full GPU suite, every bench line, launch list, ncu --set full of the top kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_30_smi.txt
timeout 2400 python -u -m pytest -x -q --timeout 600 --timeout-method thread tests -m gpu > gpurun_out/r2_30_pytest_all.log 2>&1
tail -4 gpurun_out/r2_30_pytest_all.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_30_bench_default.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_30_bench_reference.log 2>&1
for w in srgan_c3_vgg ae_c2 fsrgan pix2pix_c4 infer_fsrgan_1080p infer_ae_1080p; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_30_bench_$w.log 2>&1
done
grep -h '"metric"\|"impl"' gpurun_out/r2_30_bench_*.log | cut -c1-260
timeout 300 python tools/bench_conv.py --iters 24 > gpurun_out/r2_30_bench_conv.jsonl 2> gpurun_out/r2_30_bench_conv.err
cat gpurun_out/r2_30_bench_conv.jsonl | head -8
timeout 300 python tools/step_profile.py > gpurun_out/r2_30_step_profile_srgan.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_30_launches.csv python tools/step_profile.py > gpurun_out/r2_30_ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel|umma_wgrad_kernel|wgrad_reduce" -c 10 -o /tmp/r2_30_body python tools/bench_conv.py --only body_fwd,body_dgrad,body_wgrad,body_wgrad_x4 --iters 2 --graph 0 > gpurun_out/r2_30_ncu_body.log 2>&1
ncu -i /tmp/r2_30_body.ncu-rep --page raw --csv > gpurun_out/r2_30_body_raw.csv 2>/dev/null
DG_PROBE_MODES="res+bn" timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_conv_kernel" -c 4 -o /tmp/r2_30_fused python tools/dgrad_fused_probe.py > gpurun_out/r2_30_ncu_fused.log 2>&1
ncu -i /tmp/r2_30_fused.ncu-rep --page raw --csv > gpurun_out/r2_30_fused_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dw3x3_tma_kernel" -c 3 -o /tmp/r2_30_dw python tools/infer_profile.py --model fsrgan > gpurun_out/r2_30_ncu_dw.log 2>&1
ncu -i /tmp/r2_30_dw.ncu-rep --page raw --csv > gpurun_out/r2_30_dw_raw.csv 2>/dev/null
ls -la gpurun_out | grep r2_30
