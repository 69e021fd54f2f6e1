#!/bin/bash
# round 2, job 63: end-of-round evidence at HEAD (fourth part of the round): full GPU suite, every bench line, reference arm, frame profiles, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_63_smi.txt
timeout 1500 python -u -m pytest -x -q --timeout 600 --timeout-method thread tests -m gpu > gpurun_out/r2_63_pytest_all.log 2>&1
tail -6 gpurun_out/r2_63_pytest_all.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_63_bench_srgan_c3.log 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_63_bench_reference.log 2>&1
for w in srgan_c3_vgg ae_c2 fsrgan pix2pix_c4 infer_fsrgan_1080p infer_ae_1080p; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_63_bench_$w.log 2>&1
done
grep -h '"metric"\|"impl"' gpurun_out/r2_63_bench_*.log | cut -c1-230
timeout 300 python tools/infer_profile.py --model fsrgan --list 3 > gpurun_out/r2_63_infer_fsrgan.log 2>&1
timeout 300 python tools/infer_profile.py --model autoencoder --list 3 > gpurun_out/r2_63_infer_ae.log 2>&1; head -14 gpurun_out/r2_63_infer_ae.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_63_smoke.log 2>&1; tail -2 gpurun_out/r2_63_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_63_launches_infer.csv python tools/infer_profile.py --model fsrgan --list 1 > gpurun_out/r2_63_ncu_launches.log 2>&1
ls -la gpurun_out/r2_63_*
