#!/bin/bash
# round 2, job 49: end-of-round evidence at HEAD (third part of the round): full GPU suite, every bench line, reference arm, frame profile
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_49_smi.txt
timeout 2400 python -u -m pytest -x -q --timeout 600 --timeout-method thread tests -m gpu > gpurun_out/r2_49_pytest_all.log 2>&1
tail -4 gpurun_out/r2_49_pytest_all.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_49_bench_srgan_c3.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_49_bench_reference.log 2>&1
for w in srgan_c3_vgg ae_c2 fsrgan pix2pix_c4 infer_fsrgan_1080p infer_ae_1080p; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_49_bench_$w.log 2>&1
done
grep -h '"metric"\|"impl"' gpurun_out/r2_49_bench_*.log | cut -c1-260
timeout 300 python tools/infer_profile.py --model fsrgan --list 3 > gpurun_out/r2_49_infer_fsrgan.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_49_smoke.log 2>&1; tail -2 gpurun_out/r2_49_smoke.log
