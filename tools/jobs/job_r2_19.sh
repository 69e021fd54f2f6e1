#!/bin/bash
# round 2, job 19: narrow fp32 store of the RGB-side convolutions; full GPU test pass at this HEAD
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -q --timeout 600 --timeout-method thread"
timeout 300 python -u -m pytest -x -v --timeout 300 tests/test_kernels_gpu.py -k "narrow" > gpurun_out/r2_19_pytest_new.log 2>&1
grep -E "PASSED|FAILED|SKIPPED|Error|assert" gpurun_out/r2_19_pytest_new.log | tail -12
timeout 2400 $PYT tests -m gpu > gpurun_out/r2_19_pytest_all.log 2>&1
tail -15 gpurun_out/r2_19_pytest_all.log
timeout 300 python tools/infer_profile.py --model fsrgan --list 8 > gpurun_out/r2_19_infer_fsrgan.log 2>&1
head -14 gpurun_out/r2_19_infer_fsrgan.log
for w in infer_fsrgan_1080p infer_ae_1080p srgan_c3 ae_c2 fsrgan pix2pix_c4; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_19_bench_$w.log 2>&1
done
grep -h '"value"' gpurun_out/r2_19_bench*.log | cut -c1-200
