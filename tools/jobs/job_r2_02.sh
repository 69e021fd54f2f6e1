#!/bin/bash
# round 2, job 2: timelines at HEAD, PDL A/B, new bench.py workloads, full GPU test suite
set -x
mkdir -p gpurun_out
for f in 0 1 2 4; do DG_ALL_ROLES=1 DG_FLAGS=$f timeout 120 python tools/conv_timeline.py > gpurun_out/r2_02_conv_timeline_f$f.log 2>&1; done
timeout 120 python tools/wgrad_timeline.py 16 96 96 64 64 > gpurun_out/r2_02_wgrad_timeline.log 2>&1
DG_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_02_bench_pdl.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_02_bench.log 2>&1
for w in srgan_c3_vgg ae_c2 fsrgan pix2pix_c4 infer_fsrgan_1080p infer_ae_1080p; do
  timeout 400 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_02_bench_$w.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_02_pytest.log 2>&1
tail -3 gpurun_out/r2_02_pytest.log
