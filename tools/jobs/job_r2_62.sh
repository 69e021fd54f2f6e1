#!/bin/bash
# round 2, job 62: PReLU / skip-add behind a folded BatchNorm in the staged conv epilogue (dg_umma_conv2d_fwd_res_prelu): tests, A/B, profile
set -x
mkdir -p gpurun_out
timeout 300 python -u -m pytest -x -q --timeout 120 --timeout-method thread tests/test_kernels_gpu.py -k "res_prelu or d2s or tapsum or narrow" > gpurun_out/r2_62_pytest_k.log 2>&1
rc=$?; tail -25 gpurun_out/r2_62_pytest_k.log
if [ $rc -eq 0 ]; then
  timeout 600 python -u -m pytest -x -q -s --timeout 300 --timeout-method thread tests/test_infer_gpu.py > gpurun_out/r2_62_pytest_infer.log 2>&1; tail -12 gpurun_out/r2_62_pytest_infer.log
  timeout 300 python tools/infer_profile.py --model fsrgan --list 2 > gpurun_out/r2_62_infer_fsrgan.log 2>&1; head -12 gpurun_out/r2_62_infer_fsrgan.log
  timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_62_bench_infer_fsrgan.log 2>&1
  DG_FUSE_RES_EPILOGUE=0 timeout 300 python bench.py --workload infer_fsrgan_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_62_bench_infer_fsrgan_nofuse.log 2>&1
  grep -h '"metric"' gpurun_out/r2_62_bench_*.log | cut -c1-200
fi
# U-Net concats padded to 32-channel multiples (Engine.concat_pad32): autoencoder tests, A/B of the train step and of the 1080p frame
timeout 600 python -u -m pytest -x -q --timeout 300 --timeout-method thread tests -m gpu -k "autoencoder or ae_" > gpurun_out/r2_62_pytest_ae.log 2>&1; tail -5 gpurun_out/r2_62_pytest_ae.log
for v in 1 0; do
  DG_CONCAT_PAD32=$v timeout 300 python bench.py --workload ae_c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_62_bench_ae_c2_pad$v.log 2>&1
  DG_CONCAT_PAD32=$v timeout 300 python bench.py --workload infer_ae_1080p --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_62_bench_infer_ae_pad$v.log 2>&1
done
grep -h '"metric"' gpurun_out/r2_62_bench_ae_c2_pad*.log gpurun_out/r2_62_bench_infer_ae_pad*.log | cut -c1-200
timeout 300 python tools/infer_profile.py --model autoencoder --list 3 > gpurun_out/r2_62_infer_ae.log 2>&1; head -8 gpurun_out/r2_62_infer_ae.log; tail -4 gpurun_out/r2_62_infer_ae.log
