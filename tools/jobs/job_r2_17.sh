#!/bin/bash
# round 2, job 17: where the Fast-SRGAN frame / train step and the pix2pix step spend their time
set -x
mkdir -p gpurun_out
timeout 300 python tools/infer_profile.py --model fsrgan --list 25 > gpurun_out/r2_17_infer_fsrgan.log 2>&1
cat gpurun_out/r2_17_infer_fsrgan.log
timeout 300 python tools/infer_profile.py --model autoencoder --list 12 > gpurun_out/r2_17_infer_ae.log 2>&1
cat gpurun_out/r2_17_infer_ae.log
timeout 300 python tools/step_profile.py --model fsrgan --batch 16 --crop 384 > gpurun_out/r2_17_step_fsrgan.log 2>&1
cat gpurun_out/r2_17_step_fsrgan.log
timeout 300 python tools/step_profile.py --model pix2pix --batch 32 --crop 256 > gpurun_out/r2_17_step_pix2pix.log 2>&1
cat gpurun_out/r2_17_step_pix2pix.log
