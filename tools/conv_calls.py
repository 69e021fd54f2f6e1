#!/usr/bin/env python
"""Per-call timing of the conv families of one eager train step (CUDA events around every conv / wgrad call, GPU parked
first so the events bracket device time only): lists the slowest calls with their algorithmic FLOPs and TFLOP/s.
usage: python tools/conv_calls.py [--model srgan|fsrgan|autoencoder|pix2pix] [--batch B] [--crop C] [--top N]"""
import argparse
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="pix2pix", choices=["srgan", "fsrgan", "autoencoder", "pix2pix"])
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--crop", type=int, default=256)
ap.add_argument("--top", type=int, default=40)
args = ap.parse_args()
from denoise_gan_b200.dataloader import synthetic_pair  # noqa: E402

ns = SimpleNamespace(crop_size=args.crop, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0, retrain=0)
if args.model == "srgan":
    from denoise_gan_b200.srgan import SRGAN as M
    from denoise_gan_b200.train_srgan import train_step
    scale = 4
elif args.model == "fsrgan":
    from denoise_gan_b200.fsrgan import FastSRGAN as M
    from denoise_gan_b200.train_fsrgan import train_step
    scale = 4
elif args.model == "pix2pix":
    from denoise_gan_b200.pix2pix import Pix2Pix as M
    from denoise_gan_b200.train_pix2pix import train_step
    scale = 1
else:
    from denoise_gan_b200.autoencoder import Autoencoder as M
    from denoise_gan_b200.train_autoencoder import train_step
    scale = 1
model = M(ns)
E = model.engine
E.wgrad_overlap = False          # one stream: every event pair brackets its own call
x, y = synthetic_pair(args.batch, args.crop, scale, step=0)
x, y = x.cuda(), y.cuda()
for _ in range(3):
    train_step(model, x, y)
torch.cuda.synchronize()
os.environ["DG_DEBUG_CONFIG"] = "0"
E.prof = []
torch.cuda._sleep(int(60e6))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
train_step(model, x, y)
e1.record()
torch.cuda.synchronize()
rows = [(kind, flops, a.elapsed_time(b)) for kind, flops, a, b in E.prof]
E.prof = None
tot = {}
for kind, flops, ms in rows:
    t = tot.setdefault(kind, [0, 0.0, 0.0]); t[0] += 1; t[1] += ms; t[2] += flops
for kind, (n, ms, fl) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{kind:12s} n={n:4d} {ms:8.3f} ms  {fl / 1e9:9.1f} GFLOP  {fl / ms / 1e9 if ms > 0 else 0:8.1f} TFLOP/s")
print(f"-- the {args.top} slowest calls (in tape order index)")
for i, (kind, flops, ms) in sorted(enumerate(rows), key=lambda t: -t[1][2])[:args.top]:
    print(f"#{i:4d} {kind:12s} {ms * 1e3:9.1f} us  {flops / 1e9:9.2f} GFLOP  {flops / ms / 1e9 if ms > 0 else 0:8.1f} TFLOP/s")
