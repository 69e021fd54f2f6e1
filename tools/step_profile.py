#!/usr/bin/env python
"""Warm, in-order timing of one eager SRGAN train step by C-ABI entry point (CUDA events around every call).
Complements the ncu launch list (cold-cache, serialised): this one sees the L2-resident behaviour of the step.
usage: python tools/step_profile.py [--fp16 0|1] [--batch 16] [--crop 384]"""
import argparse
import collections
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--fp16", type=int, default=1)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--crop", type=int, default=384)
ap.add_argument("--model", default="srgan", choices=["srgan", "fsrgan", "autoencoder", "pix2pix"])
args = ap.parse_args()

from denoise_gan_b200.dataloader import synthetic_pair  # noqa: E402

ns = SimpleNamespace(crop_size=args.crop, scale=4, lr=1e-3, fp16=args.fp16, vgg=0, seed=0, retrain=0)
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
x, y = synthetic_pair(args.batch, args.crop, scale, step=0)
x, y = x.cuda(), y.cuda()
for _ in range(3):
    train_step(model, x, y)
torch.cuda.synchronize()
model.engine.prof_calls = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
train_step(model, x, y)
e1.record()
torch.cuda.synchronize()
calls = model.engine.prof_calls
model.engine.prof_calls = None
tot = collections.defaultdict(lambda: [0, 0.0])
for name, a, b in calls:
    tot[name][0] += 1
    tot[name][1] += a.elapsed_time(b)
s = sum(v[1] for v in tot.values())
print(f"{len(calls)} calls, {s:.3f} ms inside calls, {e0.elapsed_time(e1):.3f} ms step wall (eager, events add launch gaps)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} n={v[0]:4d} {v[1]:8.3f} ms {100 * v[1] / s:5.1f}%  avg {1e3 * v[1] / v[0]:8.1f} us")
