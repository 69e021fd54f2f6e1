#!/usr/bin/env python
"""Warm, in-order timing of one 1080p inference frame (FrameRunner.video_frame, infer_video.py:138-159) by C-ABI entry point.
usage: python tools/infer_profile.py [--model fsrgan|autoencoder]"""
import argparse
import collections
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="fsrgan", choices=["fsrgan", "autoencoder"])
ap.add_argument("--list", type=int, default=0, help="print the N slowest single calls")
args = ap.parse_args()

from denoise_gan_b200.infer import FrameRunner  # noqa: E402

ns = SimpleNamespace(crop_size=256, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0, retrain=0)
if args.model == "fsrgan":
    from denoise_gan_b200.fsrgan import FastSRGAN as M
    up = 4
else:
    from denoise_gan_b200.autoencoder import Autoencoder as M
    up = 1
model = M(ns)
runner = FrameRunner(model, upscale=up)
frame = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(1080, 1920, 3), dtype=np.uint8)).cuda()
for _ in range(3):
    runner.video_frame(frame, to_host=False)
torch.cuda.synchronize()
model.engine.prof_calls = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
runner.video_frame(frame, to_host=False)
e1.record()
torch.cuda.synchronize()
calls = model.engine.prof_calls
model.engine.prof_calls = None
tot = collections.defaultdict(lambda: [0, 0.0])
for name, a, b in calls:
    tot[name][0] += 1
    tot[name][1] += a.elapsed_time(b)
s = sum(v[1] for v in tot.values())
print(f"{len(calls)} calls, {s:.3f} ms inside calls, {e0.elapsed_time(e1):.3f} ms frame wall (eager, events add launch gaps)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} n={v[0]:4d} {v[1]:8.3f} ms {100 * v[1] / s:5.1f}%  avg {1e3 * v[1] / v[0]:8.1f} us")
if args.list:
    print(f"-- the {args.list} slowest calls (call order index)")
    for i, (name, a, b) in sorted(enumerate(calls), key=lambda t: -t[1][1].elapsed_time(t[1][2]))[:args.list]:
        print(f"#{i:4d} {name:34s} {1e3 * a.elapsed_time(b):9.1f} us")
