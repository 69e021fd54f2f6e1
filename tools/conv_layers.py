#!/usr/bin/env python
"""Per-launch table of the convolution calls of one warm, single-stream SRGAN train step (CUDA events on the
launching stream): kind, GFLOP, us, TFLOP/s, grouped by identical (kind, GFLOP).  Shows which layers hold the
conv families' time.  usage: python tools/conv_layers.py [--model srgan] [--batch 16]"""
import argparse
import collections
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--crop", type=int, default=384)
args = ap.parse_args()

from denoise_gan_b200.dataloader import synthetic_pair  # noqa: E402
from denoise_gan_b200.srgan import SRGAN  # noqa: E402
from denoise_gan_b200.train_srgan import train_step  # noqa: E402

ns = SimpleNamespace(crop_size=args.crop, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0, retrain=0)
model = SRGAN(ns)
x, y = synthetic_pair(args.batch, args.crop, 4, step=0)
x, y = x.cuda(), y.cuda()
E = model.engine
E.wgrad_overlap = False
for _ in range(3):
    train_step(model, x, y)
torch.cuda.synchronize()
E.prof = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
train_step(model, x, y)
e1.record()
torch.cuda.synchronize()
prof, E.prof = E.prof, None
groups = collections.OrderedDict()
for kind, flops, a, b in prof:
    g = groups.setdefault((kind, round(flops / 1e9, 3)), [0, 0.0])
    g[0] += 1
    g[1] += a.elapsed_time(b) * 1e3
tot = sum(g[1] for g in groups.values())
print(f"{len(prof)} conv calls, {tot / 1e3:.3f} ms inside, step wall {e0.elapsed_time(e1):.3f} ms (eager, single stream)")
for (kind, gf), (n, us) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print(f"{kind:12s} {gf:9.3f} GFLOP  n={n:3d}  avg {us / n:8.1f} us  total {us / 1e3:7.3f} ms {100 * us / tot:5.1f}%  {gf * n / us * 1e3:7.1f} TFLOP/s")
