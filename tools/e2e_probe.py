#!/usr/bin/env python
"""Breaks the end-to-end step (host batch -> losses on the host) into its parts under the prefetching feed."""
import os, sys, time
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200.dataloader import synthetic_pair
from denoise_gan_b200.graph import GraphedStep, DevicePrefetcher
from denoise_gan_b200.srgan import SRGAN
from denoise_gan_b200.train_srgan import train_step

ns = SimpleNamespace(crop_size=384, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0, retrain=0)
model = SRGAN(ns)
x_h, y_h = synthetic_pair(16, 384, 4, step=0)
x_h, y_h = x_h.pin_memory(), y_h.pin_memory()
step = GraphedStep(model, train_step, x_h, y_h, warmup=2)
K = 20
def timeit(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(K): fn()
    torch.cuda.synchronize(); print(f"{name:40s} {(time.perf_counter() - t0) / K * 1e3:8.3f} ms/iter", flush=True)
timeit("graph only", lambda: step())
timeit("graph + sync each", lambda: (step(), torch.cuda.synchronize()))
timeit("h2d same stream + graph + sync", lambda: (step(x_h, y_h), torch.cuda.synchronize()))
timeit("+ 7 float reads", lambda: [float(v) for v in step(x_h, y_h)])
timeit("+ stacked read", lambda: torch.stack([v.detach().float().reshape(()) for v in step(x_h, y_h)]).tolist())
cs = torch.cuda.Stream()
xd2, yd2 = torch.empty_like(step.x), torch.empty_like(step.y)
def h2d_side():
    with torch.cuda.stream(cs):
        xd2.copy_(x_h, non_blocking=True); yd2.copy_(y_h, non_blocking=True)
timeit("h2d alone on side stream", h2d_side)
def overlapped():
    h2d_side(); step(); torch.cuda.synchronize()
timeit("h2d on side stream || graph + sync", overlapped)
def feed_loop():
    feed = DevicePrefetcher(((x_h, y_h) for _ in range(K)), torch.device("cuda", 0))
    for xd, yd in feed:
        out = step(xd, yd)
        torch.stack([v.detach().float().reshape(()) for v in out]).tolist()
torch.cuda.synchronize(); t0 = time.perf_counter(); feed_loop(); torch.cuda.synchronize()
print(f"{'prefetcher loop':40s} {(time.perf_counter() - t0) / K * 1e3:8.3f} ms/iter")
t0 = time.perf_counter(); feed_loop(); torch.cuda.synchronize()
print(f"{'prefetcher loop (2nd)':40s} {(time.perf_counter() - t0) / K * 1e3:8.3f} ms/iter")
