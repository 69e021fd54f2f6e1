#!/usr/bin/env python
"""Times dg_pair_synthesis (crop + bicubic + JPEG round trip + normalise on the device) for the batch shapes of the BASELINE
configurations, next to the CPU oracle on one sample (the reference does this work in tf.data map calls on host threads)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200.dataloader import GpuPairSynth  # noqa: E402
from oracle import pairs as P  # noqa: E402

rng = np.random.default_rng(0)
images = torch.from_numpy(rng.integers(0, 256, (32, 720, 1280, 3), dtype=np.uint8))
for name, batch, crop, scale in [("C3 SRGAN 96->384", 16, 384, 4), ("C2 autoencoder 256", 64, 256, 1), ("C4 pix2pix 256", 32, 256, 1)]:
    feed = GpuPairSynth(images, batch, crop, scale, 50, seed=1)
    for k in range(3):
        feed.batch(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(20):
        feed.batch(3 + k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out_bytes = batch * (crop * crop + (crop // scale) ** 2) * 3 * 4
    t0 = time.perf_counter()
    P.synth_pair(images[0].numpy(), 0, 0, crop, scale, 50)
    cpu = time.perf_counter() - t0
    print(f"{name:22s} batch {batch:3d}: {ms * 1e3:8.1f} us per batch on the device ({batch / ms * 1e3:9.0f} pairs/s, {out_bytes / ms / 1e6:6.1f} GB/s of fp32 "
          f"output); numpy oracle {cpu * 1e3:7.1f} ms per SAMPLE on one core")
