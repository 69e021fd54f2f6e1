#!/usr/bin/env python
"""Compares the parameter gradients of one bf16 SRGAN step with the weight-gradient batching on and off (debug aid)."""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200.dataloader import synthetic_pair  # noqa: E402
from denoise_gan_b200.srgan import SRGAN  # noqa: E402
from denoise_gan_b200.train_srgan import train_step  # noqa: E402

crop, batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 2
res = {}
for b in (1, 4):
    model = SRGAN(SimpleNamespace(crop_size=crop, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0))
    model.engine.wgrad_batch = b
    x, y = synthetic_pair(batch, crop, 4, step=100)
    train_step(model, x.cuda(), y.cuda())
    torch.cuda.synchronize()
    res[b] = (model.gen_params.grads(), model.disc_params.grads())
for net in (0, 1):
    worst = []
    for name, t in res[1][net].items():
        u = res[4][net][name]
        e = ((t.double() - u.double()).norm() / t.double().norm().clamp_min(1e-30)).item()
        worst.append((e, name))
    worst.sort(reverse=True)
    print("net", "GD"[net], "worst:", [(f"{e:.2e}", n) for e, n in worst[:6]])
