#!/usr/bin/env python
"""Debug aid: isolates which backward op of D(fake) deviates in the fp32 model-level test."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200 import params as P
from denoise_gan_b200.dataloader import synthetic_pair
from denoise_gan_b200.srgan import SRGAN
from denoise_gan_b200.train_common import gan_step

crop, batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 4
model = SRGAN(SimpleNamespace(crop_size=crop, scale=4, lr=1e-3, fp16=0, vgg=0, seed=0))
x, y = synthetic_pair(batch, crop, 4, step=0)
E = model.engine
rec, grec = {}, {}
E.record, E.grad_record = rec, grec
snap = {k: v.cuda().double() for k, v in model.disc_params.export().items()}   # weights BEFORE the Adam update
gan_step(model, x.cuda(), y.cuda(), from_logits=True, disc_scale=1.0)
torch.cuda.synchronize()
by_out = {n.seq: n for n in E.tape}


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


for i in range(8, 1, -1):
    bn_node = by_out[rec[f"d/lrelu{i}"].seq]
    conv_out = bn_node.inputs[0]
    conv_node = by_out[conv_out.seq]
    xin = conv_node.inputs[0]
    g_act = grec[rec[f"d/lrelu{i}"].seq]          # dL/d(lrelu_i)
    g_conv = grec.get(conv_out.seq)                # dL/d(conv_i raw) = bn backward output
    g_in = grec.get(xin.seq)                       # dL/d(conv_i input) = conv dgrad output
    w = snap[f"d/conv{i}/kernel"]
    gamma = snap[f"d/bn{i}/gamma"]
    # reference BN backward in float64 from our own inputs
    xr = conv_out.t.double().requires_grad_(True)
    mean = xr.mean(dim=(0, 1, 2)); var = xr.var(dim=(0, 1, 2), unbiased=False)
    t = gamma * (xr - mean) * torch.rsqrt(var + 1e-3) + snap[f"d/bn{i}/beta"]
    out = torch.where(t >= 0, t, 0.2 * t)
    (out * g_act.double()).sum().backward()
    e_bn = rel(g_conv, xr.grad) if g_conv is not None else None
    # reference conv dgrad from our own bn-backward output
    stride = 2 if i % 2 == 0 else 1
    xi = xin.t.double().requires_grad_(True)
    H = xi.shape[1]
    out_sz = -(-H // stride)
    tot = max((out_sz - 1) * stride + 3 - H, 0)
    pb = tot // 2
    yy = F.conv2d(F.pad(xi.permute(0, 3, 1, 2), (pb, tot - pb, pb, tot - pb)), w.permute(3, 2, 0, 1).contiguous(), stride=stride)
    (yy.permute(0, 2, 3, 1) * g_conv.double()).sum().backward()
    e_dg = rel(g_in, xi.grad) if g_in is not None else None
    print(f"layer {i}: shape {tuple(conv_out.t.shape)} bn_bwd err {e_bn}  conv_dgrad err {e_dg}")
