#!/usr/bin/env python
"""Lists, for one model, every conv layer and whether the tensor-core kernels have a tile configuration for it
(forward / input gradient / weight gradient), with the library's reason when they do not."""
import ctypes as C
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200 import _lib as L  # noqa: E402

lib = L.load(); ctx = L.ctx(0)


def probe(name, N, H, W, cin, cout, k, stride, pt, pl, Ho, Wo):
    cin_p, cout_p = -(-cin // 16) * 16, -(-cout // 16) * 16
    x = L.DgTensor(1 << 20, L.DG_BF16, N, H, W, cin_p, cin_p, 0)
    y = L.DgTensor(1 << 20, L.DG_BF16, N, Ho, Wo, cout_p, cout_p, 0)
    cp = L.DgConvParams(k, k, stride, pt, pl, 0, 0.0)
    res = []
    for fn, a, b in ((lib.dg_umma_conv2d_fwd_supported, x, y), (lib.dg_umma_conv2d_dgrad_supported, y, x)):
        ok = bool(fn(ctx, C.byref(a), C.byref(b), C.byref(cp)))
        res.append("yes" if ok else "NO (" + lib.dg_last_error().decode()[:90] + ")")
    nb = lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(x), C.byref(y), C.byref(cp))
    res.append("yes" if nb > 0 else "NO (" + lib.dg_last_error().decode()[:90] + ")")
    print(f"{name:22s} N{N} {H}x{W} {cin}->{cout} k{k} s{stride}: fwd {res[0]} | dgrad {res[1]} | wgrad {res[2]}")


B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
# pix2pix generator down path (pix2pix.py:147-156): 4x4 stride 2 SAME
chans = [3, 64, 128, 256, 512, 512, 512, 512, 512]
h = 256
for i in range(8):
    probe(f"g/down{i}", B, h, h, chans[i], chans[i + 1], 4, 2, 1, 1, h // 2, h // 2)
    h //= 2
# up path: Conv2DTranspose = dgrad of a 4x4 s2 conv f: [N,2h,2h,Cout] -> [N,h,h,Cin]
ups = [(512, 512), (1024, 512), (1024, 512), (1024, 512), (1024, 256), (512, 128), (256, 64), (128, 3)]
h = 1
for i, (ci, co) in enumerate(ups):
    probe(f"g/up{i} (as conv {co}->{ci})", B, 2 * h, 2 * h, co, ci, 4, 2, 1, 1, h, h)
    h *= 2
probe("d/down0", B, 256, 256, 6, 64, 4, 2, 1, 1, 128, 128)
probe("d/down1", B, 128, 128, 64, 128, 4, 2, 1, 1, 64, 64)
probe("d/down2", B, 64, 64, 128, 256, 4, 2, 1, 1, 32, 32)
probe("d/conv4 (pad1 valid)", B, 32, 32, 256, 512, 4, 1, 1, 1, 31, 31)
probe("d/last (pad1 valid)", B, 31, 31, 512, 1, 4, 1, 1, 1, 30, 30)
