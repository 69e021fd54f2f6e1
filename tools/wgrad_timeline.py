#!/usr/bin/env python
"""Prints the clock64 timeline of CTA 0 of one tensor-core weight-gradient launch (debug aid)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200 import _lib as L  # noqa: E402

lib = L.load(); ctx = L.ctx(0); st = L.stream_ptr()
N, H, W, cin, cout, k = 16, 96, 96, 64, 64, 3
if len(sys.argv) > 1:
    N, H, W, cin, cout = [int(v) for v in sys.argv[1:6]]
x = torch.randn(N, H, W, cin, device="cuda").to(torch.bfloat16)
dy = torch.randn(N, H, W, cout, device="cuda").to(torch.bfloat16)
dw = torch.empty(k, k, cin, cout, device="cuda")
cp = L.DgConvParams(k, k, 1, 1, 1, 0, 0.0)
tx, ty = L.tensor(x), L.tensor(dy)
nb = lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(ty), C.byref(cp))
ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
run = lambda: L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tx), C.byref(ty), dw.data_ptr(), None, C.byref(cp), 0, ws.data_ptr(), nb, st))
for _ in range(3):
    run()
dbg = torch.zeros(3 * 16 * 4, dtype=torch.int64, device="cuda")
lib.dg_debug_wgrad_timeline(dbg.data_ptr())
run()
torch.cuda.synchronize()
lib.dg_debug_wgrad_timeline(None)
t = dbg.cpu().view(3, 16, 4)
t0 = int(t[t > 0].min())
print("mma warp: tile start / operands landed / issued+commit")
for i in range(16):
    if t[1, i, 0] > 0:
        print(f"  tile {i}: " + " ".join(str(int(v) - t0) for v in t[1, i, :3]))
print("epilogue: wait start / accumulators done / dump finished:", " ".join(str(int(v) - t0) for v in t[2, 0, :3]))
