#!/usr/bin/env python
"""Times dg_umma_conv2d_dgrad_fused (skip-add + BatchNorm-backward sums in the epilogue) against the plain dgrad on the
generator trunk shape, with the debug flags that switch parts of the epilogue off (32: no transposed warp sum, 64: no
global loads of the skip gradient / BatchNorm input), and prints the epilogue timeline of CTA 0 (debug aid)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200 import _lib as L  # noqa: E402

lib = L.load(); ctx = L.ctx(0); st = L.stream_ptr()
_st0 = st
N, H, W, cin, cout, k = 16, 96, 96, 64, 64, 3
gz = torch.randn(N, H, W, cout, device="cuda").to(torch.bfloat16)
yb = torch.randn(N, H, W, cin, device="cuda").to(torch.bfloat16)
res = torch.randn(N, H, W, cin, device="cuda").to(torch.bfloat16)
dx = torch.empty(N, H, W, cin, device="cuda", dtype=torch.bfloat16)
w = torch.randn(k, k, cin, cout, device="cuda") * 0.05
pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
L.check(lib.dg_umma_pack_weights(ctx, w.data_ptr(), pk.data_ptr(), k, k, cin, cout, 1, st))
cp = L.DgConvParams(k, k, 1, 1, 1, 0, 0.0)
tg, tdx, tyb, tres = L.tensor(gz), L.tensor(dx), L.tensor(yb), L.tensor(res)
rows = lib.dg_umma_conv2d_dgrad_fused_blocks(ctx, C.byref(tg), C.byref(tdx), C.byref(cp))
part = torch.empty(rows, 2, cin, device="cuda")
sc = torch.rand(cin, device="cuda") + 0.5; sh = torch.randn(cin, device="cuda"); mu = torch.randn(cin, device="cuda")


def run(mode, flags):
    st = L.stream_ptr()
    lib.dg_debug_conv_flags(flags)
    bs = L.DgBnBwdStats(C.pointer(tyb), sc.data_ptr(), sh.data_ptr(), mu.data_ptr(), 1, 0.0, part.data_ptr())
    if mode == "plain":
        L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(tg), pk.data_ptr(), None, C.byref(tdx), C.byref(cp), st))
    else:
        L.check(lib.dg_umma_conv2d_dgrad_fused(ctx, C.byref(tg), pk.data_ptr(), C.byref(tdx), C.byref(cp), C.byref(tres) if "res" in mode else None,
                                               C.byref(bs) if "bn" in mode else None, st))
    lib.dg_debug_conv_flags(0)


MODES = [(m, 0) for m in os.environ.get("DG_PROBE_MODES", "plain,res,bn,res+bn").split(",")]
for mode, flags in MODES:
    for _ in range(3):
        run(mode, flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            run(mode, flags)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{mode:8s} flags {flags:3d}: {e0.elapsed_time(e1) / 20 * 1e3:7.2f} us per launch")

dbg = torch.zeros(256 + 148 * 4, dtype=torch.int64, device="cuda")
lib.dg_debug_conv_timeline(dbg.data_ptr())
run("res+bn", 0)
torch.cuda.synchronize()
lib.dg_debug_conv_timeline(None)
t = dbg.cpu()[:192].view(3, 16, 4)
t0 = int(t[t > 0].min())
for role, rn in enumerate(["producer", "mma", "epilogue"]):
    print(rn)
    for it in range(16):
        row = t[role, it]
        if int(row.max()) == 0:
            continue
        print("  tile", it, "  ".join(f"m{s}={int(row[s]) - t0}" for s in range(4) if int(row[s]) > 0))
life = dbg.cpu()[256:].view(148, 4)
print("first entry -> last exit", int(life[:, 3].max() - life[:, 2].min()), "ns")
