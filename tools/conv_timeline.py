#!/usr/bin/env python
"""Prints the per-role clock64 timeline of CTA 0 of one tensor-core conv launch (debug aid)."""
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
y = torch.empty(N, H, W, cout, device="cuda", dtype=torch.bfloat16)
w = torch.randn(k, k, cin, cout, device="cuda") * 0.05
pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
L.check(lib.dg_umma_pack_weights(ctx, w.data_ptr(), pk.data_ptr(), k, k, cin, cout, 0, st))
cp = L.DgConvParams(k, k, 1, 1, 1, 0, 0.0)
tx, ty = L.tensor(x), L.tensor(y)
for _ in range(3):
    L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk.data_ptr(), None, C.byref(ty), C.byref(cp), None, st))
import os as _os
flags = int(_os.environ.get("DG_FLAGS", "0"))
lib.dg_debug_conv_flags(flags)
print("flags", flags)
dbg = torch.zeros(256 + 148 * 4, dtype=torch.int64, device="cuda")
lib.dg_debug_conv_timeline(dbg.data_ptr())
L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk.data_ptr(), None, C.byref(ty), C.byref(cp), None, st))
torch.cuda.synchronize()
lib.dg_debug_conv_timeline(None)
life = dbg.cpu()[256:].view(148, 4)
t = dbg.cpu()[:192].view(3, 16, 4)
t0 = int(t[t > 0].min())
names = {0: ["tile start", "slot free", "loads issued", ""], 1: ["tile start", "acc free", "operands landed", "mma issued+commit"],
         2: ["tile start", "acc full", "stored+released", ""]}
lib.dg_debug_conv_flags(0)
for role, rn in enumerate(["producer", "mma", "epilogue"]):
    print(rn)
    if role != 1 and not _os.environ.get('DG_ALL_ROLES'):
        continue
    for it in range(16):
        row = t[role, it]
        if int(row.max()) == 0:
            continue
        print("  tile", it, "  ".join(f"{names[role][s]}={int(row[s]) - t0}" for s in range(4) if int(row[s]) > 0))

pro = dbg.cpu()[192:198]
e0 = int(life[0, 0])
print("prologue of CTA 0 (cycles after kernel entry): barriers initialised", int(pro[0]) - e0, "| descriptors prefetched", int(pro[1]) - e0,
      "| weight loads issued", int(pro[2]) - e0, "| TMEM allocated (warp 1)", int(pro[3]) - e0, "| past the CTA-wide sync", int(pro[4]) - e0)
cyc = (life[:, 1] - life[:, 0]).float()
ns0, ns1 = life[:, 2].min().item(), life[:, 3].max().item()
print(f"CTA lifetimes: min {cyc.min().item():.0f} mean {cyc.mean().item():.0f} max {cyc.max().item():.0f} cycles; "
      f"first entry -> last exit {ns1 - ns0} ns; entry spread {life[:, 2].max().item() - ns0} ns; CTA0 entry->first mark {t0 - int(life[0, 0])} cycles")
