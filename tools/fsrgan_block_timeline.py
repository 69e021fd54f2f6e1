#!/usr/bin/env python
"""clock64 timeline of CTA 0 of one dg_fsrgan_block_infer launch at the 1080p frame size (debug aid): per tile, when the compute
warps start E (expand accumulators -> shared memory), finish it, start D (depthwise stage), finish it, and what the control
warp sees."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200 import _lib as L  # noqa: E402

lib = L.load(); ctx = L.ctx(0); st = L.stream_ptr()
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1280, 2048)
g = torch.Generator().manual_seed(0)
x = torch.randn(1, H, W, 32, generator=g).to(torch.bfloat16).cuda()
y = torch.empty_like(x)
w1 = (torch.randn(192, 32, generator=g) * 0.2).to(torch.bfloat16).cuda(); w2 = (torch.randn(32, 192, generator=g) * 0.1).to(torch.float16).cuda()
b1, wd, bd, b2 = [t.cuda() for t in (torch.randn(192, generator=g), torch.randn(9, 192, generator=g) * 0.3, torch.randn(192, generator=g), torch.randn(32, generator=g))]
tx, ty = L.tensor(x), L.tensor(y)


def run():
    L.check(lib.dg_fsrgan_block_infer(ctx, C.byref(tx), w1.data_ptr(), b1.data_ptr(), wd.data_ptr(), bd.data_ptr(), w2.data_ptr(), b2.data_ptr(), C.byref(ty), st))


for _ in range(3):
    run()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10):
    run()
ev[1].record(); torch.cuda.synchronize()
print(f"{ev[0].elapsed_time(ev[1]) * 100:.1f} us per launch ({H}x{W}, {(H // 8) * (W // 16)} tiles)")
dbg = torch.zeros(16 * 40, dtype=torch.int64, device="cuda")
lib.dg_debug_fsrgan_block_timeline(dbg.data_ptr())
run()
torch.cuda.synchronize()
lib.dg_debug_fsrgan_block_timeline(None)
t = dbg.cpu().view(16, 40)
t0 = int(t[0, 0])
names = ["E start", "E done", "D start", "D done", "ctl: E read", "ctl: P(i-1) done", "ctl: A ready"]
for it in range(16):
    print(f"tile {it:2d}  " + "  ".join(f"{names[s]}={int(t[it, s]) - t0}" for s in range(7)))
d = t[1:15]
print("mean cycles: E", float((d[:, 1] - d[:, 0]).float().mean()), " sync+wait", float((d[:, 2] - d[:, 1]).float().mean()),
      " D", float((d[:, 3] - d[:, 2]).float().mean()), " tile period", float((t[2:15, 0] - t[1:14, 0]).float().mean()))
print("per compute warp, tiles 4..7: D start and D done relative to the tile's E start")
for it in range(4, 8):
    print(f"tile {it}: D start " + " ".join(str(int(t[it, 8 + w]) - int(t[it, 0])) for w in range(12)))
    print(f"        D done  " + " ".join(str(int(t[it, 24 + w]) - int(t[it, 0])) for w in range(12)))
if os.environ.get("DG_FSRGAN_BLOCK_WS", "1") != "0":
    print("warp-specialised kernel, CTA 0, cycles after the first mark: E warp (wait acc | acc ready | tile copy free | E done), D warp 0 (top | tile full | A free | D done), "
          "P warp (top | acc ready), control (x landed | acc free), project control (A ready)")
    t0 = int(t[0][t[0] > 0].min())
    for it in range(4, 12):
        r = lambda sl: " ".join(str(int(t[it, k]) - t0) for k in sl)
        print(f"tile {it:2d}  E {r(range(0, 4))}   D {r(range(8, 12))}   P {r(range(16, 18))}   X {r(range(24, 26))}   CP {r(range(28, 29))}")
