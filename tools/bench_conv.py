#!/usr/bin/env python
"""Micro-benchmark of single convolution launches through the C ABI (CUDA-event timing; inputs rotate
through enough buffers to exceed the 126 MB L2).  Used to fill the per-kernel roofline table in
DESIGN.md and as the short command for `ncu --set full` captures.

    python tools/bench_conv.py [--only body_fwd] [--iters 20]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200 import _lib as L  # noqa: E402

CASES = {
    # name: (kind, N, H, W, cin, cout, k, stride)
    "body_fwd": ("fwd", 16, 96, 96, 64, 64, 3, 1),
    "body_dgrad": ("dgrad", 16, 96, 96, 64, 64, 3, 1),
    "body_wgrad": ("wgrad", 16, 96, 96, 64, 64, 3, 1),
    "body_wgrad_x4": ("wgrad4", 16, 96, 96, 64, 64, 3, 1),      # four trunk layers per launch (dg_umma_conv2d_wgrad_batch): us is per LAYER
    "up1_fwd": ("fwd", 16, 96, 96, 64, 256, 3, 1),
    "up2_fwd": ("fwd", 16, 192, 192, 64, 256, 3, 1),
    "up2_dgrad": ("dgrad", 16, 192, 192, 64, 256, 3, 1),
    "up2_wgrad": ("wgrad", 16, 192, 192, 64, 256, 3, 1),
    "d3_fwd": ("fwd", 16, 192, 192, 32, 32, 3, 1),
    "d2_fwd_s2": ("fwd", 16, 384, 384, 32, 32, 3, 2),
    "d3_wgrad": ("wgrad", 16, 192, 192, 32, 32, 3, 1),
    "vgg_256": ("fwd", 16, 96, 96, 256, 256, 3, 1),
    "up1_dgrad": ("dgrad", 16, 96, 96, 64, 256, 3, 1),
    "up1_wgrad": ("wgrad", 16, 96, 96, 64, 256, 3, 1),
    "d3_dgrad": ("dgrad", 16, 192, 192, 32, 32, 3, 1),
    "d2_dgrad_s2": ("dgrad", 16, 384, 384, 32, 32, 3, 2),
    "d2_wgrad_s2": ("wgrad", 16, 384, 384, 32, 32, 3, 2),
    "d5_fwd": ("fwd", 16, 96, 96, 32, 64, 3, 1),
    "d1_fwd": ("fwd", 16, 384, 384, 16, 32, 3, 1),
    "d1_dgrad": ("dgrad", 16, 384, 384, 16, 32, 3, 1),
    "d1_wgrad": ("wgrad", 16, 384, 384, 16, 32, 3, 1),
    "gout_fwd": ("fwd", 16, 384, 384, 64, 16, 1, 1),
    "gout_dgrad": ("dgrad", 16, 384, 384, 64, 16, 1, 1),
    "gout_wgrad": ("wgrad", 16, 384, 384, 64, 16, 1, 1),
    "gin_fwd": ("fwd", 16, 96, 96, 16, 64, 3, 1),
    "gin_wgrad": ("wgrad", 16, 96, 96, 16, 64, 3, 1),
    "d4_fwd_s2": ("fwd", 16, 192, 192, 32, 32, 3, 2),
    "d4_dgrad_s2": ("dgrad", 16, 192, 192, 32, 32, 3, 2),
    "d7_fwd": ("fwd", 16, 48, 48, 64, 64, 3, 1),
    "d7_wgrad": ("wgrad", 16, 48, 48, 64, 64, 3, 1),
    "d6_fwd_s2": ("fwd", 16, 96, 96, 64, 64, 3, 2),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sets", type=int, default=6)
    ap.add_argument("--graph", type=int, default=1, help="1: time a CUDA graph of the launches (no host launch gaps), 0: eager launches")
    args = ap.parse_args()
    lib = L.load(); ctx = L.ctx(0); st = L.stream_ptr()
    peak = 1393.4
    if os.path.exists("MEASURED_PEAKS.json"):
        peak = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"]
    for name, (kind, N, H, W, cin, cout, k, s) in CASES.items():
        if args.only and name not in args.only.split(","):
            continue
        Ho, Wo = H // s, W // s
        pt = max((Ho - 1) * s + k - H, 0) // 2
        cp = L.DgConvParams(k, k, s, pt, pt, 0, 0.0)
        xs = [torch.randn(N, H, W, cin, device="cuda").to(torch.bfloat16) for _ in range(args.sets)]
        ys = [torch.randn(N, Ho, Wo, cout, device="cuda").to(torch.bfloat16) for _ in range(args.sets)]
        w = torch.randn(k, k, cin, cout, device="cuda") * 0.05
        pk0 = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda"); pk1 = torch.empty_like(pk0)
        L.check(lib.dg_umma_pack_weights(ctx, w.data_ptr(), pk0.data_ptr(), k, k, cin, cout, 0, st))
        L.check(lib.dg_umma_pack_weights(ctx, w.data_ptr(), pk1.data_ptr(), k, k, cin, cout, 1, st))
        dw = torch.empty_like(w); db = torch.empty(cout, device="cuda")
        tx0, ty0 = L.tensor(xs[0]), L.tensor(ys[0])
        nb = lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tx0), C.byref(ty0), C.byref(cp))
        wk = torch.empty(max(nb, 16), dtype=torch.uint8, device="cuda")
        if kind == "wgrad4":
            dw4 = [torch.empty_like(w) for _ in range(4)]
            wk4 = torch.empty(max(lib.dg_umma_conv2d_wgrad_batch_workspace_bytes(4, C.byref(tx0), C.byref(ty0), C.byref(cp)), 16), dtype=torch.uint8, device="cuda")

        def run(i):
            tx, ty = L.tensor(xs[i % args.sets]), L.tensor(ys[i % args.sets])
            if kind == "fwd":
                L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk0.data_ptr(), None, C.byref(ty), C.byref(cp), None, L.stream_ptr()))
            elif kind == "dgrad":
                L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(ty), pk1.data_ptr(), None, C.byref(tx), C.byref(cp), L.stream_ptr()))
            elif kind == "wgrad4":
                tt = [(L.tensor(xs[(i + j) % args.sets]), L.tensor(ys[(i + j) % args.sets])) for j in range(4)]
                px = (C.POINTER(L.DgTensor) * 4)(*[C.pointer(t[0]) for t in tt]); pd = (C.POINTER(L.DgTensor) * 4)(*[C.pointer(t[1]) for t in tt])
                pw = (C.c_void_p * 4)(*[d.data_ptr() for d in dw4]); pa = (C.c_int * 4)(0, 0, 0, 0)
                L.check(lib.dg_umma_conv2d_wgrad_batch(ctx, 4, px, pd, pw, None, C.byref(cp), pa, wk4.data_ptr(), wk4.numel(), L.stream_ptr()))
            else:
                L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tx), C.byref(ty), dw.data_ptr(), None if os.environ.get('DG_BENCH_NO_DB') else db.data_ptr(), C.byref(cp), 0, wk.data_ptr(), nb, L.stream_ptr()))

        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run(0)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    for i in range(args.iters):
                        run(i)
            torch.cuda.current_stream().wait_stream(side)
            gr.replay(); torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                gr.replay()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (5 * args.iters)
        else:
            e0.record()
            for i in range(args.iters):
                run(i)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / args.iters
        if kind == "wgrad4":
            us /= 4
        flops = 2.0 * N * Ho * Wo * k * k * cin * cout
        byts = 2.0 * N * (H * W * cin + Ho * Wo * cout)
        print(json.dumps({"case": name, "us": round(us, 2), "tflops": round(flops / us / 1e6, 1), "frac_of_bf16_peak": round(flops / us / 1e6 / peak, 3),
                          "algo_GBs": round(byts / us / 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
