#!/usr/bin/env python
"""Per-layer roofline table from `tools/bench_conv.py` lines: for every case the algorithmic FLOPs and bytes, which roof
bounds it (arithmetic intensity against the ridge of the measured peaks) and the fraction of that roof it reaches.
usage: python tools/roofline_table.py profiles/bench_conv_r1_final_per_layer.jsonl > profiles/roofline_per_layer_r1.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.bench_conv import CASES  # noqa: E402  (name -> kind, N, H, W, cin, cout, k, stride)

peaks = {"bf16_tflops": 1655.9, "hbm_gbs": 6438.8}
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peaks.update(json.load(open(pk)))
TF, BW = peaks["bf16_tflops"], peaks["hbm_gbs"]
ridge = TF * 1e12 / (BW * 1e9)

rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
print(f"Per-layer roofline, B200, bf16 (peaks: {TF:.1f} TFLOP/s burst, {BW:.1f} GB/s; ridge {ridge:.0f} FLOP/B).  Times are CUDA-graph")
print("replays over buffers rotated through > L2 (`tools/bench_conv.py`); wgrad rows include the partial-reduce launch and the bias")
print("gradient.  Bytes are algorithmic: input + output activations once (bf16) + fp32 dW for wgrad.\n")
print("| layer (SRGAN C3, B=16) | kind | GFLOP | MB | FLOP/B | bound | us | TFLOP/s | GB/s | fraction of its roof |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    kind, N, H, W, cin, cout, k, s = CASES[r["case"]]
    Ho, Wo = H // s, W // s
    flops = 2.0 * N * Ho * Wo * k * k * cin * cout
    byts = 2.0 * N * (H * W * cin + Ho * Wo * cout) + (4.0 * k * k * cin * cout if kind == "wgrad" else 2.0 * k * k * cin * cout)
    ai = flops / byts
    us = r["us"]
    tf, gb = flops / us / 1e6, byts / us / 1e3
    bound = "tensor" if ai > ridge else "hbm"
    frac = tf / TF if bound == "tensor" else gb / BW
    note = f" ({r['note']})" if r.get("note") else ""
    print(f"| {r['case']} {cin}->{cout} k{k} s{s} @{H}x{W}{note} | {kind} | {flops / 1e9:.2f} | {byts / 1e6:.1f} | {ai:.0f} | {bound} | {us:.1f} | "
          f"{tf:.0f} | {gb:.0f} | {frac:.2f} |")
