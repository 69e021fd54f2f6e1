#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (C2 autoencoder, Fast-SRGAN, C4 pix2pix train steps and the C5
1080p inference path) on one GPU — same method as bench.py (CUDA-graph replays, CUDA events), one JSON line each.
bench.py stays the headline (C3); this fills the per-config table in DESIGN.md section 6.

    python tools/bench_models.py [--only ae_c2,fsrgan,pix2pix_c4,infer_fsrgan,infer_ae] [--steps 10] [--fp16 1]
"""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200.dataloader import synthetic_pair  # noqa: E402
from denoise_gan_b200.graph import GraphedStep  # noqa: E402

# name: (crop, scale, per-GPU batch, step GFLOP/img G+D)  
TRAIN = {"ae_c2": (256, 1, 64, 53.64), "fsrgan": (384, 4, 16, 32.48), "pix2pix_c4": (256, 1, 32, 116.73), "srgan_c3": (384, 4, 16, 125.30)}
INFER = {"infer_fsrgan": 1466.0, "infer_ae": 615.7}   # GFLOP per 1080p frame


def build(name, fp16, batch):
    crop, scale, _, _ = TRAIN[name]
    ns = SimpleNamespace(crop_size=crop, scale=scale, lr=1e-3, fp16=fp16, vgg=0, seed=0, retrain=0)
    if name == "ae_c2":
        from denoise_gan_b200.autoencoder import Autoencoder as M
        from denoise_gan_b200.train_autoencoder import train_step
    elif name == "fsrgan":
        from denoise_gan_b200.fsrgan import FastSRGAN as M
        from denoise_gan_b200.train_fsrgan import train_step
    elif name == "pix2pix_c4":
        from denoise_gan_b200.pix2pix import Pix2Pix as M
        from denoise_gan_b200.train_pix2pix import train_step
    else:
        from denoise_gan_b200.srgan import SRGAN as M
        from denoise_gan_b200.train_srgan import train_step
    return M(ns), train_step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--fp16", type=int, default=1)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    peak = 1655.9
    if os.path.exists("MEASURED_PEAKS.json"):
        peak = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"]
    sel = args.only.split(",") if args.only else list(TRAIN)[:3] + list(INFER)
    for name in sel:
        if name in TRAIN:
            crop, scale, batch, gf = TRAIN[name]
            batch = args.batch or batch
            model, train_step = build(name, args.fp16, batch)
            x, y = synthetic_pair(batch, crop, scale, step=0)
            if args.no_graph:
                xd, yd = x.cuda(), y.cuda()
                run = lambda: train_step(model, xd, yd)
                for _ in range(3):
                    run()
            else:
                step = GraphedStep(model, train_step, x, y, warmup=2)
                run = step
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            ips = batch / (ms * 1e-3)
            print(json.dumps({"config": name, "kind": "train step (G+D)", "per_gpu_batch": batch, "dtype": "bf16" if args.fp16 else "f32",
                              "ms_per_step": round(ms, 3), "images_per_s": round(ips, 1), "step_tflops": round(ips * gf / 1e3, 1),
                              "frac_of_bf16_burst": round(ips * gf / 1e3 / peak, 4), "cuda_graph": not args.no_graph,
                              "mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}), flush=True)
            del model, run
            torch.cuda.empty_cache()
        else:
            from denoise_gan_b200.infer import FrameRunner
            ns = SimpleNamespace(crop_size=256, scale=4, lr=1e-3, fp16=args.fp16, vgg=0, seed=0, retrain=0)
            if name == "infer_fsrgan":
                from denoise_gan_b200.fsrgan import FastSRGAN
                runner = FrameRunner(FastSRGAN(ns), upscale=4)
            else:
                from denoise_gan_b200.autoencoder import Autoencoder
                runner = FrameRunner(Autoencoder(ns), upscale=1)
            frame = np.random.default_rng(0).integers(0, 256, size=(1080, 1920, 3), dtype=np.uint8)
            for _ in range(2):
                out = runner.video_frame(frame)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = max(args.steps // 2, 3)
            for _ in range(n):
                out = runner.video_frame(frame)      # uint8 H2D, pre, forward, post, uint8 D2H: the reference's per-frame loop
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            print(json.dumps({"config": name, "kind": "1080p frame end to end (uint8 in, uint8 out)", "dtype": "bf16" if args.fp16 else "f32",
                              "ms_per_frame": round(dt * 1e3, 2), "frames_per_s": round(1 / dt, 2), "out_shape": list(out.shape),
                              "tflops": round(INFER[name] / dt / 1e3, 1), "frac_of_bf16_burst": round(INFER[name] / dt / 1e3 / peak, 4),
                              "mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}), flush=True)
            del runner
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
