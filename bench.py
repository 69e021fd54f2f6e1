#!/usr/bin/env python
"""Headline benchmark: SRGAN 4x generator+discriminator train step (BASELINE.json north star),
96->384 px, batch 16 per GPU, bf16 tensor-core path, synthetic noisy/clean pairs.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8 ...
    python bench.py --impl reference ...      # the reference algorithm's CPU restatement (oracle)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Every configuration BASELINE.json names (BASELINE.md section 2 for the algorithmic GFLOP per image / frame).
# kind "train": one G+D optimisation step on a batch; kind "infer": one 1080p frame through infer_video.py's per-frame path.
WORKLOADS = {
    "srgan_c3": dict(kind="train", model="srgan", crop=384, scale=4, batch=16, vgg=0, gflop=125.30, cpu_batch=16,
                     label="SRGAN 4x G+D train step 96->384 px (train_srgan.py:61-118)"),
    "srgan_c3_vgg": dict(kind="train", model="srgan", crop=384, scale=4, batch=16, vgg=1, gflop=469.29, cpu_batch=4,
                         label="SRGAN 4x G+D+VGG19 content-loss train step 96->384 px, the literal reference step (train_srgan.py:61-118, srgan.py:69-93)"),
    "ae_c2": dict(kind="train", model="autoencoder", crop=256, scale=1, batch=64, vgg=0, gflop=53.64, cpu_batch=16,
                  label="conv autoencoder G+D train step 256x256 (train_autoencoder.py:66-112)"),
    "ae_c1": dict(kind="train", model="autoencoder", crop=128, scale=1, batch=8, vgg=0, gflop=13.41, cpu_batch=8,
                  label="conv autoencoder G+D train step 128x128, batch 8: the reference's CPU-runnable case (train_autoencoder.py:66-112)"),
    "fsrgan": dict(kind="train", model="fsrgan", crop=384, scale=4, batch=16, vgg=0, gflop=32.48, cpu_batch=16,
                   label="Fast-SRGAN 4x G+D train step 96->384 px (train_fsrgan.py:61-120)"),
    "pix2pix_c4": dict(kind="train", model="pix2pix", crop=256, scale=1, batch=32, vgg=0, gflop=116.73, cpu_batch=8,
                       label="pix2pix U-Net + PatchGAN train step 256x256, two generator passes (train_pix2pix.py:33-71)"),
    "infer_fsrgan_1080p": dict(kind="infer", model="fsrgan", upscale=4, gflop=1466.0,
                               label="Fast-SRGAN 1080p frame, uint8 BGR in -> 4320x7680 uint8 RGB out (infer_video.py:138-159)"),
    "infer_ae_1080p": dict(kind="infer", model="autoencoder", upscale=1, gflop=615.7,
                           label="autoencoder 1080p frame, uint8 BGR in -> uint8 RGB out (infer_video.py:138-159)"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p["bf16_tflops"], p["bf16_tflops_sustained"], p["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons while the timed region runs: in-process NVML (initialised BEFORE the timed
    region; a query costs microseconds), falling back to spawning nvidia-smi.  Spawning nvidia-smi inside the region was
    measured to disturb it: each start-up initialises NVML against the driver and stalled the host-synchronous
    end-to-end loop by milliseconds per step on some boxes."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0])); self.max_mhz = float(out[1])
        for n, v in zip(names, out[2:]):
            if v.strip().lower().startswith("active"):
                self.reasons.add(n)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.02 if self.nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_traffic(wl, kernel):
    """`roofline.traffic`: dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from an
    `ncu --set full` capture of THIS code, as summarised by tools/ncu_extract.py into profiles/ncu_traffic.json.  It is not
    measured in the bench run itself (ncu replays kernels), so the source capture is named next to it; null without one."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture summarised for this workload"
    try:
        with open(path) as f:
            t = json.load(f)
        ent = t.get(wl, {}).get(kernel)
        if ent:
            return ent["dram_bytes_per_launch"], f"ncu capture {ent['capture']} ({ent['launch']}), not this run"
    except Exception:
        pass
    return None, "no ncu capture summarised for this workload"


def model_and_step(wl, fp16=1):
    """(model, train_step) of a training workload through the repo's public, reference-shaped API."""
    W = WORKLOADS[wl]
    ns = SimpleNamespace(crop_size=W["crop"], scale=W["scale"], lr=1e-3, fp16=fp16, vgg=W["vgg"], seed=0, retrain=0)
    if W["model"] == "srgan":
        from denoise_gan_b200.srgan import SRGAN as M
        from denoise_gan_b200.train_srgan import train_step
    elif W["model"] == "fsrgan":
        from denoise_gan_b200.fsrgan import FastSRGAN as M
        from denoise_gan_b200.train_fsrgan import train_step
    elif W["model"] == "pix2pix":
        from denoise_gan_b200.pix2pix import Pix2Pix as M
        from denoise_gan_b200.train_pix2pix import train_step
    else:
        from denoise_gan_b200.autoencoder import Autoencoder as M
        from denoise_gan_b200.train_autoencoder import train_step
    return M(ns), train_step


def cpu_step_time(wl, batch, steps, warmup):
    """Times the oracle's restatement of the reference's path on the host cores (float32, all threads): the
    reference's TensorFlow CPU implementation is not installable here (no TF wheel), so this arm is a *port*."""
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.dataloader import synthetic_pair
    from oracle import models as OM
    from oracle import ops_np as ON
    from oracle import ops_torch as OT
    from oracle import steps as OS
    W = WORKLOADS[wl]
    torch.set_num_threads(os.cpu_count())
    if W["kind"] == "infer":
        # one 256 x 256 block of the padded 1280 x 2048 frame (40 such blocks per frame), training=False
        g = P.init_fsrgan_generator(0) if W["model"] == "fsrgan" else P.init_autoencoder_generator(0)
        gen = OM.fsrgan_generator if W["model"] == "fsrgan" else OM.autoencoder_generator
        x = torch.rand(1, 256, 256, 3) * 2 - 1
        times = []
        with torch.no_grad():
            for s in range(warmup + steps):
                t0 = time.perf_counter()
                gen(g, x, False)
                if s >= warmup:
                    times.append(time.perf_counter() - t0)
        return sum(times) / len(times) * 40.0        # seconds per 1280 x 2048 frame
    crop, scale = W["crop"], W["scale"]
    v = P.init_vgg19_synthetic() if W["vgg"] else None
    if W["model"] == "pix2pix":
        g, d = P.init_pix2pix(0)
        go, do = OT.KerasAdam(2e-4, beta_1=0.5), OT.KerasAdam(2e-4, beta_1=0.5)
    else:
        g = {"srgan": lambda: P.init_srgan_generator(0, scale), "fsrgan": lambda: P.init_fsrgan_generator(0),
             "autoencoder": lambda: P.init_autoencoder_generator(0)}[W["model"]]()
        d = P.init_patch_discriminator(1)
        go, do = OT.KerasAdam(1e-3, decay_steps=100000), OT.KerasAdam(5e-3, decay_steps=100000)
    times = []
    for s in range(warmup + steps):
        x, y = synthetic_pair(batch, crop, scale, step=s)
        t0 = time.perf_counter()
        if W["model"] == "srgan":
            OS.srgan_train_step(g, d, v, go, do, x, y)
        elif W["model"] == "fsrgan":
            OS.srgan_train_step(g, d, v, go, do, x, y, fsrgan=True)
        elif W["model"] == "autoencoder":
            OS.autoencoder_train_step(g, d, v, go, do, x, y)
        else:
            shapes = [(batch, 2 << i, 2 << i, 512) for i in range(3)]
            masks = [[torch.from_numpy(ON.dropout_keep_mask(7, (p * 3 + i) << 24, int(torch.tensor(sh).prod()))).reshape(sh)
                      for i, sh in enumerate(shapes)] for p in range(2)]
            OS.pix2pix_train_step(g, d, v, go, do, x, y, masks[0], masks[1])
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def metric_of(W):
    return ("inference frames/sec (1080p, frames sharded across GPUs)", "frames/s") if W["kind"] == "infer" \
        else ("train images/sec (G+D step)", "images/s")


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation on the host cores.  TensorFlow is not installable in
    this image, so the arm times the oracle port (kind "port") on all host threads, at the workload's full per-GPU batch
    unless that would not finish in minutes (the sample is stated)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    W = WORKLOADS[wl]
    metric, unit = metric_of(W)
    if W["kind"] == "infer":
        t = cpu_step_time(wl, 1, args.steps, args.warmup)
        val, sample = 1.0 / t, "one 256x256 block of the padded 1280x2048 frame per timed step, scaled x40 to a frame"
        cfg = {"workload": W["label"]}
    else:
        sample_batch = args.cpu_batch or W["cpu_batch"]
        t = cpu_step_time(wl, sample_batch, args.steps, args.warmup)
        val = sample_batch / t
        sample = (f"batch {sample_batch} of the {W['batch']}-image step per timed step" if sample_batch != W["batch"]
                  else f"the full {W['batch']}-image step per timed step")
        cfg = {"workload": W["label"], "per_gpu_batch": W["batch"], "cpu_batch": sample_batch,
               "content_loss": "vgg19-synthetic" if W["vgg"] else "off (G+D step)"}
    line = {
        "impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": unit, "cores": os.cpu_count(), "kind": "port",
                         "sample": sample + "; torch-CPU (oneDNN) oracle restatement, TensorFlow not installable"},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def profile_families(model, step_fn, x, y):
    """Conv-kernel families of ONE extra eager step, timed live with CUDA events around every launch (single stream, GPU
    parked first so that the launch queue is pre-filled and every event pair brackets device time only).  FLOPs are the
    ALGORITHMIC ones of each launch (2 * pixels * taps * Cin * Cout with the layer's true channel counts)."""
    from denoise_gan_b200 import _lib as _L
    E = model.engine
    E.prof = []
    calls0 = _L.CALLS
    comm, model.comm = model.comm, None      # profile step is rank-0 only: no exchange, or the other ranks would be missed
    overlap, E.wgrad_overlap = E.wgrad_overlap, False   # one stream: per-kernel times are not stretched by a concurrent kernel
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.025 * 1.9e9))
    step_fn(model, x, y)
    torch.cuda.synchronize()
    model.comm = comm
    E.wgrad_overlap = overlap
    fam = {}
    for kind, flops, a, b in E.prof:
        f = fam.setdefault(kind, [0.0, 0.0, 0])
        f[0] += flops; f[1] += a.elapsed_time(b); f[2] += 1
    E.prof = None
    return fam, _L.CALLS - calls0


def roofline_block(fam, wl, ms_step, sustained, peak_src):
    kinds = {k: {"launches": v[2], "ms": round(v[1], 4), "gflop": round(v[0] / 1e9, 1),
                 "tflops": round(v[0] / (v[1] * 1e-3) / 1e12, 1) if v[1] > 0 else None,
                 "frac_of_sustained": round(v[0] / (v[1] * 1e-3) / 1e12 / sustained, 4) if v[1] > 0 else None}
             for k, v in fam.items()}
    if not fam:
        return {"bound": "tensor", "achieved": None, "peak": sustained, "unit": "TFLOP/s", "frac": None, "traffic": None}
    top = max(fam.items(), key=lambda kv: kv[1][1])
    ach = top[1][0] / (top[1][1] * 1e-3) / 1e12
    traffic, tsrc = measured_traffic(wl, top[0])
    conv_ms = sum(v[1] for v in fam.values())
    return {"bound": "tensor", "kernel": top[0], "achieved": ach, "peak": sustained, "unit": "TFLOP/s", "frac": ach / sustained,
            "traffic": traffic, "traffic_source": tsrc,
            "flops": "algorithmic (true channel counts; zero-padded channels are not counted)",
            "peak_source": peak_src + ", sustained figure (kernels timed inside a step)",
            "conv_families": kinds,
            "conv_share_of_step": round(conv_ms / ms_step, 4),
            "conv_share_note": "sum of the conv-family launch times of one single-stream eager step / the CUDA-graph step time "
                               "(weight gradients overlap the main chain on a side stream in the graph, so the share can exceed their critical-path share)"}


def run_infer(args, rank, world, local):
    """C5: 1080p frames through FrameRunner (infer_video.py's per-frame path).  `value`: frames/s with the uint8 frame already
    on the device and the uint8 result left there (FrameRunner.device_frame: the frame's launches as one CUDA-graph replay); `e2e`: FrameRunner.video() over HOST frames (pinned uint8 H2D, forward,
    uint8 D2H of every result).  Frames shard round-robin across ranks, no collective."""
    import numpy as np
    import torch.distributed as dist
    from denoise_gan_b200.infer import FrameRunner
    wl = args.workload
    W = WORKLOADS[wl]
    ns = SimpleNamespace(crop_size=256, scale=4, lr=1e-3, fp16=args.fp16, vgg=0, seed=0, retrain=0)
    if W["model"] == "fsrgan":
        from denoise_gan_b200.fsrgan import FastSRGAN as M
    else:
        from denoise_gan_b200.autoencoder import Autoencoder as M
    model = M(ns)
    runner = FrameRunner(model, upscale=W["upscale"])
    rng = np.random.default_rng(rank)
    per_rank = args.steps
    frames = [rng.integers(0, 256, size=(1080, 1920, 3), dtype=np.uint8) for _ in range(2)]
    dev = torch.from_numpy(frames[0]).cuda()
    for _ in range(max(args.warmup, 3)):
        out = runner.device_frame(dev, 0)          # the third call captures the frame as a CUDA graph, later calls replay it

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(per_rank):
        out = runner.device_frame(dev, 0)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / per_rank
    seq = [frames[k & 1] for k in range(per_rank * world)]     # the global frame list; this rank takes k = rank (mod world)
    for _ in runner.video((seq * 6)[:6 * world], rank, world):      # warm-up: both staging slots seen three times, so their frames are captured (CUDA graphs) before the timed loop
        pass
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    n_out = 0
    for _, host in runner.video(seq, rank, world, copy=False):
        n_out += 1
    e3.record()
    barrier()
    assert n_out == per_rank
    ms_e2e = e2.elapsed_time(e3) / per_rank
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank != 0:
        return
    burst, sustained, hbm, peak_src = peaks()
    E = model.engine
    E.prof = []
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.025 * 1.9e9))
    runner.video_frame(dev, to_host=False)
    torch.cuda.synchronize()
    fam = {}
    for kind, flops, a, b in E.prof:
        f = fam.setdefault(kind, [0.0, 0.0, 0])
        f[0] += flops; f[1] += a.elapsed_time(b); f[2] += 1
    E.prof = None
    metric, unit = metric_of(W)
    fps = world * 1e3 / ms
    in_b, out_b = 1080 * 1920 * 3, int(out.numel())
    # W["gflop"] counts the reference's computation (the frame padded to a multiple of 256, infer_video.py:79-83); a generator with a
    # declared receptive field runs on frame + margin (FrameRunner.compute_size), so the EXECUTED work is smaller by the pixel ratio
    from denoise_gan_b200.infer import padded_size
    (ch, cw), (ph, pw) = runner.compute_size(1080, 1920), padded_size(1080, 1920)
    gflop_exec = W["gflop"] * (ch * cw) / float(ph * pw)
    line = {
        "metric": metric, "value": fps, "unit": unit, "n_gpus": world, "steps": per_rank, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.fp16 else "f32", "data": "synthetic",
        "config": {"workload": W["label"], "frames_per_gpu": per_rank, "parallelism": f"frames round-robin over {world} GPU(s), no collective",
                   "compute_size": [ch, cw], "reference_padded_size": [ph, pw],
                   "compute_size_note": "frame + receptive-field margin when the generator declares one (cropped output identical bit for bit), else the reference's padding",
                   "l2": "no flush needed: one frame streams several GB of activations through a 126 MB L2"},
        "step_tflops": fps * gflop_exec / 1e3 / world, "step_frac_of_bf16_burst": fps * gflop_exec / 1e3 / world / burst,
        "step_gflop_executed": gflop_exec, "step_gflop_reference": W["gflop"],
        "roofline": roofline_block(fam, wl, ms, sustained, peak_src),
        "e2e": {"value": world * 1e3 / ms_e2e, "unit": unit, "h2d_bytes_per_step": in_b, "d2h_bytes_per_step": out_b, "ms_per_step": ms_e2e},
        "gpu_launches": len(E.tape) and None, "clocks": clocks,
    }
    from denoise_gan_b200 import _lib as _L
    c0 = _L.CALLS
    runner.video_frame(dev, to_host=False)
    line["gpu_launches"] = (_L.CALLS - c0) * per_rank
    line["abi_calls_per_step"] = _L.CALLS - c0
    if world == 1 and not args.no_cpu:
        t_cpu = cpu_step_time(wl, 1, 2, 1)
        line["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": unit, "cores": os.cpu_count(), "kind": "port",
                                "sample": "2 timed 256x256 blocks of the padded 1280x2048 frame, scaled x40; torch-CPU oracle restatement"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="srgan_c3", choices=list(WORKLOADS))
    ap.add_argument("--vgg", type=int, default=0, help="(legacy) 1 = --workload srgan_c3_vgg")
    ap.add_argument("--fp16", type=int, default=1, help="1: bf16 tensor-core path (default); 0: fp32 CUDA-core parity tier")
    ap.add_argument("--cpu-batch", type=int, default=0, help="images per CPU-baseline step (0: the workload's default; the full batch for srgan_c3)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the short +VGG variant measurement of the default workload")
    args = ap.parse_args()
    if args.vgg and args.workload == "srgan_c3":
        args.workload = "srgan_c3_vgg"
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("DG_BENCH_WATCHDOG_S", "420")), exit=True)   # a hung rank reports where, then dies

    def trace(msg):
        if os.environ.get("DG_BENCH_TRACE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)
    import torch.distributed as dist
    numa = None
    if world > 1:
        from denoise_gan_b200.parallel import pin_to_numa_node
        numa = pin_to_numa_node(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))     # before any pinned allocation / worker thread
        trace(f"cpu affinity: {numa}")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = args.workload
    W = WORKLOADS[wl]
    if W["kind"] == "infer":
        run_infer(args, rank, world, local)
        faulthandler.cancel_dump_traceback_later()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
        return

    from denoise_gan_b200.dataloader import synthetic_pair
    from denoise_gan_b200.graph import GraphedStep

    crop, scale, batch, gf_img = W["crop"], W["scale"], W["batch"], W["gflop"]
    model, train_step = model_and_step(wl, args.fp16)
    if world > 1:
        from denoise_gan_b200.parallel import GradAllReduce
        model.comm = GradAllReduce(model.device)
        model.world_size = world
    trace("model built")
    x_h, y_h = synthetic_pair(batch, crop, scale, step=0, rank=rank)
    x_h, y_h = x_h.pin_memory(), y_h.pin_memory()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    dot = os.path.join(ROOT, "gpurun_out", f"step_graph_rank{rank}.dot") if rank == 0 else None
    if args.no_graph:
        xd, yd = x_h.cuda(), y_h.cuda()
        run = lambda xx=None, yy=None: train_step(model, xd if xx is None else xx.cuda(non_blocking=True), yd if yy is None else yy.cuda(non_blocking=True))
        kernel_nodes = None
        for _ in range(2):
            run()
    else:
        step = GraphedStep(model, train_step, x_h, y_h, warmup=2, debug_dot=dot)
        run = step
        kernel_nodes = step.kernel_nodes
    trace("step captured")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        run()
    trace("warmup enqueued")
    # ---- device-resident throughput (inputs already in HBM)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    trace(f"device-resident timing done: {ms:.3f} ms/step")
    # ---- end-to-end through train_step with HOST batches: H2D of the batch + D2H of the losses every step
    from denoise_gan_b200.graph import DevicePrefetcher
    for xd_, yd_ in DevicePrefetcher(((x_h, y_h) for _ in range(5)), torch.device("cuda", local), depth=3):   # untimed: first-use costs of the feed path
        torch.stack([v.detach().float().reshape(()) for v in run(xd_, yd_)]).tolist()
    # diagnostic (untimed): one batch over PCIe with the GPU otherwise idle
    hx0, hx1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xs_, ys_ = torch.empty_like(x_h, device="cuda"), torch.empty_like(y_h, device="cuda")
    torch.cuda.synchronize()
    hx0.record(); xs_.copy_(x_h, non_blocking=True); ys_.copy_(y_h, non_blocking=True); hx1.record()
    torch.cuda.synchronize()
    h2d_ms_idle = hx0.elapsed_time(hx1)
    del xs_, ys_
    # everything the timed loop needs is allocated BEFORE the region (cudaMalloc / cudaHostAlloc may synchronise the device
    # and take milliseconds), and the Python garbage collector is parked: a single ~100 ms host stall inside a 40-step
    # host-driven loop showed up as +2-3 ms per step on some runs
    import gc
    n_out = len(run.out) if not args.no_graph else len(run())
    LAG = 4                                     # the host reads step k's losses after it has launched step k+LAG
    host_bufs = [torch.empty(n_out, dtype=torch.float32).pin_memory() for _ in range(LAG + 1)]
    evs = [torch.cuda.Event() for _ in range(LAG + 1)]
    feed = DevicePrefetcher(((x_h, y_h) for _ in range(args.steps)), torch.device("cuda", local), depth=3).preallocate(x_h, y_h)
    gc.collect()
    gc.disable()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    d2h = 0
    # the batch of step k+1 crosses PCIe on a copy stream while step k runs (DevicePrefetcher = the reference pipeline's
    # dataset.prefetch); every one of the K host->device copies is enqueued and completed inside the timed region
    # Losses come back through a small ring of pinned host buffers: the copy of step k's scalars is enqueued behind step k and
    # the host consumes it after it has launched step k+4 (every step's result is read on the host inside the timed
    # region; the last one before the closing event), so neither PCIe direction nor the graph launch idles the GPU.
    inflight, host_losses = [], []
    for k, (xd_, yd_) in enumerate(feed):
        out = run(xd_, yd_)
        s_ = k % (LAG + 1)
        packed = run.packed if not args.no_graph else torch.stack([v.detach().float().reshape(()) for v in out])
        host_bufs[s_].copy_(packed, non_blocking=True)
        evs[s_].record()
        inflight.append(s_)
        if len(inflight) > LAG:
            o_ = inflight.pop(0)
            evs[o_].synchronize()
            host_losses.append(host_bufs[o_].tolist())
    for o_ in inflight:
        evs[o_].synchronize()
        host_losses.append(host_bufs[o_].tolist())
    d2h = 4 * n_out
    assert len(host_losses) == args.steps and all(len(h) == n_out for h in host_losses)
    assert feed.h2d_bytes == args.steps * (x_h.numel() * 4 + y_h.numel() * 4)
    e3.record()
    barrier()
    gc.enable()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    trace("e2e timing done")
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    if rank == 0:
        burst, sustained, hbm, peak_src = peaks()
        # ---- roofline of the dominant kernel family, measured live with CUDA events (one eager step)
        fam, abi_calls = profile_families(model, train_step, run.x if not args.no_graph else x_h.cuda(), run.y if not args.no_graph else y_h.cuda())
        metric, unit = metric_of(W)
        images_s = world * batch / (ms * 1e-3)
        line = {
            "metric": metric, "value": images_s, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.fp16 else "f32", "data": "synthetic",
            "config": {"workload": W["label"], "per_gpu_batch": batch, "global_batch": batch * world, "parallelism": f"dp{world}",
                       "content_loss": "vgg19-synthetic" if W["vgg"] else "off (G+D step)",
                       "l2": "no flush needed: one step streams >10 GB of activations through a 126 MB L2",
                       "cuda_graph": not args.no_graph,
                       **({"exchange": "dg_comm_allreduce (NCCL through the C ABI), reverse-layer-order buckets launched from the backward pass"
                           if getattr(model.comm, "_comm", None) is not None else "torch.distributed all_reduce", "cpu_affinity": numa} if world > 1 else {})},
            "step_tflops": images_s * gf_img / 1e3 / world,
            "step_frac_of_bf16_burst": images_s * gf_img / 1e3 / world / burst,
            "roofline": roofline_block(fam, wl, ms, sustained, peak_src),
            "e2e": {"value": world * batch / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": x_h.numel() * 4 + y_h.numel() * 4,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "h2d_ms_per_batch_gpu_idle": h2d_ms_idle},
            "gpu_launches": (kernel_nodes if kernel_nodes else abi_calls) * args.steps,
            "abi_calls_per_step": abi_calls,
            "kernel_nodes_per_step": kernel_nodes,
            "clocks": clocks,
        }
        if wl == "srgan_c3" and world == 1 and not args.no_variants and not args.no_graph:
            # the literal reference step also runs VGG19 twice (content loss, train_srgan.py:86): reported next to the G+D headline
            try:
                del run, step
                model = None
                torch.cuda.empty_cache()
                mv, ts = model_and_step("srgan_c3_vgg", args.fp16)
                sv = GraphedStep(mv, ts, x_h, y_h, warmup=2)
                for _ in range(3):
                    sv()
                v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); v0.record()
                for _ in range(10):
                    sv()
                v1.record(); torch.cuda.synchronize()
                msv = v0.elapsed_time(v1) / 10
                gfv = WORKLOADS["srgan_c3_vgg"]["gflop"]
                line["variants"] = {"srgan_c3_vgg": {"ms_per_step": msv, "images_per_s": batch / (msv * 1e-3), "steps": 10,
                                                     "step_tflops": batch / (msv * 1e-3) * gfv / 1e3,
                                                     "step_frac_of_bf16_burst": batch / (msv * 1e-3) * gfv / 1e3 / burst,
                                                     "kernel_nodes_per_step": sv.kernel_nodes,
                                                     "note": "same step with the VGG19 content loss (synthetic weights), device-resident"}}
            except Exception as exc:      # the variant is informational: never lose the headline line to it
                line["variants"] = {"srgan_c3_vgg": {"error": repr(exc)}}
        if world == 1 and not args.no_cpu:
            cb = args.cpu_batch or W["cpu_batch"]
            t_cpu = cpu_step_time(wl, cb, 2, 1)
            line["cpu_baseline"] = {"value": cb / t_cpu, "unit": unit, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"2 timed steps of batch {cb} (the step's batch is {batch}), torch-CPU oracle restatement"}
        print(json.dumps(line), flush=True)
    trace("done")
    faulthandler.cancel_dump_traceback_later()
    if world > 1:
        # The captured step graph holds NCCL kernels; tearing the communicator down underneath it can block
        # forever.  Everything is synchronised and printed: leave without the teardown.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
