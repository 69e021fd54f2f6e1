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

WORKLOADS = {
    # name: (model, crop, scale, batch/GPU, step GFLOP/img G+D, +VGG)   
    "srgan_c3": ("srgan", 384, 4, 16, 125.30, 469.29),
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


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant family's most frequent launch (64->64 3x3 conv of the
# generator body, 66 of the 116 umma_conv calls of a step) from the `ncu --set full` capture in
# profiles/ncu_body_conv_r1h_raw.csv: 19.04 MB read, 0 written.  The algorithmic bytes of that launch are 37.8 MB (18.9 in +
# 18.9 out): the input is read from DRAM exactly once and the output stays in the 126 MB L2 for the next kernel.
CONV_TRAFFIC_BYTES = {"srgan_c3": 19.04e6}


def make_model(wl, fp16=1, vgg=0):
    from denoise_gan_b200.srgan import SRGAN
    _, crop, scale, _, _, _ = WORKLOADS[wl]
    return SRGAN(SimpleNamespace(crop_size=crop, scale=scale, lr=1e-3, fp16=fp16, vgg=vgg, seed=0))


def cpu_step_time(wl, batch, steps, warmup, vgg):
    """Times the oracle's restatement of the reference train_step on the host cores (float32,
    all threads): the reference's TensorFlow CPU path is not installable here (no TF wheel)."""
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.dataloader import synthetic_pair
    from oracle import ops_torch as OT
    from oracle import steps as OS
    _, crop, scale, _, _, _ = WORKLOADS[wl]
    torch.set_num_threads(os.cpu_count())
    g = P.init_srgan_generator(0, scale); d = P.init_patch_discriminator(1)
    v = P.init_vgg19_synthetic() if vgg else None
    go, do = OT.KerasAdam(1e-3, decay_steps=100000), OT.KerasAdam(5e-3, decay_steps=100000)
    times = []
    for s in range(warmup + steps):
        x, y = synthetic_pair(batch, crop, scale, step=s)
        t0 = time.perf_counter()
        OS.srgan_train_step(g, d, v, go, do, x, y)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    _, crop, scale, batch, gf, gfv = WORKLOADS[wl]
    sample_batch = args.cpu_batch
    t = cpu_step_time(wl, sample_batch, args.steps, args.warmup, args.vgg)
    val = sample_batch / t
    line = {
        "impl": "reference", "metric": "train images/sec (G+D step)", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SRGAN 4x G+D train step {crop // scale}->{crop} px (train_srgan.py:61-118)",
                   "per_gpu_batch": batch, "content_loss": "vgg19-synthetic" if args.vgg else "off (G+D step)"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"batch {sample_batch} of the {batch}-image step per timed step; torch-CPU (oneDNN) oracle restatement, "
                                   "TensorFlow not installable"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="srgan_c3", choices=list(WORKLOADS))
    ap.add_argument("--vgg", type=int, default=0, help="1: include the VGG19 content loss (synthetic weights)")
    ap.add_argument("--fp16", type=int, default=1, help="1: bf16 tensor-core path (default); 0: fp32 CUDA-core parity tier")
    ap.add_argument("--cpu-batch", type=int, default=4, help="images per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from denoise_gan_b200.dataloader import synthetic_pair
    from denoise_gan_b200.graph import GraphedStep
    from denoise_gan_b200.train_srgan import train_step

    wl = args.workload
    _, crop, scale, batch, gflop_img, gflop_img_vgg = WORKLOADS[wl]
    gf_img = gflop_img_vgg if args.vgg else gflop_img
    model = make_model(wl, args.fp16, args.vgg)
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
    # ---- end-to-end through train_step with HOST batches: H2D of the batch + D2H of the 7 losses every step
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
    n_out = len(run.out) if not args.no_graph else 7
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
    # Losses come back through a small ring of pinned host buffers: the copy of step k's seven scalars is enqueued behind step k and
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
        E = model.engine
        E.prof = []
        from denoise_gan_b200 import _lib as _L
        calls0 = _L.CALLS
        comm, model.comm = model.comm, None      # profile step is rank-0 only: no exchange, or the other ranks would be missed
        overlap, E.wgrad_overlap = E.wgrad_overlap, False   # one stream: per-kernel times are not stretched by a concurrent kernel
        # Park the GPU for ~25 ms first: Python needs ~12 ms to enqueue the eager step, longer than the GPU needs to run it, and
        # an event pair around a call would otherwise include the host's launch latency.  With the queue pre-filled every
        # event pair brackets device time only.
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.025 * 1.9e9))
        train_step(model, run.x if not args.no_graph else x_h.cuda(), run.y if not args.no_graph else y_h.cuda())
        torch.cuda.synchronize()
        model.comm = comm
        E.wgrad_overlap = overlap
        fam = {}
        for kind, flops, a, b in E.prof:
            f = fam.setdefault(kind, [0.0, 0.0, 0])
            f[0] += flops; f[1] += a.elapsed_time(b); f[2] += 1
        E.prof = None
        abi_calls = _L.CALLS - calls0
        conv_ms = sum(v[1] for v in fam.values())
        top = max(fam.items(), key=lambda kv: kv[1][1])
        kinds = {k: {"launches": v[2], "ms": round(v[1], 4), "tflops": round(v[0] / (v[1] * 1e-3) / 1e12, 1) if v[1] > 0 else None}
                 for k, v in fam.items()}
        ach = top[1][0] / (top[1][1] * 1e-3) / 1e12
        images_s = world * batch / (ms * 1e-3)
        line = {
            "metric": "train images/sec (G+D step)", "value": images_s, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.fp16 else "f32", "data": "synthetic",
            "config": {"workload": f"SRGAN 4x G+D train step {crop // scale}->{crop} px (train_srgan.py:61-118)",
                       "per_gpu_batch": batch, "global_batch": batch * world, "parallelism": f"dp{world}",
                       "content_loss": "vgg19-synthetic" if args.vgg else "off (G+D step)",
                       "l2": "no flush needed: one step streams >10 GB of activations through a 126 MB L2",
                       "cuda_graph": not args.no_graph},
            "step_tflops": images_s * gf_img / 1e3 / world,
            "step_frac_of_bf16_burst": images_s * gf_img / 1e3 / world / burst,
            "roofline": {"bound": "tensor", "kernel": top[0], "achieved": ach, "peak": sustained, "unit": "TFLOP/s",
                         "frac": ach / sustained, "traffic": CONV_TRAFFIC_BYTES.get(wl), "peak_source": peak_src + ", sustained figure (kernel timed inside a step)",
                         "conv_families": kinds, "conv_share_of_step": conv_ms / ms if args.no_graph else None},
            "e2e": {"value": world * batch / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": x_h.numel() * 4 + y_h.numel() * 4,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "h2d_ms_per_batch_gpu_idle": h2d_ms_idle},
            "gpu_launches": (kernel_nodes if kernel_nodes else abi_calls) * args.steps,
            "abi_calls_per_step": abi_calls,
            "kernel_nodes_per_step": kernel_nodes,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            t_cpu = cpu_step_time(wl, args.cpu_batch, 2, 1, args.vgg)
            line["cpu_baseline"] = {"value": args.cpu_batch / t_cpu, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"2 timed steps of batch {args.cpu_batch} (of the {batch}-image step), torch-CPU oracle restatement"}
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
