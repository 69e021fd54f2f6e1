"""Data parallelism across the GPUs of one box: one process per GPU, batch sharded by rank,
per-replica BatchNorm statistics (the reference's single-GPU behaviour at the per-GPU batch), and ONE
exchange step per train step — a sum all-reduce of the flat generator / discriminator gradient arenas
over NCCL (NVLink 5 / NVSwitch).  1/world_size is folded into the fused Adam kernel.
The reference has no distributed path; this is new work (§8e).

Data plane: on CUDA the all-reduce is `dg_comm_allreduce` of the C ABI (include/dg_b200.h: ncclAllReduce on the caller's
stream, NCCL resolved with dlopen), so a host that is not PyTorch reaches the same collective; torch.distributed is used
only to ship the NCCL unique id to the other ranks (and as the data plane of the CPU/gloo tests).

Overlap: gradients are exchanged in BUCKETS = contiguous ranges of the flat arena in REVERSE layer order (the backward
pass finishes the last layers first).  `poll()` is called by the engine after every tape node of the backward pass and
launches the all-reduce of every bucket whose variables have all been written, on the communication stream, ordered after
the main and the weight-gradient stream; the discriminator arena is one more bucket, started when the discriminator
backward pass ends, so it overlaps the whole generator backward pass.  Bucket size ~ bucket_mb (SRGAN: 6.07 MB -> 3
buckets; pix2pix: 218 MB -> 9).  Events per bucket let the optimiser of one network start while the other network's
buckets are still in flight."""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib


class GradAllReduce:
    def __init__(self, device, group=None, bucket_mb: float | None = None):
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=device) if self.device.type == "cuda" else None
        self.bucket_bytes = int(float(os.environ.get("DG_BUCKET_MB", bucket_mb if bucket_mb is not None else 2.5)) * (1 << 20))
        self._buckets: dict = {}        # id(pset) -> [(offset, count, frozenset(names))] in reverse layer order
        self._next: dict = {}           # id(pset) -> index of the first bucket not yet launched this step
        self._events: dict = {}         # id(pset) -> event recorded after the pset's last bucket
        self.launched = 0               # all-reduce calls issued (tests)
        self._comm = None
        if self.device.type == "cuda" and os.environ.get("DG_COMM", "1") != "0":
            self._init_nccl()

    # ------------------------------------------------------------------ NCCL through the C ABI
    def _init_nccl(self):
        lib = _lib.load()
        nbytes = int(lib.dg_comm_unique_id_bytes())
        blob = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(nbytes)
            _lib.check(lib.dg_comm_unique_id(buf))
            blob[0] = bytes(buf.raw)
        dist.broadcast_object_list(blob, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        h = C.c_void_p()
        idbuf = C.create_string_buffer(blob[0], nbytes)
        _lib.check(lib.dg_comm_init(C.byref(h), idbuf, self.rank, self.world_size, self.device.index or 0))
        self._comm, self._lib = h, lib

    def close(self):
        if self._comm is not None:
            self._lib.dg_comm_destroy(self._comm)
            self._comm = None

    def _allreduce(self, flat: torch.Tensor, offset: int, count: int):
        view = flat[offset:offset + count]
        if self._comm is not None:
            _lib.check(self._lib.dg_comm_allreduce(self._comm, view.data_ptr(), int(count), torch.cuda.current_stream().cuda_stream))
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        self.launched += 1

    # ------------------------------------------------------------------ whole-arena exchange
    def start(self, flat_grad: torch.Tensor, also_after=()):
        """All-reduce the whole arena on the communication stream, ordered after work already enqueued on the current
        stream (and on the streams in `also_after`)."""
        if self.stream is None:
            self._allreduce(flat_grad, 0, flat_grad.numel())
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        for s in also_after:
            self.stream.wait_stream(s)
        with torch.cuda.stream(self.stream):
            self._allreduce(flat_grad, 0, flat_grad.numel())

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)

    def allreduce_grads(self, model):
        """Fallback single-phase exchange (both arenas after backward)."""
        self.start(model.disc_params.grad)
        self.start(model.gen_params.grad)
        self.wait()

    # ------------------------------------------------------------------ bucketed exchange overlapped with the backward pass
    def buckets(self, pset):
        """[(offset, count, names)] covering the arena, LAST variables first, each ~ bucket_bytes."""
        b = self._buckets.get(id(pset))
        if b is None:
            b, names, hi = [], [], pset.numel
            for p in reversed(list(pset.params.values())):
                names.append(p.name)
                if (hi - p.offset) * 4 >= self.bucket_bytes:
                    b.append((p.offset, hi - p.offset, frozenset(names)))
                    names, hi = [], p.offset
            if names:
                b.append((0, hi, frozenset(names)))
            self._buckets[id(pset)] = b
        return b

    def begin(self, pset):
        self._next[id(pset)] = 0

    def poll(self, pset, written, side_streams=()):
        """Launches every not-yet-launched bucket (in order) whose variables are all in `written`."""
        b = self.buckets(pset)
        i = self._next.get(id(pset), 0)
        while i < len(b) and b[i][2] <= written:
            self._launch_bucket(pset, b[i], side_streams)
            i += 1
        self._next[id(pset)] = i

    def finish(self, pset, side_streams=()):
        """Launches whatever is left (variables that received no gradient this step stay zero) and records the arena's event."""
        b = self.buckets(pset)
        for j in range(self._next.get(id(pset), 0), len(b)):
            self._launch_bucket(pset, b[j], side_streams)
        self._next[id(pset)] = len(b)
        if self.stream is not None:
            ev = self._events.get(id(pset))
            if ev is None:
                ev = self._events[id(pset)] = torch.cuda.Event()
            ev.record(self.stream)

    def _launch_bucket(self, pset, bucket, side_streams):
        off, cnt, _ = bucket
        if self.stream is None:
            self._allreduce(pset.grad, off, cnt)
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        for s in side_streams:
            self.stream.wait_stream(s)
        with torch.cuda.stream(self.stream):
            self._allreduce(pset.grad, off, cnt)

    def wait_for(self, pset):
        """The current stream waits for the exchange of this arena only (finish() must have been called)."""
        if self.stream is None:
            return
        ev = self._events.get(id(pset))
        if ev is None:
            self.wait()
        else:
            torch.cuda.current_stream().wait_event(ev)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank's contiguous slice of a global batch (dim 0)."""
    per = x.shape[0] // world
    return x[rank * per:(rank + 1) * per]


def pin_to_numa_node(local_rank: int, world_local: int | None = None) -> str:
    """Restricts the calling process to the CPU cores of the NUMA node its GPU hangs off (PCIe root), so that the pinned staging
    buffers of the input feed and the launch thread stay node-local (round-1 N=8 runs showed the per-batch H2D copy going from
    0.56 to 1.24 ms with eight unpinned ranks).  Falls back to an even split of the visible cores when the topology cannot be
    read.  Returns a one-line description (bench.py reports it); never raises."""
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        return "affinity unsupported"
    node = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank]) if os.environ.get("CUDA_VISIBLE_DEVICES") else local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                want = set()
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    want.update(range(int(a), int(b or a) + 1))
            mine = [c for c in cores if c in want]
            if mine:
                os.sched_setaffinity(0, mine)
                return f"NUMA node {node}: {len(mine)} cores"
    except Exception:
        pass
    n = world_local or 1
    if n > 1 and len(cores) >= n:
        per = len(cores) // n
        mine = cores[local_rank * per:(local_rank + 1) * per]
        try:
            os.sched_setaffinity(0, mine)
            return f"even split: {len(mine)} cores (NUMA node unknown)"
        except Exception:
            pass
    return "unpinned"
