"""CUDA-graph capture of a whole train step.

The reference wraps train_step in @tf.function (train_srgan.py:60) so one traced graph is launched
per batch; here the ~10^3 kernel launches of a step are captured once into a CUDA graph and
replayed, with the batch copied into static device buffers first."""
from __future__ import annotations

import torch


class DevicePrefetcher:
    """Hands out the batches of a host `(input, target)` iterable as device tensors whose host->device copy was
    enqueued up to `depth` BATCHES AHEAD on a copy stream, so the transfers of batches k+1, k+2 run under train step k —
    the role of `dataset.prefetch(AUTOTUNE)` in the reference input pipeline (dataloader.py:204-221).  `depth + 1` device
    buffers per tensor; pinned host batches are copied asynchronously as they are, pageable ones through pinned staging.
    (depth 2: on some boxes a transfer running next to the step graph takes longer than one step.)"""

    def __init__(self, dataset, device, depth: int = 2):
        self.it = iter(dataset)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = max(1, int(depth))
        self.slots = [None] * (self.depth + 1)     # per slot: (x_dev, y_dev, x_pin, y_pin)
        self.copied = [None] * (self.depth + 1)    # per slot: event recorded after the copies that READ its pinned staging
        self.k = 0
        self.pending = []             # (x_dev, y_dev, event) of the batches whose copies are in flight, oldest first
        self.exhausted = False
        self.h2d_bytes = 0

    def _slot(self, i, x, y):
        if self.slots[i] is None or self.slots[i][0].shape != x.shape or self.slots[i][1].shape != y.shape:
            self.slots[i] = (torch.empty(x.shape, dtype=x.dtype, device=self.device), torch.empty(y.shape, dtype=y.dtype, device=self.device),
                             None if x.is_pinned() else torch.empty(x.shape, dtype=x.dtype).pin_memory(),
                             None if y.is_pinned() else torch.empty(y.shape, dtype=y.dtype).pin_memory())
        return self.slots[i]

    def preallocate(self, x, y):
        """Allocates every slot for batches shaped like (x, y) now, so that no cudaMalloc / cudaHostAlloc (both may
        synchronise the device) happens while the pipeline is running."""
        for i in range(len(self.slots)):
            self._slot(i, x, y)
        return self

    def _enqueue(self):
        try:
            x, y = next(self.it)
        except StopIteration:
            self.exhausted = True
            return
        i = self.k % len(self.slots)
        self.k += 1
        xd, yd, xp, yp = self._slot(i, x, y)
        if xp is not None:
            # pageable batch: it goes through this slot's pinned staging.  Only the copy that last READ that staging
            # (depth + 1 batches ago) has to be finished before the host overwrites it -- not the whole copy stream,
            # and certainly not the train step just enqueued on the caller's stream.
            if self.copied[i] is not None:
                self.copied[i].synchronize()
            xp.copy_(x); yp.copy_(y)
            x, y = xp, yp
        # the slot's device buffers were consumed by work already enqueued on the caller's stream (depth + 1 batches ago)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            xd.copy_(x, non_blocking=True)
            yd.copy_(y, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.copied[i] = ev
        self.h2d_bytes += x.numel() * x.element_size() + y.numel() * y.element_size()
        self.pending.append((xd, yd, ev))

    def __iter__(self):
        return self

    def __next__(self):
        # keep `depth` transfers in flight behind the batch handed out now
        while not self.exhausted and len(self.pending) < self.depth + 1:
            self._enqueue()
        if not self.pending:
            raise StopIteration
        xd, yd, ev = self.pending.pop(0)
        torch.cuda.current_stream(self.device).wait_event(ev)
        return xd, yd


class GraphedStep:
    def __init__(self, model, step_fn, x_example: torch.Tensor, y_example: torch.Tensor, warmup: int = 2, debug_dot: str | None = None):
        self.model = model
        dev = model.device
        self.x = torch.empty_like(x_example, device=dev)
        self.y = torch.empty_like(y_example, device=dev)
        self.x.copy_(x_example); self.y.copy_(y_example)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # allocates every pooled buffer, packs weights
                self.out = step_fn(model, self.x, self.y)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        try:
            self.graph = torch.cuda.CUDAGraph(keep_graph=True)      # keeps the cudaGraph_t so that its nodes can be counted
        except TypeError:
            self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = step_fn(model, self.x, self.y)
            # the step's scalar results as ONE fp32 vector (same order), so that a caller reads them with a single copy
            scal = [v for v in self.out if isinstance(v, torch.Tensor) and v.numel() == 1]
            self.packed = self._pack(scal)
        self.kernel_nodes = self._count_kernel_nodes()

    @staticmethod
    def _pack(scal):
        """The scalars as one fp32 vector: when they already are consecutive elements of one buffer (train_common.gan_step's loss
        terms) that buffer itself, without another kernel in the graph; else a stacked copy."""
        if not scal:
            return None
        first = scal[0]
        if (all(v.dtype == torch.float32 and v.device == first.device for v in scal) and
                all(v.data_ptr() == first.data_ptr() + 4 * i for i, v in enumerate(scal))):
            return torch.as_strided(first.detach(), (len(scal),), (1,))
        return torch.stack([v.detach().float().reshape(()) for v in scal])

    def _count_kernel_nodes(self):
        """Number of kernel nodes of the captured step (bench.py's `gpu_launches`), through the CUDA runtime bindings."""
        try:
            from cuda.bindings import runtime as cudart
            raw = self.graph.raw_cuda_graph()
            err, _, n = cudart.cudaGraphGetNodes(raw, 0)
            if int(err) != 0 or n == 0:
                return None
            err, nodes, n = cudart.cudaGraphGetNodes(raw, n)
            count = 0
            for nd in nodes[:n]:
                err, ty = cudart.cudaGraphNodeGetType(nd)
                if int(err) == 0 and ty == cudart.cudaGraphNodeType.cudaGraphNodeTypeKernel:
                    count += 1
            return count or None
        except Exception:
            return None

    def __call__(self, x: torch.Tensor | None = None, y: torch.Tensor | None = None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        self.model.iterations += 1
        return self.out
