"""CUDA-graph capture of a whole train step.

The reference wraps train_step in @tf.function (train_srgan.py:60) so one traced graph is launched
per batch; here the ~10^3 kernel launches of a step are captured once into a CUDA graph and
replayed, with the batch copied into static device buffers first."""
from __future__ import annotations

import torch


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
        self.graph = torch.cuda.CUDAGraph()
        if debug_dot:
            self.graph.enable_debug_mode()
        with torch.cuda.graph(self.graph):
            self.out = step_fn(model, self.x, self.y)
        self.kernel_nodes = None
        if debug_dot:
            try:
                self.graph.debug_dump(debug_dot)
                with open(debug_dot) as f:
                    txt = f.read()
                self.kernel_nodes = txt.count("KERNEL") or None
            except Exception:
                self.kernel_nodes = None

    def __call__(self, x: torch.Tensor | None = None, y: torch.Tensor | None = None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        self.model.iterations += 1
        return self.out
