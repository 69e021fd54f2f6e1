"""Data parallelism across the GPUs of one box: one process per GPU, batch sharded by rank,
per-replica BatchNorm statistics (the reference's single-GPU behaviour at the per-GPU batch), and ONE
exchange step per train step — a sum all-reduce of the flat generator / discriminator gradient arenas
over NCCL (NVLink 5 / NVSwitch).  The discriminator bucket is launched on a side stream as soon as
the discriminator backward pass has produced it, so it overlaps the whole generator backward pass;
the generator bucket follows the last wgrad.  1/world_size is folded into the fused Adam kernel.
The reference has no distributed path; this is new work (§8e)."""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradAllReduce:
    def __init__(self, device, group=None):
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.stream = torch.cuda.Stream(device=device) if torch.device(device).type == "cuda" else None
        self._pending = None

    def start(self, flat_grad: torch.Tensor):
        """All-reduce `flat_grad` on the communication stream, ordered after work already enqueued."""
        if self.stream is None:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)

    def allreduce_grads(self, model):
        """Fallback single-phase exchange (both buckets after backward)."""
        self.start(model.disc_params.grad)
        self.start(model.gen_params.grad)
        self.wait()


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank's contiguous slice of a global batch (dim 0)."""
    per = x.shape[0] // world
    return x[rank * per:(rank + 1) * per]
