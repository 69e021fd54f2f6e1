"""Synthetic stand-in for the reference DataLoader (dataloader.py:188-229): yields the same batch
contract — (input [B, crop/scale, crop/scale, 3], target [B, crop, crop, 3]) float32 NHWC in [-1,1] —
without file or JPEG I/O (out of scope)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def synthetic_pair(batch: int, crop: int, scale: int, step: int = 0, rank: int = 0, device="cpu"):
    """clean y ~ U[-1,1]; degraded x = clip(avgpool_s(y) + 0.1 N(0,1), -1, 1)   (SURVEY.md §8d)."""
    gen = torch.Generator().manual_seed(1234 + step + 100003 * rank)
    y = torch.rand((batch, crop, crop, 3), generator=gen, dtype=torch.float32) * 2 - 1
    x = y
    if scale > 1:
        x = F.avg_pool2d(y.permute(0, 3, 1, 2), scale).permute(0, 2, 3, 1).contiguous()
    x = (x + 0.1 * torch.randn(x.shape, generator=gen, dtype=torch.float32)).clamp_(-1, 1)
    return x.to(device), y.to(device)


class DataLoader:
    """`DataLoader(args).dataset()` iterable with the reference's signature; `args` needs batch_size,
    crop_size, scale (and optionally steps_per_epoch)."""

    def __init__(self, args, device="cpu", rank=0):
        self.batch_size, self.crop_size = args.batch_size, args.crop_size
        self.scale = getattr(args, "scale", 1)
        self.steps = getattr(args, "steps_per_epoch", 8)
        self.device, self.rank = device, rank

    def dataset(self):
        for s in range(self.steps):
            yield synthetic_pair(self.batch_size, self.crop_size, self.scale, s, self.rank, self.device)
