"""The reference DataLoader's batch contract (dataloader.py:188-229): (input [B, crop/scale, crop/scale, 3], target
[B, crop, crop, 3]) float32 NHWC in [-1, 1].

* `synthetic_pair` / `DataLoader`: seeded synthetic pairs without file I/O (the bench and the parity tests use these).
* `GpuPairSynth`: the reference's per-sample map chain -- stack_crop, scale_image (bicubic), adjust_jpeg_quality, normalize --
  as device kernels over decoded uint8 images resident in HBM (dg_pair_synthesis, csrc/pairs.cu).  Decoding image FILES
  (tf.io.read_file / decode_jpeg, dataloader.py:39-41) stays with the caller: out of scope, SURVEY.md section 2."""
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


class GpuPairSynth:
    """Training pairs synthesised on the GPU from decoded images (dataloader.py:205-219: stack_crop -> scale_image ->
    adjust_jpeg_quality -> normalize; cache().shuffle().batch(drop_remainder=True) is the caller's index order).

    `images`: uint8 [n, H, W, 3] tensor (host or device; uploaded once -- the reference's `dataset.cache()`).  `batch(k)` draws the
    k-th batch: image indices from a seeded permutation (shuffle), crop offsets uniform as tf.image.random_crop, both from a
    torch CPU generator (a few integers per sample), everything else on the device.  Returns (input, target) device tensors that
    stay valid until the next-but-one call (two alternating output buffers, so a prefetching feed can hold one batch)."""

    def __init__(self, images: torch.Tensor, batch_size: int, crop_size: int, scale: int = 1, jpeg_quality: int = 50, seed: int = 0,
                 device=None):
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[3] == 3, "images: uint8 [n, H, W, 3]"
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.images = images.to(self.device).contiguous()
        self.n, self.H, self.W = (int(v) for v in self.images.shape[:3])
        self.batch_size, self.crop, self.scale, self.quality = int(batch_size), int(crop_size), int(scale), int(jpeg_quality)
        assert self.crop <= self.H and self.crop <= self.W and self.crop % self.scale == 0 and (self.crop // self.scale) % 16 == 0
        self.seed = seed
        self.ctx = _lib.ctx(self.device.index or 0)
        nbytes = self.lib.dg_pair_synthesis_workspace_bytes(self.batch_size, self.crop, self.scale)
        self.ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
        lr = self.crop // self.scale
        self.out = [(torch.empty(self.batch_size, lr, lr, 3, device=self.device), torch.empty(self.batch_size, self.crop, self.crop, 3, device=self.device))
                    for _ in range(2)]
        self.coords = [torch.empty(3, self.batch_size, dtype=torch.int32, device=self.device) for _ in range(2)]
        self.calls = 0

    def draw(self, k: int) -> torch.Tensor:
        """int32 [3, batch]: (image index, top, left) of batch k -- a pure function of (seed, k)."""
        gen = torch.Generator().manual_seed(self.seed * 1000003 + k)
        idx = torch.randint(0, self.n, (self.batch_size,), generator=gen, dtype=torch.int32)
        top = torch.randint(0, self.H - self.crop + 1, (self.batch_size,), generator=gen, dtype=torch.int32)
        left = torch.randint(0, self.W - self.crop + 1, (self.batch_size,), generator=gen, dtype=torch.int32)
        return torch.stack([idx, top, left])

    def batch(self, k: int):
        slot = self.calls & 1
        self.calls += 1
        coords = self.coords[slot]
        coords.copy_(self.draw(k), non_blocking=True)
        x, y = self.out[slot]
        self._lib.check(self.lib.dg_pair_synthesis(
            self.ctx, self.images.data_ptr(), self.n, self.H, self.W, coords[0].data_ptr(), coords[1].data_ptr(), coords[2].data_ptr(),
            self.batch_size, self.crop, self.scale, self.quality, x.data_ptr(), y.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
            self._lib.stream_ptr()))
        return x, y

    def dataset(self, steps: int):
        for k in range(steps):
            yield self.batch(k)
