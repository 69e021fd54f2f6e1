"""GPU: dg_pair_synthesis (csrc/pairs.cu) through the C ABI against oracle/pairs.py -- bit for bit (the JPEG round trip is integer
arithmetic, the float steps use explicitly rounded operations in the oracle's order)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pairs as P  # noqa: E402


@pytest.mark.parametrize("crop,scale,quality", [(64, 4, 50), (96, 1, 25), (128, 2, 90), (384, 4, 75)])
def test_pair_synthesis_bit_exact(crop, scale, quality):
    from denoise_gan_b200 import _lib as L
    lib = L.load(); ctx = L.ctx(0); st = L.stream_ptr()
    rng = np.random.default_rng(crop + scale)
    n_src, H, W, B = 3, crop + 37, crop + 52, 5
    yy, xx = np.mgrid[0:H, 0:W]
    src = np.stack([np.clip(np.stack([127 + 90 * np.sin(xx / (5.0 + k) + c) + 70 * np.cos(yy / (6.0 + c) - k) for c in range(3)], -1) +
                            rng.normal(0, 20, (H, W, 3)), 0, 255) for k in range(n_src)]).astype(np.uint8)
    idx = rng.integers(0, n_src, B).astype(np.int32)
    top = rng.integers(0, H - crop + 1, B).astype(np.int32); left = rng.integers(0, W - crop + 1, B).astype(np.int32)
    top[0], left[0] = 0, 0; top[1], left[1] = H - crop, W - crop              # windows touching the image corners
    srcd = torch.from_numpy(src).cuda()
    co = torch.from_numpy(np.stack([idx, top, left])).cuda()
    lr = crop // scale
    x = torch.full((B, lr, lr, 3), 9.0, device="cuda"); y = torch.full((B, crop, crop, 3), 9.0, device="cuda")
    nb = lib.dg_pair_synthesis_workspace_bytes(B, crop, scale)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    L.check(lib.dg_pair_synthesis(ctx, srcd.data_ptr(), n_src, H, W, co[0].data_ptr(), co[1].data_ptr(), co[2].data_ptr(), B, crop, scale,
                                  quality, x.data_ptr(), y.data_ptr(), ws.data_ptr(), nb, st))
    torch.cuda.synchronize()
    for b in range(B):
        rx, ry = P.synth_pair(src[idx[b]], int(top[b]), int(left[b]), crop, scale, quality)
        assert np.array_equal(y[b].cpu().numpy(), ry), f"target {b}"
        gx = x[b].cpu().numpy()
        bad = np.argwhere(gx != rx)
        assert bad.size == 0, f"input {b}: {len(bad)} elements differ, first {bad[:3].tolist()}, max {np.abs(gx - rx).max()}"


def test_pair_synthesis_rejects_bad_geometry():
    from denoise_gan_b200 import _lib as L
    lib = L.load(); ctx = L.ctx(0); st = L.stream_ptr()
    src = torch.zeros(1, 64, 64, 3, dtype=torch.uint8, device="cuda")
    co = torch.zeros(3, 1, dtype=torch.int32, device="cuda")
    x = torch.empty(1, 24, 24, 3, device="cuda"); y = torch.empty(1, 48, 48, 3, device="cuda")
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    rc = lib.dg_pair_synthesis(ctx, src.data_ptr(), 1, 64, 64, co[0].data_ptr(), co[1].data_ptr(), co[2].data_ptr(), 1, 48, 2, 50,
                               x.data_ptr(), y.data_ptr(), ws.data_ptr(), ws.numel(), st)
    assert rc != 0 and b"MCU" in lib.dg_last_error()                          # 24 is not a multiple of the 16x16 MCU


def test_gpu_pair_synth_feeds_a_train_step():
    """GpuPairSynth -> train_step: the batch contract of DataLoader.dataset() (dataloader.py:188-229), batches are a pure
    function of (seed, k)."""
    from types import SimpleNamespace
    from denoise_gan_b200.dataloader import GpuPairSynth
    from denoise_gan_b200.srgan import SRGAN
    from denoise_gan_b200.train_srgan import train_step
    rng = np.random.default_rng(0)
    images = torch.from_numpy(rng.integers(0, 256, (4, 160, 200, 3), dtype=np.uint8))
    feed = GpuPairSynth(images, batch_size=2, crop_size=128, scale=4, jpeg_quality=50, seed=3)
    x0, y0 = [t.clone() for t in feed.batch(0)]
    x1, y1 = feed.batch(1)
    assert tuple(x0.shape) == (2, 32, 32, 3) and tuple(y0.shape) == (2, 128, 128, 3) and not torch.equal(y0, y1)
    xa, ya = GpuPairSynth(images, 2, 128, 4, 50, seed=3).batch(0)
    assert torch.equal(xa, x0) and torch.equal(ya, y0)
    c = feed.draw(0).numpy()
    rx, ry = P.synth_pair(images[c[0, 1]].numpy(), int(c[1, 1]), int(c[2, 1]), 128, 4, 50)
    assert np.array_equal(x0[1].cpu().numpy(), rx) and np.array_equal(y0[1].cpu().numpy(), ry)
    model = SRGAN(SimpleNamespace(crop_size=128, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0, retrain=0))
    losses = train_step(model, x0, y0)
    assert all(np.isfinite(float(v)) for v in losses)
