"""DevicePrefetcher: batches arrive on the device in order and intact, pinned or pageable (graph.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pinned", [True, False])
def test_prefetcher_order_and_values(pinned):
    from denoise_gan_b200.graph import DevicePrefetcher

    host = []
    for k in range(5):
        x = torch.full((2, 4, 4, 3), float(k)); y = torch.arange(2 * 8 * 8 * 3, dtype=torch.float32).reshape(2, 8, 8, 3) + k
        host.append((x.pin_memory(), y.pin_memory()) if pinned else (x, y))
    feed = DevicePrefetcher(iter(host), torch.device("cuda", 0))
    seen = []
    for xd, yd in feed:
        assert xd.is_cuda and yd.is_cuda
        seen.append((xd.cpu().clone(), yd.cpu().clone()))     # consume on the current stream before the slot is reused
    assert len(seen) == 5
    for (xs, ys), (xh, yh) in zip(seen, host):
        assert torch.equal(xs, xh) and torch.equal(ys, yh)
    assert feed.h2d_bytes == sum(x.numel() * 4 + y.numel() * 4 for x, y in host)
