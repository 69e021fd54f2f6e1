"""GPU: dg_image_summary (train_srgan.py:27-59, 153-172) against oracle/summaries.py, bit for bit, and the train loop with the
image summaries switched on."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import summaries as OS  # noqa: E402


@pytest.mark.parametrize("kind", list(range(7)))
def test_image_summary_bit_exact(kind):
    from denoise_gan_b200.summaries import image_summary
    rng = np.random.default_rng(kind)
    a = (rng.random((2, 37, 29, 3), dtype=np.float32) * 2.4 - 1.2).astype(np.float32)      # beyond [-1, 1]: renorm clips
    b = (rng.random((2, 37, 29, 3), dtype=np.float32) * 2 - 1).astype(np.float32)
    for sub in (None, b):
        out = image_summary(kind, torch.from_numpy(a).cuda(), None if sub is None else torch.from_numpy(sub).cuda())
        ref = OS.summary_u8(kind, a[0], None if sub is None else sub[0])
        assert tuple(out.shape) == ref.shape
        assert np.array_equal(out.cpu().numpy(), ref), (kind, sub is not None, int(np.abs(out.cpu().numpy().astype(int) - ref).max()))


def test_constant_image_autoscale_is_zero():
    from denoise_gan_b200.summaries import ABS, image_summary
    a = torch.full((1, 8, 8, 3), 0.25, device="cuda")
    assert (image_summary(ABS, a) == 0).all()                   # ptp = 0: the reference divides by zero; zeros here


def test_train_loop_with_image_summaries_leaves_the_losses_unchanged():
    from denoise_gan_b200.dataloader import synthetic_pair
    from denoise_gan_b200.srgan import SRGAN
    from denoise_gan_b200.summaries import SUMMARIES
    from denoise_gan_b200.train_loop import train
    from denoise_gan_b200.train_srgan import train_step
    ns = SimpleNamespace(crop_size=64, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0, retrain=0, save_iter=2)
    batches = [synthetic_pair(2, 64, 4, step=s) for s in range(6)]
    runs = []
    for on in (False, True):
        model = SRGAN(ns)
        sink, scal = [], []
        last = train(model, batches, ns, writer=lambda t, v, s: scal.append((t, s)) if np.isscalar(v) or isinstance(v, float) else None,
                     train_step=train_step, image_summaries=on, image_sink=sink)
        runs.append((last, sink))
    (l0, s0), (l1, s1) = runs
    assert l0 == l1, (l0, l1)                                    # the extra inference forward does not disturb the captured step
    assert not s0 and [it for it, _ in s1] == [2, 4, 6]
    tags = [t for t, *_ in SUMMARIES]
    for _, images in s1:
        assert list(images) == tags
        assert tuple(images["Images/Generated"].shape) == (64, 64, 3) and tuple(images["Images/Input"].shape) == (16, 16, 3)
        assert tuple(images["Image Gradients/dx Target"].shape) == (63, 63, 3)
        assert images["Error/Absolute Error (MAE)"].max().item() == 255
