"""`train(model, dataset, args, writer)` (train_loop.py; reference train_srgan.py:120-176): prefetching feed + captured step
+ lagged scalar logging give the same losses, at the same iterations, as calling train_step batch by batch."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_matches_step_by_step():
    from denoise_gan_b200.dataloader import synthetic_pair
    from denoise_gan_b200.srgan import SRGAN
    from denoise_gan_b200.train_loop import SRGAN_TAGS, train
    from denoise_gan_b200.train_srgan import train_step

    def make():
        return SRGAN(SimpleNamespace(crop_size=64, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0))

    batches = [synthetic_pair(2, 64, 4, step=s) for s in range(6)]           # host batches, as a dataloader yields them
    a = make()
    ref = [[float(v) for v in train_step(a, x.cuda(), y.cuda())] for x, y in batches]
    rows = []
    b = make()
    last = train(b, iter(batches), SimpleNamespace(save_iter=2), writer=lambda tag, v, step: rows.append((tag, v, step)),
                 train_step=train_step)
    assert b.iterations == a.iterations == 6
    assert sorted({s for _, _, s in rows}) == [2, 4, 6] and len(rows) == 3 * len(SRGAN_TAGS)
    for tag, v, step in rows:
        assert v == pytest.approx(ref[step - 1][SRGAN_TAGS.index(tag)], rel=1e-5, abs=1e-7), (tag, step)
    assert list(last) == pytest.approx(ref[5], rel=1e-5, abs=1e-7)
    for pa, pb in ((a.gen_params, b.gen_params), (a.disc_params, b.disc_params)):
        assert torch.allclose(pa.theta, pb.theta, rtol=1e-5, atol=1e-7) and torch.equal(pa.opt_state, pb.opt_state)
