"""Resume from a checkpoint continues the run exactly: two SRGAN steps, save, load into a differently initialised model,
third step on both — identical losses and parameters (checkpoint.py; train_srgan.py:220-227 restarts from saved weights)."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fp16", [1, 0])
def test_resume_is_exact(tmp_path, fp16):
    from denoise_gan_b200 import checkpoint as CK
    from denoise_gan_b200.dataloader import synthetic_pair
    from denoise_gan_b200.srgan import SRGAN
    from denoise_gan_b200.train_srgan import train_step

    def make(seed):
        return SRGAN(SimpleNamespace(crop_size=64, scale=4, lr=1e-3, fp16=fp16, vgg=0, seed=seed))

    batches = [tuple(t.cuda() for t in synthetic_pair(2, 64, 4, step=s)) for s in range(3)]
    a = make(0)
    for s in range(2):
        train_step(a, *batches[s])
    path = str(tmp_path / "srgan.npz")
    CK.save(a, path)
    b = make(11)
    assert CK.load(b, path) == []
    assert b.iterations == a.iterations == 2
    la = [float(v) for v in train_step(a, *batches[2])]
    lb = [float(v) for v in train_step(b, *batches[2])]
    assert la == lb, (la, lb)
    for pa, pb in ((a.gen_params, b.gen_params), (a.disc_params, b.disc_params)):
        assert torch.equal(pa.theta, pb.theta) and torch.equal(pa.state, pb.state)
        assert torch.equal(pa.m, pb.m) and torch.equal(pa.v, pb.v) and torch.equal(pa.opt_state, pb.opt_state)
    # loading into a model that has already stepped re-packs the tensor-core weight copies in place
    c = make(3)
    train_step(c, *batches[0])
    CK.load(c, path)
    lc = [float(v) for v in train_step(c, *batches[2])]
    assert lc == la, (lc, la)
