"""CPU: the oracle's network restatements and the product's initialisers against the reference's
architecture facts (parameter counts and output shapes derived from the reference source, SURVEY.md
§8a / Appendix A), plus golden fixtures that pin the oracle itself against regressions."""
import os

import numpy as np
import pytest
import torch

from denoise_gan_b200 import params as P
from oracle import models as OM
from oracle import ops_np as ON
from oracle import ops_torch as OT
from oracle import steps as OS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def n_trainable(p):
    return sum(v.numel() for k, v in p.items() if not k.endswith(("moving_mean", "moving_variance")))


def test_parameter_counts_match_reference_architectures():
    assert n_trainable(P.init_srgan_generator()) == 1_518_403          # srgan.py:129-185
    assert n_trainable(P.init_patch_discriminator()) == 158_689        # srgan.py:232-272
    assert n_trainable(P.init_autoencoder_generator()) == 1_267_167    # autoencoder.py:89-188
    assert n_trainable(P.init_fsrgan_generator()) == 163_043           # fsrgan.py:99-220
    g, d = P.init_pix2pix()
    assert n_trainable(g) == 54_414_979 and n_trainable(d) == 2_768_641  # pix2pix.py:144-220


def _f64(p):
    return {k: v.double() for k, v in p.items()}


def test_output_shapes():
    x = torch.rand(1, 8, 8, 3, dtype=torch.float64) * 2 - 1
    assert OM.srgan_generator(_f64(P.init_srgan_generator()), x).shape == (1, 32, 32, 3)
    assert OM.fsrgan_generator(_f64(P.init_fsrgan_generator()), x).shape == (1, 32, 32, 3)
    y = torch.rand(1, 32, 32, 3, dtype=torch.float64)
    assert OM.patch_discriminator(_f64(P.init_patch_discriminator()), y).shape == (1, 2, 2, 1)
    assert OM.autoencoder_generator(_f64(P.init_autoencoder_generator()), y).shape == (1, 32, 32, 3)
    feats = OM.vgg19_features(_f64(P.init_vgg19_synthetic()), OM.vgg_preprocess(y))
    assert feats.shape == (1, 2, 2, 512)


@pytest.mark.slow
def test_pix2pix_shapes():
    g, d = P.init_pix2pix()
    x = torch.rand(1, 256, 256, 3) * 2 - 1
    masks = [torch.ones(1, 2 ** (i + 1), 2 ** (i + 1), 512, dtype=torch.bool) for i in range(3)]
    out = OM.pix2pix_generator(g, x, True, {}, None, masks)
    assert out.shape == (1, 256, 256, 3)
    assert OM.pix2pix_discriminator(d, x, out, True, {}).shape == (1, 30, 30, 1)


def test_vgg_preprocess_caffe_mode():
    x = torch.tensor([[[[1.0, 0.0, -1.0]]]], dtype=torch.float64)     # R=255, G=127.5, B=0
    out = OM.vgg_preprocess(x)[0, 0, 0]
    np.testing.assert_allclose(out.numpy(), [0 - 103.939, 127.5 - 116.779, 255 - 123.68], atol=1e-9)


def test_train_step_return_orders_and_moving_stats():
    g, d = _f64(P.init_srgan_generator()), _f64(P.init_patch_discriminator())
    x = torch.rand(2, 8, 8, 3, dtype=torch.float64) * 2 - 1
    y = torch.rand(2, 32, 32, 3, dtype=torch.float64) * 2 - 1
    mm0 = d["d/bn2/moving_mean"].clone()
    out = OS.srgan_train_step(g, d, None, OT.KerasAdam(1e-3), OT.KerasAdam(5e-3), x, y)
    gen_loss, adv, mae, mse, content, disc, var = [float(v) for v in out]
    assert abs(gen_loss - (content + adv + mae)) < 1e-12 and content == 0.0     # train_srgan.py:91
    assert not torch.equal(d["d/bn2/moving_mean"], mm0)                          # updated (twice) by the step
    out8 = OS.srgan_train_step(_f64(P.init_fsrgan_generator()), _f64(P.init_patch_discriminator()), None, OT.KerasAdam(1e-3),
                               OT.KerasAdam(5e-3), x, y, fsrgan=True)
    assert len(out8) == 8 and float(out8[0]) == float(out8[1])                   # train_fsrgan.py:120


def test_golden_fixture_srgan_step():
    """Fixture written by tests/golden/make_golden.py from this oracle (the reference ships no vectors)."""
    z = np.load(os.path.join(GOLDEN, "srgan_step_small.npz"))
    g, d = _f64(P.init_srgan_generator(seed=0)), _f64(P.init_patch_discriminator(seed=1))
    x, y = torch.from_numpy(z["x"]).double(), torch.from_numpy(z["y"]).double()
    out = {}
    losses = OS.srgan_train_step(g, d, None, OT.KerasAdam(1e-3, decay_steps=100000), OT.KerasAdam(5e-3, decay_steps=100000), x, y, out=out)
    np.testing.assert_allclose([float(v) for v in losses], z["losses"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out["gen_output"].numpy(), z["gen_output"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g["g/conv_out/kernel"].numpy(), z["g_conv_out_kernel_after"], rtol=1e-9, atol=1e-12)
