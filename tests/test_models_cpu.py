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


def test_tapsum_form_equals_conv2d():
    """oracle.ops_np.conv3x3_tapsum (the algebra and tiling of csrc/conv_tapsum.cu) == Conv2D(3x3, SAME), ragged tiles and images
    smaller than a tile included (fsrgan.py:216-217)."""
    import numpy as np
    from oracle import ops_np as ON
    rng = np.random.default_rng(3)
    for (n, h, w, co) in [(1, 5, 7, 3), (2, 14, 30, 3), (1, 31, 45, 1), (1, 29, 61, 2)]:
        x = rng.standard_normal((n, h, w, 32))
        k = rng.standard_normal((3, 3, 32, co)) * 0.1
        b = rng.standard_normal(co)
        ref = ON.conv2d(x, k, b, stride=1, padding="same")
        got = ON.conv3x3_tapsum(x, k, b)
        assert np.abs(got - ref).max() < 1e-12, (n, h, w, co)


def test_fsrgan_frame_plus_margin_equals_full_padding_in_the_crop():
    """The property FrameRunner.compute_size rests on, checked on the fp64 oracle: the Fast-SRGAN generator at inference has a
    receptive-field radius of 9.75 input pixels (fsrgan.py:99-220: stride-1 convolutions, per-pixel BatchNorm), so running the frame
    with a 16-pixel margin gives, inside the centre crop, exactly what the reference's padding to a multiple of 256 gives
    (infer_video.py:79-83,141,152) -- and a margin below the radius does not."""
    import numpy as np
    import torch
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.infer import tight_size
    from oracle import frames as F
    from oracle import models as OM
    g = {k: v.double() for k, v in P.init_fsrgan_generator(0).items()}
    gen = torch.Generator().manual_seed(5)
    for k in g:                                   # non-trivial inference statistics, biases and slopes
        if k.endswith("moving_mean"):
            g[k] = torch.randn(g[k].shape, generator=gen).double() * 0.2
        elif k.endswith("moving_variance"):
            g[k] = torch.rand(g[k].shape, generator=gen).double() * 1.5 + 0.25
        elif k.endswith(("bias", "beta", "alpha")):
            g[k] = torch.randn(g[k].shape, generator=gen).double() * 0.1
    fh, fw = 20, 36
    f = np.random.default_rng(1).integers(0, 256, size=(fh, fw, 3), dtype=np.uint8)

    def run(nh, nw):
        x = torch.from_numpy(F.video_pre(f, nh, nw))[None].double()
        y = OM.fsrgan_generator(g, x, training=False)[0].numpy()
        oy, ox = (4 * nh - 4 * fh) // 2, (4 * nw - 4 * fw) // 2
        return y[oy:oy + 4 * fh, ox:ox + 4 * fw]
    assert tight_size(fh, fw, 9.75) == (52, 68)
    ref = run(128, 128)                           # a stand-in for the reference's 256 x 256 (any padding >= the radius is equivalent)
    tight = run(52, 68)
    assert np.abs(tight - ref).max() < 1e-9
    short = run(fh + 16, fw + 16)                 # margin 8 < 9.75: the border of the computed region reaches the crop
    assert np.abs(short - ref).max() > 1e-6


def test_srgan_frame_plus_margin_equals_full_padding_in_the_crop():
    """Same property for the SRGAN generator (srgan.py:129-185): radius 35.5 input pixels from its kernel sizes, margin 40."""
    import numpy as np
    import torch
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.infer import tight_size
    from oracle import frames as F
    from oracle import models as OM
    g = {k: v.double() for k, v in P.init_srgan_generator(0, 4).items()}
    gen = torch.Generator().manual_seed(6)
    for k in g:
        if k.endswith("moving_mean"):
            g[k] = torch.randn(g[k].shape, generator=gen).double() * 0.2
        elif k.endswith("moving_variance"):
            g[k] = torch.rand(g[k].shape, generator=gen).double() * 1.5 + 0.25
    half = lambda name: (g[name].shape[0] - 1) / 2.0
    radius = (half("g/conv_in/kernel") + sum(half(f"g/res{i}/conv1/kernel") + half(f"g/res{i}/conv2/kernel") for i in range(16)) +
              half("g/conv_post/kernel") + half("g/up0/conv/kernel") + half("g/up1/conv/kernel") / 2 + half("g/conv_out/kernel") / 4)
    assert radius == 35.5
    fh, fw = 12, 20
    f = np.random.default_rng(2).integers(0, 256, size=(fh, fw, 3), dtype=np.uint8)

    def run(nh, nw):
        x = torch.from_numpy(F.video_pre(f, nh, nw))[None].double()
        y = OM.srgan_generator(g, x, training=False)[0].numpy()
        oy, ox = (4 * nh - 4 * fh) // 2, (4 * nw - 4 * fw) // 2
        return y[oy:oy + 4 * fh, ox:ox + 4 * fw]
    assert tight_size(fh, fw, radius) == (92, 100)
    ref = run(108, 116)                           # 48 pixels of padding on every side: more than the radius, as the reference's 256 is
    tight = run(92, 100)
    assert np.abs(tight - ref).max() < 1e-9
