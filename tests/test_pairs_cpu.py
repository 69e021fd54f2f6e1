"""CPU: the oracle of the training-pair synthesis (oracle/pairs.py, dataloader.py:188-229) against real libjpeg -- the committed
fixture (tests/golden/jpeg_roundtrip.npz, made by tests/golden/make_jpeg_golden.py with Pillow / libjpeg-turbo) and, where Pillow
is importable, a live encode/decode -- and the analytic properties of TensorFlow's bicubic weights."""
import io
import os

import numpy as np
import pytest

from oracle import pairs as P

GOLD = os.path.join(os.path.dirname(__file__), "golden", "jpeg_roundtrip.npz")


def test_jpeg_roundtrip_matches_libjpeg_fixture_bit_for_bit():
    z = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in z.files})
    assert names == ["flat", "noise", "smooth"]
    for name in names:
        src = z[f"{name}/src"]
        for q in (10, 50, 75, 95):
            assert np.array_equal(P.jpeg_roundtrip_u8(src, q), z[f"{name}/q{q}"]), (name, q)


def test_jpeg_roundtrip_matches_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (64, 32, 3), dtype=np.uint8)
    img[16:48, 8:24] = (img[16:48, 8:24] // 8) + 100          # a low-contrast patch: small coefficients, many quantised to zero
    for q in (1, 25, 60, 90, 100):
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="JPEG", quality=q, subsampling=2)
        ref = np.array(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
        assert np.array_equal(P.jpeg_roundtrip_u8(img, q), ref), q


def test_quant_tables_follow_the_libjpeg_quality_scaling():
    l50, c50 = P.quant_tables(50)
    assert l50[0, 0] == 16 and c50[0, 0] == 17 and l50[7, 7] == 99           # quality 50 = the Annex-K tables themselves
    l100, _ = P.quant_tables(100)
    assert (l100 == 1).all()
    l10, _ = P.quant_tables(10)
    assert l10[0, 0] == 80 and l10.max() == 255                               # scale 500 %, clamped to baseline 8 bits


def test_dct_pair_is_the_identity_without_quantisation():
    rng = np.random.default_rng(0)
    b = rng.integers(-128, 128, (50, 8, 8)).astype(np.int64)
    one = np.ones((8, 8), dtype=np.int64)                                     # quality 100: every divisor is 1 (x8 for the DCT scaling)
    rec = P.idct_islow(P.quantize(P.fdct_islow(b), one)) - 128
    assert np.abs(rec - b).max() <= 2 and (rec != b).mean() < 0.5


def test_saturating_uint8_conversion():
    x = np.array([-0.5, 0.0, 1.0 / 255, 0.5, 254.9 / 255, 1.0, 1.7], dtype=np.float32)
    assert P.to_u8_saturate(x).tolist() == [0, 0, 1, 127, 255, 255, 255]


def test_bicubic_weights_keys_half_pixel():
    idx, w = P.bicubic_weights(24, 96)                                        # the SRGAN 4x pair: every tap inside the image
    assert np.allclose(w, np.array([-0.0625, 0.5625, 0.5625, -0.0625], dtype=np.float32)[None, :], atol=1e-7)
    assert (idx[:, 0] == 4 * np.arange(24)).all() and (idx[:, 3] == 4 * np.arange(24) + 3).all()
    idx, w = P.bicubic_weights(10, 25)                                        # non-integer scale: borders renormalised
    assert np.allclose(w.sum(1), 1.0, atol=1e-6) and idx.min() == 0 and idx.max() == 24
    assert w[0, 0] == 0.0 or idx[0, 0] == 0


def test_bicubic_resize_reproduces_constants_and_planes():
    x = np.full((32, 32, 3), 0.37, dtype=np.float32)
    assert np.allclose(P.bicubic_resize(x, 8, 8), 0.37, atol=1e-6)
    yy, xx = np.mgrid[0:32, 0:32].astype(np.float32)
    ramp = np.stack([xx, yy, xx + yy], -1) / 64
    out = P.bicubic_resize(ramp, 8, 8)                                        # cubic convolution reproduces linear functions
    oy, ox = np.mgrid[0:8, 0:8].astype(np.float32)
    cx, cy = (ox + 0.5) * 4 - 0.5, (oy + 0.5) * 4 - 0.5
    assert np.allclose(out, np.stack([cx, cy, cx + cy], -1) / 64, atol=1e-5)


def test_synth_pair_contract():
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (80, 100, 3), dtype=np.uint8)
    x, y = P.synth_pair(src, 5, 7, 64, 4, 50)
    assert x.shape == (16, 16, 3) and y.shape == (64, 64, 3) and x.dtype == np.float32 and y.dtype == np.float32
    assert x.min() >= -1 and x.max() <= 1 and y.min() >= -1 and y.max() <= 1
    assert np.array_equal(y, src[5:69, 7:71].astype(np.float32) * np.float32(1 / 255.0) * 2 - 1)
    yy, xx = np.mgrid[0:80, 0:100]
    smooth = np.clip(np.stack([127 + 100 * np.sin(xx / 9.0 + c) + 20 * np.cos(yy / 7.0) for c in range(3)], -1), 0, 255).astype(np.uint8)
    x1, y1 = P.synth_pair(smooth, 5, 7, 64, 1, 90)                            # denoising pair: same size, mild artefacts
    assert x1.shape == (64, 64, 3) and 0 < np.abs(x1 - y1).mean() < 0.05


def test_bicubic_kernel_and_border_rule_match_pillow_on_enlargement():
    """Pillow's BICUBIC is the same Keys a = -0.5 kernel on half-pixel centres with out-of-image taps dropped and the rest renormalised;
    it differs from tf.image.resize(antialias=False) only when SHRINKING (Pillow widens the kernel).  On enlargement the oracle must
    agree: exactly for a scale whose phases fall on TensorFlow's 1/1024 weight table (x2), to the table's quantisation otherwise."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    x = rng.random((13, 17)).astype(np.float32)
    im = Image.fromarray(x, mode="F")
    for (oh, ow), tol in (((26, 34), 1e-6), ((39, 51), 2e-3), ((20, 30), 2e-3)):
        ref = np.array(im.resize((ow, oh), Image.BICUBIC), dtype=np.float32)
        ours = P.bicubic_resize(x[..., None], oh, ow)[..., 0]
        assert np.abs(ref - ours).max() < tol, ((oh, ow), float(np.abs(ref - ours).max()))
