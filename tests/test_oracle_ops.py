"""CPU: tier-2 oracle (torch) against tier-1 (direct-definition numpy) on tiny shapes, plus the TF
semantics the survey flags as parity traps (Appendix B)."""
import numpy as np
import pytest
import torch

from oracle import ops_np as N
from oracle import ops_torch as T

rng = np.random.default_rng(0)


def t(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float64))


@pytest.mark.parametrize("k,s,h", [(3, 1, 6), (3, 2, 6), (4, 2, 8), (1, 1, 5), (3, 2, 7)])
def test_conv2d_same(k, s, h):
    x = rng.standard_normal((2, h, h + 2, 3)); w = rng.standard_normal((k, k, 3, 5)); b = rng.standard_normal(5)
    np.testing.assert_allclose(T.conv2d(t(x), t(w), t(b), stride=s).numpy(), N.conv2d(x, w, b, stride=s), atol=1e-10)


def test_same_pads_asymmetric():
    assert N.same_pads(6, 3, 2) == (0, 1)      # k3 s2 even input: (0,1), not PyTorch's symmetric 1
    assert N.same_pads(8, 4, 2) == (1, 1)
    assert N.same_pads(6, 3, 1) == (1, 1)
    assert N.same_pads(7, 3, 2) == (1, 1)


def test_conv2d_explicit_pad_valid():
    x = rng.standard_normal((1, 6, 6, 2)); w = rng.standard_normal((4, 4, 2, 3))
    a = T.conv2d(t(x), t(w), None, stride=1, padding=((1, 1), (1, 1))).numpy()
    np.testing.assert_allclose(a, N.conv2d(x, w, None, 1, ((1, 1), (1, 1))), atol=1e-10)
    assert a.shape == (1, 5, 5, 3)             # ZeroPadding2D + VALID k4: n+2-4+1


@pytest.mark.parametrize("k,s", [(4, 2), (3, 2), (2, 2)])
def test_conv2d_transpose(k, s):
    x = rng.standard_normal((2, 3, 4, 5)); w = rng.standard_normal((k, k, 6, 5)); b = rng.standard_normal(6)
    a = T.conv2d_transpose(t(x), t(w), t(b), stride=s).numpy()
    np.testing.assert_allclose(a, N.conv2d_transpose(x, w, b, stride=s), atol=1e-10)
    assert a.shape == (2, 3 * s, 4 * s, 6)


def test_conv2d_transpose_is_conv_input_gradient():
    # definition check: <conv(y, W), x> == <y, convT(x, W)> for the SAME stride-2 conv
    y = torch.randn(1, 8, 8, 6, dtype=torch.float64, requires_grad=True)
    w = torch.randn(4, 4, 6, 5, dtype=torch.float64); x = torch.randn(1, 4, 4, 5, dtype=torch.float64)
    (T.conv2d(y, w, None, stride=2) * x).sum().backward()
    np.testing.assert_allclose(y.grad.numpy(), T.conv2d_transpose(x, w, None, stride=2).numpy(), atol=1e-10)


def test_depthwise():
    x = rng.standard_normal((2, 5, 6, 4)); w = rng.standard_normal((3, 3, 4, 1)); b = rng.standard_normal(4)
    np.testing.assert_allclose(T.depthwise_conv2d(t(x), t(w), t(b)).numpy(), N.depthwise_conv2d(x, w, b), atol=1e-10)


def test_depth_to_space_dcr_order():
    x = rng.standard_normal((2, 3, 4, 8))
    a = T.depth_to_space(t(x)).numpy()
    np.testing.assert_allclose(a, N.depth_to_space(x))
    # out[b,2h+i,2w+j,c] = in[b,h,w,(2i+j)*C+c]; differs from torch.nn.PixelShuffle's CRD order
    assert a[1, 2 * 1 + 1, 2 * 2 + 0, 1] == x[1, 1, 2, (2 * 1 + 0) * 2 + 1]
    ps = torch.nn.functional.pixel_shuffle(t(x).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).numpy()
    assert not np.allclose(ps, a)


def test_batch_norm_train_and_moving():
    x = rng.standard_normal((3, 4, 5, 6)) * 2 + 1; g = rng.standard_normal(6); b = rng.standard_normal(6)
    p = {"bn/gamma": t(g), "bn/beta": t(b), "bn/moving_mean": torch.zeros(6, dtype=torch.float64),
         "bn/moving_variance": torch.ones(6, dtype=torch.float64)}
    st = {}
    y = T.batch_norm(t(x), p, "bn", True, st, momentum=0.8, eps=1e-3).numpy()
    yn, mean, var = N.batch_norm_train(x, g, b, 1e-3)
    np.testing.assert_allclose(y, yn, atol=1e-10)
    np.testing.assert_allclose(st["bn/moving_mean"].numpy(), N.moving_update(0.0, mean, 0.8), atol=1e-12)
    # moving variance: Bessel-corrected batch variance (Keras fused path, `_bessels_correction_test_only = True`)
    np.testing.assert_allclose(st["bn/moving_variance"].numpy(), N.moving_update(1.0, N.bessel(var, 3 * 4 * 5), 0.8), atol=1e-12)
    yi = T.batch_norm(t(x), p, "bn", False, None).numpy()
    np.testing.assert_allclose(yi, N.batch_norm_infer(x, g, b, 0.0, 1.0), atol=1e-10)


def test_batch_norm_moving_variance_small_batch_keras_formula():
    """pix2pix bottleneck maps (2x2 and 1x1 pixels at batch 1-4): the unbiased estimate differs from the biased one by
    P/(P-1), up to 33 % at P = 4; P = 1 keeps the biased value (0) instead of dividing by zero."""
    for shape in ((1, 2, 2, 5), (2, 1, 1, 5), (1, 1, 1, 5)):
        x = rng.standard_normal(shape)
        P_ = shape[0] * shape[1] * shape[2]
        p = {"bn/gamma": torch.ones(5, dtype=torch.float64), "bn/beta": torch.zeros(5, dtype=torch.float64),
             "bn/moving_mean": torch.zeros(5, dtype=torch.float64), "bn/moving_variance": torch.ones(5, dtype=torch.float64)}
        st = {}
        T.batch_norm(t(x), p, "bn", True, st, momentum=0.99, eps=1e-3)
        var_b = x.reshape(-1, 5).var(axis=0)                       # biased
        var_u = x.reshape(-1, 5).var(axis=0, ddof=1) if P_ > 1 else var_b
        np.testing.assert_allclose(st["bn/moving_variance"].numpy(), 0.99 * 1.0 + 0.01 * var_u, atol=1e-12)
        if P_ == 4:
            np.testing.assert_allclose(var_u / np.maximum(var_b, 1e-30), 4.0 / 3.0, rtol=1e-9)


def test_activations_pool_upsample():
    x = rng.standard_normal((2, 4, 6, 3)); a = rng.standard_normal(3)
    np.testing.assert_allclose(T.prelu(t(x), t(a)).numpy(), N.prelu(x, a))
    np.testing.assert_allclose(T.leaky_relu(t(x), 0.3).numpy(), N.leaky_relu(x, 0.3))
    np.testing.assert_allclose(T.max_pool2x2(t(x)).numpy(), N.max_pool2x2(x))
    np.testing.assert_allclose(T.upsample2x_nearest(t(x)).numpy(), N.upsample2x_nearest(x))


def test_losses():
    x = rng.standard_normal((2, 3, 3, 1)) * 3; y = rng.standard_normal((2, 4, 5, 3)); z = rng.standard_normal((2, 4, 5, 3))
    for tv in (0.0, 1.0):
        np.testing.assert_allclose(T.bce_from_logits(t(x), tv).item(), N.bce_from_logits(x, tv), rtol=1e-12)
        p = N.sigmoid(x)
        np.testing.assert_allclose(T.bce_from_probs(t(p), tv).item(), N.bce_from_probs(p, tv), rtol=1e-12)
        # the two BCE forms agree unless saturated (survey row 15)
        assert abs(N.bce_from_probs(p, tv) - N.bce_from_logits(x, tv)) < 1e-5
    np.testing.assert_allclose(T.mse(t(y), t(z)).item(), N.mse(y, z), rtol=1e-12)
    np.testing.assert_allclose(T.mae(t(y), t(z)).item(), N.mae(y, z), rtol=1e-12)
    np.testing.assert_allclose(T.total_variation_mean(t(y)).item(), N.total_variation_mean(y), rtol=1e-12)


def test_keras_adam_matches_definition_and_differs_from_torch():
    th = rng.standard_normal(7); g1 = rng.standard_normal(7); g2 = rng.standard_normal(7)
    opt = T.KerasAdam(1e-3, decay_steps=100000)
    p = {"w": t(th).clone()}
    opt.apply(p, {"w": t(g1)}); opt.apply(p, {"w": t(g2)})
    a, m, v = N.adam_step(th, g1, 0, 0, 1, 1e-3)
    a, m, v = N.adam_step(a, g2, m, v, 2, 1e-3)
    np.testing.assert_allclose(p["w"].numpy(), a, rtol=1e-12)
    assert N.exponential_decay_staircase(1e-3, 99999) == 1e-3
    assert abs(N.exponential_decay_staircase(1e-3, 100000) - 1e-4) < 1e-18


def test_dropout_mask_is_deterministic_and_fair():
    m = N.dropout_keep_mask(7, 0, 1 << 16)
    assert np.array_equal(m, N.dropout_keep_mask(7, 0, 1 << 16))
    assert 0.48 < m.mean() < 0.52
    assert np.array_equal(N.dropout_keep_mask(7, 100, 50), N.dropout_keep_mask(7, 0, 150)[100:])
