"""ORACLE (test infrastructure, never imported by the product path) — tier 1.

Direct-definition numpy restatements, with explicit loops over taps, of every operator the
reference's generator/discriminator networks call into TensorFlow/Keras for.  These are the
root of trust for `oracle/ops_torch.py` (tier 2), which in turn checks the CUDA kernels.

PARITY UNPINNED: TensorFlow is not installable in the build container and the reference
ships no golden vectors, so each function restates the *documented*
TF/Keras semantics of the call site it cites; nothing here was compared with a TF run.

All activations are NHWC.  Kernel layouts are Keras': Conv2D `[kh,kw,Cin,Cout]`,
Conv2DTranspose `[kh,kw,Cout,Cin]`, DepthwiseConv2D `[kh,kw,C,1]`.
"""
from __future__ import annotations

import numpy as np


def same_pads(in_size: int, k: int, s: int) -> tuple[int, int]:
    """TF 'SAME' padding (before, after): out = ceil(in/s), total = max((out-1)*s+k-in, 0),
    before = total//2 (srgan.py:246 k3 s2 on even input -> (0,1); pix2pix.py:115 k4 s2 -> (1,1))."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


def conv2d(x, w, b=None, stride=1, padding="same"):
    """keras.layers.Conv2D (srgan.py:154, autoencoder.py:95, pix2pix.py:115,207): cross-correlation."""
    n, h, wd, c = x.shape
    kh, kw, ci, co = w.shape
    assert ci == c
    if padding == "same":
        (pt, pb), (pl, pr) = same_pads(h, kh, stride), same_pads(wd, kw, stride)
    elif padding == "valid":
        pt = pb = pl = pr = 0
    else:
        (pt, pb), (pl, pr) = padding
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    ho = (h + pt + pb - kh) // stride + 1
    wo = (wd + pl + pr - kw) // stride + 1
    y = np.zeros((n, ho, wo, co), dtype=np.float64)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i:i + (ho - 1) * stride + 1:stride, j:j + (wo - 1) * stride + 1:stride, :]
            y += np.einsum("nhwc,co->nhwo", patch.astype(np.float64), w[i, j].astype(np.float64))
    if b is not None:
        y += b
    return y


def conv3x3_tapsum(x, w, b=None, tile_h=14, tile_w=30):
    """The tap-sum form of Conv2D(3x3, stride 1, SAME) that csrc/conv_tapsum.cu computes for the generator's image convolution
    (fsrgan.py:216-217), restated tile by tile: per 30 x 14 output tile ONE product of the zero-filled (tile + 2)-pixel halo box
    against all nine taps, D[q, tap, co] = sum_ci x[q, ci] w[tap, ci, co], then y[p, co] = b[co] + sum_tap D[p + offset(tap), tap, co]
    with the three taps of a filter row combined first (the kernel's shuffle step), then the three rows.  Test infrastructure: it
    pins the algebra (and the tiling / border handling) of the kernel against conv2d() on the CPU."""
    n, h, wd, c = x.shape
    kh, kw, ci, co = w.shape
    assert (kh, kw) == (3, 3) and ci == c
    y = np.zeros((n, h, wd, co), dtype=np.float64)
    wf = w.astype(np.float64).reshape(9, ci, co)
    for b_i in range(n):
        for h0 in range(0, h, tile_h):
            for w0 in range(0, wd, tile_w):
                box = np.zeros((tile_h + 2, tile_w + 2, c), dtype=np.float64)          # out-of-image pixels: zero (TMA fill)
                ys, xs = max(h0 - 1, 0), max(w0 - 1, 0)
                ye, xe = min(h0 + tile_h + 1, h), min(w0 + tile_w + 1, wd)
                box[ys - (h0 - 1):ye - (h0 - 1), xs - (w0 - 1):xe - (w0 - 1)] = x[b_i, ys:ye, xs:xe]
                d = np.einsum("hwc,tco->hwto", box, wf)                                  # [halo_h, halo_w, 9, co]
                rows = [sum(d[:, kx:kx + tile_w, ky * 3 + kx] for kx in range(3)) for ky in range(3)]    # horizontal taps per filter row
                out = sum(rows[ky][ky:ky + tile_h] for ky in range(3))                                 # then the three rows
                if b is not None:
                    out = out + b
                th, tw = min(tile_h, h - h0), min(tile_w, wd - w0)
                y[b_i, h0:h0 + th, w0:w0 + tw] = out[:th, :tw]
    return y


def conv2d_transpose(x, w, b=None, stride=2):
    """keras.layers.Conv2DTranspose(padding='same') (pix2pix.py:130,169): the input-gradient of the
    SAME forward conv; out = in*stride and y[s*i + k - pad_before] += x[i] * W[k]; kernel [kh,kw,Cout,Cin]."""
    n, h, wd, ci = x.shape
    kh, kw, co, ci2 = w.shape
    assert ci2 == ci
    ho, wo = h * stride, wd * stride
    pt, _ = same_pads(ho, kh, stride)
    pl, _ = same_pads(wo, kw, stride)
    full = np.zeros((n, (h - 1) * stride + kh, (wd - 1) * stride + kw, co), dtype=np.float64)
    for i in range(kh):
        for j in range(kw):
            full[:, i:i + (h - 1) * stride + 1:stride, j:j + (wd - 1) * stride + 1:stride, :] += np.einsum(
                "nhwc,oc->nhwo", x.astype(np.float64), w[i, j].astype(np.float64))
    y = full[:, pt:pt + ho, pl:pl + wo, :]
    if b is not None:
        y = y + b
    return y


def depthwise_conv2d(x, w, b=None):
    """keras.layers.DepthwiseConv2D(3, strides=1, padding='same') (fsrgan.py:149-154); kernel [kh,kw,C,1]."""
    n, h, wd, c = x.shape
    kh, kw, c2, _ = w.shape
    (pt, pb), (pl, pr) = same_pads(h, kh, 1), same_pads(wd, kw, 1)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0))).astype(np.float64)
    y = np.zeros((n, h, wd, c), dtype=np.float64)
    for i in range(kh):
        for j in range(kw):
            y += xp[:, i:i + h, j:j + wd, :] * w[i, j, :, 0]
    if b is not None:
        y += b
    return y


def batch_norm_train(x, gamma, beta, eps=1e-3):
    """BatchNormalization(training=True) (srgan.py:155): batch mean and BIASED variance over N,H,W."""
    x = x.astype(np.float64)
    mean = x.mean(axis=(0, 1, 2))
    var = x.var(axis=(0, 1, 2))
    return gamma * (x - mean) / np.sqrt(var + eps) + beta, mean, var


def batch_norm_infer(x, gamma, beta, mean, var, eps=1e-3):
    return gamma * (x - mean) / np.sqrt(var + eps) + beta


def moving_update(moving, batch, momentum):
    """Keras moving statistic: moving*m + batch*(1-m).  For the moving VARIANCE pass bessel(var, P): the fused path
    (4-D NHWC inputs) feeds the Bessel-corrected estimate."""
    return moving * momentum + batch * (1.0 - momentum)


def bessel(var, n_reduced):
    """Unbiased variance from the biased one over `n_reduced` = N*H*W samples (what Keras' fused BatchNormalization
    stores into moving_variance; `_bessels_correction_test_only = True` leaves FusedBatchNorm's correction in place)."""
    return var * (n_reduced / max(n_reduced - 1, 1))


def depth_to_space(x, block=2):
    """tf.nn.depth_to_space (srgan.py:145): out[b,2h+i,2w+j,c] = in[b,h,w,(2i+j)*C+c] (DCR order)."""
    n, h, w, c = x.shape
    co = c // (block * block)
    y = np.zeros((n, h * block, w * block, co), dtype=x.dtype)
    for i in range(block):
        for j in range(block):
            y[:, i::block, j::block, :] = x[:, :, :, (i * block + j) * co:(i * block + j + 1) * co]
    return y


def prelu(x, alpha):
    """PReLU(shared_axes=[1,2]) (srgan.py:146): max(0,x) + alpha*min(0,x), alpha per channel."""
    return np.maximum(x, 0) + alpha * np.minimum(x, 0)


def leaky_relu(x, alpha):
    """LeakyReLU(alpha=0.2) (srgan.py:250) / LeakyReLU() default alpha=0.3 (pix2pix.py:121)."""
    return np.where(x >= 0, x, alpha * x)


def relu(x):
    return np.maximum(x, 0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def max_pool2x2(x):
    """MaxPool2D(2, 2, 'same') on even sizes (autoencoder.py:110)."""
    n, h, w, c = x.shape
    return x.reshape(n, h // 2, 2, w // 2, 2, c).max(axis=(2, 4))


def upsample2x_nearest(x):
    """UpSampling2D(2, 'nearest') (autoencoder.py:122)."""
    return x.repeat(2, axis=1).repeat(2, axis=2)


def bce_from_logits(logits, target):
    """BinaryCrossentropy(from_logits=True) (train_srgan.py:71): mean(max(x,0) - x*z + log1p(exp(-|x|)))."""
    x = logits.astype(np.float64)
    return np.mean(np.maximum(x, 0) - x * target + np.log1p(np.exp(-np.abs(x))))


def bce_from_probs(p, target, eps=1e-7):
    """BinaryCrossentropy() on sigmoid outputs (train_autoencoder.py:79): clip to [eps,1-eps],
    -mean(z*log(p+eps) + (1-z)*log(1-p+eps))."""
    p = np.clip(p.astype(np.float64), eps, 1 - eps)
    return -np.mean(target * np.log(p + eps) + (1 - target) * np.log(1 - p + eps))


def mse(a, b):
    return np.mean((a.astype(np.float64) - b) ** 2)


def mae(a, b):
    return np.mean(np.abs(a.astype(np.float64) - b))


def total_variation_mean(x):
    """tf.reduce_mean(tf.image.total_variation(x)) (train_srgan.py:90): per-image SUM of |dh|+|dw|, batch mean."""
    x = x.astype(np.float64)
    dh = np.abs(x[:, 1:] - x[:, :-1]).sum(axis=(1, 2, 3))
    dw = np.abs(x[:, :, 1:] - x[:, :, :-1]).sum(axis=(1, 2, 3))
    return np.mean(dh + dw)


def adam_step(theta, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras OptimizerV2 Adam (srgan.py:49): t = iterations+1; lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    theta -= lr_t * m / (sqrt(v) + eps)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return theta - lr_t * m / (np.sqrt(v) + eps), m, v


def exponential_decay_staircase(lr0, step, decay_steps=100000, decay_rate=0.1):
    """ExponentialDecay(staircase=True) (srgan.py:35-40)."""
    return lr0 * decay_rate ** (step // decay_steps)


def dropout_keep_mask(seed: int, offset: int, numel: int) -> np.ndarray:
    """Counter-based Bernoulli(0.5) keep mask shared with the CUDA dropout kernel
    (csrc/elementwise.cu: dropout_hash).  Keras Dropout(0.5) (pix2pix.py:138) draws from TF's
    stateful RNG, which cannot be reproduced; only the distribution is kept."""
    idx = (np.arange(numel, dtype=np.uint64) + np.uint64(offset)) & np.uint64(0xFFFFFFFF)
    x = (idx.astype(np.uint32) ^ np.uint32(seed & 0xFFFFFFFF)).astype(np.uint32)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint32(16))) * np.uint32(0x7FEB352D)
        x = (x ^ (x >> np.uint32(15))) * np.uint32(0x846CA68B)
        x = x ^ (x >> np.uint32(16))
    return (x >> np.uint32(31)) == 0
