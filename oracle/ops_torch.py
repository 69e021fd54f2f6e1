"""ORACLE (test infrastructure, never imported by the product path) — tier 2.

torch-CPU restatements (autograd-capable, float32 or float64) of the TF/Keras operators on the
reference's hot path, validated against the direct-definition numpy tier (`ops_np.py`) by
`tests/test_oracle_ops.py`.  Activations are NHWC; kernels are in Keras layouts.

PARITY UNPINNED: see `ops_np.py` — no TensorFlow here, no golden vectors in the reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .ops_np import same_pads


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def conv2d(x, w, b=None, stride=1, padding="same"):
    """Conv2D, kernel [kh,kw,Cin,Cout] (srgan.py:154,246; autoencoder.py:95; pix2pix.py:115,207)."""
    n, h, wd, c = x.shape
    kh, kw = w.shape[0], w.shape[1]
    if padding == "same":
        (pt, pb), (pl, pr) = same_pads(h, kh, stride), same_pads(wd, kw, stride)
    elif padding == "valid":
        pt = pb = pl = pr = 0
    else:
        (pt, pb), (pl, pr) = padding
    xn = F.pad(_nchw(x), (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1).contiguous(), b, stride=stride)
    return _nhwc(y)


def conv2d_transpose(x, w, b=None, stride=2):
    """Conv2DTranspose(padding='same'), kernel [kh,kw,Cout,Cin] (pix2pix.py:130,169)."""
    n, h, wd, ci = x.shape
    kh, kw = w.shape[0], w.shape[1]
    ho, wo = h * stride, wd * stride
    pt, _ = same_pads(ho, kh, stride)
    pl, _ = same_pads(wo, kw, stride)
    full = F.conv_transpose2d(_nchw(x), w.permute(3, 2, 0, 1).contiguous(), None, stride=stride)
    y = full[:, :, pt:pt + ho, pl:pl + wo]
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    return _nhwc(y)


def depthwise_conv2d(x, w, b=None):
    """DepthwiseConv2D(3, padding='same'), kernel [kh,kw,C,1] (fsrgan.py:149-154)."""
    n, h, wd, c = x.shape
    kh, kw = w.shape[0], w.shape[1]
    (pt, pb), (pl, pr) = same_pads(h, kh, 1), same_pads(wd, kw, 1)
    xn = F.pad(_nchw(x), (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(2, 3, 0, 1).contiguous(), b, groups=c)
    return _nhwc(y)


def batch_norm(x, p, prefix, training, state_out, momentum=0.99, eps=1e-3):
    """BatchNormalization (srgan.py:155,248; fsrgan.py:140; pix2pix.py:119).  Training: normalise with the batch mean and
    the BIASED variance; moving <- moving*m + batch*(1-m), where the moving VARIANCE takes the Bessel-corrected estimate
    var*P/(P-1): 4-D NHWC inputs go through Keras' fused path, whose FusedBatchNorm op returns the unbiased variance and
    whose `_bessels_correction_test_only = True` default leaves it in place (keras/layers/normalization.py, TF 2.1-2.4)."""
    gamma, beta = p[prefix + "/gamma"], p[prefix + "/beta"]
    if training:
        mean = x.mean(dim=(0, 1, 2))
        var = x.var(dim=(0, 1, 2), unbiased=False)
        if state_out is not None:
            mm = state_out.get(prefix + "/moving_mean", p[prefix + "/moving_mean"])
            mv = state_out.get(prefix + "/moving_variance", p[prefix + "/moving_variance"])
            state_out[prefix + "/moving_mean"] = (mm * momentum + mean.detach() * (1 - momentum)).detach()
            n = x.shape[0] * x.shape[1] * x.shape[2]
            var_u = var.detach() * (n / max(n - 1, 1))
            state_out[prefix + "/moving_variance"] = (mv * momentum + var_u * (1 - momentum)).detach()
    else:
        mean, var = p[prefix + "/moving_mean"], p[prefix + "/moving_variance"]
    return gamma * (x - mean) * torch.rsqrt(var + eps) + beta


def depth_to_space(x, block=2):
    """tf.nn.depth_to_space (srgan.py:145, fsrgan.py:188), DCR order."""
    n, h, w, c = x.shape
    co = c // (block * block)
    x = x.reshape(n, h, w, block, block, co)
    x = x.permute(0, 1, 3, 2, 4, 5)
    return x.reshape(n, h * block, w * block, co)


def prelu(x, alpha):
    return torch.clamp(x, min=0) + alpha * torch.clamp(x, max=0)


def leaky_relu(x, alpha):
    return torch.where(x >= 0, x, alpha * x)


def max_pool2x2(x):
    return _nhwc(F.max_pool2d(_nchw(x), 2, 2))


def upsample2x_nearest(x):
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


def bce_from_logits(x, target_value: float):
    """BinaryCrossentropy(from_logits=True) against a constant target (train_srgan.py:87,94-95)."""
    return (torch.clamp(x, min=0) - x * target_value + torch.log1p(torch.exp(-x.abs()))).mean()


def bce_from_probs(p, target_value: float, eps=1e-7):
    """BinaryCrossentropy() on probabilities (train_autoencoder.py:79,92,100-101)."""
    p = torch.clamp(p, eps, 1 - eps)
    return -(target_value * torch.log(p + eps) + (1 - target_value) * torch.log(1 - p + eps)).mean()


def mse(a, b):
    return ((a - b) ** 2).mean()


def mae(a, b):
    return (a - b).abs().mean()


def total_variation_mean(x):
    dh = (x[:, 1:] - x[:, :-1]).abs().sum(dim=(1, 2, 3))
    dw = (x[:, :, 1:] - x[:, :, :-1]).abs().sum(dim=(1, 2, 3))
    return (dh + dw).mean()


class KerasAdam:
    """Keras OptimizerV2 Adam restated (srgan.py:49-50, pix2pix.py:30-31); see ops_np.adam_step."""

    def __init__(self, lr, beta1=0.9, beta2=0.999, eps=1e-7, decay_steps=None, decay_rate=0.1):
        self.lr0, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.decay_steps, self.decay_rate = decay_steps, decay_rate
        self.iterations = 0
        self.m, self.v = {}, {}

    def lr(self):
        if self.decay_steps is None:
            return self.lr0
        return self.lr0 * self.decay_rate ** (self.iterations // self.decay_steps)

    @torch.no_grad()
    def apply(self, params: dict, grads: dict):
        t = self.iterations + 1
        lr_t = self.lr() * (1 - self.b2 ** t) ** 0.5 / (1 - self.b1 ** t)
        for k, g in grads.items():
            if g is None:
                continue
            th = params[k]
            m = self.m.setdefault(k, torch.zeros_like(th))
            v = self.v.setdefault(k, torch.zeros_like(th))
            m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            th.sub_(lr_t * m / (v.sqrt() + self.eps))
        self.iterations += 1
