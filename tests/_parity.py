"""Shared parity helpers.

`noise_bound`: GAN graphs contain discontinuous derivatives (LeakyReLU, |x|, max-pool argmax) and saturating
logs, so a float32 implementation — the reference's own TensorFlow fp32 path included — differs from exact
arithmetic by an input-dependent amount.  We therefore run the ORACLE twice, in float64 (truth) and in
float32, and require   |ours - truth| <= floor + k * |oracle32 - truth|   : the CUDA path must be as close
to the truth as a straightforward fp32 evaluation of the same graph, up to a small factor."""
import torch


def relerr(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def relerr_l2(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def noise_bound(ours, truth, fp32, floor, k=3.0, metric=relerr_l2):
    """Returns (error, bound)."""
    return metric(ours, truth), floor + k * metric(fp32, truth)
