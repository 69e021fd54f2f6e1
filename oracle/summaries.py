"""TEST INFRASTRUCTURE (oracle): the image summaries of the reference's training loop (train_srgan.py:27-59, 153-172) in numpy
float32, in the operation order the device kernel uses (csrc/summaries.cu).  Only tests/ import this module.

    renorm(x)      = clip((x + 1) / 2, 0, 1)                                  train_srgan.py:27-28
    autoscale(x)   = (x - min(x)) / ptp(x)                                    train_srgan.py:30-31
    tf2image(x)    = uint8(255 * renorm(x))  |  uint8(255 * autoscale(x))     train_srgan.py:33-39 (first image of the batch)
    sobel_variation(x) = sqrt((s0/4)^2 + (s1/4)^2), s = tf.image.sobel_edges(renorm(x))  (REFLECT padding)   :41-46
    high_pass_x_y / total_variation                                           :48-56
"""
from __future__ import annotations

import numpy as np

IMAGE, SQUARE, ABS, SOBEL, DX, DY, TV = range(7)
F = np.float32


def renorm(x):
    return np.clip((x.astype(F) + F(1)) * F(0.5), F(0), F(1))


def sobel_variation(x):
    """x: float32 [H, W, C] -> [H, W, C]; taps accumulated row-major, zeros skipped."""
    r = np.pad(renorm(x), ((1, 1), (1, 1), (0, 0)), mode="reflect")
    H, W = x.shape[:2]

    def tap(dy, dx):
        return r[dy:dy + H, dx:dx + W]
    # kernel 0 (d/dy): [[-1,-2,-1],[0,0,0],[1,2,1]]; kernel 1 (d/dx): its transpose
    s0 = ((((-tap(0, 0) - F(2) * tap(0, 1)) - tap(0, 2)) + tap(2, 0)) + F(2) * tap(2, 1)) + tap(2, 2)
    s1 = ((((-tap(0, 0) + tap(0, 2)) - F(2) * tap(1, 0)) + F(2) * tap(1, 2)) - tap(2, 0)) + tap(2, 2)
    a, b = s0 * F(0.25), s1 * F(0.25)
    return np.sqrt(a * a + b * b).astype(F)


def high_pass_x_y(x):
    xv = x[:, 1:, :] - x[:, :-1, :]
    yv = x[1:, :, :] - x[:-1, :, :]
    return xv[:-1, :, :], yv[:, :-1, :]


def values(kind: int, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    """The float32 image a summary kind shows, before the uint8 conversion (a, b: [H, W, C]; b: the image subtracted first)."""
    x = a.astype(F) if b is None else (a.astype(F) - b.astype(F))
    if kind == IMAGE:
        return renorm(x)
    if kind == SQUARE:
        return x * x
    if kind == ABS:
        return np.abs(x)
    if kind == SOBEL:
        return sobel_variation(x)
    dx, dy = high_pass_x_y(x)
    if kind == DX:
        return dx
    if kind == DY:
        return dy
    return np.abs(dx) + np.abs(dy)


def summary_u8(kind: int, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    v = values(kind, a, b)
    if kind != IMAGE:
        mn, mx = v.min(), v.max()
        ptp = F(mx - mn)
        v = (v - mn) / ptp if ptp > 0 else np.zeros_like(v)
    return (F(255) * v).astype(np.uint8)
