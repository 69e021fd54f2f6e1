"""Drop-in for the reference's `train_autoencoder.train_step` (train_autoencoder.py:66-112)."""
from __future__ import annotations

from .train_common import gan_step


def train_step(model, x, y):
    """x, y: [B, crop, crop, 3] float32 NHWC CUDA tensors in [-1,1].  BCE is taken on the sigmoid
    PROBABILITIES (keras BinaryCrossentropy() default, :79).  Returns (disc_loss, adv_loss, content_loss,
    mse_loss, mae_loss), the order of :112."""
    r = gan_step(model, x, y, from_logits=False, disc_scale=1.0)
    return r["disc_loss"], r["adv_loss"], r["content_loss"], r["mse_loss"], r["mae_loss"]
