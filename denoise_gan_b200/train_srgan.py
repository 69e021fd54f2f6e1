"""Drop-in for the reference's `train_srgan.train_step` (train_srgan.py:61-118)."""
from __future__ import annotations

from .train_common import gan_step


def train_step(model, img_input, img_target):
    """One simultaneous generator + discriminator update.

    img_input  [B, crop/scale, crop/scale, 3] and img_target [B, crop, crop, 3]: float32 NHWC CUDA
    tensors in [-1, 1] (the reference DataLoader's batch contract, dataloader.py:161-229).
    Returns (gen_loss, adv_loss, mae_loss, mse_loss, content_loss, disc_loss, var_loss) as 0-d device
    tensors, the order of train_srgan.py:118.
    """
    r = gan_step(model, img_input, img_target, from_logits=True, disc_scale=1.0)
    return r["gen_loss"], r["adv_loss"], r["mae_loss"], r["mse_loss"], r["content_loss"], r["disc_loss"], r["var_loss"]
