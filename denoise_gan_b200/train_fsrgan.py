"""Drop-in for the reference's `train_fsrgan.train_step` (train_fsrgan.py:61-120)."""
from __future__ import annotations

from .train_common import gan_step


def train_step(model, x, y):
    """As train_srgan.train_step but disc_loss = 0.5*(valid+fake) (:96).  Returns (gen_loss, gen_loss,
    disc_loss, adv_loss, content_loss, mse_loss, mae_loss, var_loss), the order of :120."""
    r = gan_step(model, x, y, from_logits=True, disc_scale=0.5)
    return (r["gen_loss"], r["gen_loss"], r["disc_loss"], r["adv_loss"], r["content_loss"], r["mse_loss"], r["mae_loss"],
            r["var_loss"])
