"""Drop-in for the reference's `autoencoder.Autoencoder` (autoencoder.py:4-229)."""
from __future__ import annotations

from . import params as P
from .nets import AutoencoderGenerator, PatchDiscriminator
from .params import ParamSet
from .srgan import AdamConfig, _GanBase


class Autoencoder(_GanBase):
    """Denoising autoencoder GAN (reference: autoencoder.py:7-61)."""

    def __init__(self, args, device=None, weights=None):
        self.hr_height = self.hr_width = args.crop_size
        self.lr_height, self.lr_width = self.hr_height, self.hr_width       # autoencoder.py:11-12 (no down-scaling)
        self.lr_shape = (self.lr_height, self.lr_width, 3)
        self.hr_shape = (self.hr_height, self.hr_width, 3)
        self.retrain = bool(getattr(args, "retrain", 0))
        self._setup(args, device)
        self.gen_optimizer = AdamConfig(args.lr, decay_steps=100000, decay_rate=0.1)       # :26-42
        self.disc_optimizer = AdamConfig(args.lr * 5, decay_steps=100000, decay_rate=0.1)
        patch = int(self.hr_height / 2 ** 4)
        self.disc_patch = (patch, patch, 1)                                 # :44-45
        self.gf = self.df = 32
        self._build_vgg(args)
        g_init = (weights or {}).get("g") or P.init_autoencoder_generator(seed=getattr(args, "seed", 0))
        d_init = (weights or {}).get("d") or P.init_patch_discriminator(seed=getattr(args, "seed", 0) + 1)
        self.gen_params = ParamSet("g", g_init, self.device)
        self.disc_params = ParamSet("d", d_init, self.device)
        self.generator = AutoencoderGenerator(self.engine, self.gen_params)
        self.discriminator = PatchDiscriminator(self.engine, self.disc_params, sigmoid=True)
