"""Drop-in for the reference's `fsrgan.FastSRGAN` (fsrgan.py:5-258)."""
from __future__ import annotations

from . import params as P
from .nets import FastSRGANGenerator, PatchDiscriminator
from .params import ParamSet
from .srgan import AdamConfig, _GanBase


class FastSRGAN(_GanBase):
    """Fast-SRGAN: MobileNet-style generator (reference: fsrgan.py:8-70)."""

    def __init__(self, args, device=None, weights=None):
        self.scale = args.scale
        self.hr_height = self.hr_width = args.crop_size
        self.lr_height = self.hr_height // self.scale
        self.lr_width = self.hr_width // self.scale
        self.lr_shape = (self.lr_height, self.lr_width, 3)
        self.hr_shape = (self.hr_height, self.hr_width, 3)
        self.n_residual_blocks = 6                                          # fsrgan.py:21
        self.gf = self.df = 32                                              # :51-52
        self._setup(args, device)
        self.gen_optimizer = AdamConfig(args.lr, decay_steps=100000, decay_rate=0.1)
        self.disc_optimizer = AdamConfig(args.lr * 5, decay_steps=100000, decay_rate=0.1)
        self._build_vgg(args)
        g_init = (weights or {}).get("g") or P.init_fsrgan_generator(seed=getattr(args, "seed", 0), gf=self.gf, n_blocks=6)
        d_init = (weights or {}).get("d") or P.init_patch_discriminator(seed=getattr(args, "seed", 0) + 1)
        self.gen_params = ParamSet("g", g_init, self.device)
        self.disc_params = ParamSet("d", d_init, self.device)
        self.generator = FastSRGANGenerator(self.engine, self.gen_params, self.n_residual_blocks)
        self.discriminator = PatchDiscriminator(self.engine, self.disc_params, sigmoid=False)
