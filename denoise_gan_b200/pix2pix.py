"""Drop-in for the reference's `pix2pix.Pix2Pix` (pix2pix.py:4-226)."""
from __future__ import annotations

from . import params as P
from .nets import Pix2PixDiscriminator, Pix2PixGenerator
from .params import ParamSet
from .srgan import AdamConfig, _GanBase


class Pix2Pix(_GanBase):
    """Denoising pix2pix (reference: pix2pix.py:7-43)."""

    def __init__(self, args, device=None, weights=None):
        self.hr_height = self.hr_width = args.crop_size
        self.lr_height, self.lr_width = self.hr_height, self.hr_width
        self.lr_shape = (self.lr_height, self.lr_width, 3)
        self.hr_shape = (self.hr_height, self.hr_width, 3)
        self.retrain = bool(getattr(args, "retrain", 0))
        assert args.crop_size == 256, "the reference's discriminator hard-codes 256x256 inputs (pix2pix.py:197-198)"
        self._setup(args, device)
        self.gen_optimizer = AdamConfig(2e-4, beta_1=0.5)      # pix2pix.py:30-31
        self.disc_optimizer = AdamConfig(2e-4, beta_1=0.5)
        self.gf = self.df = 32
        self._build_vgg(args)
        g_init, d_init = P.init_pix2pix(seed=getattr(args, "seed", 0))
        g_init = (weights or {}).get("g") or g_init
        d_init = (weights or {}).get("d") or d_init
        self.gen_params = ParamSet("g", g_init, self.device)
        self.disc_params = ParamSet("d", d_init, self.device)
        self.generator = Pix2PixGenerator(self.engine, self.gen_params, dropout_seed=getattr(args, "dropout_seed", 7))
        self.discriminator = Pix2PixDiscriminator(self.engine, self.disc_params)

    def generator_loss(self, disc_generated_output, gen_output, target):
        """pix2pix.py:74-94: total = 1e-3*BCE(1, D) + mse + content + 1e-5*mean(TV(target-gen)) + mae + identity, where the
        identity term runs the generator again on `target` with training=True (BN moving statistics and dropout advance, as
        in the reference).  Value-only; train_pix2pix.train_step computes the same terms fused with their gradients.
        Returns (total, gan, l1, l2, content, var, identity) device scalars."""
        E = self.engine
        adv, l1, l2, cont, var = self._loss_terms(disc_generated_output, gen_output, target)
        tgt = target.t if hasattr(target, "deps") else target
        ident_out = self.generator(tgt, training=True, pass_id=1)
        ident = E.image_losses(ident_out, tgt, 0.0, 0.0, 0.0, key="api_ident")[0][0]
        return adv + l2 + cont + var + l1 + ident, adv, l1, l2, cont, var, ident
