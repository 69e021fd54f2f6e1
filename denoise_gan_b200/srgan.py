"""Drop-in for the reference's `srgan.SRGAN` (srgan.py:8-272): same constructor arguments and
attributes, networks executed by the B200 kernels."""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import params as P
from .engine import Engine
from .nets import PatchDiscriminator, SRGANGenerator, VGG19Features
from .params import ParamSet


class AdamConfig(SimpleNamespace):
    """Stand-in for tf.keras.optimizers.Adam(+ExponentialDecay staircase) hyper-parameters."""

    def __init__(self, lr, beta_1=0.9, beta_2=0.999, epsilon=1e-7, decay_steps=0, decay_rate=0.1):
        super().__init__(lr=lr, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, decay_steps=decay_steps, decay_rate=decay_rate)

    def apply(self, engine: Engine, pset: ParamSet, grad_scale=1.0):
        engine.adam(pset, self.lr, self.beta_1, self.beta_2, self.epsilon, self.decay_steps, self.decay_rate, grad_scale)


class _GanBase:
    """Common plumbing of the four model classes."""

    def _setup(self, args, device=None):
        self.iterations = 0
        self.epochs = 0
        self.fp16 = bool(getattr(args, "fp16", 0))
        self.engine = Engine(device, bf16=self.fp16)
        self.device = self.engine.device
        self.use_vgg = bool(getattr(args, "vgg", True))
        self.world_size = 1
        self.comm = None

    def _build_vgg(self, args):
        if not self.use_vgg:
            self.vgg = None
            return
        weights = getattr(args, "vgg_weights", None)
        tensors = P.init_vgg19_synthetic() if weights is None else weights
        self.vgg_params = ParamSet("vgg", tensors, self.device, trainable=False)
        self.vgg = VGG19Features(self.engine, self.vgg_params)

    def content_loss(self, target, gen_output, key="content"):
        """srgan.py:69-75: MSE of VGG19 block5_conv4 features / 12.75 (caffe preprocessing).
        Returns (loss device scalar, seed gradient for gen_output's feature map, feature Var)."""
        E = self.engine
        tgt = target if hasattr(target, "deps") else E.input(target)
        gf = self.vgg(E.vgg_preprocess(gen_output))
        tf_ = self.vgg(E.vgg_preprocess(tgt))
        loss, dgf = E.feature_mse(gf, tf_, 1.0 / 12.75, key=key)
        return loss, dgf, gf


    # ---- value-only loss methods kept for API completeness: the train steps compute the same
    # terms fused with their gradients (train_common.gan_step / train_pix2pix.train_step).
    def _var(self, t):
        return t if hasattr(t, "deps") else self.engine.input(t)

    def discriminator_loss(self, disc_real_output, disc_generated_output):
        """srgan.py:120-127 / pix2pix.py:96-103: BCE_logits(1, D(real)) + BCE_logits(0, D(fake)); device scalar."""
        E = self.engine
        real, _ = E.bce(self._var(disc_real_output), 1.0, True, 0.0, key="api_real")
        fake, _ = E.bce(self._var(disc_generated_output), 0.0, True, 0.0, key="api_fake")
        return real + fake

    def _loss_terms(self, disc_generated_output, gen_output, target):
        E = self.engine
        adv, _ = E.bce(self._var(disc_generated_output), 1.0, True, 0.0, key="api_adv")
        tgt = target.t if hasattr(target, "deps") else target
        out3, _ = E.image_losses(self._var(gen_output), tgt, 0.0, 0.0, 0.0, key="api_img")
        cont = self.content_loss(target, self._var(gen_output), key="api_content")[0] if self.vgg is not None else torch.zeros(1, device=self.device)
        return 1e-3 * adv[0], out3[0], out3[1], cont.reshape(()), 1e-5 * out3[2]


class SRGAN(_GanBase):
    """SRGAN for super resolution (reference: srgan.py:8-67)."""

    def __init__(self, args, device=None, weights=None):
        self.scale = args.scale
        self.hr_height = self.hr_width = args.crop_size
        self.lr_height = self.hr_height // self.scale
        self.lr_width = self.hr_width // self.scale
        self.lr_shape = [self.lr_height, self.lr_width, 3]
        self.hr_shape = [self.hr_height, self.hr_width, 3]
        self._setup(args, device)
        # learning-rate schedules, srgan.py:35-50 (TTUR: discriminator lr x5)
        self.gen_optimizer = AdamConfig(args.lr, decay_steps=100000, decay_rate=0.1)
        self.disc_optimizer = AdamConfig(args.lr * 5, decay_steps=100000, decay_rate=0.1)
        self._build_vgg(args)
        g_init = (weights or {}).get("g") or P.init_srgan_generator(seed=getattr(args, "seed", 0), scale=self.scale)
        d_init = (weights or {}).get("d") or P.init_patch_discriminator(seed=getattr(args, "seed", 0) + 1)
        self.gen_params = ParamSet("g", g_init, self.device)
        self.disc_params = ParamSet("d", d_init, self.device)
        self.generator = SRGANGenerator(self.engine, self.gen_params, self.scale)
        self.discriminator = PatchDiscriminator(self.engine, self.disc_params, sigmoid=False)

    def generator_loss(self, disc_generated_output, gen_output, target):
        """srgan.py:97-117 (not called by train_srgan.py, whose inline loss is authoritative): total = adv + l2 + content.
        Returns (total, adv, l1, l2, content, var) device scalars."""
        adv, l1, l2, cont, var = self._loss_terms(disc_generated_output, gen_output, target)
        return adv + l2 + cont, adv, l1, l2, cont, var
