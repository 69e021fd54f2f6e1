"""denoise_gan_b200 — B200-native hot path of pmcbride/denoise-gan (conv forward/backward of the
generator and discriminator networks) behind the reference's Python surface.

    from denoise_gan_b200.srgan import SRGAN
    from denoise_gan_b200.train_srgan import train_step
"""
__version__ = "0.1.0"
