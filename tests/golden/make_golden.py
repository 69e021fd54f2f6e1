"""Writes the golden fixtures.  The reference has no golden vectors and TensorFlow is not installable
in the build container, so these pin the ORACLE against silent regressions;
they do not pin it against TensorFlow ("parity unpinned").  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from denoise_gan_b200 import params as P  # noqa: E402
from denoise_gan_b200.dataloader import synthetic_pair  # noqa: E402
from oracle import ops_torch as OT  # noqa: E402
from oracle import steps as OS  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
g = {k: v.double() for k, v in P.init_srgan_generator(seed=0).items()}
d = {k: v.double() for k, v in P.init_patch_discriminator(seed=1).items()}
x, y = synthetic_pair(2, 32, 4, step=0)
out = {}
losses = OS.srgan_train_step(g, d, None, OT.KerasAdam(1e-3, decay_steps=100000), OT.KerasAdam(5e-3, decay_steps=100000),
                             x.double(), y.double(), out=out)
np.savez_compressed(os.path.join(here, "srgan_step_small.npz"), x=x.numpy(), y=y.numpy(),
                    losses=np.array([float(v) for v in losses]), gen_output=out["gen_output"].numpy(),
                    g_conv_out_kernel_after=g["g/conv_out/kernel"].numpy())
print("wrote srgan_step_small.npz", [float(v) for v in losses])
