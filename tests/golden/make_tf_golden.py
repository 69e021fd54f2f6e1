#!/usr/bin/env python
"""Writes the TensorFlow-pinned fixtures tests/golden/tf/*.npz by running the LITERAL reference (oracle/tf_hook.py).

Needs TensorFlow 2.x and a checkout of pmcbride/denoise-gan; neither exists in the build container, so the fixtures are
produced elsewhere with ONE command from the repo root and committed:

    DG_REFERENCE_DIR=/path/to/denoise-gan python tests/golden/make_tf_golden.py [--only srgan,fsrgan,autoencoder,pix2pix]

Each file holds the seeded inputs (this repo's `synthetic_pair`), a checksum of the injected weights (the initialisers of
denoise_gan_b200/params.py are deterministic), and the reference's per-layer activations, activation gradients, parameter
gradients, losses over the steps, and every variable after the last step.  tests/test_tf_golden.py replays them against
`oracle/` (CPU) and against the CUDA path (GPU).  Small shapes: the files stay below a few MB."""
import argparse
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from denoise_gan_b200 import params as P  # noqa: E402
from denoise_gan_b200.dataloader import synthetic_pair  # noqa: E402
from oracle import tf_hook  # noqa: E402

# kind: (batch, crop, scale, steps, with_vgg)
CASES = {"srgan": (2, 64, 4, 3, False), "srgan_vgg": (2, 64, 4, 1, True), "fsrgan": (2, 64, 4, 2, False),
         "autoencoder": (2, 64, 1, 2, False), "pix2pix": (1, 256, 1, 1, False)}


def checksum(tensors):
    c = 0
    for k, v in tensors.items():
        c = zlib.crc32(np.ascontiguousarray(v.numpy() if hasattr(v, "numpy") else v, dtype=np.float32).tobytes(), zlib.crc32(k.encode(), c))
    return np.array(c, dtype=np.uint32)


def weights(kind, scale):
    base = kind.split("_")[0]
    if base == "pix2pix":
        return P.init_pix2pix(0)
    g = {"srgan": lambda: P.init_srgan_generator(0, scale), "fsrgan": lambda: P.init_fsrgan_generator(0),
         "autoencoder": lambda: P.init_autoencoder_generator(0)}[base]()
    return g, P.init_patch_discriminator(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=",".join(CASES))
    args = ap.parse_args()
    if not tf_hook.tf_available():
        raise SystemExit("TensorFlow is not importable here: run this script on a machine with TensorFlow 2.x (see the docstring)")
    out_dir = os.path.join(ROOT, "tests", "golden", "tf")
    os.makedirs(out_dir, exist_ok=True)
    for name in args.only.split(","):
        batch, crop, scale, steps, with_vgg = CASES[name]
        g, d = weights(name, scale)
        x, y = synthetic_pair(batch, crop, scale, step=0)
        v = P.init_vgg19_synthetic() if with_vgg else None
        res = tf_hook.run_step(name.split("_")[0], {k: t.numpy() for k, t in g.items()}, {k: t.numpy() for k, t in d.items()},
                               x.numpy(), y.numpy(), crop=crop, scale=scale, steps=steps,
                               vgg_tensors=None if v is None else {k: t.numpy() for k, t in v.items()})
        res.update(x=x.numpy(), y=y.numpy(), g_checksum=checksum(g), d_checksum=checksum(d), case=np.array([batch, crop, scale, steps, int(with_vgg)]))
        path = os.path.join(out_dir, f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"wrote {path}: losses {res['losses'].tolist()} (TensorFlow {res['tf_version']})")


if __name__ == "__main__":
    main()
