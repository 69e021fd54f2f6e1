#!/usr/bin/env python
"""Writes tests/golden/jpeg_roundtrip.npz: small uint8 images and their REAL libjpeg-turbo round trips (Pillow: baseline JPEG,
4:2:0 chroma subsampling, the given quality, default decoder = 'islow' IDCT + fancy up-sampling) -- what
tf.image.adjust_jpeg_quality (dataloader.py:138) does between its two dtype conversions.  Run wherever Pillow exists:
    python tests/golden/make_jpeg_golden.py
The fixture pins oracle/pairs.py::jpeg_roundtrip_u8 on machines without Pillow (the GPU box)."""
import io
import os

import numpy as np
from PIL import Image

rng = np.random.default_rng(7)
yy, xx = np.mgrid[0:48, 0:64]
smooth = np.stack([127 + 100 * np.sin(xx / 9.0 + c) + 60 * np.cos(yy / 7.0 - c) for c in range(3)], -1) + rng.normal(0, 12, (48, 64, 3))
images = {"noise": rng.integers(0, 256, (32, 48, 3), dtype=np.uint8), "smooth": np.clip(smooth, 0, 255).astype(np.uint8),
          "flat": np.full((16, 16, 3), 200, dtype=np.uint8)}
out = {}
for name, img in images.items():
    out[f"{name}/src"] = img
    for q in (10, 50, 75, 95):
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="JPEG", quality=q, subsampling=2)
        out[f"{name}/q{q}"] = np.array(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jpeg_roundtrip.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("src")})
