#!/usr/bin/env python
"""Keras `.h5` -> checkpoint `.npz` converter.

The reference saves its generators (and only those) as Keras HDF5 files (train_srgan.py:256-259 `model.generator.save(...)`,
autoencoder.py:141-146 loads them back when `args.retrain`), and VGG19's ImageNet weights ship as a Keras-applications `.h5`.
This script turns such a file into the `.npz` that `denoise_gan_b200.checkpoint.load(model, path, strict=False)` takes: every
tensor under the package's own name, in its Keras layout (no transposition: the arenas store Keras layouts).

    python tools/h5_to_npz.py --model srgan|fsrgan|autoencoder|pix2pix|vgg19 [--net gen|disc] in.h5 out.npz

Runs wherever `h5py` exists (it is not installed on the GPU image; the reference's own environment has it).  Keras gives its
layers automatic names (conv2d_17, batch_normalization_3, p_re_lu_1 ...) that depend on how many layers the process created
before, so tensors are matched by ORDER and SHAPE, not by name: the weighted layers of the file, in the file's `layer_names`
order (= the order the builder created them), are walked next to the package's parameter groups, which are declared in the same
builder order (params.py cites the reference lines); every shape is checked, a mismatch aborts.  `match()` is that walk and
has no h5py dependency (tests/test_h5_convert_cpu.py)."""
from __future__ import annotations

import argparse
import os
import sys
from collections import OrderedDict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# Keras variable name (last path component, without ':0') -> the package's suffix
_SUFFIX = {"kernel": "kernel", "depthwise_kernel": "kernel", "bias": "bias", "gamma": "gamma", "beta": "beta",
           "moving_mean": "moving_mean", "moving_variance": "moving_variance", "alpha": "alpha"}


def target_groups(spec: "OrderedDict[str, tuple]"):
    """[(layer prefix, {suffix: (full name, shape)})] in declaration order; a layer = the names sharing everything up to the last '/'."""
    groups: "OrderedDict[str, dict]" = OrderedDict()
    for name, shape in spec.items():
        prefix, suffix = name.rsplit("/", 1)
        groups.setdefault(prefix, {})[suffix] = (name, tuple(shape))
    return list(groups.items())


def match(keras_layers, spec: "OrderedDict[str, tuple]") -> "OrderedDict[str, np.ndarray]":
    """keras_layers: [(layer name, [(variable name, array)])] for the layers of the file that HAVE weights, in file order.
    Returns {package name: array}.  Raises ValueError on any count / suffix / shape disagreement."""
    groups = target_groups(spec)
    layers = [(n, w) for n, w in keras_layers if len(w) > 0]
    if len(layers) != len(groups):
        raise ValueError(f"the file has {len(layers)} weighted layers, the model declares {len(groups)} "
                         f"(first file layers: {[n for n, _ in layers[:4]]}, first model layers: {[g for g, _ in groups[:4]]})")
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for (lname, weights), (prefix, want) in zip(layers, groups):
        got = {}
        for vname, arr in weights:
            base = vname.split("/")[-1].split(":")[0]
            if base not in _SUFFIX:
                raise ValueError(f"{lname}: unknown variable {vname}")
            got[_SUFFIX[base]] = np.asarray(arr)
        if set(got) != set(want):
            raise ValueError(f"file layer {lname} holds {sorted(got)}, model layer {prefix} expects {sorted(want)}")
        for suffix, (full, shape) in want.items():
            a = got[suffix]
            if tuple(a.shape) != shape:
                # PReLU(shared_axes=[1,2]) stores alpha as (1,1,C); DepthwiseConv2D kernels are (kh,kw,C,1) in both
                if a.size == int(np.prod(shape)) and suffix == "alpha":
                    a = a.reshape(shape)
                else:
                    raise ValueError(f"{lname}/{suffix} has shape {tuple(a.shape)}, {full} expects {shape}")
            out[full] = a.astype(np.float32, copy=False)
    return out


def model_spec(model: str, net: str) -> "OrderedDict[str, tuple]":
    from denoise_gan_b200 import params as P
    if model == "vgg19":
        t = P.init_vgg19_synthetic()
    elif net == "disc":
        t = P.init_pix2pix()[1] if model == "pix2pix" else P.init_patch_discriminator()
    else:
        t = {"srgan": P.init_srgan_generator, "fsrgan": P.init_fsrgan_generator, "autoencoder": P.init_autoencoder_generator,
             "pix2pix": lambda: P.init_pix2pix()[0]}[model]()
    return OrderedDict((k, tuple(v.shape)) for k, v in t.items())


def read_h5(path: str):
    """[(layer name, [(variable name, array)])] from a Keras HDF5 file (full model or save_weights)."""
    import h5py  # not on the GPU image: run this where the reference's environment lives
    with h5py.File(path, "r") as f:
        g = f["model_weights"] if "model_weights" in f else f
        names = [n.decode() if isinstance(n, bytes) else n for n in g.attrs["layer_names"]]
        layers = []
        for n in names:
            wn = [w.decode() if isinstance(w, bytes) else w for w in g[n].attrs.get("weight_names", [])]
            layers.append((n, [(w, np.asarray(g[n][w])) for w in wn]))
    return layers


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--model", required=True, choices=["srgan", "fsrgan", "autoencoder", "pix2pix", "vgg19"])
    ap.add_argument("--net", default="gen", choices=["gen", "disc"])
    ap.add_argument("h5")
    ap.add_argument("npz")
    a = ap.parse_args()
    arrays = match(read_h5(a.h5), model_spec(a.model, a.net))
    prefix = "vgg" if a.model == "vgg19" else a.net
    np.savez(a.npz, **{f"{prefix}/{k}": v for k, v in arrays.items()})
    print(f"{a.npz}: {len(arrays)} tensors, {sum(v.size for v in arrays.values())} values")


if __name__ == "__main__":
    main()
