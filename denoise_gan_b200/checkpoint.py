"""Checkpoint / resume of a model's training state.

The reference saves the generator once per epoch as a Keras `.h5` and restarts from it when `args.retrain` is set
(train_srgan.py:220-227,244-245,256-259; autoencoder.py:141-146) — weights only, so its optimisers restart cold.  Here a
checkpoint is one `.npz` holding, for the generator and the discriminator, every tensor BY NAME IN ITS KERAS LAYOUT (Conv2D
`[kh,kw,Cin,Cout]`, Conv2DTranspose `[kh,kw,Cout,Cin]`, BN gamma/beta/moving_mean/moving_variance, PReLU alpha — the names of
DESIGN.md "Parameter naming"), the Adam first/second moments under the same names, both optimisers' iteration counters and the
model's `iterations` / `epochs`: resuming continues the run bit for bit (the dropout stream and the ExponentialDecay schedule
are functions of the stored counters).  A weights-only file — e.g. converted from a Keras `.h5` wherever h5py exists:
`np.savez(path, **{"gen/" + name: array})` — loads with `strict=False`.

Storage is host-side numpy; nothing here touches the hot path.  After loading into a model that has already run a step the
bf16 K-major weight copies of the tensor-core kernels are re-packed in place (their addresses are baked into captured CUDA
graphs)."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .params import ParamSet

_SETS = (("gen", "gen_params"), ("disc", "disc_params"))


def _opt_arrays(ps: ParamSet, prefix: str, out: dict):
    for name, p in ps.params.items():
        sl = slice(p.offset, p.offset + p.numel)
        out[f"{prefix}_opt/m/{name}"] = ps.m[sl].view(p.shape).detach().cpu().numpy().copy()
        out[f"{prefix}_opt/v/{name}"] = ps.v[sl].view(p.shape).detach().cpu().numpy().copy()
    out[f"{prefix}_opt/state"] = ps.opt_state.detach().cpu().numpy().copy()     # int64: iterations | lr bits (dg_adam_step)


def state_dict(model) -> "OrderedDict[str, np.ndarray]":
    """Everything needed to continue training `model`, as named host arrays."""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for prefix, attr in _SETS:
        ps: ParamSet = getattr(model, attr)
        for name, t in ps.export().items():
            out[f"{prefix}/{name}"] = t.numpy().copy()
        if ps.trainable:
            _opt_arrays(ps, prefix, out)
    out["meta/iterations"] = np.asarray(int(getattr(model, "iterations", 0)), dtype=np.int64)
    out["meta/epochs"] = np.asarray(int(getattr(model, "epochs", 0)), dtype=np.int64)
    return out


def load_state_dict(model, sd, strict: bool = True) -> list[str]:
    """Restores `state_dict(model)` output (or a weights-only subset with strict=False).  Returns the names that were
    expected but absent (always empty when strict).

    Two phases: (1) every key and shape of BOTH networks, the optimiser moments and the counters is validated and staged as
    host tensors without touching the model; (2) only then is the staged state copied in.  A mismatched or partial file
    therefore raises with the model exactly as it was (no half-loaded generator next to a stale discriminator)."""
    missing: list[str] = []
    known = set()
    staged = []            # (destination tensor, host source tensor)
    meta = {}

    def stage(key, shape, dst, dtype=np.float32):
        known.add(key)
        if key not in sd:
            missing.append(key)
            return
        a = np.asarray(sd[key])
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{key}: checkpoint shape {tuple(a.shape)} != model shape {tuple(shape)}")
        staged.append((dst, torch.from_numpy(np.ascontiguousarray(a.astype(dtype, copy=False)))))

    # ---- phase 1: validate + stage (no writes into the model)
    for prefix, attr in _SETS:
        ps: ParamSet = getattr(model, attr)
        for name in list(ps.params) + list(ps.states):
            stage(f"{prefix}/{name}", ps[name].shape, ps[name].data)
        if ps.trainable:
            for name, p in ps.params.items():
                sl = slice(p.offset, p.offset + p.numel)
                stage(f"{prefix}_opt/m/{name}", p.shape, ps.m[sl].view(p.shape))
                stage(f"{prefix}_opt/v/{name}", p.shape, ps.v[sl].view(p.shape))
            stage(f"{prefix}_opt/state", tuple(ps.opt_state.shape), ps.opt_state, np.int64)
    for key, attr in (("meta/iterations", "iterations"), ("meta/epochs", "epochs")):
        known.add(key)
        if key in sd:
            a = np.asarray(sd[key])
            if a.size != 1:
                raise ValueError(f"{key}: expected a scalar, got shape {tuple(a.shape)}")
            meta[attr] = int(a.reshape(()))
        else:
            missing.append(key)
    if strict:
        unexpected = sorted(k for k in sd if k not in known)
        if missing or unexpected:
            raise KeyError(f"checkpoint does not match the model: missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                           f"unexpected {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}")
    # ---- phase 2: commit
    for dst, src in staged:
        dst.copy_(src)
    for attr, v in meta.items():
        setattr(model, attr, v)
    for _, attr in _SETS:
        _repack_in_place(model, getattr(model, attr))
    return missing


def _repack_in_place(model, ps: ParamSet):
    """Refreshes the packed bf16 copies of kernels that already have one (no-op on CPU or before the first step)."""
    eng = getattr(model, "engine", None)
    if eng is None or not any(p.packed_fwd is not None or p.packed_dgrad is not None for p in ps.params.values()):
        return
    ps.repack(eng.lib, eng.ctx, eng.st)


def save(model, path: str) -> None:
    """Writes `state_dict(model)` as an uncompressed .npz (train_srgan.py:244-245 saves the generator every epoch)."""
    np.savez(path, **state_dict(model))


def load(model, path: str, strict: bool = True) -> list[str]:
    """Restores a checkpoint written by `save` (train_srgan.py:220-227: restart from saved weights when `args.retrain`)."""
    with np.load(path) as z:
        return load_state_dict(model, {k: z[k] for k in z.files}, strict=strict)
