"""Checkpoint / resume (denoise_gan_b200/checkpoint.py) on CPU arenas: round trip of weights in Keras
layouts, BatchNorm moving statistics, Adam moments and counters; shape and name checking; weights-only files."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from denoise_gan_b200 import checkpoint as CK
from denoise_gan_b200 import params as P
from denoise_gan_b200.params import ParamSet


def _model(seed):
    m = SimpleNamespace(gen_params=ParamSet("g", P.init_srgan_generator(seed, 4), "cpu"),
                        disc_params=ParamSet("d", P.init_patch_discriminator(seed + 1), "cpu"), iterations=0, epochs=0)
    return m


def _scramble(m, seed):
    g = torch.Generator().manual_seed(seed)
    for ps in (m.gen_params, m.disc_params):
        ps.m.copy_(torch.randn(ps.m.shape, generator=g)); ps.v.copy_(torch.rand(ps.v.shape, generator=g))
        ps.state.copy_(torch.rand(ps.state.shape, generator=g))
        ps.opt_state.copy_(torch.tensor([1234 + seed, 987654321]))
    m.iterations, m.epochs = 1234 + seed, 7


def _same(a, b):
    for pa, pb in ((a.gen_params, b.gen_params), (a.disc_params, b.disc_params)):
        for name in list(pa.params) + list(pa.states):
            assert torch.equal(pa[name].data, pb[name].data), name
        for name, p in pa.params.items():
            sl = slice(p.offset, p.offset + p.numel)
            assert torch.equal(pa.m[sl], pb.m[sl]) and torch.equal(pa.v[sl], pb.v[sl]), name
        assert torch.equal(pa.opt_state, pb.opt_state)
    assert (a.iterations, a.epochs) == (b.iterations, b.epochs)


def test_round_trip_through_npz(tmp_path):
    a, b = _model(0), _model(5)
    _scramble(a, 3)
    path = str(tmp_path / "ckpt.npz")
    CK.save(a, path)
    assert CK.load(b, path) == []
    _same(a, b)
    sd = CK.state_dict(a)
    # Keras layouts and names: Conv2D kernels [kh, kw, Cin, Cout], 1 518 403 + 158 689 trainable values
    assert sd["gen/g/res0/conv1/kernel"].shape == (3, 3, 64, 64) and sd["gen/g/conv_out/kernel"].shape == (1, 1, 64, 3)
    n_g = sum(v.size for k, v in sd.items() if k.startswith("gen/") and not k.endswith(("moving_mean", "moving_variance")))
    n_d = sum(v.size for k, v in sd.items() if k.startswith("disc/") and not k.endswith(("moving_mean", "moving_variance")))
    assert (n_g, n_d) == (1518403, 158689)


def test_weights_only_and_mismatches():
    a, b = _model(0), _model(9)
    _scramble(b, 1)
    weights = {k: v for k, v in CK.state_dict(a).items() if k.startswith("gen/")}
    with pytest.raises(KeyError):
        CK.load_state_dict(b, weights, strict=True)
    b2 = _model(9); _scramble(b2, 1)
    missing = CK.load_state_dict(b2, weights, strict=False)
    assert "disc/d/conv1/kernel" in missing and "gen_opt/state" in missing and "meta/iterations" in missing
    for name in a.gen_params.params:
        assert torch.equal(a.gen_params[name].data, b2.gen_params[name].data)
    assert b2.iterations == 1235 and int(b2.gen_params.opt_state[0]) == 1235          # untouched: not in the file
    bad = dict(weights); bad["gen/g/res0/conv1/kernel"] = np.zeros((3, 3, 64, 32), np.float32)
    with pytest.raises(ValueError):
        CK.load_state_dict(_model(2), bad, strict=False)
    extra = CK.state_dict(a); extra["gen/not_a_layer/kernel"] = np.zeros(3, np.float32)
    with pytest.raises(KeyError):
        CK.load_state_dict(_model(2), extra, strict=True)


def test_failed_load_leaves_the_model_untouched():
    """load_state_dict validates every key and shape of BOTH networks before it copies anything: a file whose generator part
    is fine but whose discriminator part is wrong (or absent under strict) must not leave the generator half-loaded."""
    a = _model(0)
    for make_bad in ("disc_shape", "disc_missing_strict", "unexpected_strict"):
        b, ref = _model(9), _model(9)
        _scramble(b, 1); _scramble(ref, 1)
        sd = dict(CK.state_dict(a))
        if make_bad == "disc_shape":
            sd["disc/d/conv8/kernel"] = np.zeros((3, 3, 64, 32), np.float32)
            with pytest.raises(ValueError):
                CK.load_state_dict(b, sd, strict=False)
        elif make_bad == "disc_missing_strict":
            sd = {k: v for k, v in sd.items() if not k.startswith("disc")}
            with pytest.raises(KeyError):
                CK.load_state_dict(b, sd, strict=True)
        else:
            sd["gen/not_a_layer/kernel"] = np.zeros(3, np.float32)
            with pytest.raises(KeyError):
                CK.load_state_dict(b, sd, strict=True)
        _same(b, ref)


def test_keras_named_weights_only_file_round_trip(tmp_path):
    """The documented path for real weights: a weights-only .npz in Keras variable layouts (what tools/h5_to_npz.py writes from
    a Keras .h5: Conv2D kernels [kh,kw,Cin,Cout], BN gamma/beta/moving_*) loads with strict=False and leaves the optimiser cold."""
    a = _model(3)
    keras_like = {"gen/" + n: p.data.numpy().copy() for n, p in list(a.gen_params.params.items()) + list(a.gen_params.states.items())}
    path = str(tmp_path / "weights_only.npz")
    np.savez(path, **keras_like)
    b = _model(11)
    missing = CK.load(b, path, strict=False)
    assert all(not k.startswith("gen/") for k in missing) and any(k.startswith("gen_opt/") for k in missing)
    for n in list(a.gen_params.params) + list(a.gen_params.states):
        assert torch.equal(a.gen_params[n].data, b.gen_params[n].data), n
    assert float(b.gen_params.m.abs().sum()) == 0.0 and int(b.gen_params.opt_state[0]) == 0
