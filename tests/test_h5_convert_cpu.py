"""CPU: the order-and-shape matching core of tools/h5_to_npz.py (Keras `.h5` -> checkpoint `.npz`, train_srgan.py:256-259 /
autoencoder.py:141-146) on a synthetic file listing with Keras-style automatic layer names; h5py itself is not needed."""
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import h5_to_npz as H  # noqa: E402


def _fake_keras_file(spec, rng):
    """What read_h5 would return for a model whose layers were created in the spec's order, with Keras auto-names, one
    weight-less layer (an Add / Activation) between the weighted ones, and PReLU alphas stored as (1,1,C)."""
    layers, counts = [], {}
    for prefix, want in H.target_groups(spec):
        kind = ("batch_normalization" if "gamma" in want else "p_re_lu" if "alpha" in want else "conv2d")
        counts[kind] = counts.get(kind, 0) + 1
        lname = kind if counts[kind] == 1 else f"{kind}_{counts[kind] - 1}"
        ws = []
        for suffix, (full, shape) in want.items():
            a = rng.standard_normal(shape).astype(np.float32)
            if suffix == "alpha":
                a = a.reshape((1, 1) + shape)
            ws.append((f"{lname}/{suffix}:0", a))
        layers.append((lname, ws))
        layers.append((f"add_{len(layers)}", []))
    return layers


@pytest.mark.parametrize("model,net", [("srgan", "gen"), ("srgan", "disc"), ("fsrgan", "gen"), ("autoencoder", "gen"), ("pix2pix", "gen"),
                                       ("pix2pix", "disc"), ("vgg19", "gen")])
def test_match_by_order_and_shape(model, net):
    spec = H.model_spec(model, net)
    rng = np.random.default_rng(3)
    layers = _fake_keras_file(spec, rng)
    out = H.match(layers, spec)
    assert list(out) == list(spec) or set(out) == set(spec)
    flat = {}
    for lname, ws in layers:
        for vname, a in ws:
            flat[(lname, vname.split("/")[-1].split(":")[0])] = a
    # every tensor landed under the right name, bit for bit, in the declared shape
    it = iter([(ln, ws) for ln, ws in layers if ws])
    for prefix, want in H.target_groups(spec):
        lname, ws = next(it)
        for suffix, (full, shape) in want.items():
            src = dict((v.split("/")[-1].split(":")[0], a) for v, a in ws)[suffix]
            assert out[full].shape == shape and np.array_equal(out[full].reshape(-1), src.reshape(-1))


def test_mismatches_abort():
    spec = H.model_spec("srgan", "gen")
    layers = _fake_keras_file(spec, np.random.default_rng(0))
    with pytest.raises(ValueError, match="weighted layers"):
        H.match(layers[:-4], spec)
    bad = [(n, [(v, a[..., :-1] if a.ndim == 4 else a) for v, a in ws]) for n, ws in layers]
    with pytest.raises(ValueError, match="shape"):
        H.match(bad, spec)
    swapped = list(layers)
    i = [k for k, (n, ws) in enumerate(swapped) if ws][1]
    j = [k for k, (n, ws) in enumerate(swapped) if ws][2]
    swapped[i], swapped[j] = swapped[j], swapped[i]
    with pytest.raises(ValueError):
        H.match(swapped, spec)


def test_converted_file_loads_into_a_model_state(tmp_path):
    """The converter's output keys are the checkpoint's: `gen/<name>` in Keras layouts (checkpoint.load(strict=False))."""
    spec = H.model_spec("srgan", "gen")
    arrays = H.match(_fake_keras_file(spec, np.random.default_rng(1)), spec)
    path = os.path.join(tmp_path, "g.npz")
    np.savez(path, **{f"gen/{k}": v for k, v in arrays.items()})
    with np.load(path) as z:
        assert set(z.files) == {f"gen/{k}" for k in spec}
        for k, shape in spec.items():
            assert z[f"gen/{k}"].shape == shape
