"""ORACLE (test infrastructure, never imported by the product path) — the LITERAL reference under TensorFlow.

The build container has no TensorFlow (no wheel, no network: SURVEY.md §8c), so every parity claim of this repo rests on
`oracle/`'s restatement of the documented TF/Keras semantics: "parity unpinned".  This module is the other half of the
pin: on ANY machine with TensorFlow 2.x and a checkout of the reference it

  * imports the reference's own modules unmodified (`srgan.SRGAN`, `train_srgan.train_step`, ... from `DG_REFERENCE_DIR`,
    default /root/reference),
  * injects this repo's seeded weights (denoise_gan_b200/params.py initialisers, Keras layouts) into the Keras models,
    layer by layer in construction order with every shape checked,
  * feeds the repo's `synthetic_pair` batches through the reference `train_step` (eagerly, so that the gradients the
    step hands to `apply_gradients` can be recorded), and
  * returns per-layer activations, activation gradients, parameter gradients, the step's losses and the post-Adam weights
    as numpy arrays under THIS repo's parameter names.

`tests/golden/make_tf_golden.py` writes them to `tests/golden/tf/*.npz`; `tests/test_tf_golden.py` consumes the files when
they exist (oracle on CPU, CUDA path on the GPU) and reports "parity unpinned" when they do not.  NOT RUN in the build
container (TensorFlow absent): the code below is written against the TF 2.1-2.4 API the reference itself uses
(srgan.py:5 `mixed_precision.experimental`).

Appendix-B items exercised by the fixtures: 2 SAME padding asymmetry (stride-2 discriminator convs), 3
Conv2DTranspose (pix2pix), 4 depth_to_space order (SRGAN / Fast-SRGAN up-sampling), 5 BatchNormalization (batch statistics,
moving-average update incl. Bessel's correction, eps 1e-3, momentum 0.99 / 0.8 / 0.999), 8 losses (BCE from logits and from
probabilities, MSE, MAE, total variation), 10 Keras Adam (epsilon placement, bias correction, ExponentialDecay).
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np

REFERENCE_DIR = os.environ.get("DG_REFERENCE_DIR", "/root/reference")


def tf_available() -> bool:
    try:
        import tensorflow  # noqa: F401
        return True
    except Exception:
        return False


def _import_reference():
    """The reference's modules, unmodified, from REFERENCE_DIR."""
    if not os.path.isdir(REFERENCE_DIR):
        raise RuntimeError(f"reference checkout not found at {REFERENCE_DIR} (set DG_REFERENCE_DIR)")
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import importlib
    return {name: importlib.import_module(name) for name in
            ("srgan", "fsrgan", "autoencoder", "pix2pix", "train_srgan", "train_fsrgan", "train_autoencoder", "train_pix2pix")}


# ---------------------------------------------------------------------------------------------------------------------
# weight injection: this repo's names  <->  Keras layers in construction order
def _prefixes(tensors: "OrderedDict[str, np.ndarray]"):
    """Layer prefixes ('g/res0/conv1', 'g/res0/bn1', ...) in the order the initialisers create them, which is the order the
    reference builds its layers in (params.py mirrors srgan.py:129-185 etc. line by line)."""
    seen = OrderedDict()
    for name in tensors:
        seen.setdefault(name.rsplit("/", 1)[0], []).append(name.rsplit("/", 1)[1])
    return seen


_KERAS_SLOTS = {           # Keras weight order of each layer class  ->  this repo's tensor suffixes
    "Conv2D": ("kernel", "bias"), "Conv2DTranspose": ("kernel", "bias"), "DepthwiseConv2D": ("kernel", "bias"),
    "BatchNormalization": ("gamma", "beta", "moving_mean", "moving_variance"), "PReLU": ("alpha",),
}


def _weighted_layers(model):
    out = []
    for layer in model.layers:
        if hasattr(layer, "layers"):                      # nested Sequential (pix2pix downsample / upsample blocks)
            out.extend(_weighted_layers(layer))
        elif layer.weights:
            out.append(layer)
    return out


def inject(model, tensors) -> "OrderedDict[str, object]":
    """Assigns `tensors` (this repo's names, Keras layouts) to the Keras `model`; returns name -> tf.Variable.
    Every layer class and every shape is checked: a mismatch means the construction orders differ and raises."""
    layers = _weighted_layers(model)
    groups = _prefixes(tensors)
    if len(layers) != len(groups):
        raise ValueError(f"{model.name}: {len(layers)} weighted Keras layers vs {len(groups)} parameter groups")
    var_of = OrderedDict()
    for layer, (prefix, suffixes) in zip(layers, groups.items()):
        slots = _KERAS_SLOTS.get(type(layer).__name__)
        if slots is None:
            raise ValueError(f"unexpected weighted layer {type(layer).__name__} ({layer.name}) at {prefix}")
        ws = layer.weights
        names = [s for s in slots if s in suffixes]
        if len(names) != len(ws):
            raise ValueError(f"{prefix} ({type(layer).__name__}): Keras has {len(ws)} weights, the initialiser has {suffixes}")
        for v, suffix in zip(ws, names):
            a = np.asarray(tensors[f"{prefix}/{suffix}"], dtype=np.float32)
            if a.size != int(np.prod(v.shape)):
                raise ValueError(f"{prefix}/{suffix}: shape {a.shape} vs Keras {tuple(v.shape)}")
            v.assign(a.reshape(tuple(v.shape)))           # PReLU alpha (C,) -> (1,1,C); depthwise [3,3,C] -> [3,3,C,1]
            var_of[f"{prefix}/{suffix}"] = v
    return var_of


def export(var_of, like) -> "OrderedDict[str, np.ndarray]":
    return OrderedDict((n, np.asarray(v.numpy(), dtype=np.float32).reshape(np.asarray(like[n]).shape)) for n, v in var_of.items())


# ---------------------------------------------------------------------------------------------------------------------
def _activation_taps(model):
    """Outputs worth comparing layer by layer: every convolution, BatchNormalization, activation, Add and the model output."""
    keep = ("Conv2D", "Conv2DTranspose", "DepthwiseConv2D", "BatchNormalization", "PReLU", "LeakyReLU", "ReLU", "Activation", "Add",
            "Concatenate", "MaxPooling2D", "UpSampling2D", "TensorFlowOpLayer", "Lambda")
    return [l for l in model.layers if type(l).__name__ in keep]


def run_step(kind: str, g_tensors, d_tensors, x: np.ndarray, y: np.ndarray, *, lr=1e-3, crop=None, scale=4, vgg_tensors=None,
             steps: int = 1, fp16: int = 0):
    """Runs the reference's own `train_step` `steps` times on (x, y) with the injected weights.

    kind: 'srgan' | 'fsrgan' | 'autoencoder' | 'pix2pix'.  Returns a dict of numpy arrays:
      losses            [steps, n]  the tuple train_step returns, per step
      act/<layer>       activations of the generator's layers for the FIRST step's batch (training=True)
      dact/<layer>      d gen_loss / d activation for the same pass (SRGAN-family and autoencoder)
      ggrad/<name>, dgrad/<name>   the gradients the first step handed to apply_gradients, under this repo's names
      g_after/<name>, d_after/<name>   every variable (incl. BN moving statistics) after the LAST step
    """
    import tensorflow as tf
    tf.config.run_functions_eagerly(True)                 # train_step is a @tf.function: run its Python body so that hooks see values
    ref = _import_reference()
    crop = int(crop if crop is not None else y.shape[1])
    args = SimpleNamespace(crop_size=crop, scale=scale, lr=lr, fp16=fp16, retrain=0, batch_size=int(x.shape[0]), epochs=1,
                           pretrained_weights=None, pretrained=None)
    real_vgg = tf.keras.applications.VGG19

    def vgg_no_download(*a, **kw):                        # seeded synthetic VGG19 weights are injected below; never download
        kw["weights"] = None
        return real_vgg(*a, **kw)
    tf.keras.applications.VGG19 = vgg_no_download
    try:
        cls = {"srgan": ref["srgan"].SRGAN, "fsrgan": ref["fsrgan"].FastSRGAN, "autoencoder": ref["autoencoder"].Autoencoder,
               "pix2pix": ref["pix2pix"].Pix2Pix}[kind]
        model = cls(args)
    finally:
        tf.keras.applications.VGG19 = real_vgg
    step_fn = {"srgan": ref["train_srgan"], "fsrgan": ref["train_fsrgan"], "autoencoder": ref["train_autoencoder"],
               "pix2pix": ref["train_pix2pix"]}[kind].train_step
    gvars = inject(model.generator, g_tensors)
    dvars = inject(model.discriminator, d_tensors)
    if vgg_tensors is not None and getattr(model, "vgg", None) is not None:
        inject(model.vgg, vgg_tensors)
    elif getattr(model, "vgg", None) is not None and vgg_tensors is None:
        # "G+D step" fixtures: the content term must vanish exactly -> zero the VGG weights (features == 0 for both images)
        for v in model.vgg.weights:
            v.assign(tf.zeros_like(v))
    out = {}
    xt, yt = tf.constant(x, tf.float32), tf.constant(y, tf.float32)

    # ---- layer-by-layer activations and their gradients for the first batch (no variable is modified by this pass except the
    # BN moving statistics, which are restored afterwards)
    state0 = [v.numpy() for v in model.generator.weights + model.discriminator.weights]
    taps = _activation_taps(model.generator)
    probe = tf.keras.Model(model.generator.inputs, [l.output for l in taps])
    if kind != "pix2pix":
        with tf.GradientTape() as tape:
            acts = probe(xt, training=True)
            gen_out = acts[-1]
            d_fake = model.discriminator(gen_out, training=True)
            bce = tf.keras.losses.BinaryCrossentropy(from_logits=(kind != "autoencoder"))
            loss = 1e-3 * bce(tf.ones_like(d_fake), d_fake) + tf.reduce_mean(tf.abs(yt - gen_out))
            if vgg_tensors is not None:
                loss = loss + model.content_loss(yt, gen_out)
        dacts = tape.gradient(loss, acts)
        for l, a, da in zip(taps, acts, dacts):
            out[f"act/{l.name}"] = a.numpy()
            if da is not None:
                out[f"dact/{l.name}"] = da.numpy()
        out["act_order"] = np.array([l.name for l in taps])
        out["act_class"] = np.array([type(l).__name__ for l in taps])
    for v, a in zip(model.generator.weights + model.discriminator.weights, state0):
        v.assign(a)

    # ---- the literal train_step; record what it hands to apply_gradients on the first step
    rec = {}

    def hook(opt, tag, var_of):
        inner = opt.apply_gradients
        by_id = {id(v): n for n, v in var_of.items()}

        def apply(grads_and_vars, *a, **kw):
            gv = list(grads_and_vars)
            if tag not in rec:
                rec[tag] = {by_id[id(v)]: (None if g is None else np.asarray(g.numpy(), np.float32)) for g, v in gv if id(v) in by_id}
            return inner(gv, *a, **kw)
        opt.apply_gradients = apply
    hook(model.gen_optimizer, "ggrad", gvars)
    hook(model.disc_optimizer, "dgrad", dvars)
    losses = []
    for _ in range(steps):
        r = step_fn(model, xt, yt)
        losses.append([float(v) for v in r])
    out["losses"] = np.asarray(losses, np.float64)
    for tag in ("ggrad", "dgrad"):
        for n, g in rec.get(tag, {}).items():
            if g is not None:
                like = (g_tensors if tag == "ggrad" else d_tensors)[n]
                out[f"{tag}/{n}"] = g.reshape(np.asarray(like).shape)
    for n, a in export(gvars, g_tensors).items():
        out[f"g_after/{n}"] = a
    for n, a in export(dvars, d_tensors).items():
        out[f"d_after/{n}"] = a
    out["tf_version"] = np.array(tf.__version__)
    return out
