"""ORACLE (test infrastructure, never imported by the product path) — network restatements.

torch-CPU functional restatements of the reference's generator / discriminator / VGG19 graphs.
Each function takes a flat `params` dict (Keras-layout tensors, names documented in
DESIGN.md "Parameter naming") and returns the NHWC output; `acts`, when given, collects
per-layer activations for the per-layer parity tests; `state_out` collects the updated BN
moving statistics of a training-mode call.

PARITY UNPINNED: follows the reference source line by line, cannot be run
against TensorFlow here.
"""
from __future__ import annotations

import torch

from . import ops_torch as T


def _rec(acts, name, x):
    if acts is not None:
        acts[name] = x
    return x


# ----------------------------------------------------------------------------- SRGAN
def bf16_quant(t):
    """Storage rounding of the bf16 path (straight-through for autograd): what the CUDA kernels do when
    they write an activation to HBM.  Used to build a bf16-EMULATING oracle whose only differences from
    the tensor-core path are accumulation order and gradient storage rounding."""
    return t.to(torch.bfloat16).to(t.dtype)


def _ident(t):
    return t


def _wq(w, q):
    """tensor-core layers consume bf16 copies of the kernels (Cin and Cout multiples of 16)."""
    return q(w) if (w.shape[2] % 16 == 0 and w.shape[3] % 16 == 0) else w


def srgan_generator(p, x, training=True, state_out=None, acts=None, scale=4, q=None):
    """srgan.py:129-185.  `q` (optional) rounds every stored activation, emulating the bf16 path."""
    q = q or _ident
    bn = lambda t, name: T.batch_norm(t, p, name, training, state_out, momentum=0.99, eps=1e-3)
    n = q(T.conv2d(x, p["g/conv_in/kernel"]))                                # :154
    n = bn(n, "g/bn_in")                                                     # :155
    n = q(T.prelu(n, p["g/prelu_in/alpha"]))                                 # :157
    temp = _rec(acts, "g/prelu_in", n)
    for i in range(16):                                                      # :161-170
        nn = q(T.conv2d(n, _wq(p[f"g/res{i}/conv1/kernel"], q)))
        nn = q(torch.relu(bn(nn, f"g/res{i}/bn1")))
        nn = q(T.conv2d(nn, _wq(p[f"g/res{i}/conv2/kernel"], q)))
        nn = bn(nn, f"g/res{i}/bn2")
        n = _rec(acts, f"g/res{i}/add", q(n + nn))
    n2 = q(T.conv2d(n, _wq(p["g/conv_post/kernel"], q)))                     # :172
    n2 = bn(n2, "g/bn_post")
    n = _rec(acts, "g/post_add", q(n2 + temp))                               # :175
    for j in range(scale // 2):                                              # :179-180, deconv2d :134-147
        u = q(T.conv2d(n, _wq(p[f"g/up{j}/conv/kernel"], q), p[f"g/up{j}/conv/bias"]))
        u = T.depth_to_space(u, 2)
        n = _rec(acts, f"g/up{j}/prelu", q(T.prelu(u, p[f"g/up{j}/prelu/alpha"])))
    out = T.conv2d(n, p["g/conv_out/kernel"], p["g/conv_out/bias"])          # :182 (1x1)
    return _rec(acts, "g/tanh", torch.tanh(out))                             # :183 (fp32 output)


def patch_discriminator(p, x, training=True, state_out=None, acts=None, sigmoid=False, prefix="d", q=None):
    """The 9-conv discriminator shared by srgan.py:232-272, fsrgan.py:222-258 (logits) and
    autoencoder.py:190-229 (sigmoid)."""
    q = q or _ident
    filters_strides = [(32, 1), (32, 2), (32, 1), (32, 2), (64, 1), (64, 2), (64, 1), (64, 2)]
    d = x
    for i, (_, s) in enumerate(filters_strides, start=1):
        d = T.conv2d(d, _wq(p[f"{prefix}/conv{i}/kernel"], q), p[f"{prefix}/conv{i}/bias"], stride=s)
        if i > 1:
            d = T.batch_norm(q(d), p, f"{prefix}/bn{i}", training, state_out, momentum=0.8, eps=1e-3)
        d = _rec(acts, f"{prefix}/lrelu{i}", q(T.leaky_relu(d, 0.2)))
    logits = T.conv2d(d, p[f"{prefix}/logits/kernel"], p[f"{prefix}/logits/bias"])
    if sigmoid:
        logits = torch.sigmoid(logits)
    return _rec(acts, f"{prefix}/out", logits)


# ----------------------------------------------------------------------------- Autoencoder
AE_CONVS = [  # (name, cin, cout) in call order, autoencoder.py:150-186
    ("conv1", 3, 32), ("conv1b", 32, 32), ("conv2", 32, 44), ("conv3", 44, 56), ("conv4", 56, 76),
    ("conv5", 76, 100), ("conv6", 176, 152), ("conv6b", 152, 152), ("conv7", 208, 112),
    ("conv7b", 112, 112), ("conv8", 156, 84), ("conv8b", 84, 84), ("conv9", 116, 64),
    ("conv9b", 64, 64), ("conv10", 67, 64), ("conv10b", 64, 32), ("conv11", 32, 3),
]


def autoencoder_generator(p, x, training=True, state_out=None, acts=None):
    """autoencoder.py:89-188."""
    def conv(t, name, relu=True):
        y = T.conv2d(t, p[f"g/{name}/kernel"], p[f"g/{name}/bias"])
        return _rec(acts, f"g/{name}", torch.relu(y) if relu else torch.tanh(y))

    def upcat(a, b):                                                         # :113-136
        return torch.cat([torch.relu(T.upsample2x_nearest(a)), b], dim=3)

    c1 = conv(x, "conv1"); c1b = conv(c1, "conv1b"); p1 = T.max_pool2x2(c1b)
    c2 = conv(p1, "conv2"); p2 = T.max_pool2x2(c2)
    c3 = conv(p2, "conv3"); p3 = T.max_pool2x2(c3)
    c4 = conv(p3, "conv4"); p4 = T.max_pool2x2(c4)
    c5 = conv(p4, "conv5"); p5 = T.max_pool2x2(c5)
    t = conv(conv(upcat(p5, p4), "conv6"), "conv6b")
    t = conv(conv(upcat(t, p3), "conv7"), "conv7b")
    t = conv(conv(upcat(t, p2), "conv8"), "conv8b")
    t = conv(conv(upcat(t, p1), "conv9"), "conv9b")
    t = conv(conv(upcat(t, x), "conv10"), "conv10b")
    return conv(t, "conv11", relu=False)


# ----------------------------------------------------------------------------- Fast-SRGAN
def fsrgan_generator(p, x, training=True, state_out=None, acts=None, n_blocks=6):
    """fsrgan.py:99-220."""
    def bn(t, name, momentum=0.99):
        return T.batch_norm(t, p, name, training, state_out, momentum=momentum, eps=1e-3)

    c1 = T.conv2d(x, p["g/c1/kernel"], p["g/c1/bias"])                      # :198
    c1 = T.prelu(bn(c1, "g/c1_bn"), p["g/c1_prelu/alpha"])                   # :199-200
    _rec(acts, "g/c1", c1)
    r = c1
    for i in range(n_blocks):                                                # residual_block :112-176
        t = r
        if i:
            t = T.conv2d(t, p[f"g/b{i}/expand/kernel"], p[f"g/b{i}/expand/bias"])
            t = torch.relu(bn(t, f"g/b{i}/expand_bn", 0.999))
        t = T.depthwise_conv2d(t, p[f"g/b{i}/dw/kernel"], p[f"g/b{i}/dw/bias"])
        t = torch.relu(bn(t, f"g/b{i}/dw_bn", 0.999))
        t = T.conv2d(t, p[f"g/b{i}/project/kernel"], p[f"g/b{i}/project/bias"])
        t = bn(t, f"g/b{i}/project_bn", 0.999)
        r = _rec(acts, f"g/b{i}/add", r + t)
    c2 = T.conv2d(r, p["g/c2/kernel"], p["g/c2/bias"])                      # :208
    c2 = _rec(acts, "g/c2_add", bn(c2, "g/c2_bn") + c1)                      # :209-210
    u = c2
    for j in range(2):                                                       # :213-214
        u = T.conv2d(u, p[f"g/up{j}/conv/kernel"], p[f"g/up{j}/conv/bias"])
        u = T.prelu(T.depth_to_space(u, 2), p[f"g/up{j}/prelu/alpha"])
        _rec(acts, f"g/up{j}/prelu", u)
    out = T.conv2d(u, p["g/conv_out/kernel"], p["g/conv_out/bias"])          # :217 (3x3)
    return _rec(acts, "g/tanh", torch.tanh(out))


# ----------------------------------------------------------------------------- pix2pix
P2P_DOWN = [64, 128, 256, 512, 512, 512, 512, 512]
P2P_UP = [512, 512, 512, 512, 256, 128, 64]


def pix2pix_generator(p, x, training=True, state_out=None, acts=None, dropout_masks=None):
    """pix2pix.py:144-192.  `dropout_masks[i]` (bool keep-mask, NHWC) stands in for Dropout(0.5)'s
    RNG on up-blocks 0..2 when training; kept units are scaled by 2."""
    skips = []
    t = x
    for i, _ in enumerate(P2P_DOWN):                                         # downsample :110-123
        t = T.conv2d(t, p[f"g/down{i}/conv/kernel"], None, stride=2)
        if i > 0:
            t = T.batch_norm(t, p, f"g/down{i}/bn", training, state_out)
        t = _rec(acts, f"g/down{i}", T.leaky_relu(t, 0.3))
        skips.append(t)
    skips = list(reversed(skips[:-1]))
    for i, _ in enumerate(P2P_UP):                                           # upsample :125-142
        t = T.conv2d_transpose(t, p[f"g/up{i}/convt/kernel"], None, stride=2)
        t = T.batch_norm(t, p, f"g/up{i}/bn", training, state_out)
        if i < 3 and training:
            keep = dropout_masks[i].to(t.dtype)
            t = t * keep * 2.0
        t = torch.relu(t)
        t = _rec(acts, f"g/up{i}", torch.cat([t, skips[i]], dim=3))          # :188
    out = T.conv2d_transpose(t, p["g/last/kernel"], p["g/last/bias"], stride=2)
    return _rec(acts, "g/tanh", torch.tanh(out))


def pix2pix_discriminator(p, inp, tar, training=True, state_out=None, acts=None):
    """pix2pix.py:194-220 (PatchGAN on concat(input, target))."""
    t = torch.cat([inp, tar], dim=3)
    for i, _ in enumerate([64, 128, 256], start=1):
        t = T.conv2d(t, p[f"d/down{i}/conv/kernel"], None, stride=2)
        if i > 1:
            t = T.batch_norm(t, p, f"d/down{i}/bn", training, state_out)
        t = _rec(acts, f"d/down{i}", T.leaky_relu(t, 0.3))
    t = T.conv2d(t, p["d/conv4/kernel"], None, stride=1, padding=((1, 1), (1, 1)))   # ZeroPadding2D + VALID
    t = T.leaky_relu(T.batch_norm(t, p, "d/bn4", training, state_out), 0.3)
    _rec(acts, "d/conv4", t)
    t = T.conv2d(t, p["d/last/kernel"], p["d/last/bias"], stride=1, padding=((1, 1), (1, 1)))
    return _rec(acts, "d/out", t)


# ----------------------------------------------------------------------------- VGG19 trunk
VGG_CFG = [(1, 2, 64), (2, 2, 128), (3, 4, 256), (4, 4, 512), (5, 4, 512)]  # (block, convs, filters)


def vgg19_features(p, x):
    """keras.applications.VGG19(include_top=False) up to block5_conv4 (srgan.py:77-93); ReLU after
    every conv including the last (the Keras layer carries activation='relu')."""
    t = x
    for blk, n, _ in VGG_CFG:
        for c in range(1, n + 1):
            t = torch.relu(T.conv2d(t, p[f"vgg/block{blk}_conv{c}/kernel"], p[f"vgg/block{blk}_conv{c}/bias"]))
        if blk < 5:
            t = T.max_pool2x2(t)
    return t


def vgg_preprocess(x):
    """vgg19.preprocess_input(((x+1)*255)/2) in 'caffe' mode (srgan.py:71-72): RGB->BGR, mean-subtract."""
    x = ((x + 1.0) * 255.0) / 2.0
    x = x.flip(dims=(3,))
    mean = torch.tensor([103.939, 116.779, 123.68], dtype=x.dtype)
    return x - mean


def content_loss(p, target, gen_output):
    """srgan.py:69-75."""
    gf = vgg19_features(p, vgg_preprocess(gen_output)) / 12.75
    tf_ = vgg19_features(p, vgg_preprocess(target)) / 12.75
    return T.mse(tf_, gf)
