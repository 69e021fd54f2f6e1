"""Parameter arenas and the reference's initialisers.

All trainable variables of one network live in ONE flat fp32 device buffer (theta) with matching
flat grad / Adam-m / Adam-v buffers, so the optimiser is a single fused kernel and the gradient
all-reduce is one bucket.  Tensors are stored in Keras layouts (Conv2D [kh,kw,Cin,Cout],
Conv2DTranspose [kh,kw,Cout,Cin], DepthwiseConv2D [kh,kw,C,1], PReLU [C]) so trained Keras
weights can be loaded verbatim.  BatchNorm moving statistics are non-trainable state in a
separate flat buffer.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import _lib

ALIGN = 64  # elements; keeps every tensor 256-byte aligned inside the arena


class Param:
    """A named slice of a parameter arena."""

    __slots__ = ("name", "shape", "offset", "numel", "group", "owner", "packed_fwd", "packed_dgrad", "trainable", "pack_pad", "pack_seg")

    def __init__(self, name, shape, offset, group, owner, trainable=True):
        self.name, self.shape, self.offset, self.group, self.owner = name, tuple(shape), offset, group, owner
        self.numel = int(math.prod(shape))
        self.packed_fwd = None
        self.packed_dgrad = None
        self.pack_pad = None          # (cin_pad, cout_pad) when the tensor-core copies are zero-padded to 16 channels
        self.pack_seg = None          # (seg_log, seg_phys): two-segment input-channel axis (dg_umma_pack_weights_seg), (0, 0) = one
        self.trainable = trainable

    @property
    def data(self) -> torch.Tensor:
        buf = self.owner.theta if self.trainable else self.owner.state
        return buf[self.offset:self.offset + self.numel].view(self.shape)

    @property
    def grad(self) -> torch.Tensor:
        return self.owner.grad[self.offset:self.offset + self.numel].view(self.shape)


class ParamSet:
    """Flat arenas for one network ('g', 'd' or 'vgg')."""

    def __init__(self, group: str, tensors: "OrderedDict[str, torch.Tensor]", device, trainable=True):
        self.group = group
        self.params: "OrderedDict[str, Param]" = OrderedDict()
        self.states: "OrderedDict[str, Param]" = OrderedDict()
        off_t = off_s = 0
        for name, t in tensors.items():
            is_state = name.endswith(("moving_mean", "moving_variance"))
            if is_state:
                self.states[name] = Param(name, t.shape, off_s, group, self, trainable=False)
                off_s += -(-t.numel() // ALIGN) * ALIGN
            else:
                self.params[name] = Param(name, t.shape, off_t, group, self)
                off_t += -(-t.numel() // ALIGN) * ALIGN
        self.numel = off_t
        self.theta = torch.zeros(max(off_t, ALIGN), dtype=torch.float32, device=device)
        self.state = torch.zeros(max(off_s, ALIGN), dtype=torch.float32, device=device)
        self.trainable = trainable
        if trainable:
            self.grad = torch.zeros_like(self.theta)
            self.m = torch.zeros_like(self.theta)
            self.v = torch.zeros_like(self.theta)
            # int64 iterations | float lr_t | pad  (see dg_adam_step)
            self.opt_state = torch.zeros(2, dtype=torch.int64, device=device)
        self.load(tensors)

    def __getitem__(self, name) -> Param:
        return self.params[name] if name in self.params else self.states[name]

    def __contains__(self, name):
        return name in self.params or name in self.states

    def load(self, tensors):
        """Copies host (or device) tensors by name into the arenas."""
        self.version = getattr(self, "version", 0) + 1      # derived copies (BatchNorm-folded inference kernels) are rebuilt
        for name, t in tensors.items():
            p = self[name]
            assert tuple(t.shape) == p.shape, f"{name}: shape {tuple(t.shape)} != {p.shape}"
            p.data.copy_(t.to(torch.float32))

    def export(self) -> "OrderedDict[str, torch.Tensor]":
        out = OrderedDict()
        for name, p in list(self.params.items()) + list(self.states.items()):
            out[name] = p.data.detach().clone().cpu()
        return out

    def grads(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((n, p.grad.detach().clone().cpu()) for n, p in self.params.items())

    def repack(self, lib, ctx, stream):
        """Refreshes the bf16 K-major copies of every conv kernel the tensor-core path consumes — one launch
        for the whole network (device table of (src, dst, geometry) entries, rebuilt when the set changes)."""
        import struct
        ents = []
        for p in self.params.values():
            for mode, t in ((0, p.packed_fwd), (1, p.packed_dgrad)):
                if t is None:
                    continue
                kh, kw, cin, cout = p.shape
                cin_p, cout_p = p.pack_pad or (cin, cout)
                kdim = cin_p if mode == 0 else cout_p
                kc = 64 if kdim % 64 == 0 else (32 if kdim % 32 == 0 else 16)
                sl, sp = p.pack_seg or (0, 0)
                ents.append((p.data.data_ptr(), t.data_ptr(), kh * kw, cin_p, cout_p, kc, mode, cin, cout, sl | (sp << 16)))
        if not ents:
            return
        key = tuple(ents)
        if getattr(self, "_pack_key", None) != key:
            blob = b"".join(struct.pack("<QQiiiiiiii", s, d, taps, cin, cout, kc, mode, cin_s, cout_s, seg)
                            for s, d, taps, cin, cout, kc, mode, cin_s, cout_s, seg in ents)
            self._pack_table = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(self.theta.device)
            self._pack_key = key
        _lib.check(lib.dg_umma_pack_weights_batch(ctx, self._pack_table.data_ptr(), len(ents), stream))


# ------------------------------------------------------------------------------------------------
# initialisers (distributional parity with the reference; tests inject these same tensors into the oracle)

def _normal(gen, shape, mean, std):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std + mean


def _trunc_normal(gen, shape, std):
    """Keras VarianceScaling 'truncated_normal': N(0, std/0.8796) resampled into +-2 sigma."""
    s = std / 0.87962566103423978
    t = torch.randn(shape, generator=gen, dtype=torch.float32)
    for _ in range(8):
        bad = t.abs() > 2
        if not bad.any():
            break
        t = torch.where(bad, torch.randn(shape, generator=gen, dtype=torch.float32), t)
    return t.clamp_(-2, 2) * s


def _glorot_uniform(gen, shape):
    kh, kw, a, b = shape
    fan_in, fan_out = kh * kw * a, kh * kw * b
    if b == 1 and len(shape) == 4 and a > 1:  # depthwise [kh,kw,C,1]: fan_in = kh*kw, fan_out = kh*kw (Keras)
        fan_in, fan_out = kh * kw * 1, kh * kw * 1
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * lim


def _bn(out, name, c, gen=None, gamma_std=None):
    out[f"{name}/gamma"] = _normal(gen, (c,), 1.0, gamma_std) if gamma_std else torch.ones(c)
    out[f"{name}/beta"] = torch.zeros(c)
    out[f"{name}/moving_mean"] = torch.zeros(c)
    out[f"{name}/moving_variance"] = torch.ones(c)


def init_patch_discriminator(seed=1, prefix="d"):
    """srgan.py:232-272 / fsrgan.py:222-258 / autoencoder.py:190-229: Keras defaults (glorot_uniform, zero bias)."""
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    chans = [3, 32, 32, 32, 32, 64, 64, 64, 64]
    for i in range(1, 9):
        p[f"{prefix}/conv{i}/kernel"] = _glorot_uniform(gen, (3, 3, chans[i - 1], chans[i]))
        p[f"{prefix}/conv{i}/bias"] = torch.zeros(chans[i])
        if i > 1:
            _bn(p, f"{prefix}/bn{i}", chans[i])
    p[f"{prefix}/logits/kernel"] = _glorot_uniform(gen, (1, 1, 64, 1))
    p[f"{prefix}/logits/bias"] = torch.zeros(1)
    return p


def init_srgan_generator(seed=0, scale=4):
    """srgan.py:129-185: kernels N(0,.02), BN gamma N(1,.02), PReLU alpha 0, biases 0."""
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    p["g/conv_in/kernel"] = _normal(gen, (3, 3, 3, 64), 0.0, 0.02)
    _bn(p, "g/bn_in", 64, gen, 0.02)
    p["g/prelu_in/alpha"] = torch.zeros(64)
    for i in range(16):
        for j in (1, 2):
            p[f"g/res{i}/conv{j}/kernel"] = _normal(gen, (3, 3, 64, 64), 0.0, 0.02)
            _bn(p, f"g/res{i}/bn{j}", 64, gen, 0.02)
    p["g/conv_post/kernel"] = _normal(gen, (3, 3, 64, 64), 0.0, 0.02)
    _bn(p, "g/bn_post", 64, gen, 0.02)
    for j in range(scale // 2):
        p[f"g/up{j}/conv/kernel"] = _normal(gen, (3, 3, 64, 256), 0.0, 0.02)
        p[f"g/up{j}/conv/bias"] = torch.zeros(256)
        p[f"g/up{j}/prelu/alpha"] = torch.zeros(64)
    p["g/conv_out/kernel"] = _normal(gen, (1, 1, 64, 3), 0.0, 0.02)
    p["g/conv_out/bias"] = torch.zeros(3)
    return p


AE_CONVS = [("conv1", 3, 32), ("conv1b", 32, 32), ("conv2", 32, 44), ("conv3", 44, 56), ("conv4", 56, 76),
            ("conv5", 76, 100), ("conv6", 176, 152), ("conv6b", 152, 152), ("conv7", 208, 112),
            ("conv7b", 112, 112), ("conv8", 156, 84), ("conv8b", 84, 84), ("conv9", 116, 64),
            ("conv9b", 64, 64), ("conv10", 67, 64), ("conv10b", 64, 32), ("conv11", 32, 3)]


def init_autoencoder_generator(seed=0):
    """autoencoder.py:89-188: he_normal for the ReLU convs, lecun_normal for the tanh head, zero biases."""
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, cin, cout in AE_CONVS:
        fan_in = 9 * cin
        std = math.sqrt((1.0 if name == "conv11" else 2.0) / fan_in)
        p[f"g/{name}/kernel"] = _trunc_normal(gen, (3, 3, cin, cout), std)
        p[f"g/{name}/bias"] = torch.zeros(cout)
    return p


def init_fsrgan_generator(seed=0, gf=32, n_blocks=6):
    """fsrgan.py:99-220: all Keras defaults (glorot_uniform kernels, zero biases, PReLU alpha 0)."""
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    p["g/c1/kernel"] = _glorot_uniform(gen, (3, 3, 3, gf)); p["g/c1/bias"] = torch.zeros(gf)
    _bn(p, "g/c1_bn", gf)
    p["g/c1_prelu/alpha"] = torch.zeros(gf)
    for i in range(n_blocks):
        c = gf
        if i:
            p[f"g/b{i}/expand/kernel"] = _glorot_uniform(gen, (1, 1, gf, 6 * gf)); p[f"g/b{i}/expand/bias"] = torch.zeros(6 * gf)
            _bn(p, f"g/b{i}/expand_bn", 6 * gf)
            c = 6 * gf
        p[f"g/b{i}/dw/kernel"] = _glorot_uniform(gen, (3, 3, c, 1)); p[f"g/b{i}/dw/bias"] = torch.zeros(c)
        _bn(p, f"g/b{i}/dw_bn", c)
        p[f"g/b{i}/project/kernel"] = _glorot_uniform(gen, (1, 1, c, gf)); p[f"g/b{i}/project/bias"] = torch.zeros(gf)
        _bn(p, f"g/b{i}/project_bn", gf)
    p["g/c2/kernel"] = _glorot_uniform(gen, (3, 3, gf, gf)); p["g/c2/bias"] = torch.zeros(gf)
    _bn(p, "g/c2_bn", gf)
    for j in range(2):
        p[f"g/up{j}/conv/kernel"] = _glorot_uniform(gen, (3, 3, gf, 4 * gf)); p[f"g/up{j}/conv/bias"] = torch.zeros(4 * gf)
        p[f"g/up{j}/prelu/alpha"] = torch.zeros(gf)
    p["g/conv_out/kernel"] = _glorot_uniform(gen, (3, 3, gf, 3)); p["g/conv_out/bias"] = torch.zeros(3)
    return p


P2P_DOWN = [64, 128, 256, 512, 512, 512, 512, 512]
P2P_UP = [512, 512, 512, 512, 256, 128, 64]


def init_pix2pix(seed=0):
    """pix2pix.py:106-226: kernels N(0,.02), no conv biases except the two heads."""
    gen = torch.Generator().manual_seed(seed)
    g, d = OrderedDict(), OrderedDict()
    cin = 3
    for i, f in enumerate(P2P_DOWN):
        g[f"g/down{i}/conv/kernel"] = _normal(gen, (4, 4, cin, f), 0.0, 0.02)
        if i > 0:
            _bn(g, f"g/down{i}/bn", f)
        cin = f
    skips = list(reversed(P2P_DOWN[:-1]))
    for i, f in enumerate(P2P_UP):
        g[f"g/up{i}/convt/kernel"] = _normal(gen, (4, 4, f, cin), 0.0, 0.02)   # [kh,kw,Cout,Cin]
        _bn(g, f"g/up{i}/bn", f)
        cin = f + skips[i]
    g["g/last/kernel"] = _normal(gen, (4, 4, 3, cin), 0.0, 0.02)
    g["g/last/bias"] = torch.zeros(3)
    cin = 6
    for i, f in enumerate([64, 128, 256], start=1):
        d[f"d/down{i}/conv/kernel"] = _normal(gen, (4, 4, cin, f), 0.0, 0.02)
        if i > 1:
            _bn(d, f"d/down{i}/bn", f)
        cin = f
    d["d/conv4/kernel"] = _normal(gen, (4, 4, 256, 512), 0.0, 0.02)
    _bn(d, "d/bn4", 512)
    d["d/last/kernel"] = _normal(gen, (4, 4, 512, 1), 0.0, 0.02)
    d["d/last/bias"] = torch.zeros(1)
    return g, d


VGG_CFG = [(1, 2, 64), (2, 2, 128), (3, 4, 256), (4, 4, 512), (5, 4, 512)]


def init_vgg19_synthetic(seed=7):
    """Seeded he-normal stand-in for keras.applications.VGG19(weights='imagenet') (srgan.py:86): the
    ImageNet file cannot be downloaded here; load real weights with ParamSet.load for true parity."""
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    cin = 3
    for blk, n, f in VGG_CFG:
        for c in range(1, n + 1):
            p[f"vgg/block{blk}_conv{c}/kernel"] = _normal(gen, (3, 3, cin, f), 0.0, math.sqrt(2.0 / (9 * cin)))
            p[f"vgg/block{blk}_conv{c}/bias"] = torch.zeros(f)
            cin = f
    # the first layer sees inputs of magnitude ~128 (caffe preprocessing); keep activations O(1)
    p["vgg/block1_conv1/kernel"] = p["vgg/block1_conv1/kernel"] / 64.0
    return p
