"""Generator / discriminator / VGG19 graphs of the reference, expressed as calls into the tape
engine (which calls the CUDA kernels).  Each class is a callable `net(x, training=bool) -> Var`
like the Keras models it replaces.
"""
from __future__ import annotations

import torch

from .engine import Engine, Var
from .params import ParamSet


class _Net:
    def __init__(self, engine: Engine, pset: ParamSet):
        self.E, self.p = engine, pset

    @property
    def trainable_variables(self):
        return list(self.p.params.values())

    def _in(self, x):
        return x if isinstance(x, Var) else self.E.input(x)

    def _bn(self, name, training, momentum=0.99, eps=1e-3):
        """`bn=` argument of Engine.conv2d for a conv whose output feeds BatchNormalization `name` (same momentum / eps as
        the bn_act call that follows): the conv epilogue produces the batch statistics and its last CTA finalises them."""
        if training:
            return (self.p, name, momentum, eps)
        # inference (infer_video.py:146, training=False): BatchNorm is an affine map of constants -> folded into the conv
        return ("fold", self.p, name, eps)


class SRGANGenerator(_Net):
    """srgan.py:129-185."""

    def __init__(self, engine, pset, scale=4):
        super().__init__(engine, pset)
        self.scale = scale
        # receptive-field radius at inference in input pixels, from the kernel sizes (FrameRunner.compute_size): every convolution
        # is stride 1, the BatchNorms are per-pixel, an up-sampling block halves the reach of the layers behind it
        half = lambda name: (pset[name].shape[0] - 1) / 2.0
        r = half("g/conv_in/kernel") + sum(half(f"g/res{i}/conv1/kernel") + half(f"g/res{i}/conv2/kernel") for i in range(16))
        r += half("g/conv_post/kernel") + sum(half(f"g/up{j}/conv/kernel") / 2 ** j for j in range(scale // 2))
        self.receptive_radius = r + half("g/conv_out/kernel") / 2 ** (scale // 2)

    def __call__(self, x, training=True) -> Var:
        E, p = self.E, self.p
        n = E.conv2d(self._in(x), p["g/conv_in/kernel"], bn=self._bn("g/bn_in", training), post=dict(prelu=p["g/prelu_in/alpha"]))
        n = E.bn_act(n, p, "g/bn_in", training=training, prelu=p["g/prelu_in/alpha"])
        temp = E.mark("g/prelu_in", n)
        for i in range(16):
            nn = E.conv2d(n, p[f"g/res{i}/conv1/kernel"], bn=self._bn(f"g/res{i}/bn1", training), post=dict(act="relu"))
            nn = E.bn_act(nn, p, f"g/res{i}/bn1", training=training, act="relu")
            nn = E.conv2d(nn, p[f"g/res{i}/conv2/kernel"], bn=self._bn(f"g/res{i}/bn2", training), post=dict(residual=n))
            n = E.mark(f"g/res{i}/add", E.bn_act(nn, p, f"g/res{i}/bn2", training=training, residual=n))
        n2 = E.conv2d(n, p["g/conv_post/kernel"], bn=self._bn("g/bn_post", training), post=dict(residual=temp))
        n = E.mark("g/post_add", E.bn_act(n2, p, "g/bn_post", training=training, residual=temp))
        for j in range(self.scale // 2):
            n = E.mark(f"g/up{j}/prelu", E.conv2d_d2s_prelu(n, p[f"g/up{j}/conv/kernel"], p[f"g/up{j}/conv/bias"], p[f"g/up{j}/prelu/alpha"], training))
        # 1x1 conv + tanh, fp32 output ('generator_tanh', dtype float32, srgan.py:183)
        return E.mark("g/tanh", E.conv2d(n, p["g/conv_out/kernel"], p["g/conv_out/bias"], act="tanh", out_dtype=torch.float32))


class PatchDiscriminator(_Net):
    """srgan.py:232-272 = fsrgan.py:222-258 (logits) = autoencoder.py:190-229 (sigmoid)."""
    STRIDES = [1, 2, 1, 2, 1, 2, 1, 2]

    def __init__(self, engine, pset, sigmoid=False, prefix="d"):
        super().__init__(engine, pset)
        self.sigmoid, self.prefix = sigmoid, prefix

    def __call__(self, x, training=True) -> Var:
        E, p, px = self.E, self.p, self.prefix
        d = self._in(x)
        for i, s in enumerate(self.STRIDES, start=1):
            w, b = p[f"{px}/conv{i}/kernel"], p[f"{px}/conv{i}/bias"]
            if i == 1:
                d = E.conv2d(d, w, b, stride=s, act="lrelu", alpha=0.2)
            else:
                d = E.conv2d(d, w, b, stride=s, bn=self._bn(f"{px}/bn{i}", training, momentum=0.8), post=dict(act="lrelu", alpha=0.2))
                d = E.bn_act(d, p, f"{px}/bn{i}", training=training, momentum=0.8, act="lrelu", alpha=0.2)
            E.mark(f"{px}/lrelu{i}", d)
        return E.conv2d(d, p[f"{px}/logits/kernel"], p[f"{px}/logits/bias"], act="sigmoid" if self.sigmoid else None,
                        out_dtype=torch.float32)


class AutoencoderGenerator(_Net):
    """autoencoder.py:89-188."""

    def __call__(self, x, training=True) -> Var:
        E, p = self.E, self.p
        x = self._in(x)
        # bf16 tensor-core path: the odd widths (3/44/56/76/100/152/84 channels) stay zero-padded to multiples of 16 from layer
        # to layer (Engine.phys_pad) -- no pad / slice pass around the convolutions
        keep = E.use_umma and E.pad_rgb and E.phys_pad and E.act_dtype == torch.bfloat16
        xin = E.pad_input(x) if keep else E.cast(x, E.act_dtype)
        if keep:
            x = xin

        def conv(t, name, act="relu", out_dtype=None):
            return E.conv2d(t, p[f"g/{name}/kernel"], p[f"g/{name}/bias"], act=act, out_dtype=out_dtype, keep_padded=keep and act == "relu")

        def upcat(a, b):
            return E.upsample_concat(a, b)

        c1b = conv(conv(x, "conv1"), "conv1b"); p1 = E.maxpool2x2(c1b)
        p2 = E.maxpool2x2(conv(p1, "conv2"))
        p3 = E.maxpool2x2(conv(p2, "conv3"))
        p4 = E.maxpool2x2(conv(p3, "conv4"))
        p5 = E.maxpool2x2(conv(p4, "conv5"))
        t = conv(conv(upcat(p5, p4), "conv6"), "conv6b")
        t = conv(conv(upcat(t, p3), "conv7"), "conv7b")
        t = conv(conv(upcat(t, p2), "conv8"), "conv8b")
        t = conv(conv(upcat(t, p1), "conv9"), "conv9b")
        t = conv(conv(upcat(t, xin), "conv10"), "conv10b")
        return conv(t, "conv11", act="tanh", out_dtype=torch.float32)


class FastSRGANGenerator(_Net):
    """fsrgan.py:99-220."""

    def __init__(self, engine, pset, n_blocks=6):
        super().__init__(engine, pset)
        self.n_blocks = n_blocks
        # receptive-field radius at inference, in input pixels: c1 (1) + one depthwise 3x3 per block + c2 (1) + the first up-conv (1)
        # + the second at 2x (1/2) + the output conv at 4x (1/4); every other layer is per-pixel (FrameRunner.compute_size)
        self.receptive_radius = n_blocks + 3.75

    def __call__(self, x, training=True) -> Var:
        E, p = self.E, self.p
        c1 = E.conv2d(self._in(x), p["g/c1/kernel"], p["g/c1/bias"])
        c1 = E.bn_act(c1, p, "g/c1_bn", training=training, prelu=p["g/c1_prelu/alpha"])
        r = c1
        for i in range(self.n_blocks):
            t = r
            if i and not training:
                # inference: expand -> depthwise -> project -> add as one launch, the 192-channel intermediates never reach HBM
                fused = E.fsrgan_block_infer(r, p, f"g/b{i}")
                if fused is not None:
                    r = fused
                    continue
            if i:
                t = E.conv2d(t, p[f"g/b{i}/expand/kernel"], p[f"g/b{i}/expand/bias"], bn=self._bn(f"g/b{i}/expand_bn", training, momentum=0.999),
                             post=dict(act="relu"))
                t = E.bn_act(t, p, f"g/b{i}/expand_bn", training=training, momentum=0.999, act="relu")
            t = E.dwconv3x3(t, p[f"g/b{i}/dw/kernel"], p[f"g/b{i}/dw/bias"], bn=self._bn(f"g/b{i}/dw_bn", training, momentum=0.999),
                            post=dict(act="relu"))
            t = E.bn_act(t, p, f"g/b{i}/dw_bn", training=training, momentum=0.999, act="relu")
            t = E.conv2d(t, p[f"g/b{i}/project/kernel"], p[f"g/b{i}/project/bias"], bn=self._bn(f"g/b{i}/project_bn", training, momentum=0.999),
                         post=dict(residual=r))
            r = E.bn_act(t, p, f"g/b{i}/project_bn", training=training, momentum=0.999, residual=r)
        c2 = E.conv2d(r, p["g/c2/kernel"], p["g/c2/bias"], bn=self._bn("g/c2_bn", training), post=dict(residual=c1))
        u = E.bn_act(c2, p, "g/c2_bn", training=training, residual=c1)
        for j in range(2):
            u = E.conv2d_d2s_prelu(u, p[f"g/up{j}/conv/kernel"], p[f"g/up{j}/conv/bias"], p[f"g/up{j}/prelu/alpha"], training)
        if not training:
            # inference: one tensor-core product against all nine taps + nine shifted adds (and straight to the uint8 frame when asked)
            img = E.conv3x3_image_infer(u, p["g/conv_out/kernel"], p["g/conv_out/bias"], act="tanh")
            if img is not None:
                return img
        return E.conv2d(u, p["g/conv_out/kernel"], p["g/conv_out/bias"], act="tanh", out_dtype=torch.float32)


class VGG19Features(_Net):
    """keras.applications.VGG19 trunk to block5_conv4 on 'caffe'-preprocessed input (srgan.py:69-93)."""
    CFG = [(1, 2), (2, 2), (3, 4), (4, 4), (5, 4)]

    def __call__(self, x, training=False) -> Var:
        E, p = self.E, self.p
        t = self._in(x)
        for blk, n in self.CFG:
            for c in range(1, n + 1):
                last = blk == 5 and c == n
                t = E.conv2d(t, p[f"vgg/block{blk}_conv{c}/kernel"], p[f"vgg/block{blk}_conv{c}/bias"], act="relu",
                             out_dtype=torch.float32 if last else None)
            if blk < 5:
                t = E.maxpool2x2(t)
        return t


class Pix2PixGenerator(_Net):
    """pix2pix.py:144-192: 8 x (Conv 4x4 s2 [+BN] + LeakyReLU(0.3)) down, 7 x (Conv2DTranspose 4x4 s2 + BN
    [+Dropout 0.5 on the first three] + ReLU, concat skip) up, Conv2DTranspose -> 3 + tanh."""
    DOWN = [64, 128, 256, 512, 512, 512, 512, 512]
    UP = [512, 512, 512, 512, 256, 128, 64]

    def __init__(self, engine, pset, dropout_seed=7):
        super().__init__(engine, pset)
        self.dropout_seed = dropout_seed

    def __call__(self, x, training=True, pass_id=0) -> Var:
        """`pass_id` separates the dropout streams of the two generator passes of one train step
        (gen_output and the identity term, pix2pix.py:90)."""
        E, p = self.E, self.p
        t = E.cast(self._in(x), E.act_dtype)
        nd, nu = len(self.DOWN), len(self.UP)
        N, H, W, _ = t.shape
        inplace = E.inplace_concat and E.bf16 and training     # (inference folds BatchNorm into the convolutions, which own their outputs)
        cat = [None] * nu
        if inplace:
            # concat([up_k, down_{nd-2-k}]) (pix2pix.py:188): both halves are written straight into the concat buffer by the layers
            # that produce them (`out=`), and their gradients are read as slices of its gradient: no copy in either direction
            for k in range(nu):
                j = nd - 2 - k
                cat[k] = E.concat_buffer((N, H >> (j + 1), W >> (j + 1), self.UP[k] + self.DOWN[j]), E.act_dtype, (self.UP[k], self.DOWN[j]))
        skips = []
        for i in range(nd):
            w = p[f"g/down{i}/conv/kernel"]
            dst = cat[nd - 2 - i][1][1] if (inplace and i <= nd - 2) else None
            if i == 0:
                t = E.conv2d(t, w, None, stride=2, act="lrelu", alpha=0.3, out=dst)
            else:
                t = E.conv2d(t, w, None, stride=2, bn=self._bn(f"g/down{i}/bn", training), post=dict(act="lrelu", alpha=0.3))
                t = E.bn_act(t, p, f"g/down{i}/bn", training=training, act="lrelu", alpha=0.3, out=dst)
            skips.append(t)
        skips = list(reversed(skips[:-1]))
        for i in range(nu):
            t = E.conv2d_transpose(t, p[f"g/up{i}/convt/kernel"], None, stride=2)
            drop = self.dropout_seed if (i < 3 and training) else None
            t = E.bn_act(t, p, f"g/up{i}/bn", training=training, act="relu", dropout_seed=drop,
                         dropout_offset=(pass_id * 3 + i) << 24, step_counter=p.opt_state if p.trainable else None,
                         out=cat[i][1][0] if inplace else None)
            t = E.concat_views(cat[i][0], [t, skips[i]]) if inplace else E.concat([t, skips[i]])
        return E.conv2d_transpose(t, p["g/last/kernel"], p["g/last/bias"], stride=2, act="tanh", out_dtype=torch.float32)


class Pix2PixDiscriminator(_Net):
    """pix2pix.py:194-220: PatchGAN on concat(input, target); ZeroPadding2D + VALID 4x4 convs at the end."""

    def __call__(self, inputs, training=True) -> Var:
        E, p = self.E, self.p
        inp, tar = inputs
        t = E.concat([E.cast(self._in(inp), E.act_dtype), E.cast(self._in(tar), E.act_dtype)])
        t = E.conv2d(t, p["d/down1/conv/kernel"], None, stride=2, act="lrelu", alpha=0.3)
        for i in (2, 3):
            t = E.conv2d(t, p[f"d/down{i}/conv/kernel"], None, stride=2, bn=self._bn(f"d/down{i}/bn", training), post=dict(act="lrelu", alpha=0.3))
            t = E.bn_act(t, p, f"d/down{i}/bn", training=training, act="lrelu", alpha=0.3)
        t = E.conv2d(t, p["d/conv4/kernel"], None, stride=1, padding=((1, 1), (1, 1)), bn=self._bn("d/bn4", training), post=dict(act="lrelu", alpha=0.3))
        t = E.bn_act(t, p, "d/bn4", training=training, act="lrelu", alpha=0.3)
        return E.conv2d(t, p["d/last/kernel"], p["d/last/bias"], stride=1, padding=((1, 1), (1, 1)), out_dtype=torch.float32)
