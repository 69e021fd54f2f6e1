"""Tape executor for the hot path: a tiny define-by-run graph whose nodes are calls into the
dg_b200 C ABI.  It stands where TensorFlow's GradientTape stands in the reference
(train_srgan.py:73-112): `forward` ops append nodes, `backward(seeds, group)` replays them in
reverse for one variable group ('g' or 'd'), exactly the two `tape.gradient` calls of the
reference's train_step.

All activation, gradient and workspace buffers come from a persistent pool keyed by the op's
position in the step, so addresses are identical every step and a whole train step can be
captured into one CUDA graph.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import weakref

import torch

from . import _lib
from ._lib import ACT, DgConvParams, check, tensor
from .params import Param


def same_pads(in_size: int, k: int, s: int):
    """TF 'SAME' (pad_before, pad_after)."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


class Var:
    """An NHWC activation on the tape."""
    __slots__ = ("t", "deps", "seq", "bn_part", "bn_done", "bn_applied", "bn_src", "bn_folded", "bn_folded_post", "segs", "relu_out", "n_cons", "n_mask")

    def __init__(self, t: torch.Tensor, deps=frozenset(), seq=-1):
        self.t, self.deps, self.seq = t, deps, seq
        self.bn_part = None     # (partials [blocks,2,C], blocks) when the producing conv reduced the BatchNorm statistics
        self.bn_done = None     # (name, scale, shift, mean, invstd) when the producing conv ALSO finalised them (last-CTA ticket)
        self.bn_applied = None  # (y_act, act, alpha, prelu, residual) when the producing conv ALSO applied BatchNorm + activation (+ skip)
        self.bn_folded = None   # name of the inference-mode BatchNorm (+ activation) already folded into the producing convolution
        self.bn_folded_post = False   # ... together with the PReLU / skip-add behind it
        self.relu_out = False   # output of a convolution with a ReLU epilogue (values >= 0)
        self.n_cons = 0         # tape nodes (and gradient seeds) that consume this Var ...
        self.n_mask = 0         # ... of which those whose input gradient already carries the factor (t > 0): when ALL do, the ReLU
                                # convolution that produced t skips its own dg_act_bwd_from_output pass (autoencoder.py:95-136)
        self.segs = None        # physically padded channels: ((logical, physical), ...) segments of the channel axis of `t`, each
                                # holding `logical` real channels followed by zeros; None = dense (see Engine.phys_pad)
        self.bn_src = None      # on the OUTPUT of a training-mode bn_act: (seq, x, scale, shift, mean, act code, alpha) -- lets the input-gradient
                                # convolution that produces this Var's gradient also reduce the BatchNorm-backward sums (dg_umma_conv2d_dgrad_fused)

    @property
    def shape(self):
        return tuple(self.t.shape)


class Node:
    __slots__ = ("seq", "inputs", "out", "group", "bwd", "fusable", "params", "wkey")

    def __init__(self, seq, inputs, out, group, bwd, fusable=False, params=(), wkey=None):
        self.seq, self.inputs, self.out, self.group, self.bwd = seq, inputs, out, group, bwd
        self.wkey = wkey          # geometry key of the node's weight gradient (Engine._wgrad_enqueue), None: not a batchable convolution
        self.params = tuple(q for q in params if q is not None)   # variables whose gradient slice this node's bwd writes
        self.fusable = fusable    # bwd accepts fctx= (fused skip-add / BatchNorm-backward sums in the dgrad epilogue)


class _ProfLib:
    """Pass-through to the C library that, when `engine.prof_calls` is a list, brackets every dg_* call with CUDA
    events (tools/step_profile.py: warm, in-order timing of a whole step by entry point)."""

    def __init__(self, lib, engine):
        self._lib, self._engine = lib, engine

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        eng = self._engine

        def call(*a):
            rec = eng.prof_calls
            if rec is None or name.endswith(("_bytes", "_supported")) or not name.startswith("dg_"):
                return fn(*a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a)
            e1.record()
            rec.append((name, e0, e1))
            return r

        return call


class Engine:
    def __init__(self, device=None, bf16: bool = False):
        self.prof_calls = None
        self.lib = _ProfLib(_lib.load(), self)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._in_side = False
        # every Engine owns its contexts (ticket counters / reduction scratch are per context and assume ONE stream per
        # context): two engines in a process, or an engine next to direct C-ABI callers, never share counters
        self._ctx_main = _lib.new_ctx(self.device.index)
        weakref.finalize(self, _lib.destroy_ctx, self._ctx_main)
        self._ctx_side = None          # private context (reduction scratch) of the side stream, created on first use
        self.bf16 = bool(bf16)
        self.act_dtype = torch.bfloat16 if bf16 else torch.float32
        self.pool: dict = {}
        self.tape: list[Node] = []
        self.seq = 0
        self._ws = None
        self.written: set = set()      # param names whose grad slice was written this step
        self.complete: set = set()     # ... and whose every writer of the current backward() has been visited
        self.use_umma = bool(bf16) and bool(self.lib.dg_has_umma(self.ctx))
        self.use_umma_wgrad = self.use_umma
        self.fuse_bn_fwd = os.environ.get("DG_BN_FUSED_FWD", "0") == "1"   # one-launch BN forward: measured 1 % slower in the step graph
        self.fuse_conv_bn_stats = os.environ.get("DG_CONV_BN_STATS", "1") != "0"   # BN batch statistics from the conv epilogue
        # ... and their finalize by the conv's last CTA (dg_umma_conv2d_fwd_bn): OFF by default, measured slower in the step graph
        # (7.54 vs 7.45 ms, 403 vs 446 kernel nodes): the serial tail on one CTA costs more than the 8-block finalize launch
        self.fuse_conv_bn_finalize = os.environ.get("DG_CONV_BN_FINALIZE", "0") == "1"
        # conv + BatchNorm(train) + activation (+ skip-add) as ONE cooperative launch where all output tiles of a CTA fit TMEM
        # (dg_umma_conv2d_fwd_bn_act): the generator trunk and most discriminator layers lose the finalize and the apply launch.
        # OFF by default: measured SLOWER in the step graph (7.57 vs 7.42 ms, same box, job r2_03) -- the tensor pipe idles
        # through the grid barrier and pass 2, and a cooperative launch cannot overlap its neighbours' tails
        self.fuse_conv_bn_act = os.environ.get("DG_CONV_BN_ACT", "0") == "1"
        # input-gradient convolutions add the skip-connection gradient and reduce the sums of the BatchNorm backward pass in
        # their epilogue (dg_umma_conv2d_dgrad_fused + dg_bn_bwd_dx_from_partials): no add launch, one pass over dy / x
        # instead of two per BatchNorm of the generator trunk and of the stride-1 discriminator layers
        # Modes (DG_DGRAD_BN_BWD): 0 off; 1 skip-add only; 2 (default) skip-add + BatchNorm-backward sums.  Same-box A/B of the C3 step
        # (job r2_09): 7.21 ms (mode 2) vs 7.36 ms (mode 1).  The first version of mode 2 (one accumulator row per thread, 496 shuffles
        # per thread and tile) was SLOWER (7.70 vs 7.37 ms, job r2_05); the epilogue now reads the accumulator in mma-fragment layout
        self.fuse_dgrad_mode = int(os.environ.get("DG_DGRAD_BN_BWD", "2"))
        self.fuse_dgrad_bn_bwd = self.fuse_dgrad_mode != 0
        self.fuse_bn_finalize_apply = os.environ.get("DG_BN_FINALIZE_APPLY", "1") != "0"
        # odd channel counts (autoencoder.py:150-186: 44/56/76/100/152/84) stay zero-padded to multiples of 16 from layer to layer
        # instead of being padded and sliced around every convolution (Var.segs)
        self.phys_pad = os.environ.get("DG_PHYS_PAD", "1") != "0"
        self.fold_relu_bwd = os.environ.get("DG_FOLD_RELU_BWD", "1") != "0"    # see Var.n_mask
        self.narrow_store = os.environ.get("DG_NARROW_STORE", "1") != "0"
        self.fuse_d2s_infer = os.environ.get("DG_FUSE_D2S", "1") != "0"     # inference: depth_to_space + PReLU as the up-conv's store pattern
        self.inplace_concat = os.environ.get("DG_INPLACE_CONCAT", "1") != "0"   # pix2pix U-Net: layers write into their half of the concat buffer
        self.fuse_fsrgan_block = os.environ.get("DG_FSRGAN_BLOCK", "1") != "0"   # inference: a Fast-SRGAN inverted-residual block as one launch
        self.fuse_res_epilogue = os.environ.get("DG_FUSE_RES_EPILOGUE", "1") != "0"   # inference: PReLU / skip-add behind a folded BatchNorm in the conv epilogue
        self.concat_pad32 = os.environ.get("DG_CONCAT_PAD32", "1") != "0"     # U-Net concats of 16 (mod 32) channels get 16 zero channels more (32-channel K chunks)
        self.tapsum_infer = os.environ.get("DG_CONV_TAPSUM", "1") != "0"     # inference: the 32 -> 3 image convolution in tap-sum form (conv_tapsum.cu)
        self.frame_sink = None   # FrameRunner: dict(out=uint8 frame, h, w, scale, offset, clip, flip, done) consumed by conv3x3_image_infer
        # weight gradients of layers with identical geometry (the generator trunk's 32 identical convolutions, the real / fake passes of
        # one discriminator layer) are collected during backward() and launched up to `wgrad_batch` at a time (dg_umma_conv2d_wgrad_batch)
        self.wgrad_batch = max(1, min(4, int(os.environ.get("DG_WGRAD_BATCH", "4"))))
        self._wq: dict = {}             # geometry key -> [(x, dy, w, b, acc, flops, lin)]
        self._wq_names: set = set()     # parameters with a queued weight gradient
        self._complete_wait: set = set()
        self._in_backward = False
        self._bwd_tag = None
        self._wq_seen: dict = {}        # geometry key -> weight gradients seen so far in the current backward()
        self._wq_expect: dict = {}      # geometry key -> convolutions of that geometry on the tape whose variables the current backward() differentiates
        self.fold_bn_infer = os.environ.get("DG_FOLD_BN", "1") != "0"      # inference: BatchNorm folded into the producing conv / depthwise conv
        self._folded: dict = {}
        self.small_map_gemm = os.environ.get("DG_SMALL_MAP_GEMM", "1") != "0"      # weight gradients of <= 8x8 maps as one dense product (dg_im2col)   # dg_bn_act_fwd_from_partials instead of finalize + apply
        self._bwd_part: dict = {}       # (bn_act seq, tag) -> (partials, rows) left by the fused dgrad for that BatchNorm's backward
        # weight gradients run on a side stream: they only feed the optimiser, so their prologue/tail overlaps the
        # dgrad / BatchNorm chain of the backward pass (joined at the end of backward())
        self.wgrad_overlap = os.environ.get("DG_WGRAD_OVERLAP", "1") != "0"
        self._side_stream = torch.cuda.Stream(device=self.device)
        self._in_side = False
        self._side_dirty = False
        self._ws_side = None
        self.pad_rgb = True            # channel counts that are not multiples of 16 (RGB sides, the autoencoder's 44/56/76/100/67/...)
                                       # run on the tensor cores through zero padding to the next multiple of 16
        self.launches = 0
        self.record = None             # dict name -> Var when a test wants per-layer activations
        self._cap: dict = {}
        self.grad_record = None        # dict seq -> dL/dVar of the generator pass when a test wants per-layer gradients
        self.prof = None               # list of (kind, flops, ev0, ev1) when bench.py profiles a step

    def _timed(self, kind: str, flops: float, call):
        """Runs `call()`; when profiling, brackets it with CUDA events on the launching stream."""
        if self.prof is None:
            return call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = call()
        e1.record()
        self.prof.append((kind, flops, e0, e1))
        return r

    def mark(self, name: str, v: "Var") -> "Var":
        if self.record is not None:
            self.record[name] = v
        return v

    # ------------------------------------------------------------------ plumbing
    def new_step(self):
        self.tape.clear()
        self.seq = 0
        self.written.clear()
        self.launches = 0

    def buf(self, key, shape, dtype) -> torch.Tensor:
        k = (key, tuple(shape), dtype)
        t = self.pool.get(k)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.pool[k] = t
        return t

    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._in_side:
            if self._ws_side is None or self._ws_side.numel() < nbytes:
                self._ws_side = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=self.device)
            return self._ws_side
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=self.device)
        return self._ws

    @property
    def ctx(self):
        if self._in_side:
            if self._ctx_side is None:
                self._ctx_side = _lib.new_ctx(self.device.index)
                weakref.finalize(self, _lib.destroy_ctx, self._ctx_side)
            return self._ctx_side
        return self._ctx_main

    def branch(self):
        """`with E.branch(): ...` runs an independent part of the forward graph (e.g. D(real) next to the generator) on
        the side stream; call E.join() before its results are consumed."""
        return self._side()

    def join(self):
        self._join_side()

    @contextlib.contextmanager
    def _side(self):
        """Runs the enclosed weight-gradient launches on the side stream, ordered after everything enqueued so far."""
        if not self.wgrad_overlap or self._in_side:
            yield
            return
        main = torch.cuda.current_stream()
        self._side_stream.wait_stream(main)
        with torch.cuda.stream(self._side_stream):
            self._in_side = True
            try:
                yield
            finally:
                self._in_side = False
        self._side_dirty = True

    def _join_side(self):
        if self._side_dirty:
            torch.cuda.current_stream().wait_stream(self._side_stream)
            self._side_dirty = False

    def _next(self):
        self.seq += 1
        return self.seq

    def _push(self, inputs, out: Var, group, bwd, fusable=False, params=(), masks=(), wkey=None):
        """`masks`: the inputs whose gradient this node returns already multiplied by (input > 0)."""
        for v in inputs:
            v.n_cons += 1
        for v in masks:
            v.n_mask += 1
        self.tape.append(Node(out.seq, inputs, out, group, bwd, fusable, params, wkey))

    def _relu_mask_rows(self, dy_shape, dx_shape, lin) -> int:
        """> 0 when the input-gradient launch of this layer can store its result already multiplied by (input > 0)
        (dg_umma_conv2d_dgrad_relu_mask: stride 1, staged epilogue, 32- or 64-channel N block).  OFF by default (DG_RELU_MASK_DGRAD=1
        switches it on): measured SLOWER on the autoencoder, 11.72 against 11.50 ms per step -- the mma-fragment epilogue that pays for
        itself when it also produces the BatchNorm-backward sums takes ~385 us on the 256x256 layers, more than the plain input
        gradient plus a separate ReLU-backward pass (profiles/step_profile_r2b_ae_relu_mask.log)."""
        key = ("relumask", tuple(dy_shape), tuple(dx_shape), lin.kh, lin.kw, lin.stride, lin.pad_t, lin.pad_l)
        rows = self._cap.get(key)
        if rows is None:
            rows = 0
            if lin.stride == 1 and self.use_umma and self.fold_relu_bwd and os.environ.get("DG_RELU_MASK_DGRAD", "0") != "0":
                BF = _lib.DG_BF16
                d_dy = _lib.DgTensor(1 << 20, BF, dy_shape[0], dy_shape[1], dy_shape[2], dy_shape[3], dy_shape[3], 0)
                d_dx = _lib.DgTensor(1 << 20, BF, dx_shape[0], dx_shape[1], dx_shape[2], dx_shape[3], dx_shape[3], 0)
                rows = int(self.lib.dg_umma_conv2d_dgrad_fused_blocks(self.ctx, C.byref(d_dy), C.byref(d_dx), C.byref(lin)))
            self._cap[key] = rows
        return rows

    def _relu_bwd_folded(self, out: Var, act) -> bool:
        """True when every consumer of this ReLU convolution's output returns a gradient that already carries the ReLU mask."""
        return self.fold_relu_bwd and act == "relu" and out.n_cons > 0 and out.n_mask == out.n_cons

    @staticmethod
    def _deps(inputs, group=None):
        d = frozenset()
        for v in inputs:
            d = d | v.deps
        return d | {group} if group else d

    def _acc_flag(self, p: Param) -> int:
        if p.name in self.written:
            return 1
        self.written.add(p.name)
        return 0

    def input(self, t: torch.Tensor) -> Var:
        return Var(t.contiguous(), frozenset(), self._next())

    @property
    def st(self):
        return _lib.stream_ptr()

    # ------------------------------------------------------------------ convolution
    def _conv_geom(self, H, W, kh, kw, stride, padding):
        if padding == "same":
            (pt, pb), (pl, pr) = same_pads(H, kh, stride), same_pads(W, kw, stride)
        elif padding == "valid":
            pt = pb = pl = pr = 0
        else:
            (pt, pb), (pl, pr) = padding
        Ho = (H + pt + pb - kh) // stride + 1
        Wo = (W + pl + pr - kw) // stride + 1
        return pt, pl, Ho, Wo

    def _umma_ok(self, x: torch.Tensor, cin, cout, kh, kw, stride, H, W):
        return (self.use_umma and x.dtype == torch.bfloat16 and cin % 16 == 0 and cout % 16 == 0 and kh * kw <= 16
                and stride in (1, 2) and (stride == 1 or (H % 2 == 0 and W % 2 == 0)))

    def _umma_supported(self, kind: str, a: torch.Tensor, b: torch.Tensor, cp: DgConvParams) -> bool:
        """Asks the library whether the tensor-core kernel has a tile configuration for this layer (shared-memory
        fit); layers it cannot take yet (e.g. pix2pix's 4x4 stride-2 convs with >= 128 channels) are routed to the
        CUDA-core kernels EXPLICITLY here -- the C ABI itself never falls back."""
        key = (kind, tuple(a.shape), tuple(b.shape), cp.kh, cp.kw, cp.stride, cp.pad_t, cp.pad_l)
        r = self._cap.get(key)
        if r is None:
            ta, tb = tensor(a), tensor(b)
            fn = self.lib.dg_umma_conv2d_fwd_supported if kind == "fwd" else self.lib.dg_umma_conv2d_dgrad_supported
            r = bool(fn(self.ctx, C.byref(ta), C.byref(tb), C.byref(cp)))
            self._cap[key] = r
        return r

    def _packed(self, p: Param, mode: int) -> torch.Tensor:
        attr = "packed_fwd" if mode == 0 else "packed_dgrad"
        t = getattr(p, attr)
        if t is None:
            kh, kw, cin, cout = p.shape
            cin_p, cout_p = p.pack_pad or (cin, cout)
            t = torch.empty(kh * kw * cin_p * cout_p, dtype=torch.bfloat16, device=self.device)
            setattr(p, attr, t)
            sl, sp = p.pack_seg or (0, 0)
            check(self.lib.dg_umma_pack_weights_seg(self.ctx, p.data.data_ptr(), t.data_ptr(), kh, kw, cin, cout, cin_p, cout_p, sl, sp,
                                                    mode, self.st))
        return t

    @staticmethod
    def _norm_segs(segs):
        """Merges channel segments: a segment without padding joins the one that follows it ((64,64)+(3,16) = (67,80))."""
        out = []
        for l, p in segs:
            if out and out[-1][0] == out[-1][1]:
                out[-1] = (out[-1][0] + l, out[-1][1] + p)
            else:
                out.append((l, p))
        if all(l == p for l, p in out):
            return None
        return tuple(out)

    # ---- RGB-sided layers on the tensor cores through zero-padded 16-channel bf16 copies
    def _zeros(self, key, shape, dtype) -> torch.Tensor:
        k = (key, tuple(shape), dtype)
        t = self.pool.get(k)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self.pool[k] = t
        return t

    def _copy_vec(self, src: torch.Tensor, dst: torch.Tensor, n: int):
        """dst[:n] = src[:n] for 1-D fp32 vectors, through the C ABI (no torch kernel inside the captured step)."""
        ts = _lib.DgTensor(src.data_ptr(), _lib.DG_F32, 1, 1, 1, n, src.numel(), 0)
        td = _lib.DgTensor(dst.data_ptr(), _lib.DG_F32, 1, 1, 1, n, dst.numel(), 0)
        check(self.lib.dg_copy(self.ctx, C.byref(ts), C.byref(td), 0, self.st))

    def _out_or_buf(self, out, key, shape, dtype) -> torch.Tensor:
        """The caller's output view (`out=`: the layer writes straight into its channel slice of a U-Net concat buffer) or a pool buffer."""
        if out is None:
            return self.buf(key, shape, dtype)
        assert tuple(out.shape) == tuple(shape) and out.dtype == dtype, f"out= view {tuple(out.shape)} {out.dtype} != {tuple(shape)} {dtype}"
        return out

    def _conv2d_padded(self, x: Var, w: Param, b: Param | None, stride, pt, pl, Ho, Wo, act, alpha, out_dtype, keep=False, out=None):
        """conv2d when Cin or Cout is not a multiple of 16 (the 3-channel image side of srgan.py:154,182,236): operands are
        zero-padded to 16 channels in bf16 (dg_pad_channels / dg_umma_pack_weights_padded), the tcgen05 kernels run on the
        padded shapes, and results are sliced back (dg_copy views, dg_unpad_weight_grad).  Returns None when the tensor-core
        kernels have no tile configuration for the padded layer (the caller then uses the CUDA-core kernels)."""
        N, H, W, cin_p = x.shape
        kh, kw, cin, cout = w.shape
        segs = x.segs
        seg = (0, 0)
        if segs is None:
            cin_p = -(-cin // 16) * 16
        else:
            assert len(segs) <= 2 and x.t.dtype == torch.bfloat16, f"{w.name}: unsupported channel segments {segs}"
            if len(segs) == 2:
                seg = segs[0]
        cout_p = -(-cout // 16) * 16
        keep = keep and self.phys_pad and cout_p != cout and act in (None, "relu", "lrelu", "tanh")     # act(0) = 0 keeps the padding zero
        ydt = out_dtype or self.act_dtype
        lin = DgConvParams(kh, kw, stride, pt, pl, 0, 0.0)
        cp = DgConvParams(kh, kw, stride, pt, pl, ACT[act], float(alpha))
        BF = _lib.DG_BF16
        dummy = x.t.data_ptr()
        d_x = _lib.DgTensor(dummy, BF, N, H, W, cin_p, cin_p, 0)
        d_y = _lib.DgTensor(dummy, BF, N, Ho, Wo, cout_p, cout_p, 0)
        key = ("padded", N, H, W, cin_p, Ho, Wo, cout_p, kh, kw, stride, pt, pl)
        ok = self._cap.get(key)
        if ok is None:
            ok = (bool(self.lib.dg_umma_conv2d_fwd_supported(self.ctx, C.byref(d_x), C.byref(d_y), C.byref(lin))) and
                  bool(self.lib.dg_umma_conv2d_dgrad_supported(self.ctx, C.byref(d_y), C.byref(d_x), C.byref(lin))) and
                  self.lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(d_x), C.byref(d_y), C.byref(lin)) > 0)
            self._cap[key] = ok
        if not ok:
            return None
        seq = self._next()
        assert w.pack_pad in (None, (cin_p, cout_p)) and w.pack_seg in (None, seg), f"{w.name}: used with two different channel layouts"
        w.pack_pad, w.pack_seg = (cin_p, cout_p), seg
        assert out is None or (not keep and cout_p == cout), f"{w.name}: out= needs a 16-channel-multiple output"
        y = None if keep else self._out_or_buf(out, (seq, "y"), (N, Ho, Wo, cout), ydt)
        pad_in = segs is None and (cin_p != cin or x.t.dtype != torch.bfloat16)
        if pad_in:
            xin = self.buf((seq, "xpad"), (N, H, W, cin_p), torch.bfloat16)
            tsrc, tdst = tensor(x.t), tensor(xin)
            check(self.lib.dg_pad_channels(self.ctx, C.byref(tsrc), C.byref(tdst), self.st))
        else:
            xin = x.t
        # fewer than 16 output channels in fp32 (the RGB image): the epilogue stores the real channels densely (dg_umma_conv2d_fwd_narrow)
        narrow = self.narrow_store and not keep and cout_p == 16 and cout < 16 and ydt == torch.float32
        yp = y if ((cout_p == cout and not keep) or narrow) else self.buf((seq, "ypad"), (N, Ho, Wo, cout_p), ydt)
        bias = None
        if b is not None:
            if cout_p == cout:
                bias = b.data.data_ptr()
            else:
                bp = self._zeros((seq, "bias_p"), (cout_p,), torch.float32)
                self._copy_vec(b.data, bp, cout)
                bias = bp.data_ptr()
        txi, typ = tensor(xin), tensor(yp)
        flops = 2.0 * N * Ho * Wo * kh * kw * cin * cout      # ALGORITHMIC: the zero-padded channels are not counted
        pk = self._packed(w, 0)
        if narrow:
            if bias is None:
                bias = self._zeros((seq, "bias_p"), (cout_p,), torch.float32).data_ptr()
            self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_fwd_narrow(
                self.ctx, C.byref(txi), pk.data_ptr(), bias, C.byref(typ), C.byref(cp), self.st)))
        else:
            self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_fwd(
                self.ctx, C.byref(txi), pk.data_ptr(), bias, C.byref(typ), C.byref(cp), None, self.st)))
        if keep:
            y = yp
        elif yp is not y:
            tv, ty = tensor(yp, c=cout), tensor(y)
            check(self.lib.dg_copy(self.ctx, C.byref(tv), C.byref(ty), 0, self.st))
        out = Var(y, self._deps([x], w.group), seq)
        out.relu_out = act == "relu"
        if keep:
            out.segs = ((cout, cout_p),)
        # conv -> conv chain (autoencoder.py:95-104): x is the output of a ReLU convolution and this layer's input gradient is stored
        # already multiplied by (x > 0), so the producer needs no ReLU backward pass of its own (Var.n_mask)
        mask_in = (x.relu_out and not pad_in and x.t.dtype == torch.bfloat16 and
                   self._relu_mask_rows((N, Ho, Wo, cout_p), tuple(x.shape), lin) > 0)

        def bwd(gy: torch.Tensor, need_in, need_p, tag):
            dpre = gy
            if ACT[act] and not self._relu_bwd_folded(out, act):
                dpre = self.buf((seq, "dpre", tag), gy.shape, gy.dtype)
                tg, tyy, td = tensor(gy), tensor(y), tensor(dpre)
                check(self.lib.dg_act_bwd_from_output(self.ctx, C.byref(tg), C.byref(tyy), ACT[act], float(alpha), C.byref(td), self.st))
            if (cout_p != cout and not keep) or dpre.dtype != torch.bfloat16:
                dpp = self.buf((seq, "dpad", tag), (N, Ho, Wo, cout_p), torch.bfloat16)
                ts, td = tensor(dpre), tensor(dpp)
                check(self.lib.dg_pad_channels(self.ctx, C.byref(ts), C.byref(td), self.st))
            else:
                dpp = dpre
            tdp = tensor(dpp)
            if need_p:
                acc = self._acc_flag(w)
                if b is not None:
                    assert self._acc_flag(b) == acc
                with self._side():
                    nbytes = self.lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(txi), C.byref(tdp), C.byref(lin))
                    ws = self.workspace(nbytes)
                    dwp = self.buf((seq, "dw_pad", tag), (kh * kw * cin_p * cout_p,), torch.float32)
                    dbp = self.buf((seq, "db_pad", tag), (cout_p,), torch.float32) if b is not None else None
                    self._timed("umma_wgrad", flops, lambda: check(self.lib.dg_umma_conv2d_wgrad(
                        self.ctx, C.byref(txi), C.byref(tdp), dwp.data_ptr(), _lib.ptr(dbp), C.byref(lin), 0, ws.data_ptr(), nbytes, self.st)))
                    check(self.lib.dg_unpad_weight_grad_seg(self.ctx, dwp.data_ptr(), _lib.ptr(dbp), w.grad.data_ptr(),
                                                            b.grad.data_ptr() if b is not None else None, kh, kw, cin, cout, cin_p, cout_p,
                                                            seg[0], seg[1], acc, self.st))
            dx = None
            if need_in[0]:
                dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
                dxp = self.buf((seq, "dx_pad", tag), (N, H, W, cin_p), torch.bfloat16) if pad_in else dx
                tdx = tensor(dxp)
                pkd = self._packed(w, 1)
                if mask_in:
                    self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_dgrad_relu_mask(
                        self.ctx, C.byref(tdp), pkd.data_ptr(), C.byref(tdx), C.byref(lin), C.byref(txi), self.st)))
                else:
                    self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_dgrad(
                        self.ctx, C.byref(tdp), pkd.data_ptr(), None, C.byref(tdx), C.byref(lin), self.st)))
                if dxp is not dx:
                    tv, td = tensor(dxp, c=cin), tensor(dx)
                    check(self.lib.dg_copy(self.ctx, C.byref(tv), C.byref(td), 0, self.st))
            return [dx]

        self._push([x], out, w.group, bwd, params=(w, b), masks=[x] if mask_in else ())
        return out

    def _fold(self, w: Param, b: Param | None, pset, bname: str, eps: float, axis: int):
        """BatchNorm(training=False) folded into the kernel and bias that feed it: y = scale*(conv(x,w)+b) + shift with
        scale = gamma/sqrt(moving_var+eps), shift = beta - moving_mean*scale (srgan.py:155,163 at inference, infer_video.py:146).
        Returns (fp32 kernel, fp32 bias); cached until the parameter sets change (ParamSet.version).  Host-side torch
        arithmetic on a few thousand numbers, once per model load: plumbing, not on the per-frame path."""
        key = (w.name, bname)
        ver = (getattr(w.owner, "version", 0), getattr(pset, "version", 0))
        ent = self._folded.get(key)
        if ent is None or ent[0] != ver:
            scale = pset[bname + "/gamma"].data / torch.sqrt(pset[bname + "/moving_variance"].data + float(eps))
            shift = pset[bname + "/beta"].data - pset[bname + "/moving_mean"].data * scale
            shape = [1, 1, 1, 1]
            shape[axis] = -1
            wf = (w.data * scale.view(shape)).contiguous()
            bf = ((b.data * scale if b is not None else 0.0) + shift).contiguous()
            if ent is None:
                ent = [ver, wf, bf, None]
            else:           # keep the addresses (they may be baked into a captured graph)
                ent[1].copy_(wf); ent[2].copy_(bf); ent[0] = ver
                if ent[3] is not None:
                    kh, kw, cin, cout = w.shape
                    check(self.lib.dg_umma_pack_weights_padded(self.ctx, ent[1].data_ptr(), ent[3].data_ptr(), kh, kw, cin, cout, cin, cout, 0, self.st))
            self._folded[key] = ent
        return ent

    def conv2d(self, x: Var, w: Param, b: Param | None = None, *, stride=1, padding="same", act=None, alpha=0.0,
               out_dtype=None, bn: bool = False, post: dict | None = None, keep_padded: bool = False, out: torch.Tensor | None = None) -> Var:
        """keras Conv2D (+bias, +activation epilogue).  `keep_padded`: an output whose channel count is not a multiple of 16
        stays zero-padded (Var.segs) for consumers that understand it (conv2d, maxpool2x2, upsample_concat).  `bn=True`: a training-mode BatchNormalization consumes the result
        next, so the tensor-core epilogue also produces its batch-statistics partials (picked up by bn_act).
        `post` = dict(act=, alpha=, prelu=, residual=) describes what the bn_act call that FOLLOWS will do with the result; when
        the layer qualifies, convolution, statistics, normalisation, activation and skip-add run as one cooperative launch
        and bn_act only records the tape node."""
        N, H, W, Cin = x.shape
        kh, kw, cin, cout = w.shape
        if x.segs is not None:
            Cin = sum(l for l, _ in x.segs)
        assert cin == Cin, f"{w.name}: Cin {cin} != input {Cin}"
        pt, pl, Ho, Wo = self._conv_geom(H, W, kh, kw, stride, padding)
        if (self.use_umma and self.pad_rgb and (cin % 16 != 0 or cout % 16 != 0 or x.segs is not None) and kh * kw <= 16 and
                stride in (1, 2) and (stride == 1 or (H % 2 == 0 and W % 2 == 0))):
            r = self._conv2d_padded(x, w, b, stride, pt, pl, Ho, Wo, act, alpha, out_dtype, keep=keep_padded, out=out)
            if r is not None:
                return r
        assert x.segs is None, f"{w.name}: physically padded input needs the tensor-core path"
        seq = self._next()
        y = self._out_or_buf(out, (seq, "y"), (N, Ho, Wo, cout), out_dtype or self.act_dtype)
        cp = DgConvParams(kh, kw, stride, pt, pl, ACT[act], float(alpha))
        lin = DgConvParams(kh, kw, stride, pt, pl, 0, 0.0)
        umma = self._umma_ok(x.t, cin, cout, kh, kw, stride, H, W)
        umma_f = umma and self._umma_supported("fwd", x.t, y, lin)
        umma_d = umma and self._umma_supported("dgrad", y, x.t, lin)
        tx, ty = tensor(x.t), tensor(y)
        bias = _lib.ptr(b.data) if b is not None else None
        flops = 2.0 * N * Ho * Wo * kh * kw * cin * cout
        bn_part = bn_done = bn_applied = None
        fold = (isinstance(bn, tuple) and len(bn) == 4 and bn[0] == "fold")
        if fold:
            _, f_pset, f_name, f_eps = bn
            bn = False
            p_act = (post or {}).get("act")
            p_prelu, p_res = (post or {}).get("prelu"), (post or {}).get("residual")
            # a PReLU / skip-add behind the BatchNorm rides on the staged epilogue (dg_umma_conv2d_fwd_res_prelu) when the layer takes it
            extra = p_prelu is not None or p_res is not None
            extra_ok = False
            if extra and self.fold_bn_infer and self.fuse_res_epilogue and umma_f and act is None and y.dtype == torch.bfloat16 and \
                    (p_prelu is None or p_act is None) and (p_res is None or (p_res.t.dtype == torch.bfloat16 and p_res.segs is None and
                                                                               tuple(p_res.shape) == (N, Ho, Wo, cout))):
                keyb = ("bnblk", N, H, W, cin, Ho, Wo, cout, kh, kw, stride, pt, pl)
                blocks = self._cap.get(keyb)
                if blocks is None:
                    blocks = int(self.lib.dg_umma_conv2d_fwd_bn_blocks(self.ctx, C.byref(tx), C.byref(ty), C.byref(cp)))
                    self._cap[keyb] = blocks
                extra_ok = blocks > 0
            if (self.fold_bn_infer and umma_f and act is None and (not extra or extra_ok)
                    and p_act in (None, "relu", "lrelu") and y.dtype == torch.bfloat16):
                ent = self._fold(w, b, f_pset, f_name, f_eps, axis=3)
                if ent[3] is None:
                    ent[3] = torch.empty(kh * kw * cin * cout, dtype=torch.bfloat16, device=self.device)
                    check(self.lib.dg_umma_pack_weights_padded(self.ctx, ent[1].data_ptr(), ent[3].data_ptr(), kh, kw, cin, cout, cin, cout, 0, self.st))
                cpf = DgConvParams(kh, kw, stride, pt, pl, ACT[p_act], float((post or {}).get("alpha", 0.0)))
                if extra:
                    tr = tensor(p_res.t) if p_res is not None else None
                    self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_fwd_res_prelu(
                        self.ctx, C.byref(tx), ent[3].data_ptr(), ent[2].data_ptr(), C.byref(ty), C.byref(cpf),
                        C.byref(tr) if tr is not None else None, _lib.ptr(p_prelu.data) if p_prelu is not None else None, self.st)))
                else:
                    self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_fwd(
                        self.ctx, C.byref(tx), ent[3].data_ptr(), ent[2].data_ptr(), C.byref(ty), C.byref(cpf), None, self.st)))
                out = Var(y, self._deps([x] + ([p_res] if p_res is not None else []), w.group), seq)
                out.bn_folded = f_name
                out.bn_folded_post = extra
                return out          # inference only: no tape node
        if umma_f:
            pk = self._packed(w, 0)
            if bn and self.fuse_conv_bn_stats and act is None:
                key = ("bnblk", N, H, W, cin, Ho, Wo, cout, kh, kw, stride, pt, pl)
                blocks = self._cap.get(key)
                if blocks is None:
                    blocks = int(self.lib.dg_umma_conv2d_fwd_bn_blocks(self.ctx, C.byref(tx), C.byref(ty), C.byref(cp)))
                    self._cap[key] = blocks
                if blocks > 0:
                    bn_part = (self.buf((seq, "bn_part"), (blocks, 2, cout), torch.float32), blocks)
            fused_post = None
            if (post is not None and isinstance(bn, tuple) and self.fuse_conv_bn_act and act is None and y.dtype == torch.bfloat16 and
                    (post.get("residual") is None or post["residual"].t.dtype == torch.bfloat16)):
                keyp = ("bnact", N, H, W, cin, Ho, Wo, cout, kh, kw, stride, pt, pl)
                pblocks = self._cap.get(keyp)
                if pblocks is None:
                    pblocks = int(self.lib.dg_umma_conv2d_fwd_bn_act_blocks(self.ctx, C.byref(tx), C.byref(ty), C.byref(cp)))
                    self._cap[keyp] = pblocks
                if pblocks > 0:
                    fused_post = pblocks
            if fused_post:
                bp, bname, bmom, beps = bn
                part = self.buf((seq, "bn_part"), (fused_post, 2, cout), torch.float32)
                coef = [self.buf((seq, "bn_" + k), (cout,), torch.float32) for k in ("scale", "shift", "mean", "invstd")]
                y_act = self.buf((seq, "y_act"), (N, Ho, Wo, cout), torch.bfloat16)
                fz = _lib.DgBnFused(bp[bname + "/gamma"].data.data_ptr(), bp[bname + "/beta"].data.data_ptr(), float(beps), float(bmom),
                                    bp[bname + "/moving_mean"].data.data_ptr(), bp[bname + "/moving_variance"].data.data_ptr(),
                                    coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(), coef[3].data_ptr(), N * Ho * Wo)
                p_act, p_alpha, p_prelu, p_res = post.get("act"), float(post.get("alpha", 0.0)), post.get("prelu"), post.get("residual")
                a_code = ACT["prelu"] if p_prelu is not None else ACT[p_act]
                tres = tensor(p_res.t) if p_res is not None else None
                tya = tensor(y_act)
                self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_fwd_bn_act(
                    self.ctx, C.byref(tx), pk.data_ptr(), bias, C.byref(ty), C.byref(cp), part.data_ptr(), C.byref(fz), a_code, p_alpha,
                    _lib.ptr(p_prelu.data) if p_prelu is not None else None, C.byref(tres) if tres is not None else None, C.byref(tya), self.st)))
                bn_done = (bname, *coef)
                bn_part = None
                bn_applied = (y_act, p_act, p_alpha, p_prelu, p_res)
            elif bn_part and isinstance(bn, tuple) and self.fuse_conv_bn_finalize and cout <= 512:
                # (pset, name, momentum, eps): the conv's last CTA also turns the partial rows into the BatchNorm coefficients
                bp, bname, bmom, beps = bn
                coef = [self.buf((seq, "bn_" + k), (cout,), torch.float32) for k in ("scale", "shift", "mean", "invstd")]
                fz = _lib.DgBnFused(bp[bname + "/gamma"].data.data_ptr(), bp[bname + "/beta"].data.data_ptr(), float(beps), float(bmom),
                                    bp[bname + "/moving_mean"].data.data_ptr(), bp[bname + "/moving_variance"].data.data_ptr(),
                                    coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(), coef[3].data_ptr(), N * Ho * Wo)
                self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_fwd_bn(
                    self.ctx, C.byref(tx), pk.data_ptr(), bias, C.byref(ty), C.byref(cp), bn_part[0].data_ptr(), C.byref(fz), self.st)))
                bn_done = (bname, *coef)
                bn_part = None
            else:
                self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_fwd(
                    self.ctx, C.byref(tx), pk.data_ptr(), bias, C.byref(ty), C.byref(cp), _lib.ptr(bn_part[0]) if bn_part else None, self.st)))
        else:
            self._timed("simt_conv", flops, lambda: check(self.lib.dg_conv2d_fwd(
                self.ctx, C.byref(tx), w.data.data_ptr(), bias, C.byref(ty), C.byref(cp), self.st)))
        out = Var(y, self._deps([x], w.group), seq)
        out.bn_part = bn_part
        out.bn_done = bn_done
        out.bn_applied = bn_applied
        out.relu_out = act == "relu"
        mask_in = (x.relu_out and umma_d and x.t.dtype == torch.bfloat16 and y.dtype == torch.bfloat16 and
                   self._relu_mask_rows(tuple(y.shape), tuple(x.shape), lin) > 0)

        def bwd(gy: torch.Tensor, need_in, need_p, tag, fctx=None):
            dpre = gy
            if ACT[act] and not self._relu_bwd_folded(out, act):
                dpre = self.buf((seq, "dpre", tag), gy.shape, gy.dtype)
                tg, tyy, td = tensor(gy), tensor(y), tensor(dpre)
                check(self.lib.dg_act_bwd_from_output(self.ctx, C.byref(tg), C.byref(tyy), ACT[act], float(alpha), C.byref(td), self.st))
            tdp = tensor(dpre)
            if need_p:
                with self._side():
                    self._wgrad(x.t, dpre, w, b, lin, flops)
            dx = None
            if need_in[0] and mask_in:
                # the producer of x is a ReLU convolution that skips its own backward pass: dL/dx leaves this launch masked
                dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
                tdx = tensor(dx)
                pk = self._packed(w, 1)
                self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_dgrad_relu_mask(
                    self.ctx, C.byref(tdp), pk.data_ptr(), C.byref(tdx), C.byref(lin), C.byref(tx), self.st)))
                return [dx]
            if need_in[0] and fctx is not None and umma_d and stride == 1 and dpre.dtype == torch.bfloat16 and x.t.dtype == torch.bfloat16:
                # this launch produces the LAST contribution to dL/dx: the skip-connection gradient accumulated so far is added
                # in the epilogue, and when x is the output of a training-mode BatchNorm its backward sums are reduced there too
                res, src = fctx.get("residual"), (x.bn_src if self.fuse_dgrad_mode >= 2 else None)
                if res is not None and (res.dtype != torch.bfloat16 or tuple(res.shape) != tuple(x.shape)):
                    res = None
                keyq = ("dgf", N, H, W, cin, Ho, Wo, cout, kh, kw, pt, pl)
                rows = self._cap.get(keyq)
                if rows is None:
                    tdq = tensor(self.buf((seq, "dx", tag), x.shape, x.t.dtype))
                    rows = int(self.lib.dg_umma_conv2d_dgrad_fused_blocks(self.ctx, C.byref(tdp), C.byref(tdq), C.byref(lin)))
                    self._cap[keyq] = rows
                if rows > 0 and (res is not None or src is not None) and (res is not None or fctx.get("residual") is None):
                    dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
                    tdx = tensor(dx)
                    tres = tensor(res) if res is not None else None
                    st_ = None
                    if src is not None:
                        s_seq, s_x, s_scale, s_shift, s_mean, s_act, s_alpha = src
                        part = self.buf((seq, "bwd_part", tag), (rows, 2, cin), torch.float32)
                        tsx = tensor(s_x)
                        st_ = _lib.DgBnBwdStats(C.pointer(tsx), s_scale.data_ptr(), s_shift.data_ptr(), s_mean.data_ptr(), int(s_act),
                                                float(s_alpha), part.data_ptr())
                        self._bwd_part[(s_seq, tag)] = (part, rows)
                    pk = self._packed(w, 1)
                    self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_dgrad_fused(
                        self.ctx, C.byref(tdp), pk.data_ptr(), C.byref(tdx), C.byref(lin), C.byref(tres) if tres is not None else None,
                        C.byref(st_) if st_ is not None else None, self.st)))
                    fctx["fused"] = True
                    return [dx]
            if need_in[0]:
                dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
                tdx = tensor(dx)
                if umma_d and dpre.dtype != torch.bfloat16 and x.t.dtype == torch.bfloat16:
                    # fp32 output layer (VGG19 block5_conv4 feeding the fp32 feature loss, srgan.py:92): its gradient is rounded
                    # to bf16 once so that the input gradient runs on the tensor cores like every other layer
                    d16 = self.buf((seq, "dpre16", tag), dpre.shape, torch.bfloat16)
                    ts_, td_ = tensor(dpre), tensor(d16)
                    check(self.lib.dg_copy(self.ctx, C.byref(ts_), C.byref(td_), 0, self.st))
                    dpre, tdp = d16, td_
                if umma_d and dpre.dtype == torch.bfloat16:
                    pk = self._packed(w, 1)
                    self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_dgrad(
                        self.ctx, C.byref(tdp), pk.data_ptr(), None, C.byref(tdx), C.byref(lin), self.st)))
                else:
                    self._timed("simt_conv", flops, lambda: check(self.lib.dg_conv2d_dgrad(
                        self.ctx, C.byref(tdp), w.data.data_ptr(), None, C.byref(tdx), C.byref(lin), self.st)))
            return [dx]

        self._push([x], out, w.group, bwd, fusable=True, params=(w, b), masks=[x] if mask_in else (),
                   wkey=(tuple(x.shape), tuple(y.shape), kh, kw, stride, pt, pl, b is None))
        return out

    def _wgrad(self, x: torch.Tensor, dy: torch.Tensor, w: Param, b: Param | None, lin: DgConvParams, flops: float = 0.0):
        """dW (+ dbias) of the forward conv `lin` with input x and output-gradient dy, into the grad arena."""
        tx, tdy = tensor(x), tensor(dy)
        acc = self._acc_flag(w)
        if b is not None:
            accb = self._acc_flag(b)
            assert accb == acc
        dbias = b.grad.data_ptr() if b is not None else None
        kh, kw, cin, cout = lin.kh, lin.kw, x.shape[3], dy.shape[3]
        Nn, Ho, Wo = dy.shape[0], dy.shape[1], dy.shape[2]
        if (self.use_umma_wgrad and self.small_map_gemm and x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and kh * kw > 1 and
                Ho * Wo <= 64 and (Nn * Ho * Wo) % 8 == 0 and cin % 16 == 0 and cout % 16 == 0):
            # maps of at most 8x8 pixels (pix2pix.py:147-166): the halo-tile kernel would spend most of its 128-pixel tiles on
            # padding; gather the taps once (dg_im2col) and run the weight gradient as a 1x1 layer over all N*Ho*Wo pixels
            P_ = Nn * Ho * Wo
            col = self.buf(("im2col", self._in_side, P_, kh * kw * cin), (1, P_ // 8, 8, kh * kw * cin), torch.bfloat16)
            check(self.lib.dg_im2col(self.ctx, C.byref(tx), C.byref(lin), Ho, Wo, col.data_ptr(), self.st))
            tcol = tensor(col)
            tdy2 = _lib.DgTensor(dy.data_ptr(), _lib.DG_BF16, 1, P_ // 8, 8, cout, cout, 0)
            one = DgConvParams(1, 1, 1, 0, 0, 0, 0.0)
            nbytes = self.lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tcol), C.byref(tdy2), C.byref(one))
            if nbytes > 0:
                ws = self.workspace(nbytes)
                self._timed("umma_wgrad", flops, lambda: check(self.lib.dg_umma_conv2d_wgrad(
                    self.ctx, C.byref(tcol), C.byref(tdy2), w.grad.data_ptr(), dbias, C.byref(one), acc, ws.data_ptr(), nbytes, self.st)))
                return
        if (self.use_umma_wgrad and x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and
                self._umma_ok(x, cin, cout, kh, kw, lin.stride, x.shape[1], x.shape[2])):
            nbytes = self.lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(tdy), C.byref(lin))
        else:
            nbytes = 0
        if nbytes > 0 and self.wgrad_batch > 1 and self._in_backward and self._wgrad_batchable(tx, tdy, lin):
            self._wgrad_enqueue(x, dy, w, b, acc, flops, lin)
            return
        if nbytes > 0:     # 0 = the tensor-core wgrad has no tile configuration for this layer
            ws = self.workspace(nbytes)
            self._timed("umma_wgrad", flops, lambda: check(self.lib.dg_umma_conv2d_wgrad(
                self.ctx, C.byref(tx), C.byref(tdy), w.grad.data_ptr(), dbias, C.byref(lin), acc, ws.data_ptr(), nbytes, self.st)))
        else:
            nbytes = self.lib.dg_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(tdy), C.byref(lin))
            ws = self.workspace(nbytes)
            self._timed("simt_wgrad", flops, lambda: check(self.lib.dg_conv2d_wgrad(
                self.ctx, C.byref(tx), C.byref(tdy), w.grad.data_ptr(), dbias, C.byref(lin), acc, ws.data_ptr(), nbytes, self.st)))

    # ---- weight gradients of identical layers, several per launch
    def _wgrad_batchable(self, tx, tdy, lin) -> bool:
        key = ("wgb", tx.n, tx.h, tx.w, tx.c, tdy.h, tdy.w, tdy.c, lin.kh, lin.kw, lin.stride, lin.pad_t, lin.pad_l)
        ok = self._cap.get(key)
        if ok is None:
            # every group size up to wgrad_batch must have a single-launch tile configuration (stride-2 and > 128-channel
            # discriminator layers included; DG_WGRAD_BATCH_WIDE=0 restores the stride-1 / <= 128-channel rule for A/B runs)
            ok = all(bool(self.lib.dg_umma_conv2d_wgrad_batch_supported(self.ctx, k, C.byref(tx), C.byref(tdy), C.byref(lin)))
                     for k in range(2, self.wgrad_batch + 1))
            if os.environ.get("DG_WGRAD_BATCH_WIDE", "1") == "0":
                ok = ok and lin.stride == 1 and tx.c <= 128
            self._cap[key] = ok
        return ok

    def _wgrad_enqueue(self, x, dy, w, b, acc, flops, lin):
        key = (tuple(x.shape), tuple(dy.shape), lin.kh, lin.kw, lin.stride, lin.pad_t, lin.pad_l, b is None)
        q = self._wq.get(key)
        if q:
            merged = all(e[2] is q[0][2] for e in q)
            dup = any(e[2] is w for e in q)
            if (merged and len(q) > 1 and not dup) or (dup and not merged):
                self._wgrad_flush(key)          # keeps "all the same gradient" and "all distinct gradients" groups apart, and the write order per parameter
                q = None
        if not q:
            q = self._wq[key] = []
        q.append((x, dy, w, b, acc, flops, lin))
        self._wq_names.add(w.name)
        if b is not None:
            self._wq_names.add(b.name)
        # a group is launched when it is full or when no partner is left to come: backward() counted the convolutions of every
        # geometry on the tape beforehand, so a layer without partners is never held back to the end of the pass
        seen = self._wq_seen[key] = self._wq_seen.get(key, 0) + 1
        left = self._wq_expect.get(key, 0) - seen
        if len(q) >= self.wgrad_batch or left <= 0:
            self._wgrad_flush(key)

    def _wgrad_flush(self, key=None):
        """Launches the queued weight gradients of one geometry (or of all): one dg_umma_conv2d_wgrad_batch per group."""
        keys = [key] if key is not None else list(self._wq)
        for k in keys:
            q = self._wq.pop(k, None)
            if not q:
                continue
            with self._side():
                n = len(q)
                lin = q[0][6]
                tens = [(tensor(e[0]), tensor(e[1])) for e in q]
                xs = (C.POINTER(_lib.DgTensor) * n)(*[C.pointer(t[0]) for t in tens])
                dys = (C.POINTER(_lib.DgTensor) * n)(*[C.pointer(t[1]) for t in tens])
                dws = (C.c_void_p * n)(*[e[2].grad.data_ptr() for e in q])
                dbs = (C.c_void_p * n)(*[(e[3].grad.data_ptr() if e[3] is not None else None) for e in q])
                accs = (C.c_int * n)(*[int(e[4]) for e in q])
                nbytes = self.lib.dg_umma_conv2d_wgrad_batch_workspace_bytes(n, C.byref(tens[0][0]), C.byref(tens[0][1]), C.byref(lin))
                ws = self.workspace(nbytes)
                self._timed("umma_wgrad", sum(e[5] for e in q), lambda: check(self.lib.dg_umma_conv2d_wgrad_batch(
                    self.ctx, n, xs, dys, dws, dbs if q[0][3] is not None else None, C.byref(lin), accs, ws.data_ptr(), nbytes, self.st)))
            for e in q:
                for prm in (e[2], e[3]):
                    if prm is not None:
                        self._wq_names.discard(prm.name)
                        if prm.name in self._complete_wait:
                            self._complete_wait.discard(prm.name)
                            self.complete.add(prm.name)

    def conv2d_transpose(self, x: Var, w: Param, b: Param | None = None, *, stride=2, act=None, alpha=0.0,
                         out_dtype=None) -> Var:
        """keras Conv2DTranspose(padding='same'), kernel [kh,kw,Cout,Cin]: the input-gradient of the SAME conv
        f: [N,sH,sW,Cout] -> [N,H,W,Cin]."""
        N, H, W, Cin = x.shape
        kh, kw, cout, cin = w.shape
        assert cin == Cin
        Ho, Wo = H * stride, W * stride
        pt, _ = same_pads(Ho, kh, stride)
        pl, _ = same_pads(Wo, kw, stride)
        if (self.use_umma and self.pad_rgb and cout % 16 != 0 and cin % 16 == 0 and x.t.dtype == torch.bfloat16 and kh * kw <= 16
                and stride == 2):
            r = self._conv2d_transpose_padded(x, w, b, stride, pt, pl, Ho, Wo, act, alpha, out_dtype)
            if r is not None:
                return r
        seq = self._next()
        y = self.buf((seq, "y"), (N, Ho, Wo, cout), out_dtype or self.act_dtype)
        cp = DgConvParams(kh, kw, stride, pt, pl, ACT[act], float(alpha))
        lin = DgConvParams(kh, kw, stride, pt, pl, 0, 0.0)
        # as a forward conv f, the kernel is HWIO with I = cout (of the transposed conv), O = cin
        umma = self._umma_ok(x.t, cin, cout, kh, kw, stride, Ho, Wo)
        umma_up = umma and self._umma_supported("dgrad", x.t, y, lin)     # y = dgrad_f(x)
        umma_dn = umma and self._umma_supported("fwd", y, x.t, lin)       # dx = f(dy)
        tx, ty = tensor(x.t), tensor(y)
        bias = _lib.ptr(b.data) if b is not None else None
        if umma_up:
            check(self.lib.dg_umma_conv2d_dgrad(self.ctx, C.byref(tx), self._packed(w, 1).data_ptr(), bias, C.byref(ty), C.byref(cp), self.st))
        else:
            check(self.lib.dg_conv2d_dgrad(self.ctx, C.byref(tx), w.data.data_ptr(), bias, C.byref(ty), C.byref(cp), self.st))
        out = Var(y, self._deps([x], w.group), seq)

        def bwd(gy, need_in, need_p, tag):
            dpre = gy
            if ACT[act]:
                dpre = self.buf((seq, "dpre", tag), gy.shape, gy.dtype)
                tg, tyy, td = tensor(gy), tensor(y), tensor(dpre)
                check(self.lib.dg_act_bwd_from_output(self.ctx, C.byref(tg), C.byref(tyy), ACT[act], float(alpha), C.byref(td), self.st))
            if need_p:
                # dWt = wgrad of f with "input" = dY (large image) and "output grad" = x (small image)
                with self._side():
                    self._wgrad(dpre, x.t, w, None, lin)
                if b is not None:
                    tdp2 = tensor(dpre)
                    nb2 = self.lib.dg_bn_workspace_bytes(C.byref(tdp2))
                    ws2 = self.workspace(nb2)
                    check(self.lib.dg_bias_grad(self.ctx, C.byref(tdp2), b.grad.data_ptr(), self._acc_flag(b), ws2.data_ptr(), nb2, self.st))
            dx = None
            if need_in[0]:
                dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
                tdp, tdx = tensor(dpre), tensor(dx)
                if umma_dn and dpre.dtype == torch.bfloat16:
                    check(self.lib.dg_umma_conv2d_fwd(self.ctx, C.byref(tdp), self._packed(w, 0).data_ptr(), None, C.byref(tdx), C.byref(lin), None, self.st))
                else:
                    check(self.lib.dg_conv2d_fwd(self.ctx, C.byref(tdp), w.data.data_ptr(), None, C.byref(tdx), C.byref(lin), self.st))
            return [dx]

        self._push([x], out, w.group, bwd, params=(w, b))
        return out

    def _conv2d_transpose_padded(self, x: Var, w: Param, b: Param | None, stride, pt, pl, Ho, Wo, act, alpha, out_dtype):
        """Conv2DTranspose whose OUTPUT channel count is not a multiple of 16 (pix2pix.py:169-173, 128 -> 3 + tanh) on the
        tensor cores: the output side is zero-padded to 16 channels exactly as in _conv2d_padded (the transposed conv is the
        input-gradient of a forward conv f: [N,Ho,Wo,cout] -> [N,H,W,cin], so the padded side is f's INPUT)."""
        N, H, W, cin = x.shape
        kh, kw, cout, _ = w.shape
        cout_p = -(-cout // 16) * 16
        ydt = out_dtype or self.act_dtype
        lin = DgConvParams(kh, kw, stride, pt, pl, 0, 0.0)
        cp = DgConvParams(kh, kw, stride, pt, pl, ACT[act], float(alpha))
        dummy = x.t.data_ptr()
        d_big = _lib.DgTensor(dummy, _lib.DG_BF16, N, Ho, Wo, cout_p, cout_p, 0)
        d_small = _lib.DgTensor(dummy, _lib.DG_BF16, N, H, W, cin, cin, 0)
        key = ("padded_T", N, H, W, cin, Ho, Wo, cout_p, kh, kw, stride, pt, pl)
        ok = self._cap.get(key)
        if ok is None:
            ok = (bool(self.lib.dg_umma_conv2d_dgrad_supported(self.ctx, C.byref(d_small), C.byref(d_big), C.byref(lin))) and
                  bool(self.lib.dg_umma_conv2d_fwd_supported(self.ctx, C.byref(d_big), C.byref(d_small), C.byref(lin))) and
                  self.lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(d_big), C.byref(d_small), C.byref(lin)) > 0)
            self._cap[key] = ok
        if not ok:
            return None
        seq = self._next()
        w.pack_pad = (cout_p, cin)          # HWIO of f: I = cout of the transposed conv (padded), O = cin
        y = self.buf((seq, "y"), (N, Ho, Wo, cout), ydt)
        yp = self.buf((seq, "ypad"), (N, Ho, Wo, cout_p), ydt)
        bias = None
        if b is not None:
            bp = self._zeros((seq, "bias_p"), (cout_p,), torch.float32)
            self._copy_vec(b.data, bp, cout)
            bias = bp.data_ptr()
        tx, typ = tensor(x.t), tensor(yp)
        flops = 2.0 * N * H * W * kh * kw * cin * cout        # ALGORITHMIC: the zero-padded channels are not counted
        pk1 = self._packed(w, 1)
        self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_dgrad(
            self.ctx, C.byref(tx), pk1.data_ptr(), bias, C.byref(typ), C.byref(cp), self.st)))
        tv, ty = tensor(yp, c=cout), tensor(y)
        check(self.lib.dg_copy(self.ctx, C.byref(tv), C.byref(ty), 0, self.st))
        out = Var(y, self._deps([x], w.group), seq)

        def bwd(gy, need_in, need_p, tag):
            dpre = gy
            if ACT[act]:
                dpre = self.buf((seq, "dpre", tag), gy.shape, gy.dtype)
                tg, tyy, td = tensor(gy), tensor(y), tensor(dpre)
                check(self.lib.dg_act_bwd_from_output(self.ctx, C.byref(tg), C.byref(tyy), ACT[act], float(alpha), C.byref(td), self.st))
            dpp = self.buf((seq, "dpad", tag), (N, Ho, Wo, cout_p), torch.bfloat16)
            ts, tdp = tensor(dpre), tensor(dpp)
            check(self.lib.dg_pad_channels(self.ctx, C.byref(ts), C.byref(tdp), self.st))
            if need_p:
                acc = self._acc_flag(w)
                with self._side():
                    nbytes = self.lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tdp), C.byref(tx), C.byref(lin))
                    ws = self.workspace(nbytes)
                    dwp = self.buf((seq, "dw_pad", tag), (kh * kw * cout_p * cin,), torch.float32)
                    self._timed("umma_wgrad", flops, lambda: check(self.lib.dg_umma_conv2d_wgrad(
                        self.ctx, C.byref(tdp), C.byref(tx), dwp.data_ptr(), None, C.byref(lin), 0, ws.data_ptr(), nbytes, self.st)))
                    check(self.lib.dg_unpad_weight_grad(self.ctx, dwp.data_ptr(), None, w.grad.data_ptr(), None, kh, kw, cout, cin,
                                                        cout_p, cin, acc, self.st))
                if b is not None:
                    tdp2 = tensor(dpre)
                    nb2 = self.lib.dg_bn_workspace_bytes(C.byref(tdp2))
                    ws2 = self.workspace(nb2)
                    check(self.lib.dg_bias_grad(self.ctx, C.byref(tdp2), b.grad.data_ptr(), self._acc_flag(b), ws2.data_ptr(), nb2, self.st))
            dx = None
            if need_in[0]:
                dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
                tdx = tensor(dx)
                pk0 = self._packed(w, 0)
                self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_fwd(
                    self.ctx, C.byref(tdp), pk0.data_ptr(), None, C.byref(tdx), C.byref(lin), None, self.st)))
            return [dx]

        self._push([x], out, w.group, bwd, params=(w, b))
        return out

    def fsrgan_block_infer(self, x: Var, pset, prefix: str, eps: float = 1e-3) -> Var | None:
        """Inverted-residual block `prefix` (e.g. "g/b3") of the Fast-SRGAN generator with training=False as ONE launch
        (fsrgan.py:112-176; dg_fsrgan_block_infer): expand -> BN -> ReLU -> depthwise -> BN -> ReLU -> project -> BN -> + x, the
        BatchNorms folded into kernels and biases (_fold).  Returns None when the tensors do not qualify (the caller then issues
        the layer calls)."""
        if not (self.fuse_fsrgan_block and self.fold_bn_infer and self.use_umma and x.t.dtype == torch.bfloat16 and x.segs is None):
            return None
        we, be = pset[prefix + "/expand/kernel"], pset[prefix + "/expand/bias"]
        wd, bd = pset[prefix + "/dw/kernel"], pset[prefix + "/dw/bias"]
        wp, bp = pset[prefix + "/project/kernel"], pset[prefix + "/project/bias"]
        if tuple(we.shape) != (1, 1, 32, 192) or tuple(wd.shape) != (3, 3, 192, 1) or tuple(wp.shape) != (1, 1, 192, 32) or x.shape[3] != 32:
            return None
        seq = self._next()
        y = self.buf((seq, "y"), x.shape, torch.bfloat16)
        tx, ty = tensor(x.t), tensor(y)
        if not self.lib.dg_fsrgan_block_infer_supported(self.ctx, C.byref(tx), C.byref(ty)):
            return None
        fe = self._fold(we, be, pset, prefix + "/expand_bn", eps, axis=3)
        fd = self._fold(wd, bd, pset, prefix + "/dw_bn", eps, axis=2)
        fp = self._fold(wp, bp, pset, prefix + "/project_bn", eps, axis=3)
        ver = (fe[0], fd[0], fp[0])
        ent = self._folded.get(("fsrgan_block", prefix))
        if ent is None or ent[0] != ver:
            # output-channel-major copies of the two 1x1 kernels (the K-major B operands of the two products: bf16 / fp16); addresses are
            # kept across parameter updates (they may be baked into a captured graph)
            w1 = fe[1].view(32, 192).t().contiguous().to(torch.bfloat16)
            w2 = fp[1].view(192, 32).t().contiguous().to(torch.float16)
            if ent is None:
                ent = [ver, w1, w2]
            else:
                ent[1].copy_(w1); ent[2].copy_(w2); ent[0] = ver
            self._folded[("fsrgan_block", prefix)] = ent
        flops = float(x.shape[0] * x.shape[1] * x.shape[2]) * (2 * 32 * 192 + 2 * 9 * 192 + 2 * 192 * 32)   # expand + depthwise + project
        self._timed("fsrgan_block", flops, lambda: check(self.lib.dg_fsrgan_block_infer(
            self.ctx, C.byref(tx), ent[1].data_ptr(), fe[2].data_ptr(), fd[1].data_ptr(), fd[2].data_ptr(), ent[2].data_ptr(), fp[2].data_ptr(),
            C.byref(ty), self.st)))
        return Var(y, self._deps([x], we.group), seq)     # inference only: no tape node

    def conv3x3_image_infer(self, x: Var, w: Param, b: Param | None, act=None, alpha=0.0) -> Var | None:
        """The generator's last layer at inference -- Conv2D(3, 3x3, SAME) + tanh in float32 on the up-scaled image
        (fsrgan.py:216-217) -- in tap-sum form (dg_conv3x3_tapsum_fwd: one tensor-core product against all nine taps, then nine
        shifted adds).  When a frame sink is installed (FrameRunner: infer_video.py:150-159) the uint8 frame is written directly
        (dg_conv3x3_tapsum_frame) and the float image never exists.  Returns None when the layer does not qualify (the caller
        then uses conv2d)."""
        if not (self.tapsum_infer and self.use_umma and x.segs is None and x.t.dtype == torch.bfloat16):
            return None
        kh, kw, cin, cout = w.shape
        N, H, W, C_ = x.shape
        if (kh, kw, cin) != (3, 3, 32) or C_ != 32 or not 1 <= cout <= 3:
            return None
        tx = tensor(x.t)
        if not self.lib.dg_conv3x3_tapsum_supported(self.ctx, C.byref(tx), cout):
            return None
        seq = self._next()
        bias = _lib.ptr(b.data) if b is not None else None
        flops = 2.0 * N * H * W * 9 * cin * cout
        sink = self.frame_sink
        if sink is not None and not sink["done"] and cout == 3 and N == 1 and sink["h"] <= H and sink["w"] <= W:
            out = sink["out"]
            self._timed("tapsum_conv", flops * (sink["h"] * sink["w"]) / (H * W), lambda: check(self.lib.dg_conv3x3_tapsum_frame(
                self.ctx, C.byref(tx), w.data.data_ptr(), bias, ACT[act], float(alpha), float(sink["scale"]), float(sink["offset"]),
                int(sink["clip"]), int(sink["flip"]), out.data_ptr(), sink["h"], sink["w"], self.st)))
            sink["done"] = True
            return Var(out, self._deps([x], w.group), seq)      # inference only: no tape node; the uint8 frame IS the result
        y = self.buf((seq, "y"), (N, H, W, cout), torch.float32)
        ty = tensor(y)
        self._timed("tapsum_conv", flops, lambda: check(self.lib.dg_conv3x3_tapsum_fwd(
            self.ctx, C.byref(tx), w.data.data_ptr(), bias, ACT[act], float(alpha), C.byref(ty), self.st)))
        return Var(y, self._deps([x], w.group), seq)            # inference only: no tape node

    def dwconv3x3(self, x: Var, w: Param, b: Param | None, bn=False, post: dict | None = None) -> Var:
        """keras DepthwiseConv2D(3, padding='same').  `bn` / `post` as for conv2d: at inference the BatchNorm (+ ReLU) that follows
        (fsrgan.py:149-155) is folded into the kernel, the bias and the store."""
        seq = self._next()
        y = self.buf((seq, "y"), x.shape, x.t.dtype)
        tx, ty = tensor(x.t), tensor(y)
        if (isinstance(bn, tuple) and len(bn) == 4 and bn[0] == "fold" and self.fold_bn_infer and (post or {}).get("act") in (None, "relu")
                and (post or {}).get("prelu") is None and (post or {}).get("residual") is None):
            _, f_pset, f_name, f_eps = bn
            ent = self._fold(w, b, f_pset, f_name, f_eps, axis=2)          # DepthwiseConv2D kernel [3,3,C,1]
            check(self.lib.dg_dwconv3x3_fwd_act(self.ctx, C.byref(tx), ent[1].data_ptr(), ent[2].data_ptr(), ACT[(post or {}).get("act")],
                                                C.byref(ty), self.st))
            out = Var(y, self._deps([x], w.group), seq)
            out.bn_folded = f_name
            return out
        check(self.lib.dg_dwconv3x3_fwd(self.ctx, C.byref(tx), w.data.data_ptr(), _lib.ptr(b.data) if b is not None else None, C.byref(ty), self.st))
        out = Var(y, self._deps([x], w.group), seq)

        def bwd(gy, need_in, need_p, tag):
            tg = tensor(gy)
            if need_p:
                acc = self._acc_flag(w)
                if b is not None:
                    self._acc_flag(b)
                nbytes = self.lib.dg_dwconv3x3_wgrad_workspace_bytes(C.byref(tx))
                ws = self.workspace(nbytes)
                check(self.lib.dg_dwconv3x3_wgrad(self.ctx, C.byref(tx), C.byref(tg), w.grad.data_ptr(),
                                                  b.grad.data_ptr() if b is not None else None, acc, ws.data_ptr(), nbytes, self.st))
            dx = None
            if need_in[0]:
                dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
                tdx = tensor(dx)
                check(self.lib.dg_dwconv3x3_dgrad(self.ctx, C.byref(tg), w.data.data_ptr(), C.byref(tdx), self.st))
            return [dx]

        self._push([x], out, w.group, bwd, params=(w, b))
        return out

    # ------------------------------------------------------------------ batch norm (+act, +residual, +dropout)
    def bn_act(self, x: Var, pset, name: str, *, training: bool, momentum=0.99, eps=1e-3, act=None, alpha=0.0,
               prelu: Param | None = None, residual: Var | None = None, dropout_seed=None, dropout_offset=0,
               step_counter: torch.Tensor | None = None, out: torch.Tensor | None = None) -> Var:
        if not training and getattr(x, "bn_folded", None) == name:
            assert out is None, f"{name}: the folded inference form has no out= (the producing convolution owns the buffer)"
            # inference: the producing convolution already applied this BatchNorm and its activation (folded kernel / bias / epilogue),
            # and -- bn_folded_post -- the PReLU / skip-add that follow it (dg_umma_conv2d_fwd_res_prelu)
            assert dropout_seed is None and (x.bn_folded_post or (prelu is None and residual is None))
            return x
        gamma, beta = pset[name + "/gamma"], pset[name + "/beta"]
        mm, mv = pset[name + "/moving_mean"], pset[name + "/moving_variance"]
        Cc = x.shape[3]
        seq = self._next()
        bn_done = getattr(x, "bn_done", None) if training else None
        if bn_done is not None:
            assert bn_done[0] == name, f"conv finalised the statistics of {bn_done[0]}, consumed by {name}"
            _, scale, shift, mean, invstd = bn_done
        else:
            scale = self.buf((seq, "scale"), (Cc,), torch.float32)
            shift = self.buf((seq, "shift"), (Cc,), torch.float32)
            mean = self.buf((seq, "mean"), (Cc,), torch.float32)
            invstd = self.buf((seq, "invstd"), (Cc,), torch.float32)
        tx = tensor(x.t)
        nbytes = self.lib.dg_bn_workspace_bytes(C.byref(tx))
        ws = self.workspace(nbytes)
        drop = 1 if (dropout_seed is not None and training) else 0
        applied = getattr(x, "bn_applied", None) if (training and bn_done is not None and not drop) else None
        if applied is not None:
            # the producing convolution already normalised, activated and added the skip connection (one cooperative launch)
            y, f_act, f_alpha, f_prelu, f_res = applied
            assert (f_act or None) == (act or None) and f_prelu is prelu and f_res is residual and \
                (act != "lrelu" or abs(f_alpha - float(alpha)) < 1e-12), f"{name}: fused conv epilogue does not match this bn_act call"
            assert out is None
        else:
            y = self._out_or_buf(out, (seq, "y"), x.shape, x.t.dtype)
        ty = tensor(y)
        tres = tensor(residual.t) if residual is not None else None
        a_code = ACT["prelu"] if prelu is not None else ACT[act]
        fused = applied is not None
        bn_part = getattr(x, "bn_part", None)
        if bn_done is not None:
            pass    # scale/shift/mean/invstd and the moving statistics were written by the producing convolution
        elif training and bn_part is not None and not drop and self.fuse_bn_finalize_apply and (rc0 := self.lib.dg_bn_act_fwd_from_partials(
                self.ctx, C.byref(tx), bn_part[0].data_ptr(), bn_part[1], gamma.data.data_ptr(), beta.data.data_ptr(), float(eps), float(momentum),
                mm.data.data_ptr(), mv.data.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), a_code, float(alpha),
                _lib.ptr(prelu.data) if prelu is not None else None, C.byref(tres) if tres is not None else None, C.byref(ty), self.st)) != 2:
            # finalize + apply in one launch: every block of the apply pass sums the per-CTA rows of the producing convolution itself
            check(rc0)
            fused = True
        elif training and bn_part is not None:
            # the producing convolution already reduced the batch statistics to per-CTA partials
            check(self.lib.dg_bn_finalize(self.ctx, bn_part[0].data_ptr(), bn_part[1], x.shape[0] * x.shape[1] * x.shape[2], Cc,
                                          gamma.data.data_ptr(), beta.data.data_ptr(), float(eps), float(momentum), mm.data.data_ptr(),
                                          mv.data.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), self.st))
        elif training and not self.fuse_bn_fwd:
            check(self.lib.dg_bn_stats(self.ctx, C.byref(tx), gamma.data.data_ptr(), beta.data.data_ptr(), float(eps), float(momentum),
                                       mm.data.data_ptr(), mv.data.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                       invstd.data_ptr(), ws.data_ptr(), nbytes, self.st))
        elif training:
            # statistics + apply as ONE launch when the tensors qualify (rc 2: they do not, issue the two calls)
            rc = self.lib.dg_bn_train_fwd(self.ctx, C.byref(tx), gamma.data.data_ptr(), beta.data.data_ptr(), float(eps), float(momentum),
                                          mm.data.data_ptr(), mv.data.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                          invstd.data_ptr(), a_code, float(alpha), _lib.ptr(prelu.data) if prelu is not None else None,
                                          C.byref(tres) if tres is not None else None, drop, int(dropout_seed or 0), int(dropout_offset),
                                          _lib.ptr(step_counter), C.byref(ty), ws.data_ptr(), nbytes, self.st)
            if rc == 2:
                check(self.lib.dg_bn_stats(self.ctx, C.byref(tx), gamma.data.data_ptr(), beta.data.data_ptr(), float(eps), float(momentum),
                                           mm.data.data_ptr(), mv.data.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                           invstd.data_ptr(), ws.data_ptr(), nbytes, self.st))
            else:
                check(rc)
                fused = True
        else:
            check(self.lib.dg_bn_infer_affine(self.ctx, Cc, gamma.data.data_ptr(), beta.data.data_ptr(), mm.data.data_ptr(),
                                              mv.data.data_ptr(), float(eps), scale.data_ptr(), shift.data_ptr(), self.st))
        if not fused:
            check(self.lib.dg_bn_act_fwd(self.ctx, C.byref(tx), scale.data_ptr(), shift.data_ptr(), a_code, float(alpha),
                                         _lib.ptr(prelu.data) if prelu is not None else None,
                                         C.byref(tres) if tres is not None else None, drop, int(dropout_seed or 0), int(dropout_offset),
                                         _lib.ptr(step_counter), C.byref(ty), self.st))
        inputs = [x] + ([residual] if residual is not None else [])
        out = Var(y, self._deps(inputs, gamma.group), seq)
        if (training and not drop and prelu is None and a_code in (0, 1, 2) and self.fuse_dgrad_bn_bwd and x.t.dtype == torch.bfloat16
                and y.dtype == torch.bfloat16):
            out.bn_src = (seq, x.t, scale, shift, mean, a_code, float(alpha))

        def bwd(gy, need_in, need_p, tag):
            assert training, "backward through inference-mode BN is not part of the hot path"
            dx = self.buf((seq, "dx", tag), x.shape, gy.dtype)
            tg, tdx = tensor(gy), tensor(dx)
            if need_p:
                acc = self._acc_flag(gamma)
                self._acc_flag(beta)
                if prelu is not None:
                    self._acc_flag(prelu)
                dg, db = gamma.grad.data_ptr(), beta.grad.data_ptr()
                da = prelu.grad.data_ptr() if prelu is not None else None
            else:
                acc, dg, db, da = 0, None, None, None
            pre = self._bwd_part.pop((seq, tag), None)
            if pre is not None:
                # the convolution that produced gy already reduced sum g' and sum g'(x - mean) per CTA: one pass for dx
                check(self.lib.dg_bn_bwd_dx_from_partials(self.ctx, C.byref(tg), C.byref(tx), scale.data_ptr(), shift.data_ptr(),
                                                          gamma.data.data_ptr(), mean.data_ptr(), invstd.data_ptr(), a_code, float(alpha),
                                                          pre[0].data_ptr(), int(pre[1]), C.byref(tdx), dg, db, acc, self.st))
                res = [dx if need_in[0] else None]
                if residual is not None:
                    res.append(gy if need_in[1] else None)
                return res
            ws = self.workspace(nbytes)     # the workspace of the stream this runs on (not the forward pass's: side-stream branches)
            check(self.lib.dg_bn_act_bwd(self.ctx, C.byref(tg), C.byref(tx), scale.data_ptr(), shift.data_ptr(), gamma.data.data_ptr(),
                                         mean.data_ptr(), invstd.data_ptr(), a_code, float(alpha),
                                         _lib.ptr(prelu.data) if prelu is not None else None, drop, int(dropout_seed or 0),
                                         int(dropout_offset), _lib.ptr(step_counter), C.byref(tdx), dg, db, da, acc, ws.data_ptr(), nbytes,
                                         self.st))
            res = [dx if need_in[0] else None]
            if residual is not None:
                res.append(gy if need_in[1] else None)
            return res

        self._push(inputs, out, gamma.group, bwd, params=(gamma, beta, prelu))
        return out

    # ------------------------------------------------------------------ structural ops
    def conv2d_d2s_prelu(self, x: Var, w: Param, b: Param | None, prelu: Param | None, training: bool) -> Var:
        """The up-sampling block Conv2D -> depth_to_space(2) -> PReLU (srgan.py:144-146, fsrgan.py:180-186).  Inference on the tensor-core
        path: ONE launch, the convolution's epilogue stores in depth_to_space order with the PReLU applied and the pre-activation
        tensor is never written (dg_umma_conv2d_fwd_d2s_prelu).  Training (or no tensor cores): the two ops."""
        N, H, W, cin = x.shape
        kh, kw, _, cout = w.shape
        if (training or not self.fuse_d2s_infer or x.segs is not None or not self._umma_ok(x.t, cin, cout, kh, kw, 1, H, W) or cout % 128 != 0 or cout > 768):
            return self.d2s_prelu(self.conv2d(x, w, b), prelu)
        pt, pl, Ho, Wo = self._conv_geom(H, W, kh, kw, 1, "same")
        lin = DgConvParams(kh, kw, 1, pt, pl, 0, 0.0)
        d_y = _lib.DgTensor(x.t.data_ptr(), _lib.DG_BF16, N, Ho, Wo, cout, cout, 0)
        key = ("d2sfwd", N, H, W, cin, cout, kh, kw)
        ok = self._cap.get(key)
        if ok is None:
            tx0 = tensor(x.t)
            ok = bool(self.lib.dg_umma_conv2d_fwd_supported(self.ctx, C.byref(tx0), C.byref(d_y), C.byref(lin)))
            self._cap[key] = ok
        if not ok:
            return self.d2s_prelu(self.conv2d(x, w, b), prelu)
        seq = self._next()
        y = self.buf((seq, "y"), (N, 2 * Ho, 2 * Wo, cout // 4), torch.bfloat16)
        tx, ty = tensor(x.t), tensor(y)
        pk = self._packed(w, 0)
        flops = 2.0 * N * Ho * Wo * kh * kw * cin * cout
        self._timed("umma_conv", flops, lambda: check(self.lib.dg_umma_conv2d_fwd_d2s_prelu(
            self.ctx, C.byref(tx), pk.data_ptr(), _lib.ptr(b.data) if b is not None else None, C.byref(ty), C.byref(lin),
            _lib.ptr(prelu.data) if prelu is not None else None, self.st)))
        return Var(y, self._deps([x], w.group), seq)          # inference only: no tape node

    def d2s_prelu(self, u: Var, prelu: Param | None) -> Var:
        """tf.nn.depth_to_space(u, 2) then PReLU(shared_axes=[1,2])."""
        N, H, W, C4 = u.shape
        seq = self._next()
        y = self.buf((seq, "y"), (N, 2 * H, 2 * W, C4 // 4), u.t.dtype)
        tu, ty = tensor(u.t), tensor(y)
        check(self.lib.dg_d2s_prelu_fwd(self.ctx, C.byref(tu), _lib.ptr(prelu.data) if prelu is not None else None, C.byref(ty), self.st))
        group = prelu.group if prelu is not None else None
        out = Var(y, self._deps([u], group), seq)

        def bwd(gy, need_in, need_p, tag):
            du = self.buf((seq, "du", tag), u.shape, gy.dtype)
            tg, tdu = tensor(gy), tensor(du)
            da, acc = None, 0
            if need_p and prelu is not None:
                acc = self._acc_flag(prelu)
                da = prelu.grad.data_ptr()
            nbytes = self.lib.dg_bn_workspace_bytes(C.byref(tg))
            ws = self.workspace(nbytes)
            check(self.lib.dg_d2s_prelu_bwd(self.ctx, C.byref(tg), C.byref(tu), _lib.ptr(prelu.data) if prelu is not None else None,
                                            C.byref(tdu), da, acc, ws.data_ptr(), nbytes, self.st))
            return [du if need_in[0] else None]

        self._push([u], out, group, bwd, params=(prelu,))
        return out

    def add(self, a: Var, b: Var) -> Var:
        seq = self._next()
        y = self.buf((seq, "y"), a.shape, a.t.dtype)
        ta, tb, ty = tensor(a.t), tensor(b.t), tensor(y)
        check(self.lib.dg_add(self.ctx, C.byref(ta), C.byref(tb), C.byref(ty), self.st))
        out = Var(y, self._deps([a, b]), seq)
        self._push([a, b], out, None, lambda gy, need_in, need_p, tag: [gy if need_in[0] else None, gy if need_in[1] else None])
        return out

    def concat(self, parts: list[Var]) -> Var:
        """tf.concat(parts, axis=3) (pix2pix.py:188,200; autoencoder.py:135)."""
        N, H, W, _ = parts[0].shape
        ctot = sum(p.shape[3] for p in parts)
        seq = self._next()
        y = self.buf((seq, "y"), (N, H, W, ctot), parts[0].t.dtype)
        off = 0
        offs = []
        for p in parts:
            c = p.shape[3]
            tp, ty = tensor(p.t), tensor(y, c=c, coff=off)
            check(self.lib.dg_copy(self.ctx, C.byref(tp), C.byref(ty), 0, self.st))
            offs.append(off)
            off += c
        out = Var(y, self._deps(parts), seq)
        out.segs = self._norm_segs([sg for p in parts for sg in (p.segs or ((p.shape[3], p.shape[3]),))])

        def bwd(gy, need_in, need_p, tag):
            res = []
            for i, p in enumerate(parts):
                if not need_in[i]:
                    res.append(None)
                    continue
                g = self.buf((seq, "dpart", i, tag), p.shape, gy.dtype)
                tg, tgi = tensor(gy, c=p.shape[3], coff=offs[i]), tensor(g)
                check(self.lib.dg_copy(self.ctx, C.byref(tg), C.byref(tgi), 0, self.st))
                res.append(g)
            return res

        self._push(list(parts), out, None, bwd)
        return out

    def concat_buffer(self, shape, dtype, splits):
        """A concat result allocated BEFORE its parts exist, and the channel-slice views the producing layers write into
        (`out=` of conv2d / bn_act): tf.concat([up, skip], axis=3) of the pix2pix U-Net (pix2pix.py:188) without a copy."""
        seq = self._next()
        y = self.buf((seq, "cat"), shape, dtype)
        views, off = [], 0
        for c in splits:
            views.append(y[..., off:off + c])
            off += c
        assert off == shape[3]
        return y, views

    def concat_views(self, y: torch.Tensor, parts: list[Var]) -> Var:
        """tf.concat(parts, axis=3) whose parts were written in place into the channel slices of `y` (concat_buffer): no forward
        copy, and the gradient of every part is the matching channel slice of the gradient of the result (a view: the consumers
        read it with the result's pixel pitch)."""
        off, offs = 0, []
        for p_ in parts:
            assert p_.t.data_ptr() == y.data_ptr() + off * y.element_size() and tuple(p_.shape[:3]) == tuple(y.shape[:3]), \
                "part is not the slice of the concat buffer"
            offs.append(off)
            off += p_.shape[3]
        assert off == y.shape[3]
        seq = self._next()
        out = Var(y, self._deps(parts), seq)

        def bwd(gy, need_in, need_p, tag):
            return [gy[..., offs[i]:offs[i] + p_.shape[3]] if need_in[i] else None for i, p_ in enumerate(parts)]

        self._push(list(parts), out, None, bwd)
        return out

    def maxpool2x2(self, x: Var) -> Var:
        N, H, W, Cc = x.shape
        seq = self._next()
        y = self.buf((seq, "y"), (N, H // 2, W // 2, Cc), x.t.dtype)
        tx, ty = tensor(x.t), tensor(y)
        check(self.lib.dg_maxpool2x2_fwd(self.ctx, C.byref(tx), C.byref(ty), self.st))
        out = Var(y, self._deps([x]), seq)
        out.segs = x.segs        # per-channel op: zero padding stays zero (max of zeros; its gradient is dropped by the consumers)

        def bwd(gy, need_in, need_p, tag):
            if not need_in[0]:
                return [None]
            dx = self.buf((seq, "dx", tag), x.shape, gy.dtype)
            tg, tdx = tensor(gy), tensor(dx)
            fn = self.lib.dg_maxpool2x2_bwd_relu if x.relu_out else self.lib.dg_maxpool2x2_bwd
            check(fn(self.ctx, C.byref(tg), C.byref(tx), C.byref(ty), C.byref(tdx), self.st))
            return [dx]

        self._push([x], out, None, bwd, masks=[x] if x.relu_out else ())
        return out

    def upsample2x_relu(self, x: Var) -> Var:
        N, H, W, Cc = x.shape
        seq = self._next()
        y = self.buf((seq, "y"), (N, 2 * H, 2 * W, Cc), x.t.dtype)
        tx, ty = tensor(x.t), tensor(y)
        check(self.lib.dg_upsample2x_relu_fwd(self.ctx, C.byref(tx), C.byref(ty), self.st))
        out = Var(y, self._deps([x]), seq)
        out.segs = x.segs

        def bwd(gy, need_in, need_p, tag):
            if not need_in[0]:
                return [None]
            dx = self.buf((seq, "dx", tag), x.shape, gy.dtype)
            tg, tdx = tensor(gy), tensor(dx)
            check(self.lib.dg_upsample2x_relu_bwd(self.ctx, C.byref(tg), C.byref(tx), C.byref(tdx), self.st))
            return [dx]

        self._push([x], out, None, bwd, masks=[x])      # the gradient of relu(up(x)) carries (x > 0)
        return out

    def pad_input(self, x: Var) -> Var:
        """The network input (3-channel image, autoencoder.py:150) as a bf16 tensor zero-padded to 16 channels: consumed in
        place by the first convolution and by the last U-Net concat (autoencoder.py:181) instead of being padded twice."""
        N, H, W, c = x.shape
        cp = -(-c // 16) * 16
        seq = self._next()
        y = self.buf((seq, "y"), (N, H, W, cp), torch.bfloat16)
        tx, ty = tensor(x.t), tensor(y)
        check(self.lib.dg_pad_channels(self.ctx, C.byref(tx), C.byref(ty), self.st))
        out = Var(y, self._deps([x]), seq)
        out.segs = self._norm_segs([(c, cp)])

        def bwd(gy, need_in, need_p, tag):
            if not need_in[0]:
                return [None]
            dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
            tg, tdx = tensor(gy, c=c), tensor(dx)
            check(self.lib.dg_copy(self.ctx, C.byref(tg), C.byref(tdx), 0, self.st))
            return [dx]

        self._push([x], out, None, bwd)
        return out

    def upsample_concat(self, a: Var, b: Var) -> Var:
        """concat([relu(UpSampling2D(2)(a)), b], axis=3) (autoencoder.py:113-136) with the up-sampled part written straight into
        its channel slice of the result, and its gradient read straight from that slice: no pass over the large half of the
        concat buffer in either direction."""
        N, H, W, ca = a.shape
        cb = b.shape[3]
        assert tuple(b.shape[:3]) == (N, 2 * H, 2 * W) and a.t.dtype == b.t.dtype
        seq = self._next()
        # a concat of 16 (mod 32) physical channels would make its consumer walk 16-channel K chunks (32-byte rows: half the bytes per
        # TMA row and per MMA operand fetch -- the 80-channel last concat of the autoencoder ran 441 us at 1080p against ~250 for 96): the
        # buffer gets 16 more zero channels, counted as padding of the last segment; only the two slices are ever written
        tail = 16 if (self.concat_pad32 and self.phys_pad and a.t.dtype == torch.bfloat16 and (ca + cb) % 32 == 16 and ca + cb > 32) else 0
        y = (self._zeros((seq, "ycat"), (N, 2 * H, 2 * W, ca + cb + tail), a.t.dtype) if tail else
             self.buf((seq, "y"), (N, 2 * H, 2 * W, ca + cb), a.t.dtype))
        ta, tya = tensor(a.t), tensor(y, c=ca, coff=0)
        check(self.lib.dg_upsample2x_relu_fwd(self.ctx, C.byref(ta), C.byref(tya), self.st))
        tb, tyb = tensor(b.t), tensor(y, c=cb, coff=ca)
        check(self.lib.dg_copy(self.ctx, C.byref(tb), C.byref(tyb), 0, self.st))
        out = Var(y, self._deps([a, b]), seq)
        segs_b = list(b.segs or ((cb, cb),))
        if tail:
            segs_b[-1] = (segs_b[-1][0], segs_b[-1][1] + tail)
        out.segs = self._norm_segs(list(a.segs or ((ca, ca),)) + segs_b)

        def bwd(gy, need_in, need_p, tag):
            da = db = None
            if need_in[0]:
                da = self.buf((seq, "da", tag), a.shape, gy.dtype)
                tg, tda = tensor(gy, c=ca, coff=0), tensor(da)
                check(self.lib.dg_upsample2x_relu_bwd(self.ctx, C.byref(tg), C.byref(ta), C.byref(tda), self.st))
            if need_in[1]:
                db = self.buf((seq, "db", tag), b.shape, gy.dtype)
                tg, tdb = tensor(gy, c=cb, coff=ca), tensor(db)
                check(self.lib.dg_copy(self.ctx, C.byref(tg), C.byref(tdb), 0, self.st))
            return [da, db]

        self._push([a, b], out, None, bwd, masks=[a])   # da carries (a > 0)
        return out

    def cast(self, x: Var, dtype) -> Var:
        """dtype conversion (e.g. the fp32 'generator_tanh' output feeding a bf16 network)."""
        if x.t.dtype == dtype:
            return x
        seq = self._next()
        y = self.buf((seq, "y"), x.shape, dtype)
        tx, ty = tensor(x.t), tensor(y)
        check(self.lib.dg_copy(self.ctx, C.byref(tx), C.byref(ty), 0, self.st))
        out = Var(y, self._deps([x]), seq)

        def bwd(gy, need_in, need_p, tag):
            if not need_in[0]:
                return [None]
            dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
            tg, tdx = tensor(gy), tensor(dx)
            check(self.lib.dg_copy(self.ctx, C.byref(tg), C.byref(tdx), 0, self.st))
            return [dx]

        self._push([x], out, None, bwd)
        return out

    def vgg_preprocess(self, x: Var) -> Var:
        """vgg19.preprocess_input(((x+1)*255)/2) in caffe mode; kept in fp32 (values are O(128))."""
        seq = self._next()
        y = self.buf((seq, "y"), x.shape, torch.float32)
        tx, ty = tensor(x.t), tensor(y)
        check(self.lib.dg_vgg_preprocess_fwd(self.ctx, C.byref(tx), C.byref(ty), self.st))
        out = Var(y, self._deps([x]), seq)

        def bwd(gy, need_in, need_p, tag):
            if not need_in[0]:
                return [None]
            dx = self.buf((seq, "dx", tag), x.shape, x.t.dtype)
            tg, tdx = tensor(gy), tensor(dx)
            check(self.lib.dg_vgg_preprocess_bwd(self.ctx, C.byref(tg), C.byref(tdx), self.st))
            return [dx]

        self._push([x], out, None, bwd)
        return out

    # ------------------------------------------------------------------ losses (value + seed gradient)
    def image_losses(self, gen: Var, target: torch.Tensor, w_mae, w_mse, w_tv, key="img"):
        """Returns (out3 = [mae, mse, tv_mean] device tensor, dgen seed gradient)."""
        out3 = self.buf((key, "out3"), (3,), torch.float32)
        dgen = self.buf((key, "dgen"), gen.shape, gen.t.dtype)
        tg, tt, td = tensor(gen.t), tensor(target), tensor(dgen)
        nbytes = self.lib.dg_loss_workspace_bytes(C.byref(tg))
        ws = self.workspace(nbytes)
        check(self.lib.dg_image_losses(self.ctx, C.byref(tg), C.byref(tt), float(w_mae), float(w_mse), float(w_tv), out3.data_ptr(),
                                       C.byref(td), 0, ws.data_ptr(), nbytes, self.st))
        return out3, dgen

    def bce(self, x: Var, target: float, from_logits: bool, grad_scale: float, key):
        """Returns (mean BCE device scalar, seed gradient grad_scale * dBCE/dx)."""
        loss = self.buf((key, "loss"), (1,), torch.float32)
        dx = self.buf((key, "dx"), x.shape, x.t.dtype)
        tx, td = tensor(x.t), tensor(dx)
        nbytes = self.lib.dg_loss_workspace_bytes(C.byref(tx))
        ws = self.workspace(nbytes)
        check(self.lib.dg_bce_const_target(self.ctx, C.byref(tx), float(target), 1 if from_logits else 0, float(grad_scale),
                                           loss.data_ptr(), C.byref(td), ws.data_ptr(), nbytes, self.st))
        return loss, dx

    def feature_mse(self, a: Var, b: Var, inv_div: float, key="feat"):
        loss = self.buf((key, "loss"), (1,), torch.float32)
        da = self.buf((key, "da"), a.shape, a.t.dtype)
        ta, tb, td = tensor(a.t), tensor(b.t), tensor(da)
        nbytes = self.lib.dg_loss_workspace_bytes(C.byref(ta))
        ws = self.workspace(nbytes)
        check(self.lib.dg_feature_mse(self.ctx, C.byref(ta), C.byref(tb), float(inv_div), loss.data_ptr(), C.byref(td), ws.data_ptr(),
                                      nbytes, self.st))
        return loss, da

    # ------------------------------------------------------------------ reverse pass
    def backward(self, seeds, group: str, tag=None, collect=None, on_node=None):
        """Equivalent of `tape.gradient(loss, <variables of group>)`: `seeds` = [(Var, dL/dVar tensor)].
        `on_node()` is called after every visited tape node (the data-parallel exchange launches the gradient buckets
        whose variables are complete, parallel.GradAllReduce.poll)."""
        tag = tag or group
        grads: dict = {}

        def accum(v: Var, g: torch.Tensor):
            cur = grads.get(v.seq)
            if cur is None:
                grads[v.seq] = [g, False]
                return
            t, owned = cur
            if not owned:
                n = self.buf((v.seq, "gacc", tag), t.shape, t.dtype)
                ta, tb, tn = tensor(t), tensor(g), tensor(n)
                check(self.lib.dg_add(self.ctx, C.byref(ta), C.byref(tb), C.byref(tn), self.st))
                grads[v.seq] = [n, True]
            else:
                tg, tt = tensor(g), tensor(t)
                check(self.lib.dg_copy(self.ctx, C.byref(tg), C.byref(tt), 1, self.st))

        for v, g in seeds:
            v.n_cons += 1          # a gradient fed in from outside carries no ReLU mask (Var.n_mask)
            accum(v, g)
        # consumers still to be visited per Var: the visit that brings the count to zero produces the LAST contribution to that
        # Var's gradient and may fold the sum accumulated so far (skip connections) into its own epilogue
        # writers still to be visited per variable of this group: a variable's gradient is COMPLETE (self.complete) once every
        # node that accumulates into it has been visited (pix2pix's generator runs twice per step, pix2pix.py:90)
        pending: dict = {}
        for node in self.tape:
            if node.group == group:
                for q in node.params:
                    pending[q.name] = pending.get(q.name, 0) + 1
        self.complete = set()
        self._complete_wait = set()
        self._in_backward = True
        self._bwd_tag = tag
        self._wq_seen = {}
        self._wq_expect = {}
        for node in self.tape:
            if node.group == group and node.wkey is not None:
                self._wq_expect[node.wkey] = self._wq_expect.get(node.wkey, 0) + 1
        remaining: dict = {}
        fuse = self.fuse_dgrad_bn_bwd and self.use_umma
        if fuse:
            for node in self.tape:
                if group in node.out.deps:
                    for v in node.inputs:
                        if group in v.deps:
                            remaining[v.seq] = remaining.get(v.seq, 0) + 1
        for node in reversed(self.tape):
            if group not in node.out.deps:
                continue
            ent = grads.pop(node.seq, None)
            need_in = [group in v.deps for v in node.inputs]
            need_p = node.group == group
            if ent is not None and (need_p or any(need_in)):
                if collect is not None:
                    collect[node.seq] = ent[0]
                fctx = None
                if fuse and node.fusable and need_in[0] and remaining.get(node.inputs[0].seq, 0) == 1:
                    cur = grads.get(node.inputs[0].seq)
                    fctx = {"residual": cur[0] if cur is not None else None}
                    gin = node.bwd(ent[0], need_in, need_p, tag, fctx=fctx)
                else:
                    gin = node.bwd(ent[0], need_in, need_p, tag)
                if fctx is not None and fctx.get("fused"):
                    grads[node.inputs[0].seq] = [gin[0], True]      # already holds the complete gradient
                else:
                    for v, g in zip(node.inputs, gin):
                        if g is not None:
                            accum(v, g)
            if fuse:
                for v in node.inputs:
                    if v.seq in remaining:
                        remaining[v.seq] -= 1
            if node.group == group:
                for q in node.params:
                    pending[q.name] -= 1
                    if pending[q.name] == 0:
                        # a gradient still sitting in the batching queue becomes complete when its group is launched
                        (self._complete_wait if q.name in self._wq_names else self.complete).add(q.name)
            if on_node is not None and ent is not None:
                on_node()
        self._bwd_part = {k: v for k, v in self._bwd_part.items() if k[1] != tag}
        self._in_backward = False
        self._wgrad_flush()
        self._join_side()

    # ------------------------------------------------------------------ optimiser
    def adam(self, pset, lr0, beta1=0.9, beta2=0.999, eps=1e-7, decay_steps=0, decay_rate=0.1, grad_scale=1.0):
        """Keras Adam over the whole arena; then refresh the packed bf16 kernels."""
        check(self.lib.dg_adam_step(self.ctx, pset.theta.data_ptr(), pset.grad.data_ptr(), pset.m.data_ptr(), pset.v.data_ptr(),
                                    pset.numel, float(lr0), float(beta1), float(beta2), float(eps), int(decay_steps), float(decay_rate),
                                    float(grad_scale), pset.opt_state.data_ptr(), self.st))
        pset.version = getattr(pset, "version", 0) + 1
        pset.repack(self.lib, self.ctx, self.st)
