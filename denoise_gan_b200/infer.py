"""Inference entry points: the generator forward with `training=False` (BatchNorm on its moving statistics,
dropout off) wrapped in the reference's frame arithmetic.

Reference: infer_video.py:92-97 (model re-wrapped on Input((None,None,3)), training=False), :138-159 (per-frame
pre/post), :79-83 (padded size); infer.py:38-68 (still images); unit_test.py:56-86 (256x256 crops).

Only uint8 frames cross PCIe: the crop-or-pad, scaling and channel flip run on the device
(csrc/frames.cu) next to the forward; host staging buffers are pinned and double-buffered in BOTH directions, so the
upload of frame i+1 and the download of frame i-1 overlap the forward of frame i (`FrameRunner.video`), whose launches are one CUDA-graph
replay from the third frame of a staging slot on (`device_frame`).  Frames shard round-robin across ranks with no collective
(`frames_for_rank`).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib


def padded_size(fh: int, fw: int, block: int = 256, model_scale: int = 1):
    """infer_video.py:79-83."""
    m = block * model_scale
    return (fh + m) - fh % m, (fw + m) - fw % m


def tight_size(fh: int, fw: int, receptive_radius: float | None = None):
    """The padded size a video frame is run at.  The reference pads every frame to the next multiple of 256 (infer_video.py:79-83,
    141) and crops the centre of the output (:152): only output pixels whose receptive field reaches the frame can differ from
    'the network on a constant image', and everything else is cropped away.  A generator that declares its receptive-field radius
    (in input pixels: stride-1 convolutions, per-pixel inference BatchNorm, no pooling) is therefore run on the frame plus a
    margin >= that radius (a multiple of 8) instead of the full padding -- the cropped frame is the same, pixel for pixel
    (tests/test_infer_gpu.py), and a 1080p Fast-SRGAN frame is 1112 x 1952 instead of 1280 x 2048 pixels of work in every layer.
    The reference's size is kept when no radius is declared, when the padding is odd (the reference's crop is then not centred
    on the frame) or when it is smaller than the margin."""
    nh, nw = padded_size(fh, fw)
    if receptive_radius is None:
        return nh, nw
    m = -(-int(np.ceil(receptive_radius)) // 8) * 8
    if (nh - fh) % 2 or (nw - fw) % 2 or (nh - fh) // 2 < m or (nw - fw) // 2 < m:
        return nh, nw
    return fh + 2 * m, fw + 2 * m


def frames_for_rank(num_frames: int, rank: int, world: int, start: int = 0):
    """Round-robin frame shard of one rank (inference: no exchange step)."""
    return range(start + rank, num_frames, world)


def _u8(frame) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(frame)) if isinstance(frame, np.ndarray) else frame.contiguous()
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError(f"expected a uint8 [H, W, 3] frame, got {t.dtype} {tuple(t.shape)}")
    return t


class FrameRunner:
    """Runs `model.generator(x, training=False)` on uint8 frames.  `model` is any of Autoencoder / SRGAN /
    FastSRGAN / Pix2Pix; `upscale` is the generator's output/input size ratio (4 for the SR models at scale 4)."""

    def __init__(self, model, upscale: int | None = None):
        self.model, self.E = model, model.engine
        self.upscale = int(upscale if upscale is not None else getattr(model, "scale", 1))
        self.copy_stream = torch.cuda.Stream(device=self.E.device)       # uploads
        self.d2h_stream = torch.cuda.Stream(device=self.E.device)        # downloads
        self.tight_padding = os.environ.get("DG_INFER_TIGHT_PAD", "1") != "0"   # video frames: frame + receptive-field margin instead of the full padding
        self.use_graphs = os.environ.get("DG_INFER_GRAPH", "1") != "0"       # video(): frames as CUDA-graph replays
        self._graphs: dict = {}
        self._graph_seen: dict = {}
        self._graph_ver = None
        self._pin: dict = {}
        self._pin_ev: dict = {}       # pinned staging buffer -> event after the last device copy that read it

    # ---- device-side pieces
    def _to_float(self, frame_dev: torch.Tensor, out_h: int, out_w: int, flip: bool, norm_mode: int, scale: float, offset: float):
        E = self.E
        h, w = frame_dev.shape[:2]
        x = E.buf(("infer_in",), (1, out_h, out_w, 3), torch.float32)
        _lib.check(E.lib.dg_frame_to_float(E.ctx, frame_dev.data_ptr(), h, w, int(flip), norm_mode, scale, offset, _lib.tensor(x), E.st))
        return x

    def _to_frame(self, y: torch.Tensor, out_h: int, out_w: int, scale: float, offset: float, clip: bool, flip: bool, slot=0):
        E = self.E
        out = E.buf(("infer_out", slot), (out_h, out_w, 3), torch.uint8)
        _lib.check(E.lib.dg_float_to_frame(E.ctx, _lib.tensor(y), scale, offset, int(clip), int(flip), out.data_ptr(), out_h, out_w, E.st))
        return out

    def _forward_to_frame(self, x: torch.Tensor, out_h: int | None, out_w: int | None, scale: float, offset: float, clip: bool, flip: bool, slot=0):
        """forward + _to_frame.  A generator whose last layer can write the uint8 frame itself (Engine.conv3x3_image_infer: the
        Fast-SRGAN output convolution) is offered the frame as a sink: the float image is then never written and pixels outside
        the centre crop are not computed.  out_h / out_w None: the generator's own output size (infer.py:62-68)."""
        E = self.E
        if out_h is None:
            out_h, out_w = x.shape[1] * self.upscale, x.shape[2] * self.upscale
        out = E.buf(("infer_out", slot), (out_h, out_w, 3), torch.uint8)
        E.frame_sink = dict(out=out, h=out_h, w=out_w, scale=scale, offset=offset, clip=clip, flip=flip, done=False)
        try:
            y = self.forward(x)
        finally:
            sink, E.frame_sink = E.frame_sink, None
        if sink["done"]:
            return out
        return self._to_frame(y, out_h, out_w, scale, offset, clip, flip, slot)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """NHWC float [-1,1] (or [0,1] for infer.py / unit_test.py callers) -> generator output, training=False."""
        self.E.new_step()
        return self.model.generator(x, training=False).t

    def _pinned(self, key, shape):
        k = (key, tuple(shape))
        if k not in self._pin:
            self._pin[k] = torch.empty(shape, dtype=torch.uint8).pin_memory()
        return self._pin[k]

    def _h2d(self, frame: torch.Tensor, slot=0) -> torch.Tensor:
        if frame.is_cuda:
            return frame
        stage = self._pinned(("in", slot), frame.shape)
        ev = self._pin_ev.get(("in", slot))
        if ev is not None:
            ev.synchronize()          # the previous asynchronous copy OUT of this pinned buffer (video_frame(..., to_host=False))
        stage.copy_(frame)
        dev = self.E.buf(("infer_u8", slot), tuple(frame.shape), torch.uint8)
        dev.copy_(stage, non_blocking=True)
        ev = torch.cuda.Event(); ev.record()
        self._pin_ev[("in", slot)] = ev
        return dev

    def compute_size(self, fh: int, fw: int):
        """The size the generator runs on for an fh x fw video frame (see `tight_size`)."""
        r = getattr(self.model.generator, "receptive_radius", None) if self.tight_padding else None
        return tight_size(fh, fw, r)

    # ---- the three reference call sites
    def video_frame(self, frame_bgr, to_host: bool = True, out_slot: int = 0):
        """infer_video.py:138-159: BGR uint8 [fh,fw,3] -> RGB uint8 [fh*s, fw*s, 3]."""
        f = _u8(frame_bgr)
        fh, fw = f.shape[:2]
        nh, nw = self.compute_size(fh, fw)
        x = self._to_float(self._h2d(f), nh, nw, flip=True, norm_mode=0, scale=2.0, offset=-1.0)
        out = self._forward_to_frame(x, fh * self.upscale, fw * self.upscale, 0.5, 0.5, clip=True, flip=False, slot=out_slot)
        return self._d2h(out) if to_host else out

    def device_frame(self, dev: torch.Tensor, slot: int) -> torch.Tensor:
        """video_frame(dev, to_host=False, out_slot=slot) for a uint8 BGR frame already in the device buffer `dev`, as a CUDA-graph replay
        from the third frame of that buffer on: a frame is ~20 launches behind a few hundred microseconds of Python each, which
        is most of what the host does per frame (and eight ranks share one host).  Every address in the frame is fixed -- staging
        buffer, pooled activations, result slot -- so the captured frame is the eager one; `DG_INFER_GRAPH=0` keeps the eager loop."""
        ver = getattr(getattr(self.model, "gen_params", None), "version", 0)
        if ver != self._graph_ver:                   # new weights (load / optimiser step): the folded kernels a captured frame reads are rebuilt eagerly
            self._graphs.clear(); self._graph_seen.clear()
            self._graph_ver = ver
        key = (tuple(dev.shape), dev.data_ptr(), slot)
        g = self._graphs.get(key)
        if g is not None:
            g[0].replay()
            return g[1]
        seen = self._graph_seen.get(key, 0)
        self._graph_seen[key] = seen + 1
        if not self.use_graphs or seen < 2:          # eager frames first: they allocate the pooled buffers and fill the caches
            return self.video_frame(dev, to_host=False, out_slot=slot)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.video_frame(dev, to_host=False, out_slot=slot)
        self._graphs[key] = (graph, out)
        graph.replay()                               # capture records, it does not run
        return out

    def still_image(self, img_bgr, to_host: bool = True):
        """infer.py:50-68: BGR uint8 -> [0,1] RGB (float64 division) -> forward -> ((sr+1)/2)*255 -> BGR uint8."""
        f = _u8(img_bgr)
        h, w = f.shape[:2]
        out = self._forward_to_frame(self._to_float(self._h2d(f), h, w, flip=True, norm_mode=1, scale=1.0, offset=0.0), None, None, 0.5, 0.5,
                                     clip=False, flip=True)
        return self._d2h(out) if to_host else out

    def unit_image(self, img_bgr, to_host: bool = True):
        """unit_test.py:67-86: top-left 256x256 crop, float32 / 255, forward, np.uint8(((sr+1)/2)*255) (RGB)."""
        f = _u8(img_bgr)[:256, :256].contiguous()
        h, w = f.shape[:2]
        out = self._forward_to_frame(self._to_float(self._h2d(f), h, w, flip=True, norm_mode=2, scale=1.0, offset=0.0), None, None, 0.5, 0.5,
                                     clip=False, flip=False)
        return self._d2h(out) if to_host else out

    def _d2h(self, out_dev: torch.Tensor, slot=0) -> torch.Tensor:
        host = self._pinned(("out", slot), out_dev.shape)
        host.copy_(out_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.clone()

    def video(self, frames, rank: int = 0, world: int = 1, copy: bool = True):
        """Generator over (index, RGB uint8 frame) for this rank's round-robin shard of `frames` (a sequence of BGR
        uint8 frames).  Three things overlap: the upload of frame k+1 (copy stream), the forward of frame k (main stream)
        and the download of frame k-1 (second copy stream, into one of two pinned buffers); the host only ever waits for
        the download event of the frame it is about to hand out, never for the compute stream.  With `copy=False` the
        yielded tensor IS the pinned staging buffer: valid until the generator has been advanced twice more (the
        reference writes each frame to its video sink immediately, infer_video.py:160-170)."""
        idx = list(frames_for_rank(len(frames), rank, world))
        if not idx:
            return
        main = torch.cuda.current_stream()

        def upload(k):
            f = _u8(frames[idx[k]])
            slot = k & 1
            stage = self._pinned(("in", slot), f.shape)
            ev0 = self._pin_ev.get(("in", slot))
            if ev0 is not None:
                ev0.synchronize()           # the copy that last read this staging buffer (frame k-2)
            stage.copy_(f)
            dev = self.E.buf(("infer_u8", slot), tuple(f.shape), torch.uint8)
            with torch.cuda.stream(self.copy_stream):
                dev.copy_(stage, non_blocking=True)
                ev = torch.cuda.Event(); ev.record()
            self._pin_ev[("in", slot)] = ev
            return dev, ev

        pending = None                      # (frame index, pinned host buffer, download-finished event) of frame k-1
        nxt = upload(0)
        for k in range(len(idx)):
            dev, ev = nxt
            main.wait_event(ev)
            slot = k & 1
            out = self.device_frame(dev, slot)                        # device result slot: its previous download (frame k-2) was waited for below
            done = torch.cuda.Event(); done.record(main)
            if k + 1 < len(idx):
                nxt = upload(k + 1)         # other input slot
            host = self._pinned(("out", slot), out.shape)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(done)
                host.copy_(out, non_blocking=True)
                hev = torch.cuda.Event(); hev.record()
            if pending is not None:
                pi, ph, pe = pending
                pe.synchronize()
                yield pi, (ph.clone() if copy else ph)
            pending = (idx[k], host, hev)
        pi, ph, pe = pending
        pe.synchronize()
        yield pi, (ph.clone() if copy else ph)
