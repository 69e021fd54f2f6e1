"""The image summaries of the reference's training loop (train_srgan.py:150-172; same block in train_fsrgan.py,
train_autoencoder.py, train_pix2pix.py): every `log_iter` iterations the generator is run once more in inference mode on the
current batch and the FIRST image of input / target / generated is written next to error and gradient diagnostics.  The
arithmetic (renorm, autoscale, Sobel magnitude, differences, total variation, uint8 conversion) runs on the device
(dg_image_summary, csrc/summaries.cu); what comes back is one small uint8 image per tag."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

IMAGE, SQUARE, ABS, SOBEL, DX, DY, TV = range(7)

# (tag, kind, minuend, subtrahend) in the order of train_srgan.py:156-172
SUMMARIES = (
    ("Images/Input", IMAGE, "input", None),
    ("Images/Target", IMAGE, "target", None),
    ("Images/Generated", IMAGE, "gen", None),
    ("Error/Square Error (MSE)", SQUARE, "gen", "target"),
    ("Error/Absolute Error (MAE)", ABS, "gen", "target"),
    ("Error/Sobel Variation", SOBEL, "gen", "target"),
    ("Error/Total Variation", TV, "gen", "target"),
    ("Image Gradients/Sobel Input", SOBEL, "input", None),
    ("Image Gradients/Sobel Target", SOBEL, "target", None),
    ("Image Gradients/Sobel Generated", SOBEL, "gen", None),
    ("Image Gradients/dx Target", DX, "target", None),
    ("Image Gradients/dy Target", DY, "target", None),
    ("Image Gradients/dx Generated", DX, "gen", None),
    ("Image Gradients/dy Generated", DY, "gen", None),
    ("Image Gradients/Total Var Target", TV, "target", None),
    ("Image Gradients/Total Var Generated", TV, "gen", None),
)


def image_summary(kind: int, a: torch.Tensor, b: torch.Tensor | None = None, ws: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 [h', w', c] device tensor: summary `kind` of the first image of the NHWC batch a (minus b's, if given)."""
    lib = _lib.load()
    ctx = _lib.ctx(a.device.index or 0)
    n, h, w, c = a.shape
    crop = 1 if kind >= DX else 0
    out = torch.empty(h - crop, w - crop, c, dtype=torch.uint8, device=a.device)
    nbytes = lib.dg_image_summary_workspace_bytes(h, w, c)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
    ta = _lib.tensor(a.contiguous())
    tb = _lib.tensor(b.contiguous()) if b is not None else None
    _lib.check(lib.dg_image_summary(ctx, C.byref(ta), C.byref(tb) if tb is not None else None, kind, out.data_ptr(), ws.data_ptr(),
                                    ws.numel(), _lib.stream_ptr()))
    return out


def training_image_summaries(model, img_input: torch.Tensor, img_target: torch.Tensor) -> "dict[str, torch.Tensor]":
    """train_srgan.py:150-172: img_gen = model.generator(img_input, training=False), then the sixteen image summaries."""
    model.engine.new_step()
    gen = model.generator(img_input, training=False).t
    src = {"input": img_input.float(), "target": img_target.float(), "gen": gen.float()}
    h, w, c = max(tuple(t.shape[1:]) for t in src.values())
    ws = torch.empty(_lib.load().dg_image_summary_workspace_bytes(h, w, c), dtype=torch.uint8, device=gen.device)
    return {tag: image_summary(kind, src[a], src[b] if b else None, ws) for tag, kind, a, b in SUMMARIES}


def write_image_summaries(writer, images: "dict[str, torch.Tensor]", step: int):
    """Hands the images to `writer`: `add_image(tag, HWC uint8 array, step, dataformats='HWC')` (torch / TensorBoard
    SummaryWriter), a callable `(tag, array, step)`, or nothing."""
    if writer is None:
        return
    for tag, img in images.items():
        arr = img.cpu().numpy()
        if hasattr(writer, "add_image"):
            writer.add_image(tag, arr, step, dataformats="HWC")
        elif callable(writer):
            writer(tag, arr, step)
