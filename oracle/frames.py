"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the frame pre/post arithmetic
around the inference forward.  Parity unpinned (no TensorFlow here; tf.image.* semantics restated from the
documented behaviour).

References: infer_video.py:79-83 (padded size), :138-159 (per-frame pre/post), infer.py:50-68,
unit_test.py:67-86.
"""
from __future__ import annotations

import numpy as np


def padded_size(fh: int, fw: int, block: int = 256, model_scale: int = 1):
    """infer_video.py:79-83: next multiple of 256*model_scale strictly above the frame size."""
    m = block * model_scale
    return (fh + m) - fh % m, (fw + m) - fw % m


def resize_with_crop_or_pad(img: np.ndarray, th: int, tw: int) -> np.ndarray:
    """tf.image.resize_with_crop_or_pad on [H,W,C]: centre crop (offset = diff // 2) and/or centre zero pad
    (before = diff // 2, remainder after)."""
    h, w = img.shape[:2]
    if h > th:
        o = (h - th) // 2
        img = img[o:o + th]
    if w > tw:
        o = (w - tw) // 2
        img = img[:, o:o + tw]
    h, w = img.shape[:2]
    pt, pl = (th - h) // 2 if th > h else 0, (tw - w) // 2 if tw > w else 0
    out = np.zeros((th, tw) + img.shape[2:], dtype=img.dtype)
    out[pt:pt + h, pl:pl + w] = img
    return out


def video_pre(frame_bgr_u8: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """infer_video.py:140-145: BGR->RGB, convert_image_dtype(float32) (= x * float32(1/255)), crop-or-pad, *2-1."""
    rgb = frame_bgr_u8[..., ::-1].astype(np.float32) * np.float32(1.0 / 255.0)
    return resize_with_crop_or_pad(rgb, new_h, new_w) * np.float32(2) - np.float32(1)


def video_post(frame_out_f32: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """infer_video.py:149-159: *0.5+0.5, crop-or-pad to (fh*scale, fw*scale), clip [0,1], *255, astype(uint8). RGB out."""
    v = frame_out_f32.astype(np.float32) * np.float32(0.5) + np.float32(0.5)
    v = np.clip(resize_with_crop_or_pad(v, out_h, out_w), 0, 1).astype(np.float32)
    return (v * np.float32(255)).astype(np.uint8)


def still_pre(img_bgr_u8: np.ndarray) -> np.ndarray:
    """infer.py:52-58: BGR->RGB, / 255.0 in float64; model.predict casts to float32."""
    return (img_bgr_u8[..., ::-1] / 255.0).astype(np.float32)


def still_post(sr_f32: np.ndarray) -> np.ndarray:
    """infer.py:64-67: ((sr+1)/2)*255, RGB->BGR, astype(uint8)."""
    v = ((sr_f32.astype(np.float32) + np.float32(1)) / np.float32(2.0)) * np.float32(255)
    return np.clip(v, 0, 255)[..., ::-1].astype(np.uint8)


def unit_pre(img_bgr_u8: np.ndarray) -> np.ndarray:
    """unit_test.py:67-74: crop [:256,:256], BGR->RGB, float32 / 255.0."""
    return img_bgr_u8[:256, :256, ::-1].astype(np.float32) / np.float32(255.0)


def unit_post(sr_f32: np.ndarray) -> np.ndarray:
    """unit_test.py:85-86: np.uint8(((sr+1)/2)*255), RGB (converted to BGR at :90)."""
    v = ((sr_f32.astype(np.float32) + np.float32(1)) / np.float32(2.0)) * np.float32(255)
    return np.clip(v, 0, 255).astype(np.uint8)
