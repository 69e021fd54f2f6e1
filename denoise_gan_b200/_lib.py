"""ctypes binding of the dg_b200 C ABI (include/dg_b200.h).

PyTorch is used for device memory and streams only; every arithmetic kernel on the hot path is
reached through this module.  There is no CPU or eager-PyTorch fallback: a missing library or a
non-zero return code raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DG_LIB_PATH") or os.path.join(_HERE, "libdg_b200.so")   # DG_LIB_PATH: A/B experiments against another build

DG_F32, DG_BF16 = 0, 1
ACT = {None: 0, "none": 0, "linear": 0, "relu": 1, "lrelu": 2, "tanh": 3, "sigmoid": 4, "prelu": 5}


class DgTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32), ("cpitch", C.c_int32), ("coff", C.c_int32)]


class DgConvParams(C.Structure):
    _fields_ = [("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad_t", C.c_int32),
                ("pad_l", C.c_int32), ("act", C.c_int32), ("act_alpha", C.c_float)]


class DgBnFused(C.Structure):      # include/dg_b200.h: dg_bn_fused
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("momentum", C.c_float),
                ("moving_mean", C.c_void_p), ("moving_var", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("save_mean", C.c_void_p), ("save_invstd", C.c_void_p), ("pixels", C.c_longlong)]


class DgBnBwdStats(C.Structure):  # include/dg_b200.h: dg_bn_bwd_stats
    _fields_ = [("y", C.POINTER(DgTensor)), ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("act", C.c_int32),
                ("alpha", C.c_float), ("partials", C.c_void_p)]


class DgError(RuntimeError):
    pass


_lib = None
_ctx = {}

_P, _T, _CP = C.c_void_p, C.POINTER(DgTensor), C.POINTER(DgConvParams)
_i, _f, _u32, _i64, _sz = C.c_int, C.c_float, C.c_uint32, C.c_int64, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/dg_b200.h declares
SIGNATURES = {
    "dg_init": (_i, [_i, C.POINTER(_P)]),
    "dg_destroy": (None, [_P]),
    "dg_last_error": (C.c_char_p, []),
    "dg_version": (_i, []),
    "dg_has_umma": (_i, [_P]),
    "dg_comm_unique_id_bytes": (_i, []),
    "dg_comm_unique_id": (_i, [_P]),
    "dg_comm_init": (_i, [C.POINTER(_P), _P, _i, _i, _i]),
    "dg_comm_allreduce": (_i, [_P, _P, C.c_longlong, _P]),
    "dg_comm_rank": (_i, [_P]),
    "dg_comm_world": (_i, [_P]),
    "dg_comm_destroy": (None, [_P]),
    "dg_conv2d_fwd": (_i, [_P, _T, _P, _P, _T, _CP, _P]),
    "dg_conv2d_dgrad": (_i, [_P, _T, _P, _P, _T, _CP, _P]),
    "dg_conv2d_wgrad_workspace_bytes": (_sz, [_T, _T, _CP]),
    "dg_conv2d_wgrad": (_i, [_P, _T, _T, _P, _P, _CP, _i, _P, _sz, _P]),
    "dg_umma_packed_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "dg_umma_pack_weights": (_i, [_P, _P, _P, _i, _i, _i, _i, _i, _P]),
    "dg_umma_pack_weights_batch": (_i, [_P, _P, _i, _P]),
    "dg_umma_conv2d_fwd": (_i, [_P, _T, _P, _P, _T, _CP, _P, _P]),
    "dg_umma_conv2d_fwd_narrow": (_i, [_P, _T, _P, _P, _T, _CP, _P]),
    "dg_umma_conv2d_fwd_res_prelu": (_i, [_P, _T, _P, _P, _T, _CP, _T, _P, _P]),
    "dg_umma_conv2d_fwd_d2s_prelu": (_i, [_P, _T, _P, _P, _T, _CP, _P, _P]),
    "dg_umma_conv2d_dgrad": (_i, [_P, _T, _P, _P, _T, _CP, _P]),
    "dg_umma_conv2d_dgrad_fused": (_i, [_P, _T, _P, _T, _CP, _T, C.POINTER(DgBnBwdStats), _P]),
    "dg_umma_conv2d_dgrad_fused_blocks": (_i, [_P, _T, _T, _CP]),
    "dg_umma_conv2d_dgrad_relu_mask": (_i, [_P, _T, _P, _T, _CP, _T, _P]),
    "dg_bn_bwd_dx_from_partials": (_i, [_P, _T, _T, _P, _P, _P, _P, _P, _i, _f, _P, _i, _T, _P, _P, _i, _P]),
    "dg_umma_conv2d_fwd_supported": (_i, [_P, _T, _T, _CP]),
    "dg_umma_conv2d_fwd_bn_blocks": (_i, [_P, _T, _T, _CP]),
    "dg_umma_conv2d_fwd_bn": (_i, [_P, _T, _P, _P, _T, _CP, _P, C.POINTER(DgBnFused), _P]),
    "dg_umma_conv2d_fwd_bn_act": (_i, [_P, _T, _P, _P, _T, _CP, _P, C.POINTER(DgBnFused), _i, _f, _P, _T, _T, _P]),
    "dg_umma_conv2d_fwd_bn_act_blocks": (_i, [_P, _T, _T, _CP]),
    "dg_bn_finalize": (_i, [_P, _P, _i, C.c_longlong, _i, _P, _P, _f, _f, _P, _P, _P, _P, _P, _P, _P]),
    "dg_bn_act_fwd_from_partials": (_i, [_P, _T, _P, _i, _P, _P, _f, _f, _P, _P, _P, _P, _P, _P, _i, _f, _P, _T, _T, _P]),
    "dg_umma_conv2d_dgrad_supported": (_i, [_P, _T, _T, _CP]),
    "dg_bias_grad": (_i, [_P, _T, _P, _i, _P, _sz, _P]),
    "dg_pad_channels": (_i, [_P, _T, _T, _P]),
    "dg_umma_pack_weights_padded": (_i, [_P, _P, _P, _i, _i, _i, _i, _i, _i, _i, _P]),
    "dg_unpad_weight_grad": (_i, [_P, _P, _P, _P, _P, _i, _i, _i, _i, _i, _i, _i, _P]),
    "dg_umma_pack_weights_seg": (_i, [_P, _P, _P, _i, _i, _i, _i, _i, _i, _i, _i, _i, _P]),
    "dg_unpad_weight_grad_seg": (_i, [_P, _P, _P, _P, _P, _i, _i, _i, _i, _i, _i, _i, _i, _i, _P]),
    "dg_image_summary_workspace_bytes": (_sz, [_i, _i, _i]),
    "dg_image_summary": (_i, [_P, _T, _T, _i, _P, _P, _sz, _P]),
    "dg_pair_synthesis_workspace_bytes": (_sz, [_i, _i, _i]),
    "dg_pair_synthesis": (_i, [_P, _P, _i, _i, _i, _P, _P, _P, _i, _i, _i, _i, _P, _P, _P, _sz, _P]),
    "dg_frame_to_float": (_i, [_P, _P, _i, _i, _i, _i, _f, _f, _T, _P]),
    "dg_float_to_frame": (_i, [_P, _T, _f, _f, _i, _i, _P, _i, _i, _P]),
    "dg_debug_conv_timeline": (None, [_P]),
    "dg_debug_conv_flags": (None, [_i]),
    "dg_debug_wgrad_timeline": (None, [_P]),
    "dg_im2col": (_i, [_P, _T, _CP, _i, _i, _P, _P]),
    "dg_umma_conv2d_wgrad_workspace_bytes": (_sz, [_T, _T, _CP]),
    "dg_umma_conv2d_wgrad": (_i, [_P, _T, _T, _P, _P, _CP, _i, _P, _sz, _P]),
    "dg_umma_conv2d_wgrad_batch_supported": (_i, [_P, _i, _T, _T, _CP]),
    "dg_umma_conv2d_wgrad_batch_workspace_bytes": (_sz, [_i, _T, _T, _CP]),
    "dg_umma_conv2d_wgrad_batch": (_i, [_P, _i, _P, _P, _P, _P, _CP, _P, _P, _sz, _P]),
    "dg_dwconv3x3_fwd": (_i, [_P, _T, _P, _P, _T, _P]),
    "dg_fsrgan_block_infer_supported": (_i, [_P, _T, _T]),
    "dg_debug_fsrgan_block_timeline": (None, [_P]),
    "dg_fsrgan_block_infer": (_i, [_P, _T, _P, _P, _P, _P, _P, _P, _T, _P]),
    "dg_conv3x3_tapsum_supported": (_i, [_P, _T, _i]),
    "dg_conv3x3_tapsum_fwd": (_i, [_P, _T, _P, _P, _i, _f, _T, _P]),
    "dg_conv3x3_tapsum_frame": (_i, [_P, _T, _P, _P, _i, _f, _f, _f, _i, _i, _P, _i, _i, _P]),
    "dg_dwconv3x3_fwd_act": (_i, [_P, _T, _P, _P, _i, _T, _P]),
    "dg_dwconv3x3_dgrad": (_i, [_P, _T, _P, _T, _P]),
    "dg_dwconv3x3_wgrad_workspace_bytes": (_sz, [_T]),
    "dg_dwconv3x3_wgrad": (_i, [_P, _T, _T, _P, _P, _i, _P, _sz, _P]),
    "dg_bn_workspace_bytes": (_sz, [_T]),
    "dg_bn_stats": (_i, [_P, _T, _P, _P, _f, _f, _P, _P, _P, _P, _P, _P, _P, _sz, _P]),
    "dg_bn_infer_affine": (_i, [_P, _i, _P, _P, _P, _P, _f, _P, _P, _P]),
    "dg_bn_act_fwd": (_i, [_P, _T, _P, _P, _i, _f, _P, _T, _i, _u32, _u32, _P, _T, _P]),
    "dg_bn_train_fwd": (_i, [_P, _T, _P, _P, _f, _f, _P, _P, _P, _P, _P, _P, _i, _f, _P, _T, _i, _u32, _u32, _P, _T, _P, _sz, _P]),
    "dg_bn_act_bwd": (_i, [_P, _T, _T, _P, _P, _P, _P, _P, _i, _f, _P, _i, _u32, _u32, _P, _T, _P, _P, _P, _i, _P, _sz, _P]),
    "dg_act_bwd_from_output": (_i, [_P, _T, _T, _i, _f, _T, _P]),
    "dg_d2s_prelu_fwd": (_i, [_P, _T, _P, _T, _P]),
    "dg_d2s_prelu_bwd": (_i, [_P, _T, _T, _P, _T, _P, _i, _P, _sz, _P]),
    "dg_add": (_i, [_P, _T, _T, _T, _P]),
    "dg_copy": (_i, [_P, _T, _T, _i, _P]),
    "dg_maxpool2x2_fwd": (_i, [_P, _T, _T, _P]),
    "dg_maxpool2x2_bwd": (_i, [_P, _T, _T, _T, _T, _P]),
    "dg_maxpool2x2_bwd_relu": (_i, [_P, _T, _T, _T, _T, _P]),
    "dg_upsample2x_relu_fwd": (_i, [_P, _T, _T, _P]),
    "dg_upsample2x_relu_bwd": (_i, [_P, _T, _T, _T, _P]),
    "dg_vgg_preprocess_fwd": (_i, [_P, _T, _T, _P]),
    "dg_vgg_preprocess_bwd": (_i, [_P, _T, _T, _P]),
    "dg_loss_workspace_bytes": (_sz, [_T]),
    "dg_image_losses": (_i, [_P, _T, _T, _f, _f, _f, _P, _T, _i, _P, _sz, _P]),
    "dg_gan_loss_terms": (_i, [_P, _P, _P, _P, _P, _P, _f, _f, _f, _f, _P, _P]),
    "dg_bce_const_target": (_i, [_P, _T, _f, _i, _f, _P, _T, _P, _sz, _P]),
    "dg_feature_mse": (_i, [_P, _T, _T, _f, _P, _T, _P, _sz, _P]),
    "dg_adam_step": (_i, [_P, _P, _P, _P, _P, _i64, _f, _f, _f, _f, _i64, _f, _f, _P, _P]),
}


def load():
    """Loads libdg_b200.so and binds every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DgError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ctx(device: int | None = None):
    lib = load()
    if device is None:
        device = torch.cuda.current_device()
    if device not in _ctx:
        h = C.c_void_p()
        rc = lib.dg_init(int(device), C.byref(h))
        if rc != 0:
            raise DgError(lib.dg_last_error().decode())
        _ctx[device] = h
    return _ctx[device]


def new_ctx(device: int):
    """A second, private context (own reduction scratch) for launches on a second stream of the same device."""
    lib = load()
    h = C.c_void_p()
    rc = lib.dg_init(int(device), C.byref(h))
    if rc != 0:
        raise DgError(lib.dg_last_error().decode())
    return h


def destroy_ctx(h):
    """Frees a context made by new_ctx (called from the owning Engine's finalizer)."""
    try:
        if _lib is not None and h:
            _lib.dg_destroy(h)
    except Exception:
        pass


CALLS = 0  # number of C-ABI compute calls issued (each enqueues >= 1 kernel)


def check(rc: int):
    global CALLS
    CALLS += 1
    if rc != 0:
        raise DgError(_lib.dg_last_error().decode())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DG_F32
    if t.dtype == torch.bfloat16:
        return DG_BF16
    raise DgError(f"unsupported dtype {t.dtype}")


def tensor(t: torch.Tensor, c: int | None = None, coff: int = 0) -> DgTensor:
    """NHWC view descriptor of `t`: a dense [N,H,W,C] tensor (optionally a logical channel slice c / coff of it), or a torch
    view `buf[..., a:b]` of a dense NHWC buffer (the halves of an in-place U-Net concat): the pixel pitch comes from the strides."""
    assert t.is_cuda and t.dim() == 4, "expected a CUDA NHWC tensor"
    n, h, w, cc = t.shape
    if t.is_contiguous():
        cp = cc
    else:
        sn, sh, sw, sc = t.stride()
        cp = sw if w > 1 else (sh if h > 1 else sn)
        assert (sc == 1 and cp >= cc and (w == 1 or sw == cp) and (h == 1 or sh == w * cp) and (n == 1 or sn == h * w * cp)
                and c is None and coff == 0), "expected a dense NHWC tensor or a channel slice of one"
    return DgTensor(t.data_ptr(), dtype_code(t), n, h, w, cc if c is None else c, cp, coff)


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
