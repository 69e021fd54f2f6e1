"""GPU parity tests, kernel by kernel, THROUGH THE C ABI (ctypes -> libdg_b200.so), against the
CPU oracle (oracle/ops_torch.py in float64 + autograd) on identical seeded inputs.

Tolerances (north star): fp32 path 1e-5 relative to the tensor's max magnitude; bf16 tensor-core
path 2e-2 (inputs are pre-rounded to bf16 so only accumulation order and output rounding differ)."""
import ctypes as C
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ops_np as ON  # noqa: E402
from oracle import ops_torch as OT  # noqa: E402

FP32_TOL = 1e-5
BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def L():
    from denoise_gan_b200 import _lib
    _lib.load()
    _lib.ctx(0)
    return _lib


def relerr(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def dev(t, dtype=torch.float32):
    return t.to(dtype).cuda().contiguous()


def ws(nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device="cuda")


def conv_params(L, kh, kw, stride, H, W, padding, act=0, alpha=0.0):
    if padding == "same":
        pt, pl = ON.same_pads(H, kh, stride)[0], ON.same_pads(W, kw, stride)[0]
    elif padding == "valid":
        pt = pl = 0
    else:
        pt, pl = padding[0][0], padding[1][0]
    return L.DgConvParams(kh, kw, stride, pt, pl, act, alpha)


CONV_CASES = [  # (k, stride, padding, cin, cout, N, H, W)
    (3, 1, "same", 3, 32, 2, 12, 10),
    (3, 1, "same", 67, 44, 1, 9, 11),
    (3, 2, "same", 32, 32, 2, 12, 10),
    (4, 2, "same", 6, 64, 2, 8, 8),
    (1, 1, "same", 64, 3, 2, 7, 9),
    (4, 1, ((1, 1), (1, 1)), 16, 1, 1, 10, 10),
    (3, 2, "same", 5, 7, 1, 7, 9),
    (3, 1, "same", 3, 32, 4, 64, 64),      # thin-Cin layer at a size that spans many blocks
    (1, 1, "same", 64, 3, 4, 64, 64),      # thin-Cout head
    (1, 1, "same", 64, 1, 4, 8, 8),        # logits head
    (3, 1, "same", 32, 3, 2, 24, 40),      # Fast-SRGAN 3x3 head
    (3, 1, "same", 32, 64, 4, 16, 16),     # discriminator conv5 shape (Cin < tile width)
    (3, 2, "same", 64, 64, 4, 16, 16),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_simt_conv_fwd_dgrad_wgrad_fp32(L, case):
    k, s, padding, cin, cout, N, H, W = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    x = torch.randn(N, H, W, cin, generator=g, dtype=torch.float64)
    w = torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * 0.2
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    y_ref = OT.conv2d(xr, wr, br, stride=s, padding=padding)
    gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    (y_ref * gy).sum().backward()

    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, s, H, W, padding)
    xd, wd, bd, gyd = dev(x), dev(w), dev(b), dev(gy)
    y = torch.empty(y_ref.shape, device="cuda")
    tx, ty = L.tensor(xd), L.tensor(y)
    L.check(lib.dg_conv2d_fwd(ctx, C.byref(tx), wd.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp), st))
    assert relerr(y, y_ref) < FP32_TOL
    dx = torch.empty_like(xd)
    tgy, tdx = L.tensor(gyd), L.tensor(dx)
    L.check(lib.dg_conv2d_dgrad(ctx, C.byref(tgy), wd.data_ptr(), None, C.byref(tdx), C.byref(cp), st))
    assert relerr(dx, xr.grad) < FP32_TOL
    dw = torch.empty_like(wd); db = torch.empty_like(bd)
    nb = lib.dg_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(tgy), C.byref(cp))
    wk = ws(nb)
    L.check(lib.dg_conv2d_wgrad(ctx, C.byref(tx), C.byref(tgy), dw.data_ptr(), db.data_ptr(), C.byref(cp), 0, wk.data_ptr(), nb, st))
    assert relerr(dw, wr.grad) < FP32_TOL
    assert relerr(db, br.grad) < FP32_TOL
    # accumulate flag
    L.check(lib.dg_conv2d_wgrad(ctx, C.byref(tx), C.byref(tgy), dw.data_ptr(), db.data_ptr(), C.byref(cp), 1, wk.data_ptr(), nb, st))
    assert relerr(dw, 2 * wr.grad) < FP32_TOL


def test_simt_conv_epilogue_activations_and_views(L):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 6, 6, 8, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, 8, 4, generator=g, dtype=torch.float64) * 0.3
    b = torch.randn(4, generator=g, dtype=torch.float64)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    ref = OT.conv2d(x, w, b)
    for act, f in [(1, torch.relu), (2, lambda t: OT.leaky_relu(t, 0.2)), (3, torch.tanh), (4, torch.sigmoid)]:
        cp = conv_params(L, 3, 3, 1, 6, 6, "same", act, 0.2)
        # write into channels [4, 8) of a 12-channel buffer (concat slice), read x from a slice too
        xin = torch.zeros(2, 6, 6, 10, device="cuda"); xin[..., 2:] = dev(x)
        out = torch.full((2, 6, 6, 12), 7.0, device="cuda")
        tx, ty = L.tensor(xin, c=8, coff=2), L.tensor(out, c=4, coff=4)
        wd, bd = dev(w), dev(b)
        L.check(lib.dg_conv2d_fwd(ctx, C.byref(tx), wd.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp), st))
        assert relerr(out[..., 4:8], f(ref)) < FP32_TOL
        assert (out[..., :4] == 7).all() and (out[..., 8:] == 7).all()


@pytest.mark.parametrize("k,s,cin,cout", [(4, 2, 8, 6), (3, 2, 5, 4), (4, 2, 16, 3)])
def test_simt_conv_transpose_forward(L, k, s, cin, cout):
    g = torch.Generator().manual_seed(k * 10 + s)
    x = torch.randn(2, 5, 4, cin, generator=g, dtype=torch.float64)
    w = torch.randn(k, k, cout, cin, generator=g, dtype=torch.float64) * 0.3
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    ref = torch.tanh(OT.conv2d_transpose(x, w, b, stride=s))
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    Ho, Wo = 5 * s, 4 * s
    cp = conv_params(L, k, k, s, Ho, Wo, "same", 3, 0.0)
    y = torch.empty(2, Ho, Wo, cout, device="cuda")
    xd, wd, bd = dev(x), dev(w), dev(b)
    tx, ty = L.tensor(xd), L.tensor(y)
    L.check(lib.dg_conv2d_dgrad(ctx, C.byref(tx), wd.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp), st))
    assert relerr(y, ref) < FP32_TOL


UMMA_CASES = [  # (k, stride, padding, cin, cout, N, H, W)
    (3, 1, "same", 64, 64, 2, 24, 20),
    (3, 1, "same", 64, 64, 1, 48, 32),
    (3, 1, "same", 32, 32, 2, 24, 24),
    (3, 1, "same", 32, 64, 1, 16, 8),
    (3, 1, "same", 64, 256, 1, 32, 16),
    (3, 1, "same", 256, 64, 1, 32, 16),
    (3, 1, "same", 16, 48, 1, 20, 12),
    (3, 1, "same", 192, 32, 1, 16, 16),
    (1, 1, "same", 32, 192, 1, 16, 16),
    (3, 2, "same", 32, 32, 2, 24, 24),
    (3, 2, "same", 64, 64, 1, 48, 16),
    (4, 2, "same", 64, 128, 1, 32, 32),
    (4, 1, ((1, 1), (1, 1)), 64, 32, 1, 18, 18),
    (3, 1, "same", 256, 256, 1, 40, 24),
    (3, 1, "same", 128, 512, 1, 16, 16),
]


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float64)


@pytest.mark.parametrize("case", UMMA_CASES)
def test_umma_conv_fwd_dgrad_bf16(L, case):
    k, s, padding, cin, cout, N, H, W = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    x = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64))
    w = _bf16_round(torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * (1.0 / (k * np.sqrt(cin))))
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    y_ref = OT.conv2d(xr, w, b, stride=s, padding=padding)
    gy = _bf16_round(torch.randn(y_ref.shape, generator=g, dtype=torch.float64))
    (y_ref * gy).sum().backward()

    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    assert lib.dg_has_umma(ctx) == 1
    cp = conv_params(L, k, k, s, H, W, padding)
    xd, wd, bd, gyd = dev(x, torch.bfloat16), dev(w), dev(b), dev(gy, torch.bfloat16)
    pk_f = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    pk_d = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk_f.data_ptr(), k, k, cin, cout, 0, st))
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk_d.data_ptr(), k, k, cin, cout, 1, st))
    for out_dtype in (torch.bfloat16, torch.float32):
        y = torch.full(y_ref.shape, float("nan"), device="cuda", dtype=out_dtype)
        tx, ty = L.tensor(xd), L.tensor(y)
        L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk_f.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp), None, st))
        torch.cuda.synchronize()
        e = relerr(y, y_ref)
        assert e < (BF16_TOL if out_dtype == torch.bfloat16 else 1e-4), f"fwd relerr {e}"
    dx = torch.full(x.shape, float("nan"), device="cuda", dtype=torch.bfloat16)
    tgy, tdx = L.tensor(gyd), L.tensor(dx)
    L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(tgy), pk_d.data_ptr(), None, C.byref(tdx), C.byref(cp), st))
    torch.cuda.synchronize()
    e = relerr(dx, xr.grad)
    assert e < BF16_TOL, f"dgrad relerr {e}"


def test_umma_conv_transpose_forward_bf16(L):
    g = torch.Generator().manual_seed(3)
    cin, cout, k, s = 128, 64, 4, 2
    x = _bf16_round(torch.randn(2, 16, 8, cin, generator=g, dtype=torch.float64))
    w = _bf16_round(torch.randn(k, k, cout, cin, generator=g, dtype=torch.float64) * 0.02)
    ref = OT.conv2d_transpose(x, w, None, stride=s)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    # as the forward conv f the kernel is HWIO with I = cout, O = cin; Conv2DTranspose forward = dgrad of f
    cp = conv_params(L, k, k, s, 32, 16, "same")
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    wd, xd = dev(w), dev(x, torch.bfloat16)
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cout, cin, 1, st))
    y = torch.full(ref.shape, float("nan"), device="cuda", dtype=torch.bfloat16)
    tx, ty = L.tensor(xd), L.tensor(y)
    L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(tx), pk.data_ptr(), None, C.byref(ty), C.byref(cp), st))
    torch.cuda.synchronize()
    assert relerr(y, ref) < BF16_TOL


def test_umma_rejects_bad_shapes(L):
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    x = torch.zeros(1, 8, 8, 3, device="cuda", dtype=torch.bfloat16)
    y = torch.zeros(1, 8, 8, 16, device="cuda", dtype=torch.bfloat16)
    cp = conv_params(L, 3, 3, 1, 8, 8, "same")
    tx, ty = L.tensor(x), L.tensor(y)
    rc = lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), y.data_ptr(), None, C.byref(ty), C.byref(cp), None, st)
    assert rc != 0 and b"multiples of 16" in lib.dg_last_error()


@pytest.mark.parametrize("C_,act,with_res,dropout", [(64, "relu", False, False), (100, "prelu", False, False),
                                                     (32, "lrelu", False, False), (64, None, True, False),
                                                     (512, "relu", False, True), (48, "tanh", False, False)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,H,W", [(2, 6, 5), (4, 16, 16), (3, 40, 33)])
def test_bn_act_fwd_bwd(L, C_, act, with_res, dropout, dtype, N, H, W):
    g = torch.Generator().manual_seed(C_)
    rnd = (lambda t: t) if dtype == torch.float32 else _bf16_round
    x = rnd(torch.randn(N, H, W, C_, generator=g, dtype=torch.float64) * 1.5 + 0.3)
    gamma = torch.randn(C_, generator=g, dtype=torch.float64) * 0.2 + 1
    beta = torch.randn(C_, generator=g, dtype=torch.float64) * 0.2
    alpha = torch.randn(C_, generator=g, dtype=torch.float64) * 0.3
    res = rnd(torch.randn(N, H, W, C_, generator=g, dtype=torch.float64)) if with_res else None
    gy = rnd(torch.randn(N, H, W, C_, generator=g, dtype=torch.float64))
    xr, gr, br, ar = [t.clone().requires_grad_(True) for t in (x, gamma, beta, alpha)]
    p = {"bn/gamma": gr, "bn/beta": br, "bn/moving_mean": torch.zeros(C_, dtype=torch.float64), "bn/moving_variance": torch.ones(C_, dtype=torch.float64)}
    stt = {}
    t = OT.batch_norm(xr, p, "bn", True, stt, momentum=0.8, eps=1e-3)
    seed, off = 7, 11
    if dropout:
        keep = torch.from_numpy(ON.dropout_keep_mask(seed, off, x.numel())).view(x.shape).double()
        t = t * keep * 2.0
    a_code = {None: 0, "relu": 1, "lrelu": 2, "tanh": 3, "prelu": 5}[act]
    if act == "relu": t = torch.relu(t)
    elif act == "lrelu": t = OT.leaky_relu(t, 0.2)
    elif act == "tanh": t = torch.tanh(t)
    elif act == "prelu": t = OT.prelu(t, ar)
    if with_res: t = t + res
    (t * gy).sum().backward()

    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    xd, gyd = dev(x, dtype), dev(gy, dtype)
    gd, bd, ad = dev(gamma), dev(beta), dev(alpha)
    mm, mv = torch.zeros(C_, device="cuda"), torch.ones(C_, device="cuda")
    scale, shift, mean, invstd = [torch.empty(C_, device="cuda") for _ in range(4)]
    tx = L.tensor(xd)
    nb = lib.dg_bn_workspace_bytes(C.byref(tx)); wk = ws(nb)
    L.check(lib.dg_bn_stats(ctx, C.byref(tx), gd.data_ptr(), bd.data_ptr(), 1e-3, 0.8, mm.data_ptr(), mv.data_ptr(), scale.data_ptr(),
                            shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), wk.data_ptr(), nb, st))
    y = torch.empty_like(xd)
    ty = L.tensor(y)
    resd = dev(res, dtype) if with_res else None
    tres = L.tensor(resd) if with_res else None
    L.check(lib.dg_bn_act_fwd(ctx, C.byref(tx), scale.data_ptr(), shift.data_ptr(), a_code, 0.2, ad.data_ptr() if act == "prelu" else None,
                              C.byref(tres) if with_res else None, 1 if dropout else 0, seed, off, None, C.byref(ty), st))
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert relerr(y, t) < tol
    assert relerr(mm, stt["bn/moving_mean"]) < 1e-5 and relerr(mv, stt["bn/moving_variance"]) < 1e-5
    dx = torch.empty_like(xd); dgam, dbet, dalp = [torch.empty(C_, device="cuda") for _ in range(3)]
    tgy, tdx = L.tensor(gyd), L.tensor(dx)
    L.check(lib.dg_bn_act_bwd(ctx, C.byref(tgy), C.byref(tx), scale.data_ptr(), shift.data_ptr(), gd.data_ptr(), mean.data_ptr(),
                              invstd.data_ptr(), a_code, 0.2, ad.data_ptr() if act == "prelu" else None, 1 if dropout else 0, seed, off, None,
                              C.byref(tdx), dgam.data_ptr(), dbet.data_ptr(), dalp.data_ptr(), 0, wk.data_ptr(), nb, st))
    gtol = 2e-5 if dtype == torch.float32 else BF16_TOL
    assert relerr(dx, xr.grad) < gtol
    assert relerr(dgam, gr.grad) < gtol and relerr(dbet, br.grad) < gtol
    if act == "prelu":
        assert relerr(dalp, ar.grad) < gtol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_d2s_prelu_fwd_bwd(L, dtype):
    g = torch.Generator().manual_seed(2)
    rnd = (lambda t: t) if dtype == torch.float32 else _bf16_round
    u = rnd(torch.randn(2, 5, 4, 32, generator=g, dtype=torch.float64))
    alpha = torch.randn(8, generator=g, dtype=torch.float64) * 0.4
    gy = rnd(torch.randn(2, 10, 8, 8, generator=g, dtype=torch.float64))
    ur, ar = u.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
    ref = OT.prelu(OT.depth_to_space(ur, 2), ar)
    (ref * gy).sum().backward()
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    ud, gyd, ad = dev(u, dtype), dev(gy, dtype), dev(alpha)
    y = torch.empty(2, 10, 8, 8, device="cuda", dtype=dtype)
    tu, ty = L.tensor(ud), L.tensor(y)
    L.check(lib.dg_d2s_prelu_fwd(ctx, C.byref(tu), ad.data_ptr(), C.byref(ty), st))
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert relerr(y, ref) < tol
    du = torch.empty_like(ud); da = torch.empty(8, device="cuda")
    tgy, tdu = L.tensor(gyd), L.tensor(du)
    nb = lib.dg_bn_workspace_bytes(C.byref(tgy)); wk = ws(nb)
    L.check(lib.dg_d2s_prelu_bwd(ctx, C.byref(tgy), C.byref(tu), ad.data_ptr(), C.byref(tdu), da.data_ptr(), 0, wk.data_ptr(), nb, st))
    assert relerr(du, ur.grad) < tol and relerr(da, ar.grad) < (2e-5 if dtype == torch.float32 else BF16_TOL)


def test_structural_ops(L):
    g = torch.Generator().manual_seed(4)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    x = torch.randn(2, 6, 8, 5, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    # max pool
    ref = OT.max_pool2x2(xr); gy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * gy).sum().backward()
    xd, gyd = dev(x), dev(gy)
    y = torch.empty(ref.shape, device="cuda"); dx = torch.empty_like(xd)
    tx, ty, tg, tdx = L.tensor(xd), L.tensor(y), L.tensor(gyd), L.tensor(dx)
    L.check(lib.dg_maxpool2x2_fwd(ctx, C.byref(tx), C.byref(ty), st))
    L.check(lib.dg_maxpool2x2_bwd(ctx, C.byref(tg), C.byref(tx), C.byref(ty), C.byref(tdx), st))
    assert relerr(y, ref) < 1e-6 and relerr(dx, xr.grad) < 1e-6
    # upsample + relu into a concat slice
    xr.grad = None
    ref = torch.relu(OT.upsample2x_nearest(xr)); gy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * gy).sum().backward()
    wide = torch.zeros(2, 12, 16, 9, device="cuda")
    tw = L.tensor(wide, c=5, coff=0)
    L.check(lib.dg_upsample2x_relu_fwd(ctx, C.byref(tx), C.byref(tw), st))
    assert relerr(wide[..., :5], ref) < 1e-6 and (wide[..., 5:] == 0).all()
    gyd = dev(gy); tg = L.tensor(gyd)
    L.check(lib.dg_upsample2x_relu_bwd(ctx, C.byref(tg), C.byref(tx), C.byref(tdx), st))
    assert relerr(dx, xr.grad) < 1e-6
    # add, copy with accumulate + dtype conversion
    a, b = dev(torch.randn(2, 3, 3, 4, generator=g)), dev(torch.randn(2, 3, 3, 4, generator=g))
    o = torch.empty_like(a)
    ta, tb, to = L.tensor(a), L.tensor(b), L.tensor(o)
    L.check(lib.dg_add(ctx, C.byref(ta), C.byref(tb), C.byref(to), st))
    assert torch.equal(o, a + b)
    ob = torch.ones(2, 3, 3, 4, device="cuda", dtype=torch.bfloat16)
    tob = L.tensor(ob)
    L.check(lib.dg_copy(ctx, C.byref(ta), C.byref(tob), 1, st))
    assert relerr(ob, a + 1) < 1e-2
    # vgg preprocess
    img = dev(torch.rand(2, 4, 4, 3, generator=g) * 2 - 1)
    pre = torch.empty_like(img); tp, ti = L.tensor(pre), L.tensor(img)
    L.check(lib.dg_vgg_preprocess_fwd(ctx, C.byref(ti), C.byref(tp), st))
    from oracle.models import vgg_preprocess
    assert relerr(pre, vgg_preprocess(img.cpu().double())) < 1e-6


@pytest.mark.parametrize("C_", [32, 192])
def test_depthwise(L, C_):
    g = torch.Generator().manual_seed(C_)
    x = torch.randn(2, 7, 6, C_, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, C_, 1, generator=g, dtype=torch.float64); b = torch.randn(C_, generator=g, dtype=torch.float64)
    xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
    ref = OT.depthwise_conv2d(xr, wr, br); gy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * gy).sum().backward()
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    xd, wd, bd, gyd = dev(x), dev(w), dev(b), dev(gy)
    y = torch.empty_like(xd); dx = torch.empty_like(xd); dw = torch.empty_like(wd); db = torch.empty_like(bd)
    tx, ty, tg, tdx = L.tensor(xd), L.tensor(y), L.tensor(gyd), L.tensor(dx)
    L.check(lib.dg_dwconv3x3_fwd(ctx, C.byref(tx), wd.data_ptr(), bd.data_ptr(), C.byref(ty), st))
    L.check(lib.dg_dwconv3x3_dgrad(ctx, C.byref(tg), wd.data_ptr(), C.byref(tdx), st))
    nb = lib.dg_dwconv3x3_wgrad_workspace_bytes(C.byref(tx)); wk = ws(nb)
    L.check(lib.dg_dwconv3x3_wgrad(ctx, C.byref(tx), C.byref(tg), dw.data_ptr(), db.data_ptr(), 0, wk.data_ptr(), nb, st))
    assert relerr(y, ref) < FP32_TOL and relerr(dx, xr.grad) < FP32_TOL
    assert relerr(dw, wr.grad) < FP32_TOL and relerr(db, br.grad) < FP32_TOL


def test_losses_and_grads(L):
    g = torch.Generator().manual_seed(9)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    gen = (torch.rand(2, 6, 5, 3, generator=g, dtype=torch.float64) * 2 - 1)
    tgt = (torch.rand(2, 6, 5, 3, generator=g, dtype=torch.float64) * 2 - 1).float().double()
    gen = gen.float().double()
    gr = gen.clone().requires_grad_(True)
    w_mae, w_mse, w_tv = 1.0, 0.5, 1e-5
    loss = w_mae * OT.mae(tgt, gr) + w_mse * OT.mse(tgt, gr) + w_tv * OT.total_variation_mean(tgt - gr)
    loss.backward()
    out3 = torch.empty(3, device="cuda"); dgen = torch.empty(2, 6, 5, 3, device="cuda")
    gend, tgtd = dev(gen), dev(tgt)
    tg, tt, td = L.tensor(gend), L.tensor(tgtd), L.tensor(dgen)
    nb = lib.dg_loss_workspace_bytes(C.byref(tg)); wk = ws(nb)
    L.check(lib.dg_image_losses(ctx, C.byref(tg), C.byref(tt), w_mae, w_mse, w_tv, out3.data_ptr(), C.byref(td), 0, wk.data_ptr(), nb, st))
    ref3 = torch.tensor([OT.mae(tgt, gen), OT.mse(tgt, gen), OT.total_variation_mean(tgt - gen)])
    assert relerr(out3, ref3) < FP32_TOL
    assert relerr(dgen, gr.grad) < FP32_TOL
    # BCE, both forms
    x = torch.randn(2, 3, 3, 1, generator=g, dtype=torch.float64) * 2
    for from_logits in (1, 0):
        for z in (0.0, 1.0):
            xin = x if from_logits else torch.sigmoid(x)
            xin = xin.float().double()
            xr = xin.clone().requires_grad_(True)
            l = OT.bce_from_logits(xr, z) if from_logits else OT.bce_from_probs(xr, z)
            (l * 1e-3).backward()
            lo = torch.empty(1, device="cuda"); dx = torch.empty(2, 3, 3, 1, device="cuda")
            xind = dev(xin)
            tx, tdx = L.tensor(xind), L.tensor(dx)
            L.check(lib.dg_bce_const_target(ctx, C.byref(tx), z, from_logits, 1e-3, lo.data_ptr(), C.byref(tdx), wk.data_ptr(), nb, st))
            assert relerr(lo, l.detach().view(1)) < 2e-5
            assert relerr(dx, xr.grad) < 2e-5
    # feature MSE
    a = torch.randn(2, 3, 3, 16, generator=g, dtype=torch.float64).float().double(); b = torch.randn(2, 3, 3, 16, generator=g, dtype=torch.float64).float().double()
    ar = a.clone().requires_grad_(True)
    l = OT.mse(b / 12.75, ar / 12.75); l.backward()
    lo = torch.empty(1, device="cuda"); da = torch.empty(2, 3, 3, 16, device="cuda")
    a_d, b_d = dev(a), dev(b)
    ta, tb, tda = L.tensor(a_d), L.tensor(b_d), L.tensor(da)
    L.check(lib.dg_feature_mse(ctx, C.byref(ta), C.byref(tb), 1.0 / 12.75, lo.data_ptr(), C.byref(tda), wk.data_ptr(), nb, st))
    assert relerr(lo, l.detach().view(1)) < FP32_TOL and relerr(da, ar.grad) < FP32_TOL


def test_adam_matches_keras_definition(L):
    g = torch.Generator().manual_seed(1)
    n = 1000 + 3
    th = torch.randn(n, generator=g, dtype=torch.float64).float().double()
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    npad = 1024
    theta = torch.zeros(npad, device="cuda"); theta[:n] = dev(th)
    m = torch.zeros(npad, device="cuda"); v = torch.zeros(npad, device="cuda")
    state = torch.zeros(2, dtype=torch.int64, device="cuda")
    opt = OT.KerasAdam(1e-3, beta1=0.5, decay_steps=2, decay_rate=0.1)
    p = {"w": th.clone()}
    for step in range(4):
        gr = torch.randn(n, generator=g, dtype=torch.float64).float().double()
        gd = torch.zeros(npad, device="cuda"); gd[:n] = dev(gr) * 2.0   # grad_scale 0.5 undoes the x2
        L.check(lib.dg_adam_step(ctx, theta.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), npad, 1e-3, 0.5, 0.999, 1e-7, 2, 0.1,
                                 0.5, state.data_ptr(), st))
        opt.apply(p, {"w": gr})
        assert relerr(theta[:n], p["w"]) < 1e-6, f"step {step}"
    assert state[0].item() == 4


WGRAD_CASES = [  # (k, stride, padding, cin, cout, N, H, W, bias)
    (3, 1, "same", 64, 64, 2, 24, 20, False),
    (3, 1, "same", 64, 64, 1, 48, 32, True),
    (3, 1, "same", 32, 32, 2, 24, 24, True),
    (3, 1, "same", 32, 64, 1, 16, 8, True),
    (3, 1, "same", 64, 256, 1, 32, 16, True),
    (3, 1, "same", 256, 64, 1, 32, 16, False),
    (3, 1, "same", 16, 48, 1, 20, 12, True),
    (3, 1, "same", 192, 32, 1, 16, 16, True),
    (1, 1, "same", 32, 192, 1, 16, 16, True),
    (3, 2, "same", 32, 32, 2, 24, 24, True),
    (3, 2, "same", 64, 64, 1, 48, 16, True),
    (4, 2, "same", 64, 128, 1, 32, 32, False),
    (4, 1, ((1, 1), (1, 1)), 64, 32, 1, 18, 18, False),
    (3, 1, "same", 128, 128, 2, 40, 24, True),
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_umma_conv_wgrad_bf16(L, case):
    k, s, padding, cin, cout, N, H, W, bias = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    x = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64))
    w = torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64).requires_grad_(True)
    b = torch.randn(cout, generator=g, dtype=torch.float64).requires_grad_(True)
    y_ref = OT.conv2d(x, w, b, stride=s, padding=padding)
    gy = _bf16_round(torch.randn(y_ref.shape, generator=g, dtype=torch.float64))
    (y_ref * gy).sum().backward()
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, s, H, W, padding)
    xd, gyd = dev(x, torch.bfloat16), dev(gy, torch.bfloat16)
    tx, tg = L.tensor(xd), L.tensor(gyd)
    nb = lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(tg), C.byref(cp))
    assert nb > 0
    wk = ws(nb)
    dw = torch.full(w.shape, float("nan"), device="cuda"); db = torch.full((cout,), float("nan"), device="cuda")
    L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tx), C.byref(tg), dw.data_ptr(), db.data_ptr() if bias else None, C.byref(cp), 0,
                                     wk.data_ptr(), nb, st))
    torch.cuda.synchronize()
    e = relerr(dw, w.grad)
    assert e < 1e-4, f"dW relerr {e}"        # bf16 inputs are exact here; only fp32 accumulation order differs
    if bias:
        assert relerr(db, b.grad) < 1e-4
    L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tx), C.byref(tg), dw.data_ptr(), db.data_ptr() if bias else None, C.byref(cp), 1,
                                     wk.data_ptr(), nb, st))
    torch.cuda.synchronize()
    assert relerr(dw, 2 * w.grad) < 1e-4


@pytest.mark.parametrize("cin,cout,k", [(3, 32, 3), (64, 3, 1), (64, 1, 1)])
def test_thin_layers_mixed_dtypes(L, cin, cout, k):
    """bf16 model: the RGB / logits side of a thin layer is fp32, the wide side bf16 (srgan.py:183,270)."""
    g = torch.Generator().manual_seed(cin * 7 + cout)
    N, H, W = 3, 40, 24
    x_dt = torch.float32 if cin <= 4 else torch.bfloat16
    y_dt = torch.float32 if cout <= 4 else torch.bfloat16
    rx = (lambda t: t.float().double()) if x_dt == torch.float32 else _bf16_round
    ry = (lambda t: t.float().double()) if y_dt == torch.float32 else _bf16_round
    x = rx(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64))
    w = (torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * 0.2).float().double()
    b = torch.randn(cout, generator=g, dtype=torch.float64).float().double()
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = OT.conv2d(xr, wr, br)
    gy = ry(torch.randn(y_ref.shape, generator=g, dtype=torch.float64))
    (y_ref * gy).sum().backward()
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, 1, H, W, "same")
    xd, wd, bd, gyd = dev(x, x_dt), dev(w), dev(b), dev(gy, y_dt)
    y = torch.empty(y_ref.shape, device="cuda", dtype=y_dt)
    tx, ty, tg = L.tensor(xd), L.tensor(y), L.tensor(gyd)
    L.check(lib.dg_conv2d_fwd(ctx, C.byref(tx), wd.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp), st))
    assert relerr(y, y_ref) < (1e-5 if y_dt == torch.float32 else 1e-2)
    dx = torch.empty_like(xd); tdx = L.tensor(dx)
    L.check(lib.dg_conv2d_dgrad(ctx, C.byref(tg), wd.data_ptr(), None, C.byref(tdx), C.byref(cp), st))
    assert relerr(dx, xr.grad) < (1e-5 if x_dt == torch.float32 else 1e-2)
    dw = torch.empty_like(wd); db = torch.empty_like(bd)
    nb = lib.dg_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(tg), C.byref(cp)); wk = ws(nb)
    L.check(lib.dg_conv2d_wgrad(ctx, C.byref(tx), C.byref(tg), dw.data_ptr(), db.data_ptr(), C.byref(cp), 0, wk.data_ptr(), nb, st))
    assert relerr(dw, wr.grad) < 1e-5 and relerr(db, br.grad) < 1e-5


@pytest.mark.parametrize("cin,cout,k,stride,act,out_f32", [(3, 32, 3, 1, "lrelu", False), (3, 64, 3, 1, None, False),
                                                          (64, 3, 1, 1, "tanh", True), (3, 64, 4, 2, "lrelu", False)])
def test_rgb_sided_conv_on_tensor_cores(cin, cout, k, stride, act, out_f32):
    """3-channel-sided convs (srgan.py:154,182,236; pix2pix.py:147) through the zero-padded 16-channel tcgen05 path:
    forward, input gradient, kernel and bias gradients against a float64 conv on the same bf16-rounded operands."""
    from collections import OrderedDict
    from denoise_gan_b200 import _lib
    from denoise_gan_b200.engine import Engine
    from denoise_gan_b200.params import ParamSet
    from oracle import ops_torch as OT
    g = torch.Generator().manual_seed(cin * 100 + cout)
    N, H, W = 2, 40, 24
    w0 = torch.randn((k, k, cin, cout), generator=g) * 0.2
    b0 = torch.randn((cout,), generator=g) * 0.1
    x0 = torch.randn((N, H, W, cin), generator=g)
    E = Engine(bf16=True)
    ps = ParamSet("g", OrderedDict([("k", w0), ("b", b0)]), E.device)
    xin = x0.cuda() if cin == 3 else x0.cuda().to(torch.bfloat16)
    from denoise_gan_b200.engine import Var
    xv = Var(xin, frozenset({"g"}), E._next())      # depends on group "g": ask for the input gradient as well
    y = E.conv2d(xv, ps["k"], ps["b"], stride=stride, act=act, alpha=0.2, out_dtype=torch.float32 if out_f32 else None)
    assert E._cap and any(k_[0] == "padded" and v for k_, v in E._cap.items()), "padded tensor-core path was not taken"
    gy0 = torch.randn(y.shape, generator=g)
    gy = gy0.cuda().to(y.t.dtype)
    coll = {}
    E.backward([(y, gy)], "g", collect=coll)
    torch.cuda.synchronize()
    q = lambda t: t.to(torch.bfloat16).double()
    xq, wq = q(x0), q(w0)
    xr = xq.clone().requires_grad_(True); wr = wq.clone().requires_grad_(True); br = b0.double().clone().requires_grad_(True)
    pre = OT.conv2d(xr, wr, br, stride=stride)
    ref = {"lrelu": lambda t: torch.nn.functional.leaky_relu(t, 0.2), "tanh": torch.tanh, None: lambda t: t}[act](pre)
    ref.backward(gy.double().cpu())
    tol = 2e-2
    assert relerr(y.t.float().cpu(), ref.detach().float()) < tol
    dx = E.pool[[kk for kk in E.pool if isinstance(kk[0], tuple) and len(kk[0]) >= 2 and kk[0][0] == y.seq and kk[0][1] == "dx"][0]]
    assert relerr(dx.float().cpu(), xr.grad.float()) < tol
    assert relerr(ps["k"].grad.cpu(), wr.grad.float()) < tol
    assert relerr(ps["b"].grad.cpu(), br.grad.float()) < tol


def test_umma_wgrad_odd_multiples_of_16():
    """autoencoder.py conv7: 208 -> 112 channels (13 and 7 chunks of 16): N block 112, padding atoms, partial dump."""
    from denoise_gan_b200 import _lib as L
    lib, ctx, st = L.load(), L.ctx(), L.stream_ptr()
    g = torch.Generator().manual_seed(5)
    N, H, W, cin, cout, k = 12, 32, 32, 208, 112, 3       # several tiles per CTA: the second pipeline stage is used too
    x0 = torch.randn((N, H, W, cin), generator=g).to(torch.bfloat16)
    dy0 = torch.randn((N, H, W, cout), generator=g).to(torch.bfloat16)
    xd, dyd = x0.cuda(), dy0.cuda()
    cp = L.DgConvParams(k, k, 1, 1, 1, 0, 0.0)
    tx, tdy = L.tensor(xd), L.tensor(dyd)
    nbytes = lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(tdy), C.byref(cp))
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dw = torch.empty((k, k, cin, cout), device="cuda"); db = torch.empty(cout, device="cuda")
    guard = torch.full((1 << 16,), 7.0, device="cuda")      # canary right after the outputs in allocation order
    L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tx), C.byref(tdy), dw.data_ptr(), db.data_ptr(), C.byref(cp), 0, ws.data_ptr(), nbytes, st))
    torch.cuda.synchronize()
    w = torch.zeros((k, k, cin, cout), dtype=torch.float32, requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    OT.conv2d(x0.float(), w, b).backward(dy0.float())
    assert relerr(dw.cpu(), w.grad) < 2e-2 and relerr(db.cpu(), b.grad) < 2e-2
    assert bool((guard == 7.0).all())


@pytest.mark.parametrize("N,H,W,cin,cout,k,kind", [(8, 96, 96, 64, 64, 3, "fwd"), (4, 96, 96, 256, 64, 3, "fwd"),
                                                   (8, 96, 96, 32, 32, 3, "dgrad"), (6, 64, 64, 128, 128, 3, "fwd")])
def test_umma_conv_many_tiles_per_cta(N, H, W, cin, cout, k, kind):
    """Several tiles per persistent CTA: exercises the two alternating MMA-issuing warps, their per-issuer pipeline
    rings (multi-chunk K with streamed weights included) and the four TMEM accumulator buffers."""
    from denoise_gan_b200 import _lib as L
    lib, ctx, st = L.load(), L.ctx(), L.stream_ptr()
    g = torch.Generator().manual_seed(N * 1000 + cin)
    x0 = torch.randn((N, H, W, cin), generator=g).to(torch.bfloat16)
    w0 = (torch.randn((k, k, cin, cout), generator=g) * 0.05)
    xd = x0.cuda()
    wd = w0.cuda()
    pk = torch.empty(w0.numel(), dtype=torch.bfloat16, device="cuda")
    cp = L.DgConvParams(k, k, 1, k // 2, k // 2, 0, 0.0)
    wq = w0.to(torch.bfloat16).float()
    if kind == "fwd":
        L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cin, cout, 0, st))
        y = torch.empty((N, H, W, cout), dtype=torch.bfloat16, device="cuda")
        tx, ty = L.tensor(xd), L.tensor(y)
        L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk.data_ptr(), None, C.byref(ty), C.byref(cp), None, st))
        ref = OT.conv2d(x0.float(), wq)
        out = y
    else:   # dgrad: x0 plays dy [N,H,W,cin->"cout" of the conv]; conv is cout_conv=cin, cin_conv=cout
        wconv = (torch.randn((k, k, cout, cin), generator=g) * 0.05)
        pk = torch.empty(wconv.numel(), dtype=torch.bfloat16, device="cuda")
        wcd = wconv.cuda()
        L.check(lib.dg_umma_pack_weights(ctx, wcd.data_ptr(), pk.data_ptr(), k, k, cout, cin, 1, st))
        dx = torch.empty((N, H, W, cout), dtype=torch.bfloat16, device="cuda")
        tdy, tdx = L.tensor(xd), L.tensor(dx)
        L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(tdy), pk.data_ptr(), None, C.byref(tdx), C.byref(cp), st))
        xin = torch.zeros((N, H, W, cout), dtype=torch.float32, requires_grad=True)
        OT.conv2d(xin, wconv.to(torch.bfloat16).float()).backward(x0.float())
        ref = xin.grad
        out = dx
    torch.cuda.synchronize()
    assert relerr(out.float().cpu(), ref) < 2e-2


@pytest.mark.parametrize("N,H,cin,cout", [(4, 64, 128, 256), (2, 32, 256, 512), (8, 16, 512, 512), (3, 2, 512, 512)])
def test_umma_conv_k4s2_split_source_stages(N, H, cin, cout):
    """pix2pix.py:110-123 downsample convs (4x4, stride 2, SAME) with >= 128 channels: the forward kernel stages one
    parity source (halo + its four taps' weights) per pipeline slot."""
    from denoise_gan_b200 import _lib as L
    lib, ctx, st = L.load(), L.ctx(), L.stream_ptr()
    g = torch.Generator().manual_seed(H * 7 + cin)
    x0 = torch.randn((N, H, H, cin), generator=g).to(torch.bfloat16)
    w0 = torch.randn((4, 4, cin, cout), generator=g) * 0.03
    xd, wd = x0.cuda(), w0.cuda()
    pk = torch.empty(w0.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), 4, 4, cin, cout, 0, st))
    cp = L.DgConvParams(4, 4, 2, 1, 1, 0, 0.0)
    y = torch.empty((N, H // 2, H // 2, cout), dtype=torch.bfloat16, device="cuda")
    tx, ty = L.tensor(xd), L.tensor(y)
    assert lib.dg_umma_conv2d_fwd_supported(ctx, C.byref(tx), C.byref(ty), C.byref(cp))
    L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk.data_ptr(), None, C.byref(ty), C.byref(cp), None, st))
    torch.cuda.synchronize()
    ref = OT.conv2d(x0.float(), w0.to(torch.bfloat16).float(), stride=2)
    assert relerr(y.float().cpu(), ref) < 2e-2


@pytest.mark.parametrize("N,H,cin,cout", [(4, 64, 128, 256), (2, 32, 256, 512), (8, 16, 512, 512), (3, 4, 512, 512)])
def test_umma_wgrad_k4s2_one_launch_per_source(N, H, cin, cout):
    """Weight gradient of the pix2pix 4x4 stride-2 convs with >= 128 channels: the four parity sources are processed by
    separate launches (their halo boxes do not fit one pipeline stage together)."""
    from denoise_gan_b200 import _lib as L
    lib, ctx, st = L.load(), L.ctx(), L.stream_ptr()
    g = torch.Generator().manual_seed(H * 11 + cin)
    x0 = torch.randn((N, H, H, cin), generator=g).to(torch.bfloat16)
    dy0 = torch.randn((N, H // 2, H // 2, cout), generator=g).to(torch.bfloat16)
    xd, dyd = x0.cuda(), dy0.cuda()
    cp = L.DgConvParams(4, 4, 2, 1, 1, 0, 0.0)
    tx, tdy = L.tensor(xd), L.tensor(dyd)
    nbytes = lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(tdy), C.byref(cp))
    assert nbytes > 0, lib.dg_last_error().decode()
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dw = torch.empty((4, 4, cin, cout), device="cuda"); db = torch.empty(cout, device="cuda")
    L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tx), C.byref(tdy), dw.data_ptr(), db.data_ptr(), C.byref(cp), 0, ws.data_ptr(), nbytes, st))
    torch.cuda.synchronize()
    w = torch.zeros((4, 4, cin, cout), dtype=torch.float32, requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    OT.conv2d(x0.float(), w, b, stride=2).backward(dy0.float())
    assert relerr(dw.cpu(), w.grad) < 2e-2 and relerr(db.cpu(), b.grad) < 2e-2


@pytest.mark.parametrize("case", [(8, 8, 128, 64, 4, 2, True), (32, 2, 512, 512, 4, 2, False), (32, 4, 256, 512, 4, 2, False), (8, 16, 64, 32, 4, 2, True),
                                  (16, 4, 64, 48, 3, 1, True)])
def test_small_map_wgrad_as_one_dense_product(L, case):
    """Weight gradient of convolutions on maps of at most 8x8 pixels (pix2pix.py:147-166) as dg_im2col + the 1x1 form of
    dg_umma_conv2d_wgrad (one launch: the blocks of input-channel chunks run along gridDim.z), against the halo-tile launch
    and the float oracle; also with accumulate=1 (the generator runs twice per step, pix2pix.py:90)."""
    N, H, cin, cout, k, s, bias = case
    lib, ctx, st = L.load(), L.ctx(0), L.stream_ptr()
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    Ho = -(-H // s)
    x0 = torch.randn((N, H, H, cin), generator=g).to(torch.bfloat16)
    dy0 = torch.randn((N, Ho, Ho, cout), generator=g).to(torch.bfloat16)
    xd, dyd = x0.cuda(), dy0.cuda()
    cp = conv_params(L, k, k, s, H, H, "same")
    tx, tdy = L.tensor(xd), L.tensor(dyd)
    P_ = N * Ho * Ho
    assert P_ % 8 == 0
    col = torch.full((1, P_ // 8, 8, k * k * cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(lib.dg_im2col(ctx, C.byref(tx), C.byref(cp), Ho, Ho, col.data_ptr(), st))
    tcol = L.tensor(col)
    tdy2 = L.DgTensor(dyd.data_ptr(), L.DG_BF16, 1, P_ // 8, 8, cout, cout, 0)
    one = L.DgConvParams(1, 1, 1, 0, 0, 0, 0.0)
    nbytes = lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tcol), C.byref(tdy2), C.byref(one))
    assert nbytes > 0, lib.dg_last_error().decode()
    wk = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dw = torch.full((k, k, cin, cout), 3.0, device="cuda"); db = torch.full((cout,), 3.0, device="cuda")
    L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tcol), C.byref(tdy2), dw.data_ptr(), db.data_ptr() if bias else None, C.byref(one), 0,
                                     wk.data_ptr(), nbytes, st))
    torch.cuda.synchronize()
    assert not torch.isnan(col.float()).any()
    w = torch.zeros((k, k, cin, cout), dtype=torch.float32, requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    OT.conv2d(x0.float(), w, b, stride=s).backward(dy0.float())
    assert relerr(dw.cpu(), w.grad) < 2e-2
    if bias:
        assert relerr(db.cpu(), b.grad) < 2e-2
    # accumulate on top (second generator pass of the step)
    L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tcol), C.byref(tdy2), dw.data_ptr(), db.data_ptr() if bias else None, C.byref(one), 1,
                                     wk.data_ptr(), nbytes, st))
    torch.cuda.synchronize()
    assert relerr(dw.cpu(), 2.0 * w.grad) < 2e-2
    # the halo-tile launch on the original tensors gives the same gradient
    nb2 = lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(tdy), C.byref(cp))
    if nb2 > 0:
        wk2 = torch.empty(nb2, dtype=torch.uint8, device="cuda")
        dw2 = torch.empty_like(dw)
        L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tx), C.byref(tdy), dw2.data_ptr(), None, C.byref(cp), 0, wk2.data_ptr(), nb2, st))
        torch.cuda.synchronize()
        assert relerr(dw2.cpu(), w.grad) < 2e-2


# ---------------------------------------------------------------- staged (TMA-store) epilogue + fused BatchNorm statistics
@pytest.mark.parametrize("case", [(3, 1, 64, 64, 2, 24, 20), (3, 1, 64, 64, 16, 96, 96), (3, 1, 32, 32, 3, 40, 24), (3, 2, 32, 32, 2, 48, 40),
                                  (3, 1, 32, 64, 2, 33, 19), (3, 2, 64, 64, 2, 48, 16), (3, 1, 64, 16, 1, 20, 12), (1, 1, 192, 32, 2, 16, 16)])
def test_umma_conv_fused_bn_partials_and_slice_output(L, case):
    """dg_umma_conv2d_fwd with bn_partials: the per-CTA rows sum to the per-channel sum / sum of squares of the STORED
    bf16 output (what dg_bn_stats would read), dg_bn_finalize reproduces dg_bn_stats, and the staged epilogue writes a
    channel-slice view without touching its neighbours (srgan.py:154-155,247-248: Conv2D followed by BatchNormalization)."""
    k, s, cin, cout, N, H, W = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    x = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64))
    w = _bf16_round(torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * (1.0 / (k * np.sqrt(cin))))
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    y_ref = OT.conv2d(x, w, b, stride=s, padding="same")
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, s, H, W, "same")
    xd, wd, bd = dev(x, torch.bfloat16), dev(w), dev(b)
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cin, cout, 0, st))
    Ho, Wo = y_ref.shape[1], y_ref.shape[2]
    # output = channels [16, 16+cout) of a wider tensor pre-filled with a sentinel
    ybig = torch.full((N, Ho, Wo, cout + 32), 7.0, device="cuda", dtype=torch.bfloat16)
    tx, ty = L.tensor(xd), L.tensor(ybig, c=cout, coff=16)
    blocks = lib.dg_umma_conv2d_fwd_bn_blocks(ctx, C.byref(tx), C.byref(ty), C.byref(cp))
    if blocks == 0:
        assert (k, s, cin, cout) == (3, 2, 64, 64), "this layer is expected to run the staged epilogue"
        pytest.skip("four parity halos + resident weights leave no room for the staging buffers: the engine calls dg_bn_stats")
    part = torch.full((blocks, 2, cout), float("nan"), device="cuda")
    L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp), part.data_ptr(), st))
    torch.cuda.synchronize()
    y = ybig[..., 16:16 + cout]
    assert relerr(y, y_ref) < BF16_TOL
    assert torch.all(ybig[..., :16] == 7.0) and torch.all(ybig[..., 16 + cout:] == 7.0), "staged store wrote outside its channel slice"
    yd = y.double()
    sums = part.double().sum(0).cpu()
    ref_s, ref_q = yd.sum((0, 1, 2)).cpu(), (yd * yd).sum((0, 1, 2)).cpu()
    assert ((sums[0] - ref_s).abs().max() / ref_q.sqrt().max()).item() < 1e-5
    assert ((sums[1] - ref_q).abs().max() / ref_q.max()).item() < 1e-5
    # finalize == dg_bn_stats on the stored tensor
    gamma, beta = dev(torch.rand(cout, generator=g) + 0.5), dev(torch.randn(cout, generator=g))
    outs = []
    for fused in ("in-kernel", True, False):
        mm, mv = torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda")
        sc, sh, mean, inv = (torch.empty(cout, device="cuda") for _ in range(4))
        if fused == "in-kernel":
            # the conv launch itself finalises (last-CTA ticket); run it twice: the ticket must come back to zero
            fz = L.DgBnFused(gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99, mm.data_ptr(), mv.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                             mean.data_ptr(), inv.data_ptr(), N * Ho * Wo)
            for rep in range(2):
                mm.zero_(); mv.fill_(1.0)
                part.fill_(float("nan"))
                L.check(lib.dg_umma_conv2d_fwd_bn(ctx, C.byref(tx), pk.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp), part.data_ptr(),
                                                  C.byref(fz), st))
            torch.cuda.synchronize()
            assert relerr(ybig[..., 16:16 + cout], y_ref) < BF16_TOL
        elif fused:
            L.check(lib.dg_bn_finalize(ctx, part.data_ptr(), blocks, N * Ho * Wo, cout, gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99,
                                       mm.data_ptr(), mv.data_ptr(), sc.data_ptr(), sh.data_ptr(), mean.data_ptr(), inv.data_ptr(), st))
        else:
            yc = y.contiguous(); tyc = L.tensor(yc)
            nb = lib.dg_bn_workspace_bytes(C.byref(tyc)); wk = ws(nb)
            L.check(lib.dg_bn_stats(ctx, C.byref(tyc), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.99, mm.data_ptr(), mv.data_ptr(),
                                    sc.data_ptr(), sh.data_ptr(), mean.data_ptr(), inv.data_ptr(), wk.data_ptr(), nb, st))
        torch.cuda.synchronize()
        outs.append([t.clone() for t in (sc, sh, mean, inv, mm, mv)])
    for a, b_, c_ in zip(*outs):
        assert relerr(a, c_) < 1e-5 and relerr(b_, c_) < 1e-5


# ---------------------------------------------------------------- conv + BatchNorm + activation (+ Add) in one cooperative launch
@pytest.mark.parametrize("case", [(3, 1, 64, 64, 2, 24, 20, "relu", False, True), (3, 1, 64, 64, 16, 96, 96, "relu", False, False),
                                  (3, 1, 64, 64, 16, 96, 96, None, True, False), (3, 1, 16, 64, 4, 32, 32, "prelu", False, False),
                                  (3, 1, 32, 32, 3, 40, 24, "lrelu", False, True), (3, 2, 32, 32, 2, 48, 40, "lrelu", False, True),
                                  (3, 1, 32, 64, 2, 33, 19, None, True, True), (3, 1, 64, 16, 1, 20, 12, "relu", True, False),
                                  (1, 1, 192, 32, 2, 16, 16, None, True, False), (3, 1, 64, 64, 5, 96, 96, "relu", True, False)])
def test_umma_conv_bn_act_one_launch(L, case):
    """dg_umma_conv2d_fwd_bn_act (conv -> batch statistics -> grid barrier -> scale/shift -> activation (+ skip) from the
    TMEM-resident accumulators) against the three separate calls it replaces (dg_umma_conv2d_fwd with bn_partials,
    dg_bn_finalize, dg_bn_act_fwd): the raw conv output and the activated output must be BIT-identical, the BatchNorm
    coefficients and moving statistics equal to fp32 rounding; and against the float64 oracle within the bf16 bound.
    Run twice back to back: the grid-barrier counters must re-arm (srgan.py:154-157,162-169,246-250)."""
    k, s, cin, cout, N, H, W, act, with_res, bias = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    x = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64))
    w = _bf16_round(torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * (1.0 / (k * np.sqrt(cin))))
    b = torch.randn(cout, generator=g, dtype=torch.float64) if bias else None
    gamma = torch.rand(cout, generator=g, dtype=torch.float64) + 0.5
    beta = torch.randn(cout, generator=g, dtype=torch.float64) * 0.3
    alpha = torch.rand(cout, generator=g, dtype=torch.float64) * 0.4
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, s, H, W, "same")
    Ho, Wo = -(-H // s), -(-W // s)
    res = _bf16_round(torch.randn(N, Ho, Wo, cout, generator=g, dtype=torch.float64)) if with_res else None
    xd, wd = dev(x, torch.bfloat16), dev(w)
    bd = dev(b) if bias else None
    gd, bed, ad = dev(gamma), dev(beta), dev(alpha)
    resd = dev(res, torch.bfloat16) if with_res else None
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cin, cout, 0, st))
    a_code = {None: 0, "relu": 1, "lrelu": 2, "prelu": 5}[act]
    y0 = torch.empty(N, Ho, Wo, cout, device="cuda", dtype=torch.bfloat16); a0 = torch.empty_like(y0)
    y1 = torch.empty_like(y0); a1 = torch.empty_like(y0)
    tx = L.tensor(xd)
    blocks = lib.dg_umma_conv2d_fwd_bn_act_blocks(ctx, C.byref(tx), C.byref(L.tensor(y1)), C.byref(cp))
    if blocks == 0:
        assert (k, s, cin, cout) != (3, 1, 64, 64), "the fused BatchNorm phase must apply to the generator trunk layers"
        pytest.skip("no fused BatchNorm phase for this layer (tiles do not all fit TMEM, or streamed weights): the engine issues the three separate calls")
    P_ = N * Ho * Wo

    def coeffs():
        return [torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda")] + [torch.empty(cout, device="cuda") for _ in range(4)]   # mm, mv, sc, sh, mean, inv

    # ---- reference: the three separate launches
    c0 = coeffs()
    rows = lib.dg_umma_conv2d_fwd_bn_blocks(ctx, C.byref(tx), C.byref(L.tensor(y0)), C.byref(cp))
    assert rows > 0
    part0 = torch.empty(rows, 2, cout, device="cuda")
    ty0, ta0 = L.tensor(y0), L.tensor(a0)
    L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk.data_ptr(), L.ptr(bd), C.byref(ty0), C.byref(cp), part0.data_ptr(), st))
    L.check(lib.dg_bn_finalize(ctx, part0.data_ptr(), rows, P_, cout, gd.data_ptr(), bed.data_ptr(), 1e-3, 0.8, c0[0].data_ptr(), c0[1].data_ptr(),
                               c0[2].data_ptr(), c0[3].data_ptr(), c0[4].data_ptr(), c0[5].data_ptr(), st))
    tres = L.tensor(resd) if with_res else None
    L.check(lib.dg_bn_act_fwd(ctx, C.byref(ty0), c0[2].data_ptr(), c0[3].data_ptr(), a_code, 0.2, ad.data_ptr() if act == "prelu" else None,
                              C.byref(tres) if with_res else None, 0, 0, 0, None, C.byref(ta0), st))
    # ---- the fused launch, twice (moving statistics reset in between)
    part1 = torch.full((blocks, 2, cout), float("nan"), device="cuda")
    ty1, ta1 = L.tensor(y1), L.tensor(a1)
    for rep in range(2):
        c1 = coeffs()
        y1.fill_(7.0); a1.fill_(7.0)
        fz = L.DgBnFused(gd.data_ptr(), bed.data_ptr(), 1e-3, 0.8, c1[0].data_ptr(), c1[1].data_ptr(), c1[2].data_ptr(), c1[3].data_ptr(),
                         c1[4].data_ptr(), c1[5].data_ptr(), P_)
        L.check(lib.dg_umma_conv2d_fwd_bn_act(ctx, C.byref(tx), pk.data_ptr(), L.ptr(bd), C.byref(ty1), C.byref(cp), part1.data_ptr(), C.byref(fz),
                                              a_code, 0.2, ad.data_ptr() if act == "prelu" else None, C.byref(tres) if with_res else None,
                                              C.byref(ta1), st))
        torch.cuda.synchronize()
        assert torch.equal(y1, y0), "raw conv output differs from the unfused launch"
        for u, v in zip(c1, c0):
            assert relerr(u, v) < 1e-6
        if all(torch.equal(u, v) for u, v in zip(c1[2:4], c0[2:4])):
            assert torch.equal(a1, a0), "activated output differs from dg_bn_act_fwd on the stored conv output"
        else:       # coefficients differ in the last bit (different summation grouping of the partial rows): outputs within one bf16 ulp
            assert relerr(a1, a0) < 2 ** -7
    # ---- oracle (float64) on the same bf16-rounded operands
    yr = OT.conv2d(x, w, b, stride=s, padding="same")
    p = {"bn/gamma": gamma, "bn/beta": beta, "bn/moving_mean": torch.zeros(cout, dtype=torch.float64), "bn/moving_variance": torch.ones(cout, dtype=torch.float64)}
    stt = {}
    t = OT.batch_norm(yr, p, "bn", True, stt, momentum=0.8, eps=1e-3)
    if act == "relu": t = torch.relu(t)
    elif act == "lrelu": t = OT.leaky_relu(t, 0.2)
    elif act == "prelu": t = OT.prelu(t, alpha)
    if with_res: t = t + res
    assert relerr(y1, yr) < BF16_TOL and relerr(a1, t) < BF16_TOL
    assert relerr(c1[0], stt["bn/moving_mean"]) < 2e-2 and relerr(c1[1], stt["bn/moving_variance"]) < 2e-2


@pytest.mark.parametrize("case", [(64, 16, 96, 96, "relu", False), (64, 2, 24, 20, None, True), (32, 3, 40, 24, "lrelu", False), (64, 4, 32, 32, "prelu", False),
                                  (256, 1, 12, 20, "relu", True)])
def test_bn_act_fwd_from_partials(L, case):
    """dg_bn_act_fwd_from_partials (finalize folded into the apply pass) against dg_bn_finalize + dg_bn_act_fwd on the same
    per-CTA statistics rows: coefficients and moving statistics to fp32 rounding, outputs within one bf16 ulp
    (srgan.py:155-157,163-169)."""
    C_, N, H, W, act, with_res = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    x = _bf16_round(torch.randn(N, H, W, C_, generator=g, dtype=torch.float64) * 1.3 + 0.2)
    res = _bf16_round(torch.randn(N, H, W, C_, generator=g, dtype=torch.float64)) if with_res else None
    gamma = torch.rand(C_, generator=g, dtype=torch.float64) + 0.5
    beta = torch.randn(C_, generator=g, dtype=torch.float64) * 0.3
    alpha = torch.rand(C_, generator=g, dtype=torch.float64) * 0.4
    rows = 37
    P_ = N * H * W
    # synthetic per-CTA rows: split the pixels into `rows` groups and sum each (what the conv epilogue would have written)
    xf = x.reshape(P_, C_)
    idx = torch.arange(P_) % rows
    part = torch.zeros(rows, 2, C_, dtype=torch.float64)
    part[:, 0].index_add_(0, idx, xf); part[:, 1].index_add_(0, idx, xf * xf)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    xd, partd = dev(x, torch.bfloat16), dev(part)
    resd = dev(res, torch.bfloat16) if with_res else None
    gd, bed, ad = dev(gamma), dev(beta), dev(alpha)
    a_code = {None: 0, "relu": 1, "lrelu": 2, "prelu": 5}[act]
    tx = L.tensor(xd)
    tres = L.tensor(resd) if with_res else None
    outs = []
    for fused in (0, 1):
        mm, mv = torch.zeros(C_, device="cuda"), torch.ones(C_, device="cuda")
        sc, sh, mean, inv = [torch.empty(C_, device="cuda") for _ in range(4)]
        y = torch.full((N, H, W, C_), 7.0, device="cuda", dtype=torch.bfloat16)
        ty = L.tensor(y)
        if fused:
            rc = lib.dg_bn_act_fwd_from_partials(ctx, C.byref(tx), partd.data_ptr(), rows, gd.data_ptr(), bed.data_ptr(), 1e-3, 0.8, mm.data_ptr(),
                                                 mv.data_ptr(), sc.data_ptr(), sh.data_ptr(), mean.data_ptr(), inv.data_ptr(), a_code, 0.2,
                                                 ad.data_ptr() if act == "prelu" else None, C.byref(tres) if with_res else None, C.byref(ty), st)
            assert rc == 0, rc
        else:
            L.check(lib.dg_bn_finalize(ctx, partd.data_ptr(), rows, P_, C_, gd.data_ptr(), bed.data_ptr(), 1e-3, 0.8, mm.data_ptr(), mv.data_ptr(),
                                       sc.data_ptr(), sh.data_ptr(), mean.data_ptr(), inv.data_ptr(), st))
            L.check(lib.dg_bn_act_fwd(ctx, C.byref(tx), sc.data_ptr(), sh.data_ptr(), a_code, 0.2, ad.data_ptr() if act == "prelu" else None,
                                      C.byref(tres) if with_res else None, 0, 0, 0, None, C.byref(ty), st))
        torch.cuda.synchronize()
        outs.append((y, sc, sh, mean, inv, mm, mv))
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert relerr(b, a) < 1e-6
    assert relerr(outs[1][0], outs[0][0]) < 2 ** -7


# ---------------------------------------------------------------- dgrad + skip-add + BatchNorm-backward sums in the epilogue
@pytest.mark.parametrize("case", [(3, 64, 64, 2, 24, 20, "relu", True), (3, 64, 64, 16, 96, 96, "relu", False), (3, 64, 64, 16, 96, 96, None, True),
                                  (3, 32, 64, 3, 40, 24, "lrelu", False), (3, 32, 32, 2, 33, 19, "lrelu", True), (1, 192, 32, 2, 16, 16, None, True),
                                  (3, 64, 64, 5, 96, 96, None, False), (1, 32, 192, 2, 20, 12, "relu", False), (3, 64, 64, 4, 30, 50, "none_nostats", True)])
def test_umma_dgrad_fused_bn_bwd(L, case):
    """dg_umma_conv2d_dgrad_fused + dg_bn_bwd_dx_from_partials against the calls they replace (dg_umma_conv2d_dgrad, dg_add,
    dg_bn_act_bwd): the summed gradient must be bit-identical to dgrad-then-add on bf16 storage up to the single rounding the
    fusion removes (the unfused path rounds the dgrad result to bf16 BEFORE adding the skip gradient), the BatchNorm input
    gradient and dgamma / dbeta must agree within the bf16 bound, and everything must match the float64 oracle
    (autodiff of srgan.py:162-169: conv -> BN -> ReLU -> conv -> BN -> Add)."""
    k, cin, cout, N, H, W, act, with_res = case
    stats = act != "none_nostats"
    if not stats:
        act = None
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    # forward of the layer pair: yb = BN input (raw output of the previous conv), a = act(BN(yb)), z = conv(a)
    yb = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64) * 1.5 + 0.3)
    w = _bf16_round(torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * (1.0 / (k * np.sqrt(cin))))
    gz = _bf16_round(torch.randn(N, H, W, cout, generator=g, dtype=torch.float64))
    gres = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64)) if with_res else None
    gamma = torch.rand(cin, generator=g, dtype=torch.float64) + 0.5
    beta = torch.randn(cin, generator=g, dtype=torch.float64) * 0.3
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, 1, H, W, "same")
    ybd, gzd = dev(yb, torch.bfloat16), dev(gz, torch.bfloat16)
    gresd = dev(gres, torch.bfloat16) if with_res else None
    gd, bed = dev(gamma), dev(beta)
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights(ctx, dev(w).data_ptr(), pk.data_ptr(), k, k, cin, cout, 1, st))
    # BatchNorm forward coefficients of yb through the library
    a_code = {None: 0, "relu": 1, "lrelu": 2}[act]
    tyb = L.tensor(ybd)
    nb = lib.dg_bn_workspace_bytes(C.byref(tyb)); wk = ws(nb)
    sc, sh, mean, inv = [torch.empty(cin, device="cuda") for _ in range(4)]
    L.check(lib.dg_bn_stats(ctx, C.byref(tyb), gd.data_ptr(), bed.data_ptr(), 1e-3, 0.99, None, None, sc.data_ptr(), sh.data_ptr(), mean.data_ptr(),
                            inv.data_ptr(), wk.data_ptr(), nb, st))
    tgz = L.tensor(gzd)
    # ---- reference: dgrad, add, two-pass BatchNorm backward
    dx0 = torch.empty(N, H, W, cin, device="cuda", dtype=torch.bfloat16)
    tdx0 = L.tensor(dx0)
    L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(tgz), pk.data_ptr(), None, C.byref(tdx0), C.byref(cp), st))
    if with_res:
        g0 = torch.empty_like(dx0)
        tr, tg0 = L.tensor(gresd), L.tensor(g0)
        L.check(lib.dg_add(ctx, C.byref(tdx0), C.byref(tr), C.byref(tg0), st))
    else:
        g0 = dx0
    tg0 = L.tensor(g0)
    dyb0 = torch.empty_like(dx0); dga0 = torch.empty(cin, device="cuda"); dbe0 = torch.empty(cin, device="cuda")
    tdyb0 = L.tensor(dyb0)
    L.check(lib.dg_bn_act_bwd(ctx, C.byref(tg0), C.byref(tyb), sc.data_ptr(), sh.data_ptr(), gd.data_ptr(), mean.data_ptr(), inv.data_ptr(), a_code,
                              0.2, None, 0, 0, 0, None, C.byref(tdyb0), dga0.data_ptr(), dbe0.data_ptr(), None, 0, wk.data_ptr(), nb, st))
    # ---- fused: one dgrad launch (+ skip gradient, + sums), one dx pass; twice (no state may be left behind)
    g1 = torch.empty_like(dx0); tg1 = L.tensor(g1)
    rows = lib.dg_umma_conv2d_dgrad_fused_blocks(ctx, C.byref(tgz), C.byref(tg1), C.byref(cp))
    if rows == 0:
        assert (k, cin, cout) != (3, 64, 64), "the fused epilogue must apply to the generator trunk layers"
        pytest.skip("the fused BatchNorm-backward epilogue does not apply to this layer")
    part = torch.full((rows, 2, cin), float("nan"), device="cuda")
    dyb1 = torch.empty_like(dx0); dga1 = torch.empty(cin, device="cuda"); dbe1 = torch.empty(cin, device="cuda")
    tdyb1 = L.tensor(dyb1)
    for rep in range(2):
        g1.fill_(7.0); dyb1.fill_(7.0); part.fill_(float("nan"))
        bs = L.DgBnBwdStats(C.pointer(tyb), sc.data_ptr(), sh.data_ptr(), mean.data_ptr(), a_code, 0.2, part.data_ptr())
        tr = L.tensor(gresd) if with_res else None
        L.check(lib.dg_umma_conv2d_dgrad_fused(ctx, C.byref(tgz), pk.data_ptr(), C.byref(tg1), C.byref(cp), C.byref(tr) if with_res else None,
                                               C.byref(bs) if stats else None, st))
        if stats:
            L.check(lib.dg_bn_bwd_dx_from_partials(ctx, C.byref(tg1), C.byref(tyb), sc.data_ptr(), sh.data_ptr(), gd.data_ptr(), mean.data_ptr(),
                                                   inv.data_ptr(), a_code, 0.2, part.data_ptr(), rows, C.byref(tdyb1), dga1.data_ptr(),
                                                   dbe1.data_ptr(), 0, st))
        torch.cuda.synchronize()
        if not with_res:
            assert torch.equal(g1, g0), "without a skip gradient the stored gradient must equal the plain dgrad bit for bit"
        else:
            assert relerr(g1, g0) < 2 ** -7
        if stats:
            assert not torch.isnan(part).any()
            assert relerr(dbe1, dbe0) < 5e-3 and relerr(dga1, dga0) < 5e-3, (relerr(dbe1, dbe0), relerr(dga1, dga0))
            assert relerr(dyb1, dyb0) < BF16_TOL
    # ---- float64 oracle of the same sub-graph
    ybr = yb.clone().requires_grad_(True); gar = gamma.clone().requires_grad_(True); ber = beta.clone().requires_grad_(True)
    p = {"bn/gamma": gar, "bn/beta": ber, "bn/moving_mean": torch.zeros(cin, dtype=torch.float64), "bn/moving_variance": torch.ones(cin, dtype=torch.float64)}
    t = OT.batch_norm(ybr, p, "bn", True, {}, momentum=0.99, eps=1e-3)
    if act == "relu": t = torch.relu(t)
    elif act == "lrelu": t = OT.leaky_relu(t, 0.2)
    t.retain_grad()
    z = OT.conv2d(t, w, None, stride=1, padding="same")
    loss = (z * gz).sum() + ((t * gres).sum() if with_res else 0.0)
    loss.backward()
    assert relerr(g1, t.grad) < BF16_TOL
    if stats:
        assert relerr(dyb1, ybr.grad) < BF16_TOL and relerr(dga1, gar.grad) < BF16_TOL and relerr(dbe1, ber.grad) < BF16_TOL


# ---------------------------------------------------------------- K-outer mode (streamed weights reused by several sub-tiles)
@pytest.mark.parametrize("case", [(3, 256, 64, 2, 70, 20, "dgrad"), (3, 256, 64, 1, 64, 24, "fwd"), (3, 128, 256, 2, 40, 17, "fwd"),
                                  (3, 256, 256, 1, 48, 16, "fwd"), (1, 512, 64, 2, 33, 9, "fwd"), (3, 192, 96, 1, 30, 30, "dgrad")])
def test_umma_conv_k_outer_weight_ring(L, case):
    """The chunk-outer / sub-tile-inner loop order with its two-slot weight ring (conv_umma.cu `kouter`), forced through
    the debug flag so that small shapes take it: partial tiles in both directions, several N blocks, 2-8 chunks.
    (In production the cost model picks it for e.g. the 256->64 dgrad of the generator's upsampling conv, srgan.py:144.)"""
    k, cin, cout, N, H, W, kind = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    w = _bf16_round(torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * (1.0 / (k * np.sqrt(cin))))
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, 1, H, W, "same")
    wd = dev(w)
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cin, cout, 0 if kind == "fwd" else 1, st))
    lib.dg_debug_conv_flags(16)
    try:
        if kind == "fwd":
            x = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64))
            b = torch.randn(cout, generator=g, dtype=torch.float64)
            ref = OT.conv2d(x, w, b, stride=1, padding="same")
            y = torch.full(ref.shape, float("nan"), device="cuda", dtype=torch.bfloat16)
            xd, bd = dev(x, torch.bfloat16), dev(b)
            tx, ty = L.tensor(xd), L.tensor(y)
            L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp), None, st))
            out = y
        else:
            xr = torch.zeros(N, H, W, cin, dtype=torch.float64, requires_grad=True)
            gy = _bf16_round(torch.randn(N, H, W, cout, generator=g, dtype=torch.float64))
            (OT.conv2d(xr, w, None, stride=1, padding="same") * gy).sum().backward()
            ref = xr.grad
            dx = torch.full(ref.shape, float("nan"), device="cuda", dtype=torch.bfloat16)
            gyd = dev(gy, torch.bfloat16)
            tg, tdx = L.tensor(gyd), L.tensor(dx)
            L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(tg), pk.data_ptr(), None, C.byref(tdx), C.byref(cp), st))
            out = dx
        torch.cuda.synchronize()
    finally:
        lib.dg_debug_conv_flags(0)
    assert relerr(out, ref) < BF16_TOL


# ---------------------------------------------------------------- stride-2 dgrad: four output phases in one launch
@pytest.mark.parametrize("case", [(3, 32, 32, 2, 40, 24), (3, 64, 64, 2, 24, 36), (4, 64, 32, 1, 32, 16), (3, 32, 64, 3, 18, 10), (4, 128, 64, 1, 16, 16)])
def test_umma_dgrad_stride2_fused_phases(L, case):
    """dx of a stride-2 Conv2D (= Conv2DTranspose forward, pix2pix.py:130) with all four output parity phases computed by ONE
    launch when their accumulators fit TMEM (conv_umma.cu `n_phase`), against the oracle and against the one-launch-per-phase
    path it replaces; odd tile counts in both directions, bias + LeakyReLU epilogue on the transposed-conv output."""
    k, cin, cout, N, H, W = case          # forward conv: [N,H,W,cin] -> [N,H/2,W/2,cout]
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    w = _bf16_round(torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * (1.0 / (k * np.sqrt(cout))))
    gy = _bf16_round(torch.randn(N, H // 2, W // 2, cout, generator=g, dtype=torch.float64))
    xr = torch.zeros(N, H, W, cin, dtype=torch.float64, requires_grad=True)
    (OT.conv2d(xr, w, None, stride=2, padding="same") * gy).sum().backward()
    ref = xr.grad
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, 2, H, W, "same")
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    wd = dev(w)
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cin, cout, 1, st))
    gyd = dev(gy, torch.bfloat16)
    dx = torch.full(ref.shape, float("nan"), device="cuda", dtype=torch.bfloat16)
    tg, tdx = L.tensor(gyd), L.tensor(dx)
    L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(tg), pk.data_ptr(), None, C.byref(tdx), C.byref(cp), st))
    torch.cuda.synchronize()
    assert relerr(dx, ref) < BF16_TOL
    # bias + activation epilogue (Conv2DTranspose forward with bias, fp32 output)
    b = torch.randn(cin, generator=g, dtype=torch.float64)
    cpa = conv_params(L, k, k, 2, H, W, "same", act=2, alpha=0.3)
    out = torch.full(ref.shape, float("nan"), device="cuda", dtype=torch.float32)
    bd = dev(b); to = L.tensor(out)
    L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(tg), pk.data_ptr(), bd.data_ptr(), C.byref(to), C.byref(cpa), st))
    torch.cuda.synchronize()
    pre = ref + b
    assert relerr(out, torch.where(pre >= 0, pre, 0.3 * pre)) < 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_pool_upsample_vector_kernels(L, dtype):
    """8-channel-vector MaxPool2D(2,2) / UpSampling2D(2)+ReLU (autoencoder.py:110,113-136) on channel counts that are multiples of
    8, through channel-slice views (the up-sampled half is written straight into the concat buffer): bit-exact against the
    oracle on inputs that are exact in the storage type (max, relu and routing do no arithmetic; the backward sum of four
    bf16 values is rounded once)."""
    g = torch.Generator().manual_seed(11)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    x = torch.randn(3, 6, 10, 24, generator=g).to(dtype).double()
    x[0, 0, 0, :8] = x[0, 0, 1, :8]                      # ties inside a pooling window: the first position gets the gradient
    xr = x.clone().requires_grad_(True)
    ref = OT.max_pool2x2(xr); gy = torch.randn(ref.shape, generator=g).to(dtype).double()
    (ref * gy).sum().backward()
    wide_x = torch.zeros(3, 6, 10, 40, device="cuda", dtype=dtype); wide_x[..., 8:32] = x.to(dtype).cuda()
    wide_y = torch.zeros(3, 3, 5, 32, device="cuda", dtype=dtype)
    gyd = dev(gy, dtype); dx = torch.zeros(3, 6, 10, 24, device="cuda", dtype=dtype)
    tx, ty, tg, tdx = L.tensor(wide_x, c=24, coff=8), L.tensor(wide_y, c=24, coff=0), L.tensor(gyd), L.tensor(dx)
    L.check(lib.dg_maxpool2x2_fwd(ctx, C.byref(tx), C.byref(ty), st))
    L.check(lib.dg_maxpool2x2_bwd(ctx, C.byref(tg), C.byref(tx), C.byref(ty), C.byref(tdx), st))
    assert torch.equal(wide_y[..., :24].double().cpu(), ref.detach()) and (wide_y[..., 24:] == 0).all()
    assert torch.equal(dx.double().cpu(), xr.grad)
    # up-sampling + relu into a slice of a wider buffer, gradient read back from a slice
    xr.grad = None
    ref = torch.relu(OT.upsample2x_nearest(xr)); gy = torch.randn(ref.shape, generator=g).to(dtype).double()
    (ref * gy).sum().backward()
    cat = torch.zeros(3, 12, 20, 48, device="cuda", dtype=dtype)
    tc = L.tensor(cat, c=24, coff=16)
    L.check(lib.dg_upsample2x_relu_fwd(ctx, C.byref(tx), C.byref(tc), st))
    assert torch.equal(cat[..., 16:40].double().cpu(), ref.detach()) and (cat[..., :16] == 0).all() and (cat[..., 40:] == 0).all()
    gcat = torch.zeros(3, 12, 20, 48, device="cuda", dtype=dtype); gcat[..., 16:40] = gy.to(dtype).cuda()
    tgc = L.tensor(gcat, c=24, coff=16)
    dx2 = torch.zeros_like(dx); tdx2 = L.tensor(dx2)
    L.check(lib.dg_upsample2x_relu_bwd(ctx, C.byref(tgc), C.byref(tx), C.byref(tdx2), st))
    assert relerr(dx2, xr.grad) < (1e-6 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("mode", [0, 1])
def test_pack_and_unpad_with_channel_segments(L, mode):
    """dg_umma_pack_weights_seg / dg_unpad_weight_grad_seg: the input-channel axis of a kernel that consumes a U-Net concat of two
    zero-padded tensors (autoencoder.py:135: 100 of 112 channels, then 76 of 80) -- checked against dg_umma_pack_weights of the
    same kernel scattered into the padded layout on the host, and against a host gather for the gradient."""
    g = torch.Generator().manual_seed(5)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    kh = kw = 3; c1, c1p, c2, c2p, cout, cout_p = 100, 112, 76, 80, 152, 160
    cin, cin_p = c1 + c2, c1p + c2p
    w = torch.randn(kh, kw, cin, cout, generator=g)
    wp = torch.zeros(kh, kw, cin_p, cout_p)
    wp[:, :, :c1, :cout] = w[:, :, :c1]; wp[:, :, c1p:c1p + c2, :cout] = w[:, :, c1:]
    wd, wpd = dev(w), dev(wp)
    a = torch.empty(kh * kw * cin_p * cout_p, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
    L.check(lib.dg_umma_pack_weights_seg(ctx, wd.data_ptr(), a.data_ptr(), kh, kw, cin, cout, cin_p, cout_p, c1, c1p, mode, st))
    L.check(lib.dg_umma_pack_weights(ctx, wpd.data_ptr(), b.data_ptr(), kh, kw, cin_p, cout_p, mode, st))
    assert torch.equal(a, b)
    # a single padded segment is the plain padded packing
    L.check(lib.dg_umma_pack_weights_seg(ctx, wd.data_ptr(), a.data_ptr(), kh, kw, cin, cout, cin_p, cout_p, 0, 0, mode, st))
    L.check(lib.dg_umma_pack_weights_padded(ctx, wd.data_ptr(), b.data_ptr(), kh, kw, cin, cout, cin_p, cout_p, mode, st))
    assert torch.equal(a, b)
    if mode == 0:
        dwp = torch.randn(kh, kw, cin_p, cout_p, generator=g); dbp = torch.randn(cout_p, generator=g)
        dw = torch.ones(kh, kw, cin, cout, device="cuda"); db = torch.ones(cout, device="cuda")
        dwpd, dbpd = dev(dwp), dev(dbp)
        L.check(lib.dg_unpad_weight_grad_seg(ctx, dwpd.data_ptr(), dbpd.data_ptr(), dw.data_ptr(), db.data_ptr(), kh, kw, cin, cout, cin_p,
                                             cout_p, c1, c1p, 1, st))
        ref = torch.cat([dwp[:, :, :c1, :cout], dwp[:, :, c1p:c1p + c2, :cout]], dim=2) + 1
        assert torch.equal(dw.cpu(), ref) and torch.equal(db.cpu(), dbp[:cout] + 1)


@pytest.mark.parametrize("channels", [5, 24])
def test_maxpool_bwd_with_folded_relu(L, channels):
    """dg_maxpool2x2_bwd_relu = gradient of maxpool(relu(u)) with respect to u, given x = relu(u) (autoencoder.py:95-110): the ReLU
    mask of the convolution in front of the pooling is applied by the pooling's backward pass (scalar and 8-channel kernels)."""
    g = torch.Generator().manual_seed(21)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    u = torch.randn(2, 8, 6, channels, generator=g, dtype=torch.float64)
    u[0, :2, :2, 0] = -1.0                                   # a window that is zero everywhere after the ReLU: no gradient at all
    ur = u.clone().requires_grad_(True)
    ref = OT.max_pool2x2(torch.relu(ur)); gy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * gy).sum().backward()
    xd, gyd = dev(torch.relu(u)), dev(gy)
    y = torch.empty(ref.shape, device="cuda"); dx = torch.empty_like(xd)
    tx, ty, tg, tdx = L.tensor(xd), L.tensor(y), L.tensor(gyd), L.tensor(dx)
    L.check(lib.dg_maxpool2x2_fwd(ctx, C.byref(tx), C.byref(ty), st))
    L.check(lib.dg_maxpool2x2_bwd_relu(ctx, C.byref(tg), C.byref(tx), C.byref(ty), C.byref(tdx), st))
    assert relerr(y, ref) < 1e-6 and relerr(dx, ur.grad) < 1e-6
    assert (dx[0, :2, :2, 0] == 0).all()


@pytest.mark.parametrize("C_,H,W", [(32, 13, 70), (64, 8, 32), (192, 19, 45)])
def test_depthwise_bf16_tma_tiles(L, C_, H, W):
    """bf16 DepthwiseConv2D (fsrgan.py:149-154) through the TMA-fed tile kernel: partial tiles in both directions, SAME padding from
    the TMA zero fill, channel-slice views, the folded-BatchNorm inference form (bias + ReLU) and the input gradient (flipped taps).
    Inputs are exact in bf16 and the fp32 taps are used as they are, so every output is the correctly rounded fp32 sum: checked
    element by element to one bf16 ulp of the float64 oracle."""
    g = torch.Generator().manual_seed(C_ + W)
    x = torch.randn(2, H, W, C_, generator=g).bfloat16().double()
    w = torch.randn(3, 3, C_, 1, generator=g).double(); b = torch.randn(C_, generator=g).double()
    xr = x.clone().requires_grad_(True)
    ref = OT.depthwise_conv2d(xr, w, b)
    gy = torch.randn(ref.shape, generator=g).bfloat16().double()
    (OT.depthwise_conv2d(xr, w, None) * gy).sum().backward()
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    wide = torch.zeros(2, H, W, C_ + 16, device="cuda", dtype=torch.bfloat16); wide[..., 8:8 + C_] = x.to(torch.bfloat16).cuda()
    wd, bd, gyd = dev(w), dev(b), dev(gy, torch.bfloat16)
    y = torch.zeros(2, H, W, C_ + 8, device="cuda", dtype=torch.bfloat16); dx = torch.empty(2, H, W, C_, device="cuda", dtype=torch.bfloat16)
    tx, ty, tg, tdx = L.tensor(wide, c=C_, coff=8), L.tensor(y, c=C_, coff=0), L.tensor(gyd), L.tensor(dx)

    def close(a, r):
        a, r = a.double().cpu(), r.detach()
        return bool(((a - r).abs() <= 2.0 ** -8 * r.abs() + 1e-5).all())
    L.check(lib.dg_dwconv3x3_fwd(ctx, C.byref(tx), wd.data_ptr(), bd.data_ptr(), C.byref(ty), st))
    assert close(y[..., :C_], ref) and (y[..., C_:] == 0).all()
    L.check(lib.dg_dwconv3x3_fwd_act(ctx, C.byref(tx), wd.data_ptr(), bd.data_ptr(), 1, C.byref(ty), st))
    assert close(y[..., :C_], torch.relu(ref))
    L.check(lib.dg_dwconv3x3_dgrad(ctx, C.byref(tg), wd.data_ptr(), C.byref(tdx), st))
    assert close(dx, xr.grad)


@pytest.mark.parametrize("cout,act", [(3, 3), (1, 4), (5, 0)])
def test_conv_fwd_narrow_store(L, cout, act):
    """dg_umma_conv2d_fwd_narrow: a convolution with fewer than 16 output channels (the RGB image, srgan.py:182) computed on the
    tensor cores with the kernel padded to 16 output channels, the real channels stored densely in fp32 (tanh / sigmoid / none)."""
    g = torch.Generator().manual_seed(cout)
    N, H, W, cin, k = 2, 21, 19, 32, 3
    x = torch.randn(N, H, W, cin, generator=g).bfloat16().double()
    w = (torch.randn(k, k, cin, cout, generator=g) * 0.1).bfloat16().double(); b = torch.randn(cout, generator=g).double()
    ref = OT.conv2d(x, w, b, stride=1, padding="same")
    ref = torch.tanh(ref) if act == 3 else (torch.sigmoid(ref) if act == 4 else ref)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    xd, wd = dev(x, torch.bfloat16), dev(w)
    bp = torch.zeros(16, device="cuda"); bp[:cout] = b.float().cuda()
    pk = torch.empty(k * k * cin * 16, dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights_padded(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cin, cout, cin, 16, 0, st))
    guard = torch.full((N * H * W * cout + 64,), 7.0, device="cuda")
    y = guard[:N * H * W * cout].view(N, H, W, cout)
    cp = conv_params(L, k, k, 1, H, W, "same", act)
    tx, ty = L.tensor(xd), L.tensor(y)
    L.check(lib.dg_umma_conv2d_fwd_narrow(ctx, C.byref(tx), pk.data_ptr(), bp.data_ptr(), C.byref(ty), C.byref(cp), st))
    assert relerr(y, ref) < 2e-3                    # fp32 output of bf16 operands: accumulation order only
    assert (guard[N * H * W * cout:] == 7.0).all()  # nothing written past the dense tensor


@pytest.mark.parametrize("case", [(3, 32, 32, 2, 21, 19, True, True, 0), (1, 192, 32, 1, 40, 24, False, True, 0), (3, 64, 64, 1, 32, 32, True, False, 0),
                                  (3, 32, 32, 1, 16, 24, False, True, 1)])
def test_conv_fwd_res_prelu_epilogue(L, case):
    """dg_umma_conv2d_fwd_res_prelu: Conv2D (+ folded inference BatchNorm) -> [PReLU] -> [+ skip] in the staged epilogue
    (fsrgan.py:172-176, :208-210; srgan.py:166-169) against the oracle's separate ops; the skip tensor may be a channel slice."""
    k, cin, cout, N, H, W, use_prelu, use_res, act = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    x = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64))
    w = _bf16_round(torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * 0.1)
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    alpha = torch.rand(cout, generator=g, dtype=torch.float64) - 0.3
    res = _bf16_round(torch.randn(N, H, W, cout, generator=g, dtype=torch.float64))
    ref = OT.conv2d(x, w, b, stride=1, padding="same")
    if use_prelu:
        ref = OT.prelu(ref, alpha)
    elif act == 1:
        ref = torch.relu(ref)
    if use_res:
        ref = ref + res
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, 1, H, W, "same", act)
    xd, wd, bd, ad = dev(x, torch.bfloat16), dev(w), dev(b), dev(alpha)
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cin, cout, 0, st))
    y = torch.full((N, H, W, cout), 7.0, device="cuda", dtype=torch.bfloat16)
    wide = torch.randn(N, H, W, cout + 16, generator=g).bfloat16().cuda()
    wide[..., 8:8 + cout] = res.to(torch.bfloat16).cuda()
    tx, ty, tr = L.tensor(xd), L.tensor(y), L.tensor(wide, c=cout, coff=8)
    if lib.dg_umma_conv2d_fwd_bn_blocks(ctx, C.byref(tx), C.byref(ty), C.byref(cp)) <= 0:
        pytest.skip("the staged epilogue does not apply to this layer")
    L.check(lib.dg_umma_conv2d_fwd_res_prelu(ctx, C.byref(tx), pk.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp),
                                             C.byref(tr) if use_res else None, ad.data_ptr() if use_prelu else None, st))
    torch.cuda.synchronize()
    assert relerr(y, ref) < BF16_TOL
    # one rounding: the result is the bf16 rounding of the fp32 value act(conv + bias) + skip, not of a rounded intermediate
    err = (y.double().cpu() - ref).abs()
    assert (err <= 2.0 ** -8 * ref.abs() + 1e-4).all()


TAPSUM_CASES = [  # (cout, act, N, H, W, channel offset of the 32-channel input inside a wider buffer)
    (3, 3, 2, 21, 19, 0),      # smaller than one 30 x 14 tile (the halo box is wider than the image)
    (3, 3, 1, 45, 70, 8),      # 4 x 3 tiles with ragged right / bottom tiles, input = a channel slice (pixel pitch 48)
    (1, 4, 2, 28, 60, 0),      # exact multiples of the tile
    (2, 0, 1, 15, 31, 0),      # one pixel more than a tile in both directions
    (3, 0, 3, 64, 96, 0),      # more tiles than fit two rounds of a small grid
]


@pytest.mark.parametrize("case", TAPSUM_CASES)
def test_conv3x3_tapsum_fwd(L, case):
    """dg_conv3x3_tapsum_fwd (fsrgan.py:216-217 at inference): one tensor-core product against all nine taps (N = 9*cout) plus nine
    shifted fp32 adds must equal Conv2D(cout, 3, padding='same') + activation; bf16 operands, fp32 accumulation -> 1e-5."""
    cout, act, N, H, W, coff = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    x = torch.randn(N, H, W, 32, generator=g).bfloat16().double()
    w = (torch.randn(3, 3, 32, cout, generator=g) * 0.1).bfloat16().double(); b = torch.randn(cout, generator=g).double()
    ref = OT.conv2d(x, w, b, stride=1, padding="same")
    ref = torch.tanh(ref) if act == 3 else (torch.sigmoid(ref) if act == 4 else ref)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    wide = torch.randn(N, H, W, 32 + 2 * coff, generator=g).bfloat16().cuda()      # neighbouring channels must not leak in
    wide[..., coff:coff + 32] = x.to(torch.bfloat16).cuda()
    tx = L.tensor(wide, c=32, coff=coff)
    assert lib.dg_conv3x3_tapsum_supported(ctx, C.byref(tx), cout) == 1
    wd, bd = dev(w), dev(b)
    guard = torch.full((N * H * W * cout + 64,), 7.0, device="cuda")
    y = guard[:N * H * W * cout].view(N, H, W, cout)
    ty = L.tensor(y)
    L.check(lib.dg_conv3x3_tapsum_fwd(ctx, C.byref(tx), wd.data_ptr(), bd.data_ptr(), act, 0.0, C.byref(ty), st))
    err = (y.double().cpu() - ref).abs().max().item()
    assert err < FP32_TOL * max(1.0, ref.abs().max().item()), err
    assert (guard[N * H * W * cout:] == 7.0).all()
    # no bias: same call with a null pointer
    L.check(lib.dg_conv3x3_tapsum_fwd(ctx, C.byref(tx), wd.data_ptr(), None, 0, 0.0, C.byref(ty), st))
    assert (y.double().cpu() - OT.conv2d(x, w, None, stride=1, padding="same")).abs().max().item() < FP32_TOL * max(1.0, ref.abs().max().item())
    # and the implicit-GEMM kernel it replaces gives the same image (accumulation order only)
    if cout == 3 and coff == 0:
        bp = torch.zeros(16, device="cuda"); bp[:cout] = bd
        pk = torch.empty(9 * 32 * 16, dtype=torch.bfloat16, device="cuda")
        L.check(lib.dg_umma_pack_weights_padded(ctx, wd.data_ptr(), pk.data_ptr(), 3, 3, 32, cout, 32, 16, 0, st))
        y2 = torch.empty_like(y)
        cp = conv_params(L, 3, 3, 1, H, W, "same", act)
        t2 = L.tensor(y2)
        L.check(lib.dg_conv3x3_tapsum_fwd(ctx, C.byref(tx), wd.data_ptr(), bd.data_ptr(), act, 0.0, C.byref(ty), st))
        L.check(lib.dg_umma_conv2d_fwd_narrow(ctx, C.byref(tx), pk.data_ptr(), bp.data_ptr(), C.byref(t2), C.byref(cp), st))
        assert (y - y2).abs().max().item() < 2e-5


@pytest.mark.parametrize("geom", [(1, 40, 70, 40, 70), (1, 45, 70, 33, 52), (1, 64, 128, 20, 120), (1, 30, 44, 29, 43)])
@pytest.mark.parametrize("clip,flip", [(1, 0), (0, 1)])
def test_conv3x3_tapsum_frame_bit_exact(L, geom, clip, flip):
    """dg_conv3x3_tapsum_frame writes the uint8 frame of infer_video.py:150-159 / infer.py:62-68 straight from the convolution: bit for
    bit what dg_float_to_frame makes of dg_conv3x3_tapsum_fwd's float image (centre crop, (y+1)/2, clip, *255, truncation, flip)."""
    N, H, W, dh, dw = geom
    g = torch.Generator().manual_seed(zlib.crc32(str((geom, clip, flip)).encode()) & 0xFFFF)
    x = (torch.randn(N, H, W, 32, generator=g) * 1.5).bfloat16().cuda()
    w = dev(torch.randn(3, 3, 32, 3, generator=g) * 0.15); b = dev(torch.randn(3, generator=g) * 0.3)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    tx = L.tensor(x)
    y = torch.empty(N, H, W, 3, device="cuda"); ty = L.tensor(y)
    L.check(lib.dg_conv3x3_tapsum_fwd(ctx, C.byref(tx), w.data_ptr(), b.data_ptr(), 3, 0.0, C.byref(ty), st))
    ref = torch.empty(dh, dw, 3, dtype=torch.uint8, device="cuda")
    L.check(lib.dg_float_to_frame(ctx, C.byref(ty), 0.6, 0.5, clip, flip, ref.data_ptr(), dh, dw, st))     # 0.6: some values leave [0, 1]
    guard = torch.full((dh * dw * 3 + 64,), 77, dtype=torch.uint8, device="cuda")
    L.check(lib.dg_conv3x3_tapsum_frame(ctx, C.byref(tx), w.data_ptr(), b.data_ptr(), 3, 0.0, 0.6, 0.5, clip, flip, guard.data_ptr(), dh, dw, st))
    assert torch.equal(guard[:dh * dw * 3].view(dh, dw, 3), ref)
    assert (guard[dh * dw * 3:] == 77).all()
    assert len(torch.unique(ref)) > 50                      # the comparison is not between two constant images
    # a frame larger than the convolution's output is refused (crop only)
    assert lib.dg_conv3x3_tapsum_frame(ctx, C.byref(tx), w.data_ptr(), b.data_ptr(), 3, 0.0, 0.5, 0.5, 1, 0, guard.data_ptr(), H + 1, W, st) != 0


@pytest.mark.parametrize("case", [(3, 64, 64, 2, 40, 24, False, 4), (3, 64, 64, 1, 96, 96, False, 3), (3, 32, 48, 2, 24, 24, True, 2),
                                  (1, 64, 16, 1, 48, 40, True, 4)])
@pytest.mark.parametrize("merged", [False, True])
def test_umma_wgrad_batch(L, case, merged):
    """dg_umma_conv2d_wgrad_batch: the weight gradients of n layers of identical geometry in one launch, the SMs divided among the
    problems -- distinct outputs (the generator trunk's identical convolutions, srgan.py:161-172) and one shared output (the real and
    fake passes of a discriminator layer, train_srgan.py:78-79), with and without a previous content to accumulate on."""
    k, cin, cout, N, H, W, bias, n = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, 1, H, W, "same")
    xs, gys, refs_w, refs_b = [], [], [], []
    for i in range(n):
        x = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64))
        w = torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64).requires_grad_(True)
        b = torch.randn(cout, generator=g, dtype=torch.float64).requires_grad_(True)
        y_ref = OT.conv2d(x, w, b, stride=1, padding="same")
        gy = _bf16_round(torch.randn(y_ref.shape, generator=g, dtype=torch.float64))
        (y_ref * gy).sum().backward()
        xs.append(dev(x, torch.bfloat16)); gys.append(dev(gy, torch.bfloat16)); refs_w.append(w.grad); refs_b.append(b.grad)
    tens = [(L.tensor(a), L.tensor(b_)) for a, b_ in zip(xs, gys)]
    nb = lib.dg_umma_conv2d_wgrad_batch_workspace_bytes(n, C.byref(tens[0][0]), C.byref(tens[0][1]), C.byref(cp))
    assert nb > 0
    wk = ws(nb)
    n_out = 1 if merged else n
    dws = [torch.full((k, k, cin, cout), 0.5, device="cuda") for _ in range(n_out)]
    dbs = [torch.full((cout,), 0.25, device="cuda") for _ in range(n_out)]
    for acc in (0, 1):
        px = (C.POINTER(L.DgTensor) * n)(*[C.pointer(t[0]) for t in tens])
        pd = (C.POINTER(L.DgTensor) * n)(*[C.pointer(t[1]) for t in tens])
        pw = (C.c_void_p * n)(*[dws[0 if merged else i].data_ptr() for i in range(n)])
        pb = (C.c_void_p * n)(*[dbs[0 if merged else i].data_ptr() for i in range(n)])
        pa = (C.c_int * n)(*([acc] * n))
        L.check(lib.dg_umma_conv2d_wgrad_batch(ctx, n, px, pd, pw, pb if bias else None, C.byref(cp), pa, wk.data_ptr(), nb, st))
        torch.cuda.synchronize()
        if merged:
            assert relerr(dws[0], (acc + 1) * sum(refs_w)) < 1e-4
            if bias:
                assert relerr(dbs[0], (acc + 1) * sum(refs_b)) < 1e-4
        else:
            for i in range(n):
                assert relerr(dws[i], (acc + 1) * refs_w[i]) < 1e-4, (i, acc)
                if bias:
                    assert relerr(dbs[i], (acc + 1) * refs_b[i]) < 1e-4, (i, acc)


@pytest.mark.parametrize("case", [(3, 32, 32, 2, 40, 24), (3, 64, 160, 1, 32, 48), (1, 96, 64, 2, 24, 24), (3, 64, 64, 16, 96, 96)])
def test_dgrad_with_folded_relu_mask(L, case):
    """dg_umma_conv2d_dgrad_relu_mask: the input gradient of a convolution whose input was y = relu(conv(...)) (autoencoder.py:95-104),
    stored already multiplied by (y > 0) -- against dgrad followed by the oracle's ReLU backward, and bit for bit against the plain
    dgrad launch masked on the host (the mask only zeroes stored bf16 values)."""
    k, cin, cout, N, H, W = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, 1, H, W, "same")
    u = torch.randn(N, H, W, cin, generator=g, dtype=torch.float64)
    y = _bf16_round(torch.relu(u))                                   # the producing layer's stored output (about half zeros)
    w = _bf16_round(torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * 0.1)
    yr = y.clone().requires_grad_(True)
    out = OT.conv2d(yr, w, None, stride=1, padding="same")
    gy = _bf16_round(torch.randn(out.shape, generator=g, dtype=torch.float64))
    (out * gy).sum().backward()
    ref = yr.grad * (y > 0)
    yd, gyd, wd = dev(y, torch.bfloat16), dev(gy, torch.bfloat16), dev(w)
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cin, cout, 1, st))
    dx = torch.empty(N, H, W, cin, device="cuda", dtype=torch.bfloat16); dx0 = torch.empty_like(dx)
    tg, tdx, tdx0, ty = L.tensor(gyd), L.tensor(dx), L.tensor(dx0), L.tensor(yd)
    if lib.dg_umma_conv2d_dgrad_fused_blocks(ctx, C.byref(tg), C.byref(tdx), C.byref(cp)) <= 0:
        pytest.skip("the staged 32/64-channel epilogue does not apply to this layer: callers use the plain launch")
    L.check(lib.dg_umma_conv2d_dgrad_relu_mask(ctx, C.byref(tg), pk.data_ptr(), C.byref(tdx), C.byref(cp), C.byref(ty), st))
    L.check(lib.dg_umma_conv2d_dgrad(ctx, C.byref(tg), pk.data_ptr(), None, C.byref(tdx0), C.byref(cp), st))
    torch.cuda.synchronize()
    assert relerr(dx, ref) < BF16_TOL
    assert torch.equal(dx, dx0 * (yd > 0))


@pytest.mark.parametrize("case", [(3, 32, 128, 2, 21, 19, True), (3, 64, 256, 1, 32, 24, True), (3, 32, 128, 1, 16, 16, False)])
def test_conv_d2s_prelu_store(L, case):
    """dg_umma_conv2d_fwd_d2s_prelu: Conv2D + depth_to_space(2) + PReLU (srgan.py:144-146, fsrgan.py:180-186) in one launch, the
    convolution's epilogue storing in TensorFlow's DCR order -- against the oracle's three ops."""
    k, cin, cout, N, H, W, use_prelu = case
    g = torch.Generator().manual_seed(zlib.crc32(str(case).encode()) & 0xFFFF)
    x = _bf16_round(torch.randn(N, H, W, cin, generator=g, dtype=torch.float64))
    w = _bf16_round(torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * 0.1)
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    alpha = torch.rand(cout // 4, generator=g, dtype=torch.float64) - 0.3
    ref = OT.depth_to_space(OT.conv2d(x, w, b, stride=1, padding="same"), 2)
    if use_prelu:
        ref = OT.prelu(ref, alpha)
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    cp = conv_params(L, k, k, 1, H, W, "same")
    xd, wd, bd, ad = dev(x, torch.bfloat16), dev(w), dev(b), dev(alpha)
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(lib.dg_umma_pack_weights(ctx, wd.data_ptr(), pk.data_ptr(), k, k, cin, cout, 0, st))
    y = torch.full((N, 2 * H, 2 * W, cout // 4), 7.0, device="cuda", dtype=torch.bfloat16)
    tx, ty = L.tensor(xd), L.tensor(y)
    L.check(lib.dg_umma_conv2d_fwd_d2s_prelu(ctx, C.byref(tx), pk.data_ptr(), bd.data_ptr(), C.byref(ty), C.byref(cp),
                                             ad.data_ptr() if use_prelu else None, st))
    torch.cuda.synchronize()
    assert relerr(y, ref) < BF16_TOL
    # the same store into the leading channels of a wider buffer: the pixel pitch is only 16-byte aligned, so the 32-byte stores of
    # the dense case do not apply; the neighbouring channels stay untouched
    cq = cout // 4
    wide = torch.full((N, 2 * H, 2 * W, cq + 24), 7.0, device="cuda", dtype=torch.bfloat16)
    tw = L.tensor(wide, c=cq, coff=0)
    L.check(lib.dg_umma_conv2d_fwd_d2s_prelu(ctx, C.byref(tx), pk.data_ptr(), bd.data_ptr(), C.byref(tw), C.byref(cp),
                                             ad.data_ptr() if use_prelu else None, st))
    torch.cuda.synchronize()
    assert torch.equal(wide[..., :cq], y)
    assert (wide[..., cq:] == 7.0).all()


@pytest.mark.parametrize("ws", ["1", "0"])
@pytest.mark.parametrize("N,H,W", [(1, 8, 16), (2, 21, 37), (1, 64, 160), (1, 3, 5)])
def test_fsrgan_block_infer_one_launch(L, N, H, W, ws, monkeypatch):
    """dg_fsrgan_block_infer: expand 1x1 (+ReLU) -> depthwise 3x3 (+ReLU) -> project 1x1 -> + input of an inverted-residual block
    (fsrgan.py:112-176 with the inference-mode BatchNorms folded into kernels / biases) in one launch, against the float64 oracle
    with the kernel's storage roundings (the two 192-channel intermediates are fp16, saturating).  Covers a single
    tile, partial tiles in both directions with several images, many tiles per CTA, and a map smaller than one tile; the input is
    a channel slice of a wider tensor.  SAME padding of the depthwise convolution must see ZEROS outside the image, not
    relu(bias) of the expand layer: biases are large enough for the border to show it.  Both kernels: the warp-specialised one
    (16 x 4 tiles, the default) and the lock-step one (16 x 8 tiles, DG_FSRGAN_BLOCK_WS=0)."""
    monkeypatch.setenv("DG_FSRGAN_BLOCK_WS", ws)
    g = torch.Generator().manual_seed(H * 131 + W)
    bf = lambda t: t.bfloat16().double()
    x = bf(torch.randn(N, H, W, 32, generator=g))
    w1 = bf(torch.randn(1, 1, 32, 192, generator=g) * 0.2); b1 = torch.randn(192, generator=g).double() * 0.5
    wd = torch.randn(3, 3, 192, 1, generator=g).double() * 0.3; bd = torch.randn(192, generator=g).double() * 0.3
    w2 = bf(torch.randn(1, 1, 192, 32, generator=g) * 0.1); b2 = torch.randn(32, generator=g).double() * 0.2
    h16 = lambda t: t.clamp(-65504.0, 65504.0).half().double()
    e = h16(torch.relu(OT.conv2d(x, w1, b1, stride=1, padding="same")))
    d = h16(torch.relu(OT.depthwise_conv2d(e, h16(wd), h16(bd))))
    ref = OT.conv2d(d, w2, b2, stride=1, padding="same") + x
    ctx = L.ctx(0); lib = L.load(); st = L.stream_ptr()
    wide = torch.zeros(N, H, W, 48, device="cuda", dtype=torch.bfloat16); wide[..., 8:40] = x.to(torch.bfloat16).cuda()
    y = torch.full((N, H, W, 40), 3.0, device="cuda", dtype=torch.bfloat16)
    tx, ty = L.tensor(wide, c=32, coff=8), L.tensor(y, c=32, coff=0)
    assert lib.dg_fsrgan_block_infer_supported(ctx, C.byref(tx), C.byref(ty)) == 1
    w1d = w1.view(32, 192).t().contiguous().to(torch.bfloat16).cuda(); w2d = w2.view(192, 32).t().contiguous().to(torch.float16).cuda()
    b1d, wdd, bdd, b2d = dev(b1), dev(wd.view(9, 192)), dev(bd), dev(b2)
    L.check(lib.dg_fsrgan_block_infer(ctx, C.byref(tx), w1d.data_ptr(), b1d.data_ptr(), wdd.data_ptr(), bdd.data_ptr(), w2d.data_ptr(),
                                      b2d.data_ptr(), C.byref(ty), st))
    torch.cuda.synchronize()
    out = y[..., :32].double().cpu()
    assert (y[..., 32:] == 3.0).all()                 # nothing written outside the channel slice
    # the nine depthwise taps are accumulated in fp16 (<= 9 x 2^-12 of the partial sum) and an intermediate on a rounding boundary
    # may round the other way; the bound is relative to the tensor's magnitude as everywhere on the bf16 path
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < BF16_TOL, err
    assert ((out - ref).abs() <= 2.0 ** -6 * ref.abs() + 0.05).float().mean().item() > 0.999
