"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's training-pair synthesis (dataloader.py:188-229):

    stack_crop (dataloader.py:79-93)  ->  scale_image: tf.image.resize(..., method='bicubic') (dataloader.py:110-124)
    ->  adjust_jpeg_quality: tf.image.adjust_jpeg_quality (dataloader.py:126-140)  ->  normalize: x*2-1 (dataloader.py:160-178)

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.

The JPEG degradation is NOT in /root/reference: tf.image.adjust_jpeg_quality = convert_image_dtype(uint8, saturate) ->
encode_jpeg(quality, chroma_downsampling=True) -> decode_jpeg(fancy_upscaling=True, dct_method='' = JDCT_ISLOW) -> float, i.e. a
round trip through libjpeg(-turbo).  This file restates libjpeg's published baseline algorithm in integer arithmetic: RGB ->
YCbCr (jccolor.c), h2v2 chroma down-sampling (jcsample.c), the 'islow' forward DCT (jfdctint.c) and quantisation with the
quality-scaled Annex-K tables (jcparam.c, jcdctmgr.c), dequantisation + 'islow' inverse DCT (jidctint.c), h2v2 'fancy' (triangle)
chroma up-sampling (jdsample.c) and YCbCr -> RGB (jdcolor.c).  The entropy coding is lossless and is skipped.
PINNED: tests/test_pairs_cpu.py checks jpeg_roundtrip_u8 bit for bit against a real libjpeg-turbo encode/decode (Pillow, 4:2:0,
same quality) in this container and against fixtures in tests/golden/.

The bicubic resize follows TensorFlow's ResizeBicubic kernel with half_pixel_centers=True (tf.image.resize in TF2): Keys cubic
a = -0.5 read from a 1024-entry table, taps outside the image get weight 0 and the rest is renormalised, rows are interpolated
along x first.  TensorFlow is not installable here: the resize is pinned only on its analytic weights ("parity unpinned" for
that step).
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------------------------- JPEG tables
_LUMA = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99],
    dtype=np.int64).reshape(8, 8)
_CHROMA = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99],
    dtype=np.int64).reshape(8, 8)


def quant_tables(quality: int):
    """jpeg_set_quality(quality, force_baseline=TRUE): (luma, chroma) 8x8 int tables."""
    q = min(max(int(quality), 1), 100)
    scale = 5000 // q if q < 50 else 200 - 2 * q
    out = []
    for base in (_LUMA, _CHROMA):
        t = (base * scale + 50) // 100
        out.append(np.clip(t, 1, 255))
    return out[0], out[1]


_F = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069, f2053=16819,
          f2562=20995, f3072=25172)
CONST_BITS, PASS1_BITS = 13, 2


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct_1d(d, first_pass: bool):
    """One pass of jfdctint.c over the LAST axis of d (int64 [..., 8])."""
    F = _F
    d0, d1, d2, d3, d4, d5, d6, d7 = [d[..., i] for i in range(8)]
    tmp0, tmp7 = d0 + d7, d0 - d7
    tmp1, tmp6 = d1 + d6, d1 - d6
    tmp2, tmp5 = d2 + d5, d2 - d5
    tmp3, tmp4 = d3 + d4, d3 - d4
    tmp10, tmp13 = tmp0 + tmp3, tmp0 - tmp3
    tmp11, tmp12 = tmp1 + tmp2, tmp1 - tmp2
    out = [None] * 8
    if first_pass:
        out[0] = (tmp10 + tmp11) << PASS1_BITS
        out[4] = (tmp10 - tmp11) << PASS1_BITS
        sh = CONST_BITS - PASS1_BITS
    else:
        out[0] = _descale(tmp10 + tmp11, PASS1_BITS)
        out[4] = _descale(tmp10 - tmp11, PASS1_BITS)
        sh = CONST_BITS + PASS1_BITS
    z1 = (tmp12 + tmp13) * F["f0541"]
    out[2] = _descale(z1 + tmp13 * F["f0765"], sh)
    out[6] = _descale(z1 + tmp12 * (-F["f1847"]), sh)
    z1, z2, z3, z4 = tmp4 + tmp7, tmp5 + tmp6, tmp4 + tmp6, tmp5 + tmp7
    z5 = (z3 + z4) * F["f1175"]
    tmp4, tmp5, tmp6, tmp7 = tmp4 * F["f0298"], tmp5 * F["f2053"], tmp6 * F["f3072"], tmp7 * F["f1501"]
    z1, z2, z3, z4 = z1 * (-F["f0899"]), z2 * (-F["f2562"]), z3 * (-F["f1961"]), z4 * (-F["f0390"])
    z3, z4 = z3 + z5, z4 + z5
    out[7] = _descale(tmp4 + z1 + z3, sh)
    out[5] = _descale(tmp5 + z2 + z4, sh)
    out[3] = _descale(tmp6 + z2 + z3, sh)
    out[1] = _descale(tmp7 + z1 + z4, sh)
    return np.stack(out, axis=-1)


def fdct_islow(blocks):
    """jpeg_fdct_islow on int64 [..., 8(row), 8(col)] level-shifted samples; output scaled by 8."""
    t = _fdct_1d(blocks, True)                                     # rows
    return _fdct_1d(t.swapaxes(-1, -2), False).swapaxes(-1, -2)    # columns


def quantize(coef, table):
    """jcdctmgr.c: round to nearest, ties away from zero, divisor = table << 3."""
    q = table << 3
    a = np.abs(coef) + (q >> 1)
    r = np.where(a >= q, a // q, 0)
    return np.where(coef < 0, -r, r)


def _idct_1d(d, first_pass: bool):
    """One pass of jidctint.c over the LAST axis (int64 [..., 8])."""
    F = _F
    in0, in1, in2, in3, in4, in5, in6, in7 = [d[..., i] for i in range(8)]
    z2, z3 = in2, in6
    z1 = (z2 + z3) * F["f0541"]
    tmp2 = z1 + z3 * (-F["f1847"])
    tmp3 = z1 + z2 * F["f0765"]
    tmp0 = (in0 + in4) << CONST_BITS
    tmp1 = (in0 - in4) << CONST_BITS
    tmp10, tmp13 = tmp0 + tmp3, tmp0 - tmp3
    tmp11, tmp12 = tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = in7, in5, in3, in1
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F["f1175"]
    t0, t1, t2, t3 = t0 * F["f0298"], t1 * F["f2053"], t2 * F["f3072"], t3 * F["f1501"]
    z1, z2, z3, z4 = z1 * (-F["f0899"]), z2 * (-F["f2562"]), z3 * (-F["f1961"]), z4 * (-F["f0390"])
    z3, z4 = z3 + z5, z4 + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    sh = CONST_BITS - PASS1_BITS if first_pass else CONST_BITS + PASS1_BITS + 3
    out = [_descale(tmp10 + t3, sh), _descale(tmp11 + t2, sh), _descale(tmp12 + t1, sh), _descale(tmp13 + t0, sh),
           _descale(tmp13 - t0, sh), _descale(tmp12 - t1, sh), _descale(tmp11 - t2, sh), _descale(tmp10 - t3, sh)]
    return np.stack(out, axis=-1)


def idct_islow(coef):
    """jpeg_idct_islow on DEQUANTISED int64 [..., 8(row), 8(col)] coefficients -> samples 0..255."""
    t = _idct_1d(coef.swapaxes(-1, -2), True).swapaxes(-1, -2)     # pass 1: columns
    r = _idct_1d(t, False)                                         # pass 2: rows
    return np.clip(r + 128, 0, 255)


def _blocks(plane):
    """[H, W] -> [H/8, W/8, 8, 8]."""
    H, W = plane.shape
    return plane.reshape(H // 8, 8, W // 8, 8).transpose(0, 2, 1, 3)


def _unblocks(b):
    nh, nw = b.shape[:2]
    return b.transpose(0, 2, 1, 3).reshape(nh * 8, nw * 8)


def _fix(x):
    return int(x * 65536 + 0.5)


def rgb_to_ycc(rgb):
    """jccolor.c rgb_ycc_convert on int64 [H, W, 3] -> (Y, Cb, Cr)."""
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    half, off = 1 << 15, 128 << 16
    y = (_fix(0.29900) * r + _fix(0.58700) * g + _fix(0.11400) * b + half) >> 16
    cb = (-_fix(0.16874) * r - _fix(0.33126) * g + _fix(0.50000) * b + off + half - 1) >> 16
    cr = (_fix(0.50000) * r - _fix(0.41869) * g - _fix(0.08131) * b + off + half - 1) >> 16
    return y, cb, cr


def h2v2_downsample(p):
    """jcsample.c h2v2_downsample: 2x2 box with the alternating 1, 2 rounding bias along a row."""
    s = p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]
    bias = np.where(np.arange(s.shape[1]) % 2 == 0, 1, 2)[None, :]
    return (s + bias) >> 2


def h2v2_fancy_upsample(p):
    """jdsample.c h2v2_fancy_upsample: triangle filter, 3/4 nearer + 1/4 further in each direction; rows / columns beyond the
    image edge are the edge itself."""
    h, w = p.shape
    up = np.vstack([p[:1], p[:-1]])          # row above (edge replicated)
    dn = np.vstack([p[1:], p[-1:]])          # row below
    out = np.empty((2 * h, 2 * w), dtype=np.int64)
    for v, other in ((0, up), (1, dn)):
        cs = 3 * p + other                   # column sums of the two contributing rows
        last = np.hstack([cs[:, :1], cs[:, :-1]])
        nxt = np.hstack([cs[:, 1:], cs[:, -1:]])
        out[v::2, 0::2] = (3 * cs + last + 8) >> 4
        out[v::2, 1::2] = (3 * cs + nxt + 7) >> 4
    return out


def ycc_to_rgb(y, cb, cr):
    """jdcolor.c ycc_rgb_convert."""
    half = 1 << 15
    xb, xr = cb - 128, cr - 128
    r = y + ((_fix(1.40200) * xr + half) >> 16)
    g = y + ((-_fix(0.34414) * xb + half - _fix(0.71414) * xr) >> 16)
    b = y + ((_fix(1.77200) * xb + half) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255)


def jpeg_roundtrip_u8(img: np.ndarray, quality: int) -> np.ndarray:
    """uint8 [H, W, 3] (H, W multiples of 16) -> the image after a baseline 4:2:0 JPEG encode/decode at `quality`."""
    H, W, _ = img.shape
    assert H % 16 == 0 and W % 16 == 0, "image sides must be multiples of the 16x16 MCU"
    ql, qc = quant_tables(quality)
    y, cb, cr = rgb_to_ycc(img.astype(np.int64))
    planes = []
    for p, q in ((y, ql), (h2v2_downsample(cb), qc), (h2v2_downsample(cr), qc)):
        coef = quantize(fdct_islow(_blocks(p - 128)), q)
        planes.append(_unblocks(idct_islow(coef * q)))
    return ycc_to_rgb(planes[0], h2v2_fancy_upsample(planes[1]), h2v2_fancy_upsample(planes[2])).astype(np.uint8)


def to_u8_saturate(x: np.ndarray) -> np.ndarray:
    """tf.image.convert_image_dtype(float32 -> uint8, saturate=True): saturate_cast<uint8>(x * 255.5) (truncation)."""
    v = x.astype(np.float32) * np.float32(255.5)
    return np.clip(v, 0, 255).astype(np.uint8)


def adjust_jpeg_quality(x: np.ndarray, quality: int) -> np.ndarray:
    """tf.image.adjust_jpeg_quality on float32 [H, W, 3] in [0, 1] (dataloader.py:138)."""
    u = jpeg_roundtrip_u8(to_u8_saturate(x), quality)
    return u.astype(np.float32) * np.float32(1.0 / 255.0)


# ---------------------------------------------------------------------------------------------- bicubic resize
TABLE = 1024


def _coeffs_table():
    a = np.float32(-0.5)
    t = np.zeros((TABLE + 1) * 2, dtype=np.float32)
    for i in range(TABLE + 1):
        x = np.float32(i) / np.float32(TABLE)
        t[2 * i] = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + np.float32(1)
        x = x + np.float32(1)
        t[2 * i + 1] = ((a * x - np.float32(5) * a) * x + np.float32(8) * a) * x - np.float32(4) * a
    return t


_TABLE = _coeffs_table()


def bicubic_weights(out_size: int, in_size: int):
    """TensorFlow GetWeightsAndIndices<HalfPixelScaler, use_keys_cubic=true>: per output coordinate four (index, weight)."""
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.zeros((out_size, 4), dtype=np.int64)
    wts = np.zeros((out_size, 4), dtype=np.float32)
    for o in range(out_size):
        in_loc_f = (np.float32(o) + np.float32(0.5)) * scale - np.float32(0.5)
        in_loc = int(np.floor(in_loc_f))
        delta = in_loc_f - np.float32(in_loc)
        off = int(np.rint(delta * np.float32(TABLE)))
        cand = [in_loc - 1, in_loc, in_loc + 1, in_loc + 2]
        w = [_TABLE[off * 2 + 1], _TABLE[off * 2], _TABLE[(TABLE - off) * 2], _TABLE[(TABLE - off) * 2 + 1]]
        for k in range(4):
            b = min(max(cand[k], 0), in_size - 1)
            idx[o, k] = b
            wts[o, k] = w[k] if b == cand[k] else np.float32(0)
        s = np.float32(wts[o, 0] + wts[o, 1]) + np.float32(wts[o, 2] + wts[o, 3])       # (w0 + w1) + (w2 + w3) in float32
        if abs(s) >= 1000.0 * np.finfo(np.float32).tiny:
            wts[o] = wts[o] * (np.float32(1) / s)
    return idx, wts


def bicubic_resize(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """tf.image.resize(x, [out_h, out_w], method='bicubic') on float32 [H, W, C] (dataloader.py:120-122), float32 arithmetic in
    TensorFlow's order: each of the four rows is interpolated along x (((v0 w0 + v1 w1) + v2 w2) + v3 w3), then the four results
    along y in the same way."""
    H, W, _ = x.shape
    x = x.astype(np.float32)
    iy, wy = bicubic_weights(out_h, H)
    ix, wx = bicubic_weights(out_w, W)

    def interp(v0, v1, v2, v3, w):
        return ((v0 * w[..., 0:1] + v1 * w[..., 1:2]) + v2 * w[..., 2:3]) + v3 * w[..., 3:4]
    rows = x[:, ix]                                   # [H, out_w, 4, C]
    hx = interp(rows[:, :, 0], rows[:, :, 1], rows[:, :, 2], rows[:, :, 3], wx[None, :, :])      # [H, out_w, C]
    cols = hx[iy]                                     # [out_h, 4, out_w, C]
    return interp(cols[:, 0], cols[:, 1], cols[:, 2], cols[:, 3], wy[:, None, :]).astype(np.float32)


# ---------------------------------------------------------------------------------------------- the pair
def synth_pair(src_u8: np.ndarray, top: int, left: int, crop: int, scale: int, quality: int):
    """One training pair from a decoded uint8 image [H, W, 3] (dataloader.py:205-219 after load_image): target = the
    crop, input = adjust_jpeg_quality(bicubic(crop, crop/scale)), both normalised to [-1, 1]."""
    hr = src_u8[top:top + crop, left:left + crop].astype(np.float32) * np.float32(1.0 / 255.0)   # convert_image_dtype(uint8 -> float32)
    lr = hr if scale == 1 else bicubic_resize(hr, crop // scale, crop // scale)
    lr = adjust_jpeg_quality(lr, quality)
    return lr * np.float32(2) - np.float32(1), hr * np.float32(2) - np.float32(1)
