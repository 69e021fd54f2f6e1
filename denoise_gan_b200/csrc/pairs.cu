// Training-pair synthesis on the device (dataloader.py:188-229 after load_image): random crop (stack_crop, :79-93), bicubic
// down-scaling (scale_image, :110-124 = tf.image.resize(method='bicubic'), half-pixel centres, Keys a = -0.5), JPEG degradation
// (adjust_jpeg_quality, :126-140 = tf.image.adjust_jpeg_quality, a libjpeg round trip) and normalisation to [-1, 1] (:160-178).
// At a 7 ms train step a host tf.data pipeline (decode + crop + resize + JPEG round trip per sample on CPU threads) is the
// bottleneck; here a batch is produced from decoded uint8 images that already sit in HBM by five small launches.
//
// The JPEG round trip restates libjpeg's baseline algorithm in the SAME integer arithmetic (oracle/pairs.py documents each step
// and is pinned bit for bit against libjpeg-turbo): RGB -> YCbCr, 2x2 chroma box filter, 'islow' forward DCT + quantisation with
// the quality-scaled Annex-K tables, dequantisation + 'islow' inverse DCT, triangle ('fancy') chroma up-sampling, YCbCr -> RGB.
// Entropy coding is lossless and skipped.  Float steps use explicitly rounded operations (no FMA contraction) in the order of
// the oracle, so the whole path is bit-exact against it.
//
// HBM-bound byte work: one thread per 8x8 block for the transforms (a block is 64 bytes in, 64 bytes out), one thread per pixel
// elsewhere; grids are sized by the element count.
#include <stdint.h>

#include "dg_common.cuh"

namespace {

constexpr int TABLE = 1024;

__constant__ int c_luma[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                               18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
__constant__ int c_chroma[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137, F1961 = 16069,
              F2053 = 16819, F2562 = 20995, F3072 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ constexpr int fix16(double x) { return (int)(x * 65536 + 0.5); }

// one 1-D pass of jfdctint.c over d[0], d[S], ..., d[7S]
template <int S, bool FIRST>
__device__ __forceinline__ void fdct8(int* d) {
  const int tmp0 = d[0] + d[7 * S], tmp7 = d[0] - d[7 * S], tmp1 = d[S] + d[6 * S], tmp6 = d[S] - d[6 * S];
  const int tmp2 = d[2 * S] + d[5 * S], tmp5 = d[2 * S] - d[5 * S], tmp3 = d[3 * S] + d[4 * S], tmp4 = d[3 * S] - d[4 * S];
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  constexpr int SH = FIRST ? CONST_BITS - PASS1_BITS : CONST_BITS + PASS1_BITS;
  if (FIRST) {
    d[0] = (tmp10 + tmp11) << PASS1_BITS;
    d[4 * S] = (tmp10 - tmp11) << PASS1_BITS;
  } else {
    d[0] = descale(tmp10 + tmp11, PASS1_BITS);
    d[4 * S] = descale(tmp10 - tmp11, PASS1_BITS);
  }
  int z1 = (tmp12 + tmp13) * F0541;
  d[2 * S] = descale(z1 + tmp13 * F0765, SH);
  d[6 * S] = descale(z1 + tmp12 * (-F1847), SH);
  z1 = tmp4 + tmp7;
  int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
  const int z5 = (z3 + z4) * F1175;
  const int t4 = tmp4 * F0298, t5 = tmp5 * F2053, t6 = tmp6 * F3072, t7 = tmp7 * F1501;
  z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
  z3 += z5; z4 += z5;
  d[7 * S] = descale(t4 + z1 + z3, SH);
  d[5 * S] = descale(t5 + z2 + z4, SH);
  d[3 * S] = descale(t6 + z2 + z3, SH);
  d[S] = descale(t7 + z1 + z4, SH);
}

// one 1-D pass of jidctint.c
template <int S, bool FIRST>
__device__ __forceinline__ void idct8(int* d) {
  int z2 = d[2 * S], z3 = d[6 * S];
  int z1 = (z2 + z3) * F0541;
  const int tmp2 = z1 + z3 * (-F1847), tmp3 = z1 + z2 * F0765;
  const int tmp0 = (d[0] + d[4 * S]) << CONST_BITS, tmp1 = (d[0] - d[4 * S]) << CONST_BITS;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int t0 = d[7 * S], t1 = d[5 * S], t2 = d[3 * S], t3 = d[S];
  z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2;
  int z4 = t1 + t3;
  const int z5 = (z3 + z4) * F1175;
  t0 *= F0298; t1 *= F2053; t2 *= F3072; t3 *= F1501;
  z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
  z3 += z5; z4 += z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  constexpr int SH = FIRST ? CONST_BITS - PASS1_BITS : CONST_BITS + PASS1_BITS + 3;
  d[0] = descale(tmp10 + t3, SH); d[7 * S] = descale(tmp10 - t3, SH);
  d[S] = descale(tmp11 + t2, SH); d[6 * S] = descale(tmp11 - t2, SH);
  d[2 * S] = descale(tmp12 + t1, SH); d[5 * S] = descale(tmp12 - t1, SH);
  d[3 * S] = descale(tmp13 + t0, SH); d[4 * S] = descale(tmp13 - t0, SH);
}

__device__ __forceinline__ int quant_entry(const int* base, int k, int scale) {
  const int t = (base[k] * scale + 50) / 100;
  return t < 1 ? 1 : (t > 255 ? 255 : t);
}

// forward DCT, quantise, dequantise, inverse DCT of one 8x8 block held in d[64] (row-major, level-shifted samples)
__device__ __forceinline__ void block_roundtrip(int* d, const int* base, int qscale) {
#pragma unroll
  for (int r = 0; r < 8; ++r) fdct8<1, true>(d + 8 * r);
#pragma unroll
  for (int c = 0; c < 8; ++c) fdct8<8, false>(d + c);
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    const int qt = quant_entry(base, k, qscale), q = qt << 3;
    int a = d[k] < 0 ? -d[k] : d[k];
    a += q >> 1;
    a = a >= q ? a / q : 0;
    d[k] = (d[k] < 0 ? -a : a) * qt;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) idct8<8, true>(d + c);
#pragma unroll
  for (int r = 0; r < 8; ++r) idct8<1, false>(d + 8 * r);
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    const int v = d[k] + 128;
    d[k] = v < 0 ? 0 : (v > 255 ? 255 : v);
  }
}

// tf.image.convert_image_dtype(float -> uint8, saturate=True): saturate_cast<uint8>(x * 255.5)
__device__ __forceinline__ int to_u8(float x) {
  const float v = __fmul_rn(x, 255.5f);
  return v <= 0.f ? 0 : (v >= 255.f ? 255 : (int)v);
}

// ------------------------------------------------------------------ crop + normalise
// target[b,i,j,c] = (src/255)*2 - 1; with scale == 1 the degraded input starts from the same crop
__global__ void crop_kernel(const uint8_t* __restrict__ src, int src_h, int src_w, const int* __restrict__ idx, const int* __restrict__ top,
                            const int* __restrict__ left, int batch, int crop, float* __restrict__ target, uint8_t* __restrict__ lr_u8) {
  const long total = (long)batch * crop * crop * 3;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    long p = i / 3;
    const int x = (int)(p % crop);
    p /= crop;
    const int y = (int)(p % crop), b = (int)(p / crop);
    const uint8_t s = src[(((long)idx[b] * src_h + top[b] + y) * src_w + left[b] + x) * 3 + c];
    const float f = __fmul_rn((float)s, 1.0f / 255.0f);
    target[i] = __fadd_rn(__fmul_rn(f, 2.0f), -1.0f);
    if (lr_u8) lr_u8[i] = (uint8_t)to_u8(f);
  }
}

// ------------------------------------------------------------------ bicubic weights (TensorFlow GetWeightsAndIndices, Keys a = -0.5)
__device__ __forceinline__ float keys_inner(float x) {     // ((a + 2) x - (a + 3)) x x + 1
  return __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(1.5f, x), -2.5f), x), x), 1.0f);
}
__device__ __forceinline__ float keys_outer(float x) {     // ((a x - 5a) x + 8a) x - 4a, x in [1, 2]
  return __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(-0.5f, x), 2.5f), x), -4.0f), x), 2.0f);
}
__global__ void bicubic_weights_kernel(int out_size, int in_size, int* __restrict__ idx, float* __restrict__ wts) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out_size) return;
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  const float in_f = __fadd_rn(__fmul_rn(__fadd_rn((float)o, 0.5f), scale), -0.5f);
  const int in_loc = (int)floorf(in_f);
  const float delta = __fadd_rn(in_f, -(float)in_loc);
  const int off = (int)rintf(__fmul_rn(delta, (float)TABLE));
  const float x0 = __fdiv_rn((float)off, (float)TABLE), x1 = __fdiv_rn((float)(TABLE - off), (float)TABLE);
  float w[4] = {keys_outer(__fadd_rn(x0, 1.0f)), keys_inner(x0), keys_inner(x1), keys_outer(__fadd_rn(x1, 1.0f))};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cand = in_loc - 1 + k;
    const int b = cand < 0 ? 0 : (cand > in_size - 1 ? in_size - 1 : cand);
    idx[o * 4 + k] = b;
    if (b != cand) w[k] = 0.f;
  }
  const float s = __fadd_rn(__fadd_rn(w[0], w[1]), __fadd_rn(w[2], w[3]));
  if (fabsf(s) >= 1000.0f * 1.17549435e-38f) {
    const float inv = __fdiv_rn(1.0f, s);
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = __fmul_rn(w[k], inv);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) wts[o * 4 + k] = w[k];
}

__device__ __forceinline__ float interp4(float v0, float v1, float v2, float v3, const float* w) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0, w[0]), __fmul_rn(v1, w[1])), __fmul_rn(v2, w[2])), __fmul_rn(v3, w[3]));
}

// lr_u8[b,oy,ox,c] = to_u8(bicubic(src crop / 255)): rows interpolated along x first, then along y
__global__ void bicubic_kernel(const uint8_t* __restrict__ src, int src_h, int src_w, const int* __restrict__ idx, const int* __restrict__ top,
                               const int* __restrict__ left, int batch, int lr, const int* __restrict__ widx, const float* __restrict__ wts,
                               uint8_t* __restrict__ lr_u8) {
  const long total = (long)batch * lr * lr * 3;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    long p = i / 3;
    const int ox = (int)(p % lr);
    p /= lr;
    const int oy = (int)(p % lr), b = (int)(p / lr);
    const uint8_t* base = src + (((long)idx[b] * src_h + top[b]) * src_w + left[b]) * 3 + c;
    float rowv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint8_t* row = base + (long)widx[oy * 4 + r] * src_w * 3;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __fmul_rn((float)row[widx[ox * 4 + k] * 3], 1.0f / 255.0f);
      rowv[r] = interp4(v[0], v[1], v[2], v[3], wts + ox * 4);
    }
    lr_u8[i] = (uint8_t)to_u8(interp4(rowv[0], rowv[1], rowv[2], rowv[3], wts + oy * 4));
  }
}

// ------------------------------------------------------------------ JPEG: one thread per 8x8 block of Y, Cb or Cr
// planes: Y [batch, n, n], then Cb and Cr [batch, n/2, n/2] (decoded samples)
__global__ void jpeg_blocks_kernel(const uint8_t* __restrict__ rgb, int batch, int n, int quality, uint8_t* __restrict__ yp,
                                   uint8_t* __restrict__ cbp, uint8_t* __restrict__ crp) {
  const int by_n = n / 8, bc_n = n / 16;
  const int per_img = by_n * by_n + 2 * bc_n * bc_n;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * per_img) return;
  const int b = t / per_img;
  int k = t - b * per_img;
  const int q = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
  const int qscale = q < 50 ? 5000 / q : 200 - 2 * q;
  const uint8_t* img = rgb + (long)b * n * n * 3;
  int d[64];
  constexpr int HALF = 1 << 15, OFF = 128 << 16;
  if (k < by_n * by_n) {
    const int y0 = (k / by_n) * 8, x0 = (k % by_n) * 8;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const uint8_t* px = img + ((long)(y0 + (i >> 3)) * n + x0 + (i & 7)) * 3;
      d[i] = ((fix16(0.29900) * px[0] + fix16(0.58700) * px[1] + fix16(0.11400) * px[2] + HALF) >> 16) - 128;
    }
    block_roundtrip(d, c_luma, qscale);
    uint8_t* o = yp + (long)b * n * n;
#pragma unroll
    for (int i = 0; i < 64; ++i) o[(long)(y0 + (i >> 3)) * n + x0 + (i & 7)] = (uint8_t)d[i];
  } else {
    k -= by_n * by_n;
    const int which = k / (bc_n * bc_n);       // 0: Cb, 1: Cr
    k -= which * bc_n * bc_n;
    const int y0 = (k / bc_n) * 8, x0 = (k % bc_n) * 8;      // in the half-resolution plane
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const int cy = y0 + (i >> 3), cx = x0 + (i & 7);
      int s = 0;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const uint8_t* px = img + ((long)(2 * cy + dy) * n + 2 * cx + dx) * 3;
          const int r = px[0], g = px[1], bl = px[2];
          s += which == 0 ? ((-fix16(0.16874) * r - fix16(0.33126) * g + fix16(0.50000) * bl + OFF + HALF - 1) >> 16)
                          : ((fix16(0.50000) * r - fix16(0.41869) * g - fix16(0.08131) * bl + OFF + HALF - 1) >> 16);
        }
      d[i] = ((s + ((cx & 1) ? 2 : 1)) >> 2) - 128;
    }
    block_roundtrip(d, c_chroma, qscale);
    const int hn = n / 2;
    uint8_t* o = (which == 0 ? cbp : crp) + (long)b * hn * hn;
#pragma unroll
    for (int i = 0; i < 64; ++i) o[(long)(y0 + (i >> 3)) * hn + x0 + (i & 7)] = (uint8_t)d[i];
  }
}

__device__ __forceinline__ int fancy_up(const uint8_t* __restrict__ p, int hn, int y, int x) {
  const int cy = y >> 1, cx = x >> 1;
  int oy = (y & 1) ? cy + 1 : cy - 1;
  oy = oy < 0 ? 0 : (oy > hn - 1 ? hn - 1 : oy);
  int ox = (x & 1) ? cx + 1 : cx - 1;
  ox = ox < 0 ? 0 : (ox > hn - 1 ? hn - 1 : ox);
  const int cs = 3 * p[cy * hn + cx] + p[oy * hn + cx];
  const int co = 3 * p[cy * hn + ox] + p[oy * hn + ox];
  return (3 * cs + co + ((x & 1) ? 7 : 8)) >> 4;
}

// input[b,y,x,:] = (decode(Y, up(Cb), up(Cr)) / 255) * 2 - 1
__global__ void jpeg_decode_kernel(const uint8_t* __restrict__ yp, const uint8_t* __restrict__ cbp, const uint8_t* __restrict__ crp, int batch,
                                   int n, float* __restrict__ out) {
  const long total = (long)batch * n * n;
  const int hn = n / 2;
  constexpr int HALF = 1 << 15;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % n);
    long p = i / n;
    const int y = (int)(p % n), b = (int)(p / n);
    const int Y = yp[i];
    const int xb = fancy_up(cbp + (long)b * hn * hn, hn, y, x) - 128, xr = fancy_up(crp + (long)b * hn * hn, hn, y, x) - 128;
    int r = Y + ((fix16(1.40200) * xr + HALF) >> 16);
    int g = Y + ((-fix16(0.34414) * xb + HALF - fix16(0.71414) * xr) >> 16);
    int bl = Y + ((fix16(1.77200) * xb + HALF) >> 16);
    r = r < 0 ? 0 : (r > 255 ? 255 : r); g = g < 0 ? 0 : (g > 255 ? 255 : g); bl = bl < 0 ? 0 : (bl > 255 ? 255 : bl);
    float* o = out + i * 3;
    o[0] = __fadd_rn(__fmul_rn(__fmul_rn((float)r, 1.0f / 255.0f), 2.0f), -1.0f);
    o[1] = __fadd_rn(__fmul_rn(__fmul_rn((float)g, 1.0f / 255.0f), 2.0f), -1.0f);
    o[2] = __fadd_rn(__fmul_rn(__fmul_rn((float)bl, 1.0f / 255.0f), 2.0f), -1.0f);
  }
}

inline unsigned blocks_for(long total, int sm) {
  long b = (total + 255) / 256, cap = (long)sm * 32;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" size_t dg_pair_synthesis_workspace_bytes(int batch, int crop, int scale) {
  if (batch <= 0 || crop <= 0 || scale <= 0) return 0;
  const size_t lr = (size_t)(crop / scale);
  // degraded-input uint8 image | decoded Y | decoded Cb | decoded Cr | bicubic indices | bicubic weights
  return align256((size_t)batch * lr * lr * 3) + align256((size_t)batch * lr * lr) + 2 * align256((size_t)batch * (lr / 2) * (lr / 2)) +
         align256(lr * 4 * sizeof(int)) + align256(lr * 4 * sizeof(float));
}

extern "C" int dg_pair_synthesis(dg_ctx* ctx, const uint8_t* src, int n_src, int src_h, int src_w, const int* crop_index, const int* crop_top,
                                 const int* crop_left, int batch, int crop, int scale, int jpeg_quality, float* input, float* target,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  const char* name = "dg_pair_synthesis";
  DG_REQUIRE(src && crop_index && crop_top && crop_left && input && target && workspace, "%s: null argument", name);
  DG_REQUIRE(n_src > 0 && batch > 0 && scale >= 1 && crop > 0 && crop <= src_h && crop <= src_w && crop % scale == 0,
             "%s: bad geometry (crop %d of %dx%d, scale %d)", name, crop, src_h, src_w, scale);
  const int lr = crop / scale;
  DG_REQUIRE(lr % 16 == 0, "%s: the degraded image (%d px) must be a multiple of the 16x16 JPEG MCU", name, lr);
  DG_REQUIRE(jpeg_quality >= 1 && jpeg_quality <= 100, "%s: jpeg_quality must be in [1, 100]", name);
  DG_REQUIRE(workspace_bytes >= dg_pair_synthesis_workspace_bytes(batch, crop, scale), "%s: workspace too small", name);
  uint8_t* w = (uint8_t*)workspace;
  uint8_t* lr_u8 = w; w += align256((size_t)batch * lr * lr * 3);
  uint8_t* yp = w; w += align256((size_t)batch * lr * lr);
  uint8_t* cbp = w; w += align256((size_t)batch * (lr / 2) * (lr / 2));
  uint8_t* crp = w; w += align256((size_t)batch * (lr / 2) * (lr / 2));
  int* widx = (int*)w; w += align256((size_t)lr * 4 * sizeof(int));
  float* wts = (float*)w;
  const int sm = ctx->sm_count;
  crop_kernel<<<blocks_for((long)batch * crop * crop * 3, sm), 256, 0, ST>>>(src, src_h, src_w, crop_index, crop_top, crop_left, batch, crop,
                                                                           target, scale == 1 ? lr_u8 : nullptr);
  DG_CHECK_LAUNCH(name);
  if (scale > 1) {
    bicubic_weights_kernel<<<(lr + 127) / 128, 128, 0, ST>>>(lr, crop, widx, wts);
    bicubic_kernel<<<blocks_for((long)batch * lr * lr * 3, sm), 256, 0, ST>>>(src, src_h, src_w, crop_index, crop_top, crop_left, batch, lr, widx,
                                                                            wts, lr_u8);
    DG_CHECK_LAUNCH(name);
  }
  const int per_img = (lr / 8) * (lr / 8) + 2 * (lr / 16) * (lr / 16);
  jpeg_blocks_kernel<<<(batch * per_img + 63) / 64, 64, 0, ST>>>(lr_u8, batch, lr, jpeg_quality, yp, cbp, crp);
  DG_CHECK_LAUNCH(name);
  jpeg_decode_kernel<<<blocks_for((long)batch * lr * lr, sm), 256, 0, ST>>>(yp, cbp, crp, batch, lr, input);
  DG_CHECK_LAUNCH(name);
  return 0;
}
