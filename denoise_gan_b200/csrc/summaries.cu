// Image summaries of the reference's training loop (train_srgan.py:27-59, 153-172): every `log_iter` iterations the FIRST image of
// the batch is shown as renorm(x), and the error / gradient diagnostics (square and absolute error, Sobel magnitude, horizontal /
// vertical differences, total variation) are shown auto-scaled to their own range, all as uint8.  One image per call, off the
// hot path; two small launches (values + per-block min / max, then scale + convert), float operations explicitly rounded in the
// order of oracle/summaries.py so that the uint8 result is bit-exact.
#include <float.h>
#include <stdint.h>

#include "dg_common.cuh"

namespace {

enum { K_IMAGE = 0, K_SQUARE, K_ABS, K_SOBEL, K_DX, K_DY, K_TV };

struct SView {
  const void* p;
  int f32, pitch, off;
};

__device__ __forceinline__ float ld(const SView& v, long pix, int c) {
  const long i = pix * v.pitch + v.off + c;
  return v.f32 ? reinterpret_cast<const float*>(v.p)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(v.p)[i]);
}
__device__ __forceinline__ float diff(const SView& a, const SView& b, int has_b, int W, int y, int x, int c) {
  const long pix = (long)y * W + x;
  const float va = ld(a, pix, c);
  return has_b ? __fsub_rn(va, ld(b, pix, c)) : va;
}
__device__ __forceinline__ float renorm(float v) {
  const float r = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);
  return fminf(fmaxf(r, 0.0f), 1.0f);
}
__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(256) summary_values_kernel(SView a, SView b, int has_b, int kind, int H, int W, int C, int oh, int ow,
                                                             float* __restrict__ vals, float* __restrict__ part) {
  __shared__ float smn[256], smx[256];
  const long total = (long)oh * ow * C;
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long p = i / C;
    const int x = (int)(p % ow), y = (int)(p / ow);
    float v;
    if (kind == K_IMAGE) v = renorm(diff(a, b, has_b, W, y, x, c));
    else if (kind == K_SQUARE) { const float d = diff(a, b, has_b, W, y, x, c); v = __fmul_rn(d, d); }
    else if (kind == K_ABS) v = fabsf(diff(a, b, has_b, W, y, x, c));
    else if (kind == K_SOBEL) {
      float t[3][3];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) t[dy][dx] = renorm(diff(a, b, has_b, W, reflect(y + dy - 1, H), reflect(x + dx - 1, W), c));
      float s0 = __fsub_rn(-t[0][0], __fmul_rn(2.0f, t[0][1]));
      s0 = __fsub_rn(s0, t[0][2]); s0 = __fadd_rn(s0, t[2][0]); s0 = __fadd_rn(s0, __fmul_rn(2.0f, t[2][1])); s0 = __fadd_rn(s0, t[2][2]);
      float s1 = __fadd_rn(-t[0][0], t[0][2]);
      s1 = __fsub_rn(s1, __fmul_rn(2.0f, t[1][0])); s1 = __fadd_rn(s1, __fmul_rn(2.0f, t[1][2])); s1 = __fsub_rn(s1, t[2][0]); s1 = __fadd_rn(s1, t[2][2]);
      const float g0 = __fmul_rn(s0, 0.25f), g1 = __fmul_rn(s1, 0.25f);
      v = __fsqrt_rn(__fadd_rn(__fmul_rn(g0, g0), __fmul_rn(g1, g1)));
    } else {
      const float v00 = diff(a, b, has_b, W, y, x, c);
      const float ddx = __fsub_rn(diff(a, b, has_b, W, y, x + 1, c), v00), ddy = __fsub_rn(diff(a, b, has_b, W, y + 1, x, c), v00);
      v = kind == K_DX ? ddx : (kind == K_DY ? ddy : __fadd_rn(fabsf(ddx), fabsf(ddy)));
    }
    vals[i] = v;
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + s]);
      smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = smn[0]; part[2 * blockIdx.x + 1] = smx[0]; }
}

__global__ void __launch_bounds__(256) summary_convert_kernel(const float* __restrict__ vals, const float* __restrict__ part, int nblocks,
                                                              int autoscale, long total, uint8_t* __restrict__ out) {
  __shared__ float smn[256], smx[256];
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { mn = fminf(mn, part[2 * i]); mx = fmaxf(mx, part[2 * i + 1]); }
  smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + s]);
      smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + s]);
    }
    __syncthreads();
  }
  mn = smn[0]; mx = smx[0];
  const float ptp = __fsub_rn(mx, mn);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float v = vals[i];
    if (autoscale) v = ptp > 0.f ? __fdiv_rn(__fsub_rn(v, mn), ptp) : 0.f;
    const float s = __fmul_rn(255.0f, v);
    out[i] = (uint8_t)(s <= 0.f ? 0 : (s >= 255.f ? 255 : (int)s));
  }
}

inline int sum_blocks(long total, int sm) {
  long b = (total + 255) / 256, cap = (long)sm * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" size_t dg_image_summary_workspace_bytes(int h, int w, int c) {
  return (size_t)h * w * c * sizeof(float) + (size_t)160 * 8 * 2 * sizeof(float) + 256;
}

extern "C" int dg_image_summary(dg_ctx* ctx, const dg_tensor* a, const dg_tensor* b, int kind, uint8_t* out, void* workspace,
                                size_t workspace_bytes, void* stream) {
  const char* name = "dg_image_summary";
  DG_REQUIRE(dg_valid(a) && out && workspace, "%s: null argument", name);
  DG_REQUIRE(kind >= K_IMAGE && kind <= K_TV, "%s: unknown summary kind %d", name, kind);
  DG_REQUIRE(!b || (dg_valid(b) && dg_same_shape(a, b)), "%s: shape mismatch", name);
  DG_REQUIRE(a->h >= 2 && a->w >= 2, "%s: image too small", name);
  DG_REQUIRE(workspace_bytes >= dg_image_summary_workspace_bytes(a->h, a->w, a->c), "%s: workspace too small", name);
  const int cropped = kind >= K_DX;
  const int oh = a->h - cropped, ow = a->w - cropped;
  const long total = (long)oh * ow * a->c;
  float* vals = (float*)workspace;
  float* part = vals + (((size_t)a->h * a->w * a->c + 63) & ~(size_t)63);
  const int nb = sum_blocks(total, ctx->sm_count < 160 ? ctx->sm_count : 160);
  // only the FIRST image of the batch is shown (tf2image: image[0]): the views address image 0
  SView va{a->ptr, a->dtype == DG_F32, a->cpitch, a->coff}, vb{b ? b->ptr : nullptr, b ? b->dtype == DG_F32 : 1, b ? b->cpitch : 0, b ? b->coff : 0};
  summary_values_kernel<<<nb, 256, 0, ST>>>(va, vb, b ? 1 : 0, kind, a->h, a->w, a->c, oh, ow, vals, part);
  DG_CHECK_LAUNCH(name);
  summary_convert_kernel<<<nb, 256, 0, ST>>>(vals, part, nb, kind != K_IMAGE, total, out);
  DG_CHECK_LAUNCH(name);
  return 0;
}
