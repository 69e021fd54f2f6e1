// Depthwise 3x3 stride-1 SAME convolution (forward, input gradient, weight gradient).
// HBM-bound (9 MAC per element): one thread per output element, channels fastest so a warp
// touches contiguous memory; the weight gradient is a per-channel reduction of 10 values.
// Reference: keras DepthwiseConv2D fsrgan.py:149-154, kernel layout [3,3,C,1].
#include "dg_common.cuh"
#include "reduce.cuh"

namespace {
using namespace dgred;

template <typename T, bool FLIP>
__global__ void dw3x3_kernel(const T* __restrict__ x, int xp, int xo, const float* __restrict__ w,
                             const float* __restrict__ bias, T* __restrict__ y, int yp, int yo, int N, int H, int W,
                             int C) {
  long total = (long)N * H * W * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long p = i / C;
    int wq = (int)(p % W);
    long t = p / W;
    int h = (int)(t % H);
    long n = t / H;
    float acc = bias ? bias[c] : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int hh = h + a - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        int ww = wq + b - 1;
        if (ww < 0 || ww >= W) continue;
        int wi = FLIP ? ((2 - a) * 3 + (2 - b)) : (a * 3 + b);
        acc += ld_f(x + (((n * H + hh) * W + ww) * xp + xo + c)) * w[wi * C + c];
      }
    }
    st_f(y + (p * yp + yo + c), acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
dw3x3_wgrad_kernel(const T* __restrict__ x, int xp, int xo, const T* __restrict__ dy, int dp, int dof, int N, int H,
                   int W, int C, float* __restrict__ partial) {
  channel_reduce<10>((long)N * H * W, C, partial, [&](long p, int c, float* acc) {
    int wq = (int)(p % W);
    long t = p / W;
    int h = (int)(t % H);
    long n = t / H;
    float g = ld_f(dy + (p * dp + dof + c));
    acc[9] += g;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int hh = h + a - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        int ww = wq + b - 1;
        if (ww < 0 || ww >= W) continue;
        acc[a * 3 + b] += g * ld_f(x + (((n * H + hh) * W + ww) * xp + xo + c));
      }
    }
  });
}

__global__ void dw3x3_wgrad_finalize(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ dw,
                                     float* __restrict__ dbias, int accumulate) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;  // e in [0, 10*C): [k][c]
  if (e >= 10 * C) return;
  double s = 0;
  for (int b = 0; b < nblocks; ++b) s += partial[(long)b * 10 * C + e];
  int k = e / C, c = e - k * C;
  if (k < 9) dw[k * C + c] = (accumulate ? dw[k * C + c] : 0.f) + (float)s;
  else if (dbias) dbias[c] = (accumulate ? dbias[c] : 0.f) + (float)s;
}

inline unsigned blocks_for(long total, int sm) {
  long b = (total + 255) / 256, cap = (long)sm * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}
}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int dg_dwconv3x3_fwd(dg_ctx* ctx, const dg_tensor* x, const float* w, const float* bias, const dg_tensor* y,
                                void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y) && w, "dg_dwconv3x3_fwd: null argument");
  DG_REQUIRE(dg_same_shape(x, y) && x->dtype == y->dtype, "dg_dwconv3x3_fwd: shape/dtype mismatch");
  long total = dg_pixels(x) * x->c;
  DG_DISPATCH_1(x->dtype, "dg_dwconv3x3_fwd",
                dw3x3_kernel<T, false><<<blocks_for(total, ctx->sm_count), 256, 0, ST>>>(
                    (const T*)x->ptr, x->cpitch, x->coff, w, bias, (T*)y->ptr, y->cpitch, y->coff, x->n, x->h, x->w, x->c););
  DG_CHECK_LAUNCH("dg_dwconv3x3_fwd");
  return 0;
}

extern "C" int dg_dwconv3x3_dgrad(dg_ctx* ctx, const dg_tensor* dy, const float* w, const dg_tensor* dx, void* stream) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(dx) && w, "dg_dwconv3x3_dgrad: null argument");
  DG_REQUIRE(dg_same_shape(dy, dx) && dy->dtype == dx->dtype, "dg_dwconv3x3_dgrad: shape/dtype mismatch");
  long total = dg_pixels(dy) * dy->c;
  DG_DISPATCH_1(dy->dtype, "dg_dwconv3x3_dgrad",
                dw3x3_kernel<T, true><<<blocks_for(total, ctx->sm_count), 256, 0, ST>>>(
                    (const T*)dy->ptr, dy->cpitch, dy->coff, w, nullptr, (T*)dx->ptr, dx->cpitch, dx->coff, dy->n, dy->h,
                    dy->w, dy->c););
  DG_CHECK_LAUNCH("dg_dwconv3x3_dgrad");
  return 0;
}

extern "C" size_t dg_dwconv3x3_wgrad_workspace_bytes(const dg_tensor* x) {
  return (size_t)256 * 8 * 10 * x->c * sizeof(float);
}

extern "C" int dg_dwconv3x3_wgrad(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* dy, float* dw, float* dbias,
                                  int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(dy) && dw && workspace, "dg_dwconv3x3_wgrad: null argument");
  DG_REQUIRE(dg_same_shape(x, dy) && x->dtype == dy->dtype, "dg_dwconv3x3_wgrad: shape/dtype mismatch");
  DG_REQUIRE(x->c <= RED_THREADS * MAX_CPT, "dg_dwconv3x3_wgrad: C too large");
  DG_REQUIRE(workspace_bytes >= dg_dwconv3x3_wgrad_workspace_bytes(x), "dg_dwconv3x3_wgrad: workspace too small");
  long P = dg_pixels(x);
  int C = x->c;
  int blocks = red_blocks(P, C, ctx->sm_count);
  int R = RED_THREADS / red_lanes(C);
  float* partial = (float*)workspace;
  DG_DISPATCH_1(x->dtype, "dg_dwconv3x3_wgrad",
                dw3x3_wgrad_kernel<T><<<blocks, RED_THREADS, (size_t)R * 10 * C * sizeof(float), ST>>>(
                    (const T*)x->ptr, x->cpitch, x->coff, (const T*)dy->ptr, dy->cpitch, dy->coff, x->n, x->h, x->w, C,
                    partial););
  dw3x3_wgrad_finalize<<<(10 * C + 127) / 128, 128, 0, ST>>>(partial, blocks, C, dw, dbias, accumulate);
  DG_CHECK_LAUNCH("dg_dwconv3x3_wgrad");
  return 0;
}
