// Depthwise 3x3 stride-1 SAME convolution (forward, input gradient, weight gradient).
// HBM-bound (9 MAC per element): one thread per output element, channels fastest so a warp
// touches contiguous memory; the weight gradient is a per-channel reduction of 10 values.
// Reference: keras DepthwiseConv2D fsrgan.py:149-154, kernel layout [3,3,C,1].
#include <cuda.h>
#include <stdlib.h>

#include "dg_common.cuh"
#include "reduce.cuh"
#include "pointwise_vec.cuh"
#include "sm100.cuh"

namespace {
using namespace dgred;

template <typename T, bool FLIP>
__global__ void dw3x3_kernel(const T* __restrict__ x, int xp, int xo, const float* __restrict__ w,
                             const float* __restrict__ bias, T* __restrict__ y, int yp, int yo, int N, int H, int W,
                             int C, int relu) {
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
    st_f(y + (p * yp + yo + c), relu ? fmaxf(acc, 0.f) : acc);
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

// ---------------------------------------------------------------- 8-channel x 4-pixel strip kernels
// One thread owns 8 channels (one 16-byte vector) of a strip of four consecutive output pixels of a row: the 3 x 6
// input window is loaded once (18 vector loads for 4 outputs instead of 36) and stays in registers as raw vectors.
using dgvec::V8;
constexpr int DW_TW = 4;

template <typename T>
__device__ __forceinline__ void dw_load_window(const T* __restrict__ x, int xp, int xo, int n, int h, int w0, int H, int W, int c0,
                                               typename V8<T>::raw (&win)[3][DW_TW + 2], bool (&ok)[3][DW_TW + 2]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int hh = h + a - 1;
#pragma unroll
    for (int b = 0; b < DW_TW + 2; ++b) {
      const int ww = w0 + b - 1;
      ok[a][b] = hh >= 0 && hh < H && ww >= 0 && ww < W;
      if (ok[a][b]) win[a][b] = V8<T>::ldraw(x + ((((long)n * H + hh) * W + ww) * xp + xo + c0));
    }
  }
}

template <typename T, bool FLIP>
__global__ void __launch_bounds__(256)
dw3x3_strip_kernel(const T* __restrict__ x, int xp, int xo, const float* __restrict__ w, const float* __restrict__ bias,
                   T* __restrict__ y, int yp, int yo, int N, int H, int W, int C, int relu) {
  const int CV = C >> 3, SW = W / DW_TW;
  const long total = (long)N * H * SW * CV;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % CV) * 8;
    long s = i / CV;
    const int w0 = (int)(s % SW) * DW_TW;
    s /= SW;
    const int h = (int)(s % H), n = (int)(s / H);
    typename V8<T>::raw win[3][DW_TW + 2];
    bool ok[3][DW_TW + 2];
    dw_load_window<T>(x, xp, xo, n, h, w0, H, W, c0, win, ok);
    float acc[DW_TW][8];
    float b8[8];
    if (bias) dgvec::ldc8(bias + c0, b8);
#pragma unroll
    for (int t = 0; t < DW_TW; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = bias ? b8[j] : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < DW_TW + 2; ++b) {
        if (!ok[a][b]) continue;
        float v[8];
        V8<T>::cvt(win[a][b], v);
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {       // this column feeds output t = b - kb
          const int t = b - kb;
          if (t < 0 || t >= DW_TW) continue;
          float wk[8];
          dgvec::ldc8(w + (FLIP ? ((2 - a) * 3 + (2 - kb)) : (a * 3 + kb)) * C + c0, wk);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(v[j], wk[j], acc[t][j]);
        }
      }
    if (relu) {
#pragma unroll
      for (int t = 0; t < DW_TW; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = fmaxf(acc[t][j], 0.f);
    }
#pragma unroll
    for (int t = 0; t < DW_TW; ++t) V8<T>::st(y + ((((long)n * H + h) * W + w0 + t) * yp + yo + c0), acc[t]);
  }
}

// Weight/bias gradient: every thread keeps its 8 channels' ten sums (9 taps + bias) in registers over all its strips,
// the block folds them through shared memory one tap at a time, one fp32 partial per block; dw3x3_wgrad_finalize sums.
constexpr int DW_RT = 256;
template <typename T>
__global__ void __launch_bounds__(DW_RT, 1)
dw3x3_wgrad_strip_kernel(const T* __restrict__ x, int xp, int xo, const T* __restrict__ dy, int dp, int dof, int N, int H, int W,
                         int C, float* __restrict__ partial) {
  extern __shared__ float dw_red[];                 // [R][C]
  const int CV = C >> 3, R = DW_RT / CV, SW = W / DW_TW;
  const int lane = threadIdx.x % CV, row = threadIdx.x / CV, c0 = lane * 8;
  float acc[10][8];
#pragma unroll
  for (int k = 0; k < 10; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  if (row < R) {
    const long strips = (long)N * H * SW;
    for (long s = (long)blockIdx.x * R + row; s < strips; s += (long)gridDim.x * R) {
      const int w0 = (int)(s % SW) * DW_TW;
      const long r = s / SW;
      const int h = (int)(r % H), n = (int)(r / H);
      typename V8<T>::raw win[3][DW_TW + 2], gr[DW_TW];
      bool ok[3][DW_TW + 2];
#pragma unroll
      for (int t = 0; t < DW_TW; ++t) gr[t] = V8<T>::ldraw(dy + ((((long)n * H + h) * W + w0 + t) * dp + dof + c0));
      dw_load_window<T>(x, xp, xo, n, h, w0, H, W, c0, win, ok);
      float g[DW_TW][8];
#pragma unroll
      for (int t = 0; t < DW_TW; ++t) {
        V8<T>::cvt(gr[t], g[t]);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[9][j] += g[t][j];
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < DW_TW + 2; ++b) {
          if (!ok[a][b]) continue;
          float v[8];
          V8<T>::cvt(win[a][b], v);
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) {
            const int t = b - kb;
            if (t < 0 || t >= DW_TW) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[a * 3 + kb][j] = fmaf(g[t][j], v[j], acc[a * 3 + kb][j]);
          }
        }
    }
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if (row < R) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dw_red[row * C + c0 + j] = acc[k][j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += DW_RT) {
      float s = 0.f;
      for (int r = 0; r < R; ++r) s += dw_red[r * C + c];
      partial[((long)blockIdx.x * 10 + k) * C + c] = s;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- TMA-fed tile kernel (bf16, C % 32 == 0)
// The strip kernel above re-reads every input vector from L1/L2 4.5 times and its 36 weight-vector loads per thread outnumber
// the data loads 4:1; at 174 registers one block per SM keeps ~24 KB of loads in flight: 1.64 ms for the 192-channel layers of
// a 1080p Fast-SRGAN frame (2 GB of traffic, 0.31 ms at the measured HBM rate; profiles/infer_profile_r2_fsrgan.log).
// Here a persistent block walks (channel block, image, tile) triples; thread 0 keeps DWT_STAGES halo tiles
// ((8+2) x (TW+2) pixels x CB channels, out-of-image pixels zero-filled by the TMA unit = SAME padding) in flight through an
// mbarrier ring, so ~130 KB per SM are on their way without costing a register.  A thread owns 8 channels of ONE output column
// and walks down the 8 rows of the tile with a rolling 3 x 3 window in registers (3 shared-memory loads per output vector,
// conflict-free: a warp reads 512 contiguous bytes) and its 72 weights in registers for as long as the channel block lasts.
constexpr int DWT_THREADS = 256, DWT_TH = 8, DWT_STAGES = 4;

struct DwTmaParams {
  CUtensorMap xmap;
  const float* w;
  const float* bias;
  __nv_bfloat16* y;
  int yp, yo, N, H, W, C, relu, flip, tiles_w, tiles_h, cblocks;
};

template <int CB>
__global__ void __launch_bounds__(DWT_THREADS, 1) dw3x3_tma_kernel(const __grid_constant__ DwTmaParams P) {
  using namespace sm100;
  constexpr int CV = CB / 8, TW = DWT_THREADS / CV, IW = TW + 2, IH = DWT_TH + 2;
  constexpr uint32_t STAGE = (uint32_t)IH * IW * CB * 2;
  extern __shared__ uint8_t dwt_raw[];
  __shared__ __align__(8) uint64_t full[DWT_STAGES];
  const uint32_t base = (smem_u32(dwt_raw) + 127u) & ~127u;
  const int tid = threadIdx.x, cv = tid % CV, col = tid / CV;
  const int per_img = P.tiles_h * P.tiles_w, per_cb = P.N * per_img, total = per_cb * P.cblocks;
  if (tid == 0) {
    for (int s = 0; s < DWT_STAGES; ++s) mbar_init(smem_u32(&full[s]), 1);
    fence_mbar_init();
    tma_prefetch_desc(&P.xmap);
  }
  __syncthreads();
  auto issue = [&](int tile, int s) {
    const int cb = tile / per_cb, r = tile - cb * per_cb, n = r / per_img, q = r - n * per_img, th = q / P.tiles_w, tw = q - th * P.tiles_w;
    const uint32_t bar = smem_u32(&full[s]);
    mbar_expect_tx(bar, STAGE);
    tma_load_4d(base + (uint32_t)s * STAGE, &P.xmap, bar, cb * CB, tw * TW - 1, th * DWT_TH - 1, n);
  };
  if (tid == 0)
    for (int s = 0; s < DWT_STAGES; ++s) {
      const int tile = blockIdx.x + s * (int)gridDim.x;
      if (tile < total) issue(tile, s);
    }
  float wk[9][8], b8[8];
  int cur_cb = -1, k = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++k) {
    const int s = k % DWT_STAGES;
    const int cb = tile / per_cb, r = tile - cb * per_cb, n = r / per_img, q = r - n * per_img, th = q / P.tiles_w, tw = q - th * P.tiles_w;
    const int c0 = cb * CB + cv * 8;
    if (cb != cur_cb) {
      cur_cb = cb;
#pragma unroll
      for (int t = 0; t < 9; ++t) dgvec::ldc8(P.w + (P.flip ? 8 - t : t) * P.C + c0, wk[t]);
#pragma unroll
      for (int j = 0; j < 8; ++j) b8[j] = 0.f;
      if (P.bias) dgvec::ldc8(P.bias + c0, b8);
    }
    mbar_wait(smem_u32(&full[s]), (uint32_t)(k / DWT_STAGES) & 1u);
    const uint32_t sa = base + (uint32_t)s * STAGE + (uint32_t)cv * 16u;
    const int w = tw * TW + col, h0 = th * DWT_TH;
    float win[3][3][8];
    auto load_row = [&](int r_in, float (&dst)[3][8]) {
#pragma unroll
      for (int dc = 0; dc < 3; ++dc) {
        uint4 u;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                     : "r"(sa + (uint32_t)((r_in * IW + col + dc) * CB * 2)));
        dgvec::V8<__nv_bfloat16>::cvt(u, dst[dc]);
      }
    };
    load_row(0, win[0]);
    load_row(1, win[1]);
    __nv_bfloat16* yrow = P.y + ((((long)n * P.H + h0) * P.W + w) * P.yp + P.yo + c0);
#pragma unroll
    for (int h = 0; h < DWT_TH; ++h) {
      load_row(h + 2, win[(h + 2) % 3]);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = b8[j];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(win[(h + a) % 3][b][j], wk[a * 3 + b][j], acc[j]);
      if (P.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
      }
      if (w < P.W && h0 + h < P.H) dgvec::V8<__nv_bfloat16>::st(yrow + (long)h * P.W * P.yp, acc);
    }
    __syncthreads();                       // every thread has read stage s
    if (tid == 0) {
      const int nt = tile + DWT_STAGES * (int)gridDim.x;
      if (nt < total) issue(nt, s);
    }
  }
}

typedef CUresult (*DwEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline bool dw_tma_ok(const dg_ctx* ctx, const dg_tensor* x, const dg_tensor* y) {
  static const char* off = getenv("DG_DW_TMA");
  if (off && off[0] == '0') return false;
  return ctx->encode_tiled && x->dtype == DG_BF16 && y->dtype == DG_BF16 && x->c % 32 == 0 && x->cpitch % 8 == 0 && x->coff % 8 == 0 &&
         y->cpitch % 8 == 0 && y->coff % 8 == 0 && ((uintptr_t)x->ptr % 16) == 0 && ((uintptr_t)y->ptr % 16) == 0 &&
         (long)x->n * x->h * x->w < (1L << 31) / 2 && x->h >= 2 && x->w >= 2;
}

static int dw_tma_launch(const char* name, dg_ctx* ctx, const dg_tensor* x, const float* w, const float* bias, const dg_tensor* y, int relu,
                         int flip, cudaStream_t st) {
  const int CB = x->c % 64 == 0 ? 64 : 32, CV = CB / 8, TW = DWT_THREADS / CV;
  DwTmaParams P;
  memset(&P, 0, sizeof(P));
  uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
  uint64_t strides[3] = {(uint64_t)x->cpitch * 2, (uint64_t)x->cpitch * 2 * x->w, (uint64_t)x->cpitch * 2 * x->w * x->h};
  uint32_t box[4] = {(uint32_t)CB, (uint32_t)(TW + 2), (uint32_t)(DWT_TH + 2), 1}, ones[4] = {1, 1, 1, 1};
  CUresult r = ((DwEncodeFn)ctx->encode_tiled)(&P.xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (char*)x->ptr + (size_t)x->coff * 2,
                                               (const cuuint64_t*)dims, (const cuuint64_t*)strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DG_FAIL("%s: cuTensorMapEncodeTiled failed (%d)", name, (int)r);
  P.w = w; P.bias = bias; P.y = (__nv_bfloat16*)y->ptr; P.yp = y->cpitch; P.yo = y->coff;
  P.N = x->n; P.H = x->h; P.W = x->w; P.C = x->c; P.relu = relu; P.flip = flip;
  P.tiles_w = (x->w + TW - 1) / TW; P.tiles_h = (x->h + DWT_TH - 1) / DWT_TH; P.cblocks = x->c / CB;
  const long total = (long)P.cblocks * P.N * P.tiles_h * P.tiles_w;
  const uint32_t smem = (uint32_t)DWT_STAGES * (DWT_TH + 2) * (TW + 2) * CB * 2 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dw3x3_tma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dw3x3_tma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) DG_FAIL("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    attr_set = true;
  }
  const unsigned grid = (unsigned)(total < ctx->sm_count ? total : ctx->sm_count);
  if (CB == 64) dw3x3_tma_kernel<64><<<grid, DWT_THREADS, smem, st>>>(P);
  else dw3x3_tma_kernel<32><<<grid, DWT_THREADS, smem, st>>>(P);
  DG_CHECK_LAUNCH(name);
  return 0;
}

static inline bool dw_strip_ok(const dg_tensor* a, const dg_tensor* b) {
  return dgvec::vec_ok(a) && dgvec::vec_ok(b) && a->w % DW_TW == 0 && a->c <= 8 * DW_RT;
}

inline unsigned blocks_for(long total, int sm) {
  long b = (total + 255) / 256, cap = (long)sm * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}
}  // namespace

#define ST ((cudaStream_t)stream)

static int dwconv_fwd_impl(dg_ctx* ctx, const dg_tensor* x, const float* w, const float* bias, const dg_tensor* y, int relu, void* stream);

extern "C" int dg_dwconv3x3_fwd(dg_ctx* ctx, const dg_tensor* x, const float* w, const float* bias, const dg_tensor* y,
                                void* stream) {
  return dwconv_fwd_impl(ctx, x, w, bias, y, 0, stream);
}

// DepthwiseConv2D + bias + ReLU in one pass: the inference form of fsrgan.py:149-155 (DepthwiseConv2D -> BatchNormalization ->
// ReLU) once the BatchNorm affine has been folded into the kernel and the bias (infer_video.py:146, training=False).
extern "C" int dg_dwconv3x3_fwd_act(dg_ctx* ctx, const dg_tensor* x, const float* w, const float* bias, int act, const dg_tensor* y,
                                    void* stream) {
  DG_REQUIRE(act == DG_ACT_NONE || act == DG_ACT_RELU, "dg_dwconv3x3_fwd_act: activation must be none or relu");
  return dwconv_fwd_impl(ctx, x, w, bias, y, act == DG_ACT_RELU ? 1 : 0, stream);
}

static int dwconv_fwd_impl(dg_ctx* ctx, const dg_tensor* x, const float* w, const float* bias, const dg_tensor* y, int relu, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y) && w, "dg_dwconv3x3_fwd: null argument");
  DG_REQUIRE(dg_same_shape(x, y) && x->dtype == y->dtype, "dg_dwconv3x3_fwd: shape/dtype mismatch");
  long total = dg_pixels(x) * x->c;
  if (dw_tma_ok(ctx, x, y)) return dw_tma_launch("dg_dwconv3x3_fwd", ctx, x, w, bias, y, relu, 0, ST);
  if (dw_strip_ok(x, y)) {
    DG_DISPATCH_1(x->dtype, "dg_dwconv3x3_fwd",
                  dw3x3_strip_kernel<T, false><<<blocks_for(total / (8 * DW_TW), ctx->sm_count), 256, 0, ST>>>(
                      (const T*)x->ptr, x->cpitch, x->coff, w, bias, (T*)y->ptr, y->cpitch, y->coff, x->n, x->h, x->w, x->c, relu););
    DG_CHECK_LAUNCH("dg_dwconv3x3_fwd");
    return 0;
  }
  DG_DISPATCH_1(x->dtype, "dg_dwconv3x3_fwd",
                dw3x3_kernel<T, false><<<blocks_for(total, ctx->sm_count), 256, 0, ST>>>(
                    (const T*)x->ptr, x->cpitch, x->coff, w, bias, (T*)y->ptr, y->cpitch, y->coff, x->n, x->h, x->w, x->c, relu););
  DG_CHECK_LAUNCH("dg_dwconv3x3_fwd");
  return 0;
}

extern "C" int dg_dwconv3x3_dgrad(dg_ctx* ctx, const dg_tensor* dy, const float* w, const dg_tensor* dx, void* stream) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(dx) && w, "dg_dwconv3x3_dgrad: null argument");
  DG_REQUIRE(dg_same_shape(dy, dx) && dy->dtype == dx->dtype, "dg_dwconv3x3_dgrad: shape/dtype mismatch");
  long total = dg_pixels(dy) * dy->c;
  if (dw_tma_ok(ctx, dy, dx)) return dw_tma_launch("dg_dwconv3x3_dgrad", ctx, dy, w, nullptr, dx, 0, 1, ST);
  if (dw_strip_ok(dy, dx)) {
    DG_DISPATCH_1(dy->dtype, "dg_dwconv3x3_dgrad",
                  dw3x3_strip_kernel<T, true><<<blocks_for(total / (8 * DW_TW), ctx->sm_count), 256, 0, ST>>>(
                      (const T*)dy->ptr, dy->cpitch, dy->coff, w, nullptr, (T*)dx->ptr, dx->cpitch, dx->coff, dy->n, dy->h,
                      dy->w, dy->c, 0););
    DG_CHECK_LAUNCH("dg_dwconv3x3_dgrad");
    return 0;
  }
  DG_DISPATCH_1(dy->dtype, "dg_dwconv3x3_dgrad",
                dw3x3_kernel<T, true><<<blocks_for(total, ctx->sm_count), 256, 0, ST>>>(
                    (const T*)dy->ptr, dy->cpitch, dy->coff, w, nullptr, (T*)dx->ptr, dx->cpitch, dx->coff, dy->n, dy->h,
                    dy->w, dy->c, 0););
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
  if (dw_strip_ok(x, dy)) {
    const int Rr = DW_RT / (C >> 3);
    const long strips = P / DW_TW;
    long want = (strips + Rr - 1) / Rr;
    const int sblocks = (int)(want < ctx->sm_count ? (want > 0 ? want : 1) : ctx->sm_count);
    DG_DISPATCH_1(x->dtype, "dg_dwconv3x3_wgrad",
                  dw3x3_wgrad_strip_kernel<T><<<sblocks, DW_RT, (size_t)Rr * C * sizeof(float), ST>>>(
                      (const T*)x->ptr, x->cpitch, x->coff, (const T*)dy->ptr, dy->cpitch, dy->coff, x->n, x->h, x->w, C,
                      partial););
    dw3x3_wgrad_finalize<<<(10 * C + 127) / 128, 128, 0, ST>>>(partial, sblocks, C, dw, dbias, accumulate);
    DG_CHECK_LAUNCH("dg_dwconv3x3_wgrad");
    return 0;
  }
  DG_DISPATCH_1(x->dtype, "dg_dwconv3x3_wgrad",
                dw3x3_wgrad_kernel<T><<<blocks, RED_THREADS, (size_t)R * 10 * C * sizeof(float), ST>>>(
                    (const T*)x->ptr, x->cpitch, x->coff, (const T*)dy->ptr, dy->cpitch, dy->coff, x->n, x->h, x->w, C,
                    partial););
  dw3x3_wgrad_finalize<<<(10 * C + 127) / 128, 128, 0, ST>>>(partial, blocks, C, dw, dbias, accumulate);
  DG_CHECK_LAUNCH("dg_dwconv3x3_wgrad");
  return 0;
}
