// CUDA-core kernels for the "thin" convolutions: stride-1, same-size layers where one side has
// <= 4 channels (RGB first layers 3->32/64, the 64->3 / 64->1 / 32->3 heads, and their gradients).
// A tensor-core tile would be >= 75 % padding there and the layers are HBM/FFMA bound, so they get
// dedicated kernels instead of the generic implicit GEMM in conv_simt.cu:
//   expand   : thin input (<=4 ch)  -> fat output; thread = 1 pixel x 16 output channels, weights
//              broadcast from shared memory (forward of thin-Cin, dgrad of thin-Cout)
//   contract : fat input -> thin output (<=4 ch); thread = 1 pixel, 16-byte channel vectors
//              (forward of thin-Cout, dgrad of thin-Cin)
//   outer    : weight gradient dW[t][thin][fat] = sum_pixels thin (x) fat; warp = one (fat vector,
//              kernel row) role, lanes = pixels, shuffle reduction, per-block partials
// All three take the HWIO weight through (tap, thin, fat) strides, so forward and gradient forms
// share the code.  Reference call sites: srgan.py:154,182,246,268; fsrgan.py:198,217.
#pragma once
#include "dg_common.cuh"
#include "pointwise_vec.cuh"

namespace dgthin {

using dgvec::V8;

struct ThinGeom {
  int N, H, W, kh, kw;
  int dh0, dw0, dsign;     // tap (r,s) reads pixel (h + dsign*(r - dh0), w + dsign*(s - dw0))
  int CT, CF;              // thin (<=4) and fat channel counts
  int tp, to, fp, fo;      // pitch / channel offset of the thin and fat tensors
  int w_st, w_sthin, w_sfat;  // HWIO element strides for (tap, thin index, fat index)
  int act;
  float alpha;
};

constexpr int EXP_THREADS = 128;
constexpr int EXP_OG = 16;

// ---------------------------------------------------------------- expand: fat[p][o] = act(b[o] + sum_t sum_c thin[p+off_t][c] * W[t][c][o])
template <typename TT, typename TF>
__global__ void __launch_bounds__(EXP_THREADS)
thin_expand_kernel(const TT* __restrict__ thin, TF* __restrict__ fat, const float* __restrict__ w, const float* __restrict__ bias,
                   ThinGeom g) {
  __shared__ __align__(16) float ws[16 * 4][EXP_OG];
  const int taps = g.kh * g.kw;
  const int o0 = blockIdx.y * EXP_OG;
  for (int e = threadIdx.x; e < taps * g.CT * EXP_OG; e += EXP_THREADS) {
    int j = e % EXP_OG, k = e / EXP_OG;
    int c = k % g.CT, t = k / g.CT;
    ws[k][j] = w[(long)t * g.w_st + (long)c * g.w_sthin + (long)(o0 + j) * g.w_sfat];
  }
  __syncthreads();
  const uint32_t P = (uint32_t)g.N * g.H * g.W;
  const uint32_t p = blockIdx.x * EXP_THREADS + threadIdx.x;
  if (p >= P) return;
  const uint32_t t2 = p / (uint32_t)g.W;
  const int wq = (int)(p - t2 * (uint32_t)g.W);
  const long n = t2 / (uint32_t)g.H;
  const int h = (int)(t2 - (uint32_t)n * (uint32_t)g.H);
  float acc[EXP_OG];
#pragma unroll
  for (int j = 0; j < EXP_OG; ++j) acc[j] = bias ? __ldg(bias + o0 + j) : 0.f;
  for (int r = 0; r < g.kh; ++r) {
    const int hh = h + g.dsign * (r - g.dh0);
    if (hh < 0 || hh >= g.H) continue;
    for (int s = 0; s < g.kw; ++s) {
      const int ww = wq + g.dsign * (s - g.dw0);
      if (ww < 0 || ww >= g.W) continue;
      const TT* src = thin + ((long)((n * g.H + hh) * g.W + ww) * g.tp + g.to);
      const int kb = (r * g.kw + s) * g.CT;
      for (int c = 0; c < g.CT; ++c) {
        const float v = ld_f(src + c);
        const float4* wr = reinterpret_cast<const float4*>(ws[kb + c]);
#pragma unroll
        for (int q = 0; q < EXP_OG / 4; ++q) {
          float4 w4 = wr[q];
          acc[4 * q] = fmaf(v, w4.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
        }
      }
    }
  }
  float o8[8];
  TF* dst = fat + ((long)p * g.fp + g.fo + o0);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = apply_act(acc[half * 8 + j], g.act, g.alpha);
    V8<TF>::st(dst + half * 8, o8);
  }
}

// ---------------------------------------------------------------- contract: thin[p][j] = act(b[j] + sum_t sum_c fat[p+off_t][c] * W[t][c][j])
constexpr int CON_THREADS = 128;

template <typename TF, typename TT>
__global__ void __launch_bounds__(CON_THREADS)
thin_contract_kernel(const TF* __restrict__ fat, TT* __restrict__ thin, const float* __restrict__ w, const float* __restrict__ bias,
                     ThinGeom g) {
  extern __shared__ __align__(16) float4 w4s[];  // [taps*CF] : (j0..j3), zero-padded
  const int taps = g.kh * g.kw;
  for (int e = threadIdx.x; e < taps * g.CF; e += CON_THREADS) {
    int c = e % g.CF, t = e / g.CF;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < g.CT; ++j) v[j] = w[(long)t * g.w_st + (long)j * g.w_sthin + (long)c * g.w_sfat];
    w4s[e] = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  const uint32_t P = (uint32_t)g.N * g.H * g.W;
  const uint32_t p = blockIdx.x * CON_THREADS + threadIdx.x;
  if (p >= P) return;
  const uint32_t t2 = p / (uint32_t)g.W;
  const int wq = (int)(p - t2 * (uint32_t)g.W);
  const long n = t2 / (uint32_t)g.H;
  const int h = (int)(t2 - (uint32_t)n * (uint32_t)g.H);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < g.CT; ++j) acc[j] = bias ? __ldg(bias + j) : 0.f;
  for (int r = 0; r < g.kh; ++r) {
    const int hh = h + g.dsign * (r - g.dh0);
    if (hh < 0 || hh >= g.H) continue;
    for (int s = 0; s < g.kw; ++s) {
      const int ww = wq + g.dsign * (s - g.dw0);
      if (ww < 0 || ww >= g.W) continue;
      const TF* src = fat + ((long)((n * g.H + hh) * g.W + ww) * g.fp + g.fo);
      const float4* wt = w4s + (r * g.kw + s) * g.CF;
      for (int c0 = 0; c0 < g.CF; c0 += 8) {
        float v[8];
        V8<TF>::ld(src + c0, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 w4 = wt[c0 + k];
          acc[0] = fmaf(v[k], w4.x, acc[0]);
          acc[1] = fmaf(v[k], w4.y, acc[1]);
          acc[2] = fmaf(v[k], w4.z, acc[2]);
          acc[3] = fmaf(v[k], w4.w, acc[3]);
        }
      }
    }
  }
  TT* dst = thin + ((long)p * g.tp + g.to);
  for (int j = 0; j < g.CT; ++j) st_f(dst + j, apply_act(acc[j], g.act, g.alpha));
}

// ---------------------------------------------------------------- outer: dW[t][thin][fat] = sum_f fat[f][cf] * thin[f + off_t][ct]
// warp role = (fat 8-vector cv, kernel row r); lanes = pixels; accumulators [kw][CT][8].
// part layout per block: dW as HWIO strides (same as w) + bias slots behind it.
constexpr int OUT_MAX_KW = 4;

template <typename TF, typename TT, int CT, int KW>
__global__ void __launch_bounds__(512, 1) thin_outer_kernel(const TF* __restrict__ fat, const TT* __restrict__ thin, float* __restrict__ part, long part_stride,
                                  long bias_off, int bias_on_fat, int want_bias, ThinGeom g, long pix_per_block) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cv = warp, r = blockIdx.y;  // blockDim = 32 * CF/8 ; gridDim.y = kh
  const long P = (long)g.N * g.H * g.W;
  const long f_beg = (long)blockIdx.x * pix_per_block;
  const long f_end = min(P, f_beg + pix_per_block);
  float acc[KW][CT][8];
  float bsum[8];
  float tsum[CT];
#pragma unroll
  for (int s = 0; s < KW; ++s)
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[s][c][k] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) bsum[k] = 0.f;
#pragma unroll
  for (int c = 0; c < CT; ++c) tsum[c] = 0.f;
  // pixel coordinates advance incrementally (no division in the loop)
  int wq, h, n;
  {
    const uint32_t f0 = (uint32_t)(f_beg + lane);
    const uint32_t t2 = f0 / (uint32_t)g.W;
    wq = (int)(f0 - t2 * (uint32_t)g.W);
    n = (int)(t2 / (uint32_t)g.H);
    h = (int)(t2 - (uint32_t)n * (uint32_t)g.H);
  }
  for (long f = f_beg + lane; f < f_end; f += 32, wq += 32) {
    while (wq >= g.W) { wq -= g.W; if (++h == g.H) { h = 0; ++n; } }
    float v[8];
    V8<TF>::ld(fat + (f * g.fp + g.fo + cv * 8), v);
    if (r == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) bsum[k] += v[k];
      if (cv == 0) {
#pragma unroll
        for (int c = 0; c < CT; ++c) tsum[c] += ld_f(thin + (f * g.tp + g.to + c));
      }
    }
    const int hh = h + g.dsign * (r - g.dh0);
    if (hh < 0 || hh >= g.H) continue;
#pragma unroll
    for (int s = 0; s < KW; ++s) {
      const int ww = wq + g.dsign * (s - g.dw0);
      if (ww < 0 || ww >= g.W) continue;
      const TT* tp = thin + ((long)((n * g.H + hh) * g.W + ww) * g.tp + g.to);
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const float tv = ld_f(tp + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[s][c][k] = fmaf(tv, v[k], acc[s][c][k]);
      }
    }
  }
  float* out = part + (long)blockIdx.x * part_stride;
#pragma unroll
  for (int s = 0; s < KW; ++s) {
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float x = warp_sum(acc[s][c][k]);
        if (lane == 0) out[(long)(r * g.kw + s) * g.w_st + (long)c * g.w_sthin + (long)(cv * 8 + k) * g.w_sfat] = x;
      }
  }
  if (want_bias && r == 0) {
    if (bias_on_fat) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float x = warp_sum(bsum[k]);
        if (lane == 0) out[bias_off + cv * 8 + k] = x;
      }
    } else if (cv == 0) {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        float x = warp_sum(tsum[c]);
        if (lane == 0) out[bias_off + c] = x;
      }
    }
  }
}

}  // namespace dgthin
