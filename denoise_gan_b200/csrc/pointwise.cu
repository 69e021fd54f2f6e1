// HBM-bound kernels of the hot path: BatchNorm statistics / apply / backward, activation
// backward, depth_to_space(+PReLU), add, copy-into-slice, max-pool, nearest upsample(+ReLU).
// All reductions are warp-shuffle + shared-memory block reductions into per-block partials that
// a finalize kernel sums in a fixed order (deterministic; no float atomics).
//
// Reference call sites: BatchNormalization srgan.py:155,248, fsrgan.py:140-172, pix2pix.py:119,135,211;
// tf.nn.depth_to_space + PReLU srgan.py:145-146; MaxPool2D / UpSampling2D autoencoder.py:110,122-124;
// Dropout pix2pix.py:138; Add srgan.py:169,175.
#include <stdlib.h>

#include "dg_common.cuh"
#include "reduce.cuh"
#include "pointwise_vec.cuh"

namespace {

struct View {
  int pitch, off;
};

using namespace dgred;


// Sums NV per-channel partials over `nblocks` blocks (layout [block][NV][C]) for the 8 channels owned by
// this thread block: 256 threads = 32 block-lanes x 8 channels; result valid in threads with lane == 0.
template <int NV>
__device__ __forceinline__ bool sum_partials8(const float* __restrict__ partial, int nblocks, int C, double (&out)[NV], int* c_out) {
  __shared__ double sm_part[NV][32][8];
  const int cl = threadIdx.x & 7, lane = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  if (c < C)
    for (int b = lane; b < nblocks; b += 32)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] += (double)partial[((long)b * NV + v) * C + c];
#pragma unroll
  for (int v = 0; v < NV; ++v) sm_part[v][lane][cl] = acc[v];
  __syncthreads();
  *c_out = c;
  if (lane != 0 || c >= C) return false;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double t = 0.0;
    for (int l = 0; l < 32; ++l) t += sm_part[v][l][cl];
    out[v] = t;
  }
  return true;
}


// Element loop of the scalar (any channel count) kernels: 32-bit index arithmetic whenever the tensor allows it -- the
// 64-bit division per element made these kernels run at ~1 TB/s on the 3-channel fp32 images around the generator output.
#define DG_ELEM_LOOP(total, C, body)                                                                                   \
  if ((total) < (1L << 31)) {                                                                                          \
    const unsigned total32_ = (unsigned)(total), C32_ = (unsigned)(C), step32_ = gridDim.x * blockDim.x;                \
    for (unsigned i_ = blockIdx.x * blockDim.x + threadIdx.x; i_ < total32_; i_ += step32_) {                           \
      const unsigned p32_ = i_ / C32_;                                                                                 \
      body((long)p32_, (int)(i_ - p32_ * C32_));                                                                       \
    }                                                                                                                  \
  } else {                                                                                                             \
    for (long i_ = (long)blockIdx.x * blockDim.x + threadIdx.x; i_ < (total); i_ += (long)gridDim.x * blockDim.x) {     \
      const long p64_ = i_ / (C);                                                                                      \
      body(p64_, (int)(i_ - p64_ * (C)));                                                                              \
    }                                                                                                                  \
  }

// ------------------------------------------------------------------ BN statistics
template <typename T>
__global__ void __launch_bounds__(RED_THREADS) bn_stats_kernel(const T* __restrict__ x, View xv, long P, int C,
                                                               float* __restrict__ partial) {
  channel_reduce<2>(P, C, partial, [&](long p, int c, float* a) {
    float v = ld_f(x + (p * xv.pitch + xv.off + c));
    a[0] += v;
    a[1] += v * v;
  });
}

__global__ void bn_finalize_kernel(const float* __restrict__ partial, int nblocks, long P, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd) {
  double sums[2];
  int c;
  if (!sum_partials8<2>(partial, nblocks, C, sums, &c)) return;
  double mean = sums[0] / (double)P;
  double var = sums[1] / (double)P - mean * mean;
  if (var < 0.0) var = 0.0;
  float invstd = (float)(1.0 / sqrt(var + (double)eps));
  float g = gamma[c], b = beta[c];
  scale[c] = g * invstd;
  shift[c] = b - (float)mean * g * invstd;
  save_mean[c] = (float)mean;
  save_invstd[c] = invstd;
  if (moving_mean) {
    moving_mean[c] = moving_mean[c] * momentum + (float)mean * (1.f - momentum);
    moving_var[c] = moving_var[c] * momentum + (float)(var * ((double)P / (double)(P > 1 ? P - 1 : 1))) * (1.f - momentum);   // Bessel-corrected, as Keras' fused path
  }
}

__global__ void bn_infer_affine_kernel(int C, const float* gamma, const float* beta, const float* mm, const float* mv,
                                       float eps, float* scale, float* shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float invstd = rsqrtf(mv[c] + eps);
  scale[c] = gamma[c] * invstd;
  shift[c] = beta[c] - mm[c] * gamma[c] * invstd;
}

// ------------------------------------------------------------------ BN apply (+dropout +act +residual)
template <typename TI, typename TO>
__global__ void bn_act_fwd_kernel(const TI* __restrict__ x, View xv, const float* __restrict__ scale,
                                  const float* __restrict__ shift, int act, float alpha,
                                  const float* __restrict__ prelu_alpha, const TO* __restrict__ res, View rv,
                                  int dropout, uint32_t seed0, uint32_t offset, const int64_t* __restrict__ ctr, TO* __restrict__ y, View yv, long P,
                                  int C) {
  const uint32_t seed = seed0 + (ctr ? (uint32_t)(*ctr) * 0x9E3779B9u : 0u);  // a new mask every optimiser step, graph-replay safe
  long total = P * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long p = i / C;
    int c = (int)(i - p * C);
    float t = ld_f(x + (p * xv.pitch + xv.off + c)) * scale[c] + shift[c];
    if (dropout) t = dropout_keep(seed, offset + (uint32_t)i) ? 2.f * t : 0.f;
    if (act == DG_ACT_PRELU) t = t > 0.f ? t : prelu_alpha[c] * t;
    else t = apply_act(t, act, alpha);
    if (res) t += ld_f(res + (p * rv.pitch + rv.off + c));
    st_f(y + (p * yv.pitch + yv.off + c), t);
  }
}

// g = dL/d(bn output) for one element
template <typename TG, typename TX>
__device__ __forceinline__ float bn_bwd_g(const TG* dy, View dv, const TX* x, View xv, const float* scale,
                                          const float* shift, int act, float alpha, const float* prelu_alpha,
                                          int dropout, uint32_t seed, uint32_t offset, long p, int c, int C,
                                          float* t_out) {
  float xin = ld_f(x + (p * xv.pitch + xv.off + c));
  float t = xin * scale[c] + shift[c];
  float g = ld_f(dy + (p * dv.pitch + dv.off + c));
  float td = t;
  float dmul = 1.f;
  if (dropout) {
    bool keep = dropout_keep(seed, offset + (uint32_t)(p * C + c));
    td = keep ? 2.f * t : 0.f;
    dmul = keep ? 2.f : 0.f;
  }
  *t_out = td;
  float d;
  switch (act) {
    case DG_ACT_RELU: d = td > 0.f ? 1.f : 0.f; break;
    case DG_ACT_LRELU: d = td >= 0.f ? 1.f : alpha; break;
    case DG_ACT_PRELU: d = td > 0.f ? 1.f : prelu_alpha[c]; break;
    case DG_ACT_TANH: { float y = tanhf(td); d = 1.f - y * y; break; }
    case DG_ACT_SIGMOID: { float y = 1.f / (1.f + expf(-td)); d = y * (1.f - y); break; }
    default: d = 1.f;
  }
  return g * d * dmul;
}

template <typename TG, typename TX>
__global__ void __launch_bounds__(RED_THREADS)
bn_bwd_reduce_kernel(const TG* __restrict__ dy, View dv, const TX* __restrict__ x, View xv,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ invstd, int act, float alpha,
                     const float* __restrict__ prelu_alpha, int dropout, uint32_t seed0, uint32_t offset, const int64_t* __restrict__ ctr, long P, int C,
                     float* __restrict__ partial) {
  const uint32_t seed = seed0 + (ctr ? (uint32_t)(*ctr) * 0x9E3779B9u : 0u);  // a new mask every optimiser step, graph-replay safe
  channel_reduce<3>(P, C, partial, [&](long p, int c, float* a) {
    float t;
    float g = bn_bwd_g(dy, dv, x, xv, scale, shift, act, alpha, prelu_alpha, dropout, seed, offset, p, c, C, &t);
    float xhat = (ld_f(x + (p * xv.pitch + xv.off + c)) - mean[c]) * invstd[c];
    a[0] += g;
    a[1] += g * xhat;
    if (act == DG_ACT_PRELU) a[2] += ld_f(dy + (p * dv.pitch + dv.off + c)) * fminf(t, 0.f);
  });
}

// sums partials; writes dgamma/dbeta/dalpha and the per-channel means used by the dx pass
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, long P, int C,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ dalpha, int accumulate, float* __restrict__ coef) {
  double sums[3];
  int c;
  if (!sum_partials8<3>(partial, nblocks, C, sums, &c)) return;
  const double s0 = sums[0], s1 = sums[1], s2 = sums[2];
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s0;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s1;
  if (dalpha) dalpha[c] = (accumulate ? dalpha[c] : 0.f) + (float)s2;
  coef[c] = (float)(s0 / (double)P);
  coef[C + c] = (float)(s1 / (double)P);
}

template <typename TG, typename TX, typename TO>
__global__ void bn_bwd_dx_kernel(const TG* __restrict__ dy, View dv, const TX* __restrict__ x, View xv,
                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                 const float* __restrict__ gamma, const float* __restrict__ mean,
                                 const float* __restrict__ invstd, int act, float alpha,
                                 const float* __restrict__ prelu_alpha, int dropout, uint32_t seed0, uint32_t offset, const int64_t* __restrict__ ctr,
                                 const float* __restrict__ coef, TO* __restrict__ dx, View ov, long P, int C) {
  const uint32_t seed = seed0 + (ctr ? (uint32_t)(*ctr) * 0x9E3779B9u : 0u);  // a new mask every optimiser step, graph-replay safe
  long total = P * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long p = i / C;
    int c = (int)(i - p * C);
    float t;
    float g = bn_bwd_g(dy, dv, x, xv, scale, shift, act, alpha, prelu_alpha, dropout, seed, offset, p, c, C, &t);
    float xhat = (ld_f(x + (p * xv.pitch + xv.off + c)) - mean[c]) * invstd[c];
    float v = gamma[c] * invstd[c] * (g - coef[c] - xhat * coef[C + c]);
    st_f(dx + (p * ov.pitch + ov.off + c), v);
  }
}

// ------------------------------------------------------------------ activation backward from output
template <typename TG, typename TY, typename TO>
__global__ void act_bwd_kernel(const TG* __restrict__ dy, View dv, const TY* __restrict__ y, View yv, int act,
                               float alpha, TO* __restrict__ dpre, View ov, long P, int C) {
  const long total = P * C;
  auto body = [&](long p, int c) {
    float yo = ld_f(y + (p * yv.pitch + yv.off + c));
    float g = ld_f(dy + (p * dv.pitch + dv.off + c));
    float d;
    switch (act) {
      case DG_ACT_RELU: d = yo > 0.f ? 1.f : 0.f; break;
      case DG_ACT_LRELU: d = yo >= 0.f ? 1.f : alpha; break;
      case DG_ACT_TANH: d = 1.f - yo * yo; break;
      case DG_ACT_SIGMOID: d = yo * (1.f - yo); break;
      default: d = 1.f;
    }
    st_f(dpre + (p * ov.pitch + ov.off + c), g * d);
  };
  DG_ELEM_LOOP(total, C, body)
}

// ------------------------------------------------------------------ depth_to_space(2) + PReLU
// y[b,2h+i,2w+j,c] = prelu(u[b,h,w,(2i+j)*Co+c], alpha[c])      (TF "DCR" order)
template <typename T>
__global__ void d2s_prelu_fwd_kernel(const T* __restrict__ u, View uv, const float* __restrict__ alpha,
                                     T* __restrict__ y, View yv, int N, int H, int W, int Co) {
  long total = (long)N * H * W * 4 * Co;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int ch = (int)(i % (4 * Co));
    long p = i / (4 * Co);
    int w = (int)(p % W);
    long t = p / W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    int sub = ch / Co, c = ch - sub * Co;
    int di = sub >> 1, dj = sub & 1;
    float v = ld_f(u + (p * uv.pitch + uv.off + ch));
    if (alpha) v = v > 0.f ? v : alpha[c] * v;
    long q = ((long)n * 2 * H + 2 * h + di) * 2 * W + 2 * w + dj;
    st_f(y + (q * yv.pitch + yv.off + c), v);
  }
}

// du = dy_perm * (u > 0 ? 1 : alpha[c]) ; dalpha[c] = sum dy_perm * min(u, 0)
template <typename T>
__global__ void d2s_prelu_bwd_kernel(const T* __restrict__ dy, View dv, const T* __restrict__ u, View uv,
                                     const float* __restrict__ alpha, T* __restrict__ du, View ov, int N, int H, int W,
                                     int Co) {
  long total = (long)N * H * W * 4 * Co;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int ch = (int)(i % (4 * Co));
    long p = i / (4 * Co);
    int w = (int)(p % W);
    long t = p / W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    int sub = ch / Co, c = ch - sub * Co;
    long q = ((long)n * 2 * H + 2 * h + (sub >> 1)) * 2 * W + 2 * w + (sub & 1);
    float g = ld_f(dy + (q * dv.pitch + dv.off + c));
    if (alpha) {
      float uu = ld_f(u + (p * uv.pitch + uv.off + ch));
      g = uu > 0.f ? g : alpha[c] * g;
    }
    st_f(du + (p * ov.pitch + ov.off + ch), g);
  }
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
d2s_prelu_dalpha_kernel(const T* __restrict__ dy, View dv, const T* __restrict__ u, View uv, long Pout, int H2, int W2,
                        int Co, float* __restrict__ partial) {
  // reduction over OUTPUT pixels q, channel c; the source element is u[q/2][(2(i)+j)*Co + c]
  channel_reduce<1>(Pout, Co, partial, [&](long q, int c, float* a) {
    int w2 = (int)(q % W2);
    long t = q / W2;
    int h2 = (int)(t % H2);
    long n = t / H2;
    long p = (n * (H2 / 2) + h2 / 2) * (W2 / 2) + w2 / 2;
    int sub = (h2 & 1) * 2 + (w2 & 1);
    float uu = ld_f(u + (p * uv.pitch + uv.off + sub * Co + c));
    a[0] += ld_f(dy + (q * dv.pitch + dv.off + c)) * fminf(uu, 0.f);
  });
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ out,
                                    int accumulate) {
  double sums[1];
  int c;
  if (!sum_partials8<1>(partial, nblocks, C, sums, &c)) return;
  out[c] = (accumulate ? out[c] : 0.f) + (float)sums[0];
}

// ------------------------------------------------------------------ add / copy
template <typename TA, typename TO>
__global__ void add_kernel(const TA* __restrict__ a, View av, const TA* __restrict__ b, View bv, TO* __restrict__ o,
                           View ov, long P, int C) {
  const long total = P * C;
  auto body = [&](long p, int c) {
    st_f(o + (p * ov.pitch + ov.off + c),
         ld_f(a + (p * av.pitch + av.off + c)) + ld_f(b + (p * bv.pitch + bv.off + c)));
  };
  DG_ELEM_LOOP(total, C, body)
}

template <typename TI, typename TO>
__global__ void copy_kernel(const TI* __restrict__ s, View sv, TO* __restrict__ o, View ov, long P, int C,
                            int accumulate) {
  const long total = P * C;
  auto body = [&](long p, int c) {
    float v = ld_f(s + (p * sv.pitch + sv.off + c));
    TO* dst = o + (p * ov.pitch + ov.off + c);
    if (accumulate) v += ld_f(dst);
    st_f(dst, v);
  };
  DG_ELEM_LOOP(total, C, body)
}

// ------------------------------------------------------------------ max-pool 2x2 / upsample 2x
template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, View xv, T* __restrict__ y, View yv, int N, int Ho, int Wo,
                                   int C) {
  long total = (long)N * Ho * Wo * C;
  const int W = Wo * 2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long q = i / C;
    int wo = (int)(q % Wo);
    long t = q / Wo;
    int ho = (int)(t % Ho);
    long n = t / Ho;
    long p00 = (n * 2 * Ho + 2 * ho) * W + 2 * wo;
    float v = ld_f(x + (p00 * xv.pitch + xv.off + c));
    v = fmaxf(v, ld_f(x + ((p00 + 1) * xv.pitch + xv.off + c)));
    v = fmaxf(v, ld_f(x + ((p00 + W) * xv.pitch + xv.off + c)));
    v = fmaxf(v, ld_f(x + ((p00 + W + 1) * xv.pitch + xv.off + c)));
    st_f(y + (q * yv.pitch + yv.off + c), v);
  }
}

// dx = dy routed to the first window position (row-major) whose value equals the max
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ dy, View dv, const T* __restrict__ x, View xv,
                                   const T* __restrict__ y, View yv, T* __restrict__ dx, View ov, int N, int Ho, int Wo,
                                   int C, int relu) {
  long total = (long)N * Ho * Wo * C;
  const int W = Wo * 2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long q = i / C;
    int wo = (int)(q % Wo);
    long t = q / Wo;
    int ho = (int)(t % Ho);
    long n = t / Ho;
    long p00 = (n * 2 * Ho + 2 * ho) * W + 2 * wo;
    float m = ld_f(y + (q * yv.pitch + yv.off + c));
    float g = ld_f(dy + (q * dv.pitch + dv.off + c));
    long pos[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
    bool done = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = ld_f(x + (pos[k] * xv.pitch + xv.off + c));
      bool hit = !done && v == m;
      st_f(dx + (pos[k] * ov.pitch + ov.off + c), (hit && !(relu && v <= 0.f)) ? g : 0.f);
      done = done || hit;
    }
  }
}

// y[n,2h+i,2w+j,c] = relu(x[n,h,w,c])
template <typename T>
__global__ void upsample_relu_fwd_kernel(const T* __restrict__ x, View xv, T* __restrict__ y, View yv, int N, int H,
                                         int W, int C) {
  long total = (long)N * 2 * H * 2 * W * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long q = i / C;
    int w2 = (int)(q % (2 * W));
    long t = q / (2 * W);
    int h2 = (int)(t % (2 * H));
    long n = t / (2 * H);
    long p = (n * H + h2 / 2) * W + w2 / 2;
    float v = ld_f(x + (p * xv.pitch + xv.off + c));
    st_f(y + (q * yv.pitch + yv.off + c), v > 0.f ? v : 0.f);
  }
}

// dx[n,h,w,c] = (x > 0) * sum_{i,j} dy[n,2h+i,2w+j,c]
template <typename T>
__global__ void upsample_relu_bwd_kernel(const T* __restrict__ dy, View dv, const T* __restrict__ x, View xv,
                                         T* __restrict__ dx, View ov, int N, int H, int W, int C) {
  long total = (long)N * H * W * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long p = i / C;
    int w = (int)(p % W);
    long t = p / W;
    int h = (int)(t % H);
    long n = t / H;
    long q00 = (n * 2 * H + 2 * h) * 2 * W + 2 * w;
    float g = ld_f(dy + (q00 * dv.pitch + dv.off + c)) + ld_f(dy + ((q00 + 1) * dv.pitch + dv.off + c)) +
              ld_f(dy + ((q00 + 2 * W) * dv.pitch + dv.off + c)) + ld_f(dy + ((q00 + 2 * W + 1) * dv.pitch + dv.off + c));
    float v = ld_f(x + (p * xv.pitch + xv.off + c));
    st_f(dx + (p * ov.pitch + ov.off + c), v > 0.f ? g : 0.f);
  }
}

// ------------------------------------------------------------------ VGG19 'caffe' preprocessing
// out[p][c'] = (in[p][2-c'] + 1) * 127.5 - mean_bgr[c']      (srgan.py:71-72)
template <typename TI, typename TO>
__global__ void vgg_pre_fwd_kernel(const TI* __restrict__ x, View xv, TO* __restrict__ y, View yv, long P) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < P * 3; i += (long)gridDim.x * blockDim.x) {
    long p = i / 3;
    int c = (int)(i - p * 3);
    const float mean = c == 0 ? 103.939f : (c == 1 ? 116.779f : 123.68f);
    float v = ((ld_f(x + (p * xv.pitch + xv.off + 2 - c)) + 1.0f) * 255.0f) / 2.0f - mean;
    st_f(y + (p * yv.pitch + yv.off + c), v);
  }
}
// dx[p][c] = 127.5 * dy[p][2-c]
template <typename TI, typename TO>
__global__ void vgg_pre_bwd_kernel(const TI* __restrict__ dy, View dv, TO* __restrict__ dx, View ov, long P) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < P * 3; i += (long)gridDim.x * blockDim.x) {
    long p = i / 3;
    int c = (int)(i - p * 3);
    st_f(dx + (p * ov.pitch + ov.off + c), 127.5f * ld_f(dy + (p * dv.pitch + dv.off + 2 - c)));
  }
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS) bias_grad_kernel(const T* __restrict__ dy, View dv, long P, int C, float* __restrict__ partial) {
  channel_reduce<1>(P, C, partial, [&](long p, int c, float* a) { a[0] += ld_f(dy + (p * dv.pitch + dv.off + c)); });
}

inline View view_of(const dg_tensor* t) { return View{t->cpitch, t->coff}; }
inline unsigned ew_blocks(long total, int sm_count) {
  long b = (total + 255) / 256;
  long cap = (long)sm_count * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

#define ST ((cudaStream_t)stream)

// the per-channel reductions need a little over 48 KB of dynamic shared memory at C >= 512
template <typename K>
static inline void red8_optin(K kernel, size_t smem) {
  if (smem + 64 > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

extern "C" size_t dg_bn_workspace_bytes(const dg_tensor* x) {
  // partials [blocks][3][C] + coef [2][C]; blocks <= 148*8 (sized for any device <= 256 SMs)
  return ((size_t)256 * 8 * 3 * x->c + 2 * (size_t)x->c) * sizeof(float);
}

extern "C" int dg_bn_stats(dg_ctx* ctx, const dg_tensor* x, const float* gamma, const float* beta, float eps,
                           float momentum, float* moving_mean, float* moving_var, float* scale, float* shift,
                           float* save_mean, float* save_invstd, void* workspace, size_t workspace_bytes,
                           void* stream) {
  DG_REQUIRE(dg_valid(x) && gamma && beta && scale && shift && save_mean && save_invstd && workspace,
             "dg_bn_stats: null argument");
  DG_REQUIRE(x->c <= RED_THREADS * MAX_CPT, "dg_bn_stats: C=%d too large", x->c);
  DG_REQUIRE(workspace_bytes >= dg_bn_workspace_bytes(x), "dg_bn_stats: workspace too small");
  long P = dg_pixels(x);
  int C = x->c;
  int blocks = red_blocks(P, C, ctx->sm_count);
  int R = RED_THREADS / red_lanes(C);
  size_t smem = (size_t)R * 2 * C * sizeof(float);
  float* partial = (float*)workspace;
  if (dgvec::vec_ok(x)) {
    blocks = dgvec::red8_blocks(P, C, ctx->sm_count);
    DG_DISPATCH_1(x->dtype, "dg_bn_stats",
                  red8_optin(dgvec::bn_stats8_kernel<T>, dgvec::red8_smem(C, 2));
                  dg_pdl_launch(dgvec::bn_stats8_kernel<T>, dim3(blocks), dim3(dgvec::RT), dgvec::red8_smem(C, 2), ST,
                      (const T*)x->ptr, dgvec::VView{x->cpitch, x->coff}, P, C, partial, ctx->tickets, gamma, beta, eps, momentum,
                      moving_mean, moving_var, scale, shift, save_mean, save_invstd););
    DG_CHECK_LAUNCH("dg_bn_stats");
    return 0;
  } else {
    DG_DISPATCH_1(x->dtype, "dg_bn_stats",
                  bn_stats_kernel<T><<<blocks, RED_THREADS, smem, ST>>>((const T*)x->ptr, view_of(x), P, C, partial););
  }
  bn_finalize_kernel<<<(C + 7) / 8, 256, 0, ST>>>(partial, blocks, P, C, gamma, beta, eps, momentum, moving_mean,
                                                      moving_var, scale, shift, save_mean, save_invstd);
  DG_CHECK_LAUNCH("dg_bn_stats");
  return 0;
}

// Second half of dg_bn_stats for statistics whose per-block partials [nblocks][2][C] (sum, sum of squares) were produced
// by the convolution epilogue (dg_umma_conv2d_fwd with bn_partials): fixed-order double-precision sum, scale/shift,
// saved mean / invstd and the moving-statistics update.  `pixels` = N*H*W of the normalised tensor.
extern "C" int dg_bn_finalize(dg_ctx* ctx, const float* partials, int nblocks, long long pixels, int c, const float* gamma,
                              const float* beta, float eps, float momentum, float* moving_mean, float* moving_var, float* scale,
                              float* shift, float* save_mean, float* save_invstd, void* stream) {
  DG_REQUIRE(partials && nblocks > 0 && pixels > 0 && c > 0 && gamma && beta && scale && shift && save_mean && save_invstd,
             "dg_bn_finalize: null argument");
  bn_finalize_kernel<<<(c + 7) / 8, 256, 0, ST>>>(partials, nblocks, (long)pixels, c, gamma, beta, eps, momentum, moving_mean,
                                                     moving_var, scale, shift, save_mean, save_invstd);
  DG_CHECK_LAUNCH("dg_bn_finalize");
  return 0;
}

// dg_bn_finalize + dg_bn_act_fwd in ONE launch (no dropout): the blocks of the apply pass each reduce the per-CTA statistics rows
// themselves.  rc 2: the tensors do not qualify for the vector kernel (the caller issues the two separate calls).
extern "C" int dg_bn_act_fwd_from_partials(dg_ctx* ctx, const dg_tensor* x, const float* partials, int nblocks, const float* gamma,
                                           const float* beta, float eps, float momentum, float* moving_mean, float* moving_var, float* scale,
                                           float* shift, float* save_mean, float* save_invstd, int act, float act_alpha,
                                           const float* prelu_alpha, const dg_tensor* residual, const dg_tensor* y, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y) && partials && nblocks > 0 && gamma && beta && scale && shift && save_mean && save_invstd,
             "dg_bn_act_fwd_from_partials: null argument");
  DG_REQUIRE(dg_same_shape(x, y), "dg_bn_act_fwd_from_partials: shape mismatch");
  DG_REQUIRE(act != DG_ACT_PRELU || prelu_alpha, "dg_bn_act_fwd_from_partials: PReLU needs alpha");
  if (residual) DG_REQUIRE(dg_valid(residual) && dg_same_shape(residual, y) && residual->dtype == y->dtype,
                           "dg_bn_act_fwd_from_partials: residual mismatch");
  const long P = dg_pixels(x);
  const int C = x->c;
  if (!(dgvec::vec_ok(x) && dgvec::vec_ok(y) && (!residual || dgvec::vec_ok(residual)) && dgvec::RT % (C >> 3) == 0)) return 2;
  const int E = 2 * C, E4 = E >> 2, G = E4 < dgvec::RT ? dgvec::RT / E4 : 1;
  const size_t smem = (size_t)G * E * sizeof(double) + (size_t)2 * C * sizeof(float);
  if (smem > 48 * 1024) return 2;
  const int vblocks = dgvec::red8_blocks(P, C, ctx->sm_count);
  View rv = residual ? view_of(residual) : View{0, 0};
  DG_DISPATCH_2(x->dtype, y->dtype, "dg_bn_act_fwd_from_partials",
                dg_pdl_launch(dgvec::bn_fwd_part8_kernel<TI, TO>, dim3(vblocks), dim3(dgvec::RT), smem, ST, (const TI*)x->ptr,
                              dgvec::VView{x->cpitch, x->coff}, P, C, partials, nblocks, gamma, beta, eps, momentum, moving_mean, moving_var,
                              scale, shift, save_mean, save_invstd, act, act_alpha, prelu_alpha,
                              residual ? (const TO*)residual->ptr : nullptr, dgvec::VView{rv.pitch, rv.off}, (TO*)y->ptr,
                              dgvec::VView{y->cpitch, y->coff}););
  DG_CHECK_LAUNCH("dg_bn_act_fwd_from_partials");
  return 0;
}

extern "C" int dg_bn_infer_affine(dg_ctx* ctx, int c, const float* gamma, const float* beta, const float* moving_mean,
                                  const float* moving_var, float eps, float* scale, float* shift, void* stream) {
  DG_REQUIRE(c > 0 && gamma && beta && moving_mean && moving_var && scale && shift, "dg_bn_infer_affine: null argument");
  bn_infer_affine_kernel<<<(c + 127) / 128, 128, 0, ST>>>(c, gamma, beta, moving_mean, moving_var, eps, scale, shift);
  DG_CHECK_LAUNCH("dg_bn_infer_affine");
  return 0;
}

extern "C" int dg_bn_act_fwd(dg_ctx* ctx, const dg_tensor* x, const float* scale, const float* shift, int act,
                             float act_alpha, const float* prelu_alpha, const dg_tensor* residual, int dropout,
                             uint32_t seed, uint32_t offset, const int64_t* step_counter, const dg_tensor* y, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y) && scale && shift, "dg_bn_act_fwd: null argument");
  DG_REQUIRE(dg_same_shape(x, y), "dg_bn_act_fwd: shape mismatch");
  DG_REQUIRE(act != DG_ACT_PRELU || prelu_alpha, "dg_bn_act_fwd: PReLU needs alpha");
  if (residual) DG_REQUIRE(dg_valid(residual) && dg_same_shape(residual, y) && residual->dtype == y->dtype,
                           "dg_bn_act_fwd: residual mismatch");
  long P = dg_pixels(x);
  int C = x->c;
  View rv = residual ? view_of(residual) : View{0, 0};
  if (dgvec::vec_ok(x) && dgvec::vec_ok(y) && (!residual || dgvec::vec_ok(residual))) {
    DG_DISPATCH_2(x->dtype, y->dtype, "dg_bn_act_fwd",
                  dg_pdl_launch(dgvec::bn_act_fwd8_kernel<TI, TO>, dim3(dgvec::ewc_blocks(P, C, ctx->sm_count)), dim3(dgvec::ET), 0, ST,
                      (const TI*)x->ptr, dgvec::VView{x->cpitch, x->coff}, scale, shift, act, act_alpha, prelu_alpha,
                      residual ? (const TO*)residual->ptr : nullptr, dgvec::VView{rv.pitch, rv.off}, dropout, seed, offset, step_counter,
                      (TO*)y->ptr, dgvec::VView{y->cpitch, y->coff}, P, C););
    DG_CHECK_LAUNCH("dg_bn_act_fwd");
    return 0;
  }
  DG_DISPATCH_2(x->dtype, y->dtype, "dg_bn_act_fwd",
                bn_act_fwd_kernel<TI, TO><<<ew_blocks(P * C, ctx->sm_count), 256, 0, ST>>>(
                    (const TI*)x->ptr, view_of(x), scale, shift, act, act_alpha, prelu_alpha,
                    residual ? (const TO*)residual->ptr : nullptr, rv, dropout, seed, offset, step_counter, (TO*)y->ptr, view_of(y), P, C););
  DG_CHECK_LAUNCH("dg_bn_act_fwd");
  return 0;
}

// dg_bn_stats + dg_bn_act_fwd in one launch (training mode).  Returns 2 when the tensors do not qualify for the fused
// vector kernel: the caller then issues the two calls (never a silent fallback inside the library).
extern "C" int dg_bn_train_fwd(dg_ctx* ctx, const dg_tensor* x, const float* gamma, const float* beta, float eps, float momentum,
                               float* moving_mean, float* moving_var, float* scale, float* shift, float* save_mean,
                               float* save_invstd, int act, float act_alpha, const float* prelu_alpha, const dg_tensor* residual,
                               int dropout, uint32_t seed, uint32_t offset, const int64_t* step_counter, const dg_tensor* y,
                               void* workspace, size_t workspace_bytes, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y) && gamma && beta && scale && shift && save_mean && save_invstd && workspace,
             "dg_bn_train_fwd: null argument");
  DG_REQUIRE(dg_same_shape(x, y), "dg_bn_train_fwd: shape mismatch");
  DG_REQUIRE(act != DG_ACT_PRELU || prelu_alpha, "dg_bn_train_fwd: PReLU needs alpha");
  if (residual) DG_REQUIRE(dg_valid(residual) && dg_same_shape(residual, y) && residual->dtype == y->dtype,
                           "dg_bn_train_fwd: residual mismatch");
  DG_REQUIRE(workspace_bytes >= dg_bn_workspace_bytes(x), "dg_bn_train_fwd: workspace too small");
  if (!(dgvec::vec_ok(x) && dgvec::vec_ok(y) && (!residual || dgvec::vec_ok(residual)))) return 2;
  const long P = dg_pixels(x);
  const int C = x->c;
  const int blocks = dgvec::red8_blocks(P, C, ctx->sm_count);
  View rv = residual ? view_of(residual) : View{0, 0};
  DG_DISPATCH_2(x->dtype, y->dtype, "dg_bn_train_fwd", {
    red8_optin(dgvec::bn_fwd_fused8_kernel<TI, TO>, dgvec::red8_smem(C, 2));
    if (!dg_coresident(dgvec::bn_fwd_fused8_kernel<TI, TO>, dgvec::RT, dgvec::red8_smem(C, 2), blocks, ctx->sm_count)) return 2;   // grid barrier inside
    dg_coop_launch(dgvec::bn_fwd_fused8_kernel<TI, TO>, dim3(blocks), dim3(dgvec::RT), dgvec::red8_smem(C, 2), ST, (const TI*)x->ptr,
                  dgvec::VView{x->cpitch, x->coff}, P, C, (float*)workspace, ctx->tickets, gamma, beta, eps, momentum, moving_mean,
                  moving_var, scale, shift, save_mean, save_invstd, act, act_alpha, prelu_alpha,
                  residual ? (const TO*)residual->ptr : nullptr, dgvec::VView{rv.pitch, rv.off}, dropout, seed, offset, step_counter,
                  (TO*)y->ptr, dgvec::VView{y->cpitch, y->coff});
  });
  DG_CHECK_LAUNCH("dg_bn_train_fwd");
  return 0;
}

extern "C" int dg_bn_act_bwd(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* x, const float* scale,
                             const float* shift, const float* gamma, const float* save_mean, const float* save_invstd,
                             int act, float act_alpha, const float* prelu_alpha, int dropout, uint32_t seed,
                             uint32_t offset, const int64_t* step_counter, const dg_tensor* dx, float* dgamma, float* dbeta, float* dprelu_alpha,
                             int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(x) && dg_valid(dx) && scale && shift && gamma && save_mean && save_invstd &&
                 workspace, "dg_bn_act_bwd: null argument");
  DG_REQUIRE(dg_same_shape(dy, x) && dg_same_shape(dx, x), "dg_bn_act_bwd: shape mismatch");
  DG_REQUIRE(dy->dtype == dx->dtype, "dg_bn_act_bwd: dy/dx dtype mismatch");
  DG_REQUIRE(x->c <= RED_THREADS * MAX_CPT, "dg_bn_act_bwd: C too large");
  DG_REQUIRE(workspace_bytes >= dg_bn_workspace_bytes(x), "dg_bn_act_bwd: workspace too small");
  long P = dg_pixels(x);
  int C = x->c;
  int blocks = red_blocks(P, C, ctx->sm_count);
  int R = RED_THREADS / red_lanes(C);
  size_t smem = (size_t)R * 3 * C * sizeof(float);
  float* partial = (float*)workspace;
  float* coef = partial + (size_t)256 * 8 * 3 * C;
  if (dgvec::vec_ok(dy) && dgvec::vec_ok(x) && dgvec::vec_ok(dx)) {
    const int vblocks = dgvec::red8_blocks(P, C, ctx->sm_count);
    const dgvec::VView vdy{dy->cpitch, dy->coff}, vx{x->cpitch, x->coff}, vdx{dx->cpitch, dx->coff};
    static const char* env_fused = getenv("DG_BN_FUSED_BWD");
    const bool fused_bwd = !(env_fused && env_fused[0] == '0');
#define DG_BN_BWD_VEC(AM)                                                                                                        \
  if (fused_bwd && (red8_optin(dgvec::bn_bwd_fused8_kernel<TI, TO, TI, AM>, dgvec::red8_smem(C, 3)),                               \
                    dg_coresident(dgvec::bn_bwd_fused8_kernel<TI, TO, TI, AM>, dgvec::RT, dgvec::red8_smem(C, 3), vblocks, ctx->sm_count))) { \
    dg_coop_launch(dgvec::bn_bwd_fused8_kernel<TI, TO, TI, AM>, dim3(vblocks), dim3(dgvec::RT), dgvec::red8_smem(C, 3), ST,        \
        (const TI*)dy->ptr, vdy, (const TO*)x->ptr, vx, scale, shift, gamma, save_mean, save_invstd, act, act_alpha, prelu_alpha,  \
        dropout, seed, offset, step_counter, P, C, partial, ctx->tickets, dgamma, dbeta,                                         \
        act == DG_ACT_PRELU ? dprelu_alpha : nullptr, accumulate, coef, (TI*)dx->ptr, vdx);                                      \
  } else {                                                                                                                       \
    red8_optin(dgvec::bn_bwd_reduce8_kernel<TI, TO, AM>, dgvec::red8_smem(C, 3));                                                  \
    dg_pdl_launch(dgvec::bn_bwd_reduce8_kernel<TI, TO, AM>, dim3(vblocks), dim3(dgvec::RT), dgvec::red8_smem(C, 3), ST,           \
        (const TI*)dy->ptr, vdy, (const TO*)x->ptr, vx, scale, shift, save_mean, save_invstd, act, act_alpha, prelu_alpha, dropout, \
        seed, offset, step_counter, P, C, partial, ctx->tickets, dgamma, dbeta, act == DG_ACT_PRELU ? dprelu_alpha : nullptr,   \
        accumulate, coef);                                                                                                       \
    dg_pdl_launch(dgvec::bn_bwd_dx8_kernel<TI, TO, TI, AM>, dim3(dgvec::ewc_blocks(P, C, ctx->sm_count)), dim3(dgvec::ET), 0, ST, \
        (const TI*)dy->ptr, vdy, (const TO*)x->ptr, vx, scale, shift, gamma, save_mean, save_invstd, act, act_alpha, prelu_alpha,  \
        dropout, seed, offset, step_counter, coef, (TI*)dx->ptr, vdx, P, C);                                                     \
  }
    // on-chip variant: bf16 tensors, no dropout, every thread's pixel walk fits K = 16 cached iterations.  OFF by default
    // (DG_BN_BWD_CACHED=1 enables): measured SLOWER inside the SRGAN step, 7.76 vs 7.44 ms -- 128 registers with the x cache
    // leave four loads in flight per thread and 200 KB of shared memory keep every other kernel off the SM.
    {
      static const char* env_cached = getenv("DG_BN_BWD_CACHED");
      constexpr int KC = 16;
      const int Rr = dgvec::RT / (C >> 3);
      const long walk = ((long)P + (long)vblocks * Rr - 1) / ((long)vblocks * Rr);
      const int am = dgvec::act_mode(act, dropout);
      const size_t red = (dgvec::red8_smem(C, 3) + 15) & ~(size_t)15, smem_c = red + (size_t)KC * dgvec::RT * 16;
      if (fused_bwd && (env_cached && env_cached[0] == '1') && dy->dtype == DG_BF16 && x->dtype == DG_BF16 && am <= 3 && walk <= KC &&
          dgvec::RT % (C >> 3) == 0 && smem_c <= 200 * 1024) {
#define DG_BN_BWD_CACHED(AM)                                                                                                      \
  {                                                                                                                              \
    cudaFuncSetAttribute(dgvec::bn_bwd_cached8_kernel<AM, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);          \
    dg_coop_launch(dgvec::bn_bwd_cached8_kernel<AM, KC>, dim3(vblocks), dim3(dgvec::RT), smem_c, ST,                                \
        (const __nv_bfloat16*)dy->ptr, vdy, (const __nv_bfloat16*)x->ptr, vx, scale, shift, gamma, save_mean, save_invstd, act,    \
        act_alpha, prelu_alpha, P, C, partial, ctx->tickets, dgamma, dbeta, act == DG_ACT_PRELU ? dprelu_alpha : nullptr,        \
        accumulate, coef, (__nv_bfloat16*)dx->ptr, vdx, (unsigned)red);                                                          \
  }
        switch (am) {
          case 0: DG_BN_BWD_CACHED(0) break;
          case 1: DG_BN_BWD_CACHED(1) break;
          case 2: DG_BN_BWD_CACHED(2) break;
          default: DG_BN_BWD_CACHED(3) break;
        }
#undef DG_BN_BWD_CACHED
        DG_CHECK_LAUNCH("dg_bn_act_bwd");
        return 0;
      }
    }
    DG_DISPATCH_2(dy->dtype, x->dtype, "dg_bn_act_bwd", {
      switch (dgvec::act_mode(act, dropout)) {
        case 0: DG_BN_BWD_VEC(0) break;
        case 1: DG_BN_BWD_VEC(1) break;
        case 2: DG_BN_BWD_VEC(2) break;
        case 3: DG_BN_BWD_VEC(3) break;
        default: DG_BN_BWD_VEC(4) break;
      }
    });
#undef DG_BN_BWD_VEC
    DG_CHECK_LAUNCH("dg_bn_act_bwd");
    return 0;
  }
  DG_DISPATCH_2(dy->dtype, x->dtype, "dg_bn_act_bwd", {
    bn_bwd_reduce_kernel<TI, TO><<<blocks, RED_THREADS, smem, ST>>>(
        (const TI*)dy->ptr, view_of(dy), (const TO*)x->ptr, view_of(x), scale, shift, save_mean, save_invstd, act,
        act_alpha, prelu_alpha, dropout, seed, offset, step_counter, P, C, partial);
    bn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, ST>>>(partial, blocks, P, C, dgamma, dbeta,
                                                            act == DG_ACT_PRELU ? dprelu_alpha : nullptr, accumulate, coef);
    bn_bwd_dx_kernel<TI, TO, TI><<<ew_blocks(P * C, ctx->sm_count), 256, 0, ST>>>(
        (const TI*)dy->ptr, view_of(dy), (const TO*)x->ptr, view_of(x), scale, shift, gamma, save_mean, save_invstd,
        act, act_alpha, prelu_alpha, dropout, seed, offset, step_counter, coef, (TI*)dx->ptr, view_of(dx), P, C);
  });
  DG_CHECK_LAUNCH("dg_bn_act_bwd");
  return 0;
}

// dx half of the BatchNorm(+activation) backward pass from the per-CTA partial sums of dg_umma_conv2d_dgrad_fused.
extern "C" int dg_bn_bwd_dx_from_partials(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* x, const float* scale, const float* shift,
                                          const float* gamma, const float* save_mean, const float* save_invstd, int act, float act_alpha,
                                          const float* partials, int rows, const dg_tensor* dx, float* dgamma, float* dbeta, int accumulate,
                                          void* stream) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(x) && dg_valid(dx) && scale && shift && gamma && save_mean && save_invstd && partials && rows > 0,
             "dg_bn_bwd_dx_from_partials: null argument");
  DG_REQUIRE(dg_same_shape(dy, x) && dg_same_shape(dx, x), "dg_bn_bwd_dx_from_partials: shape mismatch");
  DG_REQUIRE(dy->dtype == dx->dtype, "dg_bn_bwd_dx_from_partials: dy/dx dtype mismatch");
  DG_REQUIRE(act == DG_ACT_NONE || act == DG_ACT_RELU || act == DG_ACT_LRELU, "dg_bn_bwd_dx_from_partials: activation must be none / relu / leaky relu");
  DG_REQUIRE(dgvec::vec_ok(dy) && dgvec::vec_ok(x) && dgvec::vec_ok(dx) && x->c % 8 == 0 && dgvec::RT % (x->c >> 3) == 0 && x->c <= 1024,
             "dg_bn_bwd_dx_from_partials: needs 16-byte aligned views with a channel count dividing 4096");
  const long P = dg_pixels(x);
  const int C = x->c;
  const int E = 2 * C, E4 = E >> 2, G = E4 < dgvec::RT ? dgvec::RT / E4 : 1;
  const size_t smem = (size_t)G * E * sizeof(double) + (size_t)2 * C * sizeof(float);
  const int vblocks = dgvec::red8_blocks(P, C, ctx->sm_count);
  const dgvec::VView vdy{dy->cpitch, dy->coff}, vx{x->cpitch, x->coff}, vdx{dx->cpitch, dx->coff};
#define DG_BN_DXP(AM)                                                                                                              \
  dg_pdl_launch(dgvec::bn_bwd_dx_part8_kernel<TI, TO, TI, AM>, dim3(vblocks), dim3(dgvec::RT), smem, ST, (const TI*)dy->ptr, vdy,  \
                (const TO*)x->ptr, vx, scale, shift, gamma, save_mean, save_invstd, act_alpha, partials, rows, P, C, dgamma, dbeta, \
                accumulate, (TI*)dx->ptr, vdx)
  DG_DISPATCH_2(dy->dtype, x->dtype, "dg_bn_bwd_dx_from_partials", {
    if (act == DG_ACT_NONE) DG_BN_DXP(0);
    else if (act == DG_ACT_RELU) DG_BN_DXP(1);
    else DG_BN_DXP(2);
  });
#undef DG_BN_DXP
  DG_CHECK_LAUNCH("dg_bn_bwd_dx_from_partials");
  return 0;
}

// ------------------------------------------------------------------ im2col for the small-spatial weight gradients
// out[p][t*C + c] = x[n, ho*stride + r - pad_t, wo*stride + s - pad_l, c] (0 outside the image), p = (n*Ho + ho)*Wo + wo,
// t = r*kw + s.  One 16-byte vector (8 channels) per thread.  pix2pix's bottleneck layers (pix2pix.py:147-166: 4x4 stride-2
// convolutions on 16x16 ... 1x1 maps) have so few pixels per image that the halo-tile weight gradient spends its MMAs on
// padding (a 16 x 8 tile per 2 x 2 image); gathered like this the weight gradient is ONE dense product over all N*Ho*Wo pixels.
__global__ void __launch_bounds__(256) im2col8_kernel(const __nv_bfloat16* __restrict__ x, int H, int W, int C, int pitch, int off, int kh,
                                                      int kw, int stride, int pad_t, int pad_l, int Ho, int Wo, long P,
                                                      __nv_bfloat16* __restrict__ out) {
  const int CV = C >> 3, taps = kh * kw;
  const long total = P * taps * CV;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % CV);
    const long r1 = i / CV;
    const int t = (int)(r1 % taps);
    const long p = r1 / taps;
    const int wo = (int)(p % Wo);
    const long r2 = p / Wo;
    const int ho = (int)(r2 % Ho);
    const long n = r2 / Ho;
    const int hi = ho * stride + t / kw - pad_t, wi = wo * stride + t % kw - pad_l;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = *reinterpret_cast<const uint4*>(x + ((n * H + hi) * W + wi) * pitch + off + c8 * 8);
    *reinterpret_cast<uint4*>(out + (p * taps + t) * C + c8 * 8) = v;
  }
}

extern "C" int dg_im2col(dg_ctx* ctx, const dg_tensor* x, const dg_conv_params* p, int out_h, int out_w, void* out, void* stream) {
  DG_REQUIRE(dg_valid(x) && p && out && out_h > 0 && out_w > 0, "dg_im2col: null argument");
  DG_REQUIRE(x->dtype == DG_BF16 && x->c % 8 == 0 && x->cpitch % 8 == 0 && x->coff % 8 == 0 && ((uintptr_t)x->ptr % 16) == 0 &&
                 ((uintptr_t)out % 16) == 0, "dg_im2col: needs bf16 views with 8-channel (16-byte) alignment");
  const long P = (long)x->n * out_h * out_w;
  const long total = P * p->kh * p->kw * (x->c >> 3);
  long blocks = (total + 255) / 256;
  if (blocks > (long)ctx->sm_count * 16) blocks = (long)ctx->sm_count * 16;
  im2col8_kernel<<<(unsigned)blocks, 256, 0, ST>>>((const __nv_bfloat16*)x->ptr, x->h, x->w, x->c, x->cpitch, x->coff, p->kh, p->kw, p->stride,
                                                  p->pad_t, p->pad_l, out_h, out_w, P, (__nv_bfloat16*)out);
  DG_CHECK_LAUNCH("dg_im2col");
  return 0;
}

extern "C" int dg_act_bwd_from_output(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* y, int act, float act_alpha,
                                      const dg_tensor* dpre, void* stream) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(y) && dg_valid(dpre), "dg_act_bwd_from_output: null argument");
  DG_REQUIRE(dg_same_shape(dy, y) && dg_same_shape(dpre, y), "dg_act_bwd_from_output: shape mismatch");
  DG_REQUIRE(dy->dtype == dpre->dtype, "dg_act_bwd_from_output: dy/dpre dtype mismatch");
  DG_REQUIRE(act != DG_ACT_PRELU, "dg_act_bwd_from_output: PReLU needs the pre-activation");
  long P = dg_pixels(y);
  int C = y->c;
  if (dgvec::vec_ok(dy) && dgvec::vec_ok(y) && dgvec::vec_ok(dpre)) {
    DG_DISPATCH_2(dy->dtype, y->dtype, "dg_act_bwd_from_output",
                  dgvec::act_bwd8_kernel<TI, TO, TI><<<dgvec::ew8_blocks(P * (C / 8), ctx->sm_count), dgvec::VT, 0, ST>>>(
                      (const TI*)dy->ptr, dgvec::VView{dy->cpitch, dy->coff}, (const TO*)y->ptr, dgvec::VView{y->cpitch, y->coff}, act,
                      act_alpha, (TI*)dpre->ptr, dgvec::VView{dpre->cpitch, dpre->coff}, P, C););
    DG_CHECK_LAUNCH("dg_act_bwd_from_output");
    return 0;
  }
  DG_DISPATCH_2(dy->dtype, y->dtype, "dg_act_bwd_from_output",
                act_bwd_kernel<TI, TO, TI><<<ew_blocks(P * C, ctx->sm_count), 256, 0, ST>>>(
                    (const TI*)dy->ptr, view_of(dy), (const TO*)y->ptr, view_of(y), act, act_alpha, (TI*)dpre->ptr,
                    view_of(dpre), P, C););
  DG_CHECK_LAUNCH("dg_act_bwd_from_output");
  return 0;
}

extern "C" int dg_d2s_prelu_fwd(dg_ctx* ctx, const dg_tensor* u, const float* prelu_alpha, const dg_tensor* y,
                                void* stream) {
  DG_REQUIRE(dg_valid(u) && dg_valid(y), "dg_d2s_prelu_fwd: null argument");
  DG_REQUIRE(u->c % 4 == 0 && y->c * 4 == u->c && y->h == 2 * u->h && y->w == 2 * u->w && y->n == u->n,
             "dg_d2s_prelu_fwd: shape mismatch");
  DG_REQUIRE(u->dtype == y->dtype, "dg_d2s_prelu_fwd: dtype mismatch");
  long total = dg_pixels(u) * u->c;
  if (dgvec::vec_ok(u) && dgvec::vec_ok(y)) {
    DG_DISPATCH_1(u->dtype, "dg_d2s_prelu_fwd",
                  dgvec::d2s_prelu8_kernel<T, false><<<dgvec::ew8_blocks(total / 8, ctx->sm_count), dgvec::VT, 0, ST>>>(
                      (const T*)u->ptr, dgvec::VView{u->cpitch, u->coff}, nullptr, dgvec::VView{0, 0}, prelu_alpha, (T*)y->ptr,
                      dgvec::VView{y->cpitch, y->coff}, u->n, u->h, u->w, y->c););
    DG_CHECK_LAUNCH("dg_d2s_prelu_fwd");
    return 0;
  }
  DG_DISPATCH_1(u->dtype, "dg_d2s_prelu_fwd",
                d2s_prelu_fwd_kernel<T><<<ew_blocks(total, ctx->sm_count), 256, 0, ST>>>(
                    (const T*)u->ptr, view_of(u), prelu_alpha, (T*)y->ptr, view_of(y), u->n, u->h, u->w, y->c););
  DG_CHECK_LAUNCH("dg_d2s_prelu_fwd");
  return 0;
}

extern "C" int dg_d2s_prelu_bwd(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* u, const float* prelu_alpha,
                                const dg_tensor* du, float* dprelu_alpha, int accumulate, void* workspace,
                                size_t workspace_bytes, void* stream) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(u) && dg_valid(du), "dg_d2s_prelu_bwd: null argument");
  DG_REQUIRE(dy->c * 4 == u->c && dy->h == 2 * u->h && dy->w == 2 * u->w && dg_same_shape(u, du),
             "dg_d2s_prelu_bwd: shape mismatch");
  DG_REQUIRE(dy->dtype == u->dtype && du->dtype == u->dtype, "dg_d2s_prelu_bwd: dtype mismatch");
  long total = dg_pixels(u) * u->c;
  int Co = dy->c;
  const bool vec = dgvec::vec_ok(dy) && dgvec::vec_ok(u) && dgvec::vec_ok(du);
  if (vec && prelu_alpha && dprelu_alpha) {   // one pass: du and the slope gradient together
    DG_REQUIRE(workspace && workspace_bytes >= dg_bn_workspace_bytes(dy), "dg_d2s_prelu_bwd: workspace too small");
    long Pout = dg_pixels(dy);
    int vblocks = dgvec::red8_blocks(Pout, Co, ctx->sm_count);
    DG_DISPATCH_1(u->dtype, "dg_d2s_prelu_bwd",
                  dgvec::d2s_prelu_bwd_fused8_kernel<T><<<vblocks, dgvec::RT, dgvec::red8_smem(Co, 1), ST>>>(
                      (const T*)dy->ptr, dgvec::VView{dy->cpitch, dy->coff}, (const T*)u->ptr, dgvec::VView{u->cpitch, u->coff},
                      prelu_alpha, (T*)du->ptr, dgvec::VView{du->cpitch, du->coff}, Pout, dy->h, dy->w, Co, (float*)workspace,
                      ctx->tickets, dprelu_alpha, accumulate););
    DG_CHECK_LAUNCH("dg_d2s_prelu_bwd");
    return 0;
  }
  DG_DISPATCH_1(u->dtype, "dg_d2s_prelu_bwd", {
    if (vec)
      dgvec::d2s_prelu8_kernel<T, true><<<dgvec::ew8_blocks(total / 8, ctx->sm_count), dgvec::VT, 0, ST>>>(
          (const T*)dy->ptr, dgvec::VView{dy->cpitch, dy->coff}, (const T*)u->ptr, dgvec::VView{u->cpitch, u->coff}, prelu_alpha,
          (T*)du->ptr, dgvec::VView{du->cpitch, du->coff}, u->n, u->h, u->w, Co);
    else
      d2s_prelu_bwd_kernel<T><<<ew_blocks(total, ctx->sm_count), 256, 0, ST>>>(
          (const T*)dy->ptr, view_of(dy), (const T*)u->ptr, view_of(u), prelu_alpha, (T*)du->ptr, view_of(du), u->n, u->h,
          u->w, Co);
    if (prelu_alpha && dprelu_alpha && vec) {
      DG_REQUIRE(workspace && workspace_bytes >= dg_bn_workspace_bytes(dy), "dg_d2s_prelu_bwd: workspace too small");
      long Pout = dg_pixels(dy);
      int vblocks = dgvec::red8_blocks(Pout, Co, ctx->sm_count);
      float* partial = (float*)workspace;
      dgvec::d2s_dalpha8_kernel<T><<<vblocks, dgvec::RT, dgvec::red8_smem(Co, 1), ST>>>(
          (const T*)dy->ptr, dgvec::VView{dy->cpitch, dy->coff}, (const T*)u->ptr, dgvec::VView{u->cpitch, u->coff}, Pout, dy->h,
          dy->w, Co, partial, ctx->tickets, dprelu_alpha, accumulate);
    } else if (prelu_alpha && dprelu_alpha) {
      DG_REQUIRE(workspace && workspace_bytes >= dg_bn_workspace_bytes(dy), "dg_d2s_prelu_bwd: workspace too small");
      long Pout = dg_pixels(dy);
      int blocks = red_blocks(Pout, Co, ctx->sm_count);
      int R = RED_THREADS / red_lanes(Co);
      float* partial = (float*)workspace;
      d2s_prelu_dalpha_kernel<T><<<blocks, RED_THREADS, (size_t)R * Co * sizeof(float), ST>>>(
          (const T*)dy->ptr, view_of(dy), (const T*)u->ptr, view_of(u), Pout, dy->h, dy->w, Co, partial);
      sum_partials_kernel<<<(Co + 7) / 8, 256, 0, ST>>>(partial, blocks, Co, dprelu_alpha, accumulate);
    }
  });
  DG_CHECK_LAUNCH("dg_d2s_prelu_bwd");
  return 0;
}

extern "C" int dg_add(dg_ctx* ctx, const dg_tensor* a, const dg_tensor* b, const dg_tensor* out, void* stream) {
  DG_REQUIRE(dg_valid(a) && dg_valid(b) && dg_valid(out), "dg_add: null argument");
  DG_REQUIRE(dg_same_shape(a, b) && dg_same_shape(a, out) && a->dtype == b->dtype, "dg_add: shape/dtype mismatch");
  long P = dg_pixels(a);
  if (dgvec::vec_ok(a) && dgvec::vec_ok(b) && dgvec::vec_ok(out)) {
    DG_DISPATCH_2(a->dtype, out->dtype, "dg_add",
                  dgvec::ew8_kernel<TI, TO, 1><<<dgvec::ew8_blocks(P * (a->c / 8), ctx->sm_count), dgvec::VT, 0, ST>>>(
                      (const TI*)a->ptr, dgvec::VView{a->cpitch, a->coff}, (const TI*)b->ptr, dgvec::VView{b->cpitch, b->coff},
                      (TO*)out->ptr, dgvec::VView{out->cpitch, out->coff}, P, a->c););
    DG_CHECK_LAUNCH("dg_add");
    return 0;
  }
  DG_DISPATCH_2(a->dtype, out->dtype, "dg_add",
                add_kernel<TI, TO><<<ew_blocks(P * a->c, ctx->sm_count), 256, 0, ST>>>(
                    (const TI*)a->ptr, view_of(a), (const TI*)b->ptr, view_of(b), (TO*)out->ptr, view_of(out), P, a->c););
  DG_CHECK_LAUNCH("dg_add");
  return 0;
}

extern "C" int dg_copy(dg_ctx* ctx, const dg_tensor* src, const dg_tensor* out, int accumulate, void* stream) {
  DG_REQUIRE(dg_valid(src) && dg_valid(out), "dg_copy: null argument");
  DG_REQUIRE(dg_same_shape(src, out), "dg_copy: shape mismatch");
  long P = dg_pixels(src);
  if (dgvec::vec_ok(src) && dgvec::vec_ok(out)) {
    const dgvec::VView vs{src->cpitch, src->coff}, vo{out->cpitch, out->coff};
    const unsigned nblk = dgvec::ew8_blocks(P * (src->c / 8), ctx->sm_count);
    DG_DISPATCH_2(src->dtype, out->dtype, "dg_copy", {
      if (accumulate)
        dgvec::ew8_kernel<TI, TO, 2><<<nblk, dgvec::VT, 0, ST>>>((const TI*)src->ptr, vs, nullptr, vs, (TO*)out->ptr, vo, P, src->c);
      else
        dgvec::ew8_kernel<TI, TO, 0><<<nblk, dgvec::VT, 0, ST>>>((const TI*)src->ptr, vs, nullptr, vs, (TO*)out->ptr, vo, P, src->c);
    });
    DG_CHECK_LAUNCH("dg_copy");
    return 0;
  }
  DG_DISPATCH_2(src->dtype, out->dtype, "dg_copy",
                copy_kernel<TI, TO><<<ew_blocks(P * src->c, ctx->sm_count), 256, 0, ST>>>(
                    (const TI*)src->ptr, view_of(src), (TO*)out->ptr, view_of(out), P, src->c, accumulate););
  DG_CHECK_LAUNCH("dg_copy");
  return 0;
}

extern "C" int dg_maxpool2x2_fwd(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* y, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y), "dg_maxpool2x2_fwd: null argument");
  DG_REQUIRE(x->h == 2 * y->h && x->w == 2 * y->w && x->c == y->c && x->n == y->n && x->dtype == y->dtype,
             "dg_maxpool2x2_fwd: shape mismatch (even sizes only)");
  long total = dg_pixels(y) * y->c;
  if (dgvec::vec_ok(x) && dgvec::vec_ok(y)) {
    const dgvec::VView vx{x->cpitch, x->coff}, vy{y->cpitch, y->coff};
    DG_DISPATCH_1(x->dtype, "dg_maxpool2x2_fwd",
                  dgvec::maxpool_fwd8_kernel<T><<<dgvec::ew8_blocks(total / 8, ctx->sm_count), dgvec::VT, 0, ST>>>(
                      (const T*)x->ptr, vx, (T*)y->ptr, vy, y->n, y->h, y->w, y->c););
    DG_CHECK_LAUNCH("dg_maxpool2x2_fwd");
    return 0;
  }
  DG_DISPATCH_1(x->dtype, "dg_maxpool2x2_fwd",
                maxpool_fwd_kernel<T><<<ew_blocks(total, ctx->sm_count), 256, 0, ST>>>(
                    (const T*)x->ptr, view_of(x), (T*)y->ptr, view_of(y), y->n, y->h, y->w, y->c););
  DG_CHECK_LAUNCH("dg_maxpool2x2_fwd");
  return 0;
}

static int maxpool_bwd_impl(const char* name, dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* x, const dg_tensor* y, const dg_tensor* dx,
                            int relu, void* stream) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(x) && dg_valid(y) && dg_valid(dx), "%s: null argument", name);
  DG_REQUIRE(dg_same_shape(dy, y) && dg_same_shape(dx, x) && x->h == 2 * y->h && x->w == 2 * y->w, "%s: shape mismatch", name);
  DG_REQUIRE(dy->dtype == x->dtype && y->dtype == x->dtype && dx->dtype == x->dtype, "%s: dtype mismatch", name);
  long total = dg_pixels(y) * y->c;
  if (dgvec::vec_ok(dy) && dgvec::vec_ok(x) && dgvec::vec_ok(y) && dgvec::vec_ok(dx)) {
    const dgvec::VView vd{dy->cpitch, dy->coff}, vx{x->cpitch, x->coff}, vy{y->cpitch, y->coff}, vo{dx->cpitch, dx->coff};
    DG_DISPATCH_1(x->dtype, name,
                  dgvec::maxpool_bwd8_kernel<T><<<dgvec::ew8_blocks(total / 8, ctx->sm_count), dgvec::VT, 0, ST>>>(
                      (const T*)dy->ptr, vd, (const T*)x->ptr, vx, (const T*)y->ptr, vy, (T*)dx->ptr, vo, y->n, y->h, y->w, y->c, relu););
    DG_CHECK_LAUNCH(name);
    return 0;
  }
  DG_DISPATCH_1(x->dtype, name,
                maxpool_bwd_kernel<T><<<ew_blocks(total, ctx->sm_count), 256, 0, ST>>>(
                    (const T*)dy->ptr, view_of(dy), (const T*)x->ptr, view_of(x), (const T*)y->ptr, view_of(y),
                    (T*)dx->ptr, view_of(dx), y->n, y->h, y->w, y->c, relu););
  DG_CHECK_LAUNCH(name);
  return 0;
}

extern "C" int dg_maxpool2x2_bwd(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* x, const dg_tensor* y,
                                 const dg_tensor* dx, void* stream) {
  return maxpool_bwd_impl("dg_maxpool2x2_bwd", ctx, dy, x, y, dx, 0, stream);
}

extern "C" int dg_maxpool2x2_bwd_relu(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* x, const dg_tensor* y,
                                      const dg_tensor* dx, void* stream) {
  return maxpool_bwd_impl("dg_maxpool2x2_bwd_relu", ctx, dy, x, y, dx, 1, stream);
}

extern "C" int dg_upsample2x_relu_fwd(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* y, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y), "dg_upsample2x_relu_fwd: null argument");
  DG_REQUIRE(y->h == 2 * x->h && y->w == 2 * x->w && x->c == y->c && x->n == y->n && x->dtype == y->dtype,
             "dg_upsample2x_relu_fwd: shape mismatch");
  long total = dg_pixels(y) * y->c;
  if (dgvec::vec_ok(x) && dgvec::vec_ok(y)) {
    const dgvec::VView vx{x->cpitch, x->coff}, vy{y->cpitch, y->coff};
    DG_DISPATCH_1(x->dtype, "dg_upsample2x_relu_fwd",
                  dgvec::upsample_relu_fwd8_kernel<T><<<dgvec::ew8_blocks(total / 32, ctx->sm_count), dgvec::VT, 0, ST>>>(
                      (const T*)x->ptr, vx, (T*)y->ptr, vy, x->n, x->h, x->w, x->c););
    DG_CHECK_LAUNCH("dg_upsample2x_relu_fwd");
    return 0;
  }
  DG_DISPATCH_1(x->dtype, "dg_upsample2x_relu_fwd",
                upsample_relu_fwd_kernel<T><<<ew_blocks(total, ctx->sm_count), 256, 0, ST>>>(
                    (const T*)x->ptr, view_of(x), (T*)y->ptr, view_of(y), x->n, x->h, x->w, x->c););
  DG_CHECK_LAUNCH("dg_upsample2x_relu_fwd");
  return 0;
}

extern "C" int dg_upsample2x_relu_bwd(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* x, const dg_tensor* dx,
                                      void* stream) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(x) && dg_valid(dx), "dg_upsample2x_relu_bwd: null argument");
  DG_REQUIRE(dy->h == 2 * x->h && dy->w == 2 * x->w && dy->c == x->c && dg_same_shape(dx, x) &&
                 dy->dtype == x->dtype && dx->dtype == x->dtype, "dg_upsample2x_relu_bwd: shape mismatch");
  long total = dg_pixels(x) * x->c;
  if (dgvec::vec_ok(dy) && dgvec::vec_ok(x) && dgvec::vec_ok(dx)) {
    const dgvec::VView vd{dy->cpitch, dy->coff}, vx{x->cpitch, x->coff}, vo{dx->cpitch, dx->coff};
    DG_DISPATCH_1(x->dtype, "dg_upsample2x_relu_bwd",
                  dgvec::upsample_relu_bwd8_kernel<T><<<dgvec::ew8_blocks(total / 8, ctx->sm_count), dgvec::VT, 0, ST>>>(
                      (const T*)dy->ptr, vd, (const T*)x->ptr, vx, (T*)dx->ptr, vo, x->n, x->h, x->w, x->c););
    DG_CHECK_LAUNCH("dg_upsample2x_relu_bwd");
    return 0;
  }
  DG_DISPATCH_1(x->dtype, "dg_upsample2x_relu_bwd",
                upsample_relu_bwd_kernel<T><<<ew_blocks(total, ctx->sm_count), 256, 0, ST>>>(
                    (const T*)dy->ptr, view_of(dy), (const T*)x->ptr, view_of(x), (T*)dx->ptr, view_of(dx), x->n, x->h,
                    x->w, x->c););
  DG_CHECK_LAUNCH("dg_upsample2x_relu_bwd");
  return 0;
}

extern "C" int dg_vgg_preprocess_fwd(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* y, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y), "dg_vgg_preprocess_fwd: null argument");
  DG_REQUIRE(x->c == 3 && dg_same_shape(x, y), "dg_vgg_preprocess_fwd: expects 3-channel images");
  long P = dg_pixels(x);
  DG_DISPATCH_2(x->dtype, y->dtype, "dg_vgg_preprocess_fwd",
                vgg_pre_fwd_kernel<TI, TO><<<ew_blocks(P * 3, ctx->sm_count), 256, 0, ST>>>(
                    (const TI*)x->ptr, view_of(x), (TO*)y->ptr, view_of(y), P););
  DG_CHECK_LAUNCH("dg_vgg_preprocess_fwd");
  return 0;
}

extern "C" int dg_vgg_preprocess_bwd(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* dx, void* stream) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(dx), "dg_vgg_preprocess_bwd: null argument");
  DG_REQUIRE(dy->c == 3 && dg_same_shape(dy, dx), "dg_vgg_preprocess_bwd: expects 3-channel images");
  long P = dg_pixels(dy);
  DG_DISPATCH_2(dy->dtype, dx->dtype, "dg_vgg_preprocess_bwd",
                vgg_pre_bwd_kernel<TI, TO><<<ew_blocks(P * 3, ctx->sm_count), 256, 0, ST>>>(
                    (const TI*)dy->ptr, view_of(dy), (TO*)dx->ptr, view_of(dx), P););
  DG_CHECK_LAUNCH("dg_vgg_preprocess_bwd");
  return 0;
}

// dbias[c] (+)= sum over pixels of dy[.., c]  (bias gradient of Conv2DTranspose, pix2pix.py:169-173)
extern "C" int dg_bias_grad(dg_ctx* ctx, const dg_tensor* dy, float* dbias, int accumulate, void* workspace, size_t workspace_bytes,
                            void* stream) {
  DG_REQUIRE(dg_valid(dy) && dbias && workspace, "dg_bias_grad: null argument");
  DG_REQUIRE(dy->c <= RED_THREADS * MAX_CPT, "dg_bias_grad: C too large");
  DG_REQUIRE(workspace_bytes >= dg_bn_workspace_bytes(dy), "dg_bias_grad: workspace too small");
  long P = dg_pixels(dy);
  int C = dy->c;
  int blocks = red_blocks(P, C, ctx->sm_count);
  int R = RED_THREADS / red_lanes(C);
  float* partial = (float*)workspace;
  DG_DISPATCH_1(dy->dtype, "dg_bias_grad",
                bias_grad_kernel<T><<<blocks, RED_THREADS, (size_t)R * C * sizeof(float), ST>>>((const T*)dy->ptr, view_of(dy), P, C, partial););
  sum_partials_kernel<<<(C + 7) / 8, 256, 0, ST>>>(partial, blocks, C, dbias, accumulate);
  DG_CHECK_LAUNCH("dg_bias_grad");
  return 0;
}

// ------------------------------------------------------------------ channel padding for the tensor-core path
// Layers with 3 (RGB) channels on one side run on the tensor cores through a zero-padded 16-channel bf16 copy of
// that side (srgan.py:154 conv 3->64, :182 conv 64->3, :236-246 discriminator conv 3->32).
namespace {
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
pad_channels_kernel(const TI* __restrict__ src, View sv, TO* __restrict__ dst, View dv, long P, int cs, int cd) {
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
    const TI* s = src + p * sv.pitch + sv.off;
    TO* d = dst + p * dv.pitch + dv.off;
    for (int c = 0; c < cd; ++c) st_f<TO>(d + c, c < cs ? ld_f<TI>(s + c) : 0.f);
  }
}

// 16 bf16 output channels per pixel written as two 16-byte stores
template <typename TI>
__global__ void __launch_bounds__(256)
pad_channels16_kernel(const TI* __restrict__ src, View sv, __nv_bfloat16* __restrict__ dst, long P, int cs) {
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
    const TI* s = src + p * sv.pitch + sv.off;
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = c < cs ? ld_f<TI>(s + c) : 0.f;
    uint32_t w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * c], v[2 * c + 1]);
      w[c] = *reinterpret_cast<uint32_t*>(&hh);
    }
    uint4* d = reinterpret_cast<uint4*>(dst + p * 16);
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// dst[t][c][o] (+)= src[t][c][o] for c < cin, o < cout  (src rows are cin_p x cout_p);  dbias likewise
// Two-segment input-channel axis (physically padded U-Net concat, see dg_umma_pack_weights_seg): source channel c lives at
// physical row c (c < seg_log) or seg_phys + (c - seg_log); seg_phys == 0: one segment.
__global__ void unpad_weight_grad_kernel(const float* __restrict__ src, const float* __restrict__ bsrc, float* __restrict__ dst,
                                         float* __restrict__ bdst, int taps, int cin, int cout, int cin_p, int cout_p,
                                         int accumulate, int seg_log, int seg_phys) {
  const long n = (long)taps * cin * cout;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n + (bdst ? cout : 0); i += (long)gridDim.x * blockDim.x) {
    if (i < n) {
      const int o = (int)(i % cout);
      const long r = i / cout;
      const int c = (int)(r % cin), t = (int)(r / cin);
      const int cp = (seg_phys == 0 || c < seg_log) ? c : seg_phys + (c - seg_log);
      const float v = src[((long)t * cin_p + cp) * cout_p + o];
      dst[i] = accumulate ? dst[i] + v : v;
    } else {
      const int o = (int)(i - n);
      bdst[o] = accumulate ? bdst[o] + bsrc[o] : bsrc[o];
    }
  }
}
}  // namespace

extern "C" int dg_pad_channels(dg_ctx* ctx, const dg_tensor* src, const dg_tensor* dst, void* stream) {
  DG_REQUIRE(dg_valid(src) && dg_valid(dst), "dg_pad_channels: null argument");
  DG_REQUIRE(src->n == dst->n && src->h == dst->h && src->w == dst->w && src->c <= dst->c, "dg_pad_channels: shape mismatch");
  const long P = dg_pixels(src);
  if (dst->dtype == DG_BF16 && dst->c == 16 && dst->cpitch == 16 && dst->coff == 0 && ((uintptr_t)dst->ptr % 16) == 0) {
    DG_DISPATCH_1(src->dtype, "dg_pad_channels",
                  pad_channels16_kernel<T><<<ew_blocks(P, ctx->sm_count), 256, 0, ST>>>((const T*)src->ptr, view_of(src),
                                                                                        (__nv_bfloat16*)dst->ptr, P, src->c););
  } else {
    DG_DISPATCH_2(src->dtype, dst->dtype, "dg_pad_channels",
                  pad_channels_kernel<TI, TO><<<ew_blocks(P, ctx->sm_count), 256, 0, ST>>>(
                      (const TI*)src->ptr, view_of(src), (TO*)dst->ptr, view_of(dst), P, src->c, dst->c););
  }
  DG_CHECK_LAUNCH("dg_pad_channels");
  return 0;
}

extern "C" int dg_unpad_weight_grad(dg_ctx* ctx, const float* dw_padded, const float* dbias_padded, float* dw, float* dbias, int kh,
                                    int kw, int cin, int cout, int cin_pad, int cout_pad, int accumulate, void* stream) {
  DG_REQUIRE(dw_padded && dw && cin <= cin_pad && cout <= cout_pad && (!dbias || dbias_padded), "dg_unpad_weight_grad: bad argument");
  const long n = (long)kh * kw * cin * cout + (dbias ? cout : 0);
  unpad_weight_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST>>>(dw_padded, dbias_padded, dw, dbias, kh * kw, cin, cout,
                                                                        cin_pad, cout_pad, accumulate, 0, 0);
  DG_CHECK_LAUNCH("dg_unpad_weight_grad");
  return 0;
}

extern "C" int dg_unpad_weight_grad_seg(dg_ctx* ctx, const float* dw_padded, const float* dbias_padded, float* dw, float* dbias, int kh,
                                        int kw, int cin, int cout, int cin_pad, int cout_pad, int seg_log, int seg_phys, int accumulate,
                                        void* stream) {
  DG_REQUIRE(dw_padded && dw && cin <= cin_pad && cout <= cout_pad && (!dbias || dbias_padded), "dg_unpad_weight_grad_seg: bad argument");
  DG_REQUIRE(seg_phys == 0 || (seg_log >= 0 && seg_log <= seg_phys && seg_log <= cin && seg_phys <= cin_pad &&
                               cin - seg_log <= cin_pad - seg_phys), "dg_unpad_weight_grad_seg: bad channel segments");
  const long n = (long)kh * kw * cin * cout + (dbias ? cout : 0);
  unpad_weight_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST>>>(dw_padded, dbias_padded, dw, dbias, kh * kw, cin, cout,
                                                                        cin_pad, cout_pad, accumulate, seg_log, seg_phys);
  DG_CHECK_LAUNCH("dg_unpad_weight_grad_seg");
  return 0;
}
