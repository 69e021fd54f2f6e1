// 8-channel-vector versions of the HBM-bound kernels in pointwise.cu: every thread moves 16 bytes
// (bf16) or 32 bytes (fp32) per tensor per step, so a warp touches fully used 128-byte lines.
// Used whenever channel count, pitch and offset are multiples of 8 (all 16/32/64/256-channel
// activations of the networks); the scalar kernels in pointwise.cu remain for odd shapes (RGB images,
// the autoencoder's 44/56/76/100-channel layers).
#pragma once
#include "dg_common.cuh"

namespace dgvec {

constexpr int VT = 256;

template <typename T>
struct V8;
struct RawF32 {
  float4 a, b;
};
template <>
struct V8<float> {
  typedef RawF32 raw;
  static __device__ __forceinline__ raw ldraw(const float* p) {
    raw r;
    r.a = *reinterpret_cast<const float4*>(p);
    r.b = *reinterpret_cast<const float4*>(p + 4);
    return r;
  }
  static __device__ __forceinline__ void cvt(const raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct V8<__nv_bfloat16> {
  typedef uint4 raw;
  static __device__ __forceinline__ raw ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void cvt(const raw& u, float (&v)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ void ldc8(const float* __restrict__ p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

struct VView {
  int pitch, off;
};

// ---------------------------------------------------------------- per-channel reduction skeleton (8 channels / thread)
// One block of RT threads per SM.  A thread owns 8 fixed channels (so per-channel constants live in registers) and walks
// pixels with U independent 16-byte loads per tensor in flight: `load(p, c0)` only fetches, `accum(raw, p, c0, acc)`
// only computes.  Every block writes one partial [NV][C]; the LAST block to finish (ticket counter) sums the partials
// in a fixed order in double precision and calls fin(c, sums) per channel: no finalize launch, and the result does
// not depend on block scheduling (deterministic).
constexpr int RT = 512;

// Second half of channel_reduce8: block reduce of the per-thread sums, per-block partial, last-block finalize.
template <int NV, typename Fin>
__device__ __forceinline__ bool channel_reduce8_finish(float (&acc)[NV][8], long P, int C, float* __restrict__ partial,
                                                       unsigned* __restrict__ ticket, Fin fin) {
  extern __shared__ __align__(16) float red8[];
  __shared__ int is_last;
  const int CV = C >> 3, R = RT / CV;
  const int lane = threadIdx.x % CV, row = threadIdx.x / CV;
  const int c0 = lane * 8;
  const int E = NV * C;
  // ---- block reduce: rows sharing a warp first (shuffles), then across warps through shared memory
  int rows_out;
  if (CV <= 32 && (CV & (CV - 1)) == 0) {       // RT % CV == 0: every thread is active, warps hold 32/CV whole rows
    for (int o = CV; o < 32; o <<= 1)
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[v][j] += __shfl_xor_sync(0xffffffffu, acc[v][j], o);
    if ((int)(threadIdx.x & 31) < CV) {
      const int wp = threadIdx.x >> 5;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) red8[wp * E + v * C + c0 + j] = acc[v][j];
    }
    rows_out = RT / 32;
  } else {
    if (row < R) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) red8[row * E + v * C + c0 + j] = acc[v][j];
    }
    rows_out = R;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += RT) {
    float s = 0.f;
    for (int r = 0; r < rows_out; ++r) s += red8[r * E + e];
    partial[(long)blockIdx.x * E + e] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  // ---- last block: thread = (float4 column of the partial, group of blocks); four independent loads in flight
  double* fsum = reinterpret_cast<double*>(red8);   // [E]; group sums [G][E] live behind it
  double* gsum = fsum + E;
  const int nb = (int)gridDim.x;
  const int E4 = E >> 2;
  const int G = E4 < RT ? RT / E4 : 1;
  const float4* part4 = reinterpret_cast<const float4*>(partial);
  for (int idx = threadIdx.x; idx < E4 * G; idx += RT) {
    const int e4 = idx % E4, grp = idx / E4;
    double s[4] = {0, 0, 0, 0};
    int bI = grp;
    for (; bI + 3 * G < nb; bI += 4 * G) {
      const float4 v0 = __ldcg(part4 + (long)bI * E4 + e4), v1 = __ldcg(part4 + (long)(bI + G) * E4 + e4);
      const float4 v2 = __ldcg(part4 + (long)(bI + 2 * G) * E4 + e4), v3 = __ldcg(part4 + (long)(bI + 3 * G) * E4 + e4);
      s[0] += (double)v0.x; s[1] += (double)v0.y; s[2] += (double)v0.z; s[3] += (double)v0.w;
      s[0] += (double)v1.x; s[1] += (double)v1.y; s[2] += (double)v1.z; s[3] += (double)v1.w;
      s[0] += (double)v2.x; s[1] += (double)v2.y; s[2] += (double)v2.z; s[3] += (double)v2.w;
      s[0] += (double)v3.x; s[1] += (double)v3.y; s[2] += (double)v3.z; s[3] += (double)v3.w;
    }
    for (; bI < nb; bI += G) {
      const float4 v = __ldcg(part4 + (long)bI * E4 + e4);
      s[0] += (double)v.x; s[1] += (double)v.y; s[2] += (double)v.z; s[3] += (double)v.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) gsum[(long)grp * E + e4 * 4 + k] = s[k];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += RT) {
    double t = 0;
    for (int g2 = 0; g2 < G; ++g2) t += gsum[(long)g2 * E + e];
    fsum[e] = t;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += RT) fin(c, fsum);
  if (threadIdx.x == 0) *ticket = 0u;
  return true;
}

template <int NV, int U, typename Raw, typename L, typename A, typename Fin>
__device__ __forceinline__ bool channel_reduce8(long P, int C, float* __restrict__ partial, unsigned* __restrict__ ticket, L load,
                                                A accum, Fin fin) {
  const int CV = C >> 3, R = RT / CV;
  const int lane = threadIdx.x % CV, row = threadIdx.x / CV;
  const int c0 = lane * 8;
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[v][j] = 0.f;
  if (row < R) {
    const long stride = (long)gridDim.x * R;
    long p = (long)blockIdx.x * R + row;
    for (; p + (U - 1) * stride < P; p += U * stride) {
      Raw r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) r[u] = load(p + u * stride, c0);
#pragma unroll
      for (int u = 0; u < U; ++u) accum(r[u], p + u * stride, c0, acc);
    }
    for (; p < P; p += stride) {
      Raw r = load(p, c0);
      accum(r, p, c0, acc);
    }
  }
  return channel_reduce8_finish<NV>(acc, P, C, partial, ticket, fin);
}

// Grid-wide "results are ready" barrier for kernels whose blocks are all co-resident (grid <= SM count, one block per
// SM): the block that finished the reduction (`last`) raises `flag`, everybody waits for it, and the last block to
// pass lowers it again for the next kernel.  Lets a reduction and the element-wise pass that consumes its result be
// ONE launch: on the 19 MB generator tensors a kernel boundary (launch + ramp + tail) costs as much as the pass itself.
__device__ __forceinline__ void grid_flag_barrier(bool last, unsigned* flag, unsigned* passed) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (last) {
      __threadfence();
      atomicExch(flag, 1u);
    }
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    } while (v == 0u);
    __threadfence();
    if (atomicAdd(passed, 1u) == gridDim.x - 1) {
      atomicExch(passed, 0u);
      atomicExch(flag, 0u);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void ldcg8(const float* __restrict__ p, float (&v)[8]) {   // values written earlier in this kernel
  float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

static inline int red8_blocks(long P, int C, int sm_count) {
  int R = RT / (C >> 3);
  long want = (P + R - 1) / R, cap = (long)sm_count;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
static inline size_t red8_smem(int C, int NV) {
  const int CV = C >> 3, E = NV * C, E4 = E >> 2, G = E4 < RT ? RT / E4 : 1;
  const int rows = (CV <= 32 && (CV & (CV - 1)) == 0) ? RT / 32 : RT / CV;
  size_t a = (size_t)rows * E * sizeof(float), b = (size_t)(E + (size_t)G * E) * sizeof(double);
  return a > b ? a : b;
}

// ---------------------------------------------------------------- element-wise skeleton with the same thread->channel map
// `body(p, c0)`-style kernels below use ET threads; a thread keeps its 8 channels and walks pixels EU at a time.
constexpr int ET = 256;
constexpr int EU = 4;
static inline unsigned ewc_blocks(long P, int C, int sm_count) {
  int R = ET / (C >> 3);
  long want = (P + (long)R * EU - 1) / ((long)R * EU), cap = (long)sm_count * 4;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

static inline bool vec_ok(const dg_tensor* t) {
  int esz = t->dtype == DG_F32 ? 4 : 2;
  if ((long)t->n * t->h * t->w * t->cpitch >= (1L << 31)) return false;  // kernels index with 32 bits
  return t->c % 8 == 0 && t->cpitch % 8 == 0 && t->coff % 8 == 0 && t->c <= 1024 && ((uintptr_t)t->ptr % 16) == 0 &&
         (t->cpitch * esz) % 16 == 0;
}

template <typename T>
__global__ void __launch_bounds__(RT, 1)
bn_stats8_kernel(const T* __restrict__ x, VView xv, long P, int C, float* __restrict__ partial, unsigned* __restrict__ ticket,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                 float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ scale, float* __restrict__ shift,
                 float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  pdl_trigger();
  pdl_wait();
  typedef typename V8<T>::raw Raw;
  channel_reduce8<2, (sizeof(Raw) <= 16 ? 16 : 8), Raw>(
      P, C, partial, ticket,
      [&](long p, int c0) { return V8<T>::ldraw(x + (p * xv.pitch + xv.off + c0)); },
      [&](const Raw& r, long, int, float (&a)[2][8]) {
        float v[8];
        V8<T>::cvt(r, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[0][j] += v[j]; a[1][j] = fmaf(v[j], v[j], a[1][j]); }
      },
      [&](int c, const double* sums) {
        double mean = sums[c] / (double)P;
        double var = sums[C + c] / (double)P - mean * mean;
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
      });
}

// activation applied to t = bn(x) (after dropout); shared by forward and backward
__device__ __forceinline__ float act_deriv(float tt, int act, float alpha, float al) {
  switch (act) {
    case DG_ACT_RELU: return tt > 0.f ? 1.f : 0.f;
    case DG_ACT_LRELU: return tt >= 0.f ? 1.f : alpha;
    case DG_ACT_PRELU: return tt > 0.f ? 1.f : al;
    case DG_ACT_TANH: { float yy = tanhf(tt); return 1.f - yy * yy; }
    case DG_ACT_SIGMOID: { float yy = 1.f / (1.f + expf(-tt)); return yy * (1.f - yy); }
    default: return 1.f;
  }
}

// AM = compile-time activation mode of the backward kernels: 0 none, 1 relu, 2 leaky relu, 3 prelu (straight-line code,
// no dropout); 4 = anything else decided at run time (tanh, sigmoid, dropout)
template <int AM>
__device__ __forceinline__ float act_deriv_t(float tt, int act, float alpha, float al) {
  if (AM == 0) return 1.f;
  if (AM == 1) return tt > 0.f ? 1.f : 0.f;
  if (AM == 2) return tt >= 0.f ? 1.f : alpha;
  if (AM == 3) return tt > 0.f ? 1.f : al;
  return act_deriv(tt, act, alpha, al);
}
static inline int act_mode(int act, int dropout) {
  if (dropout) return 4;
  switch (act) {
    case DG_ACT_NONE: return 0;
    case DG_ACT_RELU: return 1;
    case DG_ACT_LRELU: return 2;
    case DG_ACT_PRELU: return 3;
    default: return 4;
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(ET)
bn_act_fwd8_kernel(const TI* __restrict__ x, VView xv, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                   float alpha, const float* __restrict__ prelu_alpha, const TO* __restrict__ res, VView rv, int dropout,
                   uint32_t seed0, uint32_t offset, const int64_t* __restrict__ ctr, TO* __restrict__ y, VView yv, long P, int C) {
  pdl_trigger();
  pdl_wait();
  const uint32_t seed = seed0 + (ctr ? (uint32_t)(*ctr) * 0x9E3779B9u : 0u);  // a new mask every optimiser step, graph-replay safe
  const int CV = C >> 3, R = ET / CV;
  const int row = threadIdx.x / CV, c0 = (threadIdx.x % CV) * 8;
  if (row >= R) return;
  float sc[8], sh[8], al[8];
  ldc8(scale + c0, sc); ldc8(shift + c0, sh);
  if (act == DG_ACT_PRELU) ldc8(prelu_alpha + c0, al);
  const long stride = (long)gridDim.x * R;
  typedef typename V8<TI>::raw RawI;
  typedef typename V8<TO>::raw RawO;
  auto one = [&](const RawI& rx, const RawO& rr, long p) {
    float v[8];
    V8<TI>::cvt(rx, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    if (dropout) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = dropout_keep(seed, offset + (uint32_t)(p * C + c0 + j)) ? 2.f * v[j] : 0.f;
    }
    if (act == DG_ACT_PRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : al[j] * v[j];
    } else if (act == DG_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == DG_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] >= 0.f ? v[j] : alpha * v[j];
    } else if (act != DG_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], act, alpha);
    }
    if (res) {
      float r[8];
      V8<TO>::cvt(rr, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    V8<TO>::st(y + (p * yv.pitch + yv.off + c0), v);
  };
  long p = (long)blockIdx.x * R + row;
  for (; p + (EU - 1) * stride < P; p += EU * stride) {
    RawI rx[EU];
    RawO rr[EU];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      rx[u] = V8<TI>::ldraw(x + ((p + u * stride) * xv.pitch + xv.off + c0));
      if (res) rr[u] = V8<TO>::ldraw(res + ((p + u * stride) * rv.pitch + rv.off + c0));
    }
#pragma unroll
    for (int u = 0; u < EU; ++u) one(rx[u], rr[u], p + u * stride);
  }
  for (; p < P; p += stride) {
    RawI rx = V8<TI>::ldraw(x + (p * xv.pitch + xv.off + c0));
    RawO rr;
    if (res) rr = V8<TO>::ldraw(res + (p * rv.pitch + rv.off + c0));
    one(rx, rr, p);
  }
}

// Training-mode BatchNorm forward in ONE launch: batch statistics (as bn_stats8_kernel, incl. the moving-statistics
// update), grid flag barrier, then normalise + dropout + activation + skip-add (as bn_act_fwd8_kernel) by the same blocks.
template <typename TI, typename TO>
__global__ void __launch_bounds__(RT, 1)
bn_fwd_fused8_kernel(const TI* __restrict__ x, VView xv, long P, int C, float* __restrict__ partial, unsigned* __restrict__ ticket,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                     float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ scale, float* __restrict__ shift,
                     float* __restrict__ save_mean, float* __restrict__ save_invstd, int act, float alpha,
                     const float* __restrict__ prelu_alpha, const TO* __restrict__ res, VView rv, int dropout, uint32_t seed0,
                     uint32_t offset, const int64_t* __restrict__ ctr, TO* __restrict__ y, VView yv) {
  pdl_trigger();
  pdl_wait();
  typedef typename V8<TI>::raw Raw;
  const bool last = channel_reduce8<2, (sizeof(Raw) <= 16 ? 16 : 8), Raw>(
      P, C, partial, ticket,
      [&](long p, int c0) { return V8<TI>::ldraw(x + (p * xv.pitch + xv.off + c0)); },
      [&](const Raw& r, long, int, float (&a)[2][8]) {
        float v[8];
        V8<TI>::cvt(r, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[0][j] += v[j]; a[1][j] = fmaf(v[j], v[j], a[1][j]); }
      },
      [&](int c, const double* sums) {
        double mean = sums[c] / (double)P;
        double var = sums[C + c] / (double)P - mean * mean;
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
      });
  grid_flag_barrier(last, ticket + 1, ticket + 2);
  const uint32_t seed = seed0 + (ctr ? (uint32_t)(*ctr) * 0x9E3779B9u : 0u);
  const int CV = C >> 3, R = RT / CV;
  const int row = threadIdx.x / CV, c0 = (threadIdx.x % CV) * 8;
  if (row >= R) return;
  float sc[8], sh[8], al[8];
  ldcg8(scale + c0, sc); ldcg8(shift + c0, sh);
  if (act == DG_ACT_PRELU) ldc8(prelu_alpha + c0, al);
  const long stride = (long)gridDim.x * R;
  typedef typename V8<TO>::raw RawO;
  auto one = [&](const Raw& rx, const RawO& rr, long p) {
    float v[8];
    V8<TI>::cvt(rx, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    if (dropout) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = dropout_keep(seed, offset + (uint32_t)(p * C + c0 + j)) ? 2.f * v[j] : 0.f;
    }
    if (act == DG_ACT_PRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : al[j] * v[j];
    } else if (act == DG_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == DG_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] >= 0.f ? v[j] : alpha * v[j];
    } else if (act != DG_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], act, alpha);
    }
    if (res) {
      float r[8];
      V8<TO>::cvt(rr, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    V8<TO>::st(y + (p * yv.pitch + yv.off + c0), v);
  };
  long p = (long)blockIdx.x * R + row;
  for (; p + (EU - 1) * stride < P; p += EU * stride) {
    Raw rx[EU];
    RawO rr[EU];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      rx[u] = V8<TI>::ldraw(x + ((p + u * stride) * xv.pitch + xv.off + c0));
      if (res) rr[u] = V8<TO>::ldraw(res + ((p + u * stride) * rv.pitch + rv.off + c0));
    }
#pragma unroll
    for (int u = 0; u < EU; ++u) one(rx[u], rr[u], p + u * stride);
  }
  for (; p < P; p += stride) {
    Raw rx = V8<TI>::ldraw(x + (p * xv.pitch + xv.off + c0));
    RawO rr;
    if (res) rr = V8<TO>::ldraw(res + (p * rv.pitch + rv.off + c0));
    one(rx, rr, p);
  }
}

// Training-mode BatchNorm forward when the convolution that produced x already reduced the batch statistics to per-CTA rows
// ([nrows][2][C]: sum, sum of squares; dg_umma_conv2d_fwd with bn_partials): every block sums the rows in the same fixed order in
// double precision and derives scale / shift itself (no finalize launch, no grid barrier); block 0 publishes scale / shift / mean /
// invstd for the backward pass and updates the moving statistics; then normalise + activation + skip-add as bn_act_fwd8_kernel.
template <typename TI, typename TO>
__global__ void __launch_bounds__(RT, 1)
bn_fwd_part8_kernel(const TI* __restrict__ x, VView xv, long P, int C, const float* __restrict__ partial, int nrows,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                    float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ scale, float* __restrict__ shift,
                    float* __restrict__ save_mean, float* __restrict__ save_invstd, int act, float alpha,
                    const float* __restrict__ prelu_alpha, const TO* __restrict__ res, VView rv, TO* __restrict__ y, VView yv) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float red8[];
  const int E = 2 * C, E4 = E >> 2, G = E4 < RT ? RT / E4 : 1;
  double* gsum = reinterpret_cast<double*>(red8);                    // [G][E]
  float* coef_s = reinterpret_cast<float*>(gsum + (size_t)G * E);   // [2][C]: scale, shift
  {
    const float4* part4 = reinterpret_cast<const float4*>(partial);
    for (int idx = threadIdx.x; idx < E4 * G; idx += RT) {
      const int e4 = idx % E4, grp = idx / E4;
      double s[4] = {0, 0, 0, 0};
      int bI = grp;
      for (; bI + 3 * G < nrows; bI += 4 * G) {     // four independent loads in flight
        const float4 v0 = __ldg(part4 + (long)bI * E4 + e4), v1 = __ldg(part4 + (long)(bI + G) * E4 + e4);
        const float4 v2 = __ldg(part4 + (long)(bI + 2 * G) * E4 + e4), v3 = __ldg(part4 + (long)(bI + 3 * G) * E4 + e4);
        s[0] += (double)v0.x; s[1] += (double)v0.y; s[2] += (double)v0.z; s[3] += (double)v0.w;
        s[0] += (double)v1.x; s[1] += (double)v1.y; s[2] += (double)v1.z; s[3] += (double)v1.w;
        s[0] += (double)v2.x; s[1] += (double)v2.y; s[2] += (double)v2.z; s[3] += (double)v2.w;
        s[0] += (double)v3.x; s[1] += (double)v3.y; s[2] += (double)v3.z; s[3] += (double)v3.w;
      }
      for (; bI < nrows; bI += G) {
        const float4 v = __ldg(part4 + (long)bI * E4 + e4);
        s[0] += (double)v.x; s[1] += (double)v.y; s[2] += (double)v.z; s[3] += (double)v.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) gsum[(long)grp * E + e4 * 4 + k] = s[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += RT) {
      double s0 = 0, s1 = 0;
      for (int g2 = 0; g2 < G; ++g2) { s0 += gsum[(long)g2 * E + c]; s1 += gsum[(long)g2 * E + C + c]; }
      const double mean = s0 / (double)P;
      double var = s1 / (double)P - mean * mean;
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)eps));
      const float g = gamma[c], b = beta[c];
      const float scv = g * invstd, shv = b - (float)mean * g * invstd;
      coef_s[c] = scv;
      coef_s[C + c] = shv;
      if (blockIdx.x == 0) {
        scale[c] = scv;
        shift[c] = shv;
        save_mean[c] = (float)mean;
        save_invstd[c] = invstd;
        if (moving_mean) {
          moving_mean[c] = moving_mean[c] * momentum + (float)mean * (1.f - momentum);
          moving_var[c] = moving_var[c] * momentum + (float)(var * ((double)P / (double)(P > 1 ? P - 1 : 1))) * (1.f - momentum);   // Bessel-corrected
        }
      }
    }
    __syncthreads();
  }
  const int CV = C >> 3, R = RT / CV;
  const int row = threadIdx.x / CV, c0 = (threadIdx.x % CV) * 8;
  if (row >= R) return;
  float sc[8], sh[8], al[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = coef_s[c0 + j]; sh[j] = coef_s[C + c0 + j]; }
  if (act == DG_ACT_PRELU) ldc8(prelu_alpha + c0, al);
  const long stride = (long)gridDim.x * R;
  typedef typename V8<TI>::raw Raw;
  typedef typename V8<TO>::raw RawO;
  auto apply = [&](const Raw& rx, const RawO& rr, long p) {
    float v[8];
    V8<TI>::cvt(rx, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    if (act == DG_ACT_PRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : al[j] * v[j];
    } else if (act == DG_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == DG_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] >= 0.f ? v[j] : alpha * v[j];
    } else if (act != DG_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], act, alpha);
    }
    if (res) {
      float r[8];
      V8<TO>::cvt(rr, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    V8<TO>::st(y + (p * yv.pitch + yv.off + c0), v);
  };
  long p = (long)blockIdx.x * R + row;
  for (; p + (EU - 1) * stride < P; p += EU * stride) {
    Raw rx[EU];
    RawO rr[EU];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      rx[u] = V8<TI>::ldraw(x + ((p + u * stride) * xv.pitch + xv.off + c0));
      if (res) rr[u] = V8<TO>::ldraw(res + ((p + u * stride) * rv.pitch + rv.off + c0));
    }
#pragma unroll
    for (int u = 0; u < EU; ++u) apply(rx[u], rr[u], p + u * stride);
  }
  for (; p < P; p += stride) {
    Raw rx = V8<TI>::ldraw(x + (p * xv.pitch + xv.off + c0));
    RawO rr;
    if (res) rr = V8<TO>::ldraw(res + (p * rv.pitch + rv.off + c0));
    apply(rx, rr, p);
  }
}

template <typename TG, typename TX>
struct RawPair {
  typename V8<TG>::raw g;
  typename V8<TX>::raw x;
};

// Sums for the BatchNorm backward pass: s0 = sum g, s1 = invstd * sum g*(x-mean), s2 = sum dy*min(t,0) (PReLU slope), with
// g = dL/d(bn output) = dy * act'(t) * dropout, t = bn(x).  Finalize also leaves coef = {s0/P, s1/P} for the dx kernel.
template <typename TG, typename TX, int AM>
__global__ void __launch_bounds__(RT, 1)
bn_bwd_reduce8_kernel(const TG* __restrict__ dy, VView dv, const TX* __restrict__ x, VView xv, const float* __restrict__ scale,
                      const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd, int act,
                      float alpha, const float* __restrict__ prelu_alpha, int dropout, uint32_t seed0, uint32_t offset, const int64_t* __restrict__ ctr, long P, int C,
                      float* __restrict__ partial, unsigned* __restrict__ ticket, float* __restrict__ dgamma, float* __restrict__ dbeta,
                      float* __restrict__ dalpha, int accumulate, float* __restrict__ coef) {
  pdl_trigger();
  pdl_wait();
  const uint32_t seed = seed0 + (ctr ? (uint32_t)(*ctr) * 0x9E3779B9u : 0u);  // a new mask every optimiser step, graph-replay safe
  typedef RawPair<TG, TX> Raw;
  const int c0k = (threadIdx.x % (C >> 3)) * 8;
  float sc[8], sh[8], mu[8], al[8];
  ldc8(scale + c0k, sc); ldc8(shift + c0k, sh); ldc8(mean + c0k, mu);
  if (act == DG_ACT_PRELU) ldc8(prelu_alpha + c0k, al);
  channel_reduce8<3, (sizeof(Raw) <= 32 ? 4 : 2), Raw>(
      P, C, partial, ticket,
      [&](long p, int c0) {
        Raw r;
        r.g = V8<TG>::ldraw(dy + (p * dv.pitch + dv.off + c0));
        r.x = V8<TX>::ldraw(x + (p * xv.pitch + xv.off + c0));
        return r;
      },
      [&](const Raw& r, long p, int c0, float (&a)[3][8]) {
        float gy[8], xin[8];
        V8<TG>::cvt(r.g, gy);
        V8<TX>::cvt(r.x, xin);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float tt = fmaf(xin[j], sc[j], sh[j]);
          float g = gy[j];
          if (AM == 4 && dropout) {
            const bool keep = dropout_keep(seed, offset + (uint32_t)(p * C + c0 + j));
            tt = keep ? 2.f * tt : 0.f;
            g = keep ? 2.f * g : 0.f;
          }
          // rounded product: the reduce and dx kernels must agree bit for bit
          g = __fmul_rn(g, act_deriv_t<AM>(tt, act, alpha, (AM == 3 || (AM == 4 && act == DG_ACT_PRELU)) ? al[j] : 0.f));
          a[0][j] += g;
          a[1][j] = fmaf(g, xin[j] - mu[j], a[1][j]);
          if (AM == 3 || (AM == 4 && act == DG_ACT_PRELU)) a[2][j] = fmaf(gy[j], fminf(tt, 0.f), a[2][j]);
        }
      },
      [&](int c, const double* sums) {
        const double s0 = sums[c], s1 = sums[C + c] * (double)invstd[c], s2 = sums[2 * C + c];
        if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s0;
        if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s1;
        if (dalpha) dalpha[c] = (accumulate ? dalpha[c] : 0.f) + (float)s2;
        coef[c] = (float)(s0 / (double)P);
        coef[C + c] = (float)(s1 / (double)P);
      });
}

// dx = gamma*invstd*(g - mean_g - xhat*mean_gxhat) = A*(g - mean_g) + B*(x - mean) with per-channel constants in registers
template <typename TG, typename TX, typename TO, int AM>
__global__ void __launch_bounds__(ET)
bn_bwd_dx8_kernel(const TG* __restrict__ dy, VView dv, const TX* __restrict__ x, VView xv, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ gamma, const float* __restrict__ mean,
                  const float* __restrict__ invstd, int act, float alpha, const float* __restrict__ prelu_alpha, int dropout,
                  uint32_t seed0, uint32_t offset, const int64_t* __restrict__ ctr, const float* __restrict__ coef, TO* __restrict__ dx, VView ov, long P, int C) {
  pdl_trigger();
  pdl_wait();
  const uint32_t seed = seed0 + (ctr ? (uint32_t)(*ctr) * 0x9E3779B9u : 0u);  // a new mask every optimiser step, graph-replay safe
  const int CV = C >> 3, R = ET / CV;
  const int row = threadIdx.x / CV, c0 = (threadIdx.x % CV) * 8;
  if (row >= R) return;
  float sc[8], sh[8], al[8], A[8], B[8], mu[8], k0[8];
  ldc8(scale + c0, sc); ldc8(shift + c0, sh); ldc8(mean + c0, mu); ldc8(coef + c0, k0);
  if (act == DG_ACT_PRELU) ldc8(prelu_alpha + c0, al);
  {
    float ga[8], is[8], k1[8];
    ldc8(gamma + c0, ga); ldc8(invstd + c0, is); ldc8(coef + C + c0, k1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      A[j] = ga[j] * is[j];
      B[j] = -A[j] * is[j] * k1[j];
    }
  }
  typedef RawPair<TG, TX> Raw;
  const long stride = (long)gridDim.x * R;
  auto one = [&](const Raw& r, long p) {
    float gy[8], xin[8], o[8];
    V8<TG>::cvt(r.g, gy);
    V8<TX>::cvt(r.x, xin);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float tt = fmaf(xin[j], sc[j], sh[j]);
      float g = gy[j];
      if (AM == 4 && dropout) {
        const bool keep = dropout_keep(seed, offset + (uint32_t)(p * C + c0 + j));
        tt = keep ? 2.f * tt : 0.f;
        g = keep ? 2.f * g : 0.f;
      }
      g = __fmul_rn(g, act_deriv_t<AM>(tt, act, alpha, (AM == 3 || (AM == 4 && act == DG_ACT_PRELU)) ? al[j] : 0.f));
      o[j] = fmaf(A[j], g - k0[j], B[j] * (xin[j] - mu[j]));   // differences first: exact zero when the batch is one pixel
    }
    V8<TO>::st(dx + (p * ov.pitch + ov.off + c0), o);
  };
  long p = (long)blockIdx.x * R + row;
  for (; p + (EU - 1) * stride < P; p += EU * stride) {
    Raw r[EU];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      r[u].g = V8<TG>::ldraw(dy + ((p + u * stride) * dv.pitch + dv.off + c0));
      r[u].x = V8<TX>::ldraw(x + ((p + u * stride) * xv.pitch + xv.off + c0));
    }
#pragma unroll
    for (int u = 0; u < EU; ++u) one(r[u], p + u * stride);
  }
  for (; p < P; p += stride) {
    Raw r;
    r.g = V8<TG>::ldraw(dy + (p * dv.pitch + dv.off + c0));
    r.x = V8<TX>::ldraw(x + (p * xv.pitch + xv.off + c0));
    one(r, p);
  }
}

// BatchNorm backward in ONE launch: the per-channel sums (as bn_bwd_reduce8_kernel), a grid barrier, then dx (as
// bn_bwd_dx8_kernel) by the same 148 blocks while dy and x are still in L2.
template <typename TG, typename TX, typename TO, int AM>
__global__ void __launch_bounds__(RT, 1)
bn_bwd_fused8_kernel(const TG* __restrict__ dy, VView dv, const TX* __restrict__ x, VView xv, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ invstd, int act, float alpha, const float* __restrict__ prelu_alpha, int dropout,
                     uint32_t seed0, uint32_t offset, const int64_t* __restrict__ ctr, long P, int C, float* __restrict__ partial,
                     unsigned* __restrict__ ticket, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dalpha,
                     int accumulate, float* __restrict__ coef, TO* __restrict__ dx, VView ov) {
  pdl_trigger();
  pdl_wait();
  const uint32_t seed = seed0 + (ctr ? (uint32_t)(*ctr) * 0x9E3779B9u : 0u);
  typedef RawPair<TG, TX> Raw;
  const int CV = C >> 3, R = RT / CV;
  const int row = threadIdx.x / CV, c0 = (threadIdx.x % CV) * 8;
  float sc[8], sh[8], mu[8], al[8];
  ldc8(scale + c0, sc); ldc8(shift + c0, sh); ldc8(mean + c0, mu);
  if (act == DG_ACT_PRELU) ldc8(prelu_alpha + c0, al);
  auto gterm = [&](float gyj, float xj, int j, long p, float& tt) {
    tt = fmaf(xj, sc[j], sh[j]);
    float g = gyj;
    if (AM == 4 && dropout) {
      const bool keep = dropout_keep(seed, offset + (uint32_t)(p * C + c0 + j));
      tt = keep ? 2.f * tt : 0.f;
      g = keep ? 2.f * g : 0.f;
    }
    return __fmul_rn(g, act_deriv_t<AM>(tt, act, alpha, (AM == 3 || (AM == 4 && act == DG_ACT_PRELU)) ? al[j] : 0.f));
  };
  const bool last = channel_reduce8<3, (sizeof(Raw) <= 32 ? 4 : 2), Raw>(
      P, C, partial, ticket,
      [&](long p, int cc) {
        Raw r;
        r.g = V8<TG>::ldraw(dy + (p * dv.pitch + dv.off + cc));
        r.x = V8<TX>::ldraw(x + (p * xv.pitch + xv.off + cc));
        return r;
      },
      [&](const Raw& r, long p, int, float (&a)[3][8]) {
        float gy[8], xin[8];
        V8<TG>::cvt(r.g, gy);
        V8<TX>::cvt(r.x, xin);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float tt;
          const float g = gterm(gy[j], xin[j], j, p, tt);
          a[0][j] += g;
          a[1][j] = fmaf(g, xin[j] - mu[j], a[1][j]);
          if (AM == 3 || (AM == 4 && act == DG_ACT_PRELU)) a[2][j] = fmaf(gy[j], fminf(tt, 0.f), a[2][j]);
        }
      },
      [&](int c, const double* sums) {
        const double s0 = sums[c], s1 = sums[C + c] * (double)invstd[c], s2 = sums[2 * C + c];
        if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s0;
        if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s1;
        if (dalpha) dalpha[c] = (accumulate ? dalpha[c] : 0.f) + (float)s2;
        coef[c] = (float)(s0 / (double)P);
        coef[C + c] = (float)(s1 / (double)P);
      });
  grid_flag_barrier(last, ticket + 1, ticket + 2);
  if (row >= R) return;
  float A[8], B[8], k0[8];
  {
    float ga[8], is[8], k1[8];
    ldc8(gamma + c0, ga); ldc8(invstd + c0, is);
    ldcg8(coef + c0, k0); ldcg8(coef + C + c0, k1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      A[j] = ga[j] * is[j];
      B[j] = -A[j] * is[j] * k1[j];
    }
  }
  const long stride = (long)gridDim.x * R;
  auto one = [&](const Raw& r, long p) {
    float gy[8], xin[8], o[8];
    V8<TG>::cvt(r.g, gy);
    V8<TX>::cvt(r.x, xin);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float tt;
      const float g = gterm(gy[j], xin[j], j, p, tt);
      o[j] = fmaf(A[j], g - k0[j], B[j] * (xin[j] - mu[j]));
    }
    V8<TO>::st(dx + (p * ov.pitch + ov.off + c0), o);
  };
  long p = (long)blockIdx.x * R + row;
  for (; p + (EU - 1) * stride < P; p += EU * stride) {
    Raw r[EU];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      r[u].g = V8<TG>::ldraw(dy + ((p + u * stride) * dv.pitch + dv.off + c0));
      r[u].x = V8<TX>::ldraw(x + ((p + u * stride) * xv.pitch + xv.off + c0));
    }
#pragma unroll
    for (int u = 0; u < EU; ++u) one(r[u], p + u * stride);
  }
  for (; p < P; p += stride) {
    Raw r;
    r.g = V8<TG>::ldraw(dy + (p * dv.pitch + dv.off + c0));
    r.x = V8<TX>::ldraw(x + (p * xv.pitch + xv.off + c0));
    one(r, p);
  }
}

// BatchNorm backward, dx half, when the convolution that produced dy already reduced the two per-channel sums to per-CTA rows
// (dg_umma_conv2d_dgrad_fused: row r = [sum g' | sum g' x] over that CTA's pixels).  Every block first sums the rows in the
// same fixed order in double precision (nrows x 2C floats out of L2, ~75 KB for the generator trunk) -- no finalize launch, no grid
// barrier -- block 0 writes dgamma / dbeta, and then the blocks make ONE pass over dy and x.
template <typename TG, typename TX, typename TO, int AM>
__global__ void __launch_bounds__(RT, 1)
bn_bwd_dx_part8_kernel(const TG* __restrict__ dy, VView dv, const TX* __restrict__ x, VView xv, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ gamma, const float* __restrict__ mean,
                       const float* __restrict__ invstd, float alpha, const float* __restrict__ partial, int nrows, long P, int C,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate, TO* __restrict__ dx, VView ov) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float red8[];
  const int E = 2 * C, E4 = E >> 2, G = E4 < RT ? RT / E4 : 1;
  double* gsum = reinterpret_cast<double*>(red8);            // [G][E]
  float* coef_s = reinterpret_cast<float*>(gsum + (size_t)G * E);   // [2][C]: mean of g', invstd * mean of g'(x - mean)
  {
    const float4* part4 = reinterpret_cast<const float4*>(partial);
    for (int idx = threadIdx.x; idx < E4 * G; idx += RT) {
      const int e4 = idx % E4, grp = idx / E4;
      double s[4] = {0, 0, 0, 0};
      int bI = grp;
      for (; bI + 3 * G < nrows; bI += 4 * G) {     // four independent loads in flight
        const float4 v0 = __ldg(part4 + (long)bI * E4 + e4), v1 = __ldg(part4 + (long)(bI + G) * E4 + e4);
        const float4 v2 = __ldg(part4 + (long)(bI + 2 * G) * E4 + e4), v3 = __ldg(part4 + (long)(bI + 3 * G) * E4 + e4);
        s[0] += (double)v0.x; s[1] += (double)v0.y; s[2] += (double)v0.z; s[3] += (double)v0.w;
        s[0] += (double)v1.x; s[1] += (double)v1.y; s[2] += (double)v1.z; s[3] += (double)v1.w;
        s[0] += (double)v2.x; s[1] += (double)v2.y; s[2] += (double)v2.z; s[3] += (double)v2.w;
        s[0] += (double)v3.x; s[1] += (double)v3.y; s[2] += (double)v3.z; s[3] += (double)v3.w;
      }
      for (; bI < nrows; bI += G) {
        const float4 v = __ldg(part4 + (long)bI * E4 + e4);
        s[0] += (double)v.x; s[1] += (double)v.y; s[2] += (double)v.z; s[3] += (double)v.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) gsum[(long)grp * E + e4 * 4 + k] = s[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += RT) {
      double s0 = 0, s1 = 0;
      for (int g2 = 0; g2 < G; ++g2) { s0 += gsum[(long)g2 * E + c]; s1 += gsum[(long)g2 * E + C + c]; }
      s1 = (s1 - (double)mean[c] * s0) * (double)invstd[c];      // rows hold sum g' x: sum g'(x - mean) = sum g' x - mean sum g'
      if (blockIdx.x == 0) {
        if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s0;
        if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s1;
      }
      coef_s[c] = (float)(s0 / (double)P);
      coef_s[C + c] = (float)(s1 / (double)P);
    }
    __syncthreads();
  }
  typedef RawPair<TG, TX> Raw;
  const int CV = C >> 3, R = RT / CV;
  const int row = threadIdx.x / CV, c0 = (threadIdx.x % CV) * 8;
  if (row >= R) return;
  float sc[8], sh[8], mu[8], A[8], B[8], k0[8];
  ldc8(scale + c0, sc); ldc8(shift + c0, sh); ldc8(mean + c0, mu);
  {
    float ga[8], is[8];
    ldc8(gamma + c0, ga); ldc8(invstd + c0, is);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      k0[j] = coef_s[c0 + j];
      A[j] = ga[j] * is[j];
      B[j] = -A[j] * is[j] * coef_s[C + c0 + j];
    }
  }
  const long stride = (long)gridDim.x * R;
  auto one = [&](const Raw& r, long p) {
    float gy[8], xin[8], o[8];
    V8<TG>::cvt(r.g, gy);
    V8<TX>::cvt(r.x, xin);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float tt = fmaf(xin[j], sc[j], sh[j]);
      const float g = __fmul_rn(gy[j], act_deriv_t<AM>(tt, 0, alpha, 0.f));
      o[j] = fmaf(A[j], g - k0[j], B[j] * (xin[j] - mu[j]));
    }
    V8<TO>::st(dx + (p * ov.pitch + ov.off + c0), o);
  };
  long p = (long)blockIdx.x * R + row;
  for (; p + (EU - 1) * stride < P; p += EU * stride) {
    Raw r[EU];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      r[u].g = V8<TG>::ldraw(dy + ((p + u * stride) * dv.pitch + dv.off + c0));
      r[u].x = V8<TX>::ldraw(x + ((p + u * stride) * xv.pitch + xv.off + c0));
    }
#pragma unroll
    for (int u = 0; u < EU; ++u) one(r[u], p + u * stride);
  }
  for (; p < P; p += stride) {
    Raw r;
    r.g = V8<TG>::ldraw(dy + (p * dv.pitch + dv.off + c0));
    r.x = V8<TX>::ldraw(x + (p * xv.pitch + xv.off + c0));
    one(r, p);
  }
}

// bn_bwd_fused8_kernel for bf16 tensors whose per-block slice fits ON CHIP: in the statistics pass a thread keeps the x
// values it loaded in registers (K x 16 bytes) and parks the dy values in shared memory (K x 512 x 16 bytes), so the dx pass
// after the grid barrier touches no global memory except its store.  The 19 MB tensors of the generator trunk are 128 KB of
// x and 128 KB of dy per SM: the register file and the shared memory of a B200 SM hold exactly one of each.
template <int AM, int K>
__global__ void __launch_bounds__(RT, 1)
bn_bwd_cached8_kernel(const __nv_bfloat16* __restrict__ dy, VView dv, const __nv_bfloat16* __restrict__ x, VView xv,
                      const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ gamma,
                      const float* __restrict__ mean, const float* __restrict__ invstd, int act, float alpha,
                      const float* __restrict__ prelu_alpha, long P, int C, float* __restrict__ partial, unsigned* __restrict__ ticket,
                      float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dalpha, int accumulate,
                      float* __restrict__ coef, __nv_bfloat16* __restrict__ dx, VView ov, unsigned cache_off) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float red8[];
  uint4* dyc = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(red8) + cache_off);
  typedef V8<__nv_bfloat16> V;
  const int CV = C >> 3, R = RT / CV;
  const int row = threadIdx.x / CV, c0 = (threadIdx.x % CV) * 8;
  float sc[8], sh[8], mu[8], al[8];
  ldc8(scale + c0, sc); ldc8(shift + c0, sh); ldc8(mean + c0, mu);
  if (AM == 3) ldc8(prelu_alpha + c0, al);
  auto gterm = [&](float gyj, float xj, int j, float& tt) {
    tt = fmaf(xj, sc[j], sh[j]);
    return __fmul_rn(gyj, act_deriv_t<AM>(tt, act, alpha, AM == 3 ? al[j] : 0.f));
  };
  const long stride = (long)gridDim.x * R;
  const long p0 = (long)blockIdx.x * R + row;
  uint4 xc[K];
  float acc[3][8];
#pragma unroll
  for (int v = 0; v < 3; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[v][j] = 0.f;
  if (row < R) {
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 4) {
      uint4 g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long p = p0 + (long)(k0 + u) * stride;
        if (p < P) {
          g[u] = V::ldraw(dy + (p * dv.pitch + dv.off + c0));
          xc[k0 + u] = V::ldraw(x + (p * xv.pitch + xv.off + c0));
        } else {
          g[u] = make_uint4(0u, 0u, 0u, 0u);
          xc[k0 + u] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dyc[(k0 + u) * RT + threadIdx.x] = g[u];
        if (p0 + (long)(k0 + u) * stride < P) {
          float gy[8], xin[8];
          V::cvt(g[u], gy);
          V::cvt(xc[k0 + u], xin);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float tt;
            const float gg = gterm(gy[j], xin[j], j, tt);
            acc[0][j] += gg;
            acc[1][j] = fmaf(gg, xin[j] - mu[j], acc[1][j]);
            if (AM == 3) acc[2][j] = fmaf(gy[j], fminf(tt, 0.f), acc[2][j]);
          }
        }
      }
    }
  }
  const bool last = channel_reduce8_finish<3>(acc, P, C, partial, ticket, [&](int c, const double* sums) {
    const double s0 = sums[c], s1 = sums[C + c] * (double)invstd[c], s2 = sums[2 * C + c];
    if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s0;
    if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s1;
    if (dalpha) dalpha[c] = (accumulate ? dalpha[c] : 0.f) + (float)s2;
    coef[c] = (float)(s0 / (double)P);
    coef[C + c] = (float)(s1 / (double)P);
  });
  grid_flag_barrier(last, ticket + 1, ticket + 2);
  if (row >= R) return;
  float A[8], B[8], k0v[8];
  {
    float ga[8], is[8], k1[8];
    ldc8(gamma + c0, ga); ldc8(invstd + c0, is);
    ldcg8(coef + c0, k0v); ldcg8(coef + C + c0, k1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      A[j] = ga[j] * is[j];
      B[j] = -A[j] * is[j] * k1[j];
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const long p = p0 + (long)k * stride;
    if (p < P) {
      float gy[8], xin[8], o[8];
      V::cvt(dyc[k * RT + threadIdx.x], gy);
      V::cvt(xc[k], xin);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float tt;
        const float gg = gterm(gy[j], xin[j], j, tt);
        o[j] = fmaf(A[j], gg - k0v[j], B[j] * (xin[j] - mu[j]));
      }
      V::st(dx + (p * ov.pitch + ov.off + c0), o);
    }
  }
}

template <typename TG, typename TY, typename TO>
__global__ void __launch_bounds__(VT)
act_bwd8_kernel(const TG* __restrict__ dy, VView dv, const TY* __restrict__ y, VView yv, int act, float alpha, TO* __restrict__ dpre,
                VView ov, long P, int C) {
  const uint32_t CV = (uint32_t)C >> 3;
  const uint32_t total = (uint32_t)P * CV;   // vec_ok() guarantees < 2^31 elements
  for (uint32_t i = blockIdx.x * VT + threadIdx.x; i < total; i += gridDim.x * VT) {
    const uint32_t p = i / CV;
    const int c0 = (int)(i - p * CV) * 8;
    float g[8], yo[8];
    V8<TG>::ld(dy + (p * dv.pitch + dv.off + c0), g);
    V8<TY>::ld(y + (p * yv.pitch + yv.off + c0), yo);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d;
      switch (act) {
        case DG_ACT_RELU: d = yo[j] > 0.f ? 1.f : 0.f; break;
        case DG_ACT_LRELU: d = yo[j] >= 0.f ? 1.f : alpha; break;
        case DG_ACT_TANH: d = 1.f - yo[j] * yo[j]; break;
        case DG_ACT_SIGMOID: d = yo[j] * (1.f - yo[j]); break;
        default: d = 1.f;
      }
      g[j] *= d;
    }
    V8<TO>::st(dpre + (p * ov.pitch + ov.off + c0), g);
  }
}

// depth_to_space(2)+PReLU: one thread = 8 consecutive input channels (all in one sub-pixel block since Co % 8 == 0)
template <typename T, bool BWD>
__global__ void __launch_bounds__(VT)
d2s_prelu8_kernel(const T* __restrict__ a, VView av, const T* __restrict__ u, VView uv, const float* __restrict__ alpha,
                  T* __restrict__ o, VView ov, int N, int H, int W, int Co) {
  // FWD: a = u (input), o = y.   BWD: a = dy (output-sized), u = saved input, o = du.
  const uint32_t CV = (uint32_t)(4 * Co) >> 3;
  const uint32_t total = (uint32_t)N * H * W * CV;
  for (uint32_t i = blockIdx.x * VT + threadIdx.x; i < total; i += gridDim.x * VT) {
    const uint32_t p = i / CV;
    const int ch0 = (int)(i - p * CV) * 8;
    const uint32_t t = p / (uint32_t)W;
    const int w = (int)(p - t * (uint32_t)W);
    const uint32_t n = t / (uint32_t)H;
    const int h = (int)(t - n * (uint32_t)H);
    const int sub = ch0 / Co, c0 = ch0 - sub * Co;
    const long q = ((long)n * 2 * H + 2 * h + (sub >> 1)) * 2 * W + 2 * w + (sub & 1);
    float v[8], al[8];
    if (alpha) ldc8(alpha + c0, al);
    if (!BWD) {
      V8<T>::ld(a + (p * av.pitch + av.off + ch0), v);
      if (alpha) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : al[j] * v[j];
      }
      V8<T>::st(o + (q * ov.pitch + ov.off + c0), v);
    } else {
      V8<T>::ld(a + (q * av.pitch + av.off + c0), v);
      if (alpha) {
        float uu[8];
        V8<T>::ld(u + (p * uv.pitch + uv.off + ch0), uu);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = uu[j] > 0.f ? v[j] : al[j] * v[j];
      }
      V8<T>::st(o + (p * ov.pitch + ov.off + ch0), v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(RT, 1)
d2s_dalpha8_kernel(const T* __restrict__ dy, VView dv, const T* __restrict__ u, VView uv, long Pout, int H2, int W2, int Co,
                   float* __restrict__ partial, unsigned* __restrict__ ticket, float* __restrict__ dalpha, int accumulate) {
  typedef RawPair<T, T> Raw;
  channel_reduce8<1, (sizeof(Raw) <= 32 ? 8 : 4), Raw>(
      Pout, Co, partial, ticket,
      [&](long q, int c0) {
        int w2 = (int)(q % W2);
        long t = q / W2;
        int h2 = (int)(t % H2);
        long n = t / H2;
        long p = (n * (H2 / 2) + h2 / 2) * (W2 / 2) + w2 / 2;
        int sub = (h2 & 1) * 2 + (w2 & 1);
        Raw r;
        r.g = V8<T>::ldraw(dy + (q * dv.pitch + dv.off + c0));
        r.x = V8<T>::ldraw(u + (p * uv.pitch + uv.off + sub * Co + c0));
        return r;
      },
      [&](const Raw& r, long, int, float (&a)[1][8]) {
        float g[8], uu[8];
        V8<T>::cvt(r.g, g);
        V8<T>::cvt(r.x, uu);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[0][j] = fmaf(g[j], fminf(uu[j], 0.f), a[0][j]);
      },
      [&](int c, const double* sums) { dalpha[c] = (accumulate ? dalpha[c] : 0.f) + (float)sums[c]; });
}

// depth_to_space(2)+PReLU backward with the slope gradient in the same pass: walks OUTPUT pixels q (thread = 8 fixed
// output channels), du[p, sub*Co + c] = u > 0 ? dy : alpha*dy and dalpha[c] += dy * min(u, 0); dy and u are read once.
template <typename T>
__global__ void __launch_bounds__(RT, 1)
d2s_prelu_bwd_fused8_kernel(const T* __restrict__ dy, VView dv, const T* __restrict__ u, VView uv, const float* __restrict__ alpha,
                            T* __restrict__ du, VView ov, long Pout, int H2, int W2, int Co, float* __restrict__ partial,
                            unsigned* __restrict__ ticket, float* __restrict__ dalpha, int accumulate) {
  typedef RawPair<T, T> Raw;
  float al[8];
  ldc8(alpha + (threadIdx.x % (Co >> 3)) * 8, al);
  auto src = [&](long q, long& p, int& sub) {
    int w2 = (int)(q % W2);
    long t = q / W2;
    int h2 = (int)(t % H2);
    long n = t / H2;
    p = (n * (H2 / 2) + h2 / 2) * (W2 / 2) + w2 / 2;
    sub = (h2 & 1) * 2 + (w2 & 1);
  };
  channel_reduce8<1, (sizeof(Raw) <= 32 ? 8 : 4), Raw>(
      Pout, Co, partial, ticket,
      [&](long q, int c0) {
        long p; int sub;
        src(q, p, sub);
        Raw r;
        r.g = V8<T>::ldraw(dy + (q * dv.pitch + dv.off + c0));
        r.x = V8<T>::ldraw(u + (p * uv.pitch + uv.off + sub * Co + c0));
        return r;
      },
      [&](const Raw& r, long q, int c0, float (&a)[1][8]) {
        long p; int sub;
        src(q, p, sub);
        float g[8], uu[8], o[8];
        V8<T>::cvt(r.g, g);
        V8<T>::cvt(r.x, uu);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[0][j] = fmaf(g[j], fminf(uu[j], 0.f), a[0][j]);
          o[j] = uu[j] > 0.f ? g[j] : al[j] * g[j];
        }
        V8<T>::st(du + (p * ov.pitch + ov.off + sub * Co + c0), o);
      },
      [&](int c, const double* sums) { dalpha[c] = (accumulate ? dalpha[c] : 0.f) + (float)sums[c]; });
}

// out = a (+ b) with dtype conversion; ACC: out += a
template <typename TA, typename TO, int MODE>  // MODE 0: out = a ; 1: out = a + b ; 2: out += a
__global__ void __launch_bounds__(VT)
ew8_kernel(const TA* __restrict__ a, VView av, const TA* __restrict__ b, VView bv, TO* __restrict__ o, VView ov, long P, int C) {
  const uint32_t CV = (uint32_t)C >> 3;
  const uint32_t total = (uint32_t)P * CV;   // vec_ok() guarantees < 2^31 elements
  for (uint32_t i = blockIdx.x * VT + threadIdx.x; i < total; i += gridDim.x * VT) {
    const uint32_t p = i / CV;
    const int c0 = (int)(i - p * CV) * 8;
    float v[8];
    V8<TA>::ld(a + (p * av.pitch + av.off + c0), v);
    if (MODE == 1) {
      float w[8];
      V8<TA>::ld(b + (p * bv.pitch + bv.off + c0), w);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += w[j];
    } else if (MODE == 2) {
      float w[8];
      V8<TO>::ld(o + (p * ov.pitch + ov.off + c0), w);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += w[j];
    }
    V8<TO>::st(o + (p * ov.pitch + ov.off + c0), v);
  }
}

// ---------------------------------------------------------------- MaxPool2D(2,2) / UpSampling2D(2)+ReLU, 8 channels per thread
// (autoencoder.py:110,113-136; every activation of the bf16 path has a channel count that is a multiple of 16, the odd
// widths 44/56/76/100/152/84 being physically zero-padded).  One thread = one pixel of the SMALL map x 8 channels: the four
// pixels of its 2x2 window are four independent 16-byte accesses.
template <typename T>
__global__ void __launch_bounds__(VT)
maxpool_fwd8_kernel(const T* __restrict__ x, VView xv, T* __restrict__ y, VView yv, int N, int Ho, int Wo, int C) {
  const uint32_t CV = (uint32_t)C >> 3, total = (uint32_t)N * Ho * Wo * CV, W = 2u * Wo;
  for (uint32_t i = blockIdx.x * VT + threadIdx.x; i < total; i += gridDim.x * VT) {
    const uint32_t q = i / CV, c0 = (i - q * CV) * 8u;
    const uint32_t wo = q % Wo, t = q / Wo, ho = t % Ho, n = t / Ho;
    const uint32_t p00 = (n * 2u * Ho + 2u * ho) * W + 2u * wo;
    typename V8<T>::raw r[4];
    r[0] = V8<T>::ldraw(x + ((size_t)p00 * xv.pitch + xv.off + c0));
    r[1] = V8<T>::ldraw(x + ((size_t)(p00 + 1) * xv.pitch + xv.off + c0));
    r[2] = V8<T>::ldraw(x + ((size_t)(p00 + W) * xv.pitch + xv.off + c0));
    r[3] = V8<T>::ldraw(x + ((size_t)(p00 + W + 1) * xv.pitch + xv.off + c0));
    float m[8], v[8];
    V8<T>::cvt(r[0], m);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      V8<T>::cvt(r[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
    }
    V8<T>::st(y + ((size_t)q * yv.pitch + yv.off + c0), m);
  }
}

// dx = dy routed to the first window position (row-major) whose value equals the max
template <typename T>
__global__ void __launch_bounds__(VT)
maxpool_bwd8_kernel(const T* __restrict__ dy, VView dv, const T* __restrict__ x, VView xv, const T* __restrict__ y, VView yv,
                    T* __restrict__ dx, VView ov, int N, int Ho, int Wo, int C, int relu) {
  const uint32_t CV = (uint32_t)C >> 3, total = (uint32_t)N * Ho * Wo * CV, W = 2u * Wo;
  for (uint32_t i = blockIdx.x * VT + threadIdx.x; i < total; i += gridDim.x * VT) {
    const uint32_t q = i / CV, c0 = (i - q * CV) * 8u;
    const uint32_t wo = q % Wo, t = q / Wo, ho = t % Ho, n = t / Ho;
    const uint32_t p00 = (n * 2u * Ho + 2u * ho) * W + 2u * wo;
    const uint32_t pos[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
    typename V8<T>::raw r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = V8<T>::ldraw(x + ((size_t)pos[k] * xv.pitch + xv.off + c0));
    float m[8], g[8];
    V8<T>::ld(y + ((size_t)q * yv.pitch + yv.off + c0), m);
    V8<T>::ld(dy + ((size_t)q * dv.pitch + dv.off + c0), g);
    bool done[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) done[j] = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8], o[8];
      V8<T>::cvt(r[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool hit = !done[j] && v[j] == m[j];
        o[j] = (hit && !(relu && v[j] <= 0.f)) ? g[j] : 0.f;     // relu: x is the output of a ReLU whose mask is applied here
        done[j] = done[j] || hit;
      }
      V8<T>::st(dx + ((size_t)pos[k] * ov.pitch + ov.off + c0), o);
    }
  }
}

// y[n,2h+i,2w+j,c] = relu(x[n,h,w,c]); y may be a channel slice of the concat buffer (autoencoder.py:135)
template <typename T>
__global__ void __launch_bounds__(VT)
upsample_relu_fwd8_kernel(const T* __restrict__ x, VView xv, T* __restrict__ y, VView yv, int N, int H, int W, int C) {
  const uint32_t CV = (uint32_t)C >> 3, total = (uint32_t)N * H * W * CV, W2 = 2u * W;
  for (uint32_t i = blockIdx.x * VT + threadIdx.x; i < total; i += gridDim.x * VT) {
    const uint32_t p = i / CV, c0 = (i - p * CV) * 8u;
    const uint32_t w = p % W, t = p / W, h = t % H, n = t / H;
    float v[8];
    V8<T>::ld(x + ((size_t)p * xv.pitch + xv.off + c0), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    const uint32_t q00 = (n * 2u * H + 2u * h) * W2 + 2u * w;
    V8<T>::st(y + ((size_t)q00 * yv.pitch + yv.off + c0), v);
    V8<T>::st(y + ((size_t)(q00 + 1) * yv.pitch + yv.off + c0), v);
    V8<T>::st(y + ((size_t)(q00 + W2) * yv.pitch + yv.off + c0), v);
    V8<T>::st(y + ((size_t)(q00 + W2 + 1) * yv.pitch + yv.off + c0), v);
  }
}

// dx[n,h,w,c] = (x > 0) * sum_{i,j} dy[n,2h+i,2w+j,c]; dy may be a channel slice of the concat gradient
template <typename T>
__global__ void __launch_bounds__(VT)
upsample_relu_bwd8_kernel(const T* __restrict__ dy, VView dv, const T* __restrict__ x, VView xv, T* __restrict__ dx, VView ov, int N,
                          int H, int W, int C) {
  const uint32_t CV = (uint32_t)C >> 3, total = (uint32_t)N * H * W * CV, W2 = 2u * W;
  for (uint32_t i = blockIdx.x * VT + threadIdx.x; i < total; i += gridDim.x * VT) {
    const uint32_t p = i / CV, c0 = (i - p * CV) * 8u;
    const uint32_t w = p % W, t = p / W, h = t % H, n = t / H;
    const uint32_t q00 = (n * 2u * H + 2u * h) * W2 + 2u * w;
    typename V8<T>::raw r[4];
    r[0] = V8<T>::ldraw(dy + ((size_t)q00 * dv.pitch + dv.off + c0));
    r[1] = V8<T>::ldraw(dy + ((size_t)(q00 + 1) * dv.pitch + dv.off + c0));
    r[2] = V8<T>::ldraw(dy + ((size_t)(q00 + W2) * dv.pitch + dv.off + c0));
    r[3] = V8<T>::ldraw(dy + ((size_t)(q00 + W2 + 1) * dv.pitch + dv.off + c0));
    float xv8[8], g[8], a[8];
    V8<T>::ld(x + ((size_t)p * xv.pitch + xv.off + c0), xv8);
    V8<T>::cvt(r[0], g);
    // same left-to-right order of the four additions as the scalar kernel
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      V8<T>::cvt(r[k], a);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += a[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = xv8[j] > 0.f ? g[j] : 0.f;
    V8<T>::st(dx + ((size_t)p * ov.pitch + ov.off + c0), g);
  }
}

static inline unsigned ew8_blocks(long total_vec, int sm_count) {
  long b = (total_vec + VT - 1) / VT, cap = (long)sm_count * 32;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace dgvec
