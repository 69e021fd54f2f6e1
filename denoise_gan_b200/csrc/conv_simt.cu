// CUDA-core implicit-GEMM convolution: forward, input-gradient (= Conv2DTranspose forward) and
// weight-gradient, fp32 accumulation, fp32 or bf16 storage, arbitrary channel counts, kernel
// sizes, strides and explicit (asymmetric, TF 'SAME') padding.
//
// This is the fp32 parity tier (tcgen05 has no true-fp32 MMA) and the kernel for the layers the
// tensor-core path does not take (Cin or Cout not a multiple of 16: RGB first layers, 3- and
// 1-channel heads — all HBM-bound).
//
// Reference call sites: keras Conv2D srgan.py:154,246, autoencoder.py:95, pix2pix.py:115,207;
// Conv2DTranspose pix2pix.py:130,169; gradients via tape.gradient train_srgan.py:111-112.
#include "dg_common.cuh"
#include "conv_thin.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, NT = 256;

struct Geom {
  int N, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad_t, pad_l;
  int xp, xo;  // pitch / channel offset of the "image side" tensor (x or dx)
  int yp, yo;  // pitch / channel offset of the "output side" tensor (y or dy)
  int act;
  float alpha;
};

__device__ __forceinline__ void fma_tile(const float (&As)[TK][TM + 4], const float (&Bs)[TK][TN + 4],
                                         float (&acc)[4][4], int ty, int tx) {
#pragma unroll
  for (int kk = 0; kk < TK; ++kk) {
    float a[4], b[4];
    *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// ------------------------------------------------------------------ forward
template <typename TI, typename TO>
__global__ void __launch_bounds__(NT) conv_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, TO* __restrict__ y, Geom g) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const long M = (long)g.N * g.Ho * g.Wo;
  const int K = g.kh * g.kw * g.Cin;
  const long m0 = (long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  // A-load mapping: thread -> (4 rows m = tid/16 + 16 i, k lane = tid%16)
  const int a_kk = tid % 16;
  long a_base[4];
  int a_hi0[4], a_wi0[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long m = m0 + tid / 16 + 16 * i;
    if (m < M) {
      int wo = (int)(m % g.Wo);
      long t = m / g.Wo;
      int ho = (int)(t % g.Ho);
      int n = (int)(t / g.Ho);
      a_hi0[i] = ho * g.stride - g.pad_t;
      a_wi0[i] = wo * g.stride - g.pad_l;
      a_base[i] = (long)n * g.H * g.W;
    } else {
      a_hi0[i] = -1000000;
      a_wi0[i] = 0;
      a_base[i] = 0;
    }
  }
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    {
      int k = k0 + a_kk;
      int tap = k / g.Cin, c = k - tap * g.Cin;
      int r = tap / g.kw, s = tap - r * g.kw;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int hi = a_hi0[i] + r, wi = a_wi0[i] + s;
        float v = 0.f;
        if (k < K && hi >= 0 && hi < g.H && wi >= 0 && wi < g.W)
          v = ld_f(x + ((a_base[i] + (long)hi * g.W + wi) * g.xp + g.xo + c));
        As[a_kk][tid / 16 + 16 * i] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * NT;
      int kk = e / TN, j = e % TN;
      int k = k0 + kk;
      float v = 0.f;
      if (k < K && n0 + j < g.Cout) v = w[(long)k * g.Cout + n0 + j];
      Bs[kk][j] = v;
    }
    __syncthreads();
    fma_tile(As, Bs, acc, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int o = n0 + tx * 4 + j;
      if (o >= g.Cout) continue;
      float v = acc[i][j] + (bias ? bias[o] : 0.f);
      st_f(y + (m * g.yp + g.yo + o), apply_act(v, g.act, g.alpha));
    }
  }
}

// ------------------------------------------------------------------ dgrad / transposed-conv forward
// dx[n,hi,wi,c] = sum_{r,s,o} dy[n,(hi+pt-r)/st,(wi+pl-s)/st,o] * w[r,s,c,o]   (terms with a non-integer
// or out-of-range source position are zero)
template <typename TI, typename TO>
__global__ void __launch_bounds__(NT) conv_dgrad_kernel(const TI* __restrict__ dy, const float* __restrict__ w,
                                                        const float* __restrict__ bias, TO* __restrict__ dx, Geom g) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const long M = (long)g.N * g.H * g.W;
  const int K = g.kh * g.kw * g.Cout;
  const long m0 = (long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int a_kk = tid % 16;
  long a_base[4];
  int a_h[4], a_w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long m = m0 + tid / 16 + 16 * i;
    if (m < M) {
      int wi = (int)(m % g.W);
      long t = m / g.W;
      int hi = (int)(t % g.H);
      int n = (int)(t / g.H);
      a_h[i] = hi + g.pad_t;
      a_w[i] = wi + g.pad_l;
      a_base[i] = (long)n * g.Ho * g.Wo;
    } else {
      a_h[i] = -1000000;
      a_w[i] = 0;
      a_base[i] = 0;
    }
  }
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    {
      int k = k0 + a_kk;
      int tap = k / g.Cout, o = k - tap * g.Cout;
      int r = tap / g.kw, s = tap - r * g.kw;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int th = a_h[i] - r, tw = a_w[i] - s;
        float v = 0.f;
        if (k < K && th >= 0 && tw >= 0) {
          int ho = th / g.stride, wo = tw / g.stride;
          if (ho * g.stride == th && wo * g.stride == tw && ho < g.Ho && wo < g.Wo)
            v = ld_f(dy + ((a_base[i] + (long)ho * g.Wo + wo) * g.yp + g.yo + o));
        }
        As[a_kk][tid / 16 + 16 * i] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * NT;
      int j = e / TK, kk = e % TK;  // adjacent threads -> adjacent o (contiguous in w)
      int k = k0 + kk;
      float v = 0.f;
      if (k < K && n0 + j < g.Cin) {
        int tap = k / g.Cout, o = k - tap * g.Cout;
        v = w[((long)tap * g.Cin + n0 + j) * g.Cout + o];
      }
      Bs[kk][j] = v;
    }
    __syncthreads();
    fma_tile(As, Bs, acc, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = n0 + tx * 4 + j;
      if (c >= g.Cin) continue;
      float v = acc[i][j] + (bias ? bias[c] : 0.f);
      st_f(dx + (m * g.xp + g.xo + c), apply_act(v, g.act, g.alpha));
    }
  }
}

// ------------------------------------------------------------------ wgrad (split over pixels)
// part[z][(tap,c)][o] = sum_{pixels in slice z} x[n,ho*st+r-pt,wo*st+s-pl,c] * dy[n,ho,wo,o]
template <typename TI, typename TG>
__global__ void __launch_bounds__(NT) conv_wgrad_kernel(const TI* __restrict__ x, const TG* __restrict__ dy,
                                                        float* __restrict__ part, float* __restrict__ bias_part,
                                                        Geom g, long pix_per_split) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const long P = (long)g.N * g.Ho * g.Wo;
  const int MR = g.kh * g.kw * g.Cin;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const long p_beg = (long)blockIdx.z * pix_per_split;
  const long p_end = min(P, p_beg + pix_per_split);
  // A mapping: thread -> row m = tid % 64 (fixed), pixel lanes kk = tid/64 + 4 i
  const int am = m0 + tid % TM;
  int a_r = 0, a_s = 0, a_c = 0;
  const bool a_ok = am < MR;
  if (a_ok) {
    int tap = am / g.Cin;
    a_c = am - tap * g.Cin;
    a_r = tap / g.kw;
    a_s = tap - a_r * g.kw;
  }
  float acc[4][4] = {};
  float bsum = 0.f;
  for (long p0 = p_beg; p0 < p_end; p0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk = tid / TM + 4 * i;
      long p = p0 + kk;
      float v = 0.f;
      if (a_ok && p < p_end) {
        int wo = (int)(p % g.Wo);
        long t = p / g.Wo;
        int ho = (int)(t % g.Ho);
        int n = (int)(t / g.Ho);
        int hi = ho * g.stride - g.pad_t + a_r, wi = wo * g.stride - g.pad_l + a_s;
        if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W)
          v = ld_f(x + ((((long)n * g.H + hi) * g.W + wi) * g.xp + g.xo + a_c));
      }
      As[kk][tid % TM] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * NT;
      int kk = e / TN, j = e % TN;
      long p = p0 + kk;
      float v = 0.f;
      if (p < p_end && n0 + j < g.Cout) v = ld_f(dy + (p * g.yp + g.yo + n0 + j));
      Bs[kk][j] = v;
    }
    __syncthreads();
    fma_tile(As, Bs, acc, ty, tx);
    if (bias_part && blockIdx.x == 0 && tid < TN) {
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) bsum += Bs[kk][tid];
    }
    __syncthreads();
  }
  float* out = part + (long)blockIdx.z * MR * g.Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= MR) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int o = n0 + tx * 4 + j;
      if (o < g.Cout) out[(long)m * g.Cout + o] = acc[i][j];
    }
  }
  if (bias_part && blockIdx.x == 0 && tid < TN && n0 + tid < g.Cout)
    bias_part[(long)blockIdx.z * g.Cout + n0 + tid] = bsum;
}

// dst[i] (+)= sum_z part[z][i], fixed order (deterministic)
__global__ void reduce_splits_kernel(const float* __restrict__ part, float* __restrict__ dst, long numel, int splits,
                                     int accumulate) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numel) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(long)z * numel + i];
  dst[i] = accumulate ? dst[i] + s : s;
}

// dst[i] (+)= sum_z part[z*stride + i] for i < n0 (dst0) and n0 <= i < n0+n1 (dst1): block = 32 elements x 32 groups of
// splits, four independent loads in flight per thread, fixed summation order (deterministic).
__global__ void __launch_bounds__(1024)
reduce_strided_kernel(const float* __restrict__ part, long stride, float* __restrict__ dst0, long n0, float* __restrict__ dst1,
                      long n1, int splits, int accumulate) {
  __shared__ float sm[32][33];
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long i = (long)blockIdx.x * 32 + e;
  float s = 0.f;
  if (i < n0 + n1) {
    const float* src = part + i;
    int z = g;
    for (; z + 96 < splits; z += 128) {
      const float a = __ldg(src + (long)z * stride), b = __ldg(src + (long)(z + 32) * stride);
      const float c = __ldg(src + (long)(z + 64) * stride), d = __ldg(src + (long)(z + 96) * stride);
      s += a; s += b; s += c; s += d;
    }
    for (; z < splits; z += 32) s += __ldg(src + (long)z * stride);
  }
  sm[g][e] = s;
  __syncthreads();
  if (g == 0 && i < n0 + n1) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][e];
    float* d = i < n0 ? dst0 + i : dst1 + (i - n0);
    *d = accumulate ? *d + t : t;
  }
}

int fill_geom(const char* name, const dg_tensor* img, const dg_tensor* out, const dg_conv_params* p, Geom* g) {
  DG_REQUIRE(dg_valid(img) && dg_valid(out) && p, "%s: invalid tensor", name);
  DG_REQUIRE(p->kh >= 1 && p->kw >= 1 && p->stride >= 1, "%s: bad kernel geometry", name);
  DG_REQUIRE(img->n == out->n, "%s: batch mismatch", name);
  g->N = img->n; g->H = img->h; g->W = img->w; g->Cin = img->c;
  g->Ho = out->h; g->Wo = out->w; g->Cout = out->c;
  g->kh = p->kh; g->kw = p->kw; g->stride = p->stride; g->pad_t = p->pad_t; g->pad_l = p->pad_l;
  g->xp = img->cpitch; g->xo = img->coff; g->yp = out->cpitch; g->yo = out->coff;
  g->act = p->act; g->alpha = p->act_alpha;
  // every output position must read a window that starts inside the padded image
  DG_REQUIRE((g->Ho - 1) * g->stride - g->pad_t < g->H && (g->Wo - 1) * g->stride - g->pad_l < g->W,
             "%s: output size %dx%d inconsistent with input %dx%d", name, g->Ho, g->Wo, g->H, g->W);
  return 0;
}


// ------------------------------------------------------------------ thin-layer dispatch (conv_thin.cuh)
using dgthin::ThinGeom;

bool thin_shape_ok(const dg_tensor* img, const dg_tensor* out, const dg_conv_params* p) {
  if ((long)img->n * img->h * img->w >= (1L << 31)) return false;
  return p->stride == 1 && img->h == out->h && img->w == out->w && p->kh * p->kw <= 16 && p->kw <= dgthin::OUT_MAX_KW;
}
bool fat_ok(const dg_tensor* t, int mult) { return dgvec::vec_ok(t) && t->c % mult == 0; }

ThinGeom thin_geom(const dg_tensor* img, const dg_conv_params* p, int cin, int cout) {
  ThinGeom g;
  g.N = img->n; g.H = img->h; g.W = img->w; g.kh = p->kh; g.kw = p->kw;
  g.dh0 = p->pad_t; g.dw0 = p->pad_l; g.dsign = 1;
  g.w_st = cin * cout;
  g.act = p->act; g.alpha = p->act_alpha;
  return g;
}

// returns -1 when the layer is not "thin", else 0/1 like the public entry points
template <int MODE>  // 0 forward (in=x, out=y), 1 dgrad (in=dy, out=dx)
int try_thin(dg_ctx* ctx, const char* name, const dg_tensor* in, const float* w, const float* bias, const dg_tensor* out,
             const dg_conv_params* p, cudaStream_t st) {
  const dg_tensor* img = MODE == 0 ? in : out;   // x-side tensor
  const dg_tensor* oth = MODE == 0 ? out : in;   // y-side tensor
  if (!thin_shape_ok(img, oth, p)) return -1;
  const int cin = img->c, cout = oth->c;
  ThinGeom g = thin_geom(img, p, cin, cout);
  g.dsign = MODE == 0 ? 1 : -1;
  const long P = dg_pixels(in);
  const int in_c = in->c, out_c = out->c;
  // weight strides (tap, thin, fat) for W[t][ci][co]
  const int s_ci = cout, s_co = 1;
  if (in_c <= 4 && fat_ok(out, 16)) {           // expand: thin input -> fat output
    g.CT = in_c; g.CF = out_c; g.tp = in->cpitch; g.to = in->coff; g.fp = out->cpitch; g.fo = out->coff;
    g.w_sthin = MODE == 0 ? s_ci : s_co; g.w_sfat = MODE == 0 ? s_co : s_ci;
    dim3 grid((unsigned)((P + dgthin::EXP_THREADS - 1) / dgthin::EXP_THREADS), out_c / dgthin::EXP_OG);
    DG_DISPATCH_2(in->dtype, out->dtype, name,
                  dgthin::thin_expand_kernel<TI, TO><<<grid, dgthin::EXP_THREADS, 0, st>>>((const TI*)in->ptr, (TO*)out->ptr, w, bias, g););
    DG_CHECK_LAUNCH(name);
    return 0;
  }
  if (out_c <= 4 && fat_ok(in, 8)) {            // contract: fat input -> thin output
    g.CT = out_c; g.CF = in_c; g.tp = out->cpitch; g.to = out->coff; g.fp = in->cpitch; g.fo = in->coff;
    g.w_sthin = MODE == 0 ? s_co : s_ci; g.w_sfat = MODE == 0 ? s_ci : s_co;
    size_t smem = (size_t)g.kh * g.kw * g.CF * sizeof(float4);
    if (smem > 40 * 1024) return -1;
    unsigned grid = (unsigned)((P + dgthin::CON_THREADS - 1) / dgthin::CON_THREADS);
    DG_DISPATCH_2(in->dtype, out->dtype, name,
                  dgthin::thin_contract_kernel<TI, TO><<<grid, dgthin::CON_THREADS, smem, st>>>((const TI*)in->ptr, (TO*)out->ptr, w, bias, g););
    DG_CHECK_LAUNCH(name);
    return 0;
  }
  return -1;
}

constexpr int THIN_WGRAD_BLOCKS = 592;  // 4 x 148 partial blocks

bool thin_wgrad_applicable(const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p) {
  if (!thin_shape_ok(x, dy, p)) return false;
  const dg_tensor* fat = x->c <= 4 ? dy : x;
  const dg_tensor* thin = x->c <= 4 ? x : dy;
  if (thin->c > 4 || thin->c < 1) return false;
  if (!fat_ok(fat, 8)) return false;
  int warps = fat->c / 8;
  return warps >= 1 && warps <= 16 && (long)x->n * x->h * x->w < (1L << 31);
}

int thin_wgrad(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* dy, float* dw, float* dbias, const dg_conv_params* p,
               int accumulate, void* workspace, cudaStream_t st) {
  const char* name = "dg_conv2d_wgrad(thin)";
  const bool thin_in = x->c <= 4;
  const dg_tensor* fat = thin_in ? dy : x;
  const dg_tensor* thin = thin_in ? x : dy;
  const int cin = x->c, cout = dy->c;
  ThinGeom g = thin_geom(x, p, cin, cout);
  g.CT = thin->c; g.CF = fat->c; g.tp = thin->cpitch; g.to = thin->coff; g.fp = fat->cpitch; g.fo = fat->coff;
  g.dsign = thin_in ? 1 : -1;
  g.w_sthin = thin_in ? cout : 1;
  g.w_sfat = thin_in ? 1 : cout;
  const long P = dg_pixels(x);
  const long n_dw = (long)p->kh * p->kw * cin * cout;
  const long part_stride = n_dw + cout;
  long per = (P + THIN_WGRAD_BLOCKS - 1) / THIN_WGRAD_BLOCKS;
  per = (per + 31) / 32 * 32;
  const int blocks = (int)((P + per - 1) / per);
  const int threads = 32 * (fat->c / 8);
  const dim3 grid_o(blocks, p->kh);
  float* part = (float*)workspace;
  const int want_bias = dbias != nullptr;
#define THIN_OUTER_KW(CTN, KWN)                                                                                           \
  DG_DISPATCH_2(fat->dtype, thin->dtype, name,                                                                            \
                dgthin::thin_outer_kernel<TI, TO, CTN, KWN><<<grid_o, threads, 0, st>>>((const TI*)fat->ptr, (const TO*)thin->ptr, part, \
                                                                                       part_stride, n_dw, thin_in ? 1 : 0, want_bias, g, per);)
#define THIN_OUTER(CTN)                   \
  switch (p->kw) {                        \
    case 1: THIN_OUTER_KW(CTN, 1); break; \
    case 2: THIN_OUTER_KW(CTN, 2); break; \
    case 3: THIN_OUTER_KW(CTN, 3); break; \
    default: THIN_OUTER_KW(CTN, 4); break; \
  }
  switch (thin->c) {
    case 1: THIN_OUTER(1); break;
    case 2: THIN_OUTER(2); break;
    case 3: THIN_OUTER(3); break;
    default: THIN_OUTER(4); break;
  }
#undef THIN_OUTER
#undef THIN_OUTER_KW
  DG_CHECK_LAUNCH(name);
  const long n_bias = dbias ? cout : 0;
  reduce_strided_kernel<<<(unsigned)((n_dw + n_bias + 31) / 32), 1024, 0, st>>>(part, part_stride, dw, n_dw, dbias, n_bias, blocks, accumulate);
  DG_CHECK_LAUNCH(name);
  return 0;
}

int wgrad_splits(long P) {
  long s = (P + 4095) / 4096;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return (int)s;
}

}  // namespace

extern "C" int dg_conv2d_fwd(dg_ctx* ctx, const dg_tensor* x, const float* w, const float* bias, const dg_tensor* y,
                             const dg_conv_params* p, void* stream) {
  Geom g;
  if (fill_geom("dg_conv2d_fwd", x, y, p, &g)) return 1;
  DG_REQUIRE(w, "dg_conv2d_fwd: null weights");
  DG_REQUIRE(p->act != DG_ACT_PRELU, "dg_conv2d_fwd: PReLU is not a conv epilogue");
  {
    int r = try_thin<0>(ctx, "dg_conv2d_fwd(thin)", x, w, bias, y, p, (cudaStream_t)stream);
    if (r >= 0) return r;
  }
  long M = (long)g.N * g.Ho * g.Wo;
  dim3 grid((unsigned)((M + TM - 1) / TM), (g.Cout + TN - 1) / TN);
  DG_DISPATCH_2(x->dtype, y->dtype, "dg_conv2d_fwd",
                conv_fwd_kernel<TI, TO><<<grid, NT, 0, (cudaStream_t)stream>>>((const TI*)x->ptr, w, bias, (TO*)y->ptr, g););
  DG_CHECK_LAUNCH("dg_conv2d_fwd");
  return 0;
}

extern "C" int dg_conv2d_dgrad(dg_ctx* ctx, const dg_tensor* dy, const float* w, const float* bias,
                               const dg_tensor* dx, const dg_conv_params* p, void* stream) {
  Geom g;
  if (fill_geom("dg_conv2d_dgrad", dx, dy, p, &g)) return 1;
  DG_REQUIRE(w, "dg_conv2d_dgrad: null weights");
  DG_REQUIRE(p->act != DG_ACT_PRELU, "dg_conv2d_dgrad: PReLU is not a conv epilogue");
  {
    int r = try_thin<1>(ctx, "dg_conv2d_dgrad(thin)", dy, w, bias, dx, p, (cudaStream_t)stream);
    if (r >= 0) return r;
  }
  long M = (long)g.N * g.H * g.W;
  dim3 grid((unsigned)((M + TM - 1) / TM), (g.Cin + TN - 1) / TN);
  DG_DISPATCH_2(dy->dtype, dx->dtype, "dg_conv2d_dgrad",
                conv_dgrad_kernel<TI, TO><<<grid, NT, 0, (cudaStream_t)stream>>>((const TI*)dy->ptr, w, bias, (TO*)dx->ptr, g););
  DG_CHECK_LAUNCH("dg_conv2d_dgrad");
  return 0;
}

extern "C" size_t dg_conv2d_wgrad_workspace_bytes(const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p) {
  long P = (long)dy->n * dy->h * dy->w;
  int splits = wgrad_splits(P);
  if (thin_wgrad_applicable(x, dy, p)) splits = THIN_WGRAD_BLOCKS;  // the thin-layer path writes up to this many partials
  return (size_t)splits * ((size_t)p->kh * p->kw * x->c * dy->c + dy->c) * sizeof(float);
}

extern "C" int dg_conv2d_wgrad(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* dy, float* dw, float* dbias,
                               const dg_conv_params* p, int accumulate, void* workspace, size_t workspace_bytes,
                               void* stream) {
  Geom g;
  if (fill_geom("dg_conv2d_wgrad", x, dy, p, &g)) return 1;
  DG_REQUIRE(dw && workspace, "dg_conv2d_wgrad: null output/workspace");
  DG_REQUIRE(workspace_bytes >= dg_conv2d_wgrad_workspace_bytes(x, dy, p), "dg_conv2d_wgrad: workspace too small");
  if (thin_wgrad_applicable(x, dy, p)) return thin_wgrad(ctx, x, dy, dw, dbias, p, accumulate, workspace, (cudaStream_t)stream);
  long P = (long)g.N * g.Ho * g.Wo;
  int splits = wgrad_splits(P);
  long per = ((P + splits - 1) / splits + TK - 1) / TK * TK;
  int MR = g.kh * g.kw * g.Cin;
  float* part = (float*)workspace;
  float* bias_part = dbias ? part + (size_t)splits * MR * g.Cout : nullptr;
  dim3 grid((MR + TM - 1) / TM, (g.Cout + TN - 1) / TN, splits);
  cudaStream_t st = (cudaStream_t)stream;
  DG_DISPATCH_2(x->dtype, dy->dtype, "dg_conv2d_wgrad",
                conv_wgrad_kernel<TI, TO><<<grid, NT, 0, st>>>((const TI*)x->ptr, (const TO*)dy->ptr, part, bias_part, g, per););
  DG_CHECK_LAUNCH("dg_conv2d_wgrad");
  long numel = (long)MR * g.Cout;
  reduce_splits_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, st>>>(part, dw, numel, splits, accumulate);
  if (dbias) reduce_splits_kernel<<<(g.Cout + 255) / 256, 256, 0, st>>>(bias_part, dbias, g.Cout, splits, accumulate);
  DG_CHECK_LAUNCH("dg_conv2d_wgrad(reduce)");
  return 0;
}
