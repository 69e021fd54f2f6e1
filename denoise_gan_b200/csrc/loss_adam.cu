// Loss reductions (value + upstream gradient in one pass) and the fused Keras-Adam update.
// Reductions: warp shuffle -> shared memory -> per-block partial -> single-block final sum in
// double (fixed order, deterministic).
//
// Reference call sites: train_srgan.py:86-96 (BCE from logits, MSE, MAE, total variation),
// train_autoencoder.py:79-102 (BCE on probabilities), srgan.py:69-75 (feature MSE /12.75),
// srgan.py:35-50 + pix2pix.py:30-31 (Adam, ExponentialDecay staircase).
#include "dg_common.cuh"

namespace {

constexpr int LT = 256;

template <int NV>
__device__ __forceinline__ void block_reduce_store(float (&v)[NV], float* __restrict__ partial) {
  __shared__ float sm[NV][LT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float s = warp_sum(v[k]);
    if (lane == 0) sm[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float s = 0.f;
    for (int w = 0; w < LT / 32; ++w) s += sm[threadIdx.x][w];
    partial[(long)blockIdx.x * NV + threadIdx.x] = s;
  }
}

// out[k] = scale[k] * sum_b partial[b][k]
__global__ void final_sum_kernel(const float* __restrict__ partial, int nblocks, int nv, float s0, float s1, float s2,
                                 float* __restrict__ out) {
  __shared__ double sm[LT];
  for (int k = 0; k < nv; ++k) {
    double s = 0;
    for (int b = threadIdx.x; b < nblocks; b += LT) s += partial[(long)b * nv + k];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = LT / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = (float)(sm[0] * (double)(k == 0 ? s0 : (k == 1 ? s1 : s2)));
    __syncthreads();
  }
}

template <typename TG, typename TT, typename TD>
__global__ void __launch_bounds__(LT)
image_losses_kernel(const TG* __restrict__ gen, int gp, int go, const TT* __restrict__ tgt, int tp, int to, int N,
                    int H, int W, int C, float g_mae, float g_mse, float g_tv, TD* __restrict__ dgen, int dp, int dof,
                    int accumulate, float* __restrict__ partial) {
  const long total = (long)N * H * W * C;
  float acc[3] = {0.f, 0.f, 0.f};
  const bool idx32 = total < (1L << 31);    // 32-bit index arithmetic when the tensor allows it (three divisions per element)
  for (long i = (long)blockIdx.x * LT + threadIdx.x; i < total; i += (long)gridDim.x * LT) {
    int c, w, h;
    long p;
    if (idx32) {
      const unsigned iu = (unsigned)i, pu = iu / (unsigned)C, tu = pu / (unsigned)W;
      c = (int)(iu - pu * (unsigned)C); w = (int)(pu - tu * (unsigned)W); h = (int)(tu % (unsigned)H); p = (long)pu;
    } else {
      c = (int)(i % C); p = i / C; w = (int)(p % W); h = (int)((p / W) % H);
    }
    auto diff = [&](long pp) { return ld_f(tgt + (pp * tp + to + c)) - ld_f(gen + (pp * gp + go + c)); };
    float d = diff(p);
    acc[0] += fabsf(d);
    acc[1] += d * d;
    float grad_tv = 0.f;
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    if (h + 1 < H) {
      float e = diff(p + W) - d;
      acc[2] += fabsf(e);
      grad_tv -= sgn(e);
    }
    if (w + 1 < W) {
      float e = diff(p + 1) - d;
      acc[2] += fabsf(e);
      grad_tv -= sgn(e);
    }
    if (dgen) {
      if (g_tv != 0.f) {
        if (h > 0) grad_tv += sgn(d - diff(p - W));
        if (w > 0) grad_tv += sgn(d - diff(p - 1));
      }
      // d(loss)/d(gen) = -d(loss)/d(diff)
      float g = -(g_mae * sgn(d) + g_mse * 2.f * d + g_tv * grad_tv);
      TD* dst = dgen + (p * dp + dof + c);
      if (accumulate) g += ld_f(dst);
      st_f(dst, g);
    }
  }
  block_reduce_store<3>(acc, partial);
}

template <typename TX, typename TD>
__global__ void __launch_bounds__(LT)
bce_kernel(const TX* __restrict__ x, int xp, int xo, long P, int C, float z, int from_logits, float gscale,
           TD* __restrict__ dx, int dp, int dof, float* __restrict__ partial) {
  const long total = P * C;
  float acc[1] = {0.f};
  for (long i = (long)blockIdx.x * LT + threadIdx.x; i < total; i += (long)gridDim.x * LT) {
    long p = i / C;
    int c = (int)(i - p * C);
    float v = ld_f(x + (p * xp + xo + c));
    float loss, g;
    if (from_logits) {
      loss = fmaxf(v, 0.f) - v * z + log1pf(expf(-fabsf(v)));
      g = 1.f / (1.f + expf(-v)) - z;
    } else {
      const float eps = 1e-7f;
      float pc = fminf(fmaxf(v, eps), 1.f - eps);
      loss = -(z * logf(pc + eps) + (1.f - z) * logf(1.f - pc + eps));
      g = (v > eps && v < 1.f - eps) ? -(z / (pc + eps) - (1.f - z) / (1.f - pc + eps)) : 0.f;
    }
    acc[0] += loss;
    if (dx) st_f(dx + (p * dp + dof + c), g * gscale);
  }
  block_reduce_store<1>(acc, partial);
}

template <typename TA, typename TD>
__global__ void __launch_bounds__(LT)
feature_mse_kernel(const TA* __restrict__ a, int ap, int ao, const TA* __restrict__ b, int bp, int bo, long P, int C,
                   float gscale, TD* __restrict__ da, int dp, int dof, float* __restrict__ partial) {
  const long total = P * C;
  float acc[1] = {0.f};
  for (long i = (long)blockIdx.x * LT + threadIdx.x; i < total; i += (long)gridDim.x * LT) {
    long p = i / C;
    int c = (int)(i - p * C);
    float d = ld_f(a + (p * ap + ao + c)) - ld_f(b + (p * bp + bo + c));
    acc[0] += d * d;
    if (da) st_f(da + (p * dp + dof + c), gscale * d);
  }
  block_reduce_store<1>(acc, partial);
}

// state layout: int64 iterations | float lr_t | float pad
__global__ void adam_tick_kernel(int64_t* state, float lr0, float beta1, float beta2, int64_t decay_steps,
                                 float decay_rate) {
  int64_t it = state[0];
  double lr = (double)lr0;
  if (decay_steps > 0) lr *= pow((double)decay_rate, (double)(it / decay_steps));
  double t = (double)(it + 1);
  double lr_t = lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t));
  reinterpret_cast<float*>(state + 1)[0] = (float)lr_t;
  state[0] = it + 1;
}

__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, long numel, float beta1, float beta2, float eps, float gscale,
                            const int64_t* __restrict__ state) {
  const float lr_t = reinterpret_cast<const float*>(state + 1)[0];
  long i4 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < numel) {
    float4 g = *reinterpret_cast<const float4*>(grad + i4);
    float4 mm = *reinterpret_cast<float4*>(m + i4);
    float4 vv = *reinterpret_cast<float4*>(v + i4);
    float4 th = *reinterpret_cast<float4*>(theta + i4);
    float* gp = &g.x; float* mp = &mm.x; float* vp = &vv.x; float* tp = &th.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = gp[k] * gscale;
      mp[k] = beta1 * mp[k] + (1.f - beta1) * gg;
      vp[k] = beta2 * vp[k] + (1.f - beta2) * gg * gg;
      tp[k] -= lr_t * mp[k] / (sqrtf(vp[k]) + eps);
    }
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
    *reinterpret_cast<float4*>(theta + i4) = th;
  } else {
    for (long i = i4; i < numel; ++i) {
      float gg = grad[i] * gscale;
      float mi = beta1 * m[i] + (1.f - beta1) * gg;
      float vi = beta2 * v[i] + (1.f - beta2) * gg * gg;
      m[i] = mi;
      v[i] = vi;
      theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
  }
}

inline int loss_blocks(long total, int sm_count) {
  long b = (total + LT - 1) / LT;
  long cap = (long)sm_count * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}


// The scalar arithmetic between the loss reductions and the returned tuple (train_srgan.py:86-99, 118), one thread: explicitly rounded
// fp32 operations in the order of the reference expression so that the values equal the eager computation bit for bit.
__global__ void gan_loss_terms_kernel(const float* __restrict__ content, const float* __restrict__ adv_raw, const float* __restrict__ out3,
                                      const float* __restrict__ real_loss, const float* __restrict__ fake_loss, float w_mae, float w_mse,
                                      float tv_gain, float disc_scale, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float c = content ? content[0] : 0.f;
  const float adv = __fmul_rn(1e-3f, adv_raw[0]);
  const float mae = out3[0], mse = out3[1], var = __fmul_rn(1e-5f, out3[2]);
  float g = __fadd_rn(c, adv);
  g = __fadd_rn(g, __fmul_rn(mae, w_mae));
  g = __fadd_rn(g, __fmul_rn(mse, w_mse));
  g = __fadd_rn(g, __fmul_rn(var, tv_gain));
  out[0] = g; out[1] = adv; out[2] = mae; out[3] = mse; out[4] = c;
  out[5] = __fmul_rn(disc_scale, __fadd_rn(real_loss[0], fake_loss[0]));
  out[6] = var;
}
}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" size_t dg_loss_workspace_bytes(const dg_tensor* t) { return (size_t)256 * 8 * 3 * sizeof(float); }

extern "C" int dg_image_losses(dg_ctx* ctx, const dg_tensor* gen, const dg_tensor* target, float w_mae, float w_mse,
                               float w_tv, float* out3, const dg_tensor* dgen, int accumulate, void* workspace,
                               size_t workspace_bytes, void* stream) {
  DG_REQUIRE(dg_valid(gen) && dg_valid(target) && out3 && workspace, "dg_image_losses: null argument");
  DG_REQUIRE(dg_same_shape(gen, target), "dg_image_losses: shape mismatch");
  DG_REQUIRE(workspace_bytes >= dg_loss_workspace_bytes(gen), "dg_image_losses: workspace too small");
  if (dgen) DG_REQUIRE(dg_valid(dgen) && dg_same_shape(dgen, gen), "dg_image_losses: dgen mismatch");
  long total = dg_pixels(gen) * gen->c;
  int blocks = loss_blocks(total, ctx->sm_count);
  float inv = 1.f / (float)total, invn = 1.f / (float)gen->n;
  float* partial = (float*)workspace;
  int ddt = dgen ? dgen->dtype : DG_F32;
#define LAUNCH_IL(TG, TT, TD)                                                                                    \
  image_losses_kernel<TG, TT, TD><<<blocks, LT, 0, ST>>>(                                                        \
      (const TG*)gen->ptr, gen->cpitch, gen->coff, (const TT*)target->ptr, target->cpitch, target->coff, gen->n, \
      gen->h, gen->w, gen->c, w_mae * inv, w_mse * inv, w_tv * invn, dgen ? (TD*)dgen->ptr : nullptr,            \
      dgen ? dgen->cpitch : 0, dgen ? dgen->coff : 0, accumulate, partial)
  DG_REQUIRE(target->dtype == DG_F32, "dg_image_losses: target must be fp32");
  if (gen->dtype == DG_F32 && ddt == DG_F32) LAUNCH_IL(float, float, float);
  else if (gen->dtype == DG_F32 && ddt == DG_BF16) LAUNCH_IL(float, float, __nv_bfloat16);
  else if (gen->dtype == DG_BF16 && ddt == DG_BF16) LAUNCH_IL(__nv_bfloat16, float, __nv_bfloat16);
  else if (gen->dtype == DG_BF16 && ddt == DG_F32) LAUNCH_IL(__nv_bfloat16, float, float);
  else DG_FAIL("dg_image_losses: unsupported dtype");
#undef LAUNCH_IL
  final_sum_kernel<<<1, LT, 0, ST>>>(partial, blocks, 3, inv, inv, invn, out3);
  DG_CHECK_LAUNCH("dg_image_losses");
  return 0;
}

extern "C" int dg_bce_const_target(dg_ctx* ctx, const dg_tensor* x, float target, int from_logits, float grad_scale,
                                   float* loss_out, const dg_tensor* dx, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  DG_REQUIRE(dg_valid(x) && loss_out && workspace, "dg_bce_const_target: null argument");
  DG_REQUIRE(workspace_bytes >= dg_loss_workspace_bytes(x), "dg_bce_const_target: workspace too small");
  if (dx) DG_REQUIRE(dg_valid(dx) && dg_same_shape(dx, x), "dg_bce_const_target: dx mismatch");
  long P = dg_pixels(x);
  long total = P * x->c;
  int blocks = loss_blocks(total, ctx->sm_count);
  float inv = 1.f / (float)total;
  float* partial = (float*)workspace;
  DG_DISPATCH_2(x->dtype, dx ? dx->dtype : DG_F32, "dg_bce_const_target",
                bce_kernel<TI, TO><<<blocks, LT, 0, ST>>>((const TI*)x->ptr, x->cpitch, x->coff, P, x->c, target,
                                                         from_logits, grad_scale * inv, dx ? (TO*)dx->ptr : nullptr,
                                                         dx ? dx->cpitch : 0, dx ? dx->coff : 0, partial););
  final_sum_kernel<<<1, LT, 0, ST>>>(partial, blocks, 1, inv, 0.f, 0.f, loss_out);
  DG_CHECK_LAUNCH("dg_bce_const_target");
  return 0;
}

extern "C" int dg_feature_mse(dg_ctx* ctx, const dg_tensor* a, const dg_tensor* b, float inv_div, float* loss_out,
                              const dg_tensor* da, void* workspace, size_t workspace_bytes, void* stream) {
  DG_REQUIRE(dg_valid(a) && dg_valid(b) && loss_out && workspace, "dg_feature_mse: null argument");
  DG_REQUIRE(dg_same_shape(a, b) && a->dtype == b->dtype, "dg_feature_mse: shape/dtype mismatch");
  DG_REQUIRE(workspace_bytes >= dg_loss_workspace_bytes(a), "dg_feature_mse: workspace too small");
  if (da) DG_REQUIRE(dg_valid(da) && dg_same_shape(da, a), "dg_feature_mse: da mismatch");
  long P = dg_pixels(a);
  long total = P * a->c;
  int blocks = loss_blocks(total, ctx->sm_count);
  float s = inv_div * inv_div / (float)total;
  float* partial = (float*)workspace;
  DG_DISPATCH_2(a->dtype, da ? da->dtype : DG_F32, "dg_feature_mse",
                feature_mse_kernel<TI, TO><<<blocks, LT, 0, ST>>>((const TI*)a->ptr, a->cpitch, a->coff, (const TI*)b->ptr,
                                                                 b->cpitch, b->coff, P, a->c, 2.f * s,
                                                                 da ? (TO*)da->ptr : nullptr, da ? da->cpitch : 0,
                                                                 da ? da->coff : 0, partial););
  final_sum_kernel<<<1, LT, 0, ST>>>(partial, blocks, 1, s, 0.f, 0.f, loss_out);
  DG_CHECK_LAUNCH("dg_feature_mse");
  return 0;
}

extern "C" int dg_adam_step(dg_ctx* ctx, float* theta, const float* grad, float* m, float* v, int64_t numel, float lr0,
                            float beta1, float beta2, float eps, int64_t decay_steps, float decay_rate,
                            float grad_scale, int64_t* iterations_dev, void* stream) {
  DG_REQUIRE(theta && grad && m && v && iterations_dev && numel > 0, "dg_adam_step: null argument");
  DG_REQUIRE(((uintptr_t)theta | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) % 16 == 0,
             "dg_adam_step: arenas must be 16-byte aligned");
  adam_tick_kernel<<<1, 1, 0, ST>>>(iterations_dev, lr0, beta1, beta2, decay_steps, decay_rate);
  long threads = (numel + 3) / 4;
  adam_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ST>>>(theta, grad, m, v, numel, beta1, beta2, eps, grad_scale,
                                                                iterations_dev);
  DG_CHECK_LAUNCH("dg_adam_step");
  return 0;
}

// gen_loss = content + 1e-3 adv + w_mae mae + w_mse mse + tv_gain var, disc_loss = disc_scale (real + fake) (train_srgan.py:86-99);
// out[7] = {gen_loss, adv_loss, mae_loss, mse_loss, content_loss, disc_loss, var_loss}, the return order of train_srgan.py:118.
extern "C" int dg_gan_loss_terms(dg_ctx* ctx, const float* content, const float* adv_raw, const float* out3, const float* real_loss,
                                 const float* fake_loss, float w_mae, float w_mse, float tv_gain, float disc_scale, float* out7, void* stream) {
  DG_REQUIRE(adv_raw && out3 && real_loss && fake_loss && out7, "dg_gan_loss_terms: null argument");
  gan_loss_terms_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(content, adv_raw, out3, real_loss, fake_loss, w_mae, w_mse, tv_gain, disc_scale, out7);
  DG_CHECK_LAUNCH("dg_gan_loss_terms");
  return 0;
}
