// Frame pre/post-processing around the inference forward (infer_video.py:138-159, infer.py:50-68,
// unit_test.py:67-86): uint8 HWC frames <-> float NHWC network tensors, with the centre crop-or-pad of
// tf.image.resize_with_crop_or_pad, the [0,255] <-> [-1,1] / [0,1] scaling and the BGR<->RGB flip done in
// one pass on the device so only uint8 crosses PCIe.
#include "dg_common.cuh"

namespace {

// centre crop-or-pad offset: source coordinate = destination coordinate + off (negative => padding before)
__host__ __device__ inline int crop_or_pad_off(int src, int dst) { return src >= dst ? (src - dst) / 2 : -((dst - src) / 2); }

template <typename TO>
__global__ void __launch_bounds__(256)
frame_to_float_kernel(const uint8_t* __restrict__ src, int n, int sh, int sw, int flip, int norm, float scale, float offset,
                      TO* __restrict__ dst, int dh, int dw, int pitch, int coff) {
  const long total = (long)n * dh * dw;
  const int oy = crop_or_pad_off(sh, dh), ox = crop_or_pad_off(sw, dw);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % dw);
    const long r = i / dw;
    const int y = (int)(r % dh);
    const int b = (int)(r / dh);
    const int sy = y + oy, sx = x + ox;
    float v[3] = {0.f, 0.f, 0.f};
    if (sy >= 0 && sy < sh && sx >= 0 && sx < sw) {
      const uint8_t* p = src + (((long)b * sh + sy) * sw + sx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float u = (float)p[flip ? 2 - c : c];
        // 0: tf.image.convert_image_dtype (float32 multiply by 1/255); 1: numpy float64 '/ 255.0' then cast; 2: float32 divide
        v[c] = norm == 0 ? __fmul_rn(u, 1.0f / 255.0f) : norm == 1 ? (float)((double)u / 255.0) : __fdiv_rn(u, 255.0f);
      }
    }
    TO* q = dst + i * pitch + coff;
#pragma unroll
    for (int c = 0; c < 3; ++c) st_f<TO>(q + c, __fmaf_rn(v[c], scale, offset));
  }
}

template <typename TI>
__global__ void __launch_bounds__(256)
float_to_frame_kernel(const TI* __restrict__ src, int n, int sh, int sw, int pitch, int coff, float scale, float offset, int clip01,
                      int flip, uint8_t* __restrict__ dst, int dh, int dw) {
  const long total = (long)n * dh * dw;
  const int oy = crop_or_pad_off(sh, dh), ox = crop_or_pad_off(sw, dw);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % dw);
    const long r = i / dw;
    const int y = (int)(r % dh);
    const int b = (int)(r / dh);
    const int sy = y + oy, sx = x + ox;
    float v[3] = {0.f, 0.f, 0.f};
    if (sy >= 0 && sy < sh && sx >= 0 && sx < sw) {
      const TI* p = src + (((long)b * sh + sy) * sw + sx) * pitch + coff;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __fmaf_rn(ld_f<TI>(p + c), scale, offset);
    }
    uint8_t* q = dst + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float f = v[c];
      if (clip01) f = fminf(fmaxf(f, 0.f), 1.f);
      f = __fmul_rn(f, 255.0f);
      f = fminf(fmaxf(f, 0.f), 255.f);          // astype(uint8) truncates; out-of-range inputs saturate here
      q[flip ? 2 - c : c] = (uint8_t)(int)f;
    }
  }
}

inline int frame_blocks(long total, int sms) {
  long b = (total + 255) / 256;
  long cap = (long)sms * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" int dg_frame_to_float(dg_ctx* ctx, const uint8_t* src, int src_h, int src_w, int flip_channels, int norm_mode, float scale,
                                 float offset, const dg_tensor* dst, void* stream) {
  DG_REQUIRE(ctx && src && dg_valid(dst), "dg_frame_to_float: null argument");
  DG_REQUIRE(dst->c == 3 && src_h > 0 && src_w > 0, "dg_frame_to_float: expects 3-channel frames");
  DG_REQUIRE(norm_mode >= 0 && norm_mode <= 2, "dg_frame_to_float: unknown normalisation mode %d", norm_mode);
  const long total = dg_pixels(dst);
  DG_DISPATCH_1(dst->dtype, "dg_frame_to_float",
                frame_to_float_kernel<T><<<frame_blocks(total, ctx->sm_count), 256, 0, (cudaStream_t)stream>>>(
                    src, dst->n, src_h, src_w, flip_channels, norm_mode, scale, offset, (T*)dst->ptr, dst->h, dst->w, dst->cpitch,
                    dst->coff););
  DG_CHECK_LAUNCH("dg_frame_to_float");
  return 0;
}

extern "C" int dg_float_to_frame(dg_ctx* ctx, const dg_tensor* src, float scale, float offset, int clip01, int flip_channels,
                                 uint8_t* dst, int dst_h, int dst_w, void* stream) {
  DG_REQUIRE(ctx && dst && dg_valid(src), "dg_float_to_frame: null argument");
  DG_REQUIRE(src->c == 3 && dst_h > 0 && dst_w > 0, "dg_float_to_frame: expects 3-channel frames");
  const long total = (long)src->n * dst_h * dst_w;
  DG_DISPATCH_1(src->dtype, "dg_float_to_frame",
                float_to_frame_kernel<T><<<frame_blocks(total, ctx->sm_count), 256, 0, (cudaStream_t)stream>>>(
                    (const T*)src->ptr, src->n, src->h, src->w, src->cpitch, src->coff, scale, offset, clip01, flip_channels, dst,
                    dst_h, dst_w););
  DG_CHECK_LAUNCH("dg_float_to_frame");
  return 0;
}
