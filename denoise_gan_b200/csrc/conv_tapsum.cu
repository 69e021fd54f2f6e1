// 3x3 SAME stride-1 convolution from 32 bf16 channels to the 1..3 channels of an image at INFERENCE (the last layer of the
// Fast-SRGAN generator, fsrgan.py:216-217, as infer_video.py:146 runs it on a 4x up-scaled 1080p frame: 5120 x 8192 pixels)
// in "tap-sum" form.
//
// The implicit-GEMM kernel (conv_umma.cu) spends one 128 x 16 x 16 tcgen05.mma per filter tap and 16-channel K step: 18
// instructions per 128 pixels whose N is padded from 3 to 16, and an instruction occupies the tensor pipe 44 cycles however
// narrow N is (it re-reads its 128 x 16 A operand from shared memory: probes/umma_probe.cu) -- 792 cycles per 128 pixels,
// 1.30 ms for the frame, four times what HBM needs to deliver the input.  With so few output channels the nine taps fit the
// N dimension instead:
//     D[q, tap*Cout + co] = sum_ci x[q, ci] * w[tap, ci, co]         ONE product per pixel q of the tile's halo box,
//                                                                    N = 9*Cout <= 27 (padded to 32), K = 32: 2 instructions per 128 pixels
//     y[p, co]            = bias[co] + sum_tap D[p + offset(tap), tap*Cout + co]      27 adds per pixel on the CUDA cores
// Out-of-image pixels of the halo box are zero-filled by the TMA unit, so their D rows are zero: SAME padding.
//
// One persistent CTA per SM walks 30 x 14 output tiles (32 x 16 halo box = 512 pixels = four 128-row M blocks):
//   warp 0  TMA producer: the 32 KB halo box of a tile into a 4-stage ring (64-byte rows, 64-byte swizzle);
//   warp 1  eight tcgen05.mma per tile into one of four 128-column accumulators, tcgen05.commit frees the stage;
//   4 x 4 epilogue warps, the four groups take tiles round-robin.  A halo row is 32 pixels = the 32 accumulator lanes one warp
//           may read, so a warp holds whole halo rows (quarter q of M block mb = halo row 4 mb + q), lane = column:
//             H  the three horizontal taps of every filter row meet in the warp -- two shuffles per (filter row, channel);
//                the 3*Cout partial sums per pixel go to an fp32 scratch [3*Cout][512] in shared memory (18 KB per group);
//             V  after the group's barrier a warp owns output rows, lane = column: three scratch reads per channel (filter
//                rows, conflict-free), bias, activation, store.
//           (The first version wrote all 9*Cout columns to a 55 KB scratch and gathered 27 values per pixel with pixels dealt to
//           threads linearly: a 30-wide tile row never matches the 32 lanes, every read was a 2-way bank conflict and the kernel
//           sat at 56 % of the shared-memory pipe with half its issue slots stalled on it: 712 us.)
// Two stores: the dense fp32 NHWC tensor ('generator_tanh', dtype float32), or -- dg_conv3x3_tapsum_frame -- straight to the
// uint8 frame with the arithmetic of dg_float_to_frame (infer_video.py:150-159: (y+1)/2, clip, *255, truncate, centre crop):
// then only the tiles inside the crop window are computed, and the fp32 image (500 MB written and read back at this size)
// never exists.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "dg_common.cuh"
#include "sm100.cuh"

namespace {
using namespace sm100;

constexpr int TS_C = 32;                          // input channels (one 64-byte row per pixel)
constexpr int TS_HW = 32, TS_HH = 16;             // halo box
constexpr int TS_OW = TS_HW - 2, TS_OH = TS_HH - 2;
constexpr int TS_PIX = TS_HW * TS_HH;             // 512 rows of the product
constexpr int TS_MB = TS_PIX / 128;               // M blocks
constexpr int TS_NST = 4;                         // halo stages
constexpr int TS_N = 32;                          // accumulator columns per M block
constexpr int TS_NG = 4, TS_GW = 4;               // epilogue groups, warps per group (one per TMEM lane quarter)
constexpr int TS_THREADS = (2 + TS_NG * TS_GW) * 32;
constexpr uint32_t TS_STAGE = TS_PIX * TS_C * 2;  // 32 KB
constexpr uint32_t OFF_W = 0;                     // B operand: 32 rows (tap*Cout + co) x 64 B (bf16 over ci), 64-byte swizzle
constexpr uint32_t OFF_BIAS = 2048;
constexpr uint32_t OFF_X = 4096;
constexpr uint32_t OFF_S = OFF_X + TS_NST * TS_STAGE;
constexpr uint32_t TS_SCRATCH = 9 * TS_PIX * 4;   // per group: [3 filter rows x Cout <= 3][512] fp32
constexpr uint32_t TS_SMEM = OFF_S + TS_NG * TS_SCRATCH + 1024;

// tcgen05.wait::ld that also "modifies" the loaded registers, so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

struct TsParams {
  CUtensorMap xmap;
  const float* w;        // [3][3][32][Cout] (Keras HWIO), fp32; rounded to bf16 here like every tensor-core operand of the library
  const float* bias;     // [Cout] or null
  float* y;              // fp32 NHWC output (pixel pitch yp floats), or null when `frame` is set
  uint8_t* frame;        // [N][OH][OW][3] uint8
  int yp, N, H, W;
  int act;
  float alpha;
  int oy, ox, OH, OW;    // the window of the convolution's output that is computed (the whole image, or the frame's centre crop)
  int tiles_w, tiles_h, total;
  float scale, offset;
  int clip01, flip;
};

template <int COUT>
__global__ void __launch_bounds__(TS_THREADS, 1) conv_tapsum_kernel(const __grid_constant__ TsParams P) {
  constexpr int NCOL = 9 * COUT;
  extern __shared__ uint8_t ts_raw[];
  __shared__ __align__(8) uint64_t bar_full[TS_NST], bar_empty[TS_NST], bar_acc[TS_NG], bar_accfree[TS_NG];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(ts_raw) + 1023u) & ~1023u;
  uint8_t* gen = ts_raw + (base - smem_u32(ts_raw));
  const int n_local = ((int)blockIdx.x < P.total) ? (P.total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (tid == 0) {
    for (int s = 0; s < TS_NST; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    for (int g = 0; g < TS_NG; ++g) { mbar_init(smem_u32(&bar_acc[g]), 1); mbar_init(smem_u32(&bar_accfree[g]), 1); }
    fence_mbar_init();
    tma_prefetch_desc(&P.xmap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), TS_NG * TS_MB * TS_N);
    tmem_relinquish();
  }
  // B operand: row n = tap*COUT + co holds w[tap][0..31][co] as bf16 (K-major), rows >= 9*COUT are zero
  for (int i = tid; i < TS_N * 4; i += TS_THREADS) {
    const int n = i >> 2, c = i & 3;
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    if (n < NCOL) {
      const int tap = n / COUT, co = n - tap * COUT;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ci = c * 8 + e * 2;
        const __nv_bfloat16 lo = __float2bfloat16(P.w[(tap * TS_C + ci) * COUT + co]);
        const __nv_bfloat16 hi = __float2bfloat16(P.w[(tap * TS_C + ci + 1) * COUT + co]);
        pk[e] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      }
    }
    *reinterpret_cast<uint4*>(gen + OFF_W + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  float* bias_s = reinterpret_cast<float*>(gen + OFF_BIAS);
  if (tid < 4) bias_s[tid] = (P.bias && tid < COUT) ? P.bias[tid] : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  auto tile_coords = [&](int it, int& n, int& h0, int& w0) {
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    const int tw = tile % P.tiles_w, t2 = tile / P.tiles_w, th = t2 % P.tiles_h;
    n = t2 / P.tiles_h; h0 = P.oy + th * TS_OH; w0 = P.ox + tw * TS_OW;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    for (int it = 0; it < n_local; ++it) {
      const int s = it % TS_NST, use = it / TS_NST;
      if (use > 0) mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(use - 1) & 1u);
      if (elect_one()) {
        int n, h0, w0;
        tile_coords(it, n, h0, w0);
        const uint32_t bar = smem_u32(&bar_full[s]);
        mbar_expect_tx(bar, TS_STAGE);
        tma_load_4d(base + OFF_X + (uint32_t)s * TS_STAGE, &P.xmap, bar, 0, w0 - 1, h0 - 1, n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ tensor-core issuer
    const uint32_t idesc = make_idesc_bf16(128, TS_N, 0, 0);
    const uint64_t hi64 = make_smem_desc_hi(512, LAYOUT_SW64) << 32;
    const uint32_t lbo16 = 1u << 16;
    const uint32_t b16 = ((base + OFF_W) >> 4) | lbo16;
    for (int it = 0; it < n_local; ++it) {
      const int s = it % TS_NST, use = it / TS_NST, g = it % TS_NG;
      mbar_wait(smem_u32(&bar_full[s]), (uint32_t)use & 1u);
      if (it >= TS_NG) mbar_wait(smem_u32(&bar_accfree[g]), (uint32_t)(it / TS_NG - 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a16 = ((base + OFF_X + (uint32_t)s * TS_STAGE) >> 4) | lbo16;
#pragma unroll
        for (int mb = 0; mb < TS_MB; ++mb)
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_f16(tmem + (uint32_t)(g * TS_MB + mb) * TS_N, hi64 | (uint64_t)(a16 + (uint32_t)mb * 512u + 2u * k), hi64 | (uint64_t)(b16 + 2u * k),
                     idesc, (uint32_t)k);
        umma_commit(smem_u32(&bar_empty[s]));
        umma_commit(smem_u32(&bar_acc[g]));
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue groups (tiles round-robin)
    const int g = (warp - 2) / TS_GW, wg = (warp - 2) % TS_GW;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read: halo rows q, q + 4, q + 8, q + 12 of the tile
    float* S = reinterpret_cast<float*>(gen + OFF_S + (uint32_t)g * TS_SCRATCH);
    float bias[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) bias[c] = bias_s[c];
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(TS_GW * 32) : "memory"); };
    // H: lane = halo column x; partial[ky][co](x) = sum_kx D[(row, x + kx), (ky, kx, co)] for output column x (lanes 30, 31: unused)
    auto horizontal = [&](const uint32_t (&v)[32], int hr) {
      float* dst = S + hr * TS_HW + lane;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
          const float a = __uint_as_float(v[(ky * 3 + 0) * COUT + co]);
          const float b = __shfl_down_sync(0xffffffffu, __uint_as_float(v[(ky * 3 + 1) * COUT + co]), 1);
          const float c = __shfl_down_sync(0xffffffffu, __uint_as_float(v[(ky * 3 + 2) * COUT + co]), 2);
          dst[(ky * COUT + co) * TS_PIX] = (a + b) + c;
        }
    };
    for (int it = g; it < n_local; it += TS_NG) {
      int n, h0, w0;
      tile_coords(it, n, h0, w0);
      mbar_wait(smem_u32(&bar_acc[g]), (uint32_t)(it / TS_NG) & 1u);
      tc_fence_after();
      {
        const uint32_t t0 = tmem + (uint32_t)(g * TS_MB) * TS_N + ((uint32_t)(q * 32) << 16);
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(t0, v0);
        tmem_ld_32x32(t0 + TS_N, v1);
        tmem_ld_wait_on(v0); tmem_ld_wait_on(v1);
        horizontal(v0, q);
        horizontal(v1, 4 + q);
        tmem_ld_32x32(t0 + 2 * TS_N, v0);
        tmem_ld_32x32(t0 + 3 * TS_N, v1);
        tmem_ld_wait_on(v0); tmem_ld_wait_on(v1);
        horizontal(v0, 8 + q);
        horizontal(v1, 12 + q);
      }
      tc_fence_before();
      group_sync();
      if (wg == 0 && lane == 0) mbar_arrive(smem_u32(&bar_accfree[g]));
      // V: this warp's output rows r = wg, wg + 4, ...; lane = output column
      const int w = w0 + lane;
      const bool col_ok = lane < TS_OW && w < P.ox + P.OW;
#pragma unroll 1
      for (int r = wg; r < TS_OH; r += TS_GW) {
        const int h = h0 + r;
        const float* sp = S + r * TS_HW + lane;
        float acc[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co)
          acc[co] = ((bias[co] + sp[co * TS_PIX]) + sp[(COUT + co) * TS_PIX + TS_HW]) + sp[(2 * COUT + co) * TS_PIX + 2 * TS_HW];
        if (col_ok && h < P.oy + P.OH) {
          if (P.frame) {
            uint8_t* qd = P.frame + (((long)n * P.OH + (h - P.oy)) * P.OW + (w - P.ox)) * 3;
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
              float f = __fmaf_rn(apply_act(acc[co], P.act, P.alpha), P.scale, P.offset);
              if (P.clip01) f = fminf(fmaxf(f, 0.f), 1.f);
              f = __fmul_rn(f, 255.0f);
              f = fminf(fmaxf(f, 0.f), 255.f);
              qd[P.flip ? 2 - co : co] = (uint8_t)(int)f;
            }
          } else {
            float* qd = P.y + (((long)n * P.H + h) * P.W + w) * P.yp;
#pragma unroll
            for (int co = 0; co < COUT; ++co) qd[co] = apply_act(acc[co], P.act, P.alpha);
          }
        }
      }
      group_sync();     // the scratch is rewritten by the group's next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TS_NG * TS_MB * TS_N);
  }
}

typedef CUresult (*TsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tapsum_launch(dg_ctx* ctx, const dg_tensor* x, const float* w, const float* bias, int cout, int act, float alpha, float* y, int yp,
                  uint8_t* frame, int dst_h, int dst_w, float scale, float offset, int clip01, int flip, void* stream, const char* who) {
  TsParams P;
  memset(&P, 0, sizeof(P));
  uint64_t dims[4] = {(uint64_t)TS_C, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
  uint64_t strides[3] = {(uint64_t)x->cpitch * 2, (uint64_t)x->cpitch * 2 * x->w, (uint64_t)x->cpitch * 2 * x->w * x->h};
  uint32_t box[4] = {(uint32_t)TS_C, (uint32_t)TS_HW, (uint32_t)TS_HH, 1}, ones[4] = {1, 1, 1, 1};
  CUresult r = ((TsEncodeFn)ctx->encode_tiled)(&P.xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (char*)x->ptr + (size_t)x->coff * 2,
                                               (const cuuint64_t*)dims, (const cuuint64_t*)strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DG_FAIL("%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
  P.w = w; P.bias = bias; P.y = y; P.frame = frame; P.yp = yp;
  P.N = x->n; P.H = x->h; P.W = x->w; P.act = act; P.alpha = alpha;
  if (frame) {
    P.OH = dst_h; P.OW = dst_w; P.oy = (x->h - dst_h) / 2; P.ox = (x->w - dst_w) / 2;     // centre crop (tf.image.resize_with_crop_or_pad)
  } else {
    P.OH = x->h; P.OW = x->w;
  }
  P.scale = scale; P.offset = offset; P.clip01 = clip01; P.flip = flip;
  P.tiles_w = (P.OW + TS_OW - 1) / TS_OW; P.tiles_h = (P.OH + TS_OH - 1) / TS_OH;
  const long total = (long)P.N * P.tiles_h * P.tiles_w;
  DG_REQUIRE(total < (1L << 30), "%s: too many tiles", who);
  P.total = (int)total;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tapsum_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tapsum_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tapsum_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM);
    if (e != cudaSuccess) DG_FAIL("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
    attr_set = true;
  }
  const unsigned grid = (unsigned)(P.total < ctx->sm_count ? P.total : ctx->sm_count);
  if (cout == 1) conv_tapsum_kernel<1><<<grid, TS_THREADS, TS_SMEM, (cudaStream_t)stream>>>(P);
  else if (cout == 2) conv_tapsum_kernel<2><<<grid, TS_THREADS, TS_SMEM, (cudaStream_t)stream>>>(P);
  else conv_tapsum_kernel<3><<<grid, TS_THREADS, TS_SMEM, (cudaStream_t)stream>>>(P);
  DG_CHECK_LAUNCH(who);
  return 0;
}
}  // namespace

extern "C" int dg_conv3x3_tapsum_supported(dg_ctx* ctx, const dg_tensor* x, int cout) {
  static const char* off = getenv("DG_CONV_TAPSUM");
  if (off && off[0] == '0') return 0;
  return ctx && ctx->encode_tiled && ctx->cc_major == 10 && dg_valid(x) && x->dtype == DG_BF16 && x->c == TS_C && cout >= 1 && cout <= 3 &&
         x->cpitch % 8 == 0 && x->coff % 8 == 0 && ((uintptr_t)x->ptr % 16) == 0;
}

extern "C" int dg_conv3x3_tapsum_fwd(dg_ctx* ctx, const dg_tensor* x, const float* w_hwio, const float* bias, int act, float alpha,
                                     const dg_tensor* y, void* stream) {
  DG_REQUIRE(ctx && w_hwio && dg_valid(y), "dg_conv3x3_tapsum_fwd: null argument");
  DG_REQUIRE(dg_conv3x3_tapsum_supported(ctx, x, y->c), "dg_conv3x3_tapsum_fwd: needs a 32-channel bf16 NHWC input (16-byte aligned pixels) and 1..3 output channels on sm_100");
  DG_REQUIRE(y->dtype == DG_F32 && y->n == x->n && y->h == x->h && y->w == x->w, "dg_conv3x3_tapsum_fwd: y must be the fp32 [n,h,w,cout] result");
  return tapsum_launch(ctx, x, w_hwio, bias, y->c, act, alpha, (float*)y->ptr + y->coff, y->cpitch, nullptr, 0, 0, 1.f, 0.f, 0, 0, stream,
                       "dg_conv3x3_tapsum_fwd");
}

extern "C" int dg_conv3x3_tapsum_frame(dg_ctx* ctx, const dg_tensor* x, const float* w_hwio, const float* bias, int act, float alpha,
                                       float scale, float offset, int clip01, int flip_channels, uint8_t* dst, int dst_h, int dst_w,
                                       void* stream) {
  DG_REQUIRE(ctx && w_hwio && dst, "dg_conv3x3_tapsum_frame: null argument");
  DG_REQUIRE(dg_conv3x3_tapsum_supported(ctx, x, 3), "dg_conv3x3_tapsum_frame: needs a 32-channel bf16 NHWC input (16-byte aligned pixels) on sm_100");
  DG_REQUIRE(dst_h > 0 && dst_w > 0 && dst_h <= x->h && dst_w <= x->w, "dg_conv3x3_tapsum_frame: the frame must be a centre crop of the %d x %d output (got %d x %d)",
             x->h, x->w, dst_h, dst_w);
  return tapsum_launch(ctx, x, w_hwio, bias, 3, act, alpha, nullptr, 0, dst, dst_h, dst_w, scale, offset, clip01, flip_channels, stream,
                       "dg_conv3x3_tapsum_frame");
}
