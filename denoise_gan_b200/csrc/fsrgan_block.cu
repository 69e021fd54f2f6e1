// Fast-SRGAN inverted-residual block at inference as ONE launch (fsrgan.py:112-176 with training=False, the model
// infer_video.py:92-97,146 runs): 1x1 expand 32 -> 192 (+ BatchNorm + ReLU), depthwise 3x3 (+ BatchNorm + ReLU), 1x1 project
// 192 -> 32 (+ BatchNorm), + block input.  The three BatchNorms are folded into the kernels and biases by the caller.
//
// As three launches the 192-channel intermediates of a 1080p frame (1280 x 2048 padded, 1 GB each) cross HBM four times
// (~4.3 GB per block, ~1.0 ms: profiles/infer_profile_r2*_fsrgan.log); here they never leave the SM: the block moves
// 168 MB in and 168 MB out.
//
// One persistent CTA per SM walks 16 x 8 output tiles:
//   control warp : TMA of the (8+2) x (16+2) x 32 halo tile (2-stage ring, out-of-image pixels zero-filled), and the
//                  tcgen05 products -- expand: ones . bias (the fp32 bias as two bf16 terms), then [256 halo rows (180 used)] x 32 .
//                  32 x 192 -> TMEM (2 x 192 columns); project: [128 output pixels] x 192 . 192 x 32 -> TMEM (2 x 32 columns,
//                  double-buffered);
//   12 compute warps, per tile (E and D below); four epilogue warps (the control warp is one of them) run P a tile behind:
//     E  accumulator -> ReLU, zero for out-of-image pixels (the depthwise convolution pads ITS input with zeros, not
//        relu(bias)) -> fp16 [180][192] in shared memory (pixel pitch 400 B: conflict-free 16-byte stores);
//     D  depthwise 3x3 on CUDA cores: a thread owns a channel pair and two output rows, rolling 4 x 3 window in registers,
//        packed half-precision FMAs (HFMA2, the ninth with ReLU) -> straight into the 128-byte-swizzled K-major A operand
//        of the project product;
//     P  (tile i-1, epilogue warps) accumulator + bias + block input (re-read from L2) -> bf16 -> global; its TMEM / L2 latencies
//        never stall the compute warps.
//   The expand product of tile i+1 runs under D(i), the project product of tile i under E(i+1): the tensor pipe is never
//   waited for.  (The sub-partition that hosts the control warp finishes D ~400 cycles after the other three -- the 18 tcgen05.mma
//   it issues per tile cost it issue slots; sleeping between barrier polls changed nothing: profiles/fsrgan_block_timeline_r2_warps.log.)
// Why fp16 inside: the kernel is bound by the CUDA-core depthwise stage and by shared-memory bandwidth (~370 KB per tile through a
// 128 B/clock port), not by HBM.  B200 retires 128 FMAs per clock and SM whether they are issued as FFMA, FFMA2 (fma.rn.f32x2) or
// HFMA2 (probes/fma_rate_probe.cu, profiles/fma_rate_probe_r2.log), so half precision buys no arithmetic rate; what it removes
// is every other instruction of the stage: with fp32 accumulation a step of the window walk is 4 loads, 8 unpack operations, 18
// FFMA2, 2 ReLU-converts and 2 stores, with HFMA2 it is 4 loads, 18 HFMA2 (the ninth carries the ReLU) and 2 stores, and E needs
// no add.  Measured (tools/fsrgan_block_timeline.py, profiles/fsrgan_block_timeline_r2_*.log): D 3790 -> 2730 cycles, E 2090 ->
// 1110, a tile 6875 -> 4750 cycles.  The two intermediates are stored as fp16 (11 mantissa bits against bf16's 8 in the
// three-launch path; conversions saturate to +-65504, activations here are O(1)) -- the storage type of the reference's own
// 'mixed_float16' policy (train_fsrgan.py:314, --fp16) -- and the nine-tap sum is accumulated in fp16: its rounding (<= 9 x 2^-12
// of the partial sum) is of the order of the bf16 rounding of the three-launch path's stored result.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "dg_common.cuh"
#include "sm100.cuh"

namespace {
using namespace sm100;

constexpr int FB_C = 32, FB_E = 192;                  // block channels, expanded channels (expansion 6, fsrgan.py:121)
constexpr int FB_TW = 16, FB_TH = 8;                  // output tile
constexpr int FB_IW = FB_TW + 2, FB_IH = FB_TH + 2;   // halo tile
constexpr int FB_HALO = FB_IW * FB_IH;                // 180 rows of the expand product
constexpr int FB_COMPUTE_WARPS = 12, FB_EPI_WARPS = 4, FB_THREADS = (FB_COMPUTE_WARPS + FB_EPI_WARPS) * 32;   // warp 12: control + epilogue
constexpr uint32_t FB_PITCH = 400;                    // bytes per pixel of the expanded tile (384 + 16: bank spread)
constexpr uint32_t OFF_W1 = 0;                        // expand B operand: 192 rows x 64 B (bf16), 64-byte swizzle
constexpr uint32_t OFF_WB = 12288;                    // expand bias as a B operand: row n = {hi(b[n]), lo(b[n]), 0 ...} (bf16), same layout
constexpr uint32_t OFF_W2 = 24576;                    // project B operand: 3 K blocks x (32 rows x 128 B) (fp16), 128-byte swizzle
constexpr uint32_t OFF_ONES = 36864;                  // A operand of the bias product: 128 rows x 64 B, {1, 1, 0 ...} (bf16)
constexpr uint32_t OFF_A = 45056;                     // project A operand: 3 K blocks x (128 rows x 128 B) (fp16)
constexpr uint32_t OFF_X = OFF_A + 3 * 16384;         // halo stages: 2 x 16 KB (256 rows x 64 B addressed by the product, 180 written)
constexpr uint32_t OFF_I = OFF_X + 2 * 16384;         // expanded tile (fp16)
constexpr uint32_t OFF_B = OFF_I + ((FB_HALO * FB_PITCH + 127u) & ~127u);   // project bias [32] (fp32)
constexpr uint32_t FB_SMEM = OFF_B + FB_C * 4 + 1024;
constexpr uint32_t X_BYTES = FB_HALO * FB_C * 2;
constexpr uint32_t TM_E = 0, TM_P = 2 * FB_E;         // TMEM columns: expand accumulators (2 x 192), project accumulators (2 x 32)

struct FbParams {
  CUtensorMap xmap;
  const __nv_bfloat16* x;       // block input (residual), pixel pitch xp, channel offset already applied
  const __nv_bfloat16* w1;      // [192][32]  expand kernel, BatchNorm folded, K-major
  const __half* w2;             // [32][192]  project kernel, BatchNorm folded, K-major, fp16
  const float* b1;              // [192]
  const float* wd;              // [9][192]   depthwise kernel, BatchNorm folded
  const float* bd;              // [192]
  const float* b2;              // [32]
  __nv_bfloat16* y;
  int xp, yp, N, H, W, tiles_w, tiles_h, total;
  long long* dbg;               // clock64 marks of CTA 0 ([16 tiles][40]; tools/fsrgan_block_timeline.py), or null
};

__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t hfma2_relu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t relu_pack_h2(uint32_t lo, uint32_t hi) {     // fp32 bits -> max(v, 0), saturated to the fp16 range -> f16x2
  uint32_t d;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return d;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void fb_mark(const FbParams& P, int it, int slot) {
  if (P.dbg && blockIdx.x == 0 && it < 16) P.dbg[it * 40 + slot] = clock64();
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(FB_COMPUTE_WARPS * 32) : "memory"); }

__global__ void __launch_bounds__(FB_THREADS, 1) fsrgan_block_kernel(const __grid_constant__ FbParams P) {
  extern __shared__ uint8_t fb_raw[];
  __shared__ __align__(8) uint64_t bar_x[2], bar_e, bar_efree, bar_a, bar_p[2], bar_pfree[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(fb_raw) + 1023u) & ~1023u;
  uint8_t* gen = fb_raw + (base - smem_u32(fb_raw));
  const int n_local = ((int)blockIdx.x < P.total) ? (P.total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (tid == 0) {
    mbar_init(smem_u32(&bar_x[0]), 1); mbar_init(smem_u32(&bar_x[1]), 1);
    mbar_init(smem_u32(&bar_e), 1);
    mbar_init(smem_u32(&bar_efree), FB_COMPUTE_WARPS);
    mbar_init(smem_u32(&bar_a), FB_COMPUTE_WARPS);
    mbar_init(smem_u32(&bar_p[0]), 1); mbar_init(smem_u32(&bar_p[1]), 1);
    mbar_init(smem_u32(&bar_pfree[0]), FB_EPI_WARPS); mbar_init(smem_u32(&bar_pfree[1]), FB_EPI_WARPS);
    fence_mbar_init();
    tma_prefetch_desc(&P.xmap);
  }
  if (warp == FB_COMPUTE_WARPS) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  // ---- the two B operands (24 KB) and the biases, once per CTA: 16-byte chunks to their swizzled places
  for (int i = tid; i < FB_E * 4; i += FB_THREADS) {          // expand: row n (0..191), chunk c (0..3) of 64 B
    const int n = i >> 2, c = i & 3;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.w1) + i);
    *reinterpret_cast<uint4*>(gen + OFF_W1 + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)) = v;
  }
  for (int i = tid; i < FB_C * 24; i += FB_THREADS) {         // project: row n (0..31), 24 chunks of the 384-byte K row
    const int n = i / 24, c = i - n * 24, kb = c >> 3, cc = c & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.w2) + i);
    *reinterpret_cast<uint4*>(gen + OFF_W2 + kb * 4096 + n * 128 + ((cc ^ (n & 7)) << 4)) = v;
  }
  // the expand bias rides on the tensor pipe: accumulator = ones[128 x 16] . biasB[16 x 192] before the x . W1 product, the fp32
  // bias split into two bf16 terms (hi + lo: 16 mantissa bits) in K elements 0 and 1; E then has no per-element add
  for (int i = tid; i < FB_E * 4; i += FB_THREADS) {
    const int n = i >> 2, c = i & 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c == 0) {
      const float b = P.b1[n];
      const __nv_bfloat16 hi = __float2bfloat16(b), lo = __float2bfloat16(b - __bfloat162float(hi));
      v.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
    }
    *reinterpret_cast<uint4*>(gen + OFF_WB + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)) = v;
  }
  for (int i = tid; i < 128 * 4; i += FB_THREADS) {
    const int r = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(gen + OFF_ONES + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = make_uint4(c == 0 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
  }
  float* bias_s = reinterpret_cast<float*>(gen + OFF_B);
  for (int i = tid; i < FB_C; i += FB_THREADS) bias_s[i] = P.b2[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  auto tile_coords = [&](int it, int& n, int& h0, int& w0) {
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    const int tw = tile % P.tiles_w, t2 = tile / P.tiles_w, th = t2 % P.tiles_h;
    n = t2 / P.tiles_h; h0 = th * FB_TH; w0 = tw * FB_TW;
  };

  // P(it): one output pixel per thread of the four epilogue warps (TMEM lane quarter q = warp % 4)
  auto project_out = [&](int it, int q) {
    int n, h0, w0;
    tile_coords(it, n, h0, w0);
    uint32_t v[32];
    tmem_ld_32x32(tmem + TM_P + (uint32_t)(it & 1) * FB_C + ((uint32_t)(q * 32) << 16), v);
    const int r = q * 32 + lane, h = h0 + (r >> 4), w = w0 + (r & 15);
    const bool ok = h < P.H && w < P.W;
    const long pix = ((long)n * P.H + h) * P.W + w;
    uint4 res[4];
    if (ok) {
#pragma unroll
      for (int k = 0; k < 4; ++k) res[k] = __ldg(reinterpret_cast<const uint4*>(P.x + pix * P.xp) + k);
    }
    tmem_ld_wait();
    if (ok) {
      const float* b2 = reinterpret_cast<const float*>(gen + OFF_B);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t rr[4] = {res[k].x, res[k].y, res[k].z, res[k].w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = k * 8 + e * 2;
          const float lo = __uint_as_float(v[c]) + b2[c] + __uint_as_float(rr[e] << 16);
          const float hi = __uint_as_float(v[c + 1]) + b2[c + 1] + __uint_as_float(rr[e] & 0xffff0000u);
          __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
          o[e] = *reinterpret_cast<uint32_t*>(&pk);
        }
        reinterpret_cast<uint4*>(P.y + pix * P.yp)[k] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  };
  auto epilogue_tile = [&](int it, int q) {
    mbar_wait(smem_u32(&bar_p[it & 1]), ((uint32_t)it >> 1) & 1u);
    tc_fence_after();
    project_out(it, q);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_pfree[it & 1]));
  };

  if (warp == FB_COMPUTE_WARPS) {
    // ------------------------------------------------------------------ control warp (whole warp, tcgen05 / TMA under elect_one);
    // between its two waits per tile it is also the epilogue warp of TMEM lane quarter 0
    const uint32_t idesc_e = make_idesc_bf16(128, FB_E, 0, 0);
    const uint32_t idesc_p = (1u << 4) | ((uint32_t)(FB_C >> 3) << 17) | ((128u >> 4) << 24);   // fp16 A and B (format 0), fp32 accumulate
    const uint64_t hi64 = make_smem_desc_hi(512, LAYOUT_SW64) << 32, hi128 = make_smem_desc_hi(1024, LAYOUT_SW128) << 32;
    const uint32_t lbo16 = 1u << 16;
    auto load_x = [&](int it) {
      int n, h0, w0;
      tile_coords(it, n, h0, w0);
      const uint32_t bar = smem_u32(&bar_x[it & 1]);
      mbar_expect_tx(bar, X_BYTES);
      tma_load_4d(base + OFF_X + (uint32_t)(it & 1) * 16384u, &P.xmap, bar, 0, w0 - 1, h0 - 1, n);
    };
    auto mma_expand = [&](int it) {
      const uint32_t a16 = ((base + OFF_X + (uint32_t)(it & 1) * 16384u) >> 4) | lbo16, b16 = ((base + OFF_W1) >> 4) | lbo16;
      const uint32_t o16 = ((base + OFF_ONES) >> 4) | lbo16, wb16 = ((base + OFF_WB) >> 4) | lbo16;
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        umma_f16(tmem + TM_E + (uint32_t)mb * FB_E, hi64 | (uint64_t)o16, hi64 | (uint64_t)wb16, idesc_e, 0u);      // bias
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_f16(tmem + TM_E + (uint32_t)mb * FB_E, hi64 | (uint64_t)(a16 + (uint32_t)mb * 512u + 2u * k), hi64 | (uint64_t)(b16 + 2u * k), idesc_e, 1u);
      }
      umma_commit(smem_u32(&bar_e));
    };
    auto mma_project = [&](int it) {
      const uint32_t a16 = ((base + OFF_A) >> 4) | lbo16, b16 = ((base + OFF_W2) >> 4) | lbo16;
#pragma unroll
      for (int kb = 0; kb < 3; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem + TM_P + (uint32_t)(it & 1) * FB_C, hi128 | (uint64_t)(a16 + (uint32_t)kb * 1024u + 2u * k),
                   hi128 | (uint64_t)(b16 + (uint32_t)kb * 256u + 2u * k), idesc_p, (kb | k) != 0);
      umma_commit(smem_u32(&bar_p[it & 1]));
    };
    if (n_local > 0) {
      if (elect_one()) {
        load_x(0);
        if (n_local > 1) load_x(1);
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar_x[0]), 0);
      tc_fence_after();
      if (elect_one()) mma_expand(0);
      __syncwarp();
    }
    for (int it = 0; it < n_local; ++it) {
      mbar_wait(smem_u32(&bar_efree), (uint32_t)it & 1u);        // E(it) has read the expand accumulators
      tc_fence_after();
      if (lane == 0) fb_mark(P, it, 4);
      if (it + 1 < n_local) {
        mbar_wait(smem_u32(&bar_x[(it + 1) & 1]), ((uint32_t)(it + 1) >> 1) & 1u);
        tc_fence_after();
        if (elect_one()) mma_expand(it + 1);
        __syncwarp();
      }
      if (it + 2 < n_local && elect_one()) load_x(it + 2);       // its stage was read by the expand product of tile it: complete
      __syncwarp();
      if (it > 0) epilogue_tile(it - 1, 0);                      // well inside D(it): the A operand of tile it is not ready before
      if (lane == 0) fb_mark(P, it, 5);
      mbar_wait(smem_u32(&bar_a), (uint32_t)it & 1u);            // D(it) has written the A operand
      if (lane == 0) fb_mark(P, it, 6);
      if (it >= 2) mbar_wait(smem_u32(&bar_pfree[it & 1]), ((uint32_t)(it - 2) >> 1) & 1u);   // P(it-2) has read this accumulator
      tc_fence_after();
      if (elect_one()) mma_project(it);
      __syncwarp();
    }
    if (n_local > 0) epilogue_tile(n_local - 1, 0);
  } else if (warp > FB_COMPUTE_WARPS) {
    // ------------------------------------------------------------------ epilogue warps of lane quarters 1..3
    for (int it = 0; it < n_local; ++it) epilogue_tile(it, warp & 3);
  } else {
    // ------------------------------------------------------------------ compute warps
    const int q = warp & 3, j3 = warp >> 2;        // E: TMEM lane quarter, 64-column third
    const int g = warp % 3, rg = warp / 3;         // D: 64-channel group, output rows 2rg, 2rg+1
    const int cpair = g * 32 + lane;               // channel pair (channels 2cpair, 2cpair+1)
    uint32_t wk[9], bdw;                           // f16x2
#pragma unroll
    for (int t = 0; t < 9; ++t) wk[t] = pack_h2(P.wd[t * FB_E + 2 * cpair], P.wd[t * FB_E + 2 * cpair + 1]);
    bdw = pack_h2(P.bd[2 * cpair], P.bd[2 * cpair + 1]);
    const uint32_t inter = base + OFF_I, abuf = base + OFF_A;
    const int n_units = q * 32 < FB_HALO - 128 ? 4 : 2;            // 32-column batches of this warp: rows 180..255 are not part of the tile

    // E reads the accumulator in 32-column batches u = (M block u/2, columns u%2); the first two are requested BEFORE the barrier
    // that ends the previous tile, so their TMEM latency runs under the barrier skew, and batch u+2 is in flight while u is converted
    uint32_t v[2][32];
    auto ld_e = [&](int u, uint32_t (&dst)[32]) {
      tmem_ld_32x32(tmem + TM_E + (uint32_t)(u >> 1) * FB_E + (uint32_t)(j3 * 64 + (u & 1) * 32) + ((uint32_t)(q * 32) << 16), dst);
    };
    auto request_e = [&](int it) {
      mbar_wait(smem_u32(&bar_e), (uint32_t)it & 1u);
      tc_fence_after();
      ld_e(0, v[0]);
      ld_e(1, v[1]);
    };
    if (n_local > 0) request_e(0);
    for (int it = 0; it < n_local; ++it) {
      int n, h0, w0;
      tile_coords(it, n, h0, w0);
      // ---- E(it): expand accumulators (bias already in them) -> ReLU -> fp16 expanded tile
      if (tid == 0) fb_mark(P, it, 0);
      {
        auto conv = [&](int u, const uint32_t (&src)[32]) {
          const int row = (u >> 1) * 128 + q * 32 + lane;
          const int hh = row / FB_IW, ww = row - hh * FB_IW;
          const int h = h0 - 1 + hh, w = w0 - 1 + ww;
          if (row >= FB_HALO) return;
          const uint32_t dst = inter + (uint32_t)row * FB_PITCH + (uint32_t)j3 * 128u + (uint32_t)(u & 1) * 64u;
          if (h >= 0 && h < P.H && w >= 0 && w < P.W) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = relu_pack_h2(src[k * 8 + e * 2], src[k * 8 + e * 2 + 1]);
              sts128(dst + (uint32_t)k * 16u, make_uint4(o[0], o[1], o[2], o[3]));
            }
          } else {
            // outside the image the depthwise convolution sees ZERO padding, not relu(bias)
#pragma unroll
            for (int k = 0; k < 4; ++k) sts128(dst + (uint32_t)k * 16u, make_uint4(0u, 0u, 0u, 0u));
          }
        };
        tmem_ld_wait();                 // batches 0 and 1 were requested before the end-of-tile barrier
        conv(0, v[0]);
        if (n_units == 4) ld_e(2, v[0]);
        conv(1, v[1]);
        if (n_units == 4) {
          ld_e(3, v[1]);
          tmem_ld_wait();
          conv(2, v[0]);
          conv(3, v[1]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_efree));
      if (tid == 0) fb_mark(P, it, 1);
      compute_sync();                                              // the expanded tile is complete
      if (it > 0) mbar_wait(smem_u32(&bar_p[(it - 1) & 1]), ((uint32_t)(it - 1) >> 1) & 1u);   // project product of tile it-1 has read the A operand
      // ---- D(it): depthwise 3x3 + bias + ReLU -> A operand of the project product
      if (tid == 0) fb_mark(P, it, 2);
      if (lane == 0) fb_mark(P, it, 8 + warp);
      {
        const uint32_t src = inter + (uint32_t)(2 * rg * FB_IW) * FB_PITCH + (uint32_t)cpair * 4u;
        uint32_t win[4][4];                          // f16x2, straight from shared memory; column c+3 is loaded one step ahead
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          win[r][0] = lds32(src + (uint32_t)(r * FB_IW + 0) * FB_PITCH);
          win[r][1] = lds32(src + (uint32_t)(r * FB_IW + 1) * FB_PITCH);
          win[r][2] = lds32(src + (uint32_t)(r * FB_IW + 2) * FB_PITCH);
        }
        const uint32_t arow0 = abuf + (uint32_t)g * 16384u + (uint32_t)(2 * rg * FB_TW) * 128u + (uint32_t)(lane & 3) * 4u;
#pragma unroll
        for (int c = 0; c < FB_TW; ++c) {
          if (c + 3 < FB_IW) {
#pragma unroll
            for (int r = 0; r < 4; ++r) win[r][(c + 3) % 4] = lds32(src + (uint32_t)(r * FB_IW + c + 3) * FB_PITCH);
          }
          uint32_t a0 = bdw, a1 = bdw;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              if (a * 3 + b < 8) {
                a0 = hfma2(win[a][(c + b) % 4], wk[a * 3 + b], a0);
                a1 = hfma2(win[a + 1][(c + b) % 4], wk[a * 3 + b], a1);
              } else {                                  // the ninth tap carries the ReLU
                a0 = hfma2_relu(win[a][(c + b) % 4], wk[a * 3 + b], a0);
                a1 = hfma2_relu(win[a + 1][(c + b) % 4], wk[a * 3 + b], a1);
              }
            }
          // output pixel rows 2rg*16 + c and (2rg+1)*16 + c of the 128-row A operand; 16-byte chunk lane/4, XOR row%8
          const uint32_t r0 = (uint32_t)(2 * rg * FB_TW + c);
          sts32(arow0 + (uint32_t)c * 128u + ((((uint32_t)lane >> 2) ^ (r0 & 7u)) << 4), a0);
          sts32(arow0 + (uint32_t)(c + FB_TW) * 128u + ((((uint32_t)lane >> 2) ^ ((r0 + FB_TW) & 7u)) << 4), a1);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_a));
      if (tid == 0) fb_mark(P, it, 3);
      if (lane == 0) fb_mark(P, it, 24 + warp);
      if (it + 1 < n_local) request_e(it + 1);                     // expand product of tile it+1: issued under D(it), complete long ago
      compute_sync();                                              // everyone has read the expanded tile
    }
  }
  __syncthreads();
  if (warp == FB_COMPUTE_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------- warp-specialised variant: E of tile i+1 under D of tile i
// In the kernel above the twelve compute warps run E (accumulator -> shared memory: TMEM-read / latency bound) and D (depthwise
// stage: FMA-pipe bound) one after the other, ~1100 + ~2750 of a tile's ~4800 cycles, and no pipe is busy half the time.  Here
// the two phases belong to different warps and overlap:
//   E warps 12..15 (one per TMEM lane quarter): expand accumulator (double-buffered, 2 x 192 columns) -> fp16 expanded tile,
//   D warps 0..11: depthwise stage from the expanded tile (double-buffered in shared memory) into the project A operand (double-buffered:
//                  with one copy D(i) waited for the project product of tile i-1 -- 2530 instead of ~1500 cycles per tile),
//   P warps 16, 17: project accumulator + bias + block input -> global,
//   control warps 18 (TMA, expand products) and 19 (project products).
// The expanded tile must fit shared memory twice, so the tile is 16 x 4 output pixels (halo 18 x 6 = 108 rows: ONE 128-row expand
// product, every lane quarter carries its share; the project product runs with 64 of its 128 rows in use).  All hand-overs are
// mbarriers; there is no CTA-wide or group barrier inside the tile loop.
constexpr int W2_TW = 16, W2_TH = 4, W2_IW = W2_TW + 2, W2_IH = W2_TH + 2, W2_HALO = W2_IW * W2_IH;   // 108
constexpr int W2_D_WARPS = 12, W2_E_WARPS = 4, W2_P_WARPS = 2, W2_THREADS = (W2_D_WARPS + W2_E_WARPS + W2_P_WARPS + 2) * 32;
constexpr uint32_t W2_OFF_W1 = 0, W2_OFF_WB = 12288, W2_OFF_W2 = 24576, W2_OFF_ONES = 36864;
constexpr uint32_t W2_OFF_A = 45056, W2_A_BYTES = 3 * 8192;   // project A operand, twice: 3 K blocks x (64 rows x 128 B) at 8 KB; the product
                                                              // addresses 128 rows per K block (the upper 64 are somebody else's bytes): + 8 KB slack
constexpr uint32_t W2_OFF_X = W2_OFF_A + 2 * W2_A_BYTES + 8192;   // 3 halo stages x 8 KB (128 rows x 64 B addressed, 108 written)
constexpr uint32_t W2_I_BYTES = (W2_HALO * FB_PITCH + 127u) & ~127u;
constexpr uint32_t W2_OFF_I = W2_OFF_X + 3 * 8192;         // expanded tile, twice
constexpr uint32_t W2_OFF_B = W2_OFF_I + 2 * W2_I_BYTES;   // project bias [32] (fp32)
constexpr uint32_t W2_SMEM = W2_OFF_B + FB_C * 4 + 1024;
constexpr uint32_t W2_X_BYTES = W2_HALO * FB_C * 2;

__global__ void __launch_bounds__(W2_THREADS, 1) fsrgan_block_ws_kernel(const __grid_constant__ FbParams P) {
  extern __shared__ uint8_t fb_raw[];
  __shared__ __align__(8) uint64_t bar_x[3], bar_e[2], bar_efree[2], bar_ifull[2], bar_ifree[2], bar_a, bar_p[2], bar_pfree[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(fb_raw) + 1023u) & ~1023u;
  uint8_t* gen = fb_raw + (base - smem_u32(fb_raw));
  const int n_local = ((int)blockIdx.x < P.total) ? (P.total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  constexpr int WARP_E0 = W2_D_WARPS, WARP_P0 = WARP_E0 + W2_E_WARPS, WARP_CX = WARP_P0 + W2_P_WARPS, WARP_CP = WARP_CX + 1;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bar_x[i]), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_e[i]), 1);
      mbar_init(smem_u32(&bar_efree[i]), W2_E_WARPS);
      mbar_init(smem_u32(&bar_ifull[i]), W2_E_WARPS);
      mbar_init(smem_u32(&bar_ifree[i]), W2_D_WARPS);
      mbar_init(smem_u32(&bar_p[i]), 1);
      mbar_init(smem_u32(&bar_pfree[i]), W2_P_WARPS);
    }
    mbar_init(smem_u32(&bar_a), W2_D_WARPS);
    fence_mbar_init();
    tma_prefetch_desc(&P.xmap);
  }
  if (warp == WARP_CX) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  for (int i = tid; i < FB_E * 4; i += W2_THREADS) {          // expand kernel and its bias (hi + lo bf16 terms), 64-byte swizzle
    const int n = i >> 2, c = i & 3;
    const uint32_t sw = (uint32_t)(n * 64 + ((c ^ ((n >> 1) & 3)) << 4));
    *reinterpret_cast<uint4*>(gen + W2_OFF_W1 + sw) = __ldg(reinterpret_cast<const uint4*>(P.w1) + i);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c == 0) {
      const float b = P.b1[n];
      const __nv_bfloat16 hi = __float2bfloat16(b), lo = __float2bfloat16(b - __bfloat162float(hi));
      v.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
    }
    *reinterpret_cast<uint4*>(gen + W2_OFF_WB + sw) = v;
  }
  for (int i = tid; i < FB_C * 24; i += W2_THREADS) {         // project kernel, 128-byte swizzle
    const int n = i / 24, c = i - n * 24, kb = c >> 3, cc = c & 7;
    *reinterpret_cast<uint4*>(gen + W2_OFF_W2 + kb * 4096 + n * 128 + ((cc ^ (n & 7)) << 4)) = __ldg(reinterpret_cast<const uint4*>(P.w2) + i);
  }
  for (int i = tid; i < 128 * 4; i += W2_THREADS) {
    const int r = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(gen + W2_OFF_ONES + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = make_uint4(c == 0 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
  }
  float* bias_s = reinterpret_cast<float*>(gen + W2_OFF_B);
  for (int i = tid; i < FB_C; i += W2_THREADS) bias_s[i] = P.b2[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  auto tile_coords = [&](int it, int& n, int& h0, int& w0) {
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    const int tw = tile % P.tiles_w, t2 = tile / P.tiles_w, th = t2 % P.tiles_h;
    n = t2 / P.tiles_h; h0 = th * W2_TH; w0 = tw * W2_TW;
  };
  const uint32_t lbo16 = 1u << 16;

  if (warp == WARP_CX) {
    // ------------------------------------------------------------------ control: halo loads and expand products
    const uint32_t idesc_e = make_idesc_bf16(128, FB_E, 0, 0);
    const uint64_t hi64 = make_smem_desc_hi(512, LAYOUT_SW64) << 32;
    auto load_x = [&](int it) {
      int n, h0, w0;
      tile_coords(it, n, h0, w0);
      const uint32_t bar = smem_u32(&bar_x[it % 3]);
      mbar_expect_tx(bar, W2_X_BYTES);
      tma_load_4d(base + W2_OFF_X + (uint32_t)(it % 3) * 8192u, &P.xmap, bar, 0, w0 - 1, h0 - 1, n);
    };
    if (elect_one())
      for (int it = 0; it < 3 && it < n_local; ++it) load_x(it);
    __syncwarp();
    for (int it = 0; it < n_local; ++it) {
      mbar_wait(smem_u32(&bar_x[it % 3]), ((uint32_t)(it / 3)) & 1u);
      if (lane == 0) fb_mark(P, it, 24);
      if (it >= 2) mbar_wait(smem_u32(&bar_efree[it & 1]), ((uint32_t)(it - 2) >> 1) & 1u);   // E(it-2) has read this accumulator
      tc_fence_after();
      if (lane == 0) fb_mark(P, it, 25);
      if (elect_one()) {
        const uint32_t a16 = ((base + W2_OFF_X + (uint32_t)(it % 3) * 8192u) >> 4) | lbo16, b16 = ((base + W2_OFF_W1) >> 4) | lbo16;
        const uint32_t o16 = ((base + W2_OFF_ONES) >> 4) | lbo16, wb16 = ((base + W2_OFF_WB) >> 4) | lbo16;
        const uint32_t acc = tmem + TM_E + (uint32_t)(it & 1) * FB_E;
        umma_f16(acc, hi64 | (uint64_t)o16, hi64 | (uint64_t)wb16, idesc_e, 0u);      // bias
        umma_f16(acc, hi64 | (uint64_t)a16, hi64 | (uint64_t)b16, idesc_e, 1u);
        umma_f16(acc, hi64 | (uint64_t)(a16 + 2u), hi64 | (uint64_t)(b16 + 2u), idesc_e, 1u);
        umma_commit(smem_u32(&bar_e[it & 1]));
      }
      __syncwarp();
      if (it >= 1 && it + 2 < n_local) {
        // stage (it+2) % 3 was read by the expand product of tile it-1 (issued one iteration ago: complete in steady state)
        mbar_wait(smem_u32(&bar_e[(it - 1) & 1]), ((uint32_t)(it - 1) >> 1) & 1u);
        if (elect_one()) load_x(it + 2);
        __syncwarp();
      }
    }
  } else if (warp == WARP_CP) {
    // ------------------------------------------------------------------ control: project products
    const uint32_t idesc_p = (1u << 4) | ((uint32_t)(FB_C >> 3) << 17) | ((128u >> 4) << 24);   // fp16 A and B, fp32 accumulate
    const uint64_t hi128 = make_smem_desc_hi(1024, LAYOUT_SW128) << 32;
    for (int it = 0; it < n_local; ++it) {
      mbar_wait(smem_u32(&bar_a), (uint32_t)it & 1u);                                            // D(it) has written the A operand
      if (lane == 0) fb_mark(P, it, 28);
      if (it >= 2) mbar_wait(smem_u32(&bar_pfree[it & 1]), ((uint32_t)(it - 2) >> 1) & 1u);      // P(it-2) has read this accumulator
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a16 = ((base + W2_OFF_A + (uint32_t)(it & 1) * W2_A_BYTES) >> 4) | lbo16, b16 = ((base + W2_OFF_W2) >> 4) | lbo16;
#pragma unroll
        for (int kb = 0; kb < 3; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem + TM_P + (uint32_t)(it & 1) * FB_C, hi128 | (uint64_t)(a16 + (uint32_t)kb * 512u + 2u * k),
                     hi128 | (uint64_t)(b16 + (uint32_t)kb * 256u + 2u * k), idesc_p, (kb | k) != 0);
        umma_commit(smem_u32(&bar_p[it & 1]));
      }
      __syncwarp();
    }
  } else if (warp >= WARP_P0) {
    // ------------------------------------------------------------------ P warps: rows 0..63 of the project accumulator
    const int q = warp & 3;      // warps 16, 17 -> lane quarters 0, 1
    const float* b2 = reinterpret_cast<const float*>(gen + W2_OFF_B);
    for (int it = 0; it < n_local; ++it) {
      int n, h0, w0;
      tile_coords(it, n, h0, w0);
      const int r = q * 32 + lane, h = h0 + (r >> 4), w = w0 + (r & 15);
      const bool ok = h < P.H && w < P.W;
      const long pix = ((long)n * P.H + h) * P.W + w;
      uint4 res[4];
      if (ok) {
#pragma unroll
        for (int k = 0; k < 4; ++k) res[k] = __ldg(reinterpret_cast<const uint4*>(P.x + pix * P.xp) + k);
      }
      if (lane == 0 && q == 0) fb_mark(P, it, 16);
      mbar_wait(smem_u32(&bar_p[it & 1]), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      if (lane == 0 && q == 0) fb_mark(P, it, 17);
      uint32_t v[32];
      tmem_ld_32x32(tmem + TM_P + (uint32_t)(it & 1) * FB_C + ((uint32_t)(q * 32) << 16), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_pfree[it & 1]));
      if (ok) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t rr[4] = {res[k].x, res[k].y, res[k].z, res[k].w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = k * 8 + e * 2;
            const float lo = __uint_as_float(v[c]) + b2[c] + __uint_as_float(rr[e] << 16);
            const float hi = __uint_as_float(v[c + 1]) + b2[c + 1] + __uint_as_float(rr[e] & 0xffff0000u);
            __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
            o[e] = *reinterpret_cast<uint32_t*>(&pk);
          }
          reinterpret_cast<uint4*>(P.y + pix * P.yp)[k] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  } else if (warp >= WARP_E0) {
    // ------------------------------------------------------------------ E warps: expand accumulator -> ReLU -> fp16 expanded tile
    const int q = warp & 3;      // warps 12..15 -> lane quarters 0..3
    const int row = q * 32 + lane, hh = row / W2_IW, ww = row - hh * W2_IW;
    for (int it = 0; it < n_local; ++it) {
      int n, h0, w0;
      tile_coords(it, n, h0, w0);
      const int h = h0 - 1 + hh, w = w0 - 1 + ww;
      const bool in_img = h >= 0 && h < P.H && w >= 0 && w < P.W;
      const uint32_t dst = base + W2_OFF_I + (uint32_t)(it & 1) * W2_I_BYTES + (uint32_t)row * FB_PITCH;
      const uint32_t acc = tmem + TM_E + (uint32_t)(it & 1) * FB_E + ((uint32_t)(q * 32) << 16);
      if (lane == 0 && q == 0) fb_mark(P, it, 0);
      mbar_wait(smem_u32(&bar_e[it & 1]), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      if (lane == 0 && q == 0) fb_mark(P, it, 1);
      uint32_t v[2][32];
      tmem_ld_32x32(acc, v[0]);
      tmem_ld_32x32(acc + 32u, v[1]);
      if (it >= 2) mbar_wait(smem_u32(&bar_ifree[it & 1]), ((uint32_t)(it - 2) >> 1) & 1u);   // D(it-2) has read this copy of the tile
      if (lane == 0 && q == 0) fb_mark(P, it, 2);
      auto conv = [&](int u, const uint32_t (&src)[32]) {
        if (row >= W2_HALO) return;
        const uint32_t d = dst + (uint32_t)u * 64u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = in_img ? relu_pack_h2(src[k * 8 + e * 2], src[k * 8 + e * 2 + 1]) : 0u;   // zero padding outside the image
          sts128(d + (uint32_t)k * 16u, make_uint4(o[0], o[1], o[2], o[3]));
        }
      };
#pragma unroll
      for (int u = 0; u < 6; u += 2) {
        tmem_ld_wait();
        conv(u, v[0]);
        if (u + 2 < 6) tmem_ld_32x32(acc + (uint32_t)(u + 2) * 32u, v[0]);
        conv(u + 1, v[1]);
        if (u + 3 < 6) tmem_ld_32x32(acc + (uint32_t)(u + 3) * 32u, v[1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && q == 0) fb_mark(P, it, 3);
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar_efree[it & 1]));
        mbar_arrive(smem_u32(&bar_ifull[it & 1]));
      }
    }
  } else {
    // ------------------------------------------------------------------ D warps: depthwise 3x3 + bias + ReLU -> project A operand
    const int g = warp >> 2, sub = warp & 3, rp = sub & 1, chf = sub >> 1;   // 64-channel group; output rows 2rp, 2rp+1; columns 8chf..8chf+7
    const int cpair = g * 32 + lane;
    uint32_t wk[9], bdw;
#pragma unroll
    for (int t = 0; t < 9; ++t) wk[t] = pack_h2(P.wd[t * FB_E + 2 * cpair], P.wd[t * FB_E + 2 * cpair + 1]);
    bdw = pack_h2(P.bd[2 * cpair], P.bd[2 * cpair + 1]);
    const uint32_t src_off = (uint32_t)(2 * rp * W2_IW + 8 * chf) * FB_PITCH + (uint32_t)cpair * 4u;
    const uint32_t arow00 = base + W2_OFF_A + (uint32_t)g * 8192u + (uint32_t)(2 * rp * W2_TW + 8 * chf) * 128u + (uint32_t)(lane & 3) * 4u;
    for (int it = 0; it < n_local; ++it) {
      const uint32_t src = base + W2_OFF_I + (uint32_t)(it & 1) * W2_I_BYTES + src_off;
      const uint32_t arow0 = arow00 + (uint32_t)(it & 1) * W2_A_BYTES;
      if (tid == 0) fb_mark(P, it, 8);
      mbar_wait(smem_u32(&bar_ifull[it & 1]), ((uint32_t)it >> 1) & 1u);
      if (tid == 0) fb_mark(P, it, 9);
      uint32_t win[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        win[r][0] = lds32(src + (uint32_t)(r * W2_IW + 0) * FB_PITCH);
        win[r][1] = lds32(src + (uint32_t)(r * W2_IW + 1) * FB_PITCH);
        win[r][2] = lds32(src + (uint32_t)(r * W2_IW + 2) * FB_PITCH);
      }
      if (it >= 2) mbar_wait(smem_u32(&bar_p[it & 1]), ((uint32_t)(it - 2) >> 1) & 1u);   // the project product of tile it-2 has read this copy of the A operand
      if (tid == 0) fb_mark(P, it, 10);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c + 3 < 10) {
#pragma unroll
          for (int r = 0; r < 4; ++r) win[r][(c + 3) % 4] = lds32(src + (uint32_t)(r * W2_IW + c + 3) * FB_PITCH);
        }
        uint32_t a0 = bdw, a1 = bdw;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            if (a * 3 + b < 8) {
              a0 = hfma2(win[a][(c + b) % 4], wk[a * 3 + b], a0);
              a1 = hfma2(win[a + 1][(c + b) % 4], wk[a * 3 + b], a1);
            } else {
              a0 = hfma2_relu(win[a][(c + b) % 4], wk[a * 3 + b], a0);
              a1 = hfma2_relu(win[a + 1][(c + b) % 4], wk[a * 3 + b], a1);
            }
          }
        const uint32_t r0 = (uint32_t)(2 * rp * W2_TW + 8 * chf + c);
        sts32(arow0 + (uint32_t)c * 128u + ((((uint32_t)lane >> 2) ^ (r0 & 7u)) << 4), a0);
        sts32(arow0 + (uint32_t)(c + W2_TW) * 128u + ((((uint32_t)lane >> 2) ^ ((r0 + W2_TW) & 7u)) << 4), a1);
      }
      fence_proxy_async();
      __syncwarp();
      if (tid == 0) fb_mark(P, it, 11);
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar_a));
        mbar_arrive(smem_u32(&bar_ifree[it & 1]));
      }
    }
  }
  __syncthreads();
  if (warp == WARP_CX) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

typedef CUresult (*FbEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
}  // namespace

static long long* g_fb_dbg = nullptr;
extern "C" void dg_debug_fsrgan_block_timeline(long long* buf) { g_fb_dbg = buf; }

extern "C" int dg_fsrgan_block_infer_supported(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* y) {
  static const char* off = getenv("DG_FSRGAN_BLOCK");
  if (off && off[0] == '0') return 0;
  return ctx && ctx->encode_tiled && ctx->cc_major == 10 && dg_valid(x) && dg_valid(y) && x->dtype == DG_BF16 && y->dtype == DG_BF16 &&
         x->c == FB_C && y->c == FB_C && x->n == y->n && x->h == y->h && x->w == y->w && x->cpitch % 8 == 0 && x->coff % 8 == 0 &&
         y->cpitch % 8 == 0 && y->coff % 8 == 0 && ((uintptr_t)x->ptr % 16) == 0 && ((uintptr_t)y->ptr % 16) == 0 && x->h >= 2 && x->w >= 2 &&
         (long)x->n * x->h * x->w < (1L << 30);
}

extern "C" int dg_fsrgan_block_infer(dg_ctx* ctx, const dg_tensor* x, const void* w_expand, const float* b_expand, const float* w_dw,
                                     const float* b_dw, const void* w_project, const float* b_project, const dg_tensor* y, void* stream) {
  DG_REQUIRE(ctx && w_expand && b_expand && w_dw && b_dw && w_project && b_project, "dg_fsrgan_block_infer: null argument");
  DG_REQUIRE(dg_fsrgan_block_infer_supported(ctx, x, y), "dg_fsrgan_block_infer: needs bf16 NHWC tensors of %d channels (16-byte aligned pixels) on sm_100", FB_C);
  DG_REQUIRE(((uintptr_t)w_expand % 16) == 0 && ((uintptr_t)w_project % 16) == 0, "dg_fsrgan_block_infer: kernels must be 16-byte aligned");
  FbParams P;
  memset(&P, 0, sizeof(P));
  uint64_t dims[4] = {(uint64_t)FB_C, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
  uint64_t strides[3] = {(uint64_t)x->cpitch * 2, (uint64_t)x->cpitch * 2 * x->w, (uint64_t)x->cpitch * 2 * x->w * x->h};
  const char* wsenv = getenv("DG_FSRGAN_BLOCK_WS");     // 0: the lock-step kernel (16 x 8 tiles); default: the warp-specialised one (16 x 4); read per call (tests run both)
  const bool ws = !(wsenv && wsenv[0] == '0');
  uint32_t box[4] = {(uint32_t)FB_C, (uint32_t)FB_IW, (uint32_t)(ws ? W2_IH : FB_IH), 1}, ones[4] = {1, 1, 1, 1};
  CUresult r = ((FbEncodeFn)ctx->encode_tiled)(&P.xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (char*)x->ptr + (size_t)x->coff * 2,
                                               (const cuuint64_t*)dims, (const cuuint64_t*)strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DG_FAIL("dg_fsrgan_block_infer: cuTensorMapEncodeTiled failed (%d)", (int)r);
  P.x = (const __nv_bfloat16*)x->ptr + x->coff;
  P.y = (__nv_bfloat16*)y->ptr + y->coff;
  P.w1 = (const __nv_bfloat16*)w_expand; P.w2 = (const __half*)w_project;
  P.b1 = b_expand; P.wd = w_dw; P.bd = b_dw; P.b2 = b_project;
  P.xp = x->cpitch; P.yp = y->cpitch; P.N = x->n; P.H = x->h; P.W = x->w;
  const int TH = ws ? W2_TH : FB_TH;
  P.tiles_w = (x->w + FB_TW - 1) / FB_TW; P.tiles_h = (x->h + TH - 1) / TH;
  P.total = P.N * P.tiles_h * P.tiles_w;
  P.dbg = g_fb_dbg;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fsrgan_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fsrgan_block_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W2_SMEM);
    if (e != cudaSuccess) DG_FAIL("dg_fsrgan_block_infer: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const unsigned grid = (unsigned)(P.total < ctx->sm_count ? P.total : ctx->sm_count);
  if (ws) fsrgan_block_ws_kernel<<<grid, W2_THREADS, W2_SMEM, (cudaStream_t)stream>>>(P);
  else fsrgan_block_kernel<<<grid, FB_THREADS, FB_SMEM, (cudaStream_t)stream>>>(P);
  DG_CHECK_LAUNCH("dg_fsrgan_block_infer");
  return 0;
}
