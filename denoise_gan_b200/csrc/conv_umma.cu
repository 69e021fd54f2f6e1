// Tensor-core convolution for sm_100a: NHWC implicit GEMM on tcgen05.mma with the accumulator
// in TMEM and operands staged by TMA.
//
// Formulation ("multi-source halo conv").  One launch computes, for an output VIEW y (a dense or
// stride-2 sub-lattice of an NHWC tensor, possibly a channel slice),
//      y[n,h,w,o] = act( bias[o] + sum_taps sum_c  S_t[n, h+dh_t, w+dw_t, c] * Wt[t][o][c] )
// where every source S_t is itself a view (dense or a stride-2 parity lattice) of the input.
//   forward stride 1      : one source (x), taps (r-pad_t, s-pad_l)
//   forward stride 2      : four parity lattices of x, taps fall on one lattice each
//   dgrad stride 1        : one source (dy), flipped taps, transposed weight blocks
//   dgrad stride 2 / Conv2DTranspose forward : four output parity phases, in ONE launch when their accumulators fit TMEM
//                                              together (tap_acc / tap_first / ph_off), else one launch per phase
// A CTA owns an output tile of (16*MT) x 8 pixels.  Per 64/32/16-channel chunk it TMA-loads ONE
// halo box per source ((16*MT+ext_h) x (8+ext_w) pixels, out-of-image pixels zero-filled by TMA =
// SAME padding) and every tap's A operand is a row-shifted window of that box: the UMMA shared
// memory descriptor starts at (row shift)*pixel_bytes and uses SBO = halo_width*pixel_bytes, which
// is legal because the 128B/64B/32B swizzle is a function of the absolute shared-memory address
// (probes/umma_probe.cu, cases halo10_*).  Input traffic per tile is therefore ~1.4x the tile
// instead of taps x.  Weight blocks [Cout_blk][chunk] (K-major) stay resident in shared memory
// when they fit, else they stream with the halo stage.
//
// Warp roles (352 threads): warp 0 TMA producer, warps 1 and 6 TMEM allocator / MMA issuers (alternate tiles when four
// accumulator buffers fit TMEM), warps 2-5 and 7-10 two epilogue groups on alternate tiles (TMEM -> registers -> bias /
// activation -> shared-memory staging + TMA store, or direct global stores; optional BatchNorm statistics).  mbarrier
// pipelines: halo stages (full/empty), TMEM accumulators (full/empty), resident weights, the K-outer weight ring.
// Modes chosen on the host (launch_conv): resident / streamed / split-source / K-outer weights, 1-4 sub-tiles per tile,
// one or four output phases (fused stride-2 dgrad), staged or direct epilogue.
//
// Reference call sites: Conv2D srgan.py:154-182,246-268, fsrgan.py:134-217, autoencoder.py:95-104,
// pix2pix.py:115,207-218; Conv2DTranspose pix2pix.py:130,169; gradients train_srgan.py:111-112.
#include <cuda.h>
#include <stdlib.h>

#include "dg_common.cuh"
#include "sm100.cuh"

namespace {

using namespace sm100;

constexpr int MAX_SRC = 4;
constexpr int MAX_TAPS = 16;
constexpr int MAX_STAGES = 6;
constexpr int MAX_ACC = 16;         // accumulator barriers: 4 rotating buffers, or one per tile of the CTA in the BatchNorm-phase mode
constexpr int CONV_THREADS = 352;   // TMA producer, MMA issuer, 4 epilogue warps (group 0), second MMA issuer, 4 epilogue warps (group 1)
// The two epilogue groups take ALTERNATE tiles: one tile's TMEM -> registers -> shared memory -> store chain is a serial
// latency chain of ~1500 cycles per 128 x 64 tile, longer than the MMAs of thin layers (1x1, 16/32-channel, K = 16 layers)
constexpr uint32_t SMEM_LIMIT = 227 * 1024;

struct UmmaConvParams {
  CUtensorMap src[MAX_SRC];
  CUtensorMap wmap;
  int n_src, n_taps, n_chunks, kc, nb, mt;
  int tiles_h, tiles_w, n_img, out_h, out_w;
  int src_h0[MAX_SRC], src_w0[MAX_SRC];
  uint32_t src_off[MAX_SRC];
  uint32_t a_sbo[MAX_SRC], mt_stride[MAX_SRC];
  uint32_t tap_off[MAX_TAPS];
  int tap_src[MAX_TAPS], tap_w[MAX_TAPS];
  uint64_t tap_adesc[MAX_TAPS];   // complete A descriptor of the tap's window minus the stage base (added to the low word)
  uint32_t tap_ms16[MAX_TAPS];    // descriptor step between the m sub-tiles of the tap's source
  // Output phases: a stride-2 dgrad / Conv2DTranspose computes its four output parity phases in ONE launch.  Every tap
  // belongs to one phase; a phase has its own accumulator (column offset tap_acc) and its own output lattice offset.
  int n_phase;
  uint32_t tap_acc[MAX_TAPS], tap_first[MAX_TAPS];
  long ph_off[4];
  uint32_t stage_bytes, stage_tx, w_stage_off, w_block_bytes, w_res_bytes, w_res_tx;
  int n_stages, resident, cout_total;
  int split;        // 1: one SOURCE per pipeline stage (halo box + the weights of that source's taps): big stride-2 layers
  int src_tap0[MAX_SRC], src_ntaps[MAX_SRC];   // taps are sorted by source
  uint32_t src_tx[MAX_SRC];
  // K-outer mode for streamed weights: an output tile is P.mt sub-tiles of 16x8 pixels, each with its OWN halo stage, and
  // the loop order is chunk-outer / sub-tile-inner, so that a weight chunk (all taps), streamed through a two-slot ring
  // of its own, is loaded once per P.mt sub-tiles instead of once per 128 pixels
  int kouter;
  uint32_t wslot_bytes, wslot_tx;
  int nbuf_shift;   // log2 of the TMEM accumulator buffers: 2 (four buffers, TWO issuing warps on alternate tiles) or 1
  void* out;
  long out_sn, out_sh, out_sw;
  int out_f32;
  // depth_to_space(2) + PReLU as the STORE PATTERN of the convolution in front of them (srgan.py:144-146, fsrgan.py:180-186, inference:
  // the pre-activation is not needed again): conv channel c = blk * d2s_cq + cc lands at pixel (2h + blk / 2, 2w + blk % 2), channel cc of
  // the [n, 2h, 2w, d2s_cq] output described by out / out_s*; slope d2s_prelu[cc] (nullptr: no PReLU).  bf16 direct epilogue only.
  int d2s_cq;
  const float* d2s_prelu;
  // staged epilogue at inference (BatchNorm folded into kernel and bias): y = act(acc + bias) + residual, act = PReLU with per-channel
  // slopes epi_prelu (else `act`); epi_res: bf16 NHWC view of the output's shape (fsrgan.py:172-176, :208-210; srgan.py:166-169)
  const __nv_bfloat16* epi_res;
  long res_sn, res_sh, res_sw;
  const float* epi_prelu;
  int out_cvalid;   // > 0 (fp32 output, one 16-channel N block): only the first out_cvalid channels of a pixel are stored -- the
                    // 3-channel image side (srgan.py:182, fsrgan.py:217) written densely instead of padded to 16 and sliced
  // staged epilogue: the tile is written to shared memory in the TMA swizzle and leaves through ONE bulk tensor store
  // (dense bf16 output views with a 16/32/64-channel N block); optionally the BatchNorm batch-statistics partials
  // (per-channel sum and sum of squares of the STORED values) are accumulated from the staged tile
  CUtensorMap omap;
  int tstore;
  int d2s_ts;       // depth_to_space store through the staging buffers (epilogue_role_d2s_ts)
  uint32_t stg_off, stg_bytes, stg_mask;
  float* bn_partials;   // [gridDim.x][2][cout_total] or nullptr
  int bn_fin;           // the last CTA (ticket) turns the partial rows into scale/shift/mean/invstd + moving statistics
  dg_bn_fused bnf;
  unsigned* ticket;
  // BatchNorm phase ("bnp"): convolution + training-mode BatchNormalization + activation (+ skip-add) in ONE cooperative
  // launch (srgan.py:154-175: every conv of the generator trunk is followed by BN and ReLU / PReLU / Add).  Every output
  // tile of the CTA keeps its accumulator in TMEM (tiles per CTA x mt x nb <= 512 columns); pass 1 is the ordinary staged
  // epilogue (raw conv output y stored for the backward pass, per-CTA statistics partials), then a grid-wide barrier, every
  // CTA reduces the partial rows to scale / shift in the same fixed order, and pass 2 walks the accumulators again:
  // out2 = act(scale * bf16(y) + shift) (+ residual), bit-identical to dg_bn_act_fwd applied to the stored y.
  int bnp, bnp_act, bnp_res;
  float bnp_alpha;
  const float* bnp_prelu;
  CUtensorMap omap2, rmap;
  unsigned* gbar;       // [0] arrivals, [1] departures of the grid barrier (self-resetting)
  // BatchNorm-backward epilogue ("bwd", kernel instance umma_conv_kernel<true>): the launch is an input-gradient convolution
  // whose result is g = dL/da for the output a = act(BN(yb)) (+ skip) of a training-mode BatchNormalization (srgan.py:162-169:
  // the dgrad of every trunk convolution feeds the backward pass of the BatchNorm in front of it).  The epilogue
  //   * adds the gradient that arrives over the skip connection (bwd_res, srgan.py:169 Add) -- no separate add launch,
  //   * stores g (bf16, staged TMA store), and
  //   * accumulates from the rounded g the two per-channel sums of the BatchNorm backward pass, sum g' and sum g' yb
  //     with g' = g * act'(scale * yb + shift), into the per-CTA rows `bn_partials` ([gridDim.x][2][Cout]),
  // so that the BatchNorm backward pass is ONE read of g and yb (dg_bn_bwd_dx_from_partials) instead of two.
  int bwd_mask;                    // 1: yb is the output of a ReLU convolution (no BatchNorm): g is STORED multiplied by (yb > 0) and no sums
                                   // are produced -- the ReLU backward of a conv -> conv chain folded into this launch (autoencoder.py:95-104)
  int bwd_am;                      // -1: no statistics; 0 none / 1 relu / 2 leaky relu: activation behind the BatchNorm
  const __nv_bfloat16* bwd_y;      // yb, the BatchNorm input (same pixel grid as the output), first channel of the view
  const __nv_bfloat16* bwd_res;    // skip-connection gradient or nullptr
  long bwd_y_sn, bwd_y_sh, bwd_y_sw, bwd_r_sn, bwd_r_sh, bwd_r_sw;   // element strides
  const float *bwd_scale, *bwd_shift, *bwd_mean;
  float bwd_alpha;
  const float* bias;
  int act;
  float alpha;
  uint32_t layout, idesc;
  int dbg_flags;   // debug experiments (tools/conv_timeline.py): 1 no stores, 2 no epilogue work, 4 no TMA reloads, 8 unshifted taps; 16 (host) force K-outer; 64: direct depth_to_space store without 32-byte stores
  long long* dbg;  // optional per-role timeline of CTA 0 (tools/conv_timeline.py); nullptr in production
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void dbg_mark(const UmmaConvParams& P, int role, int it, int slot) {
  if (P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && it < 16) P.dbg[(role * 16 + it) * 4 + slot] = clock64();
}

template <int ACT>
__device__ __forceinline__ float act_fn(float v, float alpha) {
  if (ACT == DG_ACT_RELU) return v > 0.f ? v : 0.f;
  if (ACT == DG_ACT_LRELU) return v >= 0.f ? v : alpha * v;
  if (ACT == DG_ACT_TANH) return tanhf(v);
  if (ACT == DG_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}

// tcgen05.wait::ld that also "modifies" the loaded registers: no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// one 16-column group of one accumulator row: + bias, activation, store
template <int ACT, bool F32>
__device__ __forceinline__ void epi_store16(const uint32_t (&v)[16], const float* __restrict__ bs, float alpha, void* out, long elem) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
  if (bs) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] += bs[j];
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = act_fn<ACT>(f[j], alpha);
  if (F32) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + elem);
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
  } else {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + elem);
    dst[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    dst[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
  }
}

template <int ACT, bool F32>
__device__ __forceinline__ void epilogue_role(const UmmaConvParams& P, uint32_t tmem, int q, int lane, int nb0, int total_tiles,
                                              const float* __restrict__ bs, uint64_t* bar_acc_full, uint64_t* bar_acc_empty, int grp,
                                              const float* __restrict__ ps = nullptr) {
  const int m_idx = q * 32 + lane;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  int it = grp;
  for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < total_tiles; tile += 2 * (int)gridDim.x, it += 2) {
    const int b = it & ((1 << P.nbuf_shift) - 1);
    const uint32_t acc_phase = (uint32_t)(it >> P.nbuf_shift) & 1u;
    const int tw = tile % P.tiles_w;
    const int t2 = tile / P.tiles_w;
    const int th = t2 % P.tiles_h;
    const int n = t2 / P.tiles_h;
    if (q == 0 && lane == 0 && grp == 0) dbg_mark(P, 2, it >> 1, 0);
    mbar_wait(smem_u32(&bar_acc_full[b]), acc_phase);
    tc_fence_after();
    if (q == 0 && lane == 0 && grp == 0) dbg_mark(P, 2, it >> 1, 1);
    for (int pm = 0; pm < P.n_phase * P.mt; ++pm) {
      const int phs = pm / P.mt, m = pm - phs * P.mt;
      const int ph = th * 16 * P.mt + m * 16 + (m_idx >> 3);
      const int pw = tw * 8 + (m_idx & 7);
      const bool valid = ph < P.out_h && pw < P.out_w;
      const long pix = (long)n * P.out_sn + (long)ph * P.out_sh + (long)pw * P.out_sw + nb0 + P.ph_off[phs];
      const uint32_t acc = tmem + lane_base + (uint32_t)((b * P.n_phase * P.mt + pm) * P.nb);
      int c0 = 0;
      if (!F32 && P.d2s_cq > 0) {
        // depth_to_space + PReLU store: a 32-channel group lies inside one sub-pixel block (d2s_cq is a multiple of 32, see the host
        // check); slopes from shared memory (ps: broadcast reads).  A thread's 32 channels are 64 contiguous bytes: two 32-byte stores
        // (st.global.v8: full sectors, half the store instructions; 1078 -> 959 us on the 1080p up-convolution) when the row is 32-byte
        // aligned.  Requesting the next 32 accumulator columns before converting the current ones was measured slower (1065 us:
        // profiles/infer_profile_r2d_d2s_flags.log) and is not kept.
        // (Dense 32-channel outputs of a 128-channel layer leave through epilogue_role_d2s_ts instead.)
        const bool wide = !(P.dbg_flags & 64) && (((uintptr_t)P.out | (uintptr_t)(P.out_sw * 2)) & 31) == 0;
        uint32_t va[32];
        auto convert_store = [&](const uint32_t (&v)[32], int cb) {
          const int cg = nb0 + cb, blk = cg / P.d2s_cq, cc = cg - blk * P.d2s_cq;
          const long e = (long)n * P.out_sn + (long)(2 * ph + (blk >> 1)) * P.out_sh + (long)(2 * pw + (blk & 1)) * P.out_sw + cc;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + (bs ? bs[cb + j] : 0.f);
          if (ps) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 a = *reinterpret_cast<const float4*>(ps + cc + 4 * j4);
              f[4 * j4 + 0] = f[4 * j4 + 0] > 0.f ? f[4 * j4 + 0] : a.x * f[4 * j4 + 0];
              f[4 * j4 + 1] = f[4 * j4 + 1] > 0.f ? f[4 * j4 + 1] : a.y * f[4 * j4 + 1];
              f[4 * j4 + 2] = f[4 * j4 + 2] > 0.f ? f[4 * j4 + 2] : a.z * f[4 * j4 + 2];
              f[4 * j4 + 3] = f[4 * j4 + 3] > 0.f ? f[4 * j4 + 3] : a.w * f[4 * j4 + 3];
            }
          }
          uint32_t o[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) o[k] = pack_bf16x2(f[2 * k], f[2 * k + 1]);
          __nv_bfloat16* dstp = reinterpret_cast<__nv_bfloat16*>(P.out) + e;
          if (wide && (cc & 15) == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dstp + 16 * k), "r"(o[8 * k]), "r"(o[8 * k + 1]), "r"(o[8 * k + 2]),
                           "r"(o[8 * k + 3]), "r"(o[8 * k + 4]), "r"(o[8 * k + 5]), "r"(o[8 * k + 6]), "r"(o[8 * k + 7])
                           : "memory");
          } else {
            uint4* dst = reinterpret_cast<uint4*>(dstp);
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
          }
        };
        for (; c0 < P.nb; c0 += 32) {
          tmem_ld_32x32(acc + c0, va);
          tmem_ld_wait_on(va);
          if (valid) convert_store(va, c0);
        }
      }
      for (; c0 + 32 <= P.nb; c0 += 32) {   // two 16-column loads in flight per wait
        uint32_t v0[16], v1[16];
        if (P.dbg_flags & 2) continue;
        tmem_ld_32x16(acc + c0, v0);
        tmem_ld_32x16(acc + c0 + 16, v1);
        tmem_ld_wait();
        if (valid && !(P.dbg_flags & 1)) {
          epi_store16<ACT, F32>(v0, bs ? bs + c0 : nullptr, P.alpha, P.out, pix + c0);
          epi_store16<ACT, F32>(v1, bs ? bs + c0 + 16 : nullptr, P.alpha, P.out, pix + c0 + 16);
        }
      }
      if (c0 < P.nb) {
        uint32_t v0[16];
        tmem_ld_32x16(acc + c0, v0);
        tmem_ld_wait();
        if (F32 && P.out_cvalid > 0) {
          if (valid) {
            float* dst = reinterpret_cast<float*>(P.out) + pix;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < P.out_cvalid) dst[j] = act_fn<ACT>(__uint_as_float(v0[j]) + (bs ? bs[j] : 0.f), P.alpha);
          }
        } else if (valid) epi_store16<ACT, F32>(v0, bs ? bs + c0 : nullptr, P.alpha, P.out, pix + c0);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[b]));
    if (q == 0 && lane == 0 && grp == 0) dbg_mark(P, 2, it >> 1, 2);
  }
}

__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t mask) { return off ^ (((off >> 7) & mask) << 4); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// one 16-column group of one accumulator row: + bias, activation, bf16, two 16-byte chunks of the staged row
// (EXTRA: PReLU slopes `sl` instead of ACT, then + the 16 bf16 values at `rr` -- the skip connection -- when they are given)
template <int ACT, bool EXTRA = false>
__device__ __forceinline__ void epi_stage16(const uint32_t (&v)[16], const float* __restrict__ bs, float alpha, bool valid, uint32_t stg,
                                            uint32_t off, uint32_t mask, const __nv_bfloat16* __restrict__ rr = nullptr,
                                            const float* __restrict__ sl = nullptr) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
  if (bs) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] += bs[j];
  }
  if (EXTRA && sl) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = f[j] > 0.f ? f[j] : sl[j] * f[j];
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = act_fn<ACT>(f[j], alpha);
  }
  if (EXTRA && rr && valid) {
    const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rr)), r1 = __ldg(reinterpret_cast<const uint4*>(rr) + 1);
    const uint32_t r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[2 * j] += __uint_as_float(r[j] << 16);
      f[2 * j + 1] += __uint_as_float(r[j] & 0xffff0000u);
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = valid ? f[j] : 0.f;   // rows outside the image: zeros (clipped by the store, neutral in the sums)
  st_shared_v4(stg + swz(off, mask), pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  st_shared_v4(stg + swz(off + 16u, mask), pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
               pack_bf16x2(f[14], f[15]));
}

__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

// One 16-column group of pass 2: y = bf16(acc + bias) as stored by pass 1, t = scale*y + shift, activation, + residual.
template <int ACT2>
__device__ __forceinline__ void bnp_stage16(const uint32_t (&v)[16], const float* __restrict__ bs, const float* __restrict__ cs, int c0,
                                            float alpha, bool has_res, uint32_t res, uint32_t stg, uint32_t off, uint32_t mask) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
  if (bs) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] += bs[c0 + j];
  }
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 sc = *reinterpret_cast<const float4*>(cs + c0 + 4 * j4);
    const float4 sh = *reinterpret_cast<const float4*>(cs + 64 + c0 + 4 * j4);
    f[4 * j4 + 0] = fmaf(bf16_round(f[4 * j4 + 0]), sc.x, sh.x);
    f[4 * j4 + 1] = fmaf(bf16_round(f[4 * j4 + 1]), sc.y, sh.y);
    f[4 * j4 + 2] = fmaf(bf16_round(f[4 * j4 + 2]), sc.z, sh.z);
    f[4 * j4 + 3] = fmaf(bf16_round(f[4 * j4 + 3]), sc.w, sh.w);
  }
  if (ACT2 == DG_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
  } else if (ACT2 == DG_ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = f[j] >= 0.f ? f[j] : alpha * f[j];
  } else if (ACT2 == DG_ACT_PRELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = f[j] > 0.f ? f[j] : cs[128 + c0 + j] * f[j];
  }
  if (has_res) {
    uint32_t r[8];
    ld_shared_v4(res + swz(off, mask), r[0], r[1], r[2], r[3]);
    ld_shared_v4(res + swz(off + 16u, mask), r[4], r[5], r[6], r[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[2 * j] += __uint_as_float(r[j] << 16);
      f[2 * j + 1] += __uint_as_float(r[j] & 0xffff0000u);
    }
  }
  st_shared_v4(stg + swz(off, mask), pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  st_shared_v4(stg + swz(off + 16u, mask), pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
               pack_bf16x2(f[14], f[15]));
}

template <int ACT2>
__device__ __forceinline__ void bnp_pass2(const UmmaConvParams& P, uint32_t tmem, uint32_t stg, int q, int lane, int nb0, int total_tiles,
                                          const float* __restrict__ bs, const float* __restrict__ cs, uint64_t* bar_res_full,
                                          uint64_t* bar_res_empty, uint32_t stage_base, int grp) {
  const uint32_t bar_id = 1u + (uint32_t)grp;
  const int m_idx = q * 32 + lane;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t RB = (uint32_t)P.nb * 2u, mask = P.stg_mask;
  const bool leader = q == 0 && lane == 0;
  const bool has_res = P.bnp_res != 0;
  int it = grp;
  for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < total_tiles; tile += 2 * (int)gridDim.x, it += 2) {
    const int tw = tile % P.tiles_w, t2 = tile / P.tiles_w, th = t2 % P.tiles_h, n = t2 / P.tiles_h;
    const int slot = it % P.n_stages;
    const uint32_t res = stage_base + (uint32_t)slot * P.stage_bytes;
    if (leader) tma_store_wait_read<0>();       // this group's previous store has left the staging buffer
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    if (has_res) mbar_wait(smem_u32(&bar_res_full[slot]), ((uint32_t)(it / P.n_stages)) & 1u);
    for (int m = 0; m < P.mt; ++m) {
      const uint32_t acc = tmem + lane_base + (uint32_t)((it * P.mt + m) * P.nb);
      const uint32_t row_off = (uint32_t)(m * 128 + m_idx) * RB;
      int c0 = 0;
      for (; c0 + 32 <= P.nb; c0 += 32) {
        uint32_t v0[16], v1[16];
        tmem_ld_32x16(acc + c0, v0);
        tmem_ld_32x16(acc + c0 + 16, v1);
        tmem_ld_wait();
        bnp_stage16<ACT2>(v0, bs, cs, c0, P.bnp_alpha, has_res, res, stg, row_off + 2u * c0, mask);
        bnp_stage16<ACT2>(v1, bs, cs, c0 + 16, P.bnp_alpha, has_res, res, stg, row_off + 2u * c0 + 32u, mask);
      }
      if (c0 < P.nb) {
        uint32_t v0[16];
        tmem_ld_32x16(acc + c0, v0);
        tmem_ld_wait();
        bnp_stage16<ACT2>(v0, bs, cs, c0, P.bnp_alpha, has_res, res, stg, row_off + 2u * c0, mask);
      }
    }
    fence_proxy_async();
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    if (leader) {
      tma_store_4d(&P.omap2, stg, nb0, tw * 8, th * 16 * P.mt, n);
      tma_store_commit();
      if (has_res) mbar_arrive(smem_u32(&bar_res_empty[slot]));    // every thread of the group has read the residual tile
    }
  }
  if (leader) tma_store_wait<0>();
}

// BatchNorm phase of the fused convolution (see UmmaConvParams::bnp): grid barrier, fixed-order reduction of the per-CTA
// partial rows by EVERY CTA (double precision, same order everywhere, so all CTAs hold identical coefficients and the result
// does not depend on scheduling), CTA 0 publishes scale / shift / mean / invstd and updates the moving statistics, then pass 2.
template <int ACT>
__device__ __forceinline__ void bnp_phase(const UmmaConvParams& P, uint32_t tmem, uint32_t stg_base, uint32_t stg, int q, int lane, int nb0,
                                          int total_tiles, const float* __restrict__ bs, float* red_s, float* cs, uint64_t* bar_res_full,
                                          uint64_t* bar_res_empty, uint32_t stage_base, int grp) {
  const int etid = grp * 128 + q * 32 + lane;
  __threadfence();                                   // this CTA's partial row is visible device-wide before it arrives
  asm volatile("bar.sync 3, 256;" ::: "memory");
  if (etid == 0) {
    __threadfence();
    atomicAdd(P.gbar, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(P.gbar) : "memory");
    } while (v < gridDim.x);
    __threadfence();
  }
  asm volatile("bar.sync 3, 256;" ::: "memory");
  {
    const int C = P.cout_total, E = 2 * C, E4 = E >> 2, nblk = (int)gridDim.x;
    const int G = E4 < 256 ? 256 / E4 : 1;
    double* gsum = reinterpret_cast<double*>(red_s);          // [G][E] doubles = 8 KB <= the two staging buffers (idle: all stores have been waited for)
    const float4* part4 = reinterpret_cast<const float4*>(P.bn_partials);
    for (int idx = etid; idx < E4 * G; idx += 256) {
      const int e4 = idx % E4, g2 = idx / E4;
      double sacc[4] = {0.0, 0.0, 0.0, 0.0};
      int bI = g2;
      for (; bI + 3 * G < nblk; bI += 4 * G) {     // four independent loads in flight
        const float4 v0 = __ldcg(part4 + (long)bI * E4 + e4), v1 = __ldcg(part4 + (long)(bI + G) * E4 + e4);
        const float4 v2 = __ldcg(part4 + (long)(bI + 2 * G) * E4 + e4), v3 = __ldcg(part4 + (long)(bI + 3 * G) * E4 + e4);
        sacc[0] += (double)v0.x; sacc[1] += (double)v0.y; sacc[2] += (double)v0.z; sacc[3] += (double)v0.w;
        sacc[0] += (double)v1.x; sacc[1] += (double)v1.y; sacc[2] += (double)v1.z; sacc[3] += (double)v1.w;
        sacc[0] += (double)v2.x; sacc[1] += (double)v2.y; sacc[2] += (double)v2.z; sacc[3] += (double)v2.w;
        sacc[0] += (double)v3.x; sacc[1] += (double)v3.y; sacc[2] += (double)v3.z; sacc[3] += (double)v3.w;
      }
      for (; bI < nblk; bI += G) {
        const float4 v = __ldcg(part4 + (long)bI * E4 + e4);
        sacc[0] += (double)v.x; sacc[1] += (double)v.y; sacc[2] += (double)v.z; sacc[3] += (double)v.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) gsum[(long)g2 * E + e4 * 4 + k] = sacc[k];
    }
    asm volatile("bar.sync 3, 256;" ::: "memory");
    const dg_bn_fused& B = P.bnf;
    for (int c = etid; c < C; c += 256) {
      double s0d = 0.0, s1d = 0.0;
      for (int g2 = 0; g2 < G; ++g2) { s0d += gsum[(long)g2 * E + c]; s1d += gsum[(long)g2 * E + C + c]; }
      const double mean = s0d / (double)B.pixels;
      double var = s1d / (double)B.pixels - mean * mean;
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)B.eps));
      const float ga = B.gamma[c], be = B.beta[c];
      const float sc = ga * invstd, sh = be - (float)mean * ga * invstd;
      cs[c] = sc;
      cs[64 + c] = sh;
      if (P.bnp_prelu) cs[128 + c] = P.bnp_prelu[c];
      if (blockIdx.x == 0) {
        B.scale[c] = sc;
        B.shift[c] = sh;
        B.save_mean[c] = (float)mean;
        B.save_invstd[c] = invstd;
        if (B.moving_mean) {
          B.moving_mean[c] = B.moving_mean[c] * B.momentum + (float)mean * (1.f - B.momentum);
          B.moving_var[c] = B.moving_var[c] * B.momentum +
                            (float)(var * ((double)B.pixels / (double)(B.pixels > 1 ? B.pixels - 1 : 1))) * (1.f - B.momentum);   // Bessel-corrected
        }
      }
    }
    asm volatile("bar.sync 3, 256;" ::: "memory");
    if (etid == 0 && atomicAdd(P.gbar + 1, 1u) == gridDim.x - 1u) {    // last CTA past the barrier re-arms it for the next launch
      atomicExch(P.gbar + 1, 0u);
      atomicExch(P.gbar, 0u);
    }
  }
  switch (P.bnp_act) {
    case DG_ACT_RELU: bnp_pass2<DG_ACT_RELU>(P, tmem, stg, q, lane, nb0, total_tiles, bs, cs, bar_res_full, bar_res_empty, stage_base, grp); break;
    case DG_ACT_LRELU: bnp_pass2<DG_ACT_LRELU>(P, tmem, stg, q, lane, nb0, total_tiles, bs, cs, bar_res_full, bar_res_empty, stage_base, grp); break;
    case DG_ACT_PRELU: bnp_pass2<DG_ACT_PRELU>(P, tmem, stg, q, lane, nb0, total_tiles, bs, cs, bar_res_full, bar_res_empty, stage_base, grp); break;
    default: bnp_pass2<DG_ACT_NONE>(P, tmem, stg, q, lane, nb0, total_tiles, bs, cs, bar_res_full, bar_res_empty, stage_base, grp); break;
  }
}

// depth_to_space(2) + PReLU store through shared memory (32 -> 4 x 32 channels, fsrgan.py:180-186 at inference; one 128-channel N
// block, 16 x 8 tiles).  The direct store of epilogue_role gives every thread its own 128-byte lines: a warp-wide 16-byte store touches
// 32 lines, and the LSU takes a line per cycle -- 2048 cycles per tile on the four epilogue warps' stores alone.  Here a thread
// writes its four 64-byte groups into the tile's OUTPUT image in shared memory -- 32 x 16 pixels x 32 channels, rows of two pixels
// (128 bytes) in the TMA's 128-byte swizzle: chunk k of row R sits at k ^ (R & 7) and R & 7 is the pixel's column in the tile, so the
// eight columns of a warp-wide store cover all banks -- and one bulk tensor store per tile writes full lines, clipped at the image
// border by the TMA unit.  Two staging buffers (one per epilogue group): the store of tile i drains under tile i+1.
// (Bias and slope products as packed fp32 pairs -- add.rn.f32x2 / mul.rn.f32x2 -- with the slopes in registers were measured neutral,
// 742 against 725 us: the epilogue is not issue-bound; the kernel sits at ~60 % of the shared-memory port, which the tcgen05 operand
// reads of an N = 128 instruction alone saturate while they run.  Not kept.)
__device__ __forceinline__ void epilogue_role_d2s_ts(const UmmaConvParams& P, uint32_t tmem, uint32_t stg_base, int q, int lane, int total_tiles,
                                                     const float* __restrict__ bs, uint64_t* bar_acc_full, uint64_t* bar_acc_empty, int grp,
                                                     const float* __restrict__ ps) {
  const uint32_t bar_id = 1u + (uint32_t)grp;
  const int m_idx = q * 32 + lane, ph_l = m_idx >> 3, pw_l = m_idx & 7;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const bool leader = q == 0 && lane == 0;
  const uint32_t stg = stg_base + (uint32_t)grp * P.stg_bytes;
  int it = grp;
  for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < total_tiles; tile += 2 * (int)gridDim.x, it += 2) {
    const int b = it & ((1 << P.nbuf_shift) - 1);
    const uint32_t acc_phase = (uint32_t)(it >> P.nbuf_shift) & 1u;
    const int tw = tile % P.tiles_w, t2 = tile / P.tiles_w, th = t2 % P.tiles_h, n = t2 / P.tiles_h;
    if (leader) tma_store_wait_read<0>();     // this group's previous store has left the buffer
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    mbar_wait(smem_u32(&bar_acc_full[b]), acc_phase);
    tc_fence_after();
    const uint32_t acc = tmem + lane_base + (uint32_t)(b * P.nb);
#pragma unroll 1
    for (int blk = 0; blk < 4; ++blk) {
      uint32_t v[32];
      tmem_ld_32x32(acc + (uint32_t)(blk * 32), v);
      tmem_ld_wait_on(v);
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + (bs ? bs[blk * 32 + j] : 0.f);
      if (ps) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 a = *reinterpret_cast<const float4*>(ps + 4 * j4);
          f[4 * j4 + 0] = f[4 * j4 + 0] > 0.f ? f[4 * j4 + 0] : a.x * f[4 * j4 + 0];
          f[4 * j4 + 1] = f[4 * j4 + 1] > 0.f ? f[4 * j4 + 1] : a.y * f[4 * j4 + 1];
          f[4 * j4 + 2] = f[4 * j4 + 2] > 0.f ? f[4 * j4 + 2] : a.z * f[4 * j4 + 2];
          f[4 * j4 + 3] = f[4 * j4 + 3] > 0.f ? f[4 * j4 + 3] : a.w * f[4 * j4 + 3];
        }
      }
      const uint32_t row = stg + (uint32_t)(((2 * ph_l + (blk >> 1)) * 8 + pw_l) * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        st_shared_v4(row + ((uint32_t)(((blk & 1) * 4 + k) ^ pw_l) << 4), pack_bf16x2(f[8 * k], f[8 * k + 1]), pack_bf16x2(f[8 * k + 2], f[8 * k + 3]),
                     pack_bf16x2(f[8 * k + 4], f[8 * k + 5]), pack_bf16x2(f[8 * k + 6], f[8 * k + 7]));
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[b]));   // the accumulator is free as soon as it is in registers
    fence_proxy_async();                                         // generic-proxy writes -> visible to the bulk store
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    if (leader) {
      tma_store_4d(&P.omap, stg, 0, tw * 8, th * 32, n);        // (pixel pairs x 32 channels, pair column, output row, image)
      tma_store_commit();
    }
  }
  if (leader) tma_store_wait<0>();
}

// Epilogue through shared memory: TMEM -> registers -> bias/activation -> bf16 rows in the TMA swizzle (conflict-free
// 16-byte stores) -> one cp.async.bulk.tensor store per tile (full 128-byte lines instead of 32 scattered 16-byte
// segments per warp instruction), double-buffered so the store of tile i drains under the epilogue of tile i+1.
// With P.bn_partials the same staged tile feeds the BatchNorm batch statistics: a thread owns one channel pair and
// walks the rows of its warp's quarter (one 4-byte word per lane: conflict-free), so the statistics cost no extra
// pass over the tensor and no launch (srgan.py:155 BatchNormalization after every conv of the residual trunk).
template <int ACT, bool EXTRA = false>
__device__ __forceinline__ void epilogue_role_ts(const UmmaConvParams& P, uint32_t tmem, uint32_t stg_base, int q, int lane, int nb0,
                                                 int total_tiles, const float* __restrict__ bs, uint64_t* bar_acc_full,
                                                 uint64_t* bar_acc_empty, float* red_s, int grp, float* bnp_s, uint64_t* bar_res_full,
                                                 uint64_t* bar_res_empty, uint32_t stage_base) {
  const uint32_t bar_id = 1u + (uint32_t)grp;   // named barrier of this group's 128 threads (3: both groups)
#define DG_GROUP_SYNC() asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory")
  const int m_idx = q * 32 + lane;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t RB = (uint32_t)P.nb * 2u, mask = P.stg_mask;
  const int WPR = P.nb >> 1, RPR = 32 / WPR;          // 4-byte words per staged row, rows per warp-wide read
  const int sub = lane / WPR, w = lane - sub * WPR;
  const bool leader = q == 0 && lane == 0;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const uint32_t stg = stg_base + (uint32_t)grp * P.stg_bytes;   // one staging buffer per group
  int it = grp;
  for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < total_tiles; tile += 2 * (int)gridDim.x, it += 2) {
    const int b = P.bnp ? it : (it & ((1 << P.nbuf_shift) - 1));
    const uint32_t acc_phase = P.bnp ? 0u : ((uint32_t)(it >> P.nbuf_shift) & 1u);
    const int tw = tile % P.tiles_w;
    const int t2 = tile / P.tiles_w;
    const int th = t2 % P.tiles_h;
    const int n = t2 / P.tiles_h;
    if (leader) {
      if (grp == 0) dbg_mark(P, 2, it >> 1, 0);
      tma_store_wait_read<0>();     // this group's previous store has left the buffer (it drained under the other group's tile)
    }
    DG_GROUP_SYNC();
    mbar_wait(smem_u32(&bar_acc_full[b]), acc_phase);
    tc_fence_after();
    if (leader && grp == 0) dbg_mark(P, 2, it >> 1, 1);
    for (int m = 0; m < P.mt; ++m) {
      const int ph = th * 16 * P.mt + m * 16 + (m_idx >> 3);
      const int pw = tw * 8 + (m_idx & 7);
      const bool valid = ph < P.out_h && pw < P.out_w;
      const uint32_t acc = tmem + lane_base + (uint32_t)((b * P.mt + m) * P.nb);
      const uint32_t row_off = (uint32_t)(m * 128 + m_idx) * RB;
      const __nv_bfloat16* rr = (EXTRA && P.epi_res) ? P.epi_res + (long)n * P.res_sn + (long)ph * P.res_sh + (long)pw * P.res_sw + nb0 : nullptr;
      const float* sl = (EXTRA && P.epi_prelu) ? bnp_s : nullptr;
      int c0 = 0;
      for (; c0 + 32 <= P.nb; c0 += 32) {
        uint32_t v0[16], v1[16];
        tmem_ld_32x16(acc + c0, v0);
        tmem_ld_32x16(acc + c0 + 16, v1);
        tmem_ld_wait();
        epi_stage16<ACT, EXTRA>(v0, bs ? bs + c0 : nullptr, P.alpha, valid, stg, row_off + 2u * c0, mask, rr ? rr + c0 : nullptr, sl ? sl + c0 : nullptr);
        epi_stage16<ACT, EXTRA>(v1, bs ? bs + c0 + 16 : nullptr, P.alpha, valid, stg, row_off + 2u * c0 + 32u, mask, rr ? rr + c0 + 16 : nullptr,
                                sl ? sl + c0 + 16 : nullptr);
      }
      if (c0 < P.nb) {
        uint32_t v0[16];
        tmem_ld_32x16(acc + c0, v0);
        tmem_ld_wait();
        epi_stage16<ACT, EXTRA>(v0, bs ? bs + c0 : nullptr, P.alpha, valid, stg, row_off + 2u * c0, mask, rr ? rr + c0 : nullptr, sl ? sl + c0 : nullptr);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[b]));   // the accumulator is free as soon as it is in registers
    fence_proxy_async();                                         // generic-proxy writes -> visible to the bulk store
    DG_GROUP_SYNC();
    if (leader) {
      if (!(P.dbg_flags & 1)) tma_store_4d(&P.omap, stg, nb0, tw * 8, th * 16 * P.mt, n);
      tma_store_commit();
      if (grp == 0) dbg_mark(P, 2, it >> 1, 2);
    }
    if (P.bn_partials) {
      const int iters = 32 / RPR;
      for (int m = 0; m < P.mt; ++m) {
        const uint32_t r0 = (uint32_t)(m * 128 + q * 32 + sub);
#pragma unroll 8
        for (int i = 0; i < iters; ++i) {
          const uint32_t off = (r0 + (uint32_t)(i * RPR)) * RB + 4u * (uint32_t)w;
          const uint32_t word = ld_shared_u32(stg + swz(off, mask));
          const float a = __uint_as_float(word << 16), c = __uint_as_float(word & 0xffff0000u);
          s0 += a; q0 = fmaf(a, a, q0);
          s1 += c; q1 = fmaf(c, c, q1);
        }
      }
    }
  }
  if (leader) tma_store_wait<0>();
  if (P.bn_partials) {
    asm volatile("bar.sync 3, 256;" ::: "memory");   // both groups: all stores have left the staging buffers, buffer 0 is the scratch of the final sum
    for (int o = WPR; o < 32; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    }
    const int gw = grp * 4 + q;     // one row of the scratch per epilogue warp: [8][4][32] floats = 4 KB <= one staging buffer
    if (lane < WPR) {
      red_s[(gw * 4 + 0) * 32 + lane] = s0; red_s[(gw * 4 + 1) * 32 + lane] = s1;
      red_s[(gw * 4 + 2) * 32 + lane] = q0; red_s[(gw * 4 + 3) * 32 + lane] = q1;
    }
    asm volatile("bar.sync 3, 256;" ::: "memory");
    if (grp == 0 && q == 0 && lane < WPR) {
      float t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        t[k] = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t[k] += red_s[(r * 4 + k) * 32 + lane];
      }
      float* dst = P.bn_partials + (size_t)blockIdx.x * 2 * P.cout_total + nb0 + 2 * lane;
      dst[0] = t[0]; dst[1] = t[1];
      dst[P.cout_total] = t[2]; dst[P.cout_total + 1] = t[3];
    }
    if (P.bnp) bnp_phase<ACT>(P, tmem, stg_base, stg, q, lane, nb0, total_tiles, bs, red_s, bnp_s, bar_res_full, bar_res_empty, stage_base, grp);
    if (P.bn_fin) {
      // ---- last CTA of the grid: fixed-order double-precision sum of the rows, then the BatchNorm coefficients
      __shared__ int s_last;
      const int etid = grp * 128 + q * 32 + lane;     // any bijection onto 0..255
      __threadfence();
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (etid == 0) s_last = (atomicAdd(P.ticket, 1u) == gridDim.x * gridDim.y - 1u) ? 1 : 0;
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (s_last) {
        __threadfence();
        const int C = P.cout_total, E = 2 * C, E4 = E >> 2, nblk = (int)gridDim.x;
        const int G = E4 < 256 ? 256 / E4 : 1;
        double* gsum = reinterpret_cast<double*>(red_s);          // [G][E] doubles <= 8 KB = the two staging buffers
        const float4* part4 = reinterpret_cast<const float4*>(P.bn_partials);
        for (int idx = etid; idx < E4 * G; idx += 256) {
          const int e4 = idx % E4, g2 = idx / E4;
          double sacc[4] = {0.0, 0.0, 0.0, 0.0};
          int bI = g2;
          for (; bI + 3 * G < nblk; bI += 4 * G) {     // four independent loads in flight
            const float4 v0 = __ldcg(part4 + (long)bI * E4 + e4), v1 = __ldcg(part4 + (long)(bI + G) * E4 + e4);
            const float4 v2 = __ldcg(part4 + (long)(bI + 2 * G) * E4 + e4), v3 = __ldcg(part4 + (long)(bI + 3 * G) * E4 + e4);
            sacc[0] += (double)v0.x; sacc[1] += (double)v0.y; sacc[2] += (double)v0.z; sacc[3] += (double)v0.w;
            sacc[0] += (double)v1.x; sacc[1] += (double)v1.y; sacc[2] += (double)v1.z; sacc[3] += (double)v1.w;
            sacc[0] += (double)v2.x; sacc[1] += (double)v2.y; sacc[2] += (double)v2.z; sacc[3] += (double)v2.w;
            sacc[0] += (double)v3.x; sacc[1] += (double)v3.y; sacc[2] += (double)v3.z; sacc[3] += (double)v3.w;
          }
          for (; bI < nblk; bI += G) {
            const float4 v = __ldcg(part4 + (long)bI * E4 + e4);
            sacc[0] += (double)v.x; sacc[1] += (double)v.y; sacc[2] += (double)v.z; sacc[3] += (double)v.w;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) gsum[(long)g2 * E + e4 * 4 + k] = sacc[k];
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
        const dg_bn_fused& B = P.bnf;
        for (int c = etid; c < C; c += 256) {
          double s0d = 0.0, s1d = 0.0;
          for (int g2 = 0; g2 < G; ++g2) { s0d += gsum[(long)g2 * E + c]; s1d += gsum[(long)g2 * E + C + c]; }
          const double mean = s0d / (double)B.pixels;
          double var = s1d / (double)B.pixels - mean * mean;
          if (var < 0.0) var = 0.0;
          const float invstd = (float)(1.0 / sqrt(var + (double)B.eps));
          const float ga = B.gamma[c], be = B.beta[c];
          B.scale[c] = ga * invstd;
          B.shift[c] = be - (float)mean * ga * invstd;
          B.save_mean[c] = (float)mean;
          B.save_invstd[c] = invstd;
          if (B.moving_mean) {
            B.moving_mean[c] = B.moving_mean[c] * B.momentum + (float)mean * (1.f - B.momentum);
            B.moving_var[c] = B.moving_var[c] * B.momentum + (float)(var * ((double)B.pixels / (double)(B.pixels > 1 ? B.pixels - 1 : 1))) * (1.f - B.momentum);   // Bessel-corrected
          }
        }
        if (etid == 0) *P.ticket = 0u;
      }
    }
  }
}

#undef DG_GROUP_SYNC

__device__ __forceinline__ uint32_t ldg_u32(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
// (no "memory" clobber on the two helpers below: volatile keeps them ordered among themselves -- a thread reads its yb words
// before it overwrites them -- while the compiler stays free to hoist the coefficient loads and interleave the arithmetic of
// several column blocks; with the clobber every block waited for its own shared-memory loads, IPC 0.27 per scheduler)
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ uint32_t ld_shared_u32_nc(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Epilogue of the BatchNorm-backward instance (UmmaConvParams::bwd_*).
// The accumulator is read in the mma-fragment layout (tcgen05.ld 16x256b, sm100.cuh): a thread owns the column pairs
// 8j + 2(T%4) + {0,1} (j < NB/8) of FOUR rows of a 128-row sub-tile -- the pixels (4q + {0,1,2,3}, T/4) of the 16 x 8 tile, q = the
// warp's TMEM lane quarter -- so the per-channel sums of the BatchNorm backward pass accumulate IN THE THREAD over rows and over
// all tiles of the CTA, and the only cross-lane step is one transposed reduction over the 8 lanes that share T%4 at the very end
// (28 shuffles per thread and kernel).  A first version read one accumulator row per thread (32x32b) and reduced 128 rows x
// 128 values per tile with shuffles: 496 per thread and tile, 5-8 K cycles per tile against the 3.4 K the MMAs leave an epilogue
// group (job r2_05/06: 29 us per launch instead of 15).
// The BatchNorm input yb travels IN PLACE through the staging buffer: the group's leader, once its previous bulk store has been
// read out of the buffer, TMA-loads the yb tile of the group's NEXT output tile into it (same box, same swizzle as the store);
// a thread reads its 4-byte yb words (conflict-free: the 8 quads of a warp hit 8 different 16-byte chunks) and later overwrites
// exactly those words with g.  (Version 2 fetched yb with 4-byte global loads, 8 lines per warp instruction: 256 L1 wavefronts
// per warp and tile, 22-25 us per launch, job r2_09.)  The skip gradient still comes from global memory (17 of the 33 trunk
// layers have one); shared memory has no room for a second tile per group.
// Per element: g = acc + skip gradient, rounded to bf16 (the stored value); g' = g * act'(scale*yb + shift); sums g' and g'*yb
// (the mean is taken out by dg_bn_bwd_dx_from_partials: sum g'(yb - mean) = sum g' yb - mean sum g', in double precision).
template <int AM, int NB>
__device__ __forceinline__ void epilogue_role_bwd(const UmmaConvParams& P, uint32_t tmem, uint32_t stg_base, int q, int lane, int nb0,
                                                  int total_tiles, uint64_t* bar_acc_full, uint64_t* bar_acc_empty, float* red_s, int grp,
                                                  const float* __restrict__ cs, uint64_t* bar_y_full) {
  constexpr int NJ = NB / 8;          // 8-column blocks of the N block
  constexpr int V = 4 * NJ;           // per-thread accumulators: [2 sums][NJ][2 columns]
  const uint32_t bar_id = 1u + (uint32_t)grp;
  const uint32_t RB = (uint32_t)NB * 2u, mask = P.stg_mask;
  const bool leader = q == 0 && lane == 0;
  const bool has_res = P.bwd_res != nullptr;
  const uint32_t stg = stg_base + (uint32_t)grp * P.stg_bytes;
  const uint32_t ybar = smem_u32(&bar_y_full[grp]);
  const int t0 = lane & 3, t1 = lane >> 2;
  const int step_tiles = 2 * (int)gridDim.x;
  float S[V];
#pragma unroll
  for (int i = 0; i < V; ++i) S[i] = 0.f;
  int it = grp;
  const int tile_first = blockIdx.x + grp * (int)gridDim.x;
  if (AM >= 0 && leader && tile_first < total_tiles) {      // the staging buffer starts out free: yb of the group's first tile
    const int tw = tile_first % P.tiles_w, t2 = tile_first / P.tiles_w, th = t2 % P.tiles_h, n = t2 / P.tiles_h;
    mbar_expect_tx(ybar, P.stg_bytes);
    tma_load_4d(stg, &P.rmap, ybar, nb0, tw * 8, th * 16 * P.mt, n);
  }
  for (int tile = tile_first; tile < total_tiles; tile += step_tiles, it += 2) {
    const int b = it & ((1 << P.nbuf_shift) - 1);
    const uint32_t acc_phase = (uint32_t)(it >> P.nbuf_shift) & 1u;
    const int tw = tile % P.tiles_w, t2 = tile / P.tiles_w, th = t2 % P.tiles_h, n = t2 / P.tiles_h;
    const int pw = tw * 8 + t1;
    for (int mh = 0; mh < 2 * P.mt; ++mh) {          // (sub-tile m, 16-lane half hb) of this warp's quarter
      const int m = mh >> 1, hb = mh & 1;
      const int r0 = q * 32 + hb * 16 + t1;          // accumulator rows r0 and r0 + 8 of sub-tile m
      const int ph0 = th * 16 * P.mt + m * 16 + (r0 >> 3);
      const bool va = ph0 < P.out_h && pw < P.out_w, vb = ph0 + 1 < P.out_h && pw < P.out_w;
      uint32_t ra[NJ], rb[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) { ra[j] = rb[j] = 0u; }
      if (has_res) {
        const __nv_bfloat16* rp = P.bwd_res + ((long)n * P.bwd_r_sn + (long)ph0 * P.bwd_r_sh + (long)pw * P.bwd_r_sw + nb0 + 2 * t0);
        if (va) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) ra[j] = ldg_u32(rp + 8 * j);
        }
        if (vb) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) rb[j] = ldg_u32(rp + P.bwd_r_sh + 8 * j);
        }
      }
      if (mh == 0) {
        // the lines of the other half's rows (two image rows further down) start travelling to L1 under the waits below
        if (has_res && ph0 + 3 < P.out_h && pw < P.out_w) {
          const __nv_bfloat16* rp2 = P.bwd_res + ((long)n * P.bwd_r_sn + (long)(ph0 + 2) * P.bwd_r_sh + (long)pw * P.bwd_r_sw + nb0 + 2 * t0);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(rp2));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(rp2 + P.bwd_r_sh));
        }
        if (leader && grp == 0) dbg_mark(P, 2, it >> 1, 0);
        if (AM >= 0) mbar_wait(ybar, (uint32_t)(it >> 1) & 1u);      // yb tile landed in the staging buffer (the previous store has left it)
        else {
          if (leader) tma_store_wait_read<0>();     // this group's previous store has left the staging buffer
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        }
        mbar_wait(smem_u32(&bar_acc_full[b]), acc_phase);
        tc_fence_after();
        if (leader && grp == 0) dbg_mark(P, 2, it >> 1, 1);
      }
      uint32_t v[4 * NJ];
      const uint32_t acc = tmem + ((uint32_t)(q * 32 + hb * 16) << 16) + (uint32_t)((b * P.mt + m) * NB);
      if constexpr (NJ == 8) tmem_ld_16x256b_x8(acc, v);
      else tmem_ld_16x256b_x4(acc, v);
      const uint32_t offa = (uint32_t)(m * 128 + r0) * RB + 4u * (uint32_t)t0, offb = offa + 8u * RB;
      uint32_t ya[NJ], yb[NJ];
      if (AM >= 0) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          ya[j] = ld_shared_u32_nc(stg + swz(offa + 16u * j, mask));
          yb[j] = ld_shared_u32_nc(stg + swz(offb + 16u * j, mask));
        }
      }
      float2 scq[NJ], shq[NJ];
      if (AM >= 0) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          scq[j] = *reinterpret_cast<const float2*>(cs + 8 * j + 2 * t0);
          shq[j] = *reinterpret_cast<const float2*>(cs + 64 + 8 * j + 2 * t0);
        }
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        uint32_t ga = va ? pack_bf16x2(__uint_as_float(v[4 * j]) + __uint_as_float(ra[j] << 16),
                                       __uint_as_float(v[4 * j + 1]) + __uint_as_float(ra[j] & 0xffff0000u)) : 0u;
        uint32_t gb = vb ? pack_bf16x2(__uint_as_float(v[4 * j + 2]) + __uint_as_float(rb[j] << 16),
                                       __uint_as_float(v[4 * j + 3]) + __uint_as_float(rb[j] & 0xffff0000u)) : 0u;
        if (AM == 1 && P.bwd_mask) {      // ReLU backward of the producing convolution: zero where its output is not positive
          ga &= (__uint_as_float(ya[j] << 16) > 0.f ? 0x0000ffffu : 0u) | (__uint_as_float(ya[j] & 0xffff0000u) > 0.f ? 0xffff0000u : 0u);
          gb &= (__uint_as_float(yb[j] << 16) > 0.f ? 0x0000ffffu : 0u) | (__uint_as_float(yb[j] & 0xffff0000u) > 0.f ? 0xffff0000u : 0u);
        }
        st_shared_u32(stg + swz(offa + 16u * j, mask), ga);
        st_shared_u32(stg + swz(offb + 16u * j, mask), gb);
        if (AM >= 0) {
          const float scv[2] = {scq[j].x, scq[j].y}, shv[2] = {shq[j].x, shq[j].y};
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const uint32_t gw = rr ? gb : ga, yw = rr ? yb[j] : ya[j];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float g = e ? __uint_as_float(gw & 0xffff0000u) : __uint_as_float(gw << 16);
              const float yv = e ? __uint_as_float(yw & 0xffff0000u) : __uint_as_float(yw << 16);
              const float tt = fmaf(yv, scv[e], shv[e]);
              float d = 1.f;
              if (AM == 1) d = tt > 0.f ? 1.f : 0.f;
              if (AM == 2) d = tt >= 0.f ? 1.f : P.bwd_alpha;
              const float gp = __fmul_rn(g, d);          // the same rounded product as the dx pass (bn_bwd_dx_part8_kernel)
              S[2 * j + e] += gp;
              S[2 * NJ + 2 * j + e] = fmaf(gp, yv, S[2 * NJ + 2 * j + e]);
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[b]));   // the accumulator is free as soon as it is in registers
    fence_proxy_async();                                         // generic-proxy writes -> visible to the bulk store
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    if (leader) {
      tma_store_4d(&P.omap, stg, nb0, tw * 8, th * 16 * P.mt, n);
      tma_store_commit();
      if (grp == 0) dbg_mark(P, 2, it >> 1, 2);
      if (AM >= 0 && tile + step_tiles < total_tiles) {
        // as soon as the store has READ the buffer, the yb tile of this group's next output tile is loaded into it
        const int tn = tile + step_tiles;
        const int tw2 = tn % P.tiles_w, t22 = tn / P.tiles_w, th2 = t22 % P.tiles_h, n2 = t22 / P.tiles_h;
        tma_store_wait_read<0>();
        mbar_expect_tx(ybar, P.stg_bytes);
        tma_load_4d(stg, &P.rmap, ybar, nb0, tw2 * 8, th2 * 16 * P.mt, n2);
      }
    }
  }
  if (leader) tma_store_wait<0>();
  if (AM >= 0 && !P.bwd_mask) {
    // transposed sum over the 8 lanes that share T%4 (lane bits 4, 3, 2): each step halves the values a lane carries; the lane
    // ends with the V/8 values i0 .. i0 + V/8 - 1, i0 = (V/2) bit4 + (V/4) bit3 + (V/8) bit2, of S = [sum][j][e]
#pragma unroll
    for (int step = 0; step < 3; ++step) {
      const int h = V >> (step + 1), lbit = 16 >> step;
      const bool up = (lane & lbit) != 0;
#pragma unroll
      for (int i = 0; i < h; ++i) {
        const float keep = up ? S[i + h] : S[i];
        const float send = up ? S[i] : S[i + h];
        S[i] = keep + __shfl_xor_sync(0xffffffffu, send, lbit);
      }
    }
    constexpr int KV = V / 8;
    asm volatile("bar.sync 3, 256;" ::: "memory");   // both groups: all stores have left the staging buffers, buffer 0 is the scratch of the final sum
    const int gwarp = grp * 4 + q;                   // [8 warps][KV][32] floats <= 4 KB
#pragma unroll
    for (int k = 0; k < KV; ++k) red_s[(gwarp * KV + k) * 32 + lane] = S[k];
    asm volatile("bar.sync 3, 256;" ::: "memory");
    const int etid = gwarp * 32 + lane;
    if (etid < KV * 32) {
      const int k = etid >> 5, l = etid & 31;
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += red_s[(r * KV + k) * 32 + l];
      const int i = (V / 2) * ((l >> 4) & 1) + (V / 4) * ((l >> 3) & 1) + (V / 8) * ((l >> 2) & 1) + k;   // index into [sum][j][e]
      const int which = i / (2 * NJ), j = (i % (2 * NJ)) >> 1, e = i & 1;
      P.bn_partials[(size_t)blockIdx.x * 2 * P.cout_total + (size_t)which * P.cout_total + nb0 + 8 * j + 2 * (l & 3) + e] = t;
    }
  }
}

// Issues the MMAs of one pipeline stage (all taps of one channel chunk).  MT and NBK (= chunk/16) are
// compile-time so the body is straight-line: one descriptor add per operand per tcgen05.mma.
// NT > 0 additionally fixes the tap count, so every per-tap descriptor is a constant-bank load at a static offset and the
// whole stage is one straight line of tcgen05.mma (the issue thread must stay well under the ~48 cycles a
// 128 x 64 x 16 MMA occupies the tensor pipe: probes/umma_probe.cu "t2_*").
template <int MT, int NBK, int NT>
__device__ __forceinline__ void issue_stage(const UmmaConvParams& P, int t0, int n_taps, uint32_t sa16, uint32_t b_lo, uint32_t b_step,
                                            uint64_t b_hi, uint32_t acc0, uint32_t nb, uint32_t idesc, bool not_first_chunk) {
  const int nt = NT > 0 ? NT : n_taps;
#pragma unroll
  for (int t = 0; t < nt; ++t) {
    const int tt = NT > 0 ? t : t0 + t;      // (t0 != 0 only in split-source mode, which uses the generic instance)
    const uint64_t ad = P.tap_adesc[tt] + (uint64_t)sa16;
    const uint32_t ms = P.tap_ms16[tt];
    const uint32_t acc_rest = (not_first_chunk || P.tap_first[tt] == 0u) ? 1u : 0u;
    const uint32_t acc_t = acc0 + P.tap_acc[tt];
#pragma unroll
    for (int k16 = 0; k16 < NBK; ++k16) {
      const uint64_t bd = b_hi | (uint64_t)(b_lo + (uint32_t)t * b_step + 2u * k16);
      const uint32_t accf = k16 == 0 ? acc_rest : 1u;
#pragma unroll
      for (int m = 0; m < MT; ++m) umma_f16(acc_t + (uint32_t)m * nb, ad + (uint64_t)((uint32_t)m * ms + 2u * k16), bd, idesc, accf);
    }
  }
}

template <bool BWD>
__global__ void __launch_bounds__(CONV_THREADS, 1) umma_conv_kernel(const __grid_constant__ UmmaConvParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a_full[MAX_STAGES], bar_a_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[MAX_ACC], bar_acc_empty[MAX_ACC], bar_w, bar_wf[2], bar_we[2];
  __shared__ __align__(8) uint64_t bar_res_full[MAX_STAGES], bar_res_empty[MAX_STAGES];   // BatchNorm phase: residual tiles
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[256];
  __shared__ __align__(16) float bnp_s[3 * 64];   // BatchNorm phase: scale | shift | PReLU slope of the N block

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (P.dbg && tid == 0 && blockIdx.y == 0) {   // kernel-lifetime marks: [CTA][0] entry clock, [1] exit clock, [2]/[3] globaltimer ns
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    P.dbg[256 + blockIdx.x * 4 + 0] = clock64();
    P.dbg[256 + blockIdx.x * 4 + 2] = (long long)gt;
  }
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base;
  const uint32_t stage_base = base + (P.resident ? P.w_res_bytes : (P.kouter ? 2u * P.wslot_bytes : 0u));
  const int nb0 = blockIdx.y * P.nb;
  const int total_tiles = P.n_img * P.tiles_h * P.tiles_w;

  if (tid == 0) {
    for (int s = 0; s < P.n_stages; ++s) {
      mbar_init(smem_u32(&bar_a_full[s]), 1);
      mbar_init(smem_u32(&bar_a_empty[s]), 1);
    }
    for (int b = 0; b < MAX_ACC; ++b) {
      mbar_init(smem_u32(&bar_acc_full[b]), 1);
      mbar_init(smem_u32(&bar_acc_empty[b]), 4);
    }
    for (int b = 0; b < MAX_STAGES; ++b) {
      mbar_init(smem_u32(&bar_res_full[b]), 1);
      mbar_init(smem_u32(&bar_res_empty[b]), 1);
    }
    mbar_init(smem_u32(&bar_w), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_wf[b]), 1);
      mbar_init(smem_u32(&bar_we[b]), 1);
    }
    fence_mbar_init();
    if (P.dbg && blockIdx.x == 0 && blockIdx.y == 0) P.dbg[192] = clock64();
    // the resident weights start loading before the CTA-wide sync, under the TMEM allocation of warp 1
    // (Tried in round 2 and reverted, job r2_04: barrier initialisation spread over the lanes of warp 0, the producer warp only
    // ARRIVING at the CTA-wide barrier, and the first tile's halo issued BEFORE the weights.  The first MMA moved ~800 cycles
    // LATER (the weight loads queue behind the halo's first-use descriptor fetch), the conv family ran 6 % slower inside the
    // step (348 vs 369 TFLOP/s) and the step 7.64 vs 7.42 ms.)
    for (int s = 0; s < P.n_src; ++s) tma_prefetch_desc(&P.src[s]);
    tma_prefetch_desc(&P.wmap);
    if (P.tstore || P.d2s_ts) tma_prefetch_desc(&P.omap);
    if (P.dbg && blockIdx.x == 0 && blockIdx.y == 0) P.dbg[193] = clock64();
    if (P.resident) {
      const uint32_t bw = smem_u32(&bar_w);
      mbar_expect_tx(bw, P.w_res_tx);
      for (int b = 0; b < P.n_taps * P.n_chunks; ++b) {
        // resident block b = tap_slot * n_chunks + chunk; source rows come from the tap's weight index
        int t = b / P.n_chunks, kc = b - t * P.n_chunks;
        tma_load_2d(w_base + (uint32_t)b * P.w_block_bytes, &P.wmap, bw, 0,
                    (P.tap_w[t] * P.n_chunks + kc) * P.cout_total + nb0);
      }
    }
    if (P.dbg && blockIdx.x == 0 && blockIdx.y == 0) P.dbg[194] = clock64();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
    if (P.dbg && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0) P.dbg[195] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (P.dbg && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) { P.dbg[196] = clock64(); P.dbg[197] = P.dbg[256]; }
  pdl_wait();   // everything above (barriers, TMEM, the constant weights) overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // Pipeline slots form one ring per issuing warp (slot s belongs to issuer s % n_issuers), so every mbarrier has
      // exactly one producer/consumer pair walking consecutive phases.
      const int n_iss = P.nbuf_shift == 2 ? 2 : 1, half = P.n_stages / n_iss;
      int it = 0;
      if (P.kouter) {
        uint32_t wj = 0, aj = 0;     // weight-slot and halo-stage sequence numbers
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          const int tw = tile % P.tiles_w, t2 = tile / P.tiles_w, th = t2 % P.tiles_h, n = t2 / P.tiles_h;
          const int h0 = th * 16 * P.mt, w0 = tw * 8;
          for (int kc = 0; kc < P.n_chunks; ++kc) {
            const uint32_t ws = wj & 1u, wfull = smem_u32(&bar_wf[ws]);
            mbar_wait(smem_u32(&bar_we[ws]), ((wj >> 1) & 1u) ^ 1u);
            mbar_expect_tx(wfull, P.wslot_tx);
            for (int t = 0; t < P.n_taps; ++t)
              tma_load_2d(w_base + ws * P.wslot_bytes + (uint32_t)t * P.w_block_bytes, &P.wmap, wfull, 0,
                          (P.tap_w[t] * P.n_chunks + kc) * P.cout_total + nb0);
            ++wj;
            for (int m = 0; m < P.mt; ++m, ++aj) {
              const int stage = (int)(aj % (uint32_t)P.n_stages);
              const uint32_t full = smem_u32(&bar_a_full[stage]);
              mbar_wait(smem_u32(&bar_a_empty[stage]), ((aj / (uint32_t)P.n_stages) & 1u) ^ 1u);
              mbar_expect_tx(full, P.stage_tx);
              const uint32_t sa = stage_base + (uint32_t)stage * P.stage_bytes;
              for (int s = 0; s < P.n_src; ++s)
                tma_load_4d(sa + P.src_off[s], &P.src[s], full, kc * P.kc, w0 + P.src_w0[s], h0 + 16 * m + P.src_h0[s], n);
            }
          }
        }
      } else
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        int tw = tile % P.tiles_w;
        int t2 = tile / P.tiles_w;
        int th = t2 % P.tiles_h;
        int n = t2 / P.tiles_h;
        const int h0 = th * 16 * P.mt, w0 = tw * 8;
        const int pit = it;
        const int wi = it % n_iss;
        const int n_steps = P.split ? P.n_chunks * P.n_src : P.n_chunks;
        const uint32_t j0 = (uint32_t)(it / n_iss) * (uint32_t)n_steps;
        dbg_mark(P, 0, pit, 0);
        for (int step = 0; step < n_steps; ++step) {
          const int kc = P.split ? step / P.n_src : step;
          const int so = P.split ? step - kc * P.n_src : 0;
          const uint32_t j = j0 + (uint32_t)step;
          const int stage = (int)(j % (uint32_t)half) * n_iss + wi;
          const uint32_t phase = (j / (uint32_t)half) & 1u;
          const uint32_t full = smem_u32(&bar_a_full[stage]);
          mbar_wait(smem_u32(&bar_a_empty[stage]), phase ^ 1u);
          if (step == 0) dbg_mark(P, 0, pit, 1);
          if ((P.dbg_flags & 4) && pit >= P.n_stages) { mbar_arrive(full); continue; }
          const uint32_t sa = stage_base + (uint32_t)stage * P.stage_bytes;
          if (P.split) {
            mbar_expect_tx(full, P.src_tx[so]);
            tma_load_4d(sa, &P.src[so], full, kc * P.kc, w0 + P.src_w0[so], h0 + P.src_h0[so], n);
            for (int t = 0; t < P.src_ntaps[so]; ++t)
              tma_load_2d(sa + P.w_stage_off + (uint32_t)t * P.w_block_bytes, &P.wmap, full, 0,
                          (P.tap_w[P.src_tap0[so] + t] * P.n_chunks + kc) * P.cout_total + nb0);
            continue;
          }
          mbar_expect_tx(full, P.stage_tx);
          for (int s = 0; s < P.n_src; ++s)
            tma_load_4d(sa + P.src_off[s], &P.src[s], full, kc * P.kc, w0 + P.src_w0[s], h0 + P.src_h0[s], n);
          if (!P.resident)
            for (int t = 0; t < P.n_taps; ++t)
              tma_load_2d(sa + P.w_stage_off + (uint32_t)t * P.w_block_bytes, &P.wmap, full, 0,
                          (P.tap_w[t] * P.n_chunks + kc) * P.cout_total + nb0);
        }
        dbg_mark(P, 0, pit, 2);
      }
      if (P.bnp && P.bnp_res) {
        // BatchNorm phase, pass 2: the skip-connection tiles travel through the (now idle) halo stages.  All MMAs of this
        // CTA must have completed first: the last tile of each issuing warp has been committed to its accumulator barrier.
        const int n_local = it;
        if (n_local >= 1) mbar_wait(smem_u32(&bar_acc_full[n_local - 1]), 0);
        if (n_local >= 2) mbar_wait(smem_u32(&bar_acc_full[n_local - 2]), 0);
        int j = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
          const int tw = tile % P.tiles_w, t2 = tile / P.tiles_w, th = t2 % P.tiles_h, n = t2 / P.tiles_h;
          const int slot = j % P.n_stages;
          const uint32_t full = smem_u32(&bar_res_full[slot]);
          mbar_wait(smem_u32(&bar_res_empty[slot]), (((uint32_t)(j / P.n_stages)) & 1u) ^ 1u);
          mbar_expect_tx(full, P.stg_bytes);
          tma_load_4d(stage_base + (uint32_t)slot * P.stage_bytes, &P.rmap, full, nb0, tw * 8, th * 16 * P.mt, n);
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // ------------------------------------------------------------------ MMA issuer(s)
    // The WHOLE warp runs this loop with warp-uniform values (kernel parameters, loop counters) and only
    // the tcgen05 instructions sit under elect_one(): UTCHMMA takes its descriptors from UNIFORM registers,
    // and when the loop ran under `if (lane == 0)` every operand went through R2UR with a scoreboard wait
    // (~200 cycles per MMA measured with tools/conv_timeline.py; the tensor pipe needs one per 32).
    {
      const int n_taps = P.n_taps, n_chunks = P.n_chunks, mt = P.mt, nbk = P.kc / 16, n_stages = P.n_stages;
      const uint32_t nb = (uint32_t)P.nb, idesc = P.idesc;
      const uint32_t w_sbo = 8u * (uint32_t)P.kc * 2u;
      const uint64_t b_hi = make_smem_desc_hi(w_sbo, P.layout) << 32;
      const uint32_t lbo16 = 1u << 16;  // LBO field (ignored for swizzled K-major), 16 bytes
      const uint32_t wblk16 = P.w_block_bytes >> 4;
      const uint32_t wres16 = (w_base >> 4) | lbo16;
      const uint32_t wstage16 = (P.w_stage_off >> 4) | lbo16;
      const uint32_t stage16 = P.stage_bytes >> 4, sbase16 = stage_base >> 4;
      const bool resident = P.resident != 0;
      if (resident) {
        mbar_wait(smem_u32(&bar_w), 0);
        tc_fence_after();
      }
      // With four accumulator buffers two warps issue alternate tiles, so the tensor pipe always has the other
      // warp's MMAs queued while one warp walks its barriers between tiles (a single issuer let the pipe drain and
      // refill every 128-pixel tile: ~500 of 2700 cycles on the 64-channel layers, tools/conv_timeline.py).
      const int n_issuers = P.nbuf_shift == 2 ? 2 : 1;
      const int me = warp == 1 ? 0 : 1;
      if (P.kouter) {
        if (me == 0) {
          uint32_t wj = 0, aj = 0;
          const uint32_t wslot16 = P.wslot_bytes >> 4;
          int it = 0;
          for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            mbar_wait(smem_u32(&bar_acc_empty[b]), ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            for (int kc = 0; kc < n_chunks; ++kc, ++wj) {
              const uint32_t ws = wj & 1u;
              mbar_wait(smem_u32(&bar_wf[ws]), (wj >> 1) & 1u);
              const uint32_t b_lo = wres16 + ws * wslot16;
              for (int m = 0; m < mt; ++m, ++aj) {
                const int stage = (int)(aj % (uint32_t)n_stages);
                mbar_wait(smem_u32(&bar_a_full[stage]), (aj / (uint32_t)n_stages) & 1u);
                tc_fence_after();
                const uint32_t sa16 = sbase16 + (uint32_t)stage * stage16;
                const uint32_t acc0 = tmem + (uint32_t)(b * mt + m) * nb;
                if (elect_one()) {
#define DG_ISSUE_K(NBK_) \
  { if (n_taps == 9) issue_stage<1, NBK_, 9>(P, 0, 9, sa16, b_lo, wblk16, b_hi, acc0, nb, idesc, kc != 0); \
    else if (n_taps == 4) issue_stage<1, NBK_, 4>(P, 0, 4, sa16, b_lo, wblk16, b_hi, acc0, nb, idesc, kc != 0); \
    else issue_stage<1, NBK_, 0>(P, 0, n_taps, sa16, b_lo, wblk16, b_hi, acc0, nb, idesc, kc != 0); }
                  if (nbk == 4) DG_ISSUE_K(4) else if (nbk == 2) DG_ISSUE_K(2) else DG_ISSUE_K(1)
#undef DG_ISSUE_K
                  umma_commit(smem_u32(&bar_a_empty[stage]));
                  if (m == mt - 1) umma_commit(smem_u32(&bar_we[ws]));
                  if (m == mt - 1 && kc == n_chunks - 1) umma_commit(smem_u32(&bar_acc_full[b]));
                }
                __syncwarp();
              }
            }
          }
        }
      } else
      if (me < n_issuers)
      for (int it = me, tile = blockIdx.x + me * (int)gridDim.x; tile < total_tiles; tile += n_issuers * (int)gridDim.x, it += n_issuers) {
        const int b = P.bnp ? it : (it & ((1 << P.nbuf_shift) - 1));   // BatchNorm phase: one accumulator per tile, never recycled
        const uint32_t acc_phase = P.bnp ? 0u : ((uint32_t)(it >> P.nbuf_shift) & 1u);
        const uint32_t half = (uint32_t)(n_stages / n_issuers);
        const int n_steps = P.split ? n_chunks * P.n_src : n_chunks;
        const uint32_t j0 = (uint32_t)(it / n_issuers) * (uint32_t)n_steps;   // this issuer's slot sequence number
        if (lane == 0 && me == 0) dbg_mark(P, 1, it, 0);
        if (!P.bnp) mbar_wait(smem_u32(&bar_acc_empty[b]), acc_phase ^ 1u);
        tc_fence_after();
        if (lane == 0 && me == 0) dbg_mark(P, 1, it, 1);
        const uint32_t acc0 = tmem + (uint32_t)(b * P.n_phase * mt) * nb;
        for (int step = 0; step < n_steps; ++step) {
          const int kc = P.split ? step / P.n_src : step;
          const int so = P.split ? step - kc * P.n_src : 0;
          const uint32_t j = j0 + (uint32_t)step;
          const int stage = (int)(j % half) * n_issuers + me;
          const uint32_t phase = (j / half) & 1u;
          mbar_wait(smem_u32(&bar_a_full[stage]), phase);
          tc_fence_after();
          if (step == 0 && lane == 0 && me == 0) dbg_mark(P, 1, it, 2);
          const uint32_t sa16 = sbase16 + (uint32_t)stage * stage16;
          if (elect_one()) {
            const uint32_t b_lo = resident ? wres16 + (uint32_t)kc * wblk16 : sa16 + wstage16;
            const uint32_t b_step = resident ? (uint32_t)n_chunks * wblk16 : wblk16;
            const int t0 = P.split ? P.src_tap0[so] : 0, ntp = P.split ? P.src_ntaps[so] : n_taps;
#define DG_ISSUE(MT_, NBK_, NT_) issue_stage<MT_, NBK_, NT_>(P, t0, ntp, sa16, b_lo, b_step, b_hi, acc0, nb, idesc, step != 0)
#define DG_ISSUE_NT(MT_, NBK_) \
  { if (P.split) DG_ISSUE(MT_, NBK_, 0); else if (n_taps == 9) DG_ISSUE(MT_, NBK_, 9); else if (n_taps == 4) DG_ISSUE(MT_, NBK_, 4); else if (n_taps == 1) DG_ISSUE(MT_, NBK_, 1); else DG_ISSUE(MT_, NBK_, 0); }
            if (mt == 1) {
              if (nbk == 4) DG_ISSUE_NT(1, 4) else if (nbk == 2) DG_ISSUE_NT(1, 2) else DG_ISSUE_NT(1, 1)
            } else {
              if (nbk == 4) DG_ISSUE_NT(2, 4) else if (nbk == 2) DG_ISSUE_NT(2, 2) else DG_ISSUE_NT(2, 1)
            }
#undef DG_ISSUE_NT
#undef DG_ISSUE
            umma_commit(smem_u32(&bar_a_empty[stage]));
            if (step == n_steps - 1) umma_commit(smem_u32(&bar_acc_full[b]));
          }
          __syncwarp();
        }
        if (lane == 0 && me == 0) dbg_mark(P, 1, it, 3);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps)
    // activation / output type are resolved ONCE per kernel (warp-uniform switch) so the per-element code is
    // straight-line: TMEM -> registers, + bias (from shared memory), activation, pack, 16-byte stores.
    const int q = warp & 3;  // TMEM lane quarter this warp may access (any four consecutive warps cover the four quarters)
    const int grp = warp >= 7 ? 1 : 0;
    const int etid = grp * 128 + (warp - (grp ? 7 : 2)) * 32 + lane;   // 0..255 over both epilogue groups
    if (P.bias)
      for (int i = etid; i < P.nb; i += 256) bias_s[i] = __ldg(P.bias + nb0 + i);
    if (!BWD && P.epi_prelu)
      for (int i = etid; i < P.nb; i += 256) bnp_s[i] = __ldg(P.epi_prelu + nb0 + i);      // PReLU slopes of the staged epilogue (nb <= 64)
    if (!BWD && P.d2s_prelu)
      for (int i = etid; i < P.d2s_cq; i += 256) bnp_s[i] = __ldg(P.d2s_prelu + i);      // PReLU slopes of the depth_to_space store (d2s_cq <= 192)
    if (BWD && P.bwd_am >= 0)
      for (int i = etid; i < P.nb; i += 256) {
        if (P.bwd_mask) { bnp_s[i] = 1.f; bnp_s[64 + i] = 0.f; }
        else { bnp_s[i] = __ldg(P.bwd_scale + nb0 + i); bnp_s[64 + i] = __ldg(P.bwd_shift + nb0 + i); }
      }
    asm volatile("bar.sync 3, 256;" ::: "memory");  // epilogue warps only
    const float* bs = P.bias ? bias_s : nullptr;
    if (BWD) {
      float* red = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + P.stg_off);
#define DG_EPI_BWD(AM_) \
  { if (P.nb == 64) epilogue_role_bwd<AM_, 64>(P, tmem, base + P.stg_off, q, lane, nb0, total_tiles, bar_acc_full, bar_acc_empty, red, grp, bnp_s, bar_res_full); \
    else epilogue_role_bwd<AM_, 32>(P, tmem, base + P.stg_off, q, lane, nb0, total_tiles, bar_acc_full, bar_acc_empty, red, grp, bnp_s, bar_res_full); }
      switch (P.bwd_am) {
        case 0: DG_EPI_BWD(0) break;
        case 1: DG_EPI_BWD(1) break;
        case 2: DG_EPI_BWD(2) break;
        default: DG_EPI_BWD(-1) break;
      }
#undef DG_EPI_BWD
    } else {
#define DG_EPI(ACT)                                                                                     \
  if (P.d2s_ts) epilogue_role_d2s_ts(P, tmem, base + P.stg_off, q, lane, total_tiles, bs, bar_acc_full, bar_acc_empty, grp, P.d2s_prelu ? bnp_s : nullptr); \
  else if (P.tstore && (P.epi_res || P.epi_prelu)) epilogue_role_ts<ACT, true>(P, tmem, base + P.stg_off, q, lane, nb0, total_tiles, bs, bar_acc_full, bar_acc_empty, \
                                      reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + P.stg_off), grp, bnp_s, \
                                      bar_res_full, bar_res_empty, stage_base); \
  else if (P.tstore) epilogue_role_ts<ACT>(P, tmem, base + P.stg_off, q, lane, nb0, total_tiles, bs, bar_acc_full, bar_acc_empty, \
                                      reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + P.stg_off), grp, bnp_s, \
                                      bar_res_full, bar_res_empty, stage_base); \
  else if (P.out_f32) epilogue_role<ACT, true>(P, tmem, q, lane, nb0, total_tiles, bs, bar_acc_full, bar_acc_empty, grp); \
  else epilogue_role<ACT, false>(P, tmem, q, lane, nb0, total_tiles, bs, bar_acc_full, bar_acc_empty, grp, P.d2s_prelu ? bnp_s : nullptr);
    switch (P.act) {
      case DG_ACT_RELU: DG_EPI(DG_ACT_RELU) break;
      case DG_ACT_LRELU: DG_EPI(DG_ACT_LRELU) break;
      case DG_ACT_TANH: DG_EPI(DG_ACT_TANH) break;
      case DG_ACT_SIGMOID: DG_EPI(DG_ACT_SIGMOID) break;
      default: DG_EPI(DG_ACT_NONE) break;
    }
#undef DG_EPI
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
  if (P.dbg && tid == 0 && blockIdx.y == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    P.dbg[256 + blockIdx.x * 4 + 1] = clock64();
    P.dbg[256 + blockIdx.x * 4 + 3] = (long long)gt;
  }
}

// ------------------------------------------------------------------ weight packing
// mode 0 (forward): dst[((t*nch+kc)*Cout + o)*KC + j] = w[t][kc*KC + j][o]      (KC chunks over Cin)
// mode 1 (dgrad)  : dst[((t*nch+kc)*Cin  + c)*KC + j] = w[t][c][kc*KC + j]      (KC chunks over Cout)
// cin/cout are the PACKED (possibly zero-padded) dims, cin_s/cout_s the dims of the fp32 source.
// Physically padded activations (dg_umma_pack_weights_seg): the packed input-channel axis may consist of TWO zero-padded
// segments (a U-Net concat of two padded tensors, autoencoder.py:135): channels [0, seg_phys) hold seg_log source channels,
// the rest holds the remaining cin_s - seg_log.  seg_phys == 0: one segment.  Returns the source channel or -1 (zero).
__device__ __forceinline__ int seg_src_channel(int ci, int cin_s, int seg_log, int seg_phys) {
  if (seg_phys == 0) return ci < cin_s ? ci : -1;
  if (ci < seg_phys) return ci < seg_log ? ci : -1;
  const int r = ci - seg_phys;
  return r < cin_s - seg_log ? seg_log + r : -1;
}

__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int taps, int cin,
                                    int cout, int kc, int mode, int cin_s, int cout_s, int seg_log, int seg_phys) {
  long total = (long)taps * cin * cout;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int rows = mode == 0 ? cout : cin;  // rows of a block
    int kdim = mode == 0 ? cin : cout;  // contraction length
    int nch = kdim / kc;
    int j = (int)(i % kc);
    long r1 = i / kc;
    int row = (int)(r1 % rows);
    long r2 = r1 / rows;
    int ch = (int)(r2 % nch);
    int t = (int)(r2 / nch);
    int k = ch * kc + j;
    const int ci = seg_src_channel(mode == 0 ? k : row, cin_s, seg_log, seg_phys), co = mode == 0 ? row : k;
    float v = (ci >= 0 && co < cout_s) ? w[((long)t * cin_s + ci) * cout_s + co] : 0.f;
    dst[i] = __float2bfloat16(v);
  }
}

static long long* g_dbg_timeline = nullptr;
static int g_dbg_flags = getenv("DG_CONV_DBG_FLAGS") ? atoi(getenv("DG_CONV_DBG_FLAGS")) : 0;   // see UmmaConvParams::dbg_flags

struct PackEntry {   // mirrored by denoise_gan_b200/params.py (48 bytes)
  const float* src;
  __nv_bfloat16* dst;
  int taps, cin, cout, kc, mode, cin_src;   // cin/cout: packed (padded) dims; *_src: dims of the fp32 source (0 = same)
  int cout_src, seg;   // seg: seg_log | seg_phys << 16 (two-segment input-channel axis, see seg_src_channel), 0 = one segment
};

// all kernels of a network in ONE launch: blockIdx.y selects the table entry
__global__ void pack_weights_batch_kernel(const PackEntry* __restrict__ table) {
  const PackEntry E = table[blockIdx.y];
  const int rows = E.mode == 0 ? E.cout : E.cin, kdim = E.mode == 0 ? E.cin : E.cout, nch = kdim / E.kc;
  const unsigned total = (unsigned)E.taps * E.cin * E.cout;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned j = i % E.kc, r1 = i / E.kc;
    unsigned row = r1 % rows, r2 = r1 / rows;
    unsigned ch = r2 % nch, t = r2 / nch;
    unsigned k = ch * E.kc + j;
    const unsigned co = E.mode == 0 ? row : k;
    const unsigned cis = E.cin_src ? E.cin_src : E.cin, cos = E.cout_src ? E.cout_src : E.cout;
    const int ci = seg_src_channel((int)(E.mode == 0 ? k : row), (int)cis, E.seg & 0xffff, E.seg >> 16);
    float v = (ci >= 0 && co < cos) ? E.src[((long)t * cis + ci) * cos + co] : 0.f;
    E.dst[i] = __float2bfloat16(v);
  }
}

inline int kc_for(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : 16); }

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map(dg_ctx* ctx, CUtensorMap* m, void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_b,
               const uint32_t* box, int kc) {
  CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  uint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = ((EncodeFn)ctx->encode_tiled)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, (const cuuint64_t*)dims,
                                             (const cuuint64_t*)strides_b, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DG_FAIL("cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// A lattice view of an NHWC tensor: element (n,h,w,c) of the view is element
// (n, h*step+h_first, w*step+w_first, c) of the tensor.
struct Lattice {
  int step, h_first, w_first;
};

struct TapSpec {
  int src, dh, dw, widx;
  int phase = 0;   // output parity phase the tap contributes to (fused stride-2 dgrad), 0 otherwise
};

// Builds the launch description and runs the kernel.
// BatchNorm phase of a fused launch (UmmaConvParams::bnp)
struct BnPhase {
  const dg_bn_fused* bn;
  int act;
  float alpha;
  const float* prelu;
  const dg_tensor* res;    // skip connection added after the activation, or nullptr
  const dg_tensor* out2;   // act(BN(y)) (+ res)
};

// BatchNorm-backward epilogue of an input-gradient launch (UmmaConvParams::bwd_*)
struct BwdEpi {
  const dg_tensor* res;        // skip-connection gradient added to the result, or nullptr
  const dg_bn_bwd_stats* bn;   // statistics of the BatchNorm in front of the convolution, or nullptr
  const dg_tensor* relu_y = nullptr;   // output of the ReLU convolution in front of this one: store g * (relu_y > 0), no statistics
};

// residual / PReLU of the staged forward epilogue (UmmaConvParams::epi_res / epi_prelu)
struct EpiExtra {
  const dg_tensor* res;
  const float* prelu;
};

int launch_conv(dg_ctx* ctx, const char* name, const dg_tensor* in, const Lattice* src_lat, int n_src,
                const TapSpec* taps_in, int n_taps, const void* w_packed, int w_rows_per_block /*cout_total*/,
                const dg_tensor* out, Lattice out_lat, const float* bias, int act, float alpha, cudaStream_t st, bool dry = false,
                float* bn_partials = nullptr, int* bn_blocks = nullptr, int n_phase = 1, const Lattice* phase_lat = nullptr,
                const dg_bn_fused* bn_fin = nullptr, const BnPhase* bnp = nullptr, bool bnp_query = false, const BwdEpi* bwd = nullptr,
                bool bwd_query = false, int out_cvalid = 0, int d2s_cq = 0, const float* d2s_prelu = nullptr, const EpiExtra* ex = nullptr) {
  DG_REQUIRE(in->dtype == DG_BF16, "%s: tensor-core path needs bf16 input", name);
  DG_REQUIRE(d2s_cq == 0 || (d2s_cq % 32 == 0 && d2s_cq <= 192 && out->dtype == DG_BF16 && n_phase == 1 && !bn_partials && !bnp && !bwd && act == DG_ACT_NONE),
             "%s: the depth_to_space store needs a bf16 output with a multiple of 32 (<= 192) channels per sub-pixel block and no other epilogue", name);
  DG_REQUIRE(out_cvalid == 0 || (out->dtype == DG_F32 && out->c == 16 && out_cvalid < 16 && n_phase == 1 && !bn_partials && !bnp && !bwd),
             "%s: a narrow store needs an fp32 output of fewer than 16 channels", name);
  DG_REQUIRE(n_phase == 1 || (n_phase == 4 && phase_lat && n_src == 1), "%s: bad output-phase description", name);
  DG_REQUIRE(in->c % 16 == 0 && out->c % 16 == 0, "%s: channels must be multiples of 16 (got %d -> %d)", name, in->c, out->c);
  DG_REQUIRE(in->cpitch % 8 == 0 && in->coff % 8 == 0 && ((uintptr_t)in->ptr % 16) == 0, "%s: input view not 16-byte aligned", name);
  const int out_esz = out->dtype == DG_F32 ? 4 : 2;
  DG_REQUIRE(out_cvalid > 0 || (((uintptr_t)out->ptr % 16) == 0 && (out->cpitch * out_esz) % 16 == 0 && (out->coff * out_esz) % 16 == 0),
             "%s: output view not 16-byte aligned", name);
  DG_REQUIRE(n_src >= 1 && n_src <= MAX_SRC && n_taps >= 1 && n_taps <= MAX_TAPS, "%s: too many sources/taps", name);
  DG_REQUIRE(act != DG_ACT_PRELU, "%s: PReLU is not a conv epilogue", name);

  // taps grouped by source (stable), so that a stage can carry the taps of one source only (split mode)
  TapSpec taps_sorted[MAX_TAPS];
  {
    int k = 0;
    for (int s = 0; s < n_src; ++s)
      for (int t = 0; t < n_taps; ++t)
        if (taps_in[t].src == s) taps_sorted[k++] = taps_in[t];
    DG_REQUIRE(k == n_taps, "%s: tap with an invalid source", name);
  }
  const TapSpec* taps = taps_sorted;

  UmmaConvParams P;
  memset(&P, 0, sizeof(P));
  const int kc = kc_for(in->c);
  const int n_chunks = in->c / kc;
  const int cout = out->c;
  P.kc = kc; P.n_chunks = n_chunks; P.n_src = n_src; P.n_taps = n_taps; P.cout_total = w_rows_per_block;
  P.layout = kc == 64 ? LAYOUT_SW128 : (kc == 32 ? LAYOUT_SW64 : LAYOUT_SW32);

  // output view extents
  const int out_h = (out->h - out_lat.h_first + out_lat.step - 1) / out_lat.step;
  const int out_w = (out->w - out_lat.w_first + out_lat.step - 1) / out_lat.step;
  DG_REQUIRE(out_h > 0 && out_w > 0, "%s: empty output view", name);
  P.out_h = out_h; P.out_w = out_w; P.n_img = out->n;

  // per-source halo extents from the taps
  int dh_min[MAX_SRC], dh_max[MAX_SRC], dw_min[MAX_SRC], dw_max[MAX_SRC];
  bool used[MAX_SRC] = {false, false, false, false};
  for (int t = 0; t < n_taps; ++t) {
    int s = taps[t].src;
    DG_REQUIRE(s >= 0 && s < n_src, "%s: bad tap source", name);
    if (!used[s]) { dh_min[s] = dh_max[s] = taps[t].dh; dw_min[s] = dw_max[s] = taps[t].dw; used[s] = true; }
    dh_min[s] = taps[t].dh < dh_min[s] ? taps[t].dh : dh_min[s];
    dh_max[s] = taps[t].dh > dh_max[s] ? taps[t].dh : dh_max[s];
    dw_min[s] = taps[t].dw < dw_min[s] ? taps[t].dw : dw_min[s];
    dw_max[s] = taps[t].dw > dw_max[s] ? taps[t].dw : dw_max[s];
  }
  for (int s = 0; s < n_src; ++s) DG_REQUIRE(used[s], "%s: unused source %d", name, s);

  // choose (MT, NB, resident) : prefer resident weights with the widest N block
  const uint32_t budget = SMEM_LIMIT - 4096;
  auto halo_bytes = [&](int mt) {
    uint32_t tot = 0;
    for (int s = 0; s < n_src; ++s) {
      uint32_t hb = (uint32_t)(16 * mt + dh_max[s] - dh_min[s]) * (uint32_t)(8 + dw_max[s] - dw_min[s]) * kc * 2;
      tot += (hb + 1023u) & ~1023u;
    }
    return tot;
  };
  int best_nb = 0, best_mt = 0, best_res = 0;
  const long tiles_mt2 = (long)in->n * ((out_h + 31) / 32) * ((out_w + 7) / 8);
  // A wide layer with a SHORT contraction (the 1x1 expand convolutions of Fast-SRGAN, fsrgan.py:134-147: K = 32, 192 outputs) does
  // almost no tensor work per byte it stores: with one 192-channel N block it runs the direct epilogue (a warp's 16-byte stores land
  // on 32 different lines, 436 us for 1.17 GB at 1080p) and cannot carry the BatchNorm statistics.  64-channel N blocks re-read the
  // small input from L2 but leave through the staged TMA store and keep the statistics / BatchNorm-backward epilogues available.
  static const char* no_smallk = getenv("DG_DEBUG_NO_SMALLK_NB64");   // experiments only
  const bool small_k = !no_smallk && (long)n_taps * in->c <= 64 && cout > 64 && cout % 64 == 0 && out->dtype == DG_BF16 && out_lat.step == 1 &&
                       n_phase == 1 && d2s_cq == 0 &&
                       // plain launches only: the statistics / BatchNorm-backward epilogues and their row-count queries keep the layout
                       // they were validated with (one partial row per CTA of a single N block)
                       !bn_partials && !bn_blocks && !bnp && !bnp_query && !bwd && !bwd_query && !bn_fin;
  for (int res = 1; res >= 0 && !best_nb; --res)
    for (int mt = 2; mt >= 1 && !best_nb; --mt)
      for (int nb = cout > 256 ? 256 : cout; nb >= 16 && !best_nb; nb -= 16) {
        if (cout % nb != 0 || 2 * n_phase * mt * nb > 512) continue;
        if (small_k && nb > 64) continue;
        uint32_t wblk = (uint32_t)nb * kc * 2;
        uint32_t wres = (uint32_t)n_taps * n_chunks * wblk;
        uint32_t stage = halo_bytes(mt) + (res ? 0 : (uint32_t)n_taps * wblk);
        // resident weights: the taller (32-row) tile halves the per-tile issue bubble, worth it with >= 3 stages in
        // flight, enough tiles to fill the SMs about three times, and no half-empty bottom tile row
        const int min_stages = (res && mt == 2) ? 3 : 2;
        static const char* dbg_mt = getenv("DG_DEBUG_MT");   // experiments only
        if (dbg_mt && res && nb == cout) { if (mt != atoi(dbg_mt)) continue; }
        else if (res && mt == 2 && (cout > 32 || n_src > 1 || tiles_mt2 < 3L * ctx->sm_count || out_h % 32 != 0)) continue;   // measured: only narrow stride-1 layers gain
        uint32_t need = (res ? ((wres + 1023u) & ~1023u) : 0) + min_stages * ((stage + 1023u) & ~1023u);
        if (need <= budget) { best_nb = nb; best_mt = mt; best_res = res; }
      }
  // Fallback for many-tap, many-channel layers (pix2pix 4x4 stride 2 with >= 128 channels): one SOURCE per pipeline stage
  // (its halo box + the weights of its taps), so a stage is 1/n_src of the halo and of the streamed weights.
  int split = 0;
  int src_ntaps[MAX_SRC] = {0, 0, 0, 0}, max_ntaps = 0;
  for (int t = 0; t < n_taps; ++t) ++src_ntaps[taps[t].src];
  for (int s = 0; s < n_src; ++s) max_ntaps = src_ntaps[s] > max_ntaps ? src_ntaps[s] : max_ntaps;
  auto max_halo = [&](int mt) {
    uint32_t m = 0;
    for (int s = 0; s < n_src; ++s) {
      uint32_t hb = (uint32_t)(16 * mt + dh_max[s] - dh_min[s]) * (uint32_t)(8 + dw_max[s] - dw_min[s]) * kc * 2;
      hb = (hb + 1023u) & ~1023u;
      m = hb > m ? hb : m;
    }
    return m;
  };
  if (!best_nb && n_src > 1) {
    for (int min_stages = 3; min_stages >= 2 && !best_nb; --min_stages)
      for (int nb = cout > 256 ? 256 : cout; nb >= 16 && !best_nb; nb -= 16)
        for (int mt = 2; mt >= 1 && !best_nb; --mt) {
          if (cout % nb != 0 || 2 * n_phase * mt * nb > 512) continue;
          uint32_t stage = max_halo(mt) + (uint32_t)max_ntaps * nb * kc * 2;
          if ((uint32_t)min_stages * ((stage + 1023u) & ~1023u) <= budget) { best_nb = nb; best_mt = mt; best_res = 0; split = 1; }
        }
  }
  DG_REQUIRE(best_nb > 0, "%s: no tile configuration fits shared memory (Cin=%d Cout=%d taps=%d)", name, in->c, cout, n_taps);
  auto r1k = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  // K-outer mode (streamed weights, several chunks): weight chunks go through a two-slot ring of their own and are reused
  // by `kouter` 16x8 sub-tiles; chosen when the predicted shared-memory fill traffic per CTA drops by >= 10 %
  int kouter = 0;
  {
    // Cost model (cycles per CTA, MMA issue + shared-memory fill, shallow pipelines make them roughly additive):
    // a 128 x nb x 16 MMA occupies the tensor pipe max(nb/2, 32 + nb/4, 44) cycles (probes/umma_probe.cu t2_*), and with
    // every SM pulling at once L2 delivers ~19 B/clk/SM (5.6 TB/s).  The streamed K-outer configuration competes with
    // whatever the search above picked -- including RESIDENT weights bought with a narrow N block (256->64 dgrad ran as
    // two 32-channel blocks: 44-cycle MMAs for half the work and the input read twice).
    static const char* dbg_no_ko = getenv("DG_DEBUG_NO_KOUTER");   // experiments only
    if (!split && n_chunks >= 2 && !dbg_no_ko && n_phase == 1) {
      auto cyc = [](int nb_) { int c = nb_ / 2 > 32 + nb_ / 4 ? nb_ / 2 : 32 + nb_ / 4; return (double)(c < 44 ? 44 : c); };
      auto rounds = [&](int mt_, int nb_) {
        long t = (long)in->n * ((out_h + 16 * mt_ - 1) / (16 * mt_)) * ((out_w + 7) / 8), c = ctx->sm_count / (cout / nb_);
        c = c < 1 ? 1 : (c > t ? t : c);
        return (double)((t + c - 1) / c);
      };
      const double kk = (double)n_chunks * n_taps * (kc / 16);
      const uint32_t h1 = halo_bytes(1);
      const double cur = rounds(best_mt, best_nb) * (best_mt * kk * cyc(best_nb) +
                                                     n_chunks * ((double)halo_bytes(best_mt) + (best_res ? 0.0 : (double)n_taps * best_nb * kc * 2)) / 19.0);
      double best_cost = (g_dbg_flags & 16) ? 1e30 : 0.85 * cur;   // flag 16 (tests): K-outer whenever it fits
      int k_nb = 0;
      for (int nbk = cout > 256 ? 256 : cout; nbk >= best_nb && nbk >= 16; nbk -= 16) {
        if (cout % nbk != 0) continue;
        const uint32_t wchunk = (uint32_t)n_taps * nbk * kc * 2;
        if (2 * r1k(wchunk) + 3 * r1k(h1) > budget) continue;
        for (int mtk = 4; mtk >= 2; mtk -= 2) {
          if (2 * mtk * nbk > 512) continue;
          // three halo stages + the weight ring overlap fill and issue: measured ~1.2 x max(issue, fill) (256->64 dgrad at 192^2)
          const double mma = mtk * kk * cyc(nbk), fill = n_chunks * ((double)wchunk + (double)mtk * h1) / 19.0;
          const double c = rounds(mtk, nbk) * 1.2 * (mma > fill ? mma : fill);
          if (c < best_cost) { best_cost = c; kouter = mtk; k_nb = nbk; }
        }
      }
      if (kouter) { best_mt = kouter; best_nb = k_nb; best_res = 0; }
    }
  }
  const int nb = best_nb, mt = best_mt;
  const int hmt = kouter ? 1 : mt;     // sub-tiles covered by ONE halo box
  P.nb = nb; P.mt = mt; P.resident = best_res; P.split = split; P.kouter = kouter;
  P.wslot_tx = (uint32_t)n_taps * nb * kc * 2;
  P.wslot_bytes = r1k(P.wslot_tx);
  const uint32_t ring_bytes = kouter ? 2u * P.wslot_bytes : 0u;

  // pipeline depth, and whether the staged (shared memory + TMA store) epilogue fits next to it
  const uint32_t stage_bytes_pre = split ? r1k(max_halo(mt) + (uint32_t)max_ntaps * nb * kc * 2)
                                         : r1k(halo_bytes(hmt) + ((best_res || kouter) ? 0u : (uint32_t)n_taps * nb * kc * 2));
  const uint32_t w_res_pre = best_res ? r1k((uint32_t)n_taps * n_chunks * nb * kc * 2) : 0u;
  int n_stages = (int)((budget - w_res_pre - ring_bytes) / stage_bytes_pre);
  if (n_stages > MAX_STAGES) n_stages = MAX_STAGES;
  DG_REQUIRE(n_stages >= 2, "%s: internal: fewer than 2 stages", name);
  static const char* dbg_no_ts = getenv("DG_DEBUG_NO_TSTORE");   // experiments only
  bool ts = !kouter && n_phase == 1 && (!dbg_no_ts || bn_partials || bn_blocks) && out->dtype == DG_BF16 && out_lat.step == 1 && (nb == 16 || nb == 32 || nb == 64) &&
            d2s_cq == 0;
  // depth_to_space store of a 32 -> 4 x 32 channel layer through the staging buffers (epilogue_role_d2s_ts): one 128-channel N block,
  // 16 x 8 tiles, dense output pixels (two of them are one 128-byte staged row)
  static const char* dbg_no_d2s_ts = getenv("DG_DEBUG_NO_D2S_TSTORE");   // experiments only
  bool d2s_ts = !dbg_no_d2s_ts && d2s_cq == 32 && nb == 128 && cout == 128 && mt == 1 && !kouter && !split && n_phase == 1 && out->dtype == DG_BF16 &&
                out->cpitch == 32 && out->coff == 0 && ((uintptr_t)out->ptr % 128) == 0 && !bn_partials && !bn_blocks && !bnp && !bwd;
  const uint32_t stg_bytes = (uint32_t)mt * 128u * (uint32_t)nb * 2u;
  if (d2s_ts) {
    const long room = (long)budget - (long)w_res_pre - 2L * (long)stg_bytes;
    int ns = room > 0 ? (int)(room / (long)stage_bytes_pre) : 0;
    if (ns > MAX_STAGES) ns = MAX_STAGES;
    if (ns >= 4 || (ns >= 2 && n_stages < 4)) n_stages = ns; else d2s_ts = false;
  }
  if (ts) {
    const long room = (long)budget - (long)w_res_pre - 2L * (long)stg_bytes;
    int ns = room > 0 ? (int)(room / (long)stage_bytes_pre) : 0;
    if (ns > MAX_STAGES) ns = MAX_STAGES;
    // keep the two-issuer configuration (>= 4 slots and four TMEM buffers) when the layer had it without the staging buffers
    const bool had_two = 4 * n_phase * mt * nb <= 512 && n_stages >= 4;
    if (ns >= 2 && (!had_two || ns >= 4)) n_stages = ns; else ts = false;
  }
  const int out_h_ = (out->h - out_lat.h_first + out_lat.step - 1) / out_lat.step, out_w_ = (out->w - out_lat.w_first + out_lat.step - 1) / out_lat.step;
  const int total_tiles_pre = out->n * ((out_h_ + 16 * mt - 1) / (16 * mt)) * ((out_w_ + 7) / 8);
  int ctas_pre = ctx->sm_count / (cout / nb);
  if (ctas_pre < 1) ctas_pre = 1;
  if (ctas_pre > total_tiles_pre) ctas_pre = total_tiles_pre;
  if (bn_blocks) *bn_blocks = ts ? ctas_pre : 0;
  if (bnp || bnp_query) {
    // the fused BatchNorm phase needs: the staged epilogue, ONE N block holding all output channels (<= 64), and every tile
    // of a CTA resident in TMEM at once
    const int per_cta = (total_tiles_pre + ctas_pre - 1) / ctas_pre;
    // (the skip-connection tiles of pass 2 travel through the halo stages: a stage must hold one staged tile)
    const bool ok = ts && nb == cout && cout <= 64 && !split && !kouter && n_phase == 1 && per_cta <= MAX_ACC && per_cta * mt * nb <= 512 &&
                    ctas_pre <= ctx->sm_count && stage_bytes_pre >= stg_bytes;
    if (!ok) {
      if (bn_blocks) *bn_blocks = 0;
      DG_FAIL("%s: the fused BatchNorm phase does not apply to this layer (tiles per CTA %d x mt %d x nb %d columns)", name, per_cta, mt, nb);
    }
  }
  if (bwd || bwd_query) {
    // the BatchNorm-backward epilogue rides on the staged epilogue with a 32- or 64-channel N block
    const bool ok = ts && (nb == 32 || nb == 64) && !split && !kouter && n_phase == 1 && !bias && act == DG_ACT_NONE;
    if (!ok) {
      if (bn_blocks) *bn_blocks = 0;
      DG_FAIL("%s: the BatchNorm-backward epilogue does not apply to this layer (staged %d, nb %d)", name, (int)ts, nb);
    }
  }
  if (dry) return 0;   // capability query: a tile configuration exists
  DG_REQUIRE(!bn_partials || ts, "%s: fused BatchNorm statistics need the staged epilogue (dense bf16 output, N block of 16/32/64)", name);

  P.w_block_bytes = (uint32_t)nb * kc * 2;
  P.w_res_tx = best_res ? (uint32_t)n_taps * n_chunks * P.w_block_bytes : 0;
  P.w_res_bytes = (P.w_res_tx + 1023u) & ~1023u;
  P.idesc = make_idesc_bf16(128, nb, 0, 0);
  P.tiles_h = (out_h + 16 * mt - 1) / (16 * mt);
  P.tiles_w = (out_w + 7) / 8;

  // sources: tensor maps + placement in the stage
  uint32_t off = 0, tx = 0;
  for (int s = 0; s < n_src; ++s) {
    const Lattice& L = src_lat[s];
    const int HH = 16 * hmt + dh_max[s] - dh_min[s], WW = 8 + dw_max[s] - dw_min[s];
    DG_REQUIRE(HH <= 256 && WW <= 256, "%s: halo box too large", name);
    const int vh = (in->h - L.h_first + L.step - 1) / L.step, vw = (in->w - L.w_first + L.step - 1) / L.step;
    DG_REQUIRE(vh > 0 && vw > 0, "%s: empty source view", name);
    uint64_t dims[4] = {(uint64_t)in->c, (uint64_t)vw, (uint64_t)vh, (uint64_t)in->n};
    uint64_t strides[3] = {(uint64_t)in->cpitch * 2 * L.step, (uint64_t)in->cpitch * 2 * in->w * L.step,
                           (uint64_t)in->cpitch * 2 * in->w * in->h};
    uint32_t box[4] = {(uint32_t)kc, (uint32_t)WW, (uint32_t)HH, 1};
    char* p = (char*)in->ptr + ((size_t)in->coff + ((size_t)L.h_first * in->w + L.w_first) * in->cpitch) * 2;
    if (encode_map(ctx, &P.src[s], p, 4, dims, strides, box, kc)) return 1;
    P.src_h0[s] = dh_min[s]; P.src_w0[s] = dw_min[s];
    P.src_off[s] = split ? 0u : off;
    P.a_sbo[s] = (uint32_t)WW * kc * 2;
    P.mt_stride[s] = (uint32_t)16 * WW * kc * 2;
    uint32_t hb = (uint32_t)HH * WW * kc * 2;
    tx += hb;
    off += (hb + 1023u) & ~1023u;
  }
  P.w_stage_off = off;
  if (!best_res && !kouter) { off += (uint32_t)n_taps * P.w_block_bytes; tx += (uint32_t)n_taps * P.w_block_bytes; }
  P.stage_bytes = (off + 1023u) & ~1023u;
  P.stage_tx = tx;
  if (split) {
    P.w_stage_off = max_halo(mt);
    P.stage_bytes = (P.w_stage_off + (uint32_t)max_ntaps * P.w_block_bytes + 1023u) & ~1023u;
    int t0 = 0;
    for (int s = 0; s < n_src; ++s) {
      const uint32_t hb = (uint32_t)(16 * mt + dh_max[s] - dh_min[s]) * (uint32_t)(8 + dw_max[s] - dw_min[s]) * kc * 2;
      P.src_tap0[s] = t0; P.src_ntaps[s] = src_ntaps[s];
      P.src_tx[s] = hb + (uint32_t)src_ntaps[s] * P.w_block_bytes;
      t0 += src_ntaps[s];
    }
  }
  DG_REQUIRE(P.stage_bytes == stage_bytes_pre && P.w_res_bytes == w_res_pre, "%s: internal: stage accounting mismatch", name);
  P.tstore = ts ? 1 : 0;
  P.d2s_ts = d2s_ts ? 1 : 0;
  if (ex && (ex->res || ex->prelu)) {
    DG_REQUIRE(ts && !bn_partials && !bn_blocks && !bnp && !bnp_query && !bwd && !bwd_query && !bn_fin && (!ex->prelu || act == DG_ACT_NONE),
               "%s: the residual / PReLU epilogue needs the staged store (dense bf16 output, 16/32/64-channel N block) and no other epilogue", name);
    if (ex->res) {
      const dg_tensor* r = ex->res;
      DG_REQUIRE(dg_valid(r) && r->dtype == DG_BF16 && dg_same_shape(r, out) && r->cpitch % 8 == 0 && r->coff % 8 == 0 && ((uintptr_t)r->ptr % 16) == 0,
                 "%s: the residual must be a bf16 tensor of the output's shape with 16-byte aligned pixels", name);
      P.epi_res = (const __nv_bfloat16*)r->ptr + r->coff;
      P.res_sw = r->cpitch; P.res_sh = (long)r->cpitch * r->w; P.res_sn = (long)r->cpitch * r->w * r->h;
    }
    P.epi_prelu = ex->prelu;
  }
  P.stg_bytes = stg_bytes;
  P.stg_mask = nb == 64 ? 7u : (nb == 32 ? 3u : 1u);
  P.bn_partials = bn_partials;
  if (bn_fin) {
    DG_REQUIRE(bn_partials && cout <= 512 && cout % 8 == 0, "%s: in-kernel BatchNorm finalize needs bn_partials and Cout <= 512", name);
    DG_REQUIRE(bn_fin->gamma && bn_fin->beta && bn_fin->scale && bn_fin->shift && bn_fin->save_mean && bn_fin->save_invstd &&
                   bn_fin->pixels == (long long)out->n * out_h * out_w, "%s: bad dg_bn_fused", name);
    P.bn_fin = 1;
    P.bnf = *bn_fin;
    P.ticket = ctx->tickets + 8;     // own counter: the reduction kernels of this context use tickets[0..2]
  }
  {
    // two issuing warps need four TMEM accumulator buffers and at least two pipeline slots each
    static const char* dbg_single = getenv("DG_DEBUG_SINGLE_ISSUER");   // experiments only
    P.nbuf_shift = ((bnp || 4 * n_phase * mt * nb <= 512) && n_stages >= 4 && !dbg_single && !kouter) ? 2 : 1;
    if (P.nbuf_shift == 2) n_stages &= ~1;
  }
  P.n_stages = n_stages;
  for (int t = 0; t < n_taps; ++t) {
    int s = taps[t].src;
    int WW = 8 + dw_max[s] - dw_min[s];
    P.tap_src[t] = s;
    P.tap_w[t] = taps[t].widx;
    P.tap_off[t] = P.src_off[s] + (uint32_t)((taps[t].dh - dh_min[s]) * WW + (taps[t].dw - dw_min[s])) * kc * 2;
    P.tap_adesc[t] = (make_smem_desc_hi(P.a_sbo[s], P.layout) << 32) | (uint64_t)((P.tap_off[t] >> 4) | (1u << 16));
    P.tap_ms16[t] = P.mt_stride[s] >> 4;
  }
  P.n_phase = n_phase;
  {
    bool seen[4] = {false, false, false, false};
    for (int t = 0; t < n_taps; ++t) {
      const int ph = n_phase == 1 ? 0 : taps[t].phase;
      DG_REQUIRE(ph >= 0 && ph < n_phase, "%s: bad tap phase", name);
      P.tap_acc[t] = (uint32_t)(ph * mt * nb);
      P.tap_first[t] = seen[ph] ? 0u : 1u;
      seen[ph] = true;
    }
    for (int ph = 0; ph < n_phase; ++ph) {
      DG_REQUIRE(seen[ph], "%s: output phase without taps", name);
      P.ph_off[ph] = n_phase == 1 ? 0L : ((long)phase_lat[ph].h_first * out->w + phase_lat[ph].w_first) * (long)out->cpitch;
    }
  }
  // weights: 2D [rows][kc]
  {
    // total rows are not needed exactly for correctness of in-bounds loads; give the true extent
    uint64_t rows = (uint64_t)w_rows_per_block * n_chunks * MAX_TAPS;  // upper bound, never read past real data
    uint64_t dims[2] = {(uint64_t)kc, rows};
    uint64_t strides[1] = {(uint64_t)kc * 2};
    uint32_t box[2] = {(uint32_t)kc, (uint32_t)nb};
    if (encode_map(ctx, &P.wmap, const_cast<void*>(w_packed), 2, dims, strides, box, kc)) return 1;
  }
  // output
  P.out = (char*)out->ptr + ((size_t)out->coff + ((size_t)out_lat.h_first * out->w + out_lat.w_first) * out->cpitch) * out_esz;
  P.out_sw = (long)out->cpitch * out_lat.step;
  P.out_sh = (long)out->cpitch * out->w * out_lat.step;
  P.out_sn = (long)out->cpitch * out->w * out->h;
  if (d2s_cq > 0) {      // `out` describes the convolution's own grid; the stored tensor is [n, 2h, 2w, d2s_cq] with channel pitch out->cpitch
    P.out_sw = (long)out->cpitch;
    P.out_sh = (long)out->cpitch * 2 * out->w;
    P.out_sn = (long)out->cpitch * 2 * out->w * 2 * out->h;
  }
  P.out_f32 = out->dtype == DG_F32;
  P.out_cvalid = out_cvalid;
  P.d2s_cq = d2s_cq; P.d2s_prelu = d2s_prelu;
  P.bias = bias; P.act = act; P.alpha = alpha;
  P.dbg = g_dbg_timeline;
  P.dbg_flags = g_dbg_flags;

  P.stg_off = P.w_res_bytes + ring_bytes + (uint32_t)n_stages * P.stage_bytes;
  if (d2s_ts) {
    // the stored [n, 2h, 2w, 32] tensor as (pixel pair x 32 channels = 128 bytes, w pairs, 2h rows, n); one box = the tile's 32 x 16 output pixels
    uint64_t dims[4] = {64u, (uint64_t)out->w, 2ull * (uint64_t)out->h, (uint64_t)out->n};
    uint64_t strides[3] = {128u, 128ull * (uint64_t)out->w, 128ull * (uint64_t)out->w * 2ull * (uint64_t)out->h};
    uint32_t box[4] = {64u, 8u, 32u, 1u};
    if (encode_map(ctx, &P.omap, out->ptr, 4, dims, strides, box, 64)) return 1;
  }
  if (ts) {
    // output view as a 4-D tensor map (C, W, H, N); one box = the CTA's tile, clipped at the image border by the TMA unit
    uint64_t dims[4] = {(uint64_t)out->c, (uint64_t)out->w, (uint64_t)out->h, (uint64_t)out->n};
    uint64_t strides[3] = {(uint64_t)out->cpitch * 2, (uint64_t)out->cpitch * 2 * out->w, (uint64_t)out->cpitch * 2 * out->w * out->h};
    uint32_t box[4] = {(uint32_t)nb, 8u, (uint32_t)(16 * mt), 1u};
    if (encode_map(ctx, &P.omap, (char*)out->ptr + (size_t)out->coff * 2, 4, dims, strides, box, nb)) return 1;
  }
  if (bnp) {
    const dg_tensor* o2 = bnp->out2;
    DG_REQUIRE(bn_partials && bnp->bn && o2 && dg_valid(o2) && dg_same_shape(o2, out) && o2->dtype == DG_BF16 && act == DG_ACT_NONE,
               "%s: bad BatchNorm-phase arguments", name);
    DG_REQUIRE(((uintptr_t)o2->ptr % 16) == 0 && (o2->cpitch * 2) % 16 == 0 && (o2->coff * 2) % 16 == 0, "%s: out2 view not 16-byte aligned", name);
    DG_REQUIRE(bnp->act == DG_ACT_NONE || bnp->act == DG_ACT_RELU || bnp->act == DG_ACT_LRELU || (bnp->act == DG_ACT_PRELU && bnp->prelu),
               "%s: the BatchNorm phase takes none / relu / leaky relu / PReLU", name);
    DG_REQUIRE(bnp->bn->gamma && bnp->bn->beta && bnp->bn->scale && bnp->bn->shift && bnp->bn->save_mean && bnp->bn->save_invstd &&
                   bnp->bn->pixels == (long long)out->n * out_h * out_w, "%s: bad dg_bn_fused", name);
    P.bnp = 1; P.bnp_act = bnp->act; P.bnp_alpha = bnp->alpha; P.bnp_prelu = bnp->act == DG_ACT_PRELU ? bnp->prelu : nullptr;
    P.bnf = *bnp->bn;
    P.gbar = ctx->tickets + 16;
    uint32_t box[4] = {(uint32_t)nb, 8u, (uint32_t)(16 * mt), 1u};
    {
      uint64_t dims[4] = {(uint64_t)o2->c, (uint64_t)o2->w, (uint64_t)o2->h, (uint64_t)o2->n};
      uint64_t strides[3] = {(uint64_t)o2->cpitch * 2, (uint64_t)o2->cpitch * 2 * o2->w, (uint64_t)o2->cpitch * 2 * o2->w * o2->h};
      if (encode_map(ctx, &P.omap2, (char*)o2->ptr + (size_t)o2->coff * 2, 4, dims, strides, box, nb)) return 1;
    }
    if (bnp->res) {
      const dg_tensor* r = bnp->res;
      DG_REQUIRE(dg_valid(r) && dg_same_shape(r, out) && r->dtype == DG_BF16 && ((uintptr_t)r->ptr % 16) == 0 && (r->cpitch * 2) % 16 == 0 &&
                     (r->coff * 2) % 16 == 0, "%s: bad residual view", name);
      DG_REQUIRE(P.stage_bytes >= stg_bytes, "%s: internal: residual tile larger than a pipeline stage", name);
      uint64_t dims[4] = {(uint64_t)r->c, (uint64_t)r->w, (uint64_t)r->h, (uint64_t)r->n};
      uint64_t strides[3] = {(uint64_t)r->cpitch * 2, (uint64_t)r->cpitch * 2 * r->w, (uint64_t)r->cpitch * 2 * r->w * r->h};
      if (encode_map(ctx, &P.rmap, (char*)r->ptr + (size_t)r->coff * 2, 4, dims, strides, box, nb)) return 1;
      P.bnp_res = 1;
    }
  }
  P.bwd_am = -1;
  if (bwd) {
    DG_REQUIRE(!bnp && !bn_fin, "%s: the BatchNorm-backward epilogue excludes the forward BatchNorm modes", name);
    if (bwd->res) {
      const dg_tensor* r = bwd->res;
      DG_REQUIRE(dg_valid(r) && dg_same_shape(r, out) && r->dtype == DG_BF16 && ((uintptr_t)r->ptr % 16) == 0 && r->cpitch % 8 == 0 &&
                     r->coff % 8 == 0, "%s: bad skip-gradient view", name);
      P.bwd_res = (const __nv_bfloat16*)r->ptr + r->coff;
      P.bwd_r_sw = r->cpitch; P.bwd_r_sh = (long)r->cpitch * r->w; P.bwd_r_sn = (long)r->cpitch * r->w * r->h;
    }
    if (bwd->relu_y) {
      const dg_tensor* yb = bwd->relu_y;
      DG_REQUIRE(!bwd->bn && dg_valid(yb) && dg_same_shape(yb, out) && yb->dtype == DG_BF16 && ((uintptr_t)yb->ptr % 16) == 0 &&
                     yb->cpitch % 8 == 0 && yb->coff % 8 == 0, "%s: bad ReLU-mask arguments", name);
      P.bwd_am = 1; P.bwd_mask = 1; P.bwd_alpha = 0.f;
      P.bwd_y = (const __nv_bfloat16*)yb->ptr + yb->coff;
      P.bwd_y_sw = yb->cpitch; P.bwd_y_sh = (long)yb->cpitch * yb->w; P.bwd_y_sn = (long)yb->cpitch * yb->w * yb->h;
      uint64_t dims[4] = {(uint64_t)yb->c, (uint64_t)yb->w, (uint64_t)yb->h, (uint64_t)yb->n};
      uint64_t strides[3] = {(uint64_t)yb->cpitch * 2, (uint64_t)yb->cpitch * 2 * yb->w, (uint64_t)yb->cpitch * 2 * yb->w * yb->h};
      uint32_t box[4] = {(uint32_t)nb, 8u, (uint32_t)(16 * mt), 1u};
      if (encode_map(ctx, &P.rmap, (char*)yb->ptr + (size_t)yb->coff * 2, 4, dims, strides, box, nb)) return 1;
    }
    if (bwd->bn) {
      const dg_bn_bwd_stats* b = bwd->bn;
      const dg_tensor* yb = b->y;
      DG_REQUIRE(bn_partials && b->scale && b->shift && b->mean && dg_valid(yb) && dg_same_shape(yb, out) && yb->dtype == DG_BF16 &&
                     ((uintptr_t)yb->ptr % 16) == 0 && yb->cpitch % 8 == 0 && yb->coff % 8 == 0, "%s: bad BatchNorm-backward arguments", name);
      DG_REQUIRE(b->act == DG_ACT_NONE || b->act == DG_ACT_RELU || b->act == DG_ACT_LRELU, "%s: the BatchNorm-backward epilogue takes none / relu / leaky relu", name);
      P.bwd_am = b->act == DG_ACT_NONE ? 0 : (b->act == DG_ACT_RELU ? 1 : 2);
      P.bwd_alpha = b->alpha;
      P.bwd_y = (const __nv_bfloat16*)yb->ptr + yb->coff;
      P.bwd_y_sw = yb->cpitch; P.bwd_y_sh = (long)yb->cpitch * yb->w; P.bwd_y_sn = (long)yb->cpitch * yb->w * yb->h;
      P.bwd_scale = b->scale; P.bwd_shift = b->shift; P.bwd_mean = b->mean;
      // yb tiles are TMA-loaded into the staging buffer: same box and swizzle as the output store
      uint64_t dims[4] = {(uint64_t)yb->c, (uint64_t)yb->w, (uint64_t)yb->h, (uint64_t)yb->n};
      uint64_t strides[3] = {(uint64_t)yb->cpitch * 2, (uint64_t)yb->cpitch * 2 * yb->w, (uint64_t)yb->cpitch * 2 * yb->w * yb->h};
      uint32_t box[4] = {(uint32_t)nb, 8u, (uint32_t)(16 * mt), 1u};
      if (encode_map(ctx, &P.rmap, (char*)yb->ptr + (size_t)yb->coff * 2, 4, dims, strides, box, nb)) return 1;
    }
  }
  const uint32_t smem = P.w_res_bytes + ring_bytes + (uint32_t)n_stages * P.stage_bytes + ((ts || d2s_ts) ? 2u * stg_bytes : 0u) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 3072);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 3072);
    if (e != cudaSuccess) DG_FAIL("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    attr_set = true;
  }
  const int n_blocks = cout / nb;
  const int total_tiles = P.n_img * P.tiles_h * P.tiles_w;
  {
    static const char* dbg_cfg = getenv("DG_DEBUG_CONFIG");   // prints the tile configuration of every launch
    if (dbg_cfg)
      fprintf(stderr, "[%s] %dx%dx%d c%d->%d taps %d src %d: nb %d mt %d kc %d chunks %d resident %d split %d kouter %d phases %d stages %d x %u B, issuers %d, tstore %d, smem %u\n",
              name, out->n, out_h, out_w, in->c, cout, n_taps, n_src, nb, mt, kc, n_chunks, best_res, split, kouter, n_phase, n_stages,
              P.stage_bytes, P.nbuf_shift == 2 ? 2 : 1, P.tstore, smem);
  }
  int ctas = ctx->sm_count / n_blocks;
  if (ctas < 1) ctas = 1;
  if (ctas > total_tiles) ctas = total_tiles;
  dim3 grid(ctas, n_blocks);
  if (bnp) {
    DG_REQUIRE(n_blocks == 1 && ctas == ctas_pre, "%s: internal: BatchNorm-phase grid mismatch", name);
    DG_REQUIRE(dg_coresident(umma_conv_kernel<false>, CONV_THREADS, smem, ctas, ctx->sm_count), "%s: the grid does not fit the device at once", name);
    cudaError_t e = dg_coop_launch(umma_conv_kernel<false>, grid, dim3(CONV_THREADS), smem, st, P);
    if (e != cudaSuccess) DG_FAIL("%s: cooperative launch failed: %s", name, cudaGetErrorString(e));
  } else if (bwd) {
    dg_pdl_launch(umma_conv_kernel<true>, grid, dim3(CONV_THREADS), smem, st, P);
  } else {
    dg_pdl_launch(umma_conv_kernel<false>, grid, dim3(CONV_THREADS), smem, st, P);
  }
  DG_CHECK_LAUNCH(name);
  return 0;
}

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int pymod(int a, int b) { return a - floordiv(a, b) * b; }

}  // namespace

// Debug aid: when set to a device buffer of 3*16*4 int64, CTA 0 of every conv launch records clock64() marks
// (role, tile, slot); pass NULL to switch off.  Not part of the hot path.
extern "C" void dg_debug_conv_timeline(void* dev_buffer) { g_dbg_timeline = (long long*)dev_buffer; }
extern "C" void dg_debug_conv_flags(int flags) { g_dbg_flags = flags; }

extern "C" size_t dg_umma_packed_bytes(int kh, int kw, int cin, int cout, int mode) {
  return (size_t)kh * kw * cin * cout * 2;
}

extern "C" int dg_umma_pack_weights(dg_ctx* ctx, const float* w, void* packed, int kh, int kw, int cin, int cout,
                                    int mode, void* stream) {
  DG_REQUIRE(w && packed, "dg_umma_pack_weights: null argument");
  DG_REQUIRE(cin % 16 == 0 && cout % 16 == 0, "dg_umma_pack_weights: channels must be multiples of 16");
  DG_REQUIRE(mode == 0 || mode == 1, "dg_umma_pack_weights: bad mode");
  int kc = kc_for(mode == 0 ? cin : cout);
  long total = (long)kh * kw * cin * cout;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)packed, kh * kw, cin, cout, kc, mode, cin, cout, 0, 0);
  DG_CHECK_LAUNCH("dg_umma_pack_weights");
  return 0;
}

extern "C" int dg_umma_pack_weights_seg(dg_ctx* ctx, const float* w, void* packed, int kh, int kw, int cin, int cout, int cin_pad,
                                        int cout_pad, int seg_log, int seg_phys, int mode, void* stream) {
  DG_REQUIRE(w && packed, "dg_umma_pack_weights_seg: null argument");
  DG_REQUIRE(cin_pad % 16 == 0 && cout_pad % 16 == 0 && cin <= cin_pad && cout <= cout_pad && (mode == 0 || mode == 1),
             "dg_umma_pack_weights_seg: padded channels must be multiples of 16");
  DG_REQUIRE(seg_phys == 0 || (seg_log >= 0 && seg_log <= seg_phys && seg_log <= cin && seg_phys % 16 == 0 && seg_phys <= cin_pad &&
                               cin - seg_log <= cin_pad - seg_phys && cin_pad < 65536),
             "dg_umma_pack_weights_seg: bad channel segments (%d of %d, then %d of %d)", seg_log, seg_phys, cin - seg_log, cin_pad - seg_phys);
  int kc = kc_for(mode == 0 ? cin_pad : cout_pad);
  long total = (long)kh * kw * cin_pad * cout_pad;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)packed, kh * kw, cin_pad, cout_pad, kc, mode,
                                                                          cin, cout, seg_log, seg_phys);
  DG_CHECK_LAUNCH("dg_umma_pack_weights_seg");
  return 0;
}

extern "C" int dg_umma_pack_weights_padded(dg_ctx* ctx, const float* w, void* packed, int kh, int kw, int cin, int cout, int cin_pad,
                                           int cout_pad, int mode, void* stream) {
  DG_REQUIRE(w && packed, "dg_umma_pack_weights_padded: null argument");
  DG_REQUIRE(cin_pad % 16 == 0 && cout_pad % 16 == 0 && cin <= cin_pad && cout <= cout_pad && (mode == 0 || mode == 1),
             "dg_umma_pack_weights_padded: padded channels must be multiples of 16");
  int kc = kc_for(mode == 0 ? cin_pad : cout_pad);
  long total = (long)kh * kw * cin_pad * cout_pad;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)packed, kh * kw, cin_pad, cout_pad, kc, mode,
                                                                          cin, cout, 0, 0);
  DG_CHECK_LAUNCH("dg_umma_pack_weights_padded");
  return 0;
}

extern "C" int dg_umma_pack_weights_batch(dg_ctx* ctx, const void* table_dev, int n_entries, void* stream) {
  DG_REQUIRE(table_dev && n_entries > 0, "dg_umma_pack_weights_batch: empty table");
  // blockIdx.x strides over one kernel: enough blocks that a 4x4x512x512 kernel (pix2pix) is spread over all SMs
  dim3 grid(ctx && ctx->sm_count > 0 ? 2 * ctx->sm_count : 296, n_entries);
  pack_weights_batch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const PackEntry*)table_dev);
  DG_CHECK_LAUNCH("dg_umma_pack_weights_batch");
  return 0;
}

static int conv_fwd_impl(dg_ctx* ctx, const dg_tensor* x, const void* w_packed, const float* bias, const dg_tensor* y,
                         const dg_conv_params* p, void* stream, bool dry, float* bn_partials = nullptr, int* bn_blocks = nullptr,
                         const dg_bn_fused* bn_fin = nullptr, const BnPhase* bnp = nullptr, bool bnp_query = false, int out_cvalid = 0,
                         int d2s_cq = 0, const float* d2s_prelu = nullptr, const EpiExtra* ex = nullptr) {
  DG_REQUIRE(dg_valid(x) && y && y->ptr && w_packed && p, "dg_umma_conv2d_fwd: null argument");
  DG_REQUIRE(p->stride == 1 || p->stride == 2, "dg_umma_conv2d_fwd: stride must be 1 or 2");
  DG_REQUIRE(x->n == y->n, "dg_umma_conv2d_fwd: batch mismatch");
  DG_REQUIRE(p->kh * p->kw <= MAX_TAPS, "dg_umma_conv2d_fwd: kernel too large");
  TapSpec taps[MAX_TAPS];
  Lattice lat[MAX_SRC];
  int n_src = 0, n_taps = 0;
  if (p->stride == 1) {
    lat[0] = Lattice{1, 0, 0};
    n_src = 1;
    for (int r = 0; r < p->kh; ++r)
      for (int s = 0; s < p->kw; ++s) taps[n_taps++] = TapSpec{0, r - p->pad_t, s - p->pad_l, r * p->kw + s};
  } else {
    DG_REQUIRE(x->h % 2 == 0 && x->w % 2 == 0, "dg_umma_conv2d_fwd: stride 2 needs even input size");
    int src_of[2][2] = {{-1, -1}, {-1, -1}};
    for (int r = 0; r < p->kh; ++r)
      for (int s = 0; s < p->kw; ++s) {
        int ph = pymod(r - p->pad_t, 2), pw = pymod(s - p->pad_l, 2);
        if (src_of[ph][pw] < 0) { src_of[ph][pw] = n_src; lat[n_src++] = Lattice{2, ph, pw}; }
        taps[n_taps++] = TapSpec{src_of[ph][pw], floordiv(r - p->pad_t - ph, 2), floordiv(s - p->pad_l - pw, 2), r * p->kw + s};
      }
  }
  return launch_conv(ctx, "dg_umma_conv2d_fwd", x, lat, n_src, taps, n_taps, w_packed, y->c, y, Lattice{1, 0, 0}, bias,
                     p->act, p->act_alpha, (cudaStream_t)stream, dry, bn_partials, bn_blocks, 1, nullptr, bn_fin, bnp, bnp_query, nullptr, false,
                     out_cvalid, d2s_cq, d2s_prelu, ex);
}

// Conv2D + depth_to_space(2) + PReLU (srgan.py:144-146, fsrgan.py:180-186: the up-sampling blocks) as ONE launch for inference: the
// convolution's epilogue stores channel c = blk * (Cout/4) + cc of pixel (h, w) at pixel (2h + blk / 2, 2w + blk % 2), channel cc of
// y [n, 2h, 2w, Cout/4] (TensorFlow's DCR order) after bias and PReLU(shared_axes=[1,2]) with slope prelu_alpha[cc] (may be NULL).
// The pre-activation tensor (2.7 GB for the last block of a 1080p Fast-SRGAN frame) is never written.  Training keeps the two-launch form
// (the backward pass needs the pre-activation).
extern "C" int dg_umma_conv2d_fwd_d2s_prelu(dg_ctx* ctx, const dg_tensor* x, const void* w_packed, const float* bias, const dg_tensor* y,
                                            const dg_conv_params* p, const float* prelu_alpha, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y) && p && p->stride == 1 && p->act == DG_ACT_NONE, "dg_umma_conv2d_fwd_d2s_prelu: bad argument");
  DG_REQUIRE(y->dtype == DG_BF16 && y->coff == 0 && y->h % 2 == 0 && y->w % 2 == 0 && y->c % 32 == 0 && y->c <= 192,
             "dg_umma_conv2d_fwd_d2s_prelu: y must be a bf16 [n, 2h, 2w, Cout/4] tensor with a multiple of 32 (<= 192) channels");
  dg_tensor yc = *y;             // the convolution's own output grid
  yc.h = y->h / 2; yc.w = y->w / 2; yc.c = 4 * y->c;
  return conv_fwd_impl(ctx, x, w_packed, bias, &yc, p, stream, false, nullptr, nullptr, nullptr, nullptr, false, 0, y->c, prelu_alpha);
}

// Conv2D -> BatchNormalization(training=False) -> [PReLU] -> [Add skip] as ONE launch for inference (fsrgan.py:172-176 project + add,
// :208-210 post-residual conv + add; srgan.py:166-169, :174-176): the BatchNorm is folded into w_packed / bias by the caller, the staged
// epilogue applies y = act(acc + bias) + residual with act = PReLU(prelu_alpha[c]) when slopes are given (else p->act).  Fails when the
// layer does not take the staged epilogue (dg_umma_conv2d_fwd_bn_blocks() == 0): the caller then issues convolution and pointwise pass.
extern "C" int dg_umma_conv2d_fwd_res_prelu(dg_ctx* ctx, const dg_tensor* x, const void* w_packed, const float* bias, const dg_tensor* y,
                                            const dg_conv_params* p, const dg_tensor* residual, const float* prelu_alpha, void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(y) && p, "dg_umma_conv2d_fwd_res_prelu: bad argument");
  EpiExtra ex{residual, prelu_alpha};
  return conv_fwd_impl(ctx, x, w_packed, bias, y, p, stream, false, nullptr, nullptr, nullptr, nullptr, false, 0, 0, nullptr, &ex);
}

// Conv2D whose output has fewer than 16 channels (the RGB image: srgan.py:182, fsrgan.py:217, autoencoder.py:186), fp32: the packed
// kernel and the bias are zero-padded to 16 output channels (dg_umma_pack_weights_padded), the tensor cores compute all 16, and the
// epilogue stores only the y->c real ones into the DENSE tensor y -- no padded copy of the output, no slicing pass.
extern "C" int dg_umma_conv2d_fwd_narrow(dg_ctx* ctx, const dg_tensor* x, const void* w_packed, const float* bias_padded, const dg_tensor* y,
                                         const dg_conv_params* p, void* stream) {
  DG_REQUIRE(dg_valid(y) && y->dtype == DG_F32 && y->c < 16 && y->coff == 0, "dg_umma_conv2d_fwd_narrow: y must be an fp32 tensor of fewer than 16 channels");
  dg_tensor y16 = *y;
  y16.c = 16;
  return conv_fwd_impl(ctx, x, w_packed, bias_padded, &y16, p, stream, false, nullptr, nullptr, nullptr, nullptr, false, y->c);
}

// Conv2D + training-mode BatchNormalization + activation (+ skip-add) in one cooperative launch (UmmaConvParams::bnp).
// y receives the raw convolution output (the backward pass needs it), out = act(BN(y)) (+ residual); bn_partials is the
// [dg_umma_conv2d_fwd_bn_act_blocks()][2][Cout] workspace of the statistics.
extern "C" int dg_umma_conv2d_fwd_bn_act(dg_ctx* ctx, const dg_tensor* x, const void* w_packed, const float* bias, const dg_tensor* y,
                                         const dg_conv_params* p, float* bn_partials, const dg_bn_fused* bn, int act, float act_alpha,
                                         const float* prelu_alpha, const dg_tensor* residual, const dg_tensor* out, void* stream) {
  DG_REQUIRE(bn_partials && bn && out, "dg_umma_conv2d_fwd_bn_act: null argument");
  BnPhase ph{bn, act, act_alpha, prelu_alpha, residual, out};
  return conv_fwd_impl(ctx, x, w_packed, bias, y, p, stream, false, bn_partials, nullptr, nullptr, &ph);
}

// Rows of the statistics workspace when the fused BatchNorm phase applies to this layer (all tiles of a CTA fit TMEM, one
// N block of <= 64 channels, dense bf16 output), else 0: the caller then runs dg_umma_conv2d_fwd + dg_bn_finalize + dg_bn_act_fwd.
extern "C" int dg_umma_conv2d_fwd_bn_act_blocks(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* y, const dg_conv_params* p) {
  int blocks = 0;
  if (conv_fwd_impl(ctx, x, (const void*)1, nullptr, y, p, nullptr, true, nullptr, &blocks, nullptr, nullptr, true) != 0) return 0;
  return blocks;
}

extern "C" int dg_umma_conv2d_fwd(dg_ctx* ctx, const dg_tensor* x, const void* w_packed, const float* bias,
                                  const dg_tensor* y, const dg_conv_params* p, float* bn_partials, void* stream) {
  return conv_fwd_impl(ctx, x, w_packed, bias, y, p, stream, false, bn_partials);
}

extern "C" int dg_umma_conv2d_fwd_bn(dg_ctx* ctx, const dg_tensor* x, const void* w_packed, const float* bias, const dg_tensor* y,
                                     const dg_conv_params* p, float* bn_partials, const dg_bn_fused* bn, void* stream) {
  DG_REQUIRE(bn_partials && bn, "dg_umma_conv2d_fwd_bn: null argument");
  return conv_fwd_impl(ctx, x, w_packed, bias, y, p, stream, false, bn_partials, nullptr, bn);
}

// Number of per-CTA partial rows ([rows][2][Cout] floats) dg_umma_conv2d_fwd writes into `bn_partials` for this layer,
// or 0 when the layer does not run the staged epilogue (the caller then issues dg_bn_stats).
extern "C" int dg_umma_conv2d_fwd_bn_blocks(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* y, const dg_conv_params* p) {
  int blocks = 0;
  if (conv_fwd_impl(ctx, x, (const void*)1, nullptr, y, p, nullptr, true, nullptr, &blocks) != 0) return 0;
  return blocks;
}

// 1 when the tensor-core kernel has a tile configuration for this layer (shared-memory fit), else 0
extern "C" int dg_umma_conv2d_fwd_supported(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* y, const dg_conv_params* p) {
  return conv_fwd_impl(ctx, x, (const void*)1, nullptr, y, p, nullptr, true) == 0 ? 1 : 0;
}

static int conv_dgrad_impl(dg_ctx* ctx, const dg_tensor* dy, const void* w_packed, const float* bias, const dg_tensor* dx,
                           const dg_conv_params* p, void* stream, bool dry, const BwdEpi* bwd = nullptr, bool bwd_query = false,
                           int* bwd_blocks = nullptr) {
  DG_REQUIRE(dg_valid(dy) && dg_valid(dx) && w_packed && p, "dg_umma_conv2d_dgrad: null argument");
  DG_REQUIRE(p->stride == 1 || p->stride == 2, "dg_umma_conv2d_dgrad: stride must be 1 or 2");
  DG_REQUIRE(dx->n == dy->n, "dg_umma_conv2d_dgrad: batch mismatch");
  DG_REQUIRE(p->kh * p->kw <= MAX_TAPS, "dg_umma_conv2d_dgrad: kernel too large");
  Lattice dense{1, 0, 0};
  TapSpec taps[MAX_TAPS];
  if (p->stride == 1) {
    int n_taps = 0;
    for (int r = 0; r < p->kh; ++r)
      for (int s = 0; s < p->kw; ++s) taps[n_taps++] = TapSpec{0, p->pad_t - r, p->pad_l - s, r * p->kw + s};
    return launch_conv(ctx, "dg_umma_conv2d_dgrad", dy, &dense, 1, taps, n_taps, w_packed, dx->c, dx, dense, bias,
                       p->act, p->act_alpha, (cudaStream_t)stream, dry, bwd && bwd->bn ? bwd->bn->partials : nullptr, bwd_blocks, 1, nullptr,
                       nullptr, nullptr, false, bwd, bwd_query);
  }
  DG_REQUIRE(!bwd && !bwd_query, "dg_umma_conv2d_dgrad: the BatchNorm-backward epilogue needs a stride-1 convolution");
  DG_REQUIRE(dx->h % 2 == 0 && dx->w % 2 == 0, "dg_umma_conv2d_dgrad: stride 2 needs even image size");
  {
    // All four output parity phases in ONE launch when their accumulators fit TMEM together (2 buffers x 4 phases x mt x nb
    // <= 512 columns: the 32/64-channel discriminator layers): dy is read once instead of four times and one prologue /
    // tail is paid instead of four.  Otherwise (wide layers, pix2pix) one launch per phase below.
    static const char* no_fuse = getenv("DG_DEBUG_NO_PHASE_FUSION");   // experiments only
    Lattice plat[4];
    int n_all = 0;
    bool ok = !no_fuse;
    for (int a = 0; a < 2 && ok; ++a)
      for (int b = 0; b < 2 && ok; ++b) {
        plat[a * 2 + b] = Lattice{2, a, b};
        int cnt = 0;
        for (int r = 0; r < p->kh; ++r) {
          if (pymod(a + p->pad_t - r, 2) != 0) continue;
          for (int s = 0; s < p->kw; ++s) {
            if (pymod(b + p->pad_l - s, 2) != 0) continue;
            taps[n_all++] = TapSpec{0, floordiv(a + p->pad_t - r, 2), floordiv(b + p->pad_l - s, 2), r * p->kw + s, a * 2 + b};
            ++cnt;
          }
        }
        if (cnt == 0) ok = false;
      }
    if (ok && launch_conv(ctx, "dg_umma_conv2d_dgrad", dy, &dense, 1, taps, n_all, w_packed, dx->c, dx, Lattice{2, 0, 0}, bias, p->act,
                          p->act_alpha, (cudaStream_t)stream, true, nullptr, nullptr, 4, plat) == 0) {
      if (dry) return 0;
      return launch_conv(ctx, "dg_umma_conv2d_dgrad", dy, &dense, 1, taps, n_all, w_packed, dx->c, dx, Lattice{2, 0, 0}, bias, p->act,
                         p->act_alpha, (cudaStream_t)stream, false, nullptr, nullptr, 4, plat);
    }
  }
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      int n_taps = 0;
      for (int r = 0; r < p->kh; ++r) {
        if (pymod(a + p->pad_t - r, 2) != 0) continue;
        for (int s = 0; s < p->kw; ++s) {
          if (pymod(b + p->pad_l - s, 2) != 0) continue;
          taps[n_taps++] = TapSpec{0, floordiv(a + p->pad_t - r, 2), floordiv(b + p->pad_l - s, 2), r * p->kw + s};
        }
      }
      DG_REQUIRE(n_taps > 0, "dg_umma_conv2d_dgrad: output phase without taps (kernel smaller than stride)");
      if (launch_conv(ctx, "dg_umma_conv2d_dgrad", dy, &dense, 1, taps, n_taps, w_packed, dx->c, dx, Lattice{2, a, b},
                      bias, p->act, p->act_alpha, (cudaStream_t)stream, dry))
        return 1;
    }
  return 0;
}

extern "C" int dg_umma_conv2d_dgrad(dg_ctx* ctx, const dg_tensor* dy, const void* w_packed, const float* bias,
                                    const dg_tensor* dx, const dg_conv_params* p, void* stream) {
  return conv_dgrad_impl(ctx, dy, w_packed, bias, dx, p, stream, false);
}

// Input gradient + the skip-connection add + the statistics half of the BatchNorm backward pass in front of the convolution
// (UmmaConvParams::bwd_*): `residual` (may be NULL) is added to the result, `bn` (may be NULL) describes the BatchNormalization
// whose OUTPUT gradient this launch produces; its partials rows come from dg_umma_conv2d_dgrad_fused_blocks().
extern "C" int dg_umma_conv2d_dgrad_fused(dg_ctx* ctx, const dg_tensor* dy, const void* w_packed, const dg_tensor* dx, const dg_conv_params* p,
                                          const dg_tensor* residual, const dg_bn_bwd_stats* bn, void* stream) {
  BwdEpi e{residual, bn};
  return conv_dgrad_impl(ctx, dy, w_packed, nullptr, dx, p, stream, false, &e);
}

// Rows of the [rows][2][Cin] fp32 statistics workspace of dg_umma_conv2d_dgrad_fused, or 0 when the fused epilogue does not apply
// to the layer (stride 2, streamed weights, N block other than 32 / 64): the caller then issues dg_umma_conv2d_dgrad, dg_add and the
// two-pass dg_bn_act_bwd.
extern "C" int dg_umma_conv2d_dgrad_fused_blocks(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* dx, const dg_conv_params* p) {
  int blocks = 0;
  if (p->stride != 1) return 0;
  if (conv_dgrad_impl(ctx, dy, (const void*)1, nullptr, dx, p, nullptr, true, nullptr, true, &blocks) != 0) return 0;
  return blocks;
}

// Input gradient of a stride-1 convolution whose INPUT was the output y = relu(conv(...)) of another convolution (autoencoder.py:95-104,
// conv2d -> conv2d chains): dx is stored already multiplied by (y > 0), i.e. the ReLU backward pass of the producing layer is folded
// into this launch (one read of y through the epilogue's staging buffer instead of a separate pass over dx and y).  Applies where
// dg_umma_conv2d_dgrad_fused_blocks() > 0.
extern "C" int dg_umma_conv2d_dgrad_relu_mask(dg_ctx* ctx, const dg_tensor* dy, const void* w_packed, const dg_tensor* dx, const dg_conv_params* p,
                                              const dg_tensor* y_relu, void* stream) {
  DG_REQUIRE(y_relu, "dg_umma_conv2d_dgrad_relu_mask: null argument");
  BwdEpi e{nullptr, nullptr, y_relu};
  return conv_dgrad_impl(ctx, dy, w_packed, nullptr, dx, p, stream, false, &e);
}

extern "C" int dg_umma_conv2d_dgrad_supported(dg_ctx* ctx, const dg_tensor* dy, const dg_tensor* dx, const dg_conv_params* p) {
  return conv_dgrad_impl(ctx, dy, (const void*)1, nullptr, dx, p, nullptr, true) == 0 ? 1 : 0;
}

// wgrad on tensor cores: see conv_umma_wgrad.cu
