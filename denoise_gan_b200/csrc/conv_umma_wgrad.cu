// Tensor-core weight gradient for sm_100a:
//      dW[t][c][o] (+)= sum_{n,h,w} S_t[n, h+dh_t, w+dw_t, c] * dy[n,h,w,o]        db[o] (+)= sum dy[n,h,w,o]
// with the same sources/taps as the forward "multi-source halo conv" (conv_umma.cu).
//
// GEMM view: per tap D_t[M = Cin, N = Cout] = X_t^T dY with the PIXELS as the contraction.  Both
// operands are pixel-major in NHWC, i.e. "MN-major" for tcgen05: the halo box of x and the dense
// 16x8 tile of dy are TMA-loaded exactly as in the forward kernel and consumed in place through
// MN-major shared-memory descriptors (a_major = b_major = 1) -- no transpose anywhere.  A tap is a
// row-shifted window of the halo box (start address += shift * pixel_bytes).  UMMA wants M = 128:
//   * Cin >= 128: the two 64-channel chunk boxes of a tap form one M=128 operand (LBO = box pitch);
//   * Cin <= 64 : several TAPS are stacked along M instead; their windows differ by a constant
//     number of pixels, which is simply the descriptor's leading-dimension byte offset (LBO).
// Every CTA keeps all its [128 x Nblk] accumulators in TMEM (<= 512 columns) across ALL of its
// pixel tiles and writes one fp32 partial at the end; a second kernel sums the partials in CTA
// order (deterministic split-K).  The bias gradient is summed by the four epilogue warps (idle until the
// accumulators are final) straight from the staged dy tiles, one more row of the per-CTA partial.
// (Tried and dropped in round 1: summing the partials of 4-CTA clusters through distributed shared memory with
// per-thread st.shared::cluster stores took 20 K cycles per CTA against 12.7 K for the plain per-CTA dump; a bulk-copy
// version is the follow-up)
//
// Reference: tape.gradient(..., trainable_variables) train_srgan.py:111-112 for every Conv2D /
// Conv2DTranspose kernel and bias.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "dg_common.cuh"
#include "sm100.cuh"

namespace {

using namespace sm100;

constexpr int MAX_SRC = 4;
constexpr int MAX_TAPS = 16;
constexpr int MAX_GROUPS = 24;
constexpr int MAX_ATOMS = 8;
constexpr int MAX_STAGES = 6;
constexpr int MAX_PROB = 4;      // layers of identical geometry whose weight gradients share ONE launch (dg_umma_conv2d_wgrad_batch)
constexpr int WG_THREADS = 192;
constexpr uint32_t SMEM_LIMIT = 227 * 1024;

struct WgradParams {
  CUtensorMap src[MAX_PROB][MAX_SRC];
  CUtensorMap dymap[MAX_PROB];
  int splits;       // CTAs along gridDim.x PER PROBLEM: CTA blockIdx.x works on problem blockIdx.x / splits, pixel split blockIdx.x % splits
  int n_src, kc, kco, nb, n_groups, groups_per_cta, chunks, chunk0;
  int zblocks;      // group blocks along gridDim.z; the rest of gridDim.z enumerates BLOCKS OF CHUNKS (chunk0 += chunks per block)
  int tiles_h, tiles_w, n_img;
  int src_h0[MAX_SRC], src_w0[MAX_SRC];
  uint32_t src_off[MAX_SRC], chunk_bytes[MAX_SRC], a_sbo[MAX_SRC], a_kstep[MAX_SRC];
  uint32_t g_off[MAX_GROUPS], g_lbo[MAX_GROUPS];
  int g_src[MAX_GROUPS];
  int g_dst[MAX_GROUPS][MAX_ATOMS];  // first dW row (t*Cin + c) of each atom, -1 = padding
  uint32_t dy_off, dy_atom_bytes, stage_bytes, stage_tx, ones_off;
  int n_stages, has_bias, cout_total;
  long long* dbg;   // optional clock64 timeline of CTA 0 (tools/wgrad_timeline.py); nullptr in production
  int dump_cw;   // columns staged per pass of the coalesced partial dump (0: direct strided stores)
  float* part;
  long part_stride, bias_off;
  uint32_t a_layout, b_layout, idesc;
};

__device__ __forceinline__ void wdbg(const WgradParams& P, int role, int it, int slot) {
  if (P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 16) P.dbg[(role * 16 + it) * 4 + slot] = clock64();
}

template <int PROB>
__device__ __forceinline__ void producer_loop(const WgradParams& P, int bx, int xstep, int total_tiles, uint32_t base, int nb0, int chunk0,
                                              uint64_t* bar_full, uint64_t* bar_empty) {
  for (int s = 0; s < P.n_src; ++s) tma_prefetch_desc(&P.src[PROB][s]);
  tma_prefetch_desc(&P.dymap[PROB]);
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = bx; tile < total_tiles; tile += xstep) {
    int tw = tile % P.tiles_w;
    int t2 = tile / P.tiles_w;
    int th = t2 % P.tiles_h;
    int n = t2 / P.tiles_h;
    const int h0 = th * 16, w0 = tw * 8;
    const uint32_t full = smem_u32(&bar_full[stage]);
    mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
    mbar_expect_tx(full, P.stage_tx);
    const uint32_t sa = base + (uint32_t)stage * P.stage_bytes;
    for (int s = 0; s < P.n_src; ++s)
      for (int c = 0; c < P.chunks; ++c)
        tma_load_4d(sa + P.src_off[s] + (uint32_t)c * P.chunk_bytes[s], &P.src[PROB][s], full, (chunk0 + c) * P.kc,
                    w0 + P.src_w0[s], h0 + P.src_h0[s], n);
    for (int a = 0; a < P.nb / P.kco; ++a)
      tma_load_4d(sa + P.dy_off + (uint32_t)a * P.dy_atom_bytes, &P.dymap[PROB], full, nb0 + a * P.kco, w0, h0, n);
    if (++stage == P.n_stages) { stage = 0; phase ^= 1u; }
  }
}

__global__ void __launch_bounds__(WG_THREADS, 1) umma_wgrad_kernel(const __grid_constant__ WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_done;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int nb0 = blockIdx.y * P.nb;
  const int prob = (int)blockIdx.x / P.splits, bx = (int)blockIdx.x - prob * P.splits, xstep = P.splits;
  const int cblk = (int)blockIdx.z / P.zblocks;              // block of input-channel chunks handled by this CTA
  const int chunk0 = P.chunk0 + cblk * P.chunks;
  const int dst_shift = cblk * P.chunks * P.kc;              // its rows of dW
  const int g_begin = ((int)blockIdx.z % P.zblocks) * P.groups_per_cta;
  const int g_end = min(P.n_groups, g_begin + P.groups_per_cta);
  const bool do_bias = P.has_bias && blockIdx.z == 0;
  const int total_tiles = P.n_img * P.tiles_h * P.tiles_w;

  if (tid == 0) {
    for (int s = 0; s < P.n_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), do_bias ? 5 : 1);   // MMA commit (+ the four bias-summing warps)
    }
    mbar_init(smem_u32(&bar_done), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // The tensor maps of this CTA's problem must sit at COMPILE-TIME offsets of the parameter block: with a run-time index the
      // issuing thread pays a 64-bit address computation and a move into a uniform register for every TMA instruction (16 % on
      // the many-box stride-2 layers of pix2pix) -- hence one instance of the loop per problem.
      switch (prob) {
        case 1: producer_loop<1>(P, bx, xstep, total_tiles, base, nb0, chunk0, bar_full, bar_empty); break;
        case 2: producer_loop<2>(P, bx, xstep, total_tiles, base, nb0, chunk0, bar_full, bar_empty); break;
        case 3: producer_loop<3>(P, bx, xstep, total_tiles, base, nb0, chunk0, bar_full, bar_empty); break;
        default: producer_loop<0>(P, bx, xstep, total_tiles, base, nb0, chunk0, bar_full, bar_empty); break;
      }
    }
  } else if (warp == 1) {
    // Whole warp, warp-uniform values, tcgen05 under elect_one(): UTCHMMA reads its descriptors from uniform
    // registers (see conv_umma.cu).
    {
      const uint32_t b_sbo = 8u * (uint32_t)P.kco * 2u;      // between 8-pixel rows of the dense dy tile
      const uint32_t b_kstep16 = (16u * (uint32_t)P.kco * 2u) >> 4;   // 16 pixels per MMA
      const uint64_t b_hi = make_smem_desc_hi(b_sbo, P.b_layout) << 32;
      const uint32_t b_lo0 = ((P.dy_off >> 4) & 0x3FFFu) | (((P.dy_atom_bytes >> 4) & 0x3FFFu) << 16);
      const int ng = g_end - g_begin;
      const uint32_t nb = (uint32_t)P.nb, idesc = P.idesc, stage16 = P.stage_bytes >> 4, base16 = base >> 4;
      const int n_stages = P.n_stages;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t first = 1;
      int it = 0;
      for (int tile = bx; tile < total_tiles; tile += xstep, ++it) {
        if (lane == 0) wdbg(P, 1, it, 0);
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        if (lane == 0) wdbg(P, 1, it, 1);
        const uint32_t sa16 = base16 + (uint32_t)stage * stage16;
        if (elect_one()) {
          for (int g = 0; g < ng; ++g) {
            const int s = P.g_src[g_begin + g];
            const uint32_t a_lo = sa16 + (((P.g_off[g_begin + g] >> 4) & 0x3FFFu) | (((P.g_lbo[g_begin + g] >> 4) & 0x3FFFu) << 16));
            const uint64_t a_hi = make_smem_desc_hi(P.a_sbo[s], P.a_layout) << 32;
            const uint32_t ks = P.a_kstep[s] >> 4;
            const uint32_t acc = tmem + (uint32_t)g * nb;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint64_t bd = b_hi | (uint64_t)(sa16 + b_lo0 + (uint32_t)j * b_kstep16);
              umma_f16(acc, a_hi | (uint64_t)(a_lo + (uint32_t)j * ks), bd, idesc, (first && j == 0) ? 0u : 1u);
            }
          }
          umma_commit(smem_u32(&bar_empty[stage]));
        }
        __syncwarp();
        if (lane == 0) wdbg(P, 1, it, 2);
        first = 0;
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(smem_u32(&bar_done));
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    if (q == 0 && lane == 0) wdbg(P, 2, 0, 0);
    // db[o] = sum over pixels of dy: the four epilogue warps, idle until the accumulators are final, sum the columns of
    // every dy tile straight from its pipeline stage (TMA swizzle undone in the address, one 16-byte load = 8 channels of
    // one pixel) while the MMAs consume it.  This replaced an extra accumulator against a constant tile of ones.
    float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int et = tid - 64;                     // 0..127
    const int CH = P.nb >> 3;                    // 8-channel chunks of the N block
    const int RG = CH <= 128 ? 128 / CH : 1;     // row groups
    const int chunk = et % CH, rg = et / CH;
    const bool b_active = do_bias && rg < RG;
    if (do_bias) {
      const int c0 = 8 * chunk, atom_i = c0 / P.kco, cc = c0 - atom_i * P.kco;
      const uint32_t row_pitch = (uint32_t)P.kco * 2u, mask = P.kco == 64 ? 7u : (P.kco == 32 ? 3u : 1u);
      const uint32_t col_off = P.dy_off + (uint32_t)atom_i * P.dy_atom_bytes + (uint32_t)cc * 2u;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = bx; tile < total_tiles; tile += xstep) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        if (b_active) {
          const uint32_t sa = base + (uint32_t)stage * P.stage_bytes;
#pragma unroll 4
          for (int r = rg; r < 128; r += RG) {
            const uint32_t off = col_off + (uint32_t)r * row_pitch;      // col_off is a multiple of 1024 plus < row_pitch bytes
            const uint32_t rel = off - (P.dy_off + (uint32_t)atom_i * P.dy_atom_bytes);
            const uint32_t addr = sa + P.dy_off + (uint32_t)atom_i * P.dy_atom_bytes + (rel ^ (((rel >> 7) & mask) << 4));
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(addr) : "memory");
            bsum[0] += __uint_as_float(w0 << 16); bsum[1] += __uint_as_float(w0 & 0xFFFF0000u);
            bsum[2] += __uint_as_float(w1 << 16); bsum[3] += __uint_as_float(w1 & 0xFFFF0000u);
            bsum[4] += __uint_as_float(w2 << 16); bsum[5] += __uint_as_float(w2 & 0xFFFF0000u);
            bsum[6] += __uint_as_float(w3 << 16); bsum[7] += __uint_as_float(w3 & 0xFFFF0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_empty[stage]));
        if (++stage == P.n_stages) { stage = 0; phase ^= 1u; }
      }
    }
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    if (q == 0 && lane == 0) wdbg(P, 2, 0, 1);
    if (do_bias) {     // all MMAs have completed: the pipeline buffers are free, [128][8] floats of scratch
      float* red = reinterpret_cast<float*>(base_ptr);
#pragma unroll
      for (int k = 0; k < 8; ++k) red[et * 8 + k] = bsum[k];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (b_active && rg == 0) {
        float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int g = 0; g < RG; ++g)
#pragma unroll
          for (int k = 0; k < 8; ++k) t[k] += red[(g * CH + chunk) * 8 + k];
        float* o = P.part + (long)blockIdx.x * P.part_stride + P.bias_off + nb0 + 8 * chunk;
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = t[k];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the dump below stages through the same buffers
    }
    float* part = P.part + (long)blockIdx.x * P.part_stride;
    const int atom = m / P.kc, r = m - atom * P.kc;
    if (P.dump_cw > 0) {
      // Coalesced dump: each warp stages its 32 accumulator rows x cw columns in the (now idle) pipeline buffers and
      // writes them back as whole float4 rows, so a store instruction covers contiguous 512 bytes instead of 32 lines.
      const int cw = P.dump_cw, ld = cw + 4, cw4 = cw >> 2;
      const int cw4_shift = cw4 == 16 ? 4 : (cw4 == 8 ? 3 : 2);
      const int kc_shift = P.kc == 64 ? 6 : (P.kc == 32 ? 5 : 4);
      const int at_lo = (q * 32) >> kc_shift, at_hi = (q * 32 + 16) >> kc_shift;   // the (at most two) atoms of this warp's rows
      float* stg = reinterpret_cast<float*>(base_ptr) + (size_t)q * 32 * ld;
      for (int g = g_begin; g < g_end; ++g) {
        for (int cb = 0; cb < P.nb; cb += cw) {
          for (int c0 = 0; c0 < cw; c0 += 32) {      // two 16-column loads in flight per wait
            uint32_t v0[16], v1[16];
            const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - g_begin) * P.nb + cb + c0);
            tmem_ld_32x16(ta, v0);
            if (c0 + 16 < cw) tmem_ld_32x16(ta + 16, v1);
            tmem_ld_wait();
            float4* o = reinterpret_cast<float4*>(stg + lane * ld + c0);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = make_float4(__uint_as_float(v0[4 * k]), __uint_as_float(v0[4 * k + 1]),
                                                         __uint_as_float(v0[4 * k + 2]), __uint_as_float(v0[4 * k + 3]));
            if (c0 + 16 < cw) {
#pragma unroll
              for (int k = 0; k < 4; ++k) o[4 + k] = make_float4(__uint_as_float(v1[4 * k]), __uint_as_float(v1[4 * k + 1]),
                                                               __uint_as_float(v1[4 * k + 2]), __uint_as_float(v1[4 * k + 3]));
            }
          }
          __syncwarp();
          // cw4 is 4, 8 or 16 (power of two): rows per store instruction = 32 / cw4
          const int rstep = 32 >> cw4_shift, row0 = lane >> cw4_shift, c4 = lane & (cw4 - 1);
          const int dlo = at_lo < MAX_ATOMS ? P.g_dst[g][at_lo] : -1, dhi = at_hi < MAX_ATOMS ? P.g_dst[g][at_hi] : -1;
          for (int row = row0; row < 32; row += rstep) {
            const int mm = q * 32 + row;
            const int dst = row < 16 ? dlo : dhi;
            if (dst >= 0)
              *reinterpret_cast<float4*>(part + (long)(dst + dst_shift + (mm & (P.kc - 1))) * P.cout_total + nb0 + cb + c4 * 4) =
                  *reinterpret_cast<const float4*>(stg + row * ld + c4 * 4);
          }
          __syncwarp();
        }
      }
    } else
    for (int g = g_begin; g < g_end; ++g) {
      const int dst = atom < MAX_ATOMS ? P.g_dst[g][atom] : -1;
      for (int c0 = 0; c0 < P.nb; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - g_begin) * P.nb + c0), v);
        tmem_ld_wait();
        if (dst >= 0) {
          float4* o = reinterpret_cast<float4*>(part + (long)(dst + dst_shift + r) * P.cout_total + nb0 + c0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            o[k] = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]),
                               __uint_as_float(v[4 * k + 3]));
        }
      }
    }
  }
  if (warp == 2 && lane == 0) wdbg(P, 2, 0, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// dst[i] (+)= sum_s part[s][i] over the [dw | dbias] rows of the per-CTA partials (fixed order, deterministic).
// Block = 32 float4 columns x 8 groups of splits; every thread keeps eight independent 16-byte loads in flight.
struct ReduceOut {      // one output per problem of a batched launch (blockIdx.y)
  float* dw[MAX_PROB];
  float* dbias[MAX_PROB];
  int acc[MAX_PROB];
};

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, long part_stride, int splits, const __grid_constant__ ReduceOut ro, long n_dw, int n_bias) {
  __shared__ float4 sm[8][32];
  pdl_trigger();
  pdl_wait();
  part += (long)blockIdx.y * splits * part_stride;
  float* __restrict__ dw = ro.dw[blockIdx.y];
  float* __restrict__ dbias = ro.dbias[blockIdx.y];
  const int accumulate = ro.acc[blockIdx.y];
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long col = (long)blockIdx.x * 32 + e;          // float4 column
  const long total4 = (n_dw + n_bias) >> 2;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < total4) {
    const float4* src = reinterpret_cast<const float4*>(part) + col;
    const long st4 = part_stride >> 2;
    int z = g;
    for (; z + 56 < splits; z += 64) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldcg(src + (long)(z + 8 * k) * st4);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
    }
    for (; z < splits; z += 8) {
      const float4 v = __ldcg(src + (long)z * st4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  sm[g][e] = s;
  __syncthreads();
  if (g == 0 && col < total4) {
    float4 t = sm[0][e];
#pragma unroll
    for (int k = 1; k < 8; ++k) { t.x += sm[k][e].x; t.y += sm[k][e].y; t.z += sm[k][e].z; t.w += sm[k][e].w; }
    const long i = col * 4;
    float4* d = reinterpret_cast<float4*>(i < n_dw ? dw + i : dbias + (i - n_dw));
    if (accumulate) { const float4 o = *d; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
    *d = t;
  }
}

inline int kc_for(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : 16); }
inline uint32_t layout_for(int kc) { return kc == 64 ? LAYOUT_SW128 : (kc == 32 ? LAYOUT_SW64 : LAYOUT_SW32); }
inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int pymod(int a, int b) { return a - floordiv(a, b) * b; }

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode4(dg_ctx* ctx, CUtensorMap* m, void* ptr, const uint64_t* dims, const uint64_t* strides_b, const uint32_t* box,
            int kc) {
  CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  uint32_t ones[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeFn)ctx->encode_tiled)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, (const cuuint64_t*)dims,
                                             (const cuuint64_t*)strides_b, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DG_FAIL("cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

struct Tap {
  int src, dh, dw, widx;
};
struct Lat {
  int step, h_first, w_first;
};

struct Plan {
  Tap taps[MAX_TAPS];
  Lat lat[MAX_SRC];
  int n_taps, n_src;
  int kc, kco, n_chunks, nb, gpc, zblocks, yblocks, splits, has_bias, cblocks;
  int src_split;   // 1: one launch per source (big stride-2 layers whose four halo boxes do not fit one stage)
  uint32_t slack;
};

int make_plan(const char* name, int sm_count, const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p, int has_bias,
              Plan* pl) {
  DG_REQUIRE(dg_valid(x) && dg_valid(dy) && p, "%s: null argument", name);
  DG_REQUIRE(x->dtype == DG_BF16 && dy->dtype == DG_BF16, "%s: tensor-core wgrad needs bf16 operands", name);
  DG_REQUIRE(x->c % 16 == 0 && dy->c % 16 == 0, "%s: channels must be multiples of 16", name);
  DG_REQUIRE(p->stride == 1 || p->stride == 2, "%s: stride must be 1 or 2", name);
  DG_REQUIRE(p->kh * p->kw <= MAX_TAPS, "%s: kernel too large", name);
  DG_REQUIRE(x->n == dy->n, "%s: batch mismatch", name);
  pl->n_taps = 0; pl->n_src = 0;
  if (p->stride == 1) {
    pl->lat[0] = Lat{1, 0, 0};
    pl->n_src = 1;
    for (int r = 0; r < p->kh; ++r)
      for (int s = 0; s < p->kw; ++s) pl->taps[pl->n_taps++] = Tap{0, r - p->pad_t, s - p->pad_l, r * p->kw + s};
  } else {
    DG_REQUIRE(x->h % 2 == 0 && x->w % 2 == 0, "%s: stride 2 needs even input size", name);
    int src_of[2][2] = {{-1, -1}, {-1, -1}};
    for (int r = 0; r < p->kh; ++r)
      for (int s = 0; s < p->kw; ++s) {
        int ph = pymod(r - p->pad_t, 2), pw = pymod(s - p->pad_l, 2);
        if (src_of[ph][pw] < 0) { src_of[ph][pw] = pl->n_src; pl->lat[pl->n_src++] = Lat{2, ph, pw}; }
        pl->taps[pl->n_taps++] = Tap{src_of[ph][pw], floordiv(r - p->pad_t - ph, 2), floordiv(s - p->pad_l - pw, 2), r * p->kw + s};
      }
  }
  pl->kc = kc_for(x->c);
  pl->kco = kc_for(dy->c);
  pl->n_chunks = x->c / pl->kc;
  pl->has_bias = has_bias;
  const int atoms = 128 / pl->kc;
  const int chunks_per_launch = pl->n_chunks < atoms ? pl->n_chunks : atoms;
  // halo bytes of one launch stage (all sources, the launch's chunks)
  uint32_t halo_all = 0, halo_max = 0, max_chunk = 0;
  int taps_of[MAX_SRC] = {0, 0, 0, 0}, max_taps_src = 0;
  for (int s = 0; s < pl->n_src; ++s) {
    int dh0 = 1 << 30, dh1 = -(1 << 30), dw0 = 1 << 30, dw1 = -(1 << 30);
    for (int t = 0; t < pl->n_taps; ++t)
      if (pl->taps[t].src == s) {
        ++taps_of[s];
        dh0 = pl->taps[t].dh < dh0 ? pl->taps[t].dh : dh0; dh1 = pl->taps[t].dh > dh1 ? pl->taps[t].dh : dh1;
        dw0 = pl->taps[t].dw < dw0 ? pl->taps[t].dw : dw0; dw1 = pl->taps[t].dw > dw1 ? pl->taps[t].dw : dw1;
      }
    max_taps_src = taps_of[s] > max_taps_src ? taps_of[s] : max_taps_src;
    uint32_t cb = ((uint32_t)(16 + dh1 - dh0) * (8 + dw1 - dw0) * pl->kc * 2 + 1023u) & ~1023u;
    halo_all += cb * chunks_per_launch;
    halo_max = cb * chunks_per_launch > halo_max ? cb * chunks_per_launch : halo_max;
    max_chunk = cb > max_chunk ? cb : max_chunk;
  }
  const int cout = dy->c;
  const uint32_t budget = SMEM_LIMIT - 4096;
  // padding atoms (M rows beyond the valid taps/chunks) read shared memory past their group: keep a slack region
  bool padding = false;
  if (pl->n_chunks > 1) padding = (pl->n_chunks % atoms) != 0;
  else if (atoms > 2) padding = true;
  else
    for (int s = 0; s < pl->n_src; ++s) {
      int cnt = 0;
      for (int t = 0; t < pl->n_taps; ++t) cnt += pl->taps[t].src == s;
      padding = padding || (cnt % atoms) != 0;
    }
  // Padding atoms read (and discard) shared memory BEHIND the last real chunk box of the stage: with several chunks per
  // M=128 operand that is (missing atoms) x (chunk pitch) past the end of the last stage, which must still be mapped.
  pl->slack = !padding ? 0 : (pl->n_chunks > 1 ? (uint32_t)(atoms - pl->n_chunks % atoms) * max_chunk : max_chunk);
  const uint32_t fixed = (has_bias ? 128u * pl->kc * 2 : 0u) + pl->slack;  // ones tile + slack
  int best_nb = 0, n_groups = 0;
  for (int split = 0; split < 2 && !best_nb; ++split) {
    if (split && pl->n_src == 1) break;
    const uint32_t halo = split ? halo_max : halo_all;
    // worst-case number of accumulator groups of one launch
    n_groups = split ? max_taps_src : pl->n_taps;  // one group per tap when a launch spans several chunks
    if (pl->n_chunks == 1) n_groups = split ? (max_taps_src + 1) / 2 + 1 : (pl->n_taps + 1) / 2 + pl->n_src;  // upper bound for tap stacking
    auto fits = [&](int nb) { return 2 * ((halo + 128u * nb * 2 + 1023u) & ~1023u) + fixed <= budget; };
    for (int nb = cout > 256 ? 256 : cout; nb >= 16; nb -= 16) {
      if (cout % nb != 0 || nb % pl->kco != 0 || !fits(nb)) continue;
      int gpc = 512 / nb - (has_bias ? 1 : 0);
      if (gpc >= n_groups) { best_nb = nb; break; }
    }
    if (!best_nb) {
      for (int nb = cout > 64 ? 64 : cout; nb >= 16; nb -= 16)
        if (cout % nb == 0 && nb % pl->kco == 0 && fits(nb)) { best_nb = nb; break; }
    }
    pl->src_split = split;
  }
  DG_REQUIRE(best_nb > 0, "%s: no tile configuration fits shared memory (Cin=%d Cout=%d)", name, x->c, cout);
  pl->nb = best_nb;
  pl->gpc = 512 / best_nb - (has_bias ? 1 : 0);      // accumulator groups that fit TMEM
  if (pl->gpc > MAX_GROUPS) pl->gpc = MAX_GROUPS;
  pl->zblocks = (n_groups + pl->gpc - 1) / pl->gpc;
  pl->yblocks = cout / best_nb;
  const int tiles = dy->n * ((dy->h + 15) / 16) * ((dy->w + 7) / 8);
  // Chunk blocks in the grid: a layer with more input-channel chunks than one M = 128 operand holds used to run one LAUNCH per
  // block of chunks (4 launches for 512 channels, 64 for the im2col form of pix2pix's 4x4x512 bottleneck layers, each re-dumping
  // and re-reducing its partials); single-source layers now enumerate the blocks along gridDim.z of ONE launch.
  pl->cblocks = 1;
  if (pl->n_chunks > chunks_per_launch && pl->n_chunks % chunks_per_launch == 0 && !pl->src_split) pl->cblocks = pl->n_chunks / chunks_per_launch;
  int splits = sm_count / (pl->zblocks * pl->yblocks * pl->cblocks);
  if (splits < 1) splits = 1;
  if (splits > tiles) splits = tiles;
  pl->splits = splits;
  (void)atoms;
  return 0;
}

static long long* g_wgrad_dbg = nullptr;

}  // namespace

extern "C" void dg_debug_wgrad_timeline(void* dev_buffer) { g_wgrad_dbg = (long long*)dev_buffer; }

extern "C" size_t dg_umma_conv2d_wgrad_workspace_bytes(const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p) {
  // one full fp32 dW (+ bias row) per pixel split; sized for devices of up to 160 SMs and for a call with or without a
  // bias gradient (the bias accumulator changes how many groups fit TMEM, hence the split count)
  size_t per = (size_t)p->kh * p->kw * x->c * dy->c + dy->c;
  int splits = 0;
  for (int has_bias = 0; has_bias < 2; ++has_bias)
    for (int sms = 132; sms <= 160; sms += 4) {
      Plan pl;
      if (make_plan("dg_umma_conv2d_wgrad_workspace_bytes", sms, x, dy, p, has_bias, &pl)) return 0;
      splits = pl.splits > splits ? pl.splits : splits;
    }
  return (size_t)splits * per * sizeof(float);
}

// Weight gradients of `n` layers of IDENTICAL geometry in one launch (n = 1: the plain call).  The SMs are divided among the
// problems (splits = SMs / n pixel splits each), so every CTA walks n times as many tiles and the fixed costs of a launch --
// prologue, the dump of one 147 KB fp32 partial per CTA (12.9 K of a trunk layer's 34 K cycles, profiles/wgrad_timeline_r2.log),
// the partial reduce -- are paid once per n layers.  `merged`: all problems accumulate into the SAME dW (the discriminator's
// real and fake passes, train_srgan.py:78-79): one reduction over all n x splits partials.
static int wgrad_impl(const char* name, dg_ctx* ctx, int n, const dg_tensor* const* xs, const dg_tensor* const* dys, float* const* dws,
                      float* const* dbiases, const dg_conv_params* p, const int* accumulates, void* workspace, size_t workspace_bytes,
                      void* stream) {
  const dg_tensor *x = xs[0], *dy = dys[0];
  float* dw = dws[0];
  float* dbias = dbiases ? dbiases[0] : nullptr;
  bool merged = n > 1;
  for (int i = 1; i < n; ++i) merged = merged && dws[i] == dws[0];
  Plan pl;
  const int sms = ctx->sm_count < 160 ? ctx->sm_count : 160;
  if (make_plan(name, sms / n, x, dy, p, dbias != nullptr, &pl)) return 1;
  DG_REQUIRE(dw && workspace, "%s: null output/workspace", name);
  for (int i = 0; i < n; ++i) {
    const dg_tensor *xi = xs[i], *di = dys[i];
    DG_REQUIRE(dg_valid(xi) && dg_valid(di) && dws[i], "%s: null argument (problem %d)", name, i);
    DG_REQUIRE(dg_same_shape(xi, x) && dg_same_shape(di, dy) && xi->dtype == x->dtype && di->dtype == dy->dtype &&
                   ((dbiases && dbiases[i]) != 0) == (dbias != nullptr), "%s: problem %d differs in geometry from problem 0", name, i);
    DG_REQUIRE(xi->cpitch % 8 == 0 && xi->coff % 8 == 0 && di->cpitch % 8 == 0 && di->coff % 8 == 0 &&
                   ((uintptr_t)xi->ptr % 16) == 0 && ((uintptr_t)di->ptr % 16) == 0, "%s: views not 16-byte aligned", name);
    for (int j = 0; j < i && !merged; ++j) DG_REQUIRE(dws[j] != dws[i], "%s: problems must write distinct gradients (or all the same one)", name);
  }
  if (n > 1) {
    const int atoms0 = 128 / pl.kc, cpl0 = pl.n_chunks < atoms0 ? pl.n_chunks : atoms0;
    DG_REQUIRE(!pl.src_split && (pl.n_chunks <= cpl0 || pl.cblocks > 1), "%s: this layer needs several launches and cannot be batched", name);
  }
  const int cin = x->c, cout = dy->c, kc = pl.kc, kco = pl.kco;
  const long n_dw = (long)pl.n_taps * cin * cout;
  const long part_stride = n_dw + cout;
  DG_REQUIRE(workspace_bytes >= (size_t)n * pl.splits * part_stride * sizeof(float), "%s: workspace too small", name);
  cudaStream_t st = (cudaStream_t)stream;

  // halo extents per source
  int dh_min[MAX_SRC], dh_max[MAX_SRC], dw_min[MAX_SRC], dw_max[MAX_SRC];
  bool used[MAX_SRC] = {false, false, false, false};
  for (int t = 0; t < pl.n_taps; ++t) {
    const Tap& T = pl.taps[t];
    int s = T.src;
    if (!used[s]) { dh_min[s] = dh_max[s] = T.dh; dw_min[s] = dw_max[s] = T.dw; used[s] = true; }
    dh_min[s] = T.dh < dh_min[s] ? T.dh : dh_min[s]; dh_max[s] = T.dh > dh_max[s] ? T.dh : dh_max[s];
    dw_min[s] = T.dw < dw_min[s] ? T.dw : dw_min[s]; dw_max[s] = T.dw > dw_max[s] ? T.dw : dw_max[s];
  }

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 3072);
    if (e != cudaSuccess) DG_FAIL("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    attr_set = true;
  }

  const int atoms_full = 128 / kc;
  const int chunks_per_launch = pl.n_chunks < atoms_full ? pl.n_chunks : atoms_full;
  bool first_launch = true;
  for (int chunk0 = 0; chunk0 < (pl.cblocks > 1 ? 1 : pl.n_chunks); chunk0 += chunks_per_launch)
  for (int ssel = pl.src_split ? 0 : -1; ssel < (pl.src_split ? pl.n_src : 0); ++ssel) {   // -1: all sources in one launch
    const int chunks = (pl.n_chunks - chunk0) < chunks_per_launch ? (pl.n_chunks - chunk0) : chunks_per_launch;
    WgradParams P;
    memset(&P, 0, sizeof(P));
    P.splits = pl.splits;
    P.n_src = ssel < 0 ? pl.n_src : 1; P.kc = kc; P.kco = kco; P.nb = pl.nb; P.chunks = chunks; P.chunk0 = chunk0;
    P.tiles_h = (dy->h + 15) / 16; P.tiles_w = (dy->w + 7) / 8; P.n_img = dy->n;
    P.cout_total = cout; P.part = (float*)workspace; P.part_stride = part_stride; P.bias_off = n_dw;
    P.has_bias = (dbias != nullptr && first_launch) ? 1 : 0;
    first_launch = false;
    P.a_layout = layout_for(kc); P.b_layout = layout_for(kco);
    P.idesc = make_idesc_bf16(128, pl.nb, 1, 1);
    uint32_t off = 0, tx = 0;
    int WWs[MAX_SRC], slot_of[MAX_SRC] = {-1, -1, -1, -1};
    int n_slots = 0;
    for (int s = 0; s < pl.n_src; ++s) {
      if (ssel >= 0 && s != ssel) continue;
      const int si = n_slots++;      // slot of this source in the launch's parameter arrays
      slot_of[s] = si;
      const Lat& L = pl.lat[s];
      const int HH = 16 + dh_max[s] - dh_min[s], WW = 8 + dw_max[s] - dw_min[s];
      WWs[s] = WW;
      const int vh = (x->h - L.h_first + L.step - 1) / L.step, vw = (x->w - L.w_first + L.step - 1) / L.step;
      uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)vw, (uint64_t)vh, (uint64_t)x->n};
      uint64_t strides[3] = {(uint64_t)x->cpitch * 2 * L.step, (uint64_t)x->cpitch * 2 * x->w * L.step,
                             (uint64_t)x->cpitch * 2 * x->w * x->h};
      uint32_t box[4] = {(uint32_t)kc, (uint32_t)WW, (uint32_t)HH, 1};
      for (int pr = 0; pr < n; ++pr) {
        const dg_tensor* xp = xs[pr];
        uint64_t st_p[3] = {(uint64_t)xp->cpitch * 2 * L.step, (uint64_t)xp->cpitch * 2 * xp->w * L.step, (uint64_t)xp->cpitch * 2 * xp->w * xp->h};
        char* ptr = (char*)xp->ptr + ((size_t)xp->coff + ((size_t)L.h_first * xp->w + L.w_first) * xp->cpitch) * 2;
        if (encode4(ctx, &P.src[pr][si], ptr, dims, st_p, box, kc)) return 1;
      }
      P.src_h0[si] = dh_min[s]; P.src_w0[si] = dw_min[s];
      uint32_t hb = (uint32_t)HH * WW * kc * 2;
      P.chunk_bytes[si] = (hb + 1023u) & ~1023u;
      P.src_off[si] = off;
      P.a_sbo[si] = (uint32_t)WW * kc * 2;
      P.a_kstep[si] = 2u * WW * kc * 2;
      off += P.chunk_bytes[si] * chunks;
      tx += hb * chunks;
    }
    {  // dy tile: dense 16 x 8 pixels, nb/kco channel atoms
      uint64_t dims[4] = {(uint64_t)dy->c, (uint64_t)dy->w, (uint64_t)dy->h, (uint64_t)dy->n};
      uint64_t strides[3] = {(uint64_t)dy->cpitch * 2, (uint64_t)dy->cpitch * 2 * dy->w, (uint64_t)dy->cpitch * 2 * dy->w * dy->h};
      uint32_t box[4] = {(uint32_t)kco, 8, 16, 1};
      for (int pr = 0; pr < n; ++pr) {
        const dg_tensor* dp = dys[pr];
        uint64_t st_p[3] = {(uint64_t)dp->cpitch * 2, (uint64_t)dp->cpitch * 2 * dp->w, (uint64_t)dp->cpitch * 2 * dp->w * dp->h};
        if (encode4(ctx, &P.dymap[pr], (char*)dp->ptr + (size_t)dp->coff * 2, dims, st_p, box, kco)) return 1;
      }
      P.dy_off = off;
      P.dy_atom_bytes = 128u * kco * 2;
      off += P.dy_atom_bytes * (pl.nb / kco);
      tx += P.dy_atom_bytes * (pl.nb / kco);
    }
    P.stage_bytes = (off + 1023u) & ~1023u;
    P.stage_tx = tx;
    // groups
    int ng = 0;
    if (pl.n_chunks > 1) {
      for (int t = 0; t < pl.n_taps; ++t) {
        const Tap& T = pl.taps[t];
        int s = T.src;
        if (slot_of[s] < 0) continue;
        DG_REQUIRE(ng < MAX_GROUPS, "%s: too many groups", name);
        P.g_src[ng] = slot_of[s];
        P.g_off[ng] = P.src_off[slot_of[s]] + (uint32_t)((T.dh - dh_min[s]) * WWs[s] + (T.dw - dw_min[s])) * kc * 2;
        P.g_lbo[ng] = P.chunk_bytes[slot_of[s]];
        for (int a = 0; a < MAX_ATOMS; ++a) P.g_dst[ng][a] = a < chunks ? T.widx * cin + (chunk0 + a) * kc : -1;
        ++ng;
      }
    } else {
      // stack taps of the same source whose windows are uniformly spaced
      bool done[MAX_TAPS] = {false};
      for (int t = 0; t < pl.n_taps; ++t) {
        if (done[t]) continue;
        const Tap& T = pl.taps[t];
        const int s = T.src;
        if (slot_of[s] < 0) continue;
        auto row = [&](const Tap& U) { return (U.dh - dh_min[s]) * WWs[s] + (U.dw - dw_min[s]); };
        DG_REQUIRE(ng < MAX_GROUPS, "%s: too many groups", name);
        P.g_src[ng] = slot_of[s];
        P.g_off[ng] = P.src_off[slot_of[s]] + (uint32_t)row(T) * kc * 2;
        for (int a = 0; a < MAX_ATOMS; ++a) P.g_dst[ng][a] = -1;
        P.g_dst[ng][0] = T.widx * cin;
        done[t] = true;
        int delta = 0, count = 1, last = row(T);
        for (int u = t + 1; u < pl.n_taps && count < atoms_full; ++u) {
          if (done[u] || pl.taps[u].src != s) continue;
          int d = row(pl.taps[u]) - last;
          if (d <= 0) continue;
          if (count == 1) delta = d;
          if (d != delta) continue;
          P.g_dst[ng][count++] = pl.taps[u].widx * cin;
          done[u] = true;
          last = row(pl.taps[u]);
        }
        P.g_lbo[ng] = (uint32_t)(delta > 0 ? delta : 1) * kc * 2;
        ++ng;
      }
    }
    P.n_groups = ng;
    P.groups_per_cta = pl.gpc;
    const int zblocks = (ng + pl.gpc - 1) / pl.gpc;
    P.zblocks = zblocks;
    // shared memory: stages + ones tile + slack for padding atoms
    const uint32_t ones_bytes = P.has_bias ? 128u * kc * 2 : 0u;
    const uint32_t slack = pl.slack;
    const uint32_t budget = SMEM_LIMIT - 4096;
    DG_REQUIRE(2 * P.stage_bytes + ones_bytes + slack <= budget, "%s: tile does not fit shared memory", name);
    int n_stages = (int)((budget - ones_bytes - slack) / P.stage_bytes);
    if (n_stages > MAX_STAGES) n_stages = MAX_STAGES;
    P.n_stages = n_stages;
    P.dbg = g_wgrad_dbg;
    P.ones_off = (uint32_t)n_stages * P.stage_bytes;
    {
      int cw = pl.nb % 64 == 0 ? 64 : (pl.nb % 32 == 0 ? 32 : 16);   // must divide the N block (e.g. 112 = 7 x 16)
      static const char* dbg_direct = getenv("DG_DEBUG_WGRAD_DIRECT");   // experiments only
      P.dump_cw = (!dbg_direct && (size_t)4 * 32 * (cw + 4) * sizeof(float) <= (size_t)n_stages * P.stage_bytes) ? cw : 0;
    }
    const uint32_t smem = P.ones_off + ones_bytes + slack + 1024;
    dim3 grid(n * pl.splits, pl.yblocks, zblocks * pl.cblocks);
    dg_pdl_launch(umma_wgrad_kernel, grid, dim3(WG_THREADS), smem, st, P);
    DG_CHECK_LAUNCH(name);
  }
  const long total = n_dw + (dbias ? cout : 0);
  DG_REQUIRE(n_dw % 4 == 0 && total % 4 == 0 && part_stride % 4 == 0 && ((uintptr_t)dw % 16) == 0 && (!dbias || ((uintptr_t)dbias % 16) == 0),
             "dg_umma_conv2d_wgrad: gradient buffers must be 16-byte aligned");
  ReduceOut ro;
  memset(&ro, 0, sizeof(ro));
  const int n_out = merged ? 1 : n;
  for (int i = 0; i < n_out; ++i) { ro.dw[i] = dws[i]; ro.dbias[i] = dbiases ? dbiases[i] : nullptr; ro.acc[i] = accumulates[i]; }
  for (int i = 0; i < n_out; ++i)
    DG_REQUIRE(((uintptr_t)ro.dw[i] % 16) == 0 && (!ro.dbias[i] || ((uintptr_t)ro.dbias[i] % 16) == 0), "%s: gradient buffers must be 16-byte aligned", name);
  dg_pdl_launch(wgrad_reduce_kernel, dim3((unsigned)((total / 4 + 31) / 32), (unsigned)n_out), dim3(256), 0, st, (const float*)workspace, part_stride,
                merged ? n * pl.splits : pl.splits, ro, n_dw, dbias ? cout : 0);
  DG_CHECK_LAUNCH("dg_umma_conv2d_wgrad(reduce)");
  return 0;
}

extern "C" int dg_umma_conv2d_wgrad(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* dy, float* dw, float* dbias,
                                    const dg_conv_params* p, int accumulate, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  DG_REQUIRE(dg_valid(x) && dg_valid(dy) && p, "dg_umma_conv2d_wgrad: null argument");
  return wgrad_impl("dg_umma_conv2d_wgrad", ctx, 1, &x, &dy, &dw, &dbias, p, &accumulate, workspace, workspace_bytes, stream);
}

extern "C" size_t dg_umma_conv2d_wgrad_batch_workspace_bytes(int n, const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p) {
  if (n < 1 || n > MAX_PROB) return 0;
  // n x splits(SMs / n) partials; for small maps the split count is capped by the tile count, not by the SMs: n times the bound of one problem
  return (size_t)n * dg_umma_conv2d_wgrad_workspace_bytes(x, dy, p);
}

// 1 when dg_umma_conv2d_wgrad_batch takes n layers of this geometry in one launch (a single-launch tile configuration exists).
extern "C" int dg_umma_conv2d_wgrad_batch_supported(dg_ctx* ctx, int n, const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p) {
  if (n < 1 || n > MAX_PROB || !dg_valid(x) || !dg_valid(dy) || !p) return 0;
  if (x->dtype != DG_BF16 || dy->dtype != DG_BF16 || x->c % 16 != 0 || dy->c % 16 != 0 || (p->stride != 1 && p->stride != 2) ||
      p->kh * p->kw > MAX_TAPS || x->n != dy->n || (p->stride == 2 && (x->h % 2 != 0 || x->w % 2 != 0)))
    return 0;
  Plan pl;
  const int sms = ctx->sm_count < 160 ? ctx->sm_count : 160;
  for (int has_bias = 0; has_bias < 2; ++has_bias) {
    if (make_plan("dg_umma_conv2d_wgrad_batch_supported", sms / n, x, dy, p, has_bias, &pl)) return 0;
    const int atoms0 = 128 / pl.kc, cpl0 = pl.n_chunks < atoms0 ? pl.n_chunks : atoms0;
    if (pl.src_split || !(pl.n_chunks <= cpl0 || pl.cblocks > 1)) return 0;
  }
  return 1;
}

extern "C" int dg_umma_conv2d_wgrad_batch(dg_ctx* ctx, int n, const dg_tensor* const* x, const dg_tensor* const* dy, float* const* dw,
                                          float* const* dbias, const dg_conv_params* p, const int* accumulate, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  DG_REQUIRE(n >= 1 && n <= MAX_PROB && x && dy && dw && accumulate && p, "dg_umma_conv2d_wgrad_batch: bad argument (1 <= n <= %d)", MAX_PROB);
  return wgrad_impl("dg_umma_conv2d_wgrad_batch", ctx, n, x, dy, dw, dbias, p, accumulate, workspace, workspace_bytes, stream);
}
