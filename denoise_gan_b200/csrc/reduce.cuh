// Per-channel reduction skeleton shared by the BatchNorm, PReLU and depthwise-conv kernels:
// each thread owns up to MAX_CPT channels, strides over pixels, rows are combined through
// shared memory and every block writes one partial [NV][C]; a finalize kernel sums the
// partials in block order (deterministic).
#pragma once
#include "dg_common.cuh"

namespace dgred {

constexpr int RED_THREADS = 256;
constexpr int MAX_CPT = 4;  // channels per thread in the per-channel reductions (C <= 1024)

__host__ __device__ inline int red_lanes(int C) { return C < RED_THREADS ? C : RED_THREADS; }

static inline int red_blocks(long P, int C, int sm_count) {
  int L = red_lanes(C);
  int R = RED_THREADS / L;
  long want = (P + R - 1) / R;
  long cap = (long)sm_count * 8;
  return (int)(want < cap ? want : cap);
}

// Generic per-channel reduction skeleton: F(p, c, acc[NV]) accumulates NV values per (pixel, channel).
template <int NV, typename F>
__device__ __forceinline__ void channel_reduce(long P, int C, float* __restrict__ partial, F f) {
  extern __shared__ float red_smem[];
  const int L = red_lanes(C), R = RED_THREADS / L;
  const int lane = threadIdx.x % L, row = threadIdx.x / L;
  float acc[MAX_CPT][NV];
#pragma unroll
  for (int i = 0; i < MAX_CPT; ++i)
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[i][v] = 0.f;
  if (row < R) {
    for (long p = (long)blockIdx.x * R + row; p < P; p += (long)gridDim.x * R) {
#pragma unroll
      for (int i = 0; i < MAX_CPT; ++i) {
        int c = lane + i * L;
        if (c < C) f(p, c, acc[i]);
      }
    }
  }
  // reduce over rows through shared memory: layout [R][NV][C]
  for (int i = 0; i < MAX_CPT; ++i) {
    int c = lane + i * L;
    if (row < R && c < C)
      for (int v = 0; v < NV; ++v) red_smem[(row * NV + v) * C + c] = acc[i][v];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NV * C; e += RED_THREADS) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += red_smem[r * NV * C + e];
    partial[(long)blockIdx.x * NV * C + e] = s;  // [block][NV][C]
  }
}


}  // namespace dgred
