// Shared host/device helpers for the dg_b200 library.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dg_b200.h"

struct dg_ctx {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  void* encode_tiled;  // cuTensorMapEncodeTiled entry point (driver API, resolved at init)
  unsigned* tickets;   // device scratch: 'last block finishes the reduction' counters (zero between kernels; one stream per ctx)
};

void dg_set_error(const char* fmt, ...);

#define DG_FAIL(...)            \
  do {                          \
    dg_set_error(__VA_ARGS__);  \
    return 1;                   \
  } while (0)

#define DG_CHECK_LAUNCH(name)                                                     \
  do {                                                                            \
    cudaError_t e_ = cudaGetLastError();                                          \
    if (e_ != cudaSuccess) DG_FAIL("%s: launch failed: %s", name, cudaGetErrorString(e_)); \
  } while (0)

#define DG_REQUIRE(cond, ...)       \
  do {                              \
    if (!(cond)) DG_FAIL(__VA_ARGS__); \
  } while (0)

// ---- programmatic dependent launch (PDL): a kernel launched through dg_pdl_launch() may be scheduled while its
// predecessor in the stream is still draining; it must execute pdl_wait() before touching global memory (this waits
// for the predecessor's completion and memory flush, so ordering is exactly stream order) and should call
// pdl_trigger() at its top so that ITS successor can be scheduled early.  Launch latency and the kernel prologue
// (barrier init, TMEM allocation, weight loads) then overlap the predecessor's tail; ~500 kernels per train step.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
extern int g_dg_pdl;   // 1 enables the launch attribute (DG_PDL=1, default off: measured slower inside the step graph), api.cu
template <typename... KArgs, typename... Args>
static inline cudaError_t dg_pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_dg_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- cooperative launch: kernels that contain a grid-wide barrier (the one-launch BatchNorm forward / backward, the
// convolution with a fused BatchNorm phase) need EVERY block resident at the same time.  dg_coresident() checks that the
// grid fits the device at this block size / shared memory (an empty device: occupancy x SM count), and the launch carries
// cudaLaunchAttributeCooperative, so that the driver schedules the grid only when all of its blocks can be resident --
// other work sharing the device (side-stream weight gradients, another context, MPS) can delay it but cannot leave
// part of the grid spinning on blocks that never start.  Callers fall back to their multi-kernel path when the check fails.
extern int g_dg_coop;   // 0 drops the attribute (DG_COOP=0; debugging only), api.cu
template <typename K>
static inline bool dg_coresident(K kernel, int block_threads, size_t smem, long grid_blocks, int sm_count) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block_threads, smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return (long)per_sm * sm_count >= grid_blocks;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t dg_coop_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_dg_coop ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- element access for the two storage types
template <typename T>
__device__ __forceinline__ float ld_f(const T* p);
template <>
__device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__device__ __forceinline__ void st_f(T* p, float v);
template <>
__device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// value as it will read back from storage of type T (bf16 rounding made explicit)
template <typename T>
__device__ __forceinline__ float round_to(float v);
template <>
__device__ __forceinline__ float round_to<float>(float v) { return v; }
template <>
__device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16(v)); }

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  switch (act) {
    case DG_ACT_RELU: return v > 0.f ? v : 0.f;
    case DG_ACT_LRELU: return v >= 0.f ? v : alpha * v;
    case DG_ACT_TANH: return tanhf(v);
    case DG_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// Counter-based Bernoulli(0.5) keep decision shared with oracle/ops_np.py:dropout_keep_mask.
__device__ __forceinline__ bool dropout_keep(uint32_t seed, uint32_t idx) {
  uint32_t x = idx ^ seed;
  x = (x ^ (x >> 16)) * 0x7FEB352Du;
  x = (x ^ (x >> 15)) * 0x846CA68Bu;
  x = x ^ (x >> 16);
  return (x >> 31) == 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dispatch a lambda-like macro over the (input, output) storage types of two tensors
#define DG_DISPATCH_2(tin, tout, NAME, ...)                                               \
  do {                                                                                    \
    if ((tin) == DG_F32 && (tout) == DG_F32) { using TI = float; using TO = float; __VA_ARGS__ }                 \
    else if ((tin) == DG_F32 && (tout) == DG_BF16) { using TI = float; using TO = __nv_bfloat16; __VA_ARGS__ }   \
    else if ((tin) == DG_BF16 && (tout) == DG_F32) { using TI = __nv_bfloat16; using TO = float; __VA_ARGS__ }   \
    else if ((tin) == DG_BF16 && (tout) == DG_BF16) { using TI = __nv_bfloat16; using TO = __nv_bfloat16; __VA_ARGS__ } \
    else DG_FAIL("%s: unsupported dtype", NAME);                                          \
  } while (0)

#define DG_DISPATCH_1(t, NAME, ...)                                     \
  do {                                                                  \
    if ((t) == DG_F32) { using T = float; __VA_ARGS__ }                 \
    else if ((t) == DG_BF16) { using T = __nv_bfloat16; __VA_ARGS__ }   \
    else DG_FAIL("%s: unsupported dtype", NAME);                        \
  } while (0)

static inline bool dg_same_shape(const dg_tensor* a, const dg_tensor* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}
static inline int64_t dg_pixels(const dg_tensor* a) { return (int64_t)a->n * a->h * a->w; }
static inline bool dg_valid(const dg_tensor* a) {
  return a && a->ptr && a->n > 0 && a->h > 0 && a->w > 0 && a->c > 0 && a->cpitch >= a->coff + a->c && a->coff >= 0;
}
