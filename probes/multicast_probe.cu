// Hardware probe for the NEXT step of the conv kernels: how long do all CTAs of a launch take
// to pull the same resident weight block (9 x [64 x 64] bf16 = 73.7 KB) out of L2 into shared memory
//   (a) every CTA on its own (what umma_conv_kernel does today: ~3500 cycles, tools/conv_timeline.py prologue marks), and
//   (b) in thread-block clusters of 2 / 4 CTAs where each CTA loads 1/2 or 1/4 of the taps with TMA MULTICAST to all CTAs of
//       the cluster (every CTA's mbarrier expects the full byte count; the L2 reads shrink by the cluster size).
// Prints min / mean / max cycles from kernel entry to "all weights landed" over the CTAs.  Stand-alone: not on the product
// path, not built by __graft_entry__.build().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -o probes/multicast_probe probes/multicast_probe.cu
//   ./probes/multicast_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int TAPS = 9, ROWS = 64, KC = 64;             // one tap block = 64 rows x 64 bf16 = 8 KB, 128-byte swizzle
constexpr uint32_t BLOCK_BYTES = ROWS * KC * 2;
constexpr uint32_t SMEM_BYTES = 200 * 1024;             // as much as the conv kernel: one CTA per SM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
               " [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CS = cluster size (1: no cluster, plain loads).  out[cta] = cycles from entry to all TAPS blocks in shared memory.
template <int CS>
__global__ void __launch_bounds__(128, 1) weights_kernel(const __grid_constant__ CUtensorMap wmap, long long* out, unsigned* check) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const long long t0 = clock64();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(b, TAPS * BLOCK_BYTES);       // every CTA receives ALL taps, from itself and from its peers
  }
  if (CS > 1) cluster_sync_all();                // peers' barriers are armed before anybody multicasts into them
  else __syncthreads();
  if (threadIdx.x == 0) {
    if (CS == 1) {
      for (int t = 0; t < TAPS; ++t) tma_load_2d(base + t * BLOCK_BYTES, &wmap, b, 0, t * ROWS);
    } else {
      const uint32_t rank = cluster_ctarank();
      for (int t = (int)rank; t < TAPS; t += CS)   // same smem offset and same barrier offset in every CTA of the cluster
        tma_load_2d_mc(base + t * BLOCK_BYTES, &wmap, b, 0, t * ROWS, (uint16_t)((1u << CS) - 1u));
    }
  }
  while (!mbar_try_wait(b, 0)) {
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) {
    out[blockIdx.x] = t1 - t0;
    // checksum of the first word of every tap block (all CTAs must see the same data)
    unsigned s = 0;
    for (int t = 0; t < TAPS; ++t) s += *reinterpret_cast<const unsigned*>(smem_raw + (base - smem_u32(smem_raw)) + t * BLOCK_BYTES);
    check[blockIdx.x] = s;
  }
  if (CS > 1) cluster_sync_all();                // nobody exits while a peer may still multicast into its shared memory
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CS>
static void run(const CUtensorMap& map, int ctas, long long* d_out, unsigned* d_chk, const char* label) {
  CK(cudaFuncSetAttribute(weights_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas / CS * CS); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = CS > 1 ? 1 : 0;
  const int n = (int)cfg.gridDim.x;
  long long* h = (long long*)malloc(n * sizeof(long long));
  unsigned* hc = (unsigned*)malloc(n * sizeof(unsigned));
  for (int rep = 0; rep < 4; ++rep) {              // rep 0 is cold (weights come from DRAM), the rest from L2
    CK(cudaLaunchKernelEx(&cfg, weights_kernel<CS>, map, d_out, d_chk));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, n * sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc, d_chk, n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    long long mn = h[0], mx = h[0], sum = 0;
    bool same = true;
    for (int i = 0; i < n; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; sum += h[i]; same = same && hc[i] == hc[0]; }
    printf("%-28s rep %d: %d CTAs, cycles min %lld mean %lld max %lld, data %s\n", label, rep, n, mn, sum / n, mx, same ? "identical" : "MISMATCH");
  }
  free(h); free(hc);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeFn encode = (EncodeFn)fn;
  const size_t n = (size_t)TAPS * ROWS * KC;
  __nv_bfloat16* h = (__nv_bfloat16*)malloc(n * 2);
  for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16((float)((i * 2654435761u) % 1000) / 1000.f);
  __nv_bfloat16* d;
  CK(cudaMalloc(&d, n * 2));
  CK(cudaMemcpy(d, h, n * 2, cudaMemcpyHostToDevice));
  CUtensorMap map;
  uint64_t dims[2] = {KC, (uint64_t)TAPS * ROWS}, strides[1] = {KC * 2};
  uint32_t box[2] = {KC, ROWS}, ones[2] = {1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  long long* d_out; unsigned* d_chk;
  CK(cudaMalloc(&d_out, 1024 * sizeof(long long)));
  CK(cudaMalloc(&d_chk, 1024 * sizeof(unsigned)));
  printf("%d SMs; weight block %u bytes per CTA\n", sms, TAPS * BLOCK_BYTES);
  run<1>(map, sms, d_out, d_chk, "independent loads");
  run<2>(map, sms, d_out, d_chk, "cluster 2, multicast");
  run<4>(map, sms, d_out, d_chk, "cluster 4, multicast");
  return 0;
}
