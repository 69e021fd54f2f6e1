// Hardware probe for the tcgen05/TMA conventions the convolution kernels rely on.
// Each case stages operand tiles in shared memory with TMA, issues a few tcgen05.mma
// instructions from hand-built descriptors and compares the TMEM accumulator with a CPU
// expectation computed from the INTENDED semantics.  The result table (PASS/FAIL per
// hypothesis) is what DESIGN.md cites for the descriptor arithmetic used in conv_umma.cu.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probes/umma_probe probes/umma_probe.cu
// Run  :  probes/umma_probe > gpurun_out/umma_probe.log
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../denoise_gan_b200/csrc/sm100.cuh"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

struct Maps {
  CUtensorMap m[8];
};

struct ProbeCase {
  int n_loads;
  struct {
    int map, dims, c0, c1, c2, c3;
    uint32_t smem_off;
  } loads[8];
  uint32_t tx_bytes;
  int n_mma;
  struct {
    uint32_t a_off, b_off;
  } mma[48];
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_layout, b_layout;
  int a_bo_mode, b_bo_mode;
  uint32_t idesc;
  int n_cols;
  uint32_t dump_bytes;
  int repeat;  // timing mode: issue the MMA list this many times and report cycles
  int n_acc;   // timing mode: round-robin over this many independent accumulators (0/1 = one dependent chain)
  uint32_t acc_stride;
};

constexpr int SMEM_DATA = 160 * 1024;

__device__ __forceinline__ bool bounded_wait(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i)
    if (sm100::mbar_try_wait(bar, parity)) return true;
  return false;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ Maps maps, const ProbeCase* __restrict__ pcp,
             float* __restrict__ out, uint32_t* __restrict__ dump, int* __restrict__ status) {
  using namespace sm100;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int ok_flag;
  const ProbeCase& pc = *pcp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* data = smem_raw + (base - raw);
  const uint32_t bar_load = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

  for (int i = tid; i < SMEM_DATA / 4; i += 128) reinterpret_cast<uint32_t*>(data)[i] = 0;
  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
    ok_flag = 1;
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    mbar_expect_tx(bar_load, pc.tx_bytes);
    for (int i = 0; i < pc.n_loads; ++i) {
      const auto& L = pc.loads[i];
      if (L.dims == 2)
        tma_load_2d(base + L.smem_off, &maps.m[L.map], bar_load, L.c0, L.c1);
      else
        tma_load_4d(base + L.smem_off, &maps.m[L.map], bar_load, L.c0, L.c1, L.c2, L.c3);
    }
    if (!bounded_wait(bar_load, 0)) {
      ok_flag = 0;
      status[0] = 1;  // TMA never completed (tx byte mismatch?)
    } else {
      tc_fence_after();
      const int reps = pc.repeat > 0 ? pc.repeat : 1;
      long long t0 = clock64();
      if (pc.repeat > 0) {
        // timing mode: descriptors live in registers, the issue loop is nothing but tcgen05.mma
        uint64_t adq[8], bdq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int k = i < pc.n_mma ? i : 0;
          adq[i] = make_smem_desc(base + pc.mma[k].a_off, pc.a_lbo, pc.a_sbo, pc.a_layout, 0);
          bdq[i] = make_smem_desc(base + pc.mma[k].b_off, pc.b_lbo, pc.b_sbo, pc.b_layout, 0);
        }
        const uint32_t idesc = pc.idesc;
        const int nm = pc.n_mma;
        uint32_t dq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dq[i] = tmem + (pc.n_acc > 1 ? (uint32_t)(i % pc.n_acc) * pc.acc_stride : 0u);
        t0 = clock64();
        if (nm == 8) {
          for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) umma_f16(dq[i], adq[i], bdq[i], idesc, 1u);
          }
        } else {
          for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
            for (int i = 0; i < 4; ++i) umma_f16(dq[i], adq[i], bdq[i], idesc, 1u);
          }
        }
      } else
      for (int rep = 0; rep < reps; ++rep)
      for (int i = 0; i < pc.n_mma; ++i) {
        uint32_t a_addr = base + pc.mma[i].a_off, b_addr = base + pc.mma[i].b_off;
        uint32_t a_bo = pc.a_bo_mode ? ((a_addr >> 7) & 7) : 0;
        uint32_t b_bo = pc.b_bo_mode ? ((b_addr >> 7) & 7) : 0;
        uint64_t ad = make_smem_desc(a_addr, pc.a_lbo, pc.a_sbo, pc.a_layout, a_bo);
        uint64_t bd = make_smem_desc(b_addr, pc.b_lbo, pc.b_sbo, pc.b_layout, b_bo);
        umma_f16(tmem, ad, bd, pc.idesc, (i > 0 || rep > 0) ? 1u : 0u);
      }
      long long t1 = clock64();
      umma_commit(bar_mma);
      if (!bounded_wait(bar_mma, 0)) {
        ok_flag = 0;
        status[0] = 2;  // MMA never committed
      }
      long long t2 = clock64();
      status[1] = (int)(t1 - t0);
      status[2] = (int)(t2 - t0);
    }
  }
  __syncthreads();
  tc_fence_after();
  if (ok_flag) {
    for (int c = 0; c < pc.n_cols; c += 16) {
      uint32_t v[16];
      tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * pc.n_cols + c + j] = __uint_as_float(v[j]);
    }
  }
  for (uint32_t i = tid; i < pc.dump_bytes / 4; i += 128) dump[i] = reinterpret_cast<uint32_t*>(data)[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;

static CUtensorMap make_map(void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, const uint32_t* estr, CUtensorMapSwizzle sw) {
  CUtensorMap m;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, (const cuuint64_t*)dims,
                        (const cuuint64_t*)strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(3);
  }
  return m;
}

static float gen(int salt, long r, int c) {
  uint64_t x = (uint64_t)r * 0x9E3779B97F4A7C15ull + (uint64_t)c * 0xC2B2AE3D27D4EB4Full + (uint64_t)salt * 0x165667B19E3779F9ull;
  x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
  return (float)((long)(x % 7) - 3);
}

struct HostTensor {
  std::vector<__nv_bfloat16> h;
  __nv_bfloat16* d = nullptr;
  long rows;
  int C;
  int salt;
  float at(long r, int c) const { return gen(salt, r, c); }
  void init(long rows_, int C_, int salt_) {
    rows = rows_; C = C_; salt = salt_;
    h.resize(rows * C);
    for (long r = 0; r < rows; ++r)
      for (int c = 0; c < C; ++c) h[r * C + c] = __float2bfloat16(gen(salt, r, c));
    CK(cudaMalloc(&d, h.size() * 2));
    CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  }
};

int main() {
  CK(cudaSetDevice(0));
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  g_encode = (EncodeFn)fn;

  // tensors
  const int IH = 24, IW = 24, IN = 2;
  HostTensor X;   X.init((long)IN * IH * IW, 64, 1);    // NHWC image, C=64 (also a [1152][64] matrix)
  HostTensor Bm;  Bm.init(512, 64, 2);                  // weights / dy rows
  HostTensor T128; T128.init(256, 128, 3);              // 128-channel rows
  HostTensor T32; T32.init(256, 32, 4);
  HostTensor B32; B32.init(256, 32, 5);

  Maps maps;
  memset(&maps, 0, sizeof(maps));
  uint32_t ones[4] = {1, 1, 1, 1};
  {  // 0: X as 2D [rows][64], box 64 x 256
    uint64_t d[2] = {64, (uint64_t)X.rows}, s[1] = {128};
    uint32_t b[2] = {64, 256};
    maps.m[0] = make_map(X.d, 2, d, s, b, ones, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  {  // 1: Bm 2D [512][64], box 64 x 64
    uint64_t d[2] = {64, 512}, s[1] = {128};
    uint32_t b[2] = {64, 64};
    maps.m[1] = make_map(Bm.d, 2, d, s, b, ones, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  {  // 2: X as 4D, box (64,10,18,1)
    uint64_t d[4] = {64, IW, IH, IN}, s[3] = {128, 128ull * IW, 128ull * IW * IH};
    uint32_t b[4] = {64, 10, 18, 1};
    maps.m[2] = make_map(X.d, 4, d, s, b, ones, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  {  // 3: X as 4D, box (64,16,18,1)
    uint64_t d[4] = {64, IW, IH, IN}, s[3] = {128, 128ull * IW, 128ull * IW * IH};
    uint32_t b[4] = {64, 16, 18, 1};
    maps.m[3] = make_map(X.d, 4, d, s, b, ones, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  {  // 4: T128 2D [256][128], box (64 ch, 192 rows)
    uint64_t d[2] = {128, 256}, s[1] = {256};
    uint32_t b[2] = {64, 192};
    maps.m[4] = make_map(T128.d, 2, d, s, b, ones, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  {  // 5: Bm 2D, box 64 x 256 (N=256 / dy rows)
    uint64_t d[2] = {64, 512}, s[1] = {128};
    uint32_t b[2] = {64, 256};
    maps.m[5] = make_map(Bm.d, 2, d, s, b, ones, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  {  // 6: T32 / B32 share geometry; map 6 = T32 box (32,128) SW64 ; map 7 = B32 box (32,64) SW64
    uint64_t d[2] = {32, 256}, s[1] = {64};
    uint32_t b[2] = {32, 128};
    maps.m[6] = make_map(T32.d, 2, d, s, b, ones, CU_TENSOR_MAP_SWIZZLE_64B);
    uint32_t b2[2] = {32, 64};
    maps.m[7] = make_map(B32.d, 2, d, s, b2, ones, CU_TENSOR_MAP_SWIZZLE_64B);
  }
  // stride-2 map swaps into slot 3 later
  CUtensorMap map_s2;
  {
    uint64_t d[4] = {64, IW, IH, IN}, s[3] = {128, 128ull * IW, 128ull * IW * IH};
    uint32_t b[4] = {64, 16, 8, 1};
    uint32_t es[4] = {1, 2, 2, 1};
    map_s2 = make_map(X.d, 4, d, s, b, es, CU_TENSOR_MAP_SWIZZLE_128B);
  }

  float* d_out; uint32_t* d_dump; int* d_status; ProbeCase* d_pc;
  CK(cudaMalloc(&d_out, 128 * 512 * 4));   // timing cases dump up to 512 accumulator columns; the checked cases use <= 256
  CK(cudaMalloc(&d_dump, SMEM_DATA));
  CK(cudaMalloc(&d_status, 16));
  CK(cudaMalloc(&d_pc, sizeof(ProbeCase)));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DATA + 1024));

  std::vector<float> out(128 * 256);
  std::vector<uint32_t> dump(SMEM_DATA / 4);
  int n_pass = 0, n_fail = 0;

  auto run = [&](const std::string& name, ProbeCase pc, int M, std::function<float(int, int)> expect,
                 const Maps& mp, bool lane_search = false) {
    CK(cudaMemset(d_out, 0xff, 128 * 256 * 4));
    CK(cudaMemset(d_status, 0, 16));
    CK(cudaMemcpy(d_pc, &pc, sizeof(pc), cudaMemcpyHostToDevice));
    probe_kernel<<<1, 128, SMEM_DATA + 1024>>>(mp, d_pc, d_out, d_dump, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("{\"case\": \"%s\", \"result\": \"CUDA_ERROR\", \"err\": \"%s\"}\n", name.c_str(), cudaGetErrorString(e));
      fflush(stdout);
      exit(4);  // context is dead
    }
    int st = 0;
    CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out.data(), d_out, 128 * 256 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dump.data(), d_dump, SMEM_DATA, cudaMemcpyDeviceToHost));
    if (st != 0) {
      printf("{\"case\": \"%s\", \"result\": \"TIMEOUT\", \"status\": %d}\n", name.c_str(), st);
      ++n_fail;
      fflush(stdout);
      return;
    }
    int N = pc.n_cols;
    double maxdiff = 0;
    int bad = 0, first_m = -1, first_n = -1;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float ex = expect(m, n), got = out[m * N + n];
        double df = fabs((double)ex - (double)got);
        if (!(df <= 1e-3)) {
          if (!bad) { first_m = m; first_n = n; }
          ++bad;
        }
        if (df > maxdiff || df != df) maxdiff = df;
      }
    printf("{\"case\": \"%s\", \"result\": \"%s\", \"bad\": %d, \"of\": %d, \"maxdiff\": %.3f, \"first_bad\": [%d, %d]",
           name.c_str(), bad ? "FAIL" : "PASS", bad, M * N, maxdiff, first_m, first_n);
    if (bad && first_m >= 0)
      printf(", \"exp\": %.1f, \"got\": %.1f", expect(first_m, first_n), out[first_m * N + first_n]);
    if (lane_search) {
      // for every expected row find the TMEM lane that holds it
      printf(", \"row_to_lane\": [");
      for (int m = 0; m < M; ++m) {
        int found = -1;
        for (int l = 0; l < 128 && found < 0; ++l) {
          bool eq = true;
          for (int n = 0; n < N && eq; ++n) eq = fabs(expect(m, n) - out[l * N + n]) < 1e-3;
          if (eq) found = l;
        }
        printf("%s%d", m ? "," : "", found);
      }
      printf("]");
    }
    printf("}\n");
    fflush(stdout);
    bad ? ++n_fail : ++n_pass;
  };

  const uint32_t B_OFF = 96 * 1024;
  auto img = [&](int n, int h, int w, int c) -> float {
    if (h < 0 || h >= IH || w < 0 || w >= IW) return 0.f;
    return X.at(((long)n * IH + h) * IW + w, c);
  };
  auto base_case = [&]() {
    ProbeCase pc;
    memset(&pc, 0, sizeof(pc));
    pc.a_lbo = 16; pc.b_lbo = 16; pc.a_sbo = 1024; pc.b_sbo = 1024;
    pc.a_layout = sm100::LAYOUT_SW128; pc.b_layout = sm100::LAYOUT_SW128;
    pc.n_cols = 64;
    pc.idesc = sm100::make_idesc_bf16(128, 64, 0, 0);
    pc.dump_bytes = 0;
    return pc;
  };
  auto add_B64 = [&](ProbeCase& pc, int nrows_map /*1 or 5*/, int row0, uint32_t bytes) {
    auto& L = pc.loads[pc.n_loads++];
    L.map = nrows_map; L.dims = 2; L.c0 = 0; L.c1 = row0; L.smem_off = B_OFF;
    pc.tx_bytes += bytes;
  };

  // ---- A1: K-major SW128 baseline, N = 64 / 256 / 32 / 16
  for (int N : {64, 256, 32, 16}) {
    ProbeCase pc = base_case();
    pc.loads[pc.n_loads++] = {0, 2, 0, 0, 0, 0, 0};
    pc.tx_bytes = 256 * 128;
    // B rows: use the 256-row box map for all N (extra rows are simply unused)
    add_B64(pc, 5, 0, 256 * 128);
    pc.n_cols = N;
    pc.idesc = sm100::make_idesc_bf16(128, N, 0, 0);
    for (int k = 0; k < 4; ++k) pc.mma[pc.n_mma++] = {(uint32_t)k * 32, B_OFF + (uint32_t)k * 32};
    run("k_sw128_basic_N" + std::to_string(N), pc, 128,
        [&](int m, int n) { float s = 0; for (int k = 0; k < 64; ++k) s += X.at(m, k) * Bm.at(n, k); return s; }, maps);
  }
  // ---- A2: row-shifted start address, base-offset hypotheses
  for (int bo = 0; bo < 2; ++bo)
    for (int sh : {1, 2, 4, 7, 8, 9}) {
      ProbeCase pc = base_case();
      pc.loads[pc.n_loads++] = {0, 2, 0, 0, 0, 0, 0};
      pc.tx_bytes = 256 * 128;
      add_B64(pc, 1, 0, 64 * 128);
      pc.a_bo_mode = bo;
      for (int k = 0; k < 4; ++k) pc.mma[pc.n_mma++] = {(uint32_t)(sh * 128 + k * 32), B_OFF + (uint32_t)k * 32};
      run("k_rowshift_sh" + std::to_string(sh) + "_bo" + std::to_string(bo), pc, 128,
          [&](int m, int n) { float s = 0; for (int k = 0; k < 64; ++k) s += X.at(m + sh, k) * Bm.at(n, k); return s; }, maps);
    }
  // ---- A3/A4: halo tiles (4D TMA with negative coordinates), per-tap descriptor offsets
  for (int WW : {10, 16})
    for (int bo = 0; bo < 2; ++bo)
      for (int t = 0; t < 5; ++t) {
        const int taps[5][2] = {{0, 0}, {0, 1}, {1, 1}, {2, 2}, {1, 0}};
        int r = taps[t][0], s = taps[t][1];
        ProbeCase pc = base_case();
        pc.loads[pc.n_loads++] = {WW == 10 ? 2 : 3, 4, 0, -1, -1, 1, 0};
        pc.tx_bytes = 18 * WW * 128;
        add_B64(pc, 1, 0, 64 * 128);
        pc.a_sbo = WW * 128;
        pc.a_bo_mode = bo;
        for (int k = 0; k < 4; ++k)
          pc.mma[pc.n_mma++] = {(uint32_t)((r * WW + s) * 128 + k * 32), B_OFF + (uint32_t)k * 32};
        run("halo" + std::to_string(WW) + "_tap" + std::to_string(r) + std::to_string(s) + "_bo" + std::to_string(bo), pc, 128,
            [&](int m, int n) {
              int h = m / 8, w = m % 8;
              float acc = 0;
              for (int k = 0; k < 64; ++k) acc += img(1, h + r - 1, w + s - 1, k) * Bm.at(n, k);
              return acc;
            }, maps);
      }
  // ---- A5/A6: MN-major operands (wgrad form): D[m][n] = sum_p A[p][m] * B[p][n]
  for (int bo = 0; bo < 2; ++bo)
    for (int sh : {0, 1, 3, 8, 9, 11}) {
      if (sh == 0 && bo == 1) continue;
      ProbeCase pc = base_case();
      pc.loads[pc.n_loads++] = {4, 2, 0, 0, 0, 0, 0};
      pc.loads[pc.n_loads++] = {4, 2, 64, 0, 0, 0, 24576};
      pc.tx_bytes = 2 * 192 * 128;
      add_B64(pc, 5, 0, 256 * 128);
      pc.a_lbo = 24576; pc.a_sbo = 1024; pc.b_lbo = 32768; pc.b_sbo = 1024;
      pc.a_bo_mode = bo;
      pc.idesc = sm100::make_idesc_bf16(128, 64, 1, 1);
      for (int k = 0; k < 8; ++k) pc.mma[pc.n_mma++] = {(uint32_t)(sh * 128 + k * 2048), B_OFF + (uint32_t)k * 2048};
      run("mn_major_sh" + std::to_string(sh) + "_bo" + std::to_string(bo), pc, 128,
          [&](int m, int n) { float s = 0; for (int p = 0; p < 128; ++p) s += T128.at(p + sh, m) * Bm.at(p, n); return s; }, maps);
    }
  // ---- A7: mixed majors: A MN-major (x tile), B K-major and vice versa
  {
    ProbeCase pc = base_case();
    pc.loads[pc.n_loads++] = {4, 2, 0, 0, 0, 0, 0};
    pc.loads[pc.n_loads++] = {4, 2, 64, 0, 0, 0, 24576};
    pc.tx_bytes = 2 * 192 * 128;
    add_B64(pc, 1, 0, 64 * 128);
    pc.a_lbo = 24576; pc.a_sbo = 1024;
    pc.idesc = sm100::make_idesc_bf16(128, 64, 1, 0);
    for (int k = 0; k < 4; ++k) pc.mma[pc.n_mma++] = {(uint32_t)(k * 2048), B_OFF + (uint32_t)k * 32};
    run("a_mn_b_k", pc, 128,
        [&](int m, int n) { float s = 0; for (int p = 0; p < 64; ++p) s += T128.at(p, m) * Bm.at(n, p); return s; }, maps);
  }
  // ---- A8: SW64 K-major with BLOCK_K = 32
  {
    ProbeCase pc = base_case();
    pc.loads[pc.n_loads++] = {6, 2, 0, 0, 0, 0, 0};
    pc.loads[pc.n_loads++] = {7, 2, 0, 0, 0, 0, B_OFF};
    pc.tx_bytes = 128 * 64 + 64 * 64;
    pc.a_layout = pc.b_layout = sm100::LAYOUT_SW64;
    pc.a_sbo = pc.b_sbo = 512;
    for (int k = 0; k < 2; ++k) pc.mma[pc.n_mma++] = {(uint32_t)k * 32, B_OFF + (uint32_t)k * 32};
    run("k_sw64_basic", pc, 128,
        [&](int m, int n) { float s = 0; for (int k = 0; k < 32; ++k) s += T32.at(m, k) * B32.at(n, k); return s; }, maps);
  }
  // ---- A9: M = 64 (where do the rows land in TMEM?)
  {
    ProbeCase pc = base_case();
    pc.loads[pc.n_loads++] = {0, 2, 0, 0, 0, 0, 0};
    pc.tx_bytes = 256 * 128;
    add_B64(pc, 1, 0, 64 * 128);
    pc.idesc = sm100::make_idesc_bf16(64, 64, 0, 0);
    for (int k = 0; k < 4; ++k) pc.mma[pc.n_mma++] = {(uint32_t)k * 32, B_OFF + (uint32_t)k * 32};
    run("k_sw128_M64", pc, 64,
        [&](int m, int n) { float s = 0; for (int k = 0; k < 64; ++k) s += X.at(m, k) * Bm.at(n, k); return s; }, maps, true);
  }
  // ---- A10: element-stride-2 TMA box.  Which pixels land where, and how many bytes?
  for (uint32_t tx : {8u * 4u * 128u, 16u * 8u * 128u}) {
    Maps m2 = maps;
    m2.m[3] = map_s2;
    ProbeCase pc = base_case();
    pc.loads[pc.n_loads++] = {3, 4, 0, -1, -1, 0, 0};
    pc.tx_bytes = tx;
    add_B64(pc, 1, 0, 64 * 128);
    // A rows = smem rows as landed; expectation: row m=(h*8+w) holds pixel (2h-1, 2w-1)
    for (int k = 0; k < 4; ++k) pc.mma[pc.n_mma++] = {(uint32_t)k * 32, B_OFF + (uint32_t)k * 32};
    run("tma_stride2_tx" + std::to_string(tx), pc, 32,
        [&](int m, int n) {
          int h = m / 8, w = m % 8;
          float acc = 0;
          for (int k = 0; k < 64; ++k) acc += img(0, 2 * h - 1, 2 * w - 1, k) * Bm.at(n, k);
          return acc;
        }, m2);
  }

  // ---- T: tensor-pipe throughput for the descriptor styles the conv kernels use (cycles per MMA, single CTA)
  {
    struct TCase { const char* name; int N; int a_mn, b_mn; uint32_t a_off0, a_kstep, a_sbo, a_lbo, b_kstep, b_sbo, b_lbo; int nk; int M = 128; };
    const TCase tc[] = {
        {"t_M64_N64", 64, 0, 0, 0, 32, 1024, 16, 32, 1024, 16, 4, 64},
        {"t_M64_N128", 128, 0, 0, 0, 32, 1024, 16, 32, 1024, 16, 4, 64},
        {"t_M64_N256", 256, 0, 0, 0, 32, 1024, 16, 32, 1024, 16, 4, 64},
        {"t_M64_N256_Bhalo10", 256, 0, 0, 0, 32, 1024, 16, 32, 1280, 16, 4, 64},
        {"t_M128_N256_Bhalo10", 256, 0, 0, 0, 32, 1024, 16, 32, 1280, 16, 4, 128},
        {"t_M64_N192_mn", 192, 1, 1, 0, 2048, 1024, 16384, 2560, 1280, 128, 8, 64},
        {"t_kmajor_aligned_N64", 64, 0, 0, 0, 32, 1024, 16, 32, 1024, 16, 4},
        {"t_kmajor_shift1_N64", 64, 0, 0, 128, 32, 1024, 16, 32, 1024, 16, 4},
        {"t_kmajor_halo10_tap11_N64", 64, 0, 0, 11 * 128, 32, 1280, 16, 32, 1024, 16, 4},
        {"t_kmajor_aligned_N128", 128, 0, 0, 0, 32, 1024, 16, 32, 1024, 16, 4},
        {"t_kmajor_halo10_tap11_N128", 128, 0, 0, 11 * 128, 32, 1280, 16, 32, 1024, 16, 4},
        {"t_kmajor_aligned_N256", 256, 0, 0, 0, 32, 1024, 16, 32, 1024, 16, 4},
        {"t_mnmajor_aligned_N64", 64, 1, 1, 0, 2048, 1024, 24576, 2048, 1024, 32768, 8},
        {"t_mnmajor_shift11_N64", 64, 1, 1, 11 * 128, 2560, 1280, 128, 2048, 1024, 32768, 8},
    };
    for (const TCase& T : tc) {
      ProbeCase pc = base_case();
      pc.loads[pc.n_loads++] = {0, 2, 0, 0, 0, 0, 0};
      pc.tx_bytes = 256 * 128;
      add_B64(pc, 5, 0, 256 * 128);
      pc.n_cols = T.N;
      pc.idesc = sm100::make_idesc_bf16(T.M, T.N, T.a_mn, T.b_mn);
      pc.a_sbo = T.a_sbo; pc.a_lbo = T.a_lbo; pc.b_sbo = T.b_sbo; pc.b_lbo = T.b_lbo;
      for (int k = 0; k < T.nk; ++k) pc.mma[pc.n_mma++] = {T.a_off0 + (uint32_t)k * T.a_kstep, B_OFF + (uint32_t)k * T.b_kstep};
      pc.repeat = 256;
      CK(cudaMemset(d_status, 0, 16));
      CK(cudaMemcpy(d_pc, &pc, sizeof(pc), cudaMemcpyHostToDevice));
      probe_kernel<<<1, 128, SMEM_DATA + 1024>>>(maps, d_pc, d_out, d_dump, d_status);
      CK(cudaDeviceSynchronize());
      int st[4];
      CK(cudaMemcpy(st, d_status, 16, cudaMemcpyDeviceToHost));
      int n = pc.n_mma * pc.repeat;
      printf("{\"timing\": \"%s\", \"status\": %d, \"mmas\": %d, \"issue_cycles_per_mma\": %.1f, \"total_cycles_per_mma\": %.1f}\n", T.name, st[0], n,
             (double)st[1] / n, (double)st[2] / n);
      fflush(stdout);
    }
  }
  // ---- T2: is the ~98-cycle K-major floor a dependent-accumulator chain or operand fetch?  Independent accumulators
  //          and the narrower swizzles (a K=16 slice is a whole 32-byte row under SW32).
  {
    struct Off { uint32_t a, b; };
    auto timing2 = [&](const char* name, int M, int N, int a_mn, int b_mn, uint32_t layout, uint32_t a_sbo, uint32_t a_lbo,
                       uint32_t b_sbo, uint32_t b_lbo, std::vector<Off> offs, int n_acc, int grid = 1, int b_layout = -1) {
      ProbeCase pc = base_case();
      pc.loads[pc.n_loads++] = {0, 2, 0, 0, 0, 0, 0};
      pc.tx_bytes = 256 * 128;
      add_B64(pc, 5, 0, 256 * 128);
      pc.n_cols = N * (n_acc > 1 ? n_acc : 1);
      pc.idesc = sm100::make_idesc_bf16(M, N, a_mn, b_mn);
      pc.a_layout = layout;
      pc.b_layout = b_layout < 0 ? layout : (uint32_t)b_layout;
      pc.a_sbo = a_sbo; pc.a_lbo = a_lbo; pc.b_sbo = b_sbo; pc.b_lbo = b_lbo;
      for (auto& o : offs) pc.mma[pc.n_mma++] = {o.a, B_OFF + o.b};
      pc.repeat = 256; pc.n_acc = n_acc; pc.acc_stride = (uint32_t)N;
      CK(cudaMemset(d_status, 0, 16));
      CK(cudaMemcpy(d_pc, &pc, sizeof(pc), cudaMemcpyHostToDevice));
      pc.repeat = grid > 1 ? 2048 : 256;     // long enough for chip-level power management to act
      CK(cudaMemcpy(d_pc, &pc, sizeof(pc), cudaMemcpyHostToDevice));
      probe_kernel<<<grid, 128, SMEM_DATA + 1024>>>(maps, d_pc, d_out, d_dump, d_status);
      CK(cudaDeviceSynchronize());
      int st[4];
      CK(cudaMemcpy(st, d_status, 16, cudaMemcpyDeviceToHost));
      int n = pc.n_mma * pc.repeat;
      printf("{\"timing\": \"%s\", \"ctas\": %d, \"status\": %d, \"mmas\": %d, \"issue_cycles_per_mma\": %.1f, \"total_cycles_per_mma\": %.1f}\n", name, grid, st[0], n,
             (double)st[1] / n, (double)st[2] / n);
      fflush(stdout);
    };
    using sm100::LAYOUT_SW128; using sm100::LAYOUT_SW64; using sm100::LAYOUT_SW32;
    std::vector<Off> k128 = {{0, 0}, {32, 32}, {64, 64}, {96, 96}};
    std::vector<Off> k128x8 = {{0, 0}, {32, 32}, {64, 64}, {96, 96}, {0, 0}, {32, 32}, {64, 64}, {96, 96}};
    timing2("t2_k128_N64_acc1", 128, 64, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 1);
    timing2("t2_k128_N64_acc2", 128, 64, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 2);
    timing2("t2_k128_N64_acc4", 128, 64, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 4);
    timing2("t2_k128_N128_acc2", 128, 128, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 2);
    timing2("t2_k128_N32_acc4", 128, 32, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 4);
    timing2("t2_k128_M64_N256_acc2", 64, 256, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 2);
    std::vector<Off> k32 = {{0, 0}, {4096, 8192}, {8192, 16384}, {12288, 24576}, {0, 0}, {4096, 8192}, {8192, 16384}, {12288, 24576}};
    timing2("t2_ksw32_N64_acc1", 128, 64, 0, 0, LAYOUT_SW32, 256, 16, 256, 16, k32, 1);
    timing2("t2_ksw32_N64_acc2", 128, 64, 0, 0, LAYOUT_SW32, 256, 16, 256, 16, k32, 2);
    std::vector<Off> k64 = {{0, 0}, {32, 32}, {8192, 16384}, {8224, 16416}, {0, 0}, {32, 32}, {8192, 16384}, {8224, 16416}};
    timing2("t2_ksw64_N64_acc1", 128, 64, 0, 0, LAYOUT_SW64, 512, 16, 512, 16, k64, 1);
    timing2("t2_ksw64_N64_acc2", 128, 64, 0, 0, LAYOUT_SW64, 512, 16, 512, 16, k64, 2);
    std::vector<Off> mn8;
    for (uint32_t k = 0; k < 8; ++k) mn8.push_back({k * 2048, k * 2048});
    timing2("t2_mn_N64_acc1", 128, 64, 1, 1, LAYOUT_SW128, 1024, 24576, 1024, 32768, mn8, 1);
    timing2("t2_mn_N64_acc2", 128, 64, 1, 1, LAYOUT_SW128, 1024, 24576, 1024, 32768, mn8, 2);
    timing2("t2_mn_N64_acc4", 128, 64, 1, 1, LAYOUT_SW128, 1024, 24576, 1024, 32768, mn8, 4);
    timing2("t2_mn_N128_acc2", 128, 128, 1, 1, LAYOUT_SW128, 1024, 24576, 1024, 32768, mn8, 2);
    timing2("t2_k128_N128_acc1", 128, 128, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 1);
    timing2("t2_k128_N256_acc1", 128, 256, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 1);
    timing2("t2_k128_N256_acc2", 128, 256, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 2);
    timing2("t2_k128_N16_acc4", 128, 16, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 4);
    timing2("t2_k128_M64_N64_acc4", 64, 64, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 4);
    timing2("t2_k128_M64_N256_acc1", 64, 256, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 1);
    // ---- thin-layer weight gradients (conv_umma_wgrad.cu, 16-channel x against 32-channel dy: 3->32 at 384^2 measured ~75
    // cycles per 128x32x16 MMA in the kernel, tools/wgrad_timeline.py).  Timing only (operand values are irrelevant):
    //  * the kernel's operands: A = MN-major 32-byte-swizzle halo windows (pixel rows 10 x 32 B = 320 B apart, eight taps
    //    stacked along M through LBO = one pixel), B = MN-major 64-byte-swizzle dense dy tile;
    //  * the same with DENSE A rows (256 B apart = whole 32-byte-swizzle atoms): is the 320-byte row pitch the cost?
    //  * the TRANSPOSED formulation: A = dy^T padded to M = 64, B = taps x Cin = 144 columns of x.
    std::vector<Off> thin8;
    for (uint32_t k = 0; k < 8; ++k) thin8.push_back({k * 640, k * 1024});            // 16 pixels per MMA: 2 halo rows / 2 dy rows
    timing2("t2_thin_wgrad_M128_N32_halo320", 128, 32, 1, 1, LAYOUT_SW32, 320, 32, 512, 4096, thin8, 2, 1, (int)LAYOUT_SW64);
    std::vector<Off> thin8d;
    for (uint32_t k = 0; k < 8; ++k) thin8d.push_back({k * 512, k * 1024});
    timing2("t2_thin_wgrad_M128_N32_dense256", 128, 32, 1, 1, LAYOUT_SW32, 256, 32, 512, 4096, thin8d, 2, 1, (int)LAYOUT_SW64);
    std::vector<Off> thinT;
    for (uint32_t k = 0; k < 8; ++k) thinT.push_back({k * 1024, k * 640});
    timing2("t2_thin_wgradT_M64_N144", 64, 144, 1, 1, LAYOUT_SW64, 512, 4096, 320, 32, thinT, 2, 1, (int)LAYOUT_SW32);
    timing2("t2_thin_wgradT_M64_N160", 64, 160, 1, 1, LAYOUT_SW64, 512, 4096, 320, 32, thinT, 2, 1, (int)LAYOUT_SW32);
    // the same loops on every SM at once: chip-level (power) limits on the tensor pipe
    timing2("t3_allsm_k128_N64", 128, 64, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 1, 148);
    timing2("t3_allsm_k128_N128", 128, 128, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 1, 148);
    timing2("t3_allsm_k128_N256", 128, 256, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 1, 148);
    timing2("t3_allsm_mn_N64", 128, 64, 1, 1, LAYOUT_SW128, 1024, 24576, 1024, 32768, mn8, 1, 148);
    timing2("t3_allsm_k128_N32", 128, 32, 0, 0, LAYOUT_SW128, 1024, 16, 1024, 16, k128x8, 4, 148);
  }
  printf("{\"summary\": {\"pass\": %d, \"fail\": %d}}\n", n_pass, n_fail);
  return 0;
}
