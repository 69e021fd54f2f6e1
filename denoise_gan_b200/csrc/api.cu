// Context management and error reporting of the dg_b200 C ABI.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "dg_common.cuh"

static thread_local char g_err[1024] = "";

// default ON since round 2 (same-box A/B of the C3 step: 7.21 vs 7.42 ms in job r2_02, 7.40 vs 7.70 ms in job r2_05); it had been
// off in round 1, where the early-launched CTAs cost more than the hidden launch latency (8.43 vs 8.34 ms).  DG_PDL=0 disables.
int g_dg_pdl = []() { const char* e = getenv("DG_PDL"); return (e && e[0] == '0') ? 0 : 1; }();

int g_dg_coop = []() { const char* e = getenv("DG_COOP"); return (e && e[0] == '0') ? 0 : 1; }();   // cooperative-launch attribute of the grid-barrier kernels

void dg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* dg_last_error(void) { return g_err; }
extern "C" int dg_version(void) { return 1; }

extern "C" int dg_init(int device, dg_ctx** out) {
  if (!out) DG_FAIL("dg_init: null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    DG_FAIL("dg_init: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
  if (device < 0 || device >= count) DG_FAIL("dg_init: device %d out of range (%d devices)", device, count);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) DG_FAIL("dg_init: cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) DG_FAIL("dg_init: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    DG_FAIL("dg_init: device is sm_%d%d; this library is built for sm_100a (B200) only", prop.major, prop.minor);
  dg_ctx* c = new dg_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->encode_tiled = nullptr;
  cudaDriverEntryPointQueryResult q;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &c->encode_tiled, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !c->encode_tiled) {
    delete c;
    DG_FAIL("dg_init: cuTensorMapEncodeTiled not available from the driver");
  }
  if (cudaMalloc(&c->tickets, 256) != cudaSuccess || cudaMemset(c->tickets, 0, 256) != cudaSuccess) {
    delete c;
    DG_FAIL("dg_init: cannot allocate the ticket scratch");
  }
  *out = c;
  return 0;
}

extern "C" void dg_destroy(dg_ctx* ctx) {
  if (ctx) cudaFree(ctx->tickets);
  delete ctx;
}

extern "C" int dg_has_umma(dg_ctx* ctx) { return ctx && ctx->cc_major == 10 && ctx->encode_tiled ? 1 : 0; }
