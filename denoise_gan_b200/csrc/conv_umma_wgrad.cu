// Tensor-core weight gradient (placeholder until the tcgen05 kernel lands): reports an error so
// callers never silently take another path.
#include "dg_common.cuh"

extern "C" size_t dg_umma_conv2d_wgrad_workspace_bytes(const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p) {
  return 0;
}

extern "C" int dg_umma_conv2d_wgrad(dg_ctx* ctx, const dg_tensor* x, const dg_tensor* dy, float* dw, float* dbias,
                                    const dg_conv_params* p, int accumulate, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  DG_FAIL("dg_umma_conv2d_wgrad: not built in this revision");
}
