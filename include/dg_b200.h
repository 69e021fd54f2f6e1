/*
 * dg_b200 — C ABI of the B200-native convolutional hot path of pmcbride/denoise-gan.
 *
 * The reference has no FFI: its lower seam is "Keras layer call -> TensorFlow op" (SURVEY.md §8b).
 * Every entry point below replaces one of those op call sites; the citation after each prototype is
 * the reference file:line whose arithmetic it reproduces.  Conventions:
 *   - plain C, no exceptions; return 0 on success, non-zero on error with text in dg_last_error();
 *     invalid shapes/alignments are errors, never a fallback path;
 *   - all tensors are caller-allocated DEVICE memory, NHWC; the library allocates nothing per call
 *     (workspaces are passed in, sized by the *_workspace_bytes helpers);
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*), no hidden syncs,
 *     so whole train steps can be captured into CUDA graphs;
 *   - Keras kernel layouts: Conv2D [kh,kw,Cin,Cout] fp32; Conv2DTranspose [kh,kw,Cout,Cin];
 *     DepthwiseConv2D [kh,kw,C,1].
 */
#ifndef DG_B200_H
#define DG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dg_ctx dg_ctx;

enum { DG_F32 = 0, DG_BF16 = 1 };
enum { DG_ACT_NONE = 0, DG_ACT_RELU = 1, DG_ACT_LRELU = 2, DG_ACT_TANH = 3, DG_ACT_SIGMOID = 4, DG_ACT_PRELU = 5 };

/* NHWC tensor view: logical channels `c` stored at channel offset `coff` inside pixels of
 * `cpitch` elements (cpitch == c, coff == 0 for a dense tensor; a U-Net concat slice otherwise). */
typedef struct {
  void* ptr;
  int32_t dtype;
  int32_t n, h, w, c;
  int32_t cpitch, coff;
} dg_tensor;

/* Forward-convolution geometry.  pad_t / pad_l are the zero rows/cols BEFORE the image (TF 'SAME':
 * total//2, host computes; pix2pix ZeroPadding2D+VALID: 1). */
typedef struct {
  int32_t kh, kw, stride, pad_t, pad_l;
  int32_t act;      /* epilogue activation applied after bias */
  float act_alpha;  /* LeakyReLU slope */
} dg_conv_params;

int dg_init(int device, dg_ctx** out);
void dg_destroy(dg_ctx* ctx);
const char* dg_last_error(void);
int dg_version(void);
/* 1 if the library was built with tcgen05 kernels and `device` is sm_100. */
int dg_has_umma(dg_ctx* ctx);

/* ---- data-parallel gradient exchange (the reference is single-GPU, train_srgan.py:15).  A thin wrapper over
 * NCCL (resolved at run time with dlopen: no link-time dependency), so that a host that is not PyTorch can run the one
 * collective of the train step: a sum all-reduce of the flat gradient arenas of train_srgan.py:111-112's two gradient lists.
 * Rank 0 calls dg_comm_unique_id() and ships the dg_comm_unique_id_bytes() blob to the other ranks by any means; every rank
 * then calls dg_comm_init(); dg_comm_allreduce() enqueues an in-place fp32 sum on the caller's stream (CUDA-graph capturable).
 * Buckets are contiguous ranges of the flat arena: there is no pack / unpack step. */
typedef struct dg_comm dg_comm;
int dg_comm_unique_id_bytes(void);
int dg_comm_unique_id(void* id_out);
int dg_comm_init(dg_comm** out, const void* unique_id, int rank, int world, int device);
int dg_comm_allreduce(dg_comm*, float* buf, long long count, void* stream);
int dg_comm_rank(dg_comm*);
int dg_comm_world(dg_comm*);
void dg_comm_destroy(dg_comm*);

/* ---- convolution, CUDA-core implicit GEMM (fp32 accumulate; any channel count; fp32 or bf16 I/O).
 * Used for the fp32 parity tier and for the <16-channel first/last layers. */
/* y = act(conv(x, w) + bias).  keras Conv2D: srgan.py:154,246; autoencoder.py:95; pix2pix.py:115,207 */
int dg_conv2d_fwd(dg_ctx*, const dg_tensor* x, const float* w_hwio, const float* bias, const dg_tensor* y,
                  const dg_conv_params* p, void* stream);
/* dx = input-gradient of the forward conv described by p (autodiff of the call sites above,
 * train_srgan.py:111-112); also Conv2DTranspose forward with w = [kh,kw,Cout,Cin] (pix2pix.py:130,169):
 * then `dy` is the transposed conv's input, `dx` its output, bias/act apply to the output. */
int dg_conv2d_dgrad(dg_ctx*, const dg_tensor* dy, const float* w_hwio, const float* bias, const dg_tensor* dx,
                    const dg_conv_params* p, void* stream);
/* dw[kh,kw,Cin,Cout] (+)= sum x * dy ; dbias[Cout] (+)= sum dy (dbias may be NULL).  workspace:
 * dg_conv2d_wgrad_workspace_bytes().  accumulate != 0 adds into dw/dbias. */
size_t dg_conv2d_wgrad_workspace_bytes(const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p);
int dg_conv2d_wgrad(dg_ctx*, const dg_tensor* x, const dg_tensor* dy, float* dw_hwio, float* dbias,
                    const dg_conv_params* p, int accumulate, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ---- convolution on tcgen05 tensor cores (bf16 in, fp32 TMEM accumulate, bf16/fp32 out).
 * Requires Cin % 16 == 0 and Cout % 16 == 0.  Weights are pre-packed by dg_umma_pack_weights. */
/* packed size in bytes for a conv with the given geometry; mode 0 = forward, 1 = dgrad */
size_t dg_umma_packed_bytes(int kh, int kw, int cin, int cout, int mode);
int dg_umma_pack_weights(dg_ctx*, const float* w_hwio, void* packed, int kh, int kw, int cin, int cout,
                         int mode, void* stream);
/* same for a whole network in one launch: device table of n 48-byte entries
 * {const float* src; bf16* dst; int32 taps, cin, cout, kc, mode, pad; int64 pad} with kc = 64/32/16 = largest of
 * those dividing the contraction channel count (Cin for mode 0, Cout for mode 1) */
int dg_umma_pack_weights_batch(dg_ctx*, const void* table_dev, int n_entries, void* stream);
int dg_umma_conv2d_fwd(dg_ctx*, const dg_tensor* x, const void* w_packed, const float* bias,
                       const dg_tensor* y, const dg_conv_params* p, float* bn_partials, void* stream);
/* bn_partials (may be NULL): when a BatchNormalization follows the conv (srgan.py:154-155,162-163,247-248), the conv
 * epilogue also accumulates the batch statistics of the values it stores and writes one row [2][Cout] (sum, sum of
 * squares) per CTA; dg_umma_conv2d_fwd_bn_blocks() gives the row count (0: layer not eligible, call dg_bn_stats),
 * dg_bn_finalize() turns the rows into scale/shift/mean/invstd and updates the moving statistics. */
int dg_umma_conv2d_fwd_bn_blocks(dg_ctx*, const dg_tensor* x, const dg_tensor* y, const dg_conv_params* p);
/* Conv2D + the statistics half of the BatchNormalization that follows it, in ONE launch: as dg_umma_conv2d_fwd with
 * bn_partials, and the LAST CTA to finish (ticket counter) sums the per-CTA rows in a fixed order in double precision and
 * does what dg_bn_finalize does (scale/shift, saved mean/invstd, moving-statistics update). */
typedef struct {
  const float* gamma; const float* beta;
  float eps, momentum;
  float* moving_mean; float* moving_var;          /* may be NULL */
  float* scale; float* shift; float* save_mean; float* save_invstd;
  long long pixels;                               /* N*H*W of the conv output */
} dg_bn_fused;
int dg_umma_conv2d_fwd_bn(dg_ctx*, const dg_tensor* x, const void* w_packed, const float* bias, const dg_tensor* y,
                          const dg_conv_params* p, float* bn_partials, const dg_bn_fused* bn, void* stream);
/* Conv2D + training-mode BatchNormalization + activation (+ Add) in ONE cooperative launch: srgan.py:154-157 (conv, BN,
 * PReLU), :162-169 (conv, BN, ReLU, conv, BN, Add), :246-250 (conv, BN, LeakyReLU); fsrgan.py:208-210.  y = raw conv output
 * (kept for the backward pass), out = act(BN_batch(y)) (+ residual); bn receives scale/shift/mean/invstd and the
 * moving-statistics update.  dg_umma_conv2d_fwd_bn_act_blocks() = rows of the [rows][2][Cout] fp32 statistics workspace, or
 * 0 when the fused phase does not apply to the layer (the caller then issues the three separate calls). */
int dg_umma_conv2d_fwd_bn_act(dg_ctx*, const dg_tensor* x, const void* w_packed, const float* bias, const dg_tensor* y,
                              const dg_conv_params* p, float* bn_partials, const dg_bn_fused* bn, int act, float act_alpha,
                              const float* prelu_alpha, const dg_tensor* residual, const dg_tensor* out, void* stream);
int dg_umma_conv2d_fwd_bn_act_blocks(dg_ctx*, const dg_tensor* x, const dg_tensor* y, const dg_conv_params* p);
int dg_umma_conv2d_dgrad(dg_ctx*, const dg_tensor* dy, const void* w_packed_dgrad, const float* bias,
                         const dg_tensor* dx, const dg_conv_params* p, void* stream);
/* Input gradient of a stride-1 convolution whose INPUT was the output a = act(BN_batch(yb)) (+ skip) of a training-mode
 * BatchNormalization (srgan.py:162-169, 246-250): besides dx = dL/da the launch (i) adds the gradient arriving over the skip
 * connection (`residual`, may be NULL; srgan.py:169,174 Add) and (ii) accumulates, from the bf16-rounded dx it stores, the two
 * per-channel sums of the BatchNorm backward pass -- sum g' and sum g' yb, g' = dx * act'(scale*yb + shift) -- into one
 * row [2][Cin] per CTA (`bn`, may be NULL).  dg_bn_bwd_dx_from_partials then finishes the BatchNorm backward pass in ONE read
 * of dx and yb.  Replaces tape.gradient's separate Add / BatchNorm-grad reductions (train_srgan.py:111). */
typedef struct {
  const dg_tensor* y;                                   /* yb: the BatchNorm input (raw conv output), shape of dx, bf16 */
  const float* scale; const float* shift; const float* mean;   /* forward coefficients of that BatchNorm */
  int act; float alpha;                                 /* activation behind it: DG_ACT_NONE / RELU / LRELU */
  float* partials;                                      /* [dg_umma_conv2d_dgrad_fused_blocks()][2][Cin] fp32 */
} dg_bn_bwd_stats;
int dg_umma_conv2d_dgrad_fused(dg_ctx*, const dg_tensor* dy, const void* w_packed_dgrad, const dg_tensor* dx, const dg_conv_params* p,
                               const dg_tensor* residual, const dg_bn_bwd_stats* bn, void* stream);
int dg_umma_conv2d_dgrad_fused_blocks(dg_ctx*, const dg_tensor* dy, const dg_tensor* dx, const dg_conv_params* p);
/* Input gradient of a stride-1 convolution whose input was the output y_relu = relu(conv(...)) of another convolution
 * (autoencoder.py:95-104, the conv2d -> conv2d chains): dx is stored already multiplied by (y_relu > 0) -- the ReLU backward pass
 * of the producing layer folded into this launch.  Applies where dg_umma_conv2d_dgrad_fused_blocks() > 0. */
int dg_umma_conv2d_dgrad_relu_mask(dg_ctx*, const dg_tensor* dy, const void* w_packed_dgrad, const dg_tensor* dx, const dg_conv_params* p,
                                   const dg_tensor* y_relu, void* stream);
/* dx half of the BatchNorm(+activation) backward pass when dg_umma_conv2d_dgrad_fused already reduced the per-channel sums
 * (`partials` = its [rows][2][C] workspace): dx = gamma*invstd*(g' - mean(g') - xhat*mean(g' xhat)); dgamma / dbeta (may be NULL)
 * receive sum g' xhat / sum g'.  One read of dy and x instead of two (autodiff of srgan.py:155,163,167,247). */
int dg_bn_bwd_dx_from_partials(dg_ctx*, const dg_tensor* dy, const dg_tensor* x, const float* scale, const float* shift,
                               const float* gamma, const float* save_mean, const float* save_invstd, int act, float act_alpha,
                               const float* partials, int rows, const dg_tensor* dx, float* dgamma, float* dbeta, int accumulate,
                               void* stream);
/* capability queries: 1 when the tensor-core kernels have a tile configuration for the layer (shared-memory fit);
 * callers route the remaining layers to the CUDA-core kernels explicitly */
int dg_umma_conv2d_fwd_supported(dg_ctx*, const dg_tensor* x, const dg_tensor* y, const dg_conv_params* p);
int dg_umma_conv2d_dgrad_supported(dg_ctx*, const dg_tensor* dy, const dg_tensor* dx, const dg_conv_params* p);
/* debug aid (tools/conv_timeline.py): device buffer of 3*16*4 int64 receiving clock64() marks of CTA 0, or NULL */
void dg_debug_conv_timeline(void* dev_buffer);
void dg_debug_conv_flags(int flags);
void dg_debug_wgrad_timeline(void* dev_buffer);   /* same for the weight-gradient kernel */ /* debug experiments only: results are WRONG when non-zero */
/* Gather for the weight gradient of convolutions on very small maps (pix2pix.py:147-166, the 16x16 ... 1x1 bottleneck of the
 * U-Net): out[p][t*Cin + c] = x[n, ho*stride + r - pad_t, wo*stride + s - pad_l, c], p = (n*Ho + ho)*Wo + wo (bf16, zero outside
 * the image).  With x' = out viewed as [1, P/8, 8, taps*Cin] and dy' = dy viewed as [1, P/8, 8, Cout], dg_umma_conv2d_wgrad of a
 * 1x1 convolution on (x', dy') IS the weight gradient [kh,kw,Cin,Cout] of the original layer (tape.gradient, train_pix2pix.py:62). */
int dg_im2col(dg_ctx*, const dg_tensor* x, const dg_conv_params* p, int out_h, int out_w, void* out, void* stream);
size_t dg_umma_conv2d_wgrad_workspace_bytes(const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p);
int dg_umma_conv2d_wgrad(dg_ctx*, const dg_tensor* x, const dg_tensor* dy, float* dw_hwio, float* dbias,
                         const dg_conv_params* p, int accumulate, void* workspace, size_t workspace_bytes,
                         void* stream);
/* The same for n <= 4 layers of IDENTICAL geometry in ONE launch (the 32 identical 64->64 convolutions of the generator trunk,
 * srgan.py:161-172; the real and fake passes of one discriminator layer, train_srgan.py:78-79): the SMs are divided among the
 * problems and the fixed costs of a launch (prologue, one fp32 partial per CTA, the partial reduce) are paid once per n layers.
 * x[i] / dy[i] / dw[i] / dbias[i] / accumulate[i] describe problem i; the dw pointers are either all distinct or all the same
 * (then the n gradients are summed into it, accumulate[0] decides whether on top of its previous content). */
int dg_umma_conv2d_wgrad_batch_supported(dg_ctx*, int n, const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p);
size_t dg_umma_conv2d_wgrad_batch_workspace_bytes(int n, const dg_tensor* x, const dg_tensor* dy, const dg_conv_params* p);
int dg_umma_conv2d_wgrad_batch(dg_ctx*, int n, const dg_tensor* const* x, const dg_tensor* const* dy, float* const* dw,
                               float* const* dbias, const dg_conv_params* p, const int* accumulate, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ---- depthwise 3x3 s1 SAME (fsrgan.py:149-154) */
int dg_dwconv3x3_fwd(dg_ctx*, const dg_tensor* x, const float* w, const float* bias, const dg_tensor* y, void* stream);
/* DepthwiseConv2D + bias + ReLU in one pass: inference form of fsrgan.py:149-155 with the BatchNorm affine folded into the
 * kernel and bias (infer_video.py:146 runs the model with training=False).  act: DG_ACT_NONE or DG_ACT_RELU. */
int dg_dwconv3x3_fwd_act(dg_ctx*, const dg_tensor* x, const float* w_33c, const float* bias, int act, const dg_tensor* y, void* stream);
/* Inverted-residual block of the Fast-SRGAN generator at inference as ONE launch (fsrgan.py:112-176 under training=False,
 * infer_video.py:92-97,146): y = x + project(relu(depthwise3x3(relu(expand(x))))) with the three BatchNorms folded into the
 * kernels / biases by the caller.  x, y: bf16 NHWC, 32 channels.  w_expand: bf16 [192][32] (output-channel major),
 * w_dw: fp32 [3*3][192], w_project: FP16 [32][192]; biases fp32.  The 192-channel intermediates stay in shared / tensor memory
 * as fp16 (saturating), the depthwise taps are accumulated in fp16 (csrc/fsrgan_block.cu explains why).
 * dg_fsrgan_block_infer_supported() is 1 when the tensors qualify (otherwise the caller issues the three layer calls). */
int dg_fsrgan_block_infer_supported(dg_ctx*, const dg_tensor* x, const dg_tensor* y);
/* debug aid: clock64 marks of CTA 0's phases ([16 tiles][40] int64) of the next dg_fsrgan_block_infer launches; NULL turns it off */
void dg_debug_fsrgan_block_timeline(long long* buf);
int dg_fsrgan_block_infer(dg_ctx*, const dg_tensor* x, const void* w_expand, const float* b_expand, const float* w_dw, const float* b_dw,
                          const void* w_project, const float* b_project, const dg_tensor* y, void* stream);
/* Conv2D(1..3 filters, 3x3, stride 1, SAME) from a 32-channel bf16 tensor at INFERENCE -- the generator's last layer,
 * fsrgan.py:216-217 ('generator_tanh', dtype float32), which infer_video.py:146 runs on the 4x up-scaled frame -- in tap-sum
 * form (csrc/conv_tapsum.cu): ONE tcgen05 product per pixel against all nine taps (N = 9*cout), then the nine shifted adds on
 * the CUDA cores; 2 tensor instructions per 128 pixels instead of 18.  w_hwio: the Keras kernel [3][3][32][cout] in fp32
 * (rounded to bf16 by the kernel, as dg_umma_pack_weights does), bias fp32 [cout] or NULL, act/alpha as dg_conv_params.
 * _fwd writes the fp32 [n,h,w,cout] tensor y; _frame (cout = 3) writes the uint8 frame [n][dst_h][dst_w][3] directly with
 * dg_float_to_frame's arithmetic (infer_video.py:150-159, infer.py:62-68): v*scale + offset, optional clip to [0,1], *255,
 * truncation, optional channel flip, centre crop to dst_h x dst_w (<= h x w) -- pixels outside the crop are not computed and
 * the fp32 image is never written.  _supported() is 1 when x qualifies (otherwise the caller uses dg_umma_conv2d_fwd_narrow). */
int dg_conv3x3_tapsum_supported(dg_ctx*, const dg_tensor* x, int cout);
int dg_conv3x3_tapsum_fwd(dg_ctx*, const dg_tensor* x, const float* w_hwio, const float* bias, int act, float alpha, const dg_tensor* y,
                          void* stream);
int dg_conv3x3_tapsum_frame(dg_ctx*, const dg_tensor* x, const float* w_hwio, const float* bias, int act, float alpha, float scale,
                            float offset, int clip01, int flip_channels, uint8_t* dst, int dst_h, int dst_w, void* stream);
int dg_dwconv3x3_dgrad(dg_ctx*, const dg_tensor* dy, const float* w, const dg_tensor* dx, void* stream);
size_t dg_dwconv3x3_wgrad_workspace_bytes(const dg_tensor* x);
int dg_dwconv3x3_wgrad(dg_ctx*, const dg_tensor* x, const dg_tensor* dy, float* dw, float* dbias, int accumulate,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- BatchNormalization (srgan.py:155,248; fsrgan.py:140-172; pix2pix.py:119,135,211) */
size_t dg_bn_workspace_bytes(const dg_tensor* x);
/* batch statistics -> per-channel scale/shift (y = x*scale+shift), saved mean / invstd, moving-stat update */
int dg_bn_stats(dg_ctx*, const dg_tensor* x, const float* gamma, const float* beta, float eps, float momentum,
                float* moving_mean, float* moving_var, float* scale, float* shift, float* save_mean,
                float* save_invstd, void* workspace, size_t workspace_bytes, void* stream);
int dg_bn_finalize(dg_ctx*, const float* partials, int nblocks, long long pixels, int c, const float* gamma, const float* beta,
                   float eps, float momentum, float* moving_mean, float* moving_var, float* scale, float* shift,
                   float* save_mean, float* save_invstd, void* stream);
/* dg_bn_finalize + dg_bn_act_fwd as ONE launch (training mode, no dropout; srgan.py:155-157,163-169): every block of the apply
 * pass sums the per-CTA statistics rows itself.  Returns 2 (nothing launched) when the views do not qualify for the 8-channel
 * vector kernel: issue the two separate calls. */
int dg_bn_act_fwd_from_partials(dg_ctx*, const dg_tensor* x, const float* partials, int nblocks, const float* gamma, const float* beta,
                                float eps, float momentum, float* moving_mean, float* moving_var, float* scale, float* shift,
                                float* save_mean, float* save_invstd, int act, float act_alpha, const float* prelu_alpha,
                                const dg_tensor* residual, const dg_tensor* y, void* stream);
/* inference: scale/shift from the moving statistics */
int dg_bn_infer_affine(dg_ctx*, int c, const float* gamma, const float* beta, const float* moving_mean,
                       const float* moving_var, float eps, float* scale, float* shift, void* stream);
/* y = act(drop(x*scale+shift)) + residual ; drop: keep-mask from (seed, offset), kept units x2 (pix2pix.py:138).
 * step_counter (device int64, may be NULL) is mixed into the seed so that every optimiser step draws a new mask
 * even when the step is replayed from a CUDA graph: seed_eff = seed + (uint32)(*step_counter) * 0x9E3779B9. */
int dg_bn_act_fwd(dg_ctx*, const dg_tensor* x, const float* scale, const float* shift, int act, float act_alpha,
                  const float* prelu_alpha, const dg_tensor* residual, int dropout, uint32_t seed, uint32_t offset,
                  const int64_t* step_counter, const dg_tensor* y, void* stream);
/* dg_bn_stats followed by dg_bn_act_fwd as ONE launch (statistics, grid barrier, apply): training-mode BatchNormalization
 * + activation of srgan.py:155-157 etc.  Returns 0 on success, 1 on error, 2 when the tensors do not qualify for the
 * fused kernel (channels/pitch not multiples of 8, > 2^31 elements): issue the two calls instead. */
int dg_bn_train_fwd(dg_ctx*, const dg_tensor* x, const float* gamma, const float* beta, float eps, float momentum,
                    float* moving_mean, float* moving_var, float* scale, float* shift, float* save_mean, float* save_invstd,
                    int act, float act_alpha, const float* prelu_alpha, const dg_tensor* residual, int dropout, uint32_t seed,
                    uint32_t offset, const int64_t* step_counter, const dg_tensor* y, void* workspace, size_t workspace_bytes,
                    void* stream);
/* backward of the above: dx, dgamma, dbeta (and dprelu_alpha when act == PRELU) */
int dg_bn_act_bwd(dg_ctx*, const dg_tensor* dy, const dg_tensor* x, const float* scale, const float* shift,
                  const float* gamma, const float* save_mean, const float* save_invstd, int act, float act_alpha,
                  const float* prelu_alpha, int dropout, uint32_t seed, uint32_t offset, const int64_t* step_counter,
                  const dg_tensor* dx,
                  float* dgamma, float* dbeta, float* dprelu_alpha, int accumulate, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ---- activations / structural ops */
/* dpre = dy * act'(.) from the saved OUTPUT y (relu, lrelu, tanh, sigmoid) */
int dg_act_bwd_from_output(dg_ctx*, const dg_tensor* dy, const dg_tensor* y, int act, float act_alpha,
                           const dg_tensor* dpre, void* stream);
/* tf.nn.depth_to_space(u, 2) followed by PReLU(shared_axes=[1,2]) (srgan.py:145-146, fsrgan.py:188-189);
 * prelu_alpha == NULL: plain depth_to_space */
int dg_d2s_prelu_fwd(dg_ctx*, const dg_tensor* u, const float* prelu_alpha, const dg_tensor* y, void* stream);
int dg_d2s_prelu_bwd(dg_ctx*, const dg_tensor* dy, const dg_tensor* u, const float* prelu_alpha, const dg_tensor* du,
                     float* dprelu_alpha, int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* out = a + b */
int dg_add(dg_ctx*, const dg_tensor* a, const dg_tensor* b, const dg_tensor* out, void* stream);
/* out (+)= src, with dtype conversion and view (concat-slice) support; accumulate != 0 adds */
int dg_copy(dg_ctx*, const dg_tensor* src, const dg_tensor* out, int accumulate, void* stream);
/* MaxPool2D(2,2) (autoencoder.py:110) and its gradient (first max in window order gets the gradient) */
int dg_maxpool2x2_fwd(dg_ctx*, const dg_tensor* x, const dg_tensor* y, void* stream);
int dg_maxpool2x2_bwd(dg_ctx*, const dg_tensor* dy, const dg_tensor* x, const dg_tensor* y, const dg_tensor* dx, void* stream);
/* the same when x is the output of a ReLU convolution (autoencoder.py:95-110, conv2d(..., relu) -> maxpool2d): the gradient is also
 * multiplied by (x > 0), i.e. the ReLU's own backward pass is folded in and tape.gradient needs no separate pass over x */
int dg_maxpool2x2_bwd_relu(dg_ctx*, const dg_tensor* dy, const dg_tensor* x, const dg_tensor* y, const dg_tensor* dx, void* stream);
/* UpSampling2D(2,'nearest') + relu into a concat slice (autoencoder.py:113-136) and its gradient */
int dg_upsample2x_relu_fwd(dg_ctx*, const dg_tensor* x, const dg_tensor* y, void* stream);
int dg_upsample2x_relu_bwd(dg_ctx*, const dg_tensor* dy, const dg_tensor* x, const dg_tensor* dx, void* stream);

/* dbias[c] (+)= sum_pixels dy[..,c]: bias gradient of Conv2DTranspose (pix2pix.py:169-173) */
int dg_bias_grad(dg_ctx*, const dg_tensor* dy, float* dbias, int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* vgg19.preprocess_input(((x+1)*255)/2), 'caffe' mode: RGB->BGR, subtract (103.939,116.779,123.68) (srgan.py:71-72) */
int dg_vgg_preprocess_fwd(dg_ctx*, const dg_tensor* x, const dg_tensor* y, void* stream);
int dg_vgg_preprocess_bwd(dg_ctx*, const dg_tensor* dy, const dg_tensor* dx, void* stream);

/* ---- zero-padded channels for the tensor-core path: RGB-sided layers (srgan.py:154 conv 3->64, :182 conv 64->3,
 * :236-246 discriminator conv 3->32; same in the other models) run on tcgen05 through a 16-channel bf16 copy.
 * dg_pad_channels: dst[..., :src.c] = src (dtype conversion), dst[..., src.c:] = 0.
 * dg_umma_pack_weights_padded: dg_umma_pack_weights of the kernel zero-padded to [kh,kw,cin_pad,cout_pad].
 * dg_unpad_weight_grad: dw[t,c,o] (+)= dw_padded[t,c,o] (c < cin, o < cout), dbias likewise (may be NULL). */
int dg_pad_channels(dg_ctx*, const dg_tensor* src, const dg_tensor* dst, void* stream);
int dg_umma_pack_weights_padded(dg_ctx*, const float* w, void* packed, int kh, int kw, int cin, int cout, int cin_pad, int cout_pad,
                                int mode, void* stream);
int dg_unpad_weight_grad(dg_ctx*, const float* dw_padded, const float* dbias_padded, float* dw, float* dbias, int kh, int kw, int cin,
                         int cout, int cin_pad, int cout_pad, int accumulate, void* stream);
/* Conv2D + tf.nn.depth_to_space(.., 2) + PReLU(shared_axes=[1,2]) (srgan.py:144-146, fsrgan.py:180-186) as one launch for INFERENCE:
 * y is the [n, 2h, 2w, Cout/4] bf16 result; channel c = blk*(Cout/4) + cc of conv pixel (h, w) is stored at (2h + blk/2, 2w + blk%2, cc)
 * after bias and PReLU with slope prelu_alpha[cc] (may be NULL).  The pre-activation tensor is never written (training keeps
 * dg_umma_conv2d_fwd + dg_d2s_prelu_fwd: the backward pass needs it). */
int dg_umma_conv2d_fwd_d2s_prelu(dg_ctx*, const dg_tensor* x, const void* w_packed, const float* bias, const dg_tensor* y,
                                 const dg_conv_params* p, const float* prelu_alpha, void* stream);
/* Conv2D -> BatchNormalization(training=False) -> [PReLU(shared_axes=[1,2])] -> [Add skip] as ONE launch for INFERENCE
 * (fsrgan.py:172-176 project + add, :208-210 post-residual conv + add; srgan.py:166-169, :174-176): the BatchNorm is folded into
 * w_packed / bias by the caller, the staged epilogue computes y = act(conv + bias) + residual, where act is PReLU with per-channel
 * slopes prelu_alpha (NULL: p->act) and residual is a bf16 tensor of y's shape (NULL: none).  Applies to layers that take the staged
 * epilogue -- dg_umma_conv2d_fwd_bn_blocks() > 0: dense bf16 output, 16/32/64-channel blocks -- and fails otherwise. */
int dg_umma_conv2d_fwd_res_prelu(dg_ctx*, const dg_tensor* x, const void* w_packed, const float* bias, const dg_tensor* y,
                                 const dg_conv_params* p, const dg_tensor* residual, const float* prelu_alpha, void* stream);
/* Conv2D with fewer than 16 output channels (the RGB side, srgan.py:182 / fsrgan.py:217 / autoencoder.py:186), fp32 output:
 * w_packed / bias_padded are zero-padded to 16 output channels, y is the DENSE [n,h,w,c<16] result (no padded copy, no slice). */
int dg_umma_conv2d_fwd_narrow(dg_ctx*, const dg_tensor* x, const void* w_packed, const float* bias_padded, const dg_tensor* y,
                              const dg_conv_params* p, void* stream);
/* Physically padded activations (the autoencoder's 44/56/76/100/152/84-channel layers, autoencoder.py:150-186, stay zero-padded
 * to multiples of 16 from layer to layer instead of being padded and sliced around every convolution).  The input-channel axis
 * of a kernel that consumes a U-Net concat of two padded tensors (autoencoder.py:135) consists of TWO padded segments:
 * physical channels [0, seg_phys) hold the first seg_log channels of the Keras kernel, [seg_phys, cin_pad) the remaining
 * cin - seg_log; everything else is zero.  seg_phys = 0: one segment (= the functions above). */
int dg_umma_pack_weights_seg(dg_ctx*, const float* w, void* packed, int kh, int kw, int cin, int cout, int cin_pad, int cout_pad,
                             int seg_log, int seg_phys, int mode, void* stream);
int dg_unpad_weight_grad_seg(dg_ctx*, const float* dw_padded, const float* dbias_padded, float* dw, float* dbias, int kh, int kw,
                             int cin, int cout, int cin_pad, int cout_pad, int seg_log, int seg_phys, int accumulate, void* stream);

/* ---- image summaries of the training loop (train_srgan.py:27-59, 153-172): the FIRST image of batch `a` (minus the first image
 * of `b` when b is not NULL) as uint8 [h', w', c].  kind 0: uint8(255 * clip((x+1)/2, 0, 1)) (tf2image(norm=True)); the others are
 * auto-scaled to their own range, uint8(255 * (v - min v) / ptp v) (tf2image(norm=False)): 1 x^2, 2 |x|, 3 Sobel magnitude of
 * renorm(x) (tf.image.sobel_edges, REFLECT padding, /4), 4 / 5 horizontal / vertical differences and 6 total variation
 * (high_pass_x_y, total_variation: h' = h-1, w' = w-1). */
size_t dg_image_summary_workspace_bytes(int h, int w, int c);
int dg_image_summary(dg_ctx*, const dg_tensor* a, const dg_tensor* b, int kind, uint8_t* out, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---- training-pair synthesis on the device (dataloader.py:188-229 after load_image): stack_crop (:79-93) -> scale_image =
 * tf.image.resize(method='bicubic') (:110-124) -> adjust_jpeg_quality = tf.image.adjust_jpeg_quality (:126-140) -> normalize
 * x*2-1 (:160-178).  src: n_src decoded uint8 images [n_src, src_h, src_w, 3] in DEVICE memory; sample b is the crop x crop
 * window at (crop_top[b], crop_left[b]) of image crop_index[b] (three DEVICE int arrays of `batch` entries: the random offsets
 * of tf.image.random_crop are drawn by the caller).  target [batch, crop, crop, 3] and input [batch, crop/scale, crop/scale, 3]
 * are dense fp32 in [-1, 1].  The JPEG step is libjpeg's baseline round trip (4:2:0, 'islow' DCT, fancy up-sampling) in its own
 * integer arithmetic, bit-exact against libjpeg-turbo; crop/scale must be a multiple of 16. */
size_t dg_pair_synthesis_workspace_bytes(int batch, int crop, int scale);
int dg_pair_synthesis(dg_ctx*, const uint8_t* src, int n_src, int src_h, int src_w, const int* crop_index, const int* crop_top,
                      const int* crop_left, int batch, int crop, int scale, int jpeg_quality, float* input, float* target,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- frame pre/post-processing around the inference forward (infer_video.py:138-159, infer.py:50-68,
 * unit_test.py:67-86).  src/dst frames are packed uint8 [n, h, w, 3] in DEVICE memory.
 * dg_frame_to_float: centre crop-or-pad (tf.image.resize_with_crop_or_pad, infer_video.py:142) of the frame to
 *   dst's h x w, u8 -> [0,1] (norm_mode 0: float32 * (1/255) as tf.image.convert_image_dtype, infer_video.py:141;
 *   1: float64 '/ 255.0' then cast, infer.py:58; 2: float32 '/ 255', unit_test.py:74), padding pixels = 0,
 *   then v*scale + offset (infer_video.py:145 uses 2,-1), optional BGR<->RGB flip (cv2.cvtColor, :140).
 * dg_float_to_frame: v = src*scale + offset (infer_video.py:149 uses 0.5,0.5), centre crop-or-pad to dst_h x dst_w
 *   with zeros (:156), optional clip to [0,1] (:156), *255 and truncation to uint8 (:158-159), optional flip. */
int dg_frame_to_float(dg_ctx*, const uint8_t* src, int src_h, int src_w, int flip_channels, int norm_mode, float scale, float offset,
                      const dg_tensor* dst, void* stream);
int dg_float_to_frame(dg_ctx*, const dg_tensor* src, float scale, float offset, int clip01, int flip_channels, uint8_t* dst,
                      int dst_h, int dst_w, void* stream);

/* ---- losses: value sums and upstream gradients in one pass */
size_t dg_loss_workspace_bytes(const dg_tensor* t);
/* out3 = {mean|t-g|, mean(t-g)^2, mean_b TV(t-g)};  dgen (+)= w_mae*dMAE + w_mse*dMSE + w_tv*dTV
 * (train_srgan.py:88-90, pix2pix.py:78-84); dgen may be NULL */
int dg_image_losses(dg_ctx*, const dg_tensor* gen, const dg_tensor* target, float w_mae, float w_mse, float w_tv,
                    float* out3, const dg_tensor* dgen, int accumulate, void* workspace, size_t workspace_bytes,
                    void* stream);
/* The scalar arithmetic of the train step between the loss reductions and the returned tuple (train_srgan.py:86-99, 118), as one
 * launch instead of a dozen scalar tensor ops: content (device scalar, may be NULL = 0), adv_raw = BCE(ones, D(fake)), out3 =
 * {mae, mse, mean total variation} of dg_image_losses, real / fake = the discriminator's two BCE terms.
 * out7 = {gen_loss, adv_loss = 1e-3 adv_raw, mae_loss, mse_loss, content_loss, disc_loss = disc_scale (real + fake), var_loss = 1e-5 tv}
 * with gen_loss = content + adv_loss + w_mae mae + w_mse mse + tv_gain var_loss. */
int dg_gan_loss_terms(dg_ctx*, const float* content, const float* adv_raw, const float* out3, const float* real_loss,
                      const float* fake_loss, float w_mae, float w_mse, float tv_gain, float disc_scale, float* out7, void* stream);
/* BinaryCrossentropy against a constant target. from_logits=1: train_srgan.py:71; 0: clipped-probability
 * form, train_autoencoder.py:79.  loss_out = mean BCE ; dx = grad_scale * dBCE/dx (dx may be NULL) */
int dg_bce_const_target(dg_ctx*, const dg_tensor* x, float target, int from_logits, float grad_scale,
                        float* loss_out, const dg_tensor* dx, void* workspace, size_t workspace_bytes, void* stream);
/* feature MSE for the VGG content loss (srgan.py:73-75): mean((a-b)^2 / 12.75^2), da = grad wrt a */
int dg_feature_mse(dg_ctx*, const dg_tensor* a, const dg_tensor* b, float inv_div, float* loss_out,
                   const dg_tensor* da, void* workspace, size_t workspace_bytes, void* stream);

/* ---- fused Keras Adam over a flat parameter arena (srgan.py:35-50, pix2pix.py:30-31) */
/* state: int64 iterations at state[0]; advances it and updates theta/m/v for `numel` elements.
 * grad_scale multiplies the gradient first (1/world_size after an all-reduce sum). */
int dg_adam_step(dg_ctx*, float* theta, const float* grad, float* m, float* v, int64_t numel, float lr0,
                 float beta1, float beta2, float eps, int64_t decay_steps, float decay_rate, float grad_scale,
                 int64_t* iterations_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DG_B200_H */
