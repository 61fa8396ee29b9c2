/*
 * fnst.h -- C ABI of libfnst.so: hand-written sm_100a kernels for the style-transfer hot path.
 *
 * The reference (HajarHAMDOUCH01/Fast-neural-style-transfer) is pure Python and has no FFI; its
 * boundary for this path is a set of Python symbols.  Each entry point
 * below is the native operator that the drop-in Python symbol binds through ctypes; the
 * reference construct it replaces is cited as file:line into the reference tree.
 *
 * Conventions
 *   - extern "C", POD arguments only: raw device pointers, ints, one POD descriptor struct.
 *   - All device memory is caller-owned (torch-allocated); the library never allocates, frees or
 *     retains device pointers across calls.
 *   - Every launch is asynchronous on `stream` of CUDA device `device`; no hidden sync.  Kernels of the tensor-core
 *     path are launched with programmatic stream serialization (PDL): each one overlaps its set-up with the tail of
 *     the kernel in front of it on the stream and executes griddepcontrol.wait before touching global memory, so
 *     stream order is preserved; under stream capture these become programmatic CUDA-graph edges.  FNST_PDL=0 in
 *     the environment turns the launch attribute off.
 *   - Return 0 on success, <0 for argument/shape/alignment errors, >0 = cudaError_t.
 *     fnst_last_error() returns a thread-local message.  There is no fallback path: an
 *     unsupported configuration is an error.
 *   - Activations are NHWC ("pixel-major") tensors of element type FNST_F32 / FNST_F16 /
 *     FNST_BF16.  Reflection padding is materialised by the producer as a halo in the
 *     consumer's buffer, so every convolution is a stride-1 "valid" gather-GEMM:
 *         out[n,h,w,j] = sum_t sum_{c<kc} A[n, h+h0+dh[t], w+w0+dw[t], c0[t]+c] * B[j][t*kc+c]
 *     with out-of-range A coordinates reading as zero (TMA out-of-bounds fill).  Stride-2
 *     convolutions read a space-to-depth halo buffer; transposed convolutions are 2x2-tap
 *     gather-GEMMs with 4*C_out columns and a depth-to-space epilogue.
 */
#ifndef FNST_H_
#define FNST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FNST_MAX_TAPS 96

enum { FNST_F32 = 0, FNST_F16 = 1, FNST_BF16 = 2 };

/* conv epilogues */
enum {
  FNST_EPI_NHWC = 0,   /* out[n,h,w,j], j < c_out; optional bias, relu, per-(n,c) sum/sumsq      */
  FNST_EPI_D2S = 1,    /* j = (ph*2+pw)*c_out + o  ->  out[n,2h+ph,2w+pw,o]; stats indexed by o   */
  FNST_EPI_NCHW_F32 = 2, /* out[n,j,h,w] fp32 (+bias), j < c_out                                 */
  FNST_EPI_ROWSUM9 = 3   /* 9-row separable form of a 9x9 conv (tensor cores only): GEMM column kh*c_out+o holds the
                            horizontal partial T[(y,x)][kh,o] of input row y; out[n,o,h,w] = bias[o] +
                            sum_kh T[(h+kh, w)][kh,o], fp32 NCHW.  n_gemm = 32, 9*c_out <= 32, taps have dh = 0.   */
};

enum { FNST_PAD_NONE = 0, FNST_PAD_REFLECT = 1, FNST_PAD_ZERO = 2 };

/* Descriptor of one gather-GEMM convolution (see formula above).  Strides are in elements. */
typedef struct fnst_conv_desc {
  const void* a;                 /* activation view, element type `dtype`                         */
  int64_t a_stride_w, a_stride_h, a_stride_n;
  int32_t a_w, a_h, a_n, a_c;    /* logical extents of the view (bounds for zero fill)            */
  int32_t ntaps, kc;             /* taps and channels per tap (kc % 64 == 0 for tensor cores)     */
  int32_t h0, w0;
  int8_t tap_dh[FNST_MAX_TAPS];
  int8_t tap_dw[FNST_MAX_TAPS];
  int16_t tap_c0[FNST_MAX_TAPS];
  const void* b;                 /* packed weights [n_gemm][ntaps*kc], element type `dtype`       */
  int32_t n_gemm;                /* GEMM columns (multiple of 16)                                 */
  int32_t out_n, out_h, out_w;   /* output tile space                                             */
  int32_t epilogue;              /* FNST_EPI_*                                                    */
  int32_t c_out;                 /* real output channels                                          */
  int32_t relu;
  int32_t dtype;                 /* operand element type                                          */
  int32_t out_dtype;             /* element type of `out` (ignored for NCHW_F32)                  */
  void* out;
  const float* bias;             /* [c_out] or NULL                                               */
  float* stats;                  /* [out_n][c_out][2] (sum, sumsq), zeroed by the call; or NULL   */
  /* backward (dgrad) epilogue extras, FNST_EPI_NHWC only: out = (acc + addend) * (mask > 0)       */
  const void* addend;            /* same shape/dtype as out, or NULL                              */
  const void* mask;              /* forward activation (ReLU output) of dtype mask_dtype, or NULL */
  int32_t mask_dtype;
  int32_t b_image_rows;          /* 0: one weight matrix for all images; else image n uses rows [n*b_image_rows, +n_gemm) of b */
  /* wgrad only: element strides of the gradient operand g (all zero = contiguous NHWC [out_n,out_h,out_w,n_gemm]) */
  int64_t g_stride_w, g_stride_h, g_stride_n;
  int32_t flags;                 /* FNST_DESC_* bits                                              */
} fnst_conv_desc;

/* fnst_conv_desc.flags: the accumulators the call would zero (conv: `stats`; wgrad: `out`) were already zeroed by the
 * caller (e.g. one arena memset per forward), so the call issues no memset in front of the kernel and the kernel can
 * be chained to its predecessor by programmatic dependent launch. */
#define FNST_DESC_PREZEROED 1
/* fnst_conv_tc only: PIXEL-STREAM form.  The activation view is ONE row of a_w pixels (a_h == a_n == out_h == out_n == 1, NHWC
 * epilogue, no statistics) and tap (dh, dw) addresses pixel  w + dh * pitch + dw  with pitch = a_stride_h / a_stride_w; GEMM
 * tiles are 128 consecutive pixels.  A zero-haloed image batch [n][H+2z][W+2z][c] presented this way turns a zero-padded
 * convolution over the (H+2z) x (W+2z) domain into n*(H+2z)*(W+2z)/128 full tiles -- the data gradient of a 3x3 convolution
 * behind ReflectionPad2d(1) on 4 x 64 x 64 images is 145 tiles (one wave of a 148-SM GPU) instead of the 180 of 8 x 16 pixel
 * boxes on the 66 x 66 domain.  Columns beyond the image in each row of the output are don't-care values. */
#define FNST_DESC_LINEAR 2

int fnst_version(void);
const char* fnst_last_error(void);
/* Tuning knobs of the tensor-core kernels (measurement tooling; defaults are the measured best):
 * "conv_block_n" (0 = heuristic), "conv_pair" (0/1: cta_group::2 CTA pairs), "conv_stage_out" (0/1: shared-memory tile + TMA
 * store epilogue for one-tile-per-CTA launches), "conv_rowstream" (0/1: row-streaming kernel with resident weights for 3x3
 * convolutions over 64 input channels with 64 / 128 outputs and no statistics), "wgrad_waves_x2", "wgrad_bn" (0 = widest), "pdl" (0/1),
 * "inorm_bwd_tma" (0/1, default 0: InstanceNorm backward pass 1 stages image rows with TMA box loads), "inorm_bwd_blocks" (1/2 resident
 * blocks per SM of the register form of that kernel), "resize_staged" (0/1).
 * Returns 0, or -1 for an unknown name. */
int fnst_set_tuning(const char* name, int value);
/* Measurement only: while set (non-NULL), every fnst_conv_tc launch makes the MMA-issuing thread of CTA i write
 * buf[4*i .. 4*i+3] = {SM clocks, nanoseconds (globaltimer), k-blocks, tiles} of its main loops (first operands landed ->
 * last accumulator complete).  buf must hold 4 * 148 uint64.  NULL switches it off. */
int fnst_set_debug_buffer(void* device_ptr);
/* 1 when the running device can execute the tcgen05/TMA kernels (compute capability 10.x). */
int fnst_device_supports_tc(int device);

/*
 * Gather-GEMM convolution on tensor cores (tcgen05.mma, TMEM accumulators, TMA operand loads).
 * Replaces nn.Conv2d behind ReflectionPad2d (models/model.py:68-75), nn.ConvTranspose2d
 * (models/model.py:13-22) and torchvision VGG conv+ReLU (models/vgg19_net.py:38-51).
 * dtype must be FNST_F16 or FNST_BF16.
 */
int fnst_conv_tc(const fnst_conv_desc* d, int device, void* stream);

/*
 * final_conv forward on tensor cores as a row-streaming kernel: ConvLayer(32, 3, kernel=9) (models/model.py:47,64-65).
 *   act      [n][h+8][w+8][32] fp16/bf16 NHWC halo buffer (reflected border written by fnst_inorm_apply, pad 4)
 *   wpacked  [9 kw][2 channel halves][2 k-chunks][32 rows = kh*3+o, 27 used][8 channels] of the same element type
 *            (18 KB; engine.pack_final_stream)
 *   bias3    the three biases (device, fp32);  out [n][3][h][w] fp32 NCHW.
 * Each input row is loaded into shared memory once (un-swizzled core-matrix layout) and the nine horizontal taps are
 * the same buffer at start addresses 16 bytes apart; the nine-row sum runs over a TMEM ring in the epilogue.
 */
int fnst_finalconv_tc(const void* act, const void* wpacked, const float* bias3, float* out, int n, int h, int w,
                      int dtype, int device, void* stream);

/* Same operator on CUDA cores with fp32 accumulation (any dtype); the fp32-accurate path. */
int fnst_conv_simt(const fnst_conv_desc* d, int device, void* stream);

/*
 * First-layer convolution for 3-channel NCHW fp32 images: reflect/zero padding by index math.
 * Replaces ConvLayer(3,64,9,stride=2) (models/model.py:28) and VGG conv1_1 (features[0]).
 * x [n,3,h,w] fp32; wgt [3*k*k][c_out] fp32, tap-major (index ((c*k+kh)*k+kw)*c_out + o, i.e. the
 * PyTorch OIHW weight permuted to (c,kh,kw,o)); out NHWC [n,ho,wo,c_out] of out_dtype.
 * stats (optional) as above; bias/relu optional.
 */
int fnst_conv_first(const float* x, int n, int h, int w, const float* wgt, const float* bias,
                    int c_out, int k, int stride, int pad, int pad_mode, int relu,
                    void* out, int out_dtype, float* stats, int device, void* stream);

/*
 * InstanceNorm2d(affine) apply + ReLU + Dropout2d scale + residual add, writing the consumer's
 * halo buffer.  Replaces nn.InstanceNorm2d / F.relu / nn.Dropout2d / `x + y`
 * (models/model.py:51,52,60,61,86-90).
 *   raw   [n,h,w,c] dtype          conv output
 *   stats [n,c,2] fp32             per-plane sum / sum of squares from the conv epilogue
 *   drop  [n,c] fp32 or NULL       Dropout2d scale (0 or 1/0.9)
 *   res   or NULL                  residual source: NHWC buffer with halo `res_pad`
 *   out                            [n, h+2*pad, w+2*pad, c]; if s2d: [n, ceil((h+2p)/2), ceil((w+2p)/2), 4c]
 *                                  with channel = ((hp&1)*2 + (wp&1))*c + ch
 *   out_bf16 or NULL               second copy of the same activation as bfloat16, same geometry with c channels per pixel
 *                                  (also when `split`): the operand of the weight-gradient GEMM, whose other operand --
 *                                  the gradient -- is bf16 (saves a cast pass over every saved activation in training)
 *   raw_dtype                      element type of raw (== dtype, or FNST_F32 with split)
 *   split                          error-compensated fp16 pair: out (and res) hold 2c channels per pixel [hi(c) | lo(c)],
 *                                  hi = fp16(y), lo = fp16(y - hi)  (fp16x3 path: hi*hi + hi*lo + lo*hi on tensor cores)
 */
int fnst_inorm_apply(const void* raw, const float* stats, const float* gamma, const float* beta,
                     const float* drop, const void* res, int res_pad, void* out, void* out_bf16,
                     int n, int h, int w, int c, int dtype, int relu, float eps,
                     int pad, int pad_mode, int s2d, int raw_dtype, int split, int device, void* stream);

/*
 * NCHW fp32 3-channel image -> 2-byte NHWC halo buffer [n][rows][pitch][c_pad] (c_pad 4 or 8; channels >= 3 zero), interior at
 * (pad, pad) with reflect / zero border, the rest of the buffer zero.  Feeds the tensor-core form of the first-layer
 * convolutions: with 4 (8) channels per pixel a window of 16 (8) consecutive pixels is one 128-byte K row, so
 * ConvLayer(3,64,9,stride=2) (models/model.py:28) is a 9-tap and VGG conv1_1 a 3-tap gather-GEMM (taps = kernel rows).
 */
int fnst_image_to_halo(const float* x, void* out, int n, int h, int w, int pad, int pad_mode, int c_pad, int rows,
                       int pitch, int dtype, int split, int device, void* stream);
/* split (c_pad == 8 only): channels 0..2 hold fp16(x), channels 4..6 hold fp16(x - hi) (hi|lo pair of the fp16x3 path). */

/* MaxPool2d(2,2) on NHWC (torchvision features[4], [9], [18]); h, w are input extents (floor). */
int fnst_maxpool2(const void* in, void* out, int n, int h, int w, int c, int dtype, int device, void* stream);

/*
 * Gram matrix G[n] = F[n]^T F[n] for NHWC features F[n] = [h*w][c]  (losses/losses.py:6-13).
 * out fp32 [n][c][c].  use_tc: tcgen05 path (dtype F16/BF16, c % 64 == 0), else CUDA cores.
 * prezeroed != 0: `out` is already zero (the split-K partial tiles are accumulated into it); else the call memsets it.
 */
int fnst_gram(const void* feat, float* out, int n, int hw, int c, int dtype, int use_tc, int prezeroed, int device, void* stream);

/* acc[0] += sum (a-b)^2 over `count` elements (double accumulator)  (losses/losses.py:41,54).
 * b_period: b is indexed modulo b_period (target Gram broadcast over the batch). */
int fnst_sse(const void* a, const void* b, int64_t count, int64_t b_period, int dtype_a, int dtype_b,
             double* acc, int device, void* stream);

/* Loss scalars in one launch each, no memset, no host arithmetic (losses/losses.py:41,54,71 incl. their normalisation):
 *   out[0] (+)= scale * sum (a-b)^2        (accumulate != 0: add to out[0]; else overwrite)
 *   out[0]   = scale * TV(img)
 * Block partials go to `workspace` (fnst_loss_workspace_bytes() bytes, zeroed ONCE by the caller) and are summed in block
 * order by the last block to finish (deterministic), which also resets the workspace. */
int64_t fnst_loss_workspace_bytes(void);
int fnst_sse_scaled(const void* a, const void* b, int64_t count, int64_t b_period, int dtype_a, int dtype_b, float scale,
                    void* workspace, float* out, int accumulate, int device, void* stream);
int fnst_tv_scaled(const float* img, int planes, int h, int w, float scale, void* workspace, float* out, int device, void* stream);

/* acc[0] += sum of squared vertical and horizontal differences of an NCHW fp32 image
 * (losses/losses.py:62-73; the caller divides by b*c*h*w). */
int fnst_tv(const float* img, int planes, int h, int w, double* acc, int device, void* stream);

/* NHWC (dtype) -> NCHW fp32 and back (layout plumbing at the module boundary).  c_pad >= c: the
 * NHWC tensor has c_pad channels, channels c..c_pad-1 are written as zero. */
int fnst_nhwc_to_nchw(const void* in, float* out, int n, int h, int w, int c, int dtype, int device, void* stream);
int fnst_nchw_to_nhwc(const float* in, void* out, int n, int h, int w, int c, int c_pad, int dtype, int device, void* stream);

/* uint8 HWC images <-> NCHW fp32 on the device (inference.py:28-31 ToTensor, :52-60 de-normalise/clamp).
 * mean3/std3 are HOST pointers to three floats.  u8->f32: (u8/255 - mean)/std;  f32->u8: trunc(clamp(y*std + mean, 0, 1)*255)
 * (torchvision ToPILImage on a float tensor is mul(255).byte(): truncation, inference.py:57-60). */
int fnst_u8_to_nchw(const void* in, float* out, int n, int h, int w, const float* mean3, const float* std3, int device, void* stream);
int fnst_nchw_to_u8(const float* in, void* out, int n, int h, int w, const float* mean3, const float* std3, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward operators (autograd of train.py:200 `total_loss.backward()` through the same modules).
 * Data gradients (dgrad) of every convolution are gather-GEMMs again (fnst_conv_tc / fnst_conv_simt
 * with negated taps and transposed packed weights); the entry points below are the rest.
 * ------------------------------------------------------------------------------------------------ */

/* Weight gradient of a gather-GEMM: dB[j][t*kc+c] = sum_{n,h,w} g[n,h,w,j] * A[n,h+h0+dh[t],w+w0+dw[t],c0[t]+c].
 * Uses d->a (A view, d->dtype), the taps, and: d->b = g, NHWC [out_n,out_h,out_w,n_gemm] of dtype g_dtype;
 * d->out = dB fp32 [n_gemm][ntaps*kc] (zeroed by the call). */
int fnst_wgrad_simt(const fnst_conv_desc* d, int g_dtype, int device, void* stream);
/* Same on tensor cores (tcgen05, MN-major operands, split over pixels with fp32 L2 atomics):
 * operand dtypes fp16/bf16 (they may differ), kc % 64 == 0. */
int fnst_wgrad_tc(const fnst_conv_desc* d, int g_dtype, int device, void* stream);

/* Weight gradient of fnst_conv_first (reflect/zero pad by index math): dw tap-major fp32 [3*k*k][c_out],
 * zeroed by the call; g NHWC [n,ho,wo,c_out] of g_dtype. */
int fnst_conv_first_wgrad(const float* x, int n, int h, int w, const void* g, int g_dtype, int c_out, int k, int stride,
                          int pad, int pad_mode, float* dw, int device, void* stream);

/*
 * InstanceNorm2d backward, pass 1.  Forms the gradient with respect to the normalised pre-activation
 *   gy = ( fold(gsrc) + extra ) * drop * [relu ? (y > 0) : 1],    y = raw*a + b  (a, b from stats/gamma/beta)
 * where gsrc is the gradient of the consumer's halo buffer (layout pad / pad_mode / s2d exactly as written by
 * fnst_inorm_apply; fold = ReflectionPad2d backward) and extra an optional plain [n,h,w,c] gradient (residual
 * branch).  Writes gy [n,h,w,c] (g_dtype) and accumulates sums[n][c] = (sum gy, sum gy*xhat) (zeroed by the call);
 * dgb (optional, fp32 [2][c], zeroed by the call) receives d gamma = sum_n sum gy*xhat and d beta = sum_n sum gy
 * (every block of the grid then adds to the same 2c addresses: prefer passing NULL here and dgb to pass 2).
 * prezeroed != 0: sums and dgb were zeroed by the caller (no memset is issued in front of the kernel).
 */
int fnst_inorm_bwd_reduce(const void* gsrc, const void* extra, const void* raw, const float* stats,
                          const float* gamma, const float* beta, const float* drop, void* gy, float* sums,
                          float* dgb, int n, int h, int w, int c, int act_dtype, int g_dtype, int relu, float eps,
                          int pad, int pad_mode, int s2d, int gsrc_slack, int prezeroed, int device, void* stream);
/* gsrc_slack: gsrc is allocated gsrc_slack rows and columns larger than its halo extent ([n][h+2pad+slack][w+2pad+slack][c];
 * space-to-depth: slack in space-to-depth pixels) -- the output of a pixel-stream data-gradient GEMM (FNST_DESC_LINEAR).
 * Pass 2: draw = gamma*rstd*(gy - mean(gy) - xhat*mean(gy*xhat)), written NHWC [n,h,w,c] or, if out_s2d,
 * space-to-depth [n,h/2,w/2,4c] (channel = ((h&1)*2+(w&1))*c + ch; h, w even).
 * dgb (optional, fp32 [2][c], written, not accumulated): d gamma = sum_n sums[n][c][1], d beta = sum_n sums[n][c][0].
 * out_pad > 0 (plain NHWC output only): draw is [n][h+2*out_pad][w+2*out_pad][c] with the gradient at (out_pad, out_pad) and a
 * ZERO halo written by this call (operand of a pixel-stream data-gradient GEMM, FNST_DESC_LINEAR). */
int fnst_inorm_bwd_apply(const void* gy, const void* raw, const float* stats, const float* sums, const float* gamma,
                         void* draw, float* dgb, int n, int h, int w, int c, int act_dtype, int g_dtype, float eps,
                         int out_s2d, int out_pad, int device, void* stream);

/*
 * InstanceNorm2d backward in ONE pass (same semantics as fnst_inorm_bwd_reduce followed by fnst_inorm_bwd_apply; autograd of
 * nn.InstanceNorm2d / F.relu / nn.Dropout2d / ReflectionPad2d / `x + y`, models/model.py:51-61,74-75,86-90 under train.py:200).
 * A slab (image, 16 channels) is staged by TMA box loads into the shared memory of a thread-block cluster of K CTAs; the
 * per-channel sums travel through distributed shared memory, so every tensor is read once and nothing is re-read or
 * atomically accumulated (gsrc: plain halo layout only -- a space-to-depth gradient buffer needs the two-pass operators):
 *   draw    [n,h,w,c] (or space-to-depth if out_s2d)   gradient of the raw conv output
 *   gy_out  or NULL  [n,h,w,c]                           the gradient gy itself (the residual branch re-uses it)
 *   sums    [n][c][2] fp32, WRITTEN (not accumulated)     (sum gy, sum gy*xhat) per plane, for fnst_affine_grads
 * fnst_inorm_bwd_fused_parts returns K for a configuration (0: w > 256, space-to-depth gsrc, or the plane does not fit 8 CTAs'
 * shared memory -> use the two-pass operators).
 */
int fnst_inorm_bwd_fused_parts(int n, int h, int w, int c, int act_dtype, int g_dtype, int has_gsrc, int has_extra, int s2d);
int fnst_inorm_bwd_fused(const void* gsrc, const void* extra, const void* raw, const float* stats,
                         const float* gamma, const float* beta, const float* drop, void* draw, void* gy_out,
                         float* sums, int n, int h, int w, int c, int act_dtype, int g_dtype, int relu, float eps,
                         int pad, int pad_mode, int s2d, int out_s2d, int device, void* stream);
/* d gamma / d beta of `layers` InstanceNorm layers in one launch: table = device array of {int64 src, int64 dgamma,
 * int64 dbeta, int32 C, int32 pad} (element offsets: the layer's [n][C][2] block in `sums`; its d gamma / d beta in `out`).
 * out[dgamma + c] = sum_i sums[src + (i*C + c)*2 + 1], out[dbeta + c] = sum_i sums[src + (i*C + c)*2] (fixed order). */
int fnst_affine_grads(const float* sums, const void* table, int layers, int n, int max_c, float* out, int device, void* stream);

/* MaxPool2d(2,2) backward fused with the ReLU mask of its input: gin = (extra + route(gout)) * (in > 0);
 * the first maximal element of each window receives the gradient (PyTorch tie rule). */
int fnst_maxpool2_bwd(const void* in, const void* gout, const void* extra, void* gin, int n, int h, int w, int c,
                      int act_dtype, int g_dtype, int device, void* stream);

/* da = 2 * coef * scale[0] * (a - b) (b broadcast with period b_period), optionally masked by (a > 0).  scale: device
 * scalar (the incoming gradient of the loss); coef: host constant (the loss's normalisation, losses/losses.py:41,54). */
int fnst_sse_bwd(const void* a, const void* b, int64_t count, int64_t b_period, int dtype_a, int dtype_b,
                 const float* scale, float coef, void* da, int g_dtype, int relu_mask, int device, void* stream);

/* out = (g + extra) * (act > 0): ReLU backward with an optional second gradient branch (count % 8 == 0). */
int fnst_relu_mask(const void* g, const void* extra, const void* act, void* out, int64_t count, int act_dtype,
                   int g_dtype, int device, void* stream);

/* Style-loss backward factor (losses/losses.py:15-44): S[n] = scale[0]*coef*((G[n]-Gt) + (G[n]-Gt)^T) in element type
 * out_dtype; gt is broadcast with period gt_numel.  The feature gradient is then the per-image 1x1 gather-GEMM F[n] * S[n]. */
int fnst_gram_diff_sym(const float* g, const float* gt, int n, int c, int64_t gt_numel, const float* scale, float coef,
                       void* s_out, int out_dtype, int device, void* stream);

/* Gradient of fnst_tv: dimg = coef * scale[0] * d/dimg sum(dh^2 + dw^2), NCHW fp32 (scale: device scalar, coef: host constant). */
int fnst_tv_bwd(const float* img, int planes, int h, int w, const float* scale, float coef, float* dimg, int device, void* stream);

/* Element-type conversion of a contiguous tensor (count % 8 == 0).  tcgen05 kind::f16 needs both operands in the
 * same 16-bit format, so saved fp16 activations are converted to the bf16 gradient format for fnst_wgrad_tc. */
int fnst_cast(const void* in, void* out, int64_t count, int in_dtype, int out_dtype, int device, void* stream);

/* Weight re-layout: out[i] = idx[i] < 0 ? 0 : src[idx[i]] converted to out_dtype (count elements; idx int32 on the device).
 * One launch per packed operand: PyTorch OIHW / IOHW parameters -> gather-GEMM operands of include/fnst.h's conv
 * descriptor (and packed fp32 weight gradients -> parameter layout), with index maps cached by the host layer. */
int fnst_gather_cast(const void* src, int src_dtype, const int32_t* idx, void* out, int out_dtype, int64_t count,
                     int device, void* stream);

/* out[c] = sum over n,h,w of x[n,c,h,w] (fp32 NCHW); final_conv bias gradient. */
int fnst_channel_sum(const float* x, int n, int c, int hw, float* out, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer tail of the training step (train.py:203-205): torch.nn.utils.clip_grad_norm_ and
 * torch.optim.Adam(betas, eps, weight_decay = coupled L2, amsgrad=False) as multi-tensor kernels over the 58
 * parameter tensors.  Tensor lists are HOST arrays of n device pointers to contiguous fp32 tensors of numels[i]
 * elements (1 <= numels[i] < 2^31); any n (64 tensors per launch).
 * ------------------------------------------------------------------------------------------------ */

/* Bytes of the device workspace of fnst_grad_norm (8-byte aligned; zeroed ONCE by the caller at allocation: the kernel
 * leaves it zeroed again at the end of every call). */
int64_t fnst_grad_norm_workspace_bytes(void);

/* clip_grad_norm_ (train.py:203), part 1: norm_coef[0] = total 2-norm of all gradients (fp64 accumulation),
 * norm_coef[1] = min(1, max_norm / (norm + 1e-6)) -- torch's clip coefficient.  No host synchronisation. */
int fnst_grad_norm(void* const* grads, const int64_t* numels, int n, void* workspace, float max_norm, float* norm_coef,
                   int device, void* stream);

/* clip_grad_norm_, part 2: g *= coef[0] in place for every gradient (coef on the device, e.g. norm_coef + 1). */
int fnst_grad_scale(void* const* grads, const int64_t* numels, int n, const float* coef, int device, void* stream);

/* One Adam step (train.py:205 with the optimizer of train.py:135-139), the arithmetic of torch.optim.Adam:
 *   g = grad_scale[0] * g (if grad_scale != NULL: clip fused into the update);   g += weight_decay * p;
 *   m += (1 - beta1) * (g - m);   v = beta2 * v + (1 - beta2) * g * g;
 *   p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps).
 * step counts from 1.  params / exp_avg / exp_avg_sq are updated in place; grads are only read. */
int fnst_adam_step(void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                   const int64_t* numels, int n, double lr, double beta1, double beta2, double eps, double weight_decay,
                   int64_t step, const float* grad_scale, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input pre-processing (device side): the per-image transform of the reference's data pipeline,
 *   transforms.Resize((out_h, out_w)) -> transforms.ToTensor() [-> transforms.Normalize(mean, std)]
 * (train.py:92-102, inference.py:28-31, data/dataset.py:21-27).  On a PIL image, Resize is Pillow's
 * Image.resize(BILINEAR): two-pass 8-bit resampling with 22-bit fixed-point weights; reproduced bit for bit.
 * ------------------------------------------------------------------------------------------------ */

/* img_hwc: decoded RGB image, uint8 [in_h][in_w][3] with in_pitch_bytes per row, on the device.
 * out_chw (optional): float32 [3][out_h][out_w] = u8/255, then (x-mean)/std if mean3/std3 (HOST float[3]) are given.
 * out_u8_hwc (optional): the resized uint8 image [out_h][out_w][3].  Down-scaling factors up to 35 per axis. */
int fnst_resize_to_tensor(const void* img_hwc, int in_h, int in_w, int64_t in_pitch_bytes, int out_h, int out_w,
                          float* out_chw, void* out_u8_hwc, const float* mean3, const float* std3, int device, void* stream);

/* Many images per launch: `images_dev` is a DEVICE array of n descriptors (decoded uint8 RGB images, sizes may differ;
 * max_in_h / max_in_w bound them: they size the kernel's shared-memory strip and the down-scaling check).  Image i is written to
 * out_nchw[i] ([n][3][out_h][out_w] float) and / or out_u8_nhwc[i] ([n][out_h][out_w][3]); same arithmetic as
 * fnst_resize_to_tensor, bit for bit.  Replaces the per-sample transform calls behind DataLoader's collate (train.py:98-107). */
typedef struct fnst_image_desc {
  const void* data;        /* uint8 [h][w][3] (device), rows pitch_bytes apart */
  int32_t h, w;
  int64_t pitch_bytes;
} fnst_image_desc;
int fnst_resize_batch_to_tensor(const fnst_image_desc* images_dev, int n, int max_in_h, int max_in_w, int out_h, int out_w,
                                float* out_nchw, void* out_u8_nhwc, const float* mean3, const float* std3, int device,
                                void* stream);
/* Test hook (no device work): the resampling window of output index `index` along one axis, computed on the HOST by the
 * same functions the kernel runs on the device.  Returns the window capacity ksize (> 0), or < 0 on bad arguments. */
int fnst_resize_window_host(int in_size, int out_size, int index, int* first, int* len, int* kk, int kk_capacity);

#ifdef __cplusplus
}
#endif
#endif /* FNST_H_ */
