// CUDA-core convolutions (fp32 accumulate): the gather-GEMM operator of fnst.h on FFMA pipes
// (the fp32-accurate path and the functional twin of the tcgen05 kernel), plus the 3-channel
// first-layer convolution that reads NCHW fp32 images with reflect/zero padding by index math.
#include "common.cuh"

namespace fnst {

// Shared epilogue: one output element (pixel n,h,w ; GEMM column j ; value v already biased/relu'd).
template <typename TOut>
__device__ __forceinline__ void store_out(const fnst_conv_desc& d, int n, int h, int w, int j, float v) {
  if (d.epilogue == FNST_EPI_NHWC) {
    size_t idx = (((size_t)n * d.out_h + h) * d.out_w + w) * d.c_out + j;
    reinterpret_cast<TOut*>(d.out)[idx] = from_f32<TOut>(v);
  } else if (d.epilogue == FNST_EPI_D2S) {
    int phase = j / d.c_out, o = j - phase * d.c_out;
    int ho = 2 * h + (phase >> 1), wo = 2 * w + (phase & 1);
    size_t idx = (((size_t)n * (2 * d.out_h) + ho) * (2 * d.out_w) + wo) * d.c_out + o;
    reinterpret_cast<TOut*>(d.out)[idx] = from_f32<TOut>(v);
  } else {  // FNST_EPI_NCHW_F32
    size_t idx = (((size_t)n * d.c_out + j) * d.out_h + h) * d.out_w + w;
    reinterpret_cast<float*>(d.out)[idx] = v;
  }
}

constexpr int SB_M = 64, SB_N = 64, SB_K = 16;

template <typename T, typename TOut>
__global__ void __launch_bounds__(256) conv_simt_kernel(const __grid_constant__ fnst_conv_desc d) {
  __shared__ float As[SB_K][SB_M + 4];
  __shared__ float Bs[SB_K][SB_N + 4];
  __shared__ float s_sum[SB_N], s_sq[SB_N];
  __shared__ int s_dh[FNST_MAX_TAPS], s_dw[FNST_MAX_TAPS], s_c0[FNST_MAX_TAPS];

  const int tid = threadIdx.x;
  const int tiles_w = (d.out_w + 7) >> 3, tiles_h = (d.out_h + 7) >> 3;
  int tile = blockIdx.x;
  const int tw = tile % tiles_w; tile /= tiles_w;
  const int th = tile % tiles_h;
  const int n = tile / tiles_h;
  const int col0 = blockIdx.y * SB_N;

  for (int i = tid; i < d.ntaps; i += 256) { s_dh[i] = d.tap_dh[i]; s_dw[i] = d.tap_dw[i]; s_c0[i] = d.tap_c0[i]; }
  if (tid < SB_N) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
  __syncthreads();

  // loader mapping
  const int lp = tid >> 2, kq = (tid & 3) * 4;
  const int lh = th * 8 + (lp >> 3), lw = tw * 8 + (lp & 7);
  const bool lvalid = lh < d.out_h && lw < d.out_w;
  const int lcol = col0 + lp;
  const bool cvalid = lcol < d.n_gemm;
  const T* A = reinterpret_cast<const T*>(d.a);
  const T* B = reinterpret_cast<const T*>(d.b);
  const size_t ktot = (size_t)d.ntaps * d.kc;

  // compute mapping
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < d.ntaps; ++t) {
    const int ih = lh + d.h0 + s_dh[t], iw = lw + d.w0 + s_dw[t];
    const bool in_ok = lvalid && ih >= 0 && ih < d.a_h && iw >= 0 && iw < d.a_w;
    const T* ap = A + (size_t)n * d.a_stride_n + (size_t)(in_ok ? ih : 0) * d.a_stride_h +
                  (size_t)(in_ok ? iw : 0) * d.a_stride_w + s_c0[t] + kq;
    const T* bp = B + ((size_t)n * d.b_image_rows + (size_t)(cvalid ? lcol : 0)) * ktot + (size_t)t * d.kc + kq;
    for (int c = 0; c < d.kc; c += SB_K) {
      float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (in_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = to_f32<T>(ap[c + i]);
      }
      if (cvalid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bv[i] = to_f32<T>(bp[c + i]);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) { As[kq + i][lp] = av[i]; Bs[kq + i][lp] = bv[i]; }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SB_K; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  // epilogue
  float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = ty * 4 + i;
    const int h = th * 8 + (p >> 3), w = tw * 8 + (p & 7);
    if (h >= d.out_h || w >= d.out_w) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= d.n_gemm) continue;
      const int ch = d.epilogue == FNST_EPI_D2S ? col % d.c_out : col;
      if (ch >= d.c_out) continue;
      float v = acc[i][j];
      if (d.bias) v += d.bias[ch];
      if (d.relu) v = fmaxf(v, 0.f);
      if (d.epilogue == FNST_EPI_NHWC && (d.addend || d.mask)) {
        const size_t idx = (((size_t)n * d.out_h + h) * d.out_w + w) * d.c_out + col;
        if (d.addend) v += load_scalar_f32(d.addend, d.out_dtype, idx);
        if (d.mask && !(load_scalar_f32(d.mask, d.mask_dtype, idx) > 0.f)) v = 0.f;
      }
      csum[j] += v; csq[j] += v * v;
      store_out<TOut>(d, n, h, w, col, v);
    }
  }
  if (d.stats) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&s_sum[tx * 4 + j], csum[j]); atomicAdd(&s_sq[tx * 4 + j], csq[j]); }
    __syncthreads();
    if (tid < SB_N) {
      const int col = col0 + tid;
      if (col < d.n_gemm) {
        const int ch = d.epilogue == FNST_EPI_D2S ? col % d.c_out : col;
        if (ch < d.c_out) {
          atomicAdd(&d.stats[((size_t)n * d.c_out + ch) * 2 + 0], s_sum[tid]);
          atomicAdd(&d.stats[((size_t)n * d.c_out + ch) * 2 + 1], s_sq[tid]);
        }
      }
    }
  }
}

int validate_conv_desc(const fnst_conv_desc* d) {
  FNST_CHECK_ARG(d != nullptr, "null conv desc");
  FNST_CHECK_ARG(d->a && d->b && d->out, "conv: null pointer");
  FNST_CHECK_ARG(d->ntaps > 0 && d->ntaps <= FNST_MAX_TAPS, "conv: ntaps %d out of range", d->ntaps);
  FNST_CHECK_ARG(d->kc > 0 && d->kc % 16 == 0, "conv: kc %d must be a multiple of 16", d->kc);
  FNST_CHECK_ARG(d->n_gemm > 0 && d->n_gemm % 16 == 0, "conv: n_gemm %d must be a multiple of 16", d->n_gemm);
  FNST_CHECK_ARG(d->out_n > 0 && d->out_h > 0 && d->out_w > 0, "conv: empty output");
  FNST_CHECK_ARG(d->out_n == d->a_n, "conv: batch mismatch");
  FNST_CHECK_ARG(d->epilogue >= 0 && d->epilogue <= 3, "conv: bad epilogue %d", d->epilogue);
  if (d->epilogue == FNST_EPI_D2S) FNST_CHECK_ARG(d->n_gemm == 4 * d->c_out, "conv: d2s needs n_gemm == 4*c_out");
  else FNST_CHECK_ARG(d->c_out <= d->n_gemm, "conv: c_out > n_gemm");
  for (int t = 0; t < d->ntaps; ++t)
    FNST_CHECK_ARG(d->tap_c0[t] >= 0 && d->tap_c0[t] + d->kc <= d->a_c, "conv: tap %d channel window out of range", t);
  return 0;
}

}  // namespace fnst

using namespace fnst;

extern "C" int fnst_conv_simt(const fnst_conv_desc* d, int device, void* stream) {
  if (int r = validate_conv_desc(d)) return r;
  FNST_CHECK_ARG(d->epilogue != FNST_EPI_ROWSUM9, "conv_simt: the ROWSUM9 epilogue exists on the tensor-core kernel only");
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  if (d->stats && !(d->flags & FNST_DESC_PREZEROED)) FNST_CUDA(cudaMemsetAsync(d->stats, 0, sizeof(float) * 2 * (size_t)d->out_n * d->c_out, st));
  dim3 grid(((d->out_w + 7) / 8) * ((d->out_h + 7) / 8) * d->out_n, (d->n_gemm + SB_N - 1) / SB_N);
  const int odt = d->epilogue == FNST_EPI_NCHW_F32 ? FNST_F32 : d->out_dtype;
  FNST_DISPATCH_DTYPE(d->dtype, T, {
    FNST_DISPATCH_DTYPE(odt, TOut, { conv_simt_kernel<T, TOut><<<grid, 256, 0, st>>>(*d); });
  });
  return launch_status("conv_simt");
}

// ---------------------------------------------------------------------------------------------
// First-layer convolution, C_in = 3, NCHW fp32 input.
// Block = 8x8 output pixels x all c_out channels; 256 threads = 64 pixels x 4 channel groups.
// Dynamic smem: weights [3*k*k][c_out] fp32 + input patch [3][ph][pw] fp32.
// ---------------------------------------------------------------------------------------------
namespace fnst {

template <typename TOut, int CPT>   // CPT = channels per thread (c_out / 4)
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x, int n_img, int H, int W,
                                                         const float* __restrict__ wgt, const float* __restrict__ bias,
                                                         int c_out, int k, int stride, int pad, int pad_mode, int relu,
                                                         TOut* __restrict__ out, int Ho, int Wo, float* __restrict__ stats) {
  extern __shared__ float smem[];
  const int taps = 3 * k * k;
  float* ws = smem;                         // [taps][c_out]
  const int pdim = 7 * stride + k;          // patch extent for 8 outputs
  float* patch = smem + taps * c_out;       // [3][pdim][pdim]
  float* s_stat = patch + 3 * pdim * pdim;  // [2][c_out]

  const int tid = threadIdx.x;
  const int tiles_w = (Wo + 7) >> 3, tiles_h = (Ho + 7) >> 3;
  int tile = blockIdx.x;
  const int tw = tile % tiles_w; tile /= tiles_w;
  const int th = tile % tiles_h;
  const int n = tile / tiles_h;

  // weights arrive tap-major [(c*k+kh)*k+kw][o]: straight, conflict-free copy
  for (int i = tid * 4; i < taps * c_out; i += 1024)
    *reinterpret_cast<float4*>(ws + i) = *reinterpret_cast<const float4*>(wgt + i);
  for (int i = tid; i < 2 * c_out; i += 256) s_stat[i] = 0.f;
  const int h_base = th * 8 * stride - pad, w_base = tw * 8 * stride - pad;
  for (int i = tid; i < 3 * pdim * pdim; i += 256) {
    const int c = i / (pdim * pdim), r = i - c * pdim * pdim;
    int ih = h_base + r / pdim, iw = w_base + r % pdim;
    float v = 0.f;
    if (pad_mode == FNST_PAD_REFLECT) {
      // coordinates needed by valid outputs are always within the reflect range; clamp the rest
      ih = reflect_index(ih, H); iw = reflect_index(iw, W);
      ih = min(max(ih, 0), H - 1); iw = min(max(iw, 0), W - 1);
      v = x[(((size_t)n * 3 + c) * H + ih) * W + iw];
    } else if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      v = x[(((size_t)n * 3 + c) * H + ih) * W + iw];
    }
    patch[i] = v;
  }
  __syncthreads();

  const int p = tid & 63, q = tid >> 6;     // a warp shares q -> weight reads broadcast
  const int ph = p >> 3, pw = p & 7;
  const int h = th * 8 + ph, w = tw * 8 + pw;
  float acc[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
  for (int c = 0; c < 3; ++c)
    for (int kh = 0; kh < k; ++kh) {
      const float* prow = patch + (c * pdim + ph * stride + kh) * pdim + pw * stride;
      const float* wrow = ws + ((c * k + kh) * k) * c_out + q * CPT;
      for (int kw = 0; kw < k; ++kw) {
        const float xv = prow[kw];
        const float* wp = wrow + kw * c_out;
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + j);
          acc[j] = fmaf(xv, w4.x, acc[j]); acc[j + 1] = fmaf(xv, w4.y, acc[j + 1]);
          acc[j + 2] = fmaf(xv, w4.z, acc[j + 2]); acc[j + 3] = fmaf(xv, w4.w, acc[j + 3]);
        }
      }
    }
  const bool valid = h < Ho && w < Wo;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    float v = acc[j];
    if (bias) v += bias[q * CPT + j];
    if (relu) v = fmaxf(v, 0.f);
    acc[j] = valid ? v : 0.f;
  }
  if (valid) {
    TOut* op = out + (((size_t)n * Ho + h) * Wo + w) * c_out + q * CPT;
#pragma unroll
    for (int j = 0; j < CPT; j += 8) {
      float v8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = acc[j + i];
      store8<TOut>(op + j, v8);
    }
  }
  if (stats) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const float s = warp_sum(acc[j]), s2 = warp_sum(acc[j] * acc[j]);
      if ((tid & 31) == 0) { atomicAdd(&s_stat[q * CPT + j], s); atomicAdd(&s_stat[c_out + q * CPT + j], s2); }
    }
    __syncthreads();
    for (int i = tid; i < c_out; i += 256) {
      atomicAdd(&stats[((size_t)n * c_out + i) * 2 + 0], s_stat[i]);
      atomicAdd(&stats[((size_t)n * c_out + i) * 2 + 1], s_stat[c_out + i]);
    }
  }
}

}  // namespace fnst

extern "C" int fnst_conv_first(const float* x, int n, int h, int w, const float* wgt, const float* bias,
                               int c_out, int k, int stride, int pad, int pad_mode, int relu,
                               void* out, int out_dtype, float* stats, int device, void* stream) {
  FNST_CHECK_ARG(x && wgt && out, "conv_first: null pointer");
  FNST_CHECK_ARG(c_out == 64, "conv_first: c_out %d unsupported (64 only)", c_out);
  FNST_CHECK_ARG(k % 2 == 1 && k <= 9 && (stride == 1 || stride == 2), "conv_first: unsupported k=%d stride=%d", k, stride);
  FNST_CHECK_ARG(pad_mode == FNST_PAD_REFLECT || pad_mode == FNST_PAD_ZERO, "conv_first: bad pad mode");
  if (pad_mode == FNST_PAD_REFLECT) FNST_CHECK_ARG(h > pad && w > pad, "conv_first: reflect pad %d needs h,w > pad", pad);
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  FNST_CHECK_ARG(ho > 0 && wo > 0, "conv_first: empty output");
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  if (stats) FNST_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * (size_t)n * c_out, st));
  const int pdim = 7 * stride + k;
  const size_t smem = sizeof(float) * ((size_t)3 * k * k * c_out + 3 * pdim * pdim + 2 * c_out);
  dim3 grid(((wo + 7) / 8) * ((ho + 7) / 8) * n);
  FNST_DISPATCH_DTYPE(out_dtype, TOut, {
    auto kern = conv_first_kernel<TOut, 16>;
    FNST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(x, n, h, w, wgt, bias, c_out, k, stride, pad, pad_mode, relu,
                                 reinterpret_cast<TOut*>(out), ho, wo, stats);
  });
  return launch_status("conv_first");
}
