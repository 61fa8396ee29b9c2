// HBM-bound kernels of the style-transfer path: InstanceNorm apply (+ReLU, dropout scale,
// residual add, halo write), 2x2 max-pool, squared-error and total-variation reductions,
// NHWC<->NCHW layout conversion.  All use 16/32-byte vector accesses on the channel dimension.
#include "common.cuh"
#include <cstdlib>

namespace fnst {

// -------------------------------------------------------------------------------------------
// InstanceNorm apply.  One block per (padded output row, image).  256 threads = CG channel
// groups of 8 channels x PL pixel lanes; per-channel scale/shift are computed once per thread.
// -------------------------------------------------------------------------------------------
// TI: element type of the raw conv output; T: element type of the activation written (and of the residual).
// SPLIT: the activation is an error-compensated pair, stored as 2C channels per pixel [hi(C) | lo(C)] with
// hi = T(y), lo = T(y - hi) (fp16x3 path); the residual buffer then has the same split layout.
template <typename TI, typename T, int U, bool SPLIT>
__global__ void __launch_bounds__(256) inorm_apply_kernel(const TI* __restrict__ raw, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ drop, const T* __restrict__ res, int res_pad,
                                                          T* __restrict__ out, __nv_bfloat16* __restrict__ out2, int H, int W, int C,
                                                          int relu, float eps, int pad, int pad_mode, int s2d, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_ab[];          // [2][C]: per-channel scale a and shift b of this image (keeps registers low)
  const int n = blockIdx.y;
  {
    const float inv_cnt = 1.f / (float)(H * W);
    for (int c = threadIdx.x; c < C; c += 256) {
      const float* st = stats + ((size_t)n * C + c) * 2;
      const float mean = st[0] * inv_cnt;
      const float var = fmaxf(st[1] * inv_cnt - mean * mean, 0.f);
      float ai = gamma[c] * rsqrtf(var + eps);
      float bi = beta[c] - mean * ai;
      if (drop) { const float ds = drop[(size_t)n * C + c]; ai *= ds; bi *= ds; }
      s_ab[c] = ai; s_ab[C + c] = bi;
    }
  }
  __syncthreads();
  const int CG = C >> 3, PL = 256 / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  if (pl >= PL) return;
  const int c0 = cg * 8;
  const int CO = SPLIT ? 2 * C : C;        // channels per pixel of the activation buffers
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int Hp2 = (Hp + 1) >> 1, Wp2 = (Wp + 1) >> 1;

  for (int rr = 0; rr < rows_per_block; ++rr) {
    const int hp = blockIdx.x * rows_per_block + rr;
    if (hp >= Hp) break;
    int sh = hp - pad;
    bool row_zero = false;
    if (sh < 0 || sh >= H) {
      if (pad_mode == FNST_PAD_REFLECT) sh = reflect_index(sh, H); else row_zero = true;
    }
    const TI* raw_row = raw + ((size_t)n * H + sh) * W * C + c0;
    const T* res_row = res ? res + (((size_t)n * (H + 2 * res_pad) + sh + res_pad) * (W + 2 * res_pad) + res_pad) * CO + c0 : nullptr;
    for (int wp0 = pl; wp0 < Wp; wp0 += U * PL) {
      Raw8<TI> rv[U];
      Raw8<T> sv[U], sl[U];
      bool live[U], zero[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {          // U pixels in flight: issue all loads of the group first (raw registers)
        const int wp = wp0 + u * PL;
        live[u] = wp < Wp;
        int sw = wp - pad;
        zero[u] = row_zero;
        if (sw < 0 || sw >= W) {
          if (pad_mode == FNST_PAD_REFLECT) sw = reflect_index(sw, W); else zero[u] = true;
        }
        if (live[u] && !zero[u]) {
          rv[u] = load_raw8<TI>(raw_row + (size_t)sw * C);
          if (res) {
            sv[u] = load_raw8<T>(res_row + (size_t)sw * CO);
            if (SPLIT) sl[u] = load_raw8<T>(res_row + (size_t)sw * CO + C);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;
        const int wp = wp0 + u * PL;
        float vv[8];
        if (zero[u]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) vv[i] = 0.f;
        } else {
          raw8_to_f32<TI>(rv[u], vv);
          const float4 a0 = *reinterpret_cast<const float4*>(s_ab + c0), a1 = *reinterpret_cast<const float4*>(s_ab + c0 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(s_ab + C + c0), b1 = *reinterpret_cast<const float4*>(s_ab + C + c0 + 4);
          const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float y = fmaf(vv[i], a[i], b[i]);
            vv[i] = relu ? fmaxf(y, 0.f) : y;
          }
          if (res) {
            float r[8];
            raw8_to_f32<T>(sv[u], r);
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] += r[i];
            if (SPLIT) {
              raw8_to_f32<T>(sl[u], r);
#pragma unroll
              for (int i = 0; i < 8; ++i) vv[i] += r[i];
            }
          }
        }
        T* dst;
        if (!s2d) dst = out + (((size_t)n * Hp + hp) * Wp + wp) * CO + c0;
        else dst = out + (((size_t)n * Hp2 + (hp >> 1)) * Wp2 + (wp >> 1)) * (4 * CO) + ((hp & 1) * 2 + (wp & 1)) * CO + c0;
        if (out2) {
          // second copy of the same activation in bfloat16, same geometry with C channels per pixel: the weight-gradient
          // GEMM's operand (kind::f16 MMAs take ONE 16-bit format and gradients are bf16), written here instead of by a
          // separate cast pass over the tensor
          __nv_bfloat16* d2 = !s2d ? out2 + (((size_t)n * Hp + hp) * Wp + wp) * C + c0
                                   : out2 + (((size_t)n * Hp2 + (hp >> 1)) * Wp2 + (wp >> 1)) * (4 * C) + ((hp & 1) * 2 + (wp & 1)) * C + c0;
          store8<__nv_bfloat16>(d2, vv);
        }
        if (SPLIT) {
          float hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { hi[i] = to_f32<T>(from_f32<T>(vv[i])); lo[i] = vv[i] - hi[i]; }
          store8<T>(dst, hi);
          store8<T>(dst + C, lo);
        } else {
          store8<T>(dst, vv);
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool2_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                       int N, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H >> 1, Wo = W >> 1, CG = C >> 3;
  const size_t total = (size_t)N * Ho * Wo * CG;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cg = i % CG; size_t r = i / CG;
    const int wo = r % Wo; r /= Wo;
    const int ho = r % Ho; const int n = r / Ho;
    const T* p = in + (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + cg * 8;
    float v0[8], v1[8], v2[8], v3[8];
    load8<T>(p, v0); load8<T>(p + C, v1); load8<T>(p + (size_t)W * C, v2); load8<T>(p + (size_t)W * C + C, v3);
#pragma unroll
    for (int k = 0; k < 8; ++k) v0[k] = fmaxf(fmaxf(v0[k], v1[k]), fmaxf(v2[k], v3[k]));
    store8<T>(out + (((size_t)n * Ho + ho) * Wo + wo) * C + cg * 8, v0);
  }
}

__device__ __forceinline__ void block_accumulate(float v, double* acc) {
  __shared__ float s_part[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < (blockDim.x >> 5) ? s_part[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) atomicAdd(acc, (double)t);
  }
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) sse_kernel(const TA* __restrict__ a, const TB* __restrict__ b, int64_t count,
                                                  int64_t period, double* acc) {
  pdl_trigger();
  pdl_wait();
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = to_f32<TA>(a[i]) - to_f32<TB>(b[i % period]);
    s = fmaf(d, d, s);
  }
  block_accumulate(s, acc);
}

__global__ void __launch_bounds__(256) tv_kernel(const float* __restrict__ img, int planes, int H, int W, double* acc) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)planes * H * W;
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = i % W; const int y = (i / W) % H;
    const float v = img[i];
    if (y + 1 < H) { const float d = img[i + W] - v; s = fmaf(d, d, s); }
    if (x + 1 < W) { const float d = img[i + 1] - v; s = fmaf(d, d, s); }
  }
  block_accumulate(s, acc);
}

// ---- loss scalars in one launch (no memset in front, no host arithmetic behind) ---------------------------------
// workspace: double partial[LOSS_MAX_BLOCKS] followed by one unsigned counter; zeroed once by the caller.  Every block
// stores its partial, the last block to finish sums them in block order (deterministic), writes scale * sum to out[0]
// and leaves the counter at zero for the next launch.
constexpr int LOSS_MAX_BLOCKS = 148 * 16;

__device__ __forceinline__ void loss_finish(float v, double* ws, float scale, float* out, int accumulate) {
  __shared__ float s_part[32];
  __shared__ bool s_last;
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += (double)s_part[i];
    ws[blockIdx.x] = t;
    __threadfence();
    unsigned* counter = reinterpret_cast<unsigned*>(ws + LOSS_MAX_BLOCKS);
    s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last block: ordered sum of the partials (pairwise over 256 strided lanes, fixed order)
  double t = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(ws + i);
  __shared__ double s_d[256];
  s_d[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_d[threadIdx.x] += s_d[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float r = (float)(s_d[0] * (double)scale);
    out[0] = accumulate ? out[0] + r : r;
    *reinterpret_cast<unsigned*>(ws + LOSS_MAX_BLOCKS) = 0u;
  }
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) sse_scaled_kernel(const TA* __restrict__ a, const TB* __restrict__ b, int64_t count,
                                                         int64_t period, float scale, double* ws, float* out, int accumulate) {
  pdl_trigger();
  pdl_wait();
  float s = 0.f;
  if (period == count && count % 8 == 0) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count / 8; i += (int64_t)gridDim.x * blockDim.x) {
      float x[8], y[8];
      load8<TA>(a + i * 8, x);
      load8<TB>(b + i * 8, y);
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = x[k] - y[k]; s = fmaf(d, d, s); }
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
      const float d = to_f32<TA>(a[i]) - to_f32<TB>(b[i % period]);
      s = fmaf(d, d, s);
    }
  }
  loss_finish(s, ws, scale, out, accumulate);
}

__global__ void __launch_bounds__(256) tv_scaled_kernel(const float* __restrict__ img, int planes, int H, int W, float scale,
                                                        double* ws, float* out) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)planes * H * W;
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = i % W; const int y = (i / W) % H;
    const float v = img[i];
    if (y + 1 < H) { const float d = img[i + W] - v; s = fmaf(d, d, s); }
    if (x + 1 < W) { const float d = img[i + 1] - v; s = fmaf(d, d, s); }
  }
  loss_finish(s, ws, scale, out, 0);
}

// [HW][C] (T) <-> [C][HW] (fp32) per image, 32x32 smem tiles.
template <typename T, bool TO_NCHW>
__global__ void __launch_bounds__(256) transpose_kernel(const void* __restrict__ in_, void* __restrict__ out_, int HW, int C, int CP) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  if (TO_NCHW) {
    const T* in = reinterpret_cast<const T*>(in_) + (size_t)n * HW * CP;
    float* out = reinterpret_cast<float*>(out_) + (size_t)n * HW * C;
    for (int r = ty; r < 32; r += 8) {
      const int p = p0 + r, c = c0 + tx;
      tile[r][tx] = (p < HW && c < C) ? to_f32<T>(in[(size_t)p * CP + c]) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int c = c0 + r, p = p0 + tx;
      if (p < HW && c < C) out[(size_t)c * HW + p] = tile[tx][r];
    }
  } else {
    const float* in = reinterpret_cast<const float*>(in_) + (size_t)n * HW * C;
    T* out = reinterpret_cast<T*>(out_) + (size_t)n * HW * CP;
    for (int r = ty; r < 32; r += 8) {
      const int c = c0 + r, p = p0 + tx;
      tile[r][tx] = (p < HW && c < C) ? in[(size_t)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int p = p0 + r, c = c0 + tx;
      if (p < HW && c < CP) out[(size_t)p * CP + c] = from_f32<T>(tile[tx][r]);   // channels >= C come out as zero
    }
  }
}

// NCHW fp32 3-channel image -> NHWC halo buffer [n][rows][pitch][CP] (CP = 4 or 8 channels, channel 3.. zero),
// interior at (pad, pad), reflect or zero border, everything outside [-(pad), H+pad) x [-(pad), W+pad) zero.
template <typename T, int CP>
__global__ void __launch_bounds__(256) image_to_halo_kernel(const float* __restrict__ x, T* __restrict__ out, int N, int H, int W,
                                                            int pad, int reflect, int rows, int pitch, int split) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)N * rows * pitch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int wp = i % pitch; int64_t r = i / pitch;
    const int hp = r % rows; const int n = r / rows;
    int h = hp - pad, w = wp - pad;
    bool ok = h >= -pad && h < H + pad && w >= -pad && w < W + pad;
    if (ok && (h < 0 || h >= H || w < 0 || w >= W)) {
      if (reflect) { h = reflect_index(h, H); w = reflect_index(w, W); } else ok = false;
    }
    T v[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) v[c] = from_f32<T>(0.f);
    if (ok) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float f = x[(((size_t)n * 3 + c) * H + h) * W + w];
        v[c] = from_f32<T>(f);
        if (CP == 8 && split) v[4 + c] = from_f32<T>(f - to_f32<T>(v[c]));      // channels 4..6: low part (hi|lo pair)
      }
    }
    T* o = out + i * CP;
    if (CP == 4) *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(v);
    else *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(v);
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t count8) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    load8<TI>(in + i * 8, v);
    store8<TO>(out + i * 8, v);
  }
}

// out[i] = idx[i] < 0 ? 0 : src[idx[i]], converted to the output element type: every weight re-layout of the path
// (parameter layout <-> packed gather-GEMM operand, sub-pixel / space-to-depth / window forms, and their inverses for
// the gradients) is one launch of this kernel with a cached index map (engine.gather_pack).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gather_cast_kernel(const TI* __restrict__ src, const int32_t* __restrict__ idx,
                                                          TO* __restrict__ out, int64_t count) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = idx[i];
    out[i] = from_f32<TO>(j < 0 ? 0.f : to_f32<TI>(src[j]));
  }
}

// uint8 HWC images <-> NCHW fp32 (inference pre/post-processing on the device, inference.py:28-31,52-60):
//   u8 -> f32:  x[n,c,h,w] = (u8[n,h,w,c] / 255 - mean[c]) / std[c]        (mean = 0, std = 1: plain ToTensor)
//   f32 -> u8:  u8[n,h,w,c] = trunc(clamp(y[n,c,h,w] * std[c] + mean[c], 0, 1) * 255)   (ToPILImage on a float tensor is
//               pic.mul(255).byte(), i.e. truncation, inference.py:57-60)
__global__ void __launch_bounds__(256) u8_to_nchw_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t pixels,
                                                         int HW, float m0, float m1, float m2, float s0, float s1, float s2) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    const uint8_t* px = in + i * 3;
    float* o = out + n * 3 * (int64_t)HW + p;
    o[0] = (px[0] * (1.f / 255.f) - m0) / s0;
    o[HW] = (px[1] * (1.f / 255.f) - m1) / s1;
    o[2 * (int64_t)HW] = (px[2] * (1.f / 255.f) - m2) / s2;
  }
}

__global__ void __launch_bounds__(256) nchw_to_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int64_t pixels,
                                                         int HW, float m0, float m1, float m2, float s0, float s1, float s2) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    const float* y = in + n * 3 * (int64_t)HW + p;
    uint8_t* px = out + i * 3;
    px[0] = (uint8_t)__float2int_rz(fminf(fmaxf(y[0] * s0 + m0, 0.f), 1.f) * 255.f);
    px[1] = (uint8_t)__float2int_rz(fminf(fmaxf(y[HW] * s1 + m1, 0.f), 1.f) * 255.f);
    px[2] = (uint8_t)__float2int_rz(fminf(fmaxf(y[2 * (int64_t)HW] * s2 + m2, 0.f), 1.f) * 255.f);
  }
}

static int grid_for(int64_t work_items, int threads = 256) {
  int64_t b = (work_items + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace fnst

using namespace fnst;

extern "C" int fnst_inorm_apply(const void* raw, const float* stats, const float* gamma, const float* beta,
                                const float* drop, const void* res, int res_pad, void* out, void* out_bf16,
                                int n, int h, int w, int c, int dtype, int relu, float eps,
                                int pad, int pad_mode, int s2d, int raw_dtype, int split, int device, void* stream) {
  FNST_CHECK_ARG(raw_dtype == dtype || (raw_dtype == FNST_F32 && dtype == FNST_F16 && split),
                 "inorm_apply: raw/activation dtypes must match, except the fp32 -> split fp16 (hi|lo) form");
  FNST_CHECK_ARG(!split || (raw_dtype == FNST_F32 && dtype == FNST_F16), "inorm_apply: split output needs fp32 raw and fp16 activations");
  FNST_CHECK_ARG(raw && stats && gamma && beta && out, "inorm_apply: null pointer");
  FNST_CHECK_ARG(n > 0 && h > 0 && w > 0, "inorm_apply: empty tensor");
  FNST_CHECK_ARG(c % 8 == 0 && c <= 2048 && 256 % (c / 8) == 0, "inorm_apply: unsupported channel count %d", c);
  FNST_CHECK_ARG(pad >= 0 && (pad == 0 || pad_mode == FNST_PAD_REFLECT || pad_mode == FNST_PAD_ZERO), "inorm_apply: bad pad mode");
  if (pad_mode == FNST_PAD_REFLECT) FNST_CHECK_ARG(h > pad && w > pad, "inorm_apply: reflect pad %d needs h,w > pad (got %dx%d)", pad, h, w);
  FNST_CHECK_ARG(!(s2d && res), "inorm_apply: residual with space-to-depth output unsupported");
  FNST_DEVICE(device);
  // several output rows per block once the grid is large enough to fill the GPU (amortises the per-thread
  // scale/shift set-up); single rows for small problems
  const int hp = h + 2 * pad;
  int rows_per_block = 1;
  while (rows_per_block < 4 && (int64_t)n * ((hp + 2 * rows_per_block - 1) / (2 * rows_per_block)) >= 148 * 16) rows_per_block *= 2;
  dim3 grid((hp + rows_per_block - 1) / rows_per_block, n);
  if (split) {
    launch_pdl(inorm_apply_kernel<float, __half, 2, true>, dim3(grid), dim3(256), sizeof(float) * 2 * c, (cudaStream_t)stream, 
        reinterpret_cast<const float*>(raw), stats, gamma, beta, drop, reinterpret_cast<const __half*>(res), res_pad,
        reinterpret_cast<__half*>(out), reinterpret_cast<__nv_bfloat16*>(out_bf16), h, w, c, relu, eps, pad, pad_mode, s2d, rows_per_block);
    return launch_status("inorm_apply");
  }
  FNST_DISPATCH_DTYPE(dtype, T, {
    auto kern = inorm_apply_kernel<T, T, 2, false>;   // 2 pixels (x raw + residual) in flight per thread: measured best, 77-96 % of copy peak
    launch_pdl(kern, dim3(grid), dim3(256), sizeof(float) * 2 * c, (cudaStream_t)stream,
        reinterpret_cast<const T*>(raw), stats, gamma, beta, drop, reinterpret_cast<const T*>(res), res_pad,
        reinterpret_cast<T*>(out), reinterpret_cast<__nv_bfloat16*>(out_bf16), h, w, c, relu, eps, pad, pad_mode, s2d, rows_per_block);
  });
  return launch_status("inorm_apply");
}

extern "C" int fnst_maxpool2(const void* in, void* out, int n, int h, int w, int c, int dtype, int device, void* stream) {
  FNST_CHECK_ARG(in && out, "maxpool2: null pointer");
  FNST_CHECK_ARG(c % 8 == 0 && h >= 2 && w >= 2, "maxpool2: unsupported shape");
  FNST_DEVICE(device);
  const int64_t items = (int64_t)n * (h / 2) * (w / 2) * (c / 8);
  FNST_DISPATCH_DTYPE(dtype, T, {
    launch_pdl(maxpool2_kernel<T>, dim3(grid_for(items)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const T*>(in),
                                                                        reinterpret_cast<T*>(out), n, h, w, c);
  });
  return launch_status("maxpool2");
}

extern "C" int fnst_sse(const void* a, const void* b, int64_t count, int64_t b_period, int dtype_a, int dtype_b,
                        double* acc, int device, void* stream) {
  FNST_CHECK_ARG(a && b && acc && count > 0 && b_period > 0, "sse: bad arguments");
  FNST_DEVICE(device);
  FNST_DISPATCH_DTYPE(dtype_a, TA, {
    FNST_DISPATCH_DTYPE(dtype_b, TB, {
      launch_pdl(sse_kernel<TA, TB>, dim3(grid_for(count / 4)), dim3(256), 0, (cudaStream_t)stream, 
          reinterpret_cast<const TA*>(a), reinterpret_cast<const TB*>(b), count, b_period, acc);
    });
  });
  return launch_status("sse");
}

extern "C" int fnst_tv(const float* img, int planes, int h, int w, double* acc, int device, void* stream) {
  FNST_CHECK_ARG(img && acc && planes > 0 && h > 0 && w > 0, "tv: bad arguments");
  FNST_DEVICE(device);
  launch_pdl(tv_kernel, dim3(grid_for((int64_t)planes * h * w / 4)), dim3(256), 0, (cudaStream_t)stream, img, planes, h, w, acc);
  return launch_status("tv");
}

extern "C" int64_t fnst_loss_workspace_bytes(void) { return (int64_t)sizeof(double) * (LOSS_MAX_BLOCKS + 1); }

extern "C" int fnst_sse_scaled(const void* a, const void* b, int64_t count, int64_t b_period, int dtype_a, int dtype_b, float scale,
                               void* workspace, float* out, int accumulate, int device, void* stream) {
  FNST_CHECK_ARG(a && b && workspace && out && count > 0 && b_period > 0, "sse_scaled: bad arguments");
  FNST_DEVICE(device);
  FNST_DISPATCH_DTYPE(dtype_a, TA, {
    FNST_DISPATCH_DTYPE(dtype_b, TB, {
      launch_pdl(sse_scaled_kernel<TA, TB>, dim3(grid_for(count / 8)), dim3(256), 0, (cudaStream_t)stream,
          reinterpret_cast<const TA*>(a), reinterpret_cast<const TB*>(b), count, b_period, scale, reinterpret_cast<double*>(workspace),
          out, accumulate);
    });
  });
  return launch_status("sse_scaled");
}

extern "C" int fnst_tv_scaled(const float* img, int planes, int h, int w, float scale, void* workspace, float* out, int device,
                              void* stream) {
  FNST_CHECK_ARG(img && workspace && out && planes > 0 && h > 0 && w > 0, "tv_scaled: bad arguments");
  FNST_DEVICE(device);
  launch_pdl(tv_scaled_kernel, dim3(grid_for((int64_t)planes * h * w / 4)), dim3(256), 0, (cudaStream_t)stream, img, planes, h, w,
             scale, reinterpret_cast<double*>(workspace), out);
  return launch_status("tv_scaled");
}

extern "C" int fnst_nhwc_to_nchw(const void* in, float* out, int n, int h, int w, int c, int dtype, int device, void* stream) {
  FNST_CHECK_ARG(in && out && n > 0 && h > 0 && w > 0 && c > 0, "nhwc_to_nchw: bad arguments");
  FNST_DEVICE(device);
  dim3 grid((h * w + 31) / 32, (c + 31) / 32, n);
  FNST_DISPATCH_DTYPE(dtype, T, { launch_pdl(transpose_kernel<T, true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, in, out, h * w, c, c); });
  return launch_status("nhwc_to_nchw");
}

extern "C" int fnst_nchw_to_nhwc(const float* in, void* out, int n, int h, int w, int c, int c_pad, int dtype, int device, void* stream) {
  FNST_CHECK_ARG(in && out && n > 0 && h > 0 && w > 0 && c > 0 && c_pad >= c, "nchw_to_nhwc: bad arguments");
  FNST_DEVICE(device);
  dim3 grid((h * w + 31) / 32, (c_pad + 31) / 32, n);
  FNST_DISPATCH_DTYPE(dtype, T, { launch_pdl(transpose_kernel<T, false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, in, out, h * w, c, c_pad); });
  return launch_status("nchw_to_nhwc");
}

extern "C" int fnst_cast(const void* in, void* out, int64_t count, int in_dtype, int out_dtype, int device, void* stream) {
  FNST_CHECK_ARG(in && out && count > 0 && count % 8 == 0, "cast: bad arguments (count must be a multiple of 8)");
  FNST_DEVICE(device);
  FNST_DISPATCH_DTYPE(in_dtype, TI, {
    FNST_DISPATCH_DTYPE(out_dtype, TO, {
      launch_pdl(cast_kernel<TI, TO>, dim3(grid_for(count / 8)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const TI*>(in),
                                                                                  reinterpret_cast<TO*>(out), count / 8);
    });
  });
  return launch_status("cast");
}

extern "C" int fnst_gather_cast(const void* src, int src_dtype, const int32_t* idx, void* out, int out_dtype, int64_t count,
                                int device, void* stream) {
  FNST_CHECK_ARG(src && idx && out && count > 0, "gather_cast: bad arguments");
  FNST_DEVICE(device);
  FNST_DISPATCH_DTYPE(src_dtype, TI, {
    FNST_DISPATCH_DTYPE(out_dtype, TO, {
      launch_pdl(gather_cast_kernel<TI, TO>, dim3(grid_for(count)), dim3(256), 0, (cudaStream_t)stream,
                 reinterpret_cast<const TI*>(src), idx, reinterpret_cast<TO*>(out), count);
    });
  });
  return launch_status("gather_cast");
}

extern "C" int fnst_image_to_halo(const float* x, void* out, int n, int h, int w, int pad, int pad_mode, int c_pad, int rows,
                                  int pitch, int dtype, int split, int device, void* stream) {
  FNST_CHECK_ARG(!split || c_pad == 8, "image_to_halo: the hi|lo split form needs c_pad == 8");
  FNST_CHECK_ARG(x && out && n > 0 && h > 0 && w > 0, "image_to_halo: bad arguments");
  FNST_CHECK_ARG((c_pad == 4 || c_pad == 8) && (dtype == FNST_F16 || dtype == FNST_BF16), "image_to_halo: c_pad must be 4 or 8, dtype fp16/bf16");
  FNST_CHECK_ARG(rows >= h + 2 * pad && pitch >= w + 2 * pad, "image_to_halo: buffer smaller than the padded image");
  if (pad_mode == FNST_PAD_REFLECT) FNST_CHECK_ARG(h > pad && w > pad, "image_to_halo: reflect pad %d needs h,w > pad", pad);
  FNST_DEVICE(device);
  const int grid = grid_for((int64_t)n * rows * pitch);
  const int reflect = pad_mode == FNST_PAD_REFLECT;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FNST_F16) {
    if (c_pad == 4) launch_pdl(image_to_halo_kernel<__half, 4>, dim3(grid), dim3(256), 0, st, x, (__half*)out, n, h, w, pad, reflect, rows, pitch, split);
    else launch_pdl(image_to_halo_kernel<__half, 8>, dim3(grid), dim3(256), 0, st, x, (__half*)out, n, h, w, pad, reflect, rows, pitch, split);
  } else {
    if (c_pad == 4) launch_pdl(image_to_halo_kernel<__nv_bfloat16, 4>, dim3(grid), dim3(256), 0, st, x, (__nv_bfloat16*)out, n, h, w, pad, reflect, rows, pitch, split);
    else launch_pdl(image_to_halo_kernel<__nv_bfloat16, 8>, dim3(grid), dim3(256), 0, st, x, (__nv_bfloat16*)out, n, h, w, pad, reflect, rows, pitch, split);
  }
  return launch_status("image_to_halo");
}

extern "C" int fnst_u8_to_nchw(const void* in, float* out, int n, int h, int w, const float* mean3, const float* std3,
                               int device, void* stream) {
  FNST_CHECK_ARG(in && out && mean3 && std3 && n > 0 && h > 0 && w > 0, "u8_to_nchw: bad arguments");
  FNST_DEVICE(device);
  const int64_t pixels = (int64_t)n * h * w;
  launch_pdl(u8_to_nchw_kernel, dim3(grid_for(pixels)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const uint8_t*>(in), out, pixels, h * w,
                                                                         mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  return launch_status("u8_to_nchw");
}

extern "C" int fnst_nchw_to_u8(const float* in, void* out, int n, int h, int w, const float* mean3, const float* std3,
                               int device, void* stream) {
  FNST_CHECK_ARG(in && out && mean3 && std3 && n > 0 && h > 0 && w > 0, "nchw_to_u8: bad arguments");
  FNST_DEVICE(device);
  const int64_t pixels = (int64_t)n * h * w;
  launch_pdl(nchw_to_u8_kernel, dim3(grid_for(pixels)), dim3(256), 0, (cudaStream_t)stream, in, reinterpret_cast<uint8_t*>(out), pixels, h * w,
                                                                         mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  return launch_status("nchw_to_u8");
}
