// Gram matrices G[n] = F[n]^T F[n] of NHWC feature maps F[n] = [hw][c]  (losses/losses.py:6-13).
// CUDA-core version: 64x64 output tiles, split over pixel ranges, fp32 atomics into a zeroed output.
#include "common.cuh"

namespace fnst {

constexpr int GR_T = 64, GR_K = 16, GR_PIX = 1024;

template <typename T>
__global__ void __launch_bounds__(256) gram_simt_kernel(const T* __restrict__ feat, float* __restrict__ out, int HW, int C) {
  __shared__ float Xi[GR_K][GR_T + 4];
  __shared__ float Xj[GR_K][GR_T + 4];
  const int ksplit = (HW + GR_PIX - 1) / GR_PIX;
  const int n = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
  const int i0 = blockIdx.y * GR_T, j0 = blockIdx.x * GR_T;
  const int p_begin = ks * GR_PIX, p_end = min(HW, p_begin + GR_PIX);
  const T* X = feat + (size_t)n * HW * C;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 4, lc = (tid & 15) * 4;     // loader: 16 pixel rows x 64 channels
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int p0 = p_begin; p0 < p_end; p0 += GR_K) {
    const int p = p0 + lr;
    float vi[4] = {0.f, 0.f, 0.f, 0.f}, vj[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < p_end) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i0 + lc + k < C) vi[k] = to_f32<T>(X[(size_t)p * C + i0 + lc + k]);
        if (j0 + lc + k < C) vj[k] = to_f32<T>(X[(size_t)p * C + j0 + lc + k]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) { Xi[lr][lc + k] = vi[k]; Xj[lr][lc + k] = vj[k]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GR_K; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Xi[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Xj[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
  }
  float* G = out + (size_t)n * C * C;
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int i = i0 + ty * 4 + x, j = j0 + tx * 4 + y;
      if (i < C && j < C) atomicAdd(&G[(size_t)i * C + j], acc[x][y]);
    }
}

int gram_tc(const void* feat, float* out, int n, int hw, int c, int dtype, int device, cudaStream_t st);

}  // namespace fnst

using namespace fnst;

extern "C" int fnst_gram(const void* feat, float* out, int n, int hw, int c, int dtype, int use_tc, int prezeroed, int device,
                         void* stream) {
  FNST_CHECK_ARG(feat && out && n > 0 && hw > 0 && c > 0, "gram: bad arguments");
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!prezeroed) FNST_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n * c * c, st));
  if (use_tc) return gram_tc(feat, out, n, hw, c, dtype, device, st);
  const int ksplit = (hw + GR_PIX - 1) / GR_PIX;
  dim3 grid((c + GR_T - 1) / GR_T, (c + GR_T - 1) / GR_T, n * ksplit);
  FNST_DISPATCH_DTYPE(dtype, T, { gram_simt_kernel<T><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(feat), out, hw, c); });
  return launch_status("gram_simt");
}
