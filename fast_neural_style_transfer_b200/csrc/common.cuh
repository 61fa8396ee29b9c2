// Shared helpers for libfnst: error reporting, dtype dispatch, small device utilities.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <utility>

#include "../../include/fnst.h"

namespace fnst {

void set_error(const char* fmt, ...);

#define FNST_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::fnst::set_error(__VA_ARGS__);             \
      return -1;                                  \
    }                                             \
  } while (0)

#define FNST_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::fnst::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (int)e__;                                                                    \
    }                                                                                     \
  } while (0)

// Every entry point runs on the device the caller names (it may be called from autograd's backward thread, whose current
// device is not the forward thread's); the caller's current device is restored when the entry point returns.
struct DeviceGuard {
  int prev = -1;
  cudaError_t enter(int device) {
    int cur = -1;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return e;
    if (cur == device) return cudaSuccess;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) prev = cur;
    return e;
  }
  ~DeviceGuard() { if (prev >= 0) (void)cudaSetDevice(prev); }
};
#define FNST_DEVICE(device)              \
  ::fnst::DeviceGuard fnst_device_guard__; \
  FNST_CUDA(fnst_device_guard__.enter(device))

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------
// Every kernel of the tensor-core path starts with pdl_trigger() (lets the next kernel on the stream be
// scheduled as soon as all CTAs of this grid are resident) and calls pdl_wait() before its first access to
// global memory (blocks until the preceding grid has completed and its writes are visible).  The set-up in
// between -- barrier init, TMEM allocation, tensor-map prefetch, index math -- overlaps the tail of the
// predecessor.  Under stream capture the attribute becomes a programmatic edge of the CUDA graph.  A kernel
// launched through launch_pdl MUST execute pdl_wait() in every CTA, otherwise ordering is not transitive.
// FNST_PDL=0 disables the launch attribute (both device calls are then no-ops).
bool pdl_enabled();

// Tuning knobs (defaults from measurement; fnst_set_tuning / FNST_* environment variables override them).
struct Tuning {
  int conv_block_n = 0;        // 0 = heuristic; else force the column tile of conv_tc (64/128/256)
  int wgrad_waves_x2 = 2;      // wgrad_tc split-K target: tasks <= waves_x2/2 * SM count (one wave measured best: fewer L2 atomics)
  int wgrad_bn = 0;            // 0 = widest column block that divides kc; else force 64/128/256
  int pdl = 1;
  int conv_stage_out = 1;      // conv_tc epilogue: 1 = shared-memory tile + TMA store + per-column statistics where applicable, 0 = never,
                               // 2 = one-tile-per-CTA launches only (A/B of the private staging tile of the 64 / 128 column tiles)
  int conv_pair = 1;           // 1: CTA pairs (cta_group::2) where measured faster; 0: never; 2: whenever the column tile is >= 128
  int conv_rowstream = 1;      // 1: 3x3 convolutions over 64 input channels (VGG conv1_2 / conv2_1 and conv1_2's data gradient) use the row-streaming kernel
  int inorm_bwd_tma = 0;       // 1: InstanceNorm backward pass 1 stages each image row with TMA box loads where the geometry allows
                               // (alone 11.5 vs 13.5 us at batch 4; inside the training step +0.05 ms: its ~100 KB CTAs cannot share
                               // SMs with the weight-gradient GEMMs of the side stream the way the register form does -> default off)
  int inorm_bwd_blocks = 2;    // resident blocks per SM the InstanceNorm-backward reduce kernel is compiled for (16-bit types): 1 or 2
  int resize_staged = 0;       // fnst_resize_to_tensor: 1 = stage the tile's input span in shared memory with 32-bit loads
  int dbg_mode = 0;            // measurement only: bit 0 skips the per-chunk column sums, bit 1 the global statistics atomics
  unsigned long long* debug_buf = nullptr;   // measurement only: conv_tc writes per-CTA main-loop clocks / nanoseconds here
};
Tuning& tuning();

// Where griddepcontrol.launch_dependents is issued -- measured on B200 (profiles/r01_pdl_modes.md): an explicit early
// trigger makes the dependent grid resident while its predecessor still runs, and those CTAs parked in
// griddepcontrol.wait slow a small running grid down (batch-1 inference: -7 % with a trigger at kernel entry or after
// the wait, -8 % with a trigger after the MMA main loop), while the default below -- no explicit trigger, i.e. the
// implicit one at CTA exit; the dependent still skips the full launch latency and overlaps its set-up with the
// predecessor's memory flush -- is +4 % at batch 1 and neutral elsewhere.  Modes 0 (entry), 1 (after the wait) and
// 3 (after the MMA main loop of the tensor-core kernels) are kept for the A/B experiment.
#ifndef FNST_PDL_MODE
#define FNST_PDL_MODE 2
#endif
__device__ __forceinline__ void pdl_trigger_tail() {
#if FNST_PDL_MODE == 3
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#if FNST_PDL_MODE == 0
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#if FNST_PDL_MODE == 1
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// cluster_x > 1: launch as thread-block clusters of cluster_x CTAs along x (grid.x must be a multiple of it).
template <typename... KArgs, typename... Args>
inline void launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                               Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster_x; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  (void)cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);   // errors surface through launch_status()
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_pdl_cluster(kern, grid, block, smem, st, 1, std::forward<Args>(args)...);
}

inline size_t dtype_size(int dt) { return dt == FNST_F32 ? 4 : 2; }

// ---- scalar conversion helpers -------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements <-> 8 floats (16-byte vectors for 2-byte types, 2x16 for fp32).
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__half>(__half* p, const float (&v)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// Raw (unconverted) 8-element vectors: keep loads in flight in few registers, convert at the point of use.
template <typename T> struct Raw8 { uint4 u; };
template <> struct Raw8<float> { float4 a, b; };
template <typename T> __device__ __forceinline__ Raw8<T> load_raw8(const T* p) {
  Raw8<T> r; r.u = *reinterpret_cast<const uint4*>(p); return r;
}
template <> __device__ __forceinline__ Raw8<float> load_raw8<float>(const float* p) {
  Raw8<float> r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r;
}
template <typename T> __device__ __forceinline__ void raw8_to_f32(const Raw8<T>& r, float (&v)[8]);
template <> __device__ __forceinline__ void raw8_to_f32<float>(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
template <> __device__ __forceinline__ void raw8_to_f32<__half>(const Raw8<__half>& r, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void raw8_to_f32<__nv_bfloat16>(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// Runtime-typed scalar / 8-vector loads (epilogue extras whose dtype differs from the kernel's template types).
__device__ __forceinline__ float load_scalar_f32(const void* base, int dtype, size_t idx) {
  if (dtype == FNST_F32) return reinterpret_cast<const float*>(base)[idx];
  if (dtype == FNST_F16) return __half2float(reinterpret_cast<const __half*>(base)[idx]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}
__device__ __forceinline__ void load8_dyn(const void* base, int dtype, size_t idx, float (&v)[8]) {
  if (dtype == FNST_F32) load8<float>(reinterpret_cast<const float*>(base) + idx, v);
  else if (dtype == FNST_F16) load8<__half>(reinterpret_cast<const __half*>(base) + idx, v);
  else load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(base) + idx, v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ReflectionPad2d index law: -i -> i, (n-1)+i -> (n-1)-i (edge not repeated).
__host__ __device__ __forceinline__ int reflect_index(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// Dispatch a generic lambda on an element type tag.
#define FNST_DISPATCH_DTYPE(dt, T, ...)                                     \
  switch (dt) {                                                             \
    case FNST_F32: { using T = float; __VA_ARGS__; break; }                 \
    case FNST_F16: { using T = __half; __VA_ARGS__; break; }                \
    case FNST_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }        \
    default: ::fnst::set_error("bad dtype %d", (int)(dt)); return -1;       \
  }

}  // namespace fnst
