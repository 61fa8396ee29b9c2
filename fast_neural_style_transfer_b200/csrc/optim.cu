// Optimizer tail of the training step (train.py:203-205): clip_grad_norm_(max_norm) and Adam with coupled L2 weight
// decay, as three bandwidth-bound multi-tensor kernels over the 58 parameter tensors.
//
// The tensor lists are host arrays of device pointers; up to MT_MAX tensors travel by value in the kernel
// parameters (no device-side tables to keep in sync, capturable in a CUDA graph), longer lists are split over
// several launches.  A block owns one chunk of MT_CHUNK consecutive elements of ONE tensor (chunks never straddle
// tensors), found by a binary search over the prefix sum of chunk counts; 16-byte vector accesses when every pointer
// of the chunk is 16-byte aligned, scalar otherwise.
//
// Algorithmic bytes per parameter element: squared norm 4 (read g); scale 8 (read+write g);
// Adam 28 (read p, g, m, v; write p, m, v) -- 6 243 843 elements => 25 / 50 / 175 MB per step.
#include "common.cuh"

#include <cmath>

namespace fnst {

constexpr int MT_MAX = 64;        // tensors per launch
constexpr int MT_CHUNK = 4096;    // elements per block
constexpr int MT_THREADS = 256;

struct MTList {
  void* t[MT_MAX];
};
struct MTMeta {
  int numel[MT_MAX];
  int chunk_prefix[MT_MAX + 1];   // chunk_prefix[i] = number of chunks of tensors 0..i-1
  int n;
};

__device__ __forceinline__ int mt_find_tensor(const MTMeta& m, int chunk) {
  int lo = 0, hi = m.n;           // invariant: chunk_prefix[lo] <= chunk < chunk_prefix[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (m.chunk_prefix[mid] <= chunk) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ws: [0] double sum of squares, [1] unsigned block counter (both zero on entry of the first launch of a call; the
// finalising block resets them, so the workspace is zeroed once at allocation and never again).
struct SqnormWs {
  double acc;
  unsigned int done;
  unsigned int pad;
};

__global__ void __launch_bounds__(MT_THREADS) mt_sqnorm_kernel(MTList g, MTMeta meta, SqnormWs* ws, int finalize,
                                                                float max_norm, float* __restrict__ norm_coef) {
  pdl_trigger();
  pdl_wait();
  const int t = mt_find_tensor(meta, blockIdx.x);
  const int begin = (blockIdx.x - meta.chunk_prefix[t]) * MT_CHUNK;
  const int len = min(MT_CHUNK, meta.numel[t] - begin);
  const float* p = reinterpret_cast<const float*>(g.t[t]) + begin;
  float s = 0.f;
  if (aligned16(p)) {
    const int nv = len >> 2;
    for (int i = threadIdx.x; i < nv; i += MT_THREADS) {
      const float4 v = reinterpret_cast<const float4*>(p)[i];
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = (nv << 2) + threadIdx.x; i < len; i += MT_THREADS) s += p[i] * p[i];
  } else {
    for (int i = threadIdx.x; i < len; i += MT_THREADS) s += p[i] * p[i];
  }
  __shared__ float part[MT_THREADS / 32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < MT_THREADS / 32; ++i) tot += (double)part[i];
    atomicAdd(&ws->acc, tot);
    __threadfence();
    const unsigned int prev = atomicAdd(&ws->done, 1u);
    if (finalize && prev == (unsigned int)finalize - 1u) {
      // last block of the whole call (finalize = total number of blocks over all launches of this call)
      __threadfence();
      const double sq = *reinterpret_cast<volatile double*>(&ws->acc);
      const float norm = (float)sqrt(sq);
      // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to <= 1
      float coef = max_norm / (norm + 1e-6f);
      coef = coef > 1.f ? 1.f : coef;
      norm_coef[0] = norm;
      norm_coef[1] = coef;
      ws->acc = 0.0;
      ws->done = 0u;
    }
  }
}

__global__ void __launch_bounds__(MT_THREADS) mt_scale_kernel(MTList g, MTMeta meta, const float* __restrict__ coef_ptr) {
  pdl_trigger();
  pdl_wait();
  const float coef = *coef_ptr;
  const int t = mt_find_tensor(meta, blockIdx.x);
  const int begin = (blockIdx.x - meta.chunk_prefix[t]) * MT_CHUNK;
  const int len = min(MT_CHUNK, meta.numel[t] - begin);
  float* p = reinterpret_cast<float*>(g.t[t]) + begin;
  int tail = 0;
  if (aligned16(p)) {
    const int nv = len >> 2;
    for (int i = threadIdx.x; i < nv; i += MT_THREADS) {
      float4 v = reinterpret_cast<float4*>(p)[i];
      v.x *= coef; v.y *= coef; v.z *= coef; v.w *= coef;
      reinterpret_cast<float4*>(p)[i] = v;
    }
    tail = nv << 2;
  }
  for (int i = tail + threadIdx.x; i < len; i += MT_THREADS) p[i] *= coef;
}

struct AdamHyper {
  float lr_over_bc1;      // lr / (1 - beta1^step)
  float one_minus_beta1;
  float beta2;
  float one_minus_beta2;
  float bc2_sqrt;         // sqrt(1 - beta2^step)
  float eps;
  float weight_decay;
};

// torch.optim.Adam (amsgrad=False, maximize=False), the arithmetic of torch/optim/adam.py::_single_tensor_adam:
//   g   = g + wd * p
//   m   = m + (1 - beta1) * (g - m)                       (lerp)
//   v   = v * beta2 + (1 - beta2) * g * g
//   p   = p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamHyper& h, float gscale) {
  g *= gscale;
  g = g + h.weight_decay * p;
  m = m + h.one_minus_beta1 * (g - m);
  v = v * h.beta2 + h.one_minus_beta2 * g * g;
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p = p - h.lr_over_bc1 * (m / denom);
}

__global__ void __launch_bounds__(MT_THREADS) mt_adam_kernel(MTList P, MTList G, MTList M, MTList V, MTMeta meta, AdamHyper h,
                                                              const float* __restrict__ gscale_ptr) {
  pdl_trigger();
  pdl_wait();
  const float gscale = gscale_ptr ? *gscale_ptr : 1.f;
  const int t = mt_find_tensor(meta, blockIdx.x);
  const int begin = (blockIdx.x - meta.chunk_prefix[t]) * MT_CHUNK;
  const int len = min(MT_CHUNK, meta.numel[t] - begin);
  float* p = reinterpret_cast<float*>(P.t[t]) + begin;
  const float* g = reinterpret_cast<const float*>(G.t[t]) + begin;
  float* m = reinterpret_cast<float*>(M.t[t]) + begin;
  float* v = reinterpret_cast<float*>(V.t[t]) + begin;
  int tail = 0;
  if (aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v)) {
    const int nv = len >> 2;
    for (int i = threadIdx.x; i < nv; i += MT_THREADS) {
      float4 pv = reinterpret_cast<float4*>(p)[i];
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 mv = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      adam_update(pv.x, gv.x, mv.x, vv.x, h, gscale);
      adam_update(pv.y, gv.y, mv.y, vv.y, h, gscale);
      adam_update(pv.z, gv.z, mv.z, vv.z, h, gscale);
      adam_update(pv.w, gv.w, mv.w, vv.w, h, gscale);
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    tail = nv << 2;
  }
  for (int i = tail + threadIdx.x; i < len; i += MT_THREADS) {
    float pv = p[i], mv = m[i], vv = v[i];
    adam_update(pv, g[i], mv, vv, h, gscale);
    p[i] = pv; m[i] = mv; v[i] = vv;
  }
}

// Fill `meta` for tensors [first, first+count) of the list; returns the number of chunks (= blocks) or -1.
static int mt_fill_meta(const int64_t* numels, int first, int count, MTMeta& meta) {
  meta.n = count;
  meta.chunk_prefix[0] = 0;
  for (int i = 0; i < count; ++i) {
    const int64_t ne = numels[first + i];
    if (ne <= 0 || ne > 0x7fffffff) return -1;
    meta.numel[i] = (int)ne;
    meta.chunk_prefix[i + 1] = meta.chunk_prefix[i] + (int)((ne + MT_CHUNK - 1) / MT_CHUNK);
  }
  for (int i = count; i < MT_MAX; ++i) { meta.numel[i] = 0; meta.chunk_prefix[i + 1] = meta.chunk_prefix[count]; }
  return meta.chunk_prefix[count];
}

static void mt_fill_list(void* const* ptrs, int first, int count, MTList& l) {
  for (int i = 0; i < MT_MAX; ++i) l.t[i] = i < count ? ptrs[first + i] : nullptr;
}

static int mt_check(void* const* list, const int64_t* numels, int n, const char* what) {
  FNST_CHECK_ARG(list && numels && n > 0, "%s: empty tensor list", what);
  for (int i = 0; i < n; ++i) {
    FNST_CHECK_ARG(list[i] != nullptr, "%s: tensor %d is NULL", what, i);
    FNST_CHECK_ARG((reinterpret_cast<uintptr_t>(list[i]) & 3u) == 0, "%s: tensor %d is not 4-byte aligned", what, i);
    FNST_CHECK_ARG(numels[i] > 0 && numels[i] <= 0x7fffffff, "%s: tensor %d has %lld elements (1 .. 2^31-1 supported)", what, i,
                   (long long)numels[i]);
  }
  return 0;
}

}  // namespace fnst

using namespace fnst;

extern "C" int64_t fnst_grad_norm_workspace_bytes(void) { return (int64_t)sizeof(SqnormWs); }

extern "C" int fnst_grad_norm(void* const* grads, const int64_t* numels, int n, void* workspace, float max_norm,
                              float* norm_coef, int device, void* stream) {
  if (int rc = mt_check(grads, numels, n, "grad_norm")) return rc;
  FNST_CHECK_ARG(workspace && norm_coef, "grad_norm: workspace / output is NULL");
  FNST_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "grad_norm: workspace must be 8-byte aligned");
  FNST_CHECK_ARG(max_norm > 0.f, "grad_norm: max_norm must be positive");
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total_blocks = 0;
  for (int i = 0; i < n; ++i) total_blocks += (numels[i] + MT_CHUNK - 1) / MT_CHUNK;
  FNST_CHECK_ARG(total_blocks <= 0x7fffffff, "grad_norm: too many elements");
  for (int first = 0; first < n; first += MT_MAX) {
    const int count = n - first < MT_MAX ? n - first : MT_MAX;
    MTMeta meta; MTList gl;
    const int blocks = mt_fill_meta(numels, first, count, meta);
    mt_fill_list(grads, first, count, gl);
    // every launch passes the call's total block count: the block that brings the counter to it finalises
    launch_pdl(mt_sqnorm_kernel, dim3(blocks), dim3(MT_THREADS), 0, st, gl, meta, reinterpret_cast<SqnormWs*>(workspace),
               (int)total_blocks, max_norm, norm_coef);
    if (int rc = launch_status("grad_norm")) return rc;
  }
  return 0;
}

extern "C" int fnst_grad_scale(void* const* grads, const int64_t* numels, int n, const float* coef, int device, void* stream) {
  if (int rc = mt_check(grads, numels, n, "grad_scale")) return rc;
  FNST_CHECK_ARG(coef, "grad_scale: coef is NULL");
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  for (int first = 0; first < n; first += MT_MAX) {
    const int count = n - first < MT_MAX ? n - first : MT_MAX;
    MTMeta meta; MTList gl;
    const int blocks = mt_fill_meta(numels, first, count, meta);
    mt_fill_list(grads, first, count, gl);
    launch_pdl(mt_scale_kernel, dim3(blocks), dim3(MT_THREADS), 0, st, gl, meta, coef);
    if (int rc = launch_status("grad_scale")) return rc;
  }
  return 0;
}

extern "C" int fnst_adam_step(void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                              const int64_t* numels, int n, double lr, double beta1, double beta2, double eps, double weight_decay,
                              int64_t step, const float* grad_scale, int device, void* stream) {
  if (int rc = mt_check(params, numels, n, "adam_step(params)")) return rc;
  if (int rc = mt_check(grads, numels, n, "adam_step(grads)")) return rc;
  if (int rc = mt_check(exp_avg, numels, n, "adam_step(exp_avg)")) return rc;
  if (int rc = mt_check(exp_avg_sq, numels, n, "adam_step(exp_avg_sq)")) return rc;
  FNST_CHECK_ARG(step >= 1, "adam_step: step counts from 1 (got %lld)", (long long)step);
  FNST_CHECK_ARG(lr >= 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0 && weight_decay >= 0.0,
                 "adam_step: bad hyper-parameters");
  // torch's lerp_ switches formula at weight >= 0.5; only the small-weight form is implemented
  FNST_CHECK_ARG(1.0 - beta1 < 0.5, "adam_step: beta1 <= 0.5 is not supported");
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  // bias corrections in double, like the Python scalars of torch/optim/adam.py
  const double bc1 = 1.0 - std::pow(beta1, (double)step);
  const double bc2 = 1.0 - std::pow(beta2, (double)step);
  AdamHyper h;
  h.lr_over_bc1 = (float)(lr / bc1);
  h.one_minus_beta1 = (float)(1.0 - beta1);
  h.beta2 = (float)beta2;
  h.one_minus_beta2 = (float)(1.0 - beta2);
  h.bc2_sqrt = (float)std::sqrt(bc2);
  h.eps = (float)eps;
  h.weight_decay = (float)weight_decay;
  for (int first = 0; first < n; first += MT_MAX) {
    const int count = n - first < MT_MAX ? n - first : MT_MAX;
    MTMeta meta; MTList pl, gl, ml, vl;
    const int blocks = mt_fill_meta(numels, first, count, meta);
    mt_fill_list(params, first, count, pl);
    mt_fill_list(grads, first, count, gl);
    mt_fill_list(exp_avg, first, count, ml);
    mt_fill_list(exp_avg_sq, first, count, vl);
    launch_pdl(mt_adam_kernel, dim3(blocks), dim3(MT_THREADS), 0, st, pl, gl, ml, vl, meta, h, grad_scale);
    if (int rc = launch_status("adam_step")) return rc;
  }
  return 0;
}
