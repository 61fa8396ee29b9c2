// Input pre-processing of the reference's data pipeline on the GPU (device side):
//
//     transforms.Resize((256, 256)) -> transforms.ToTensor() -> transforms.Normalize(mean, std)
//     (train.py:92-102, inference.py:28-31, applied per image by data/dataset.py:21-27)
//
// On a PIL image torchvision's Resize is Pillow's `Image.resize(size, BILINEAR)`: a two-pass (horizontal, then vertical)
// 8-bit resampling with a triangle filter widened by the down-scaling factor, 22-bit fixed-point weights, round-half-up
// and clipping after EACH pass.  This kernel reproduces it bit for bit (tests: against Pillow itself):
//
//   * the per-axis coefficient windows are computed on the device in double precision with explicitly rounded
//     operations (__dmul_rn / __dadd_rn / __ddiv_rn: no FMA contraction), i.e. the same IEEE operations in the same order
//     as the C library executes on the host -- no coefficient tables to upload, no workspace;
//   * one block = one 16x16 tile of the output: warp 0 / warp 1 build the 16 horizontal / 16 vertical windows in shared
//     memory, all threads run the horizontal pass for exactly the input rows the tile's vertical windows touch into a
//     uint8 shared-memory strip (each input byte is fetched once per tile column group, the intermediate never goes to
//     HBM), then each thread finishes one output pixel with the vertical pass and writes uint8 HWC and / or the
//     normalised float CHW tensor (ToTensor's /255 and Normalize's (x-mean)/std as IEEE float32 operations in that order).
//
// Algorithmic bytes per image: read in_h*in_w*3 once, write out_h*out_w*3*4 (float) -- HBM-bound by the input image.
#include "common.cuh"

#include <atomic>

namespace fnst {

constexpr int RS_TILE = 16;
constexpr int RS_KMAX = 72;                 // taps per window: ceil(scale)*2+1 <= 72, i.e. down-scaling factors up to 35
constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

// IEEE double operations that neither nvcc nor the host compiler may contract into FMAs
#ifdef __CUDA_ARCH__
#define RS_MUL(a, b) __dmul_rn((a), (b))
#define RS_ADD(a, b) __dadd_rn((a), (b))
#define RS_SUB(a, b) __dadd_rn((a), -(b))
#define RS_DIV(a, b) __ddiv_rn((a), (b))
#else
#define RS_MUL(a, b) ((a) * (b))
#define RS_ADD(a, b) ((a) + (b))
#define RS_SUB(a, b) ((a) - (b))
#define RS_DIV(a, b) ((a) / (b))
#endif

struct AxisScale {
  double scale, filterscale, support, ss;
  int ksize;
};

__host__ __device__ inline AxisScale axis_scale(int in_size, int out_size) {
  AxisScale a;
  a.scale = RS_DIV((double)in_size, (double)out_size);
  a.filterscale = a.scale < 1.0 ? 1.0 : a.scale;
  a.support = a.filterscale;                                   // bilinear: support 1.0 * filterscale
  a.ss = RS_DIV(1.0, a.filterscale);
  a.ksize = (int)ceil(a.support) * 2 + 1;
  return a;
}

__host__ __device__ inline double triangle(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? RS_SUB(1.0, x) : 0.0;
}

// Window of output index xx: first input index, length, and fixed-point weights kk[0..len) (Pillow's precompute_coeffs +
// normalize_coeffs_8bpc for one row).  The weights are evaluated twice (sum, then normalise) instead of being stored.
__host__ __device__ inline void axis_window(const AxisScale& a, int in_size, int xx, int* first, int* len, int* kk) {
  const double center = RS_ADD(0.0, RS_MUL((double)xx + 0.5, a.scale));
  int xmin = (int)RS_ADD(RS_SUB(center, a.support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)RS_ADD(RS_ADD(center, a.support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) ww = RS_ADD(ww, triangle(RS_MUL(RS_ADD(RS_SUB((double)(x + xmin), center), 0.5), a.ss)));
  for (int x = 0; x < xmax; ++x) {
    double w = triangle(RS_MUL(RS_ADD(RS_SUB((double)(x + xmin), center), 0.5), a.ss));
    if (ww != 0.0) w = RS_DIV(w, ww);
    kk[x] = (int)RS_ADD(0.5, RS_MUL(w, (double)(1 << RS_PRECISION_BITS)));      // bilinear weights are never negative
  }
  *first = xmin;
  *len = xmax;
}

__host__ __device__ inline int clip8(int v) {
  v >>= RS_PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

struct ResizeParams {
  const uint8_t* in;
  int in_h, in_w;
  int64_t in_pitch;
  int out_h, out_w;
  float* out_f;            // [3][out_h][out_w] or null
  uint8_t* out_u8;         // [out_h][out_w][3] or null
  float mean[3], std[3];
  int normalize;
  int strip_rows;          // capacity of the shared-memory strip (rows)
  int stage_pitch;         // STAGED: bytes per staged input row (multiple of 4) and rows per staging chunk
  int stage_rows;
  const fnst_image_desc* batch;   // non-null: blockIdx.z selects the image (in / in_h / in_w / in_pitch come from the table,
                                  // the outputs are offset by z images)
};

// STAGED: the input span of the tile is first copied to shared memory with coalesced 32-bit loads (chunks of stage_rows rows),
// and the horizontal taps read shared memory; otherwise every tap is a byte load from global memory through L1.
// Measured on B200 (profiles/r01_resize_selftest.log, 1080x1920 -> 256x256, both bit-exact, 64 single-image launches back to
// back): byte loads 13.7 us per image, staged 26.3 us.  Caveat: those are per-LAUNCH times of one-image kernels and include
// the host's launch cost -- in that run the staged form also paid a cudaFuncSetAttribute call per launch (since made
// once per device) -- so they bound the kernel times from above rather than rank the two forms; until both are re-measured
// from a CUDA graph the verified un-staged form stays the default (tuning knob resize_staged = 1 selects this one).
template <bool STAGED>
__global__ void __launch_bounds__(RS_TILE * RS_TILE) resize_to_tensor_kernel(const ResizeParams p_in) {
  pdl_trigger();
  pdl_wait();
  ResizeParams p = p_in;
  if (p.batch) {
    // many images per launch (fnst_resize_batch_to_tensor): the grid's z index walks a device table of image descriptors
    const fnst_image_desc d = p.batch[blockIdx.z];
    p.in = reinterpret_cast<const uint8_t*>(d.data);
    p.in_h = d.h; p.in_w = d.w; p.in_pitch = d.pitch_bytes;
    const int64_t px = (int64_t)p.out_h * p.out_w * 3;
    if (p.out_f) p.out_f += (int64_t)blockIdx.z * px;
    if (p.out_u8) p.out_u8 += (int64_t)blockIdx.z * px;
  }
  __shared__ int kh[RS_TILE][RS_KMAX], kv[RS_TILE][RS_KMAX];
  __shared__ int bh[RS_TILE][2], bv[RS_TILE][2];
  extern __shared__ uint8_t strip[];                 // [rows][RS_TILE][3] horizontal-pass results of this tile's columns
  const int tid = threadIdx.x;
  const int X0 = blockIdx.x * RS_TILE, Y0 = blockIdx.y * RS_TILE;
  const bool need_h = p.out_w != p.in_w, need_v = p.out_h != p.in_h;

  if (tid < RS_TILE) {
    const int X = X0 + tid;
    if (X < p.out_w) {
      if (need_h) {
        const AxisScale a = axis_scale(p.in_w, p.out_w);
        axis_window(a, p.in_w, X, &bh[tid][0], &bh[tid][1], kh[tid]);
      } else { bh[tid][0] = X; bh[tid][1] = 1; }
    } else { bh[tid][0] = 0; bh[tid][1] = 0; }
  } else if (tid >= 32 && tid < 32 + RS_TILE) {
    const int l = tid - 32, Y = Y0 + l;
    if (Y < p.out_h) {
      if (need_v) {
        const AxisScale a = axis_scale(p.in_h, p.out_h);
        axis_window(a, p.in_h, Y, &bv[l][0], &bv[l][1], kv[l]);
      } else { bv[l][0] = Y; bv[l][1] = 1; }
    } else { bv[l][0] = 0x7fffffff; bv[l][1] = 0; }
  }
  __syncthreads();
  // input rows this tile reads: windows start monotonically, so the first valid lane has the smallest start
  const int r0 = bv[0][0];
  int r1 = r0;
#pragma unroll
  for (int l = 0; l < RS_TILE; ++l)
    if (bv[l][1] > 0) r1 = max(r1, bv[l][0] + bv[l][1]);
  const int rows = r1 - r0;                          // <= p.strip_rows by construction of the launch

  if (STAGED && need_h) {
    const int c0 = bh[0][0];                         // first input column of the tile (windows start monotonically)
    int c1 = c0;
#pragma unroll
    for (int l = 0; l < RS_TILE; ++l)
      if (bh[l][1] > 0) c1 = max(c1, bh[l][0] + bh[l][1]);
    const int span = (c1 - c0) * 3;                  // bytes per input row
    const int wpr = (span + 6) >> 2;                 // 32-bit words per staged row (row start up to 3 bytes past alignment)
    uint8_t* stage = strip + ((p.strip_rows * RS_TILE * 3 + 15) & ~15);
    for (int rbase = 0; rbase < rows; rbase += p.stage_rows) {
      const int nr = min(p.stage_rows, rows - rbase);
      for (int idx = tid; idx < nr * wpr; idx += RS_TILE * RS_TILE) {
        const int r = idx / wpr, wi = idx - r * wpr;
        const uint8_t* rowp = p.in + (int64_t)(r0 + rbase + r) * p.in_pitch + (int64_t)c0 * 3;
        const int mis = (int)(reinterpret_cast<uintptr_t>(rowp) & 3);
        const int total = mis + span;                // bytes counted from the aligned address below rowp
        if (4 * wi >= total) continue;
        const uint8_t* gp = rowp - mis + 4 * wi;
        uint8_t* sp = stage + r * p.stage_pitch + 4 * wi;
        const int lo = wi == 0 ? mis : 0, hi = min(4, total - 4 * wi);
        if (lo == 0 && hi == 4) {
          *reinterpret_cast<uint32_t*>(sp) = *reinterpret_cast<const uint32_t*>(gp);
        } else {                                     // partial first / last word: never touch bytes outside the row span
          for (int b = lo; b < hi; ++b) sp[b] = gp[b];
        }
      }
      __syncthreads();
      for (int idx = tid; idx < nr * RS_TILE; idx += RS_TILE * RS_TILE) {
        const int r = idx / RS_TILE, cx = idx % RS_TILE;
        const int len = bh[cx][1];
        if (len == 0) continue;
        const uint8_t* rowp = p.in + (int64_t)(r0 + rbase + r) * p.in_pitch + (int64_t)c0 * 3;
        const int mis = (int)(reinterpret_cast<uintptr_t>(rowp) & 3);
        const uint8_t* src = stage + r * p.stage_pitch + mis + (bh[cx][0] - c0) * 3;
        int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
        const int* k = kh[cx];
        for (int x = 0; x < len; ++x) {
          const int w = k[x];
          s0 += src[3 * x + 0] * w; s1 += src[3 * x + 1] * w; s2 += src[3 * x + 2] * w;
        }
        uint8_t* dst = strip + ((rbase + r) * RS_TILE + cx) * 3;
        dst[0] = (uint8_t)clip8(s0); dst[1] = (uint8_t)clip8(s1); dst[2] = (uint8_t)clip8(s2);
      }
      __syncthreads();
    }
  } else
  // horizontal pass: (row, column) pairs of the strip
  for (int idx = tid; idx < rows * RS_TILE; idx += RS_TILE * RS_TILE) {
    const int r = idx / RS_TILE, cx = idx % RS_TILE;
    const int len = bh[cx][1];
    if (len == 0) continue;
    const uint8_t* src = p.in + (int64_t)(r0 + r) * p.in_pitch + (int64_t)bh[cx][0] * 3;
    uint8_t* dst = strip + (r * RS_TILE + cx) * 3;
    if (need_h) {
      int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
      const int* k = kh[cx];
      for (int x = 0; x < len; ++x) {
        const int w = k[x];
        s0 += src[3 * x + 0] * w; s1 += src[3 * x + 1] * w; s2 += src[3 * x + 2] * w;
      }
      dst[0] = (uint8_t)clip8(s0); dst[1] = (uint8_t)clip8(s1); dst[2] = (uint8_t)clip8(s2);
    } else {
      dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
  }
  __syncthreads();

  // vertical pass: one output pixel per thread
  const int tx = tid % RS_TILE, ty = tid / RS_TILE;
  const int X = X0 + tx, Y = Y0 + ty;
  if (X >= p.out_w || Y >= p.out_h) return;
  int o[3];
  if (need_v) {
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    const int len = bv[ty][1];
    const uint8_t* col = strip + ((bv[ty][0] - r0) * RS_TILE + tx) * 3;
    const int* k = kv[ty];
    for (int y = 0; y < len; ++y) {
      const int w = k[y];
      s0 += col[y * RS_TILE * 3 + 0] * w; s1 += col[y * RS_TILE * 3 + 1] * w; s2 += col[y * RS_TILE * 3 + 2] * w;
    }
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
  } else {
    const uint8_t* px = strip + ((Y - r0) * RS_TILE + tx) * 3;
    o[0] = px[0]; o[1] = px[1]; o[2] = px[2];
  }
  if (p.out_u8) {
    uint8_t* q = p.out_u8 + ((int64_t)Y * p.out_w + X) * 3;
    q[0] = (uint8_t)o[0]; q[1] = (uint8_t)o[1]; q[2] = (uint8_t)o[2];
  }
  if (p.out_f) {
    const int64_t plane = (int64_t)p.out_h * p.out_w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __fdiv_rn((float)o[c], 255.0f);                                  // ToTensor
      if (p.normalize) v = __fdiv_rn(__fsub_rn(v, p.mean[c]), p.std[c]);         // Normalize: (x - mean) / std
      p.out_f[c * plane + (int64_t)Y * p.out_w + X] = v;
    }
  }
}

// Rows of input one tile of RS_TILE output rows can touch (upper bound, host side).
static int strip_rows_bound(int in_h, int out_h) {
  if (in_h == out_h) return RS_TILE;
  const AxisScale a = axis_scale(in_h, out_h);
  const double span = (RS_TILE - 1) * a.scale + 2.0 * a.support + 4.0;      // tests enumerate the exact need: <= 15*scale + 2*support + 3
  int rows = (int)ceil(span);
  return rows > in_h ? in_h : rows;
}

}  // namespace fnst

using namespace fnst;

extern "C" int fnst_resize_to_tensor(const void* img_hwc, int in_h, int in_w, int64_t in_pitch_bytes, int out_h, int out_w,
                                     float* out_chw, void* out_u8_hwc, const float* mean3, const float* std3, int device,
                                     void* stream) {
  FNST_CHECK_ARG(img_hwc && (out_chw || out_u8_hwc), "resize_to_tensor: null pointer");
  FNST_CHECK_ARG(in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "resize_to_tensor: bad sizes %dx%d -> %dx%d", in_h, in_w, out_h, out_w);
  FNST_CHECK_ARG(in_pitch_bytes >= (int64_t)in_w * 3, "resize_to_tensor: pitch %lld < 3*width", (long long)in_pitch_bytes);
  FNST_CHECK_ARG((mean3 == nullptr) == (std3 == nullptr), "resize_to_tensor: pass both mean and std, or neither");
  const AxisScale ah = axis_scale(in_w, out_w), av = axis_scale(in_h, out_h);
  FNST_CHECK_ARG(ah.ksize <= RS_KMAX && av.ksize <= RS_KMAX,
                 "resize_to_tensor: down-scaling factor above %d is not supported (%dx%d -> %dx%d)", (RS_KMAX - 1) / 2, in_h, in_w, out_h, out_w);
  ResizeParams p;
  p.in = reinterpret_cast<const uint8_t*>(img_hwc);
  p.in_h = in_h; p.in_w = in_w; p.in_pitch = in_pitch_bytes; p.out_h = out_h; p.out_w = out_w;
  p.out_f = out_chw; p.out_u8 = reinterpret_cast<uint8_t*>(out_u8_hwc);
  p.normalize = mean3 != nullptr;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3 ? mean3[c] : 0.f; p.std[c] = std3 ? std3[c] : 1.f; }
  p.strip_rows = strip_rows_bound(in_h, out_h);
  const bool staged = tuning().resize_staged != 0 && in_w != out_w;
  size_t smem = (size_t)p.strip_rows * RS_TILE * 3;
  p.stage_pitch = p.stage_rows = 0;
  p.batch = nullptr;
  if (staged) {
    const int span_px = strip_rows_bound(in_w, out_w);               // same bound along the width: input columns one tile touches
    p.stage_pitch = 4 * ((3 * span_px + 6) / 4);
    p.stage_rows = (32 * 1024) / p.stage_pitch;
    if (p.stage_rows > p.strip_rows) p.stage_rows = p.strip_rows;
    if (p.stage_rows < 1) p.stage_rows = 1;
    smem = ((smem + 15) & ~size_t(15)) + (size_t)p.stage_rows * p.stage_pitch;
  }
  FNST_CHECK_ARG(smem <= 160 * 1024, "resize_to_tensor: strip of %d rows does not fit shared memory", p.strip_rows);
  FNST_DEVICE(device);
  auto kern = staged ? resize_to_tensor_kernel<true> : resize_to_tensor_kernel<false>;
  if (smem > 32 * 1024) {
    // opt in to > 48 KB once per (device, variant) with the largest size this entry point can ask for, not on every launch
    static std::atomic<unsigned long long> opted[2] = {{0ull}, {0ull}};
    const unsigned long long bit = 1ull << (device & 63);
    if (!(opted[staged ? 1 : 0].load(std::memory_order_acquire) & bit)) {
      FNST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      opted[staged ? 1 : 0].fetch_or(bit, std::memory_order_release);
    }
  }
  dim3 grid((out_w + RS_TILE - 1) / RS_TILE, (out_h + RS_TILE - 1) / RS_TILE);
  launch_pdl(kern, grid, dim3(RS_TILE * RS_TILE), smem, (cudaStream_t)stream, p);
  return launch_status("resize_to_tensor");
}

// Many images in ONE launch (grid z = image): the per-image launches of fnst_resize_to_tensor are 256 blocks of 256 threads
// each -- less than one wave of a 148-SM GPU, so a single image is latency-bound (13.7 us for a 6.2 MB 1080p frame = 0.45 TB/s);
// a batch keeps every SM busy with ~16 resident tiles whose coefficient set-up and byte loads overlap.
extern "C" int fnst_resize_batch_to_tensor(const fnst_image_desc* images_dev, int n, int max_in_h, int max_in_w, int out_h,
                                           int out_w, float* out_nchw, void* out_u8_nhwc, const float* mean3, const float* std3,
                                           int device, void* stream) {
  FNST_CHECK_ARG(images_dev && (out_nchw || out_u8_nhwc) && n > 0 && n <= 65535, "resize_batch_to_tensor: bad arguments");
  FNST_CHECK_ARG(max_in_h > 0 && max_in_w > 0 && out_h > 0 && out_w > 0, "resize_batch_to_tensor: bad sizes");
  FNST_CHECK_ARG((mean3 == nullptr) == (std3 == nullptr), "resize_batch_to_tensor: pass both mean and std, or neither");
  const AxisScale ah = axis_scale(max_in_w, out_w), av = axis_scale(max_in_h, out_h);
  FNST_CHECK_ARG(ah.ksize <= RS_KMAX && av.ksize <= RS_KMAX,
                 "resize_batch_to_tensor: down-scaling factor above %d is not supported (%dx%d -> %dx%d)", (RS_KMAX - 1) / 2, max_in_h,
                 max_in_w, out_h, out_w);
  ResizeParams p;
  memset(&p, 0, sizeof(p));
  p.batch = images_dev;
  p.out_h = out_h; p.out_w = out_w;
  p.out_f = out_nchw; p.out_u8 = reinterpret_cast<uint8_t*>(out_u8_nhwc);
  p.normalize = mean3 != nullptr;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3 ? mean3[c] : 0.f; p.std[c] = std3 ? std3[c] : 1.f; }
  // strip capacity for the tallest image (the bound grows with the input height; shorter images need less)
  p.strip_rows = strip_rows_bound(max_in_h, out_h);
  if (p.strip_rows < RS_TILE) p.strip_rows = RS_TILE;
  size_t smem = (size_t)p.strip_rows * RS_TILE * 3;
  // staged form (tuning knob resize_staged): the tile's input span goes to shared memory with coalesced 32-bit loads first
  const bool staged = tuning().resize_staged != 0 && max_in_w != out_w;
  if (staged) {
    const int span_px = strip_rows_bound(max_in_w, out_w);
    p.stage_pitch = 4 * ((3 * span_px + 6) / 4);
    p.stage_rows = (32 * 1024) / p.stage_pitch;
    if (p.stage_rows > p.strip_rows) p.stage_rows = p.strip_rows;
    if (p.stage_rows < 1) p.stage_rows = 1;
    smem = ((smem + 15) & ~size_t(15)) + (size_t)p.stage_rows * p.stage_pitch;
  }
  FNST_CHECK_ARG(smem <= 160 * 1024, "resize_batch_to_tensor: strip of %d rows does not fit shared memory", p.strip_rows);
  FNST_DEVICE(device);
  auto kern = staged ? resize_to_tensor_kernel<true> : resize_to_tensor_kernel<false>;
  if (smem > 32 * 1024) {
    static std::atomic<unsigned long long> opted[2] = {{0ull}, {0ull}};
    const unsigned long long bit = 1ull << (device & 63);
    if (!(opted[staged ? 1 : 0].load(std::memory_order_acquire) & bit)) {
      FNST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      opted[staged ? 1 : 0].fetch_or(bit, std::memory_order_release);
    }
  }
  dim3 grid((out_w + RS_TILE - 1) / RS_TILE, (out_h + RS_TILE - 1) / RS_TILE, n);
  launch_pdl(kern, grid, dim3(RS_TILE * RS_TILE), smem, (cudaStream_t)stream, p);
  return launch_status("resize_batch_to_tensor");
}

// Host twin of the device arithmetic (same functions compiled for the host): used by the tests to pin the coefficient
// windows and the whole two-pass pipeline against Pillow without a GPU.  Not a fallback: it is not reachable from the
// Python product layer (no binding in _lib.EXPORTS consumers other than tests).
extern "C" int fnst_resize_window_host(int in_size, int out_size, int index, int* first, int* len, int* kk, int kk_capacity) {
  const AxisScale a = axis_scale(in_size, out_size);
  FNST_CHECK_ARG(a.ksize <= kk_capacity && index >= 0 && index < out_size, "resize_window_host: bad arguments");
  axis_window(a, in_size, index, first, len, kk);
  return a.ksize;
}
