// Weight gradients and Gram matrices on tensor cores: "pixel-contraction" GEMMs
//
//     D[j][c] (+)= sum_{pixels p} G[p][j] * A[p + tap][c]
//
// where both operands are NHWC tensors (pixel-major), i.e. MN-major for the MMA: the contraction
// index (pixel) is the slow one.  wgrad: G = output gradient, A = (halo) activation, one D per tap.
// Gram: G = A = feature map, one D per image (losses/losses.py:6-13).
//
// CTA task = (tap, 128-row block of j, BN-column block of c, pixel split).  192 threads:
//   warp 0  TMA producer: per k-block (64 pixels = TW x TH box) two {64 j, TW, TH, 1} boxes of G and
//           BN/64 boxes {64 c, TW, TH, 1} of A shifted by the tap; 128-byte swizzle, 4-stage ring.
//           Out-of-range pixels are zero-filled by TMA on both sides (= conv zero padding / ragged edges).
//   warp 1  tcgen05.mma.kind::f16 with MN-major A and B descriptors, M=128 x N=BN x K=16, fp32 in TMEM.
//   warps 2..5  epilogue: tcgen05.ld -> red.global.add.v4.f32 into the (zeroed) fp32 result; split-K
//           partial tiles of different CTAs meet in L2 atomics.
#include "tc_common.cuh"
#include <cstdlib>

namespace fnst {

int validate_conv_desc(const fnst_conv_desc* d);

struct WgradTcParams {
  int32_t out_n, out_h, out_w;
  int32_t tiles_w, tiles_h, tw_log2;         // pixel box TW x TH = 64
  int32_t kblocks_total, kblocks_per_image;
  int32_t splits, jblocks, cblocks, ntaps;
  int32_t per_image;                          // 1: one result matrix per image (Gram)
  int32_t h0, w0, kc, n_gemm;
  int64_t out_row_stride, out_image_stride;
  uint32_t idesc;
  float* out;
  int8_t tap_dh[FNST_MAX_TAPS];
  int8_t tap_dw[FNST_MAX_TAPS];
  int16_t tap_c0[FNST_MAX_TAPS];
};

constexpr int WT_THREADS = 192;
constexpr int WT_KPIX = 64;                   // pixels per k-block
constexpr int WT_G_BYTES = 2 * WT_KPIX * 128; // two 64-wide j blocks

template <int BN> struct WtCfg {
  static constexpr int A_BYTES = (BN / 64) * WT_KPIX * 128;
  static constexpr int STAGE_BYTES = WT_G_BYTES + A_BYTES;
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_a,
                const __grid_constant__ WgradTcParams p) {
  using Cfg = WtCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();

  // task decode: tap fastest (neighbouring CTAs share the same G tiles in L2), then column block, row block, split
  int task = blockIdx.x;
  const int t = task % p.ntaps; task /= p.ntaps;
  const int cb = task % p.cblocks; task /= p.cblocks;
  const int jb = task % p.jblocks; task /= p.jblocks;
  const int split = task % p.splits;
  const int img = task / p.splits;              // only meaningful when per_image
  int kb0, kb1;
  {
    const int total = p.per_image ? p.kblocks_per_image : p.kblocks_total;
    const int base = total / p.splits, rem = total % p.splits;
    kb0 = split * base + min(split, rem);
    kb1 = kb0 + base + (split < rem ? 1 : 0);
    if (p.per_image) { kb0 += img * p.kblocks_per_image; kb1 += img * p.kblocks_per_image; }
  }
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_g);
    tma_prefetch_desc(&map_a);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int TW = 1 << p.tw_log2, TH = WT_KPIX >> p.tw_log2;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int dh = p.tap_dh[t] + p.h0, dw = p.tap_dw[t] + p.w0, c0 = p.tap_c0[t] + cb * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        int r = kb;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int n = r / p.tiles_h;
        const int h = th * TH, w = tw * TW;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sg = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sa = sg + WT_G_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        tma_load_4d(sg, &map_g, &full_bar[stage], jb * 128, w, h, n);
        tma_load_4d(sg + WT_KPIX * 128, &map_g, &full_bar[stage], jb * 128 + 64, w, h, n);
#pragma unroll
        for (int i = 0; i < BN / 64; ++i)
          tma_load_4d(sa + i * WT_KPIX * 128, &map_a, &full_bar[stage], c0 + i * 64, w + dw, h + dh, n);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sg = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        // MN-major, 128-byte swizzle: LBO = distance between 64-element M/N blocks, SBO = 8 K-rows (1024 B)
        const uint64_t dg = umma_smem_desc(sg, WT_KPIX * 128, 1024);
        const uint64_t da = umma_smem_desc(sg + WT_G_BYTES, WT_KPIX * 128, 1024);
#pragma unroll
        for (int k = 0; k < WT_KPIX / 16; ++k) {
          // 16 K-rows = 2048 bytes further: +128 in the (addr >> 4) field
          umma_f16(tmem_base, dg + (uint64_t)(128 * k), da + (uint64_t)(128 * k), p.idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);
      pdl_trigger_tail();
    }
    __syncwarp();
  } else if (nkb > 0) {
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row = jb * 128 + q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    float* orow = p.out + (p.per_image ? (int64_t)img * p.out_image_stride : 0) + (int64_t)row * p.out_row_stride +
                  (int64_t)t * p.kc + cb * BN;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t raw[32];
      tmem_ld_x32(t_row + c, raw);
      tmem_ld_wait();
      if (row < p.n_gemm) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          red_add_v4(orow + c + i, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]), __uint_as_float(raw[i + 2]),
                     __uint_as_float(raw[i + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN>
static int launch_wgrad_tc(const CUtensorMap& mg, const CUtensorMap& ma, const WgradTcParams& p, int tasks, cudaStream_t st) {
  auto kern = wgrad_tc_kernel<BN>;
  FNST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WtCfg<BN>::SMEM_BYTES));
  launch_pdl(kern, dim3(tasks), dim3(WT_THREADS), WtCfg<BN>::SMEM_BYTES, st, mg, ma, p);
  return launch_status("wgrad_tc");
}

// Shared host path.  g: NHWC [n, oh, ow, n_gemm] (g_dtype); a: activation view (a_dtype); out fp32.
static int run_pixel_gemm(const fnst_conv_desc* d, const void* g, int g_dtype, float* out, int64_t out_row_stride,
                          int64_t out_image_stride, int per_image, int device, cudaStream_t st) {
  FNST_CHECK_ARG((d->dtype == FNST_F16 || d->dtype == FNST_BF16) && (g_dtype == FNST_F16 || g_dtype == FNST_BF16),
                 "wgrad_tc: operands must be fp16 or bf16");
  FNST_CHECK_ARG(d->kc % 64 == 0 && d->n_gemm % 8 == 0, "wgrad_tc: kc %d must be a multiple of 64 (n_gemm %d of 8)", d->kc, d->n_gemm);
  FNST_CHECK_ARG(d->a_stride_w % 8 == 0 && d->a_stride_h % 8 == 0 && d->a_stride_n % 8 == 0, "wgrad_tc: A strides must be multiples of 8");
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);

  WgradTcParams p;
  memset(&p, 0, sizeof(p));
  p.out_n = d->out_n; p.out_h = d->out_h; p.out_w = d->out_w;
  p.tw_log2 = d->out_w >= 16 ? 4 : 3;
  const int TW = 1 << p.tw_log2, TH = WT_KPIX >> p.tw_log2;
  p.tiles_w = (d->out_w + TW - 1) / TW;
  p.tiles_h = (d->out_h + TH - 1) / TH;
  p.kblocks_per_image = p.tiles_w * p.tiles_h;
  p.kblocks_total = p.kblocks_per_image * d->out_n;
  int bn = d->kc % 256 == 0 ? 256 : (d->kc % 128 == 0 ? 128 : 64);
  if (const int f = tuning().wgrad_bn) { if (f <= bn && d->kc % f == 0) bn = f; }
  p.jblocks = (d->n_gemm + 127) / 128;
  p.cblocks = d->kc / bn;
  p.ntaps = d->ntaps;
  p.per_image = per_image;
  const int base_tasks = p.ntaps * p.cblocks * p.jblocks * (per_image ? d->out_n : 1);
  const int kb_avail = per_image ? p.kblocks_per_image : p.kblocks_total;
  const int waves_x2 = tuning().wgrad_waves_x2;
  int splits = (waves_x2 * sms / 2) / base_tasks;                 // at most waves_x2/2 full waves of tasks (no ragged extra wave)
  if (splits > kb_avail / 4) splits = kb_avail / 4;               // keep >= 4 k-blocks per task
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.h0 = d->h0; p.w0 = d->w0; p.kc = d->kc; p.n_gemm = d->n_gemm;
  p.out_row_stride = out_row_stride; p.out_image_stride = out_image_stride;
  p.out = out;
  // instruction descriptor: A operand = g (MN-major), B operand = activation (MN-major)
  uint32_t idesc = umma_idesc_f16(0, bn, 1, 1);
  idesc &= ~((7u << 7) | (7u << 10));
  idesc |= (uint32_t)(g_dtype == FNST_BF16 ? 1 : 0) << 7;
  idesc |= (uint32_t)(d->dtype == FNST_BF16 ? 1 : 0) << 10;
  p.idesc = idesc;
  memcpy(p.tap_dh, d->tap_dh, sizeof(p.tap_dh));
  memcpy(p.tap_dw, d->tap_dw, sizeof(p.tap_dw));
  memcpy(p.tap_c0, d->tap_c0, sizeof(p.tap_c0));

  CUtensorMap mg, ma;
  {
    const uint64_t dims[4] = {(uint64_t)d->n_gemm, (uint64_t)d->out_w, (uint64_t)d->out_h, (uint64_t)d->out_n};
    uint64_t str[3] = {(uint64_t)d->n_gemm * 2, (uint64_t)d->n_gemm * d->out_w * 2, (uint64_t)d->n_gemm * d->out_w * d->out_h * 2};
    if (d->g_stride_w) { str[0] = (uint64_t)d->g_stride_w * 2; str[1] = (uint64_t)d->g_stride_h * 2; str[2] = (uint64_t)d->g_stride_n * 2; }
    const uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    if (int r = encode_tensor_map_2b(&mg, g, 4, dims, str, box)) return r;
  }
  {
    const uint64_t dims[4] = {(uint64_t)d->a_c, (uint64_t)d->a_w, (uint64_t)d->a_h, (uint64_t)d->a_n};
    const uint64_t str[3] = {(uint64_t)d->a_stride_w * 2, (uint64_t)d->a_stride_h * 2, (uint64_t)d->a_stride_n * 2};
    const uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    if (int r = encode_tensor_map_2b(&ma, d->a, 4, dims, str, box)) return r;
  }
  const int tasks = base_tasks * splits;
  switch (bn) {
    case 64: return launch_wgrad_tc<64>(mg, ma, p, tasks, st);
    case 128: return launch_wgrad_tc<128>(mg, ma, p, tasks, st);
    default: return launch_wgrad_tc<256>(mg, ma, p, tasks, st);
  }
}

int gram_tc(const void* feat, float* out, int n, int hw, int c, int dtype, int device, cudaStream_t st) {
  // Gram = pixel-contraction GEMM with G = A = feat, one C x C result per image.  The pixel axis is presented
  // as a [rows][64] grid; a ragged tail is zero-filled by TMA.
  FNST_CHECK_ARG(c % 64 == 0, "gram_tc: channel count %d must be a multiple of 64", c);
  FNST_CHECK_ARG(hw % 8 == 0, "gram_tc: h*w = %d must be a multiple of 8", hw);
  fnst_conv_desc d;
  memset(&d, 0, sizeof(d));
  const int w = hw % 16 == 0 ? 16 : 8, h = hw / w;
  d.a = feat; d.b = feat; d.out = out;
  d.a_stride_w = c; d.a_stride_h = (int64_t)c * w; d.a_stride_n = (int64_t)c * hw;
  d.a_w = w; d.a_h = h; d.a_n = n; d.a_c = c;
  d.ntaps = 1; d.kc = c; d.n_gemm = c;
  d.out_n = n; d.out_h = h; d.out_w = w;
  d.dtype = dtype;
  return run_pixel_gemm(&d, feat, dtype, out, c, (int64_t)c * c, 1, device, st);
}

}  // namespace fnst

using namespace fnst;

extern "C" int fnst_wgrad_tc(const fnst_conv_desc* d, int g_dtype, int device, void* stream) {
  if (int r = validate_conv_desc(d)) return r;
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ktot = (size_t)d->ntaps * d->kc;
  if (!(d->flags & FNST_DESC_PREZEROED)) FNST_CUDA(cudaMemsetAsync(d->out, 0, sizeof(float) * ktot * d->n_gemm, st));
  return run_pixel_gemm(d, d->b, g_dtype, reinterpret_cast<float*>(d->out), (int64_t)ktot, 0, 0, device, st);
}
