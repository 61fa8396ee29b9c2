// Backward operators of the style-transfer path (see include/fnst.h, "Backward operators"):
// weight gradients of gather-GEMM convolutions, InstanceNorm backward (with ReflectionPad2d
// fold, ReLU mask, dropout and residual handling), max-pool / squared-error / TV backward.
#include "tc_common.cuh"

namespace fnst {

int validate_conv_desc(const fnst_conv_desc* d);

// ---------------------------------------------------------------------------------------------
// wgrad: dB[j][t*kc + c] = sum_pixels g[pix][j] * A[pix + tap t][c0_t + c]
// Block = 64 (j) x 64 (c) output tile of one tap, over a slab of pixels; 256 threads x (4x4).
// grid.x = ntaps * ceil(kc/64), grid.y = ceil(n_gemm/64), grid.z = image * row-slabs.
// ---------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 16;   // output rows per pixel slab

template <typename TA, typename TG>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const __grid_constant__ fnst_conv_desc d) {
  __shared__ float Gs[16][64 + 4];
  __shared__ float As[16][64 + 4];
  const int tid = threadIdx.x;
  const int cchunks = (d.kc + 63) / 64;
  const int t = blockIdx.x / cchunks, cc = blockIdx.x % cchunks;
  const int j0 = blockIdx.y * 64;
  const int slabs = (d.out_h + WG_ROWS - 1) / WG_ROWS;
  const int n = blockIdx.z / slabs, slab = blockIdx.z % slabs;
  const int r0 = slab * WG_ROWS, r1 = min(d.out_h, r0 + WG_ROWS);
  const int npix = (r1 - r0) * d.out_w;
  const int dh = d.tap_dh[t] + d.h0, dw = d.tap_dw[t] + d.w0, c0 = d.tap_c0[t] + cc * 64;
  const int cw = min(64, d.kc - cc * 64);           // valid channels in this chunk

  const TG* G = reinterpret_cast<const TG*>(d.b);
  const int64_t gsw = d.g_stride_w ? d.g_stride_w : d.n_gemm;
  const int64_t gsh = d.g_stride_w ? d.g_stride_h : (int64_t)d.n_gemm * d.out_w;
  const int64_t gsn = d.g_stride_w ? d.g_stride_n : (int64_t)d.n_gemm * d.out_w * d.out_h;
  const TA* A = reinterpret_cast<const TA*>(d.a) + (size_t)n * d.a_stride_n;
  const int lp = tid >> 4, lq = (tid & 15) * 4;     // loader: 16 pixels x 64 (4 per thread)
  const int tx = tid & 15, ty = tid >> 4;           // compute: ty -> j, tx -> c
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;

  for (int p0 = 0; p0 < npix; p0 += 16) {
    const int p = p0 + lp;
    float gv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < npix) {
      const int h = r0 + p / d.out_w, w = p % d.out_w;
      const TG* gp = G + (size_t)n * gsn + (size_t)h * gsh + (size_t)w * gsw + j0 + lq;
#pragma unroll
      for (int i = 0; i < 4; ++i) if (j0 + lq + i < d.n_gemm) gv[i] = to_f32<TG>(gp[i]);
      const int ih = h + dh, iw = w + dw;
      if (ih >= 0 && ih < d.a_h && iw >= 0 && iw < d.a_w) {
        const TA* ap = A + (size_t)ih * d.a_stride_h + (size_t)iw * d.a_stride_w + c0 + lq;
#pragma unroll
        for (int i = 0; i < 4; ++i) if (lq + i < cw) av[i] = to_f32<TA>(ap[i]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { Gs[lp][lq + i] = gv[i]; As[lp][lq + i] = av[i]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 g4 = *reinterpret_cast<const float4*>(&Gs[k][ty * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][tx * 4]);
      const float g[4] = {g4.x, g4.y, g4.z, g4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = fmaf(g[i], a[q], acc[i][q]);
    }
  }
  float* out = reinterpret_cast<float*>(d.out);
  const size_t ktot = (size_t)d.ntaps * d.kc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + ty * 4 + i;
    if (j >= d.n_gemm) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = tx * 4 + q;
      if (c < cw) atomicAdd(&out[(size_t)j * ktot + (size_t)t * d.kc + cc * 64 + c], acc[i][q]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// First-layer wgrad (C_in = 3, NCHW fp32 image).  Block = strip of 8x8 output tiles; thread =
// (output channel o, tap group q of 4); per tile: patch + g tile in smem, broadcast patch reads.
// ---------------------------------------------------------------------------------------------
constexpr int FW_MAX_TPT = 61;   // ceil(243 / 4) taps per thread

template <typename TG>
__global__ void __launch_bounds__(256) conv_first_wgrad_kernel(const float* __restrict__ x, int H, int W,
                                                               const TG* __restrict__ g, int c_out, int k, int stride, int pad,
                                                               int pad_mode, int Ho, int Wo, int tiles_per_block, float* __restrict__ dw) {
  extern __shared__ float smem[];
  const int pdim = 7 * stride + k;
  float* patch = smem;                         // [3][pdim][pdim]
  float* gs = smem + 3 * pdim * pdim;          // [64 pixels][64 o]
  const int taps = 3 * k * k;
  const int tid = threadIdx.x, o = tid & 63, q = tid >> 6;
  const int tiles_w = (Wo + 7) >> 3, tiles_h = (Ho + 7) >> 3;
  const int tiles_img = tiles_w * tiles_h;
  const int n = blockIdx.y;
  float acc[FW_MAX_TPT];
  int off[FW_MAX_TPT];          // patch offset of each of this thread's taps (invalid taps read offset 0, never stored)
#pragma unroll
  for (int i = 0; i < FW_MAX_TPT; ++i) {
    acc[i] = 0.f;
    const int t = q + 4 * i;
    const int c = t / (k * k), r = t - c * k * k;
    off[i] = t < taps ? (c * pdim + r / k) * pdim + r % k : 0;
  }

  for (int ti = 0; ti < tiles_per_block; ++ti) {
    const int tile = blockIdx.x * tiles_per_block + ti;
    if (tile >= tiles_img) break;
    const int tw = tile % tiles_w, th = tile / tiles_w;
    const int h_base = th * 8 * stride - pad, w_base = tw * 8 * stride - pad;
    __syncthreads();
    for (int i = tid; i < 3 * pdim * pdim; i += 256) {
      const int c = i / (pdim * pdim), r = i - c * pdim * pdim;
      int ih = h_base + r / pdim, iw = w_base + r % pdim;
      float v = 0.f;
      if (pad_mode == FNST_PAD_REFLECT) {
        ih = reflect_index(ih, H); iw = reflect_index(iw, W);
        ih = min(max(ih, 0), H - 1); iw = min(max(iw, 0), W - 1);
        v = x[(((size_t)n * 3 + c) * H + ih) * W + iw];
      } else if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
        v = x[(((size_t)n * 3 + c) * H + ih) * W + iw];
      }
      patch[i] = v;
    }
    for (int i = tid; i < 64 * 64; i += 256) {
      const int p = i >> 6, oo = i & 63;
      const int h = th * 8 + (p >> 3), w = tw * 8 + (p & 7);
      gs[i] = (h < Ho && w < Wo) ? to_f32<TG>(g[(((size_t)n * Ho + h) * Wo + w) * c_out + oo]) : 0.f;
    }
    __syncthreads();
    for (int p = 0; p < 64; ++p) {
      const float gv = gs[p * 64 + o];
      const float* pp = patch + (p >> 3) * stride * pdim + (p & 7) * stride;
#pragma unroll
      for (int i = 0; i < FW_MAX_TPT; ++i) acc[i] = fmaf(gv, pp[off[i]], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < FW_MAX_TPT; ++i) {
    const int t = q + 4 * i;
    if (t < taps) atomicAdd(&dw[(size_t)t * c_out + o], acc[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm backward
// ---------------------------------------------------------------------------------------------
struct HaloLayout {
  int H, W, C, pad, reflect, s2d;
  int slack;      // the buffer is allocated `slack` rows and columns larger than the halo extent (in its own pixel units)
  __device__ __forceinline__ size_t index(int n, int hp, int wp, int c) const {
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    if (!s2d) return (((size_t)n * (Hp + slack) + hp) * (Wp + slack) + wp) * C + c;
    const int Hs = ((Hp + 1) >> 1) + slack, Ws = ((Wp + 1) >> 1) + slack;
    return (((size_t)n * Hs + (hp >> 1)) * Ws + (wp >> 1)) * (4 * C) + ((hp & 1) * 2 + (wp & 1)) * C + c;
  }
  // padded coordinates whose value was copied from interior coordinate i (extent n): returns count (<= 3)
  __device__ __forceinline__ int sources(int i, int n, int (&src)[3]) const {
    int cnt = 0;
    src[cnt++] = i + pad;
    if (reflect) {
      if (i >= 1 && i <= pad) src[cnt++] = pad - i;
      if (i <= n - 2 && i >= n - 1 - pad) src[cnt++] = pad + 2 * (n - 1) - i;
    }
    return cnt;
  }
};

// Pass 1.  Block = (row group, image); thread = (8-channel group cg, pixel lane pl).  Per-channel constants are
// computed once per block into shared memory (one channel per thread) instead of 8 channels x 5 scalars per thread;
// the per-(n,c) partial sums go lane -> smem table -> one global atomic per channel per block.
// Two blocks per SM for the 16-bit variants (128 registers; 140 unconstrained = one block per SM = 8 warps, which ncu
// showed latency-bound on long-scoreboard stalls with the 256-block batch-4 grid running as 1.73 waves).
template <typename TA, typename TG, int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS) inorm_bwd_reduce_kernel(const TG* __restrict__ gsrc, const TG* __restrict__ extra,
                                                               const TA* __restrict__ raw, const float* __restrict__ stats,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const float* __restrict__ drop, TG* __restrict__ gy,
                                                               float* __restrict__ sums, float* __restrict__ dgb, HaloLayout L,
                                                               int relu, float eps, int rows_per_block) {
  pdl_trigger();
  extern __shared__ float sm[];      // par[5][C] (a, b, mean, rstd, dropout scale) | part[PL][2C]
  const int n = blockIdx.y;
  const int C = L.C, CG = C >> 3, PL = 256 / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, c0 = cg * 8;
  float* par = sm;
  float* part = sm + 5 * C;
  {
    const float inv_cnt = 1.f / (float)(L.H * L.W);
    for (int c = threadIdx.x; c < C; c += 256) {
      const float* st = stats + ((size_t)n * C + c) * 2;
      const float mean = st[0] * inv_cnt;
      const float rstd = rsqrtf(fmaxf(st[1] * inv_cnt - mean * mean, 0.f) + eps);
      const float ai = gamma[c] * rstd;
      par[c] = ai; par[C + c] = beta[c] - mean * ai; par[2 * C + c] = mean; par[3 * C + c] = rstd;
      par[4 * C + c] = drop ? drop[(size_t)n * C + c] : 1.f;
    }
  }
  // The table above only reads tensors of the FORWARD pass (statistics, affine parameters, dropout scales): it is built while
  // the predecessor (the data-gradient GEMM that produces gsrc) drains; gradients are touched only below this wait.
  pdl_wait();
  __syncthreads();
  float a[8], b[8], mean[8], rstd[8], ds[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = par[c0 + i]; b[i] = par[C + c0 + i]; mean[i] = par[2 * C + c0 + i]; rstd[i] = par[3 * C + c0 + i]; ds[i] = par[4 * C + c0 + i];
  }
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }

  constexpr int U = 2;               // pixels in flight per thread: all loads of a group are issued before the arithmetic
  for (int rr = 0; rr < rows_per_block; ++rr) {
    const int h = blockIdx.x * rows_per_block + rr;
    if (h >= L.H) break;
    int hs[3]; const int nh = L.sources(h, L.H, hs);
    for (int w0 = pl; w0 < L.W; w0 += U * PL) {
      Raw8<TG> gc[U], ge[U];
      Raw8<TA> xr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int w = w0 + u * PL;
        if (w < L.W) {
          const size_t idx = (((size_t)n * L.H + h) * L.W + w) * C + c0;
          if (gsrc) gc[u] = load_raw8<TG>(gsrc + L.index(n, h + L.pad, w + L.pad, c0));      // the unreflected source
          if (extra) ge[u] = load_raw8<TG>(extra + idx);
          xr[u] = load_raw8<TA>(raw + idx);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int w = w0 + u * PL;
        if (w >= L.W) continue;
        float g[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = 0.f;
        if (gsrc) {
          float t[8];
          raw8_to_f32<TG>(gc[u], t);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = t[i];
          int wsrc[3]; const int nw = L.sources(w, L.W, wsrc);
          if (nh > 1 || nw > 1) {          // ReflectionPad2d fold: halo positions that were copied from (h, w) (border pixels only)
            for (int ih = 0; ih < nh; ++ih)
              for (int iw = (ih == 0 ? 1 : 0); iw < nw; ++iw) {
                load8<TG>(gsrc + L.index(n, hs[ih], wsrc[iw], c0), t);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] += t[i];
              }
          }
        }
        const size_t idx = (((size_t)n * L.H + h) * L.W + w) * C + c0;
        if (extra) {
          float t[8];
          raw8_to_f32<TG>(ge[u], t);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] += t[i];
        }
        float x[8];
        raw8_to_f32<TA>(xr[u], x);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float y = fmaf(x[i], a[i], b[i]);
          float gv = g[i] * ds[i];
          if (relu && !(y > 0.f)) gv = 0.f;
          g[i] = gv;
          s1[i] += gv;
          s2[i] = fmaf(gv, (x[i] - mean[i]) * rstd[i], s2[i]);
        }
        store8<TG>(gy + idx, g);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { part[pl * 2 * C + c0 + i] = s1[i]; part[pl * 2 * C + C + c0 + i] = s2[i]; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float t = 0.f;
    for (int q = 0; q < PL; ++q) t += part[q * 2 * C + i];
    const int c = i < C ? i : i - C, which = i < C ? 0 : 1;
    atomicAdd(&sums[((size_t)n * C + c) * 2 + which], t);
    if (dgb) atomicAdd(&dgb[which ? c : C + c], t);      // d gamma = sum gy*xhat (row 0), d beta = sum gy (row 1)
  }
}

// Pass 1, TMA-staged form (2-byte tensors, plain halo layout, a whole padded row <= 256 pixels).  The register form above keeps
// two pixels x three tensors in flight per thread (four need ~180 registers): at batch 4 it is latency-bound at 2.1-2.7 TB/s on
// L2-resident data.  Here a CTA owns ONE image row: one thread issues the row's three tiles (raw, the padded gradient row, the
// residual-branch gradient; up to 97 KB) as TMA box loads the moment the dependency wait returns, so every byte of the CTA is
// in flight at once and two CTAs per SM keep ~200 KB in flight; the arithmetic then reads shared memory (512 contiguous bytes
// per warp: conflict-free).  Same operations in the same order as the register form: gy is bit-identical.
// Measured on B200 at 4 x 64 x 64 x 256 (profiles/r02_bench_inorm_tma.json): 11.5 us against 13.5 us alone (L2-hot), but the training
// step is 0.05 ms SLOWER with it -- a kernel this small is bound by its launch / prologue / tail latencies, and its 100 KB CTAs
// cannot share SMs with the concurrent weight-gradient GEMMs as the register form's small CTAs do.  Tested option, default off.
template <typename TA, typename TG>
__global__ void __launch_bounds__(256, 2)
inorm_bwd_reduce_tma_kernel(const __grid_constant__ CUtensorMap map_raw, const __grid_constant__ CUtensorMap map_g,
                            const __grid_constant__ CUtensorMap map_e, const TG* __restrict__ gsrc, const float* __restrict__ stats,
                            const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ drop,
                            TG* __restrict__ gy, float* __restrict__ sums, float* __restrict__ dgb, HaloLayout L, int relu, float eps,
                            int has_e) {
  pdl_trigger();
  extern __shared__ uint8_t sm_raw[];
  __shared__ uint64_t bar;
  const int n = blockIdx.y, h = blockIdx.x;
  const int C = L.C, CG = C >> 3, PL = 256 / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, c0 = cg * 8;
  const int Wp = L.W + 2 * L.pad;
  float* par = reinterpret_cast<float*>(sm_raw);                                   // [5][C]
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw + 20 * C) + 127) & ~uintptr_t(127));
  const TA* t_raw = reinterpret_cast<const TA*>(tiles);                           // [W][C]
  const TG* t_g = reinterpret_cast<const TG*>(tiles + (size_t)L.W * C * sizeof(TA));   // [Wp][C] (when gsrc)
  const TG* t_e = t_g + (gsrc ? (size_t)Wp * C : 0);                              // [W][C]  (when extra)
  float* part = reinterpret_cast<float*>(tiles);                                  // [PL][2C], re-uses the tiles after the loop
  {
    const float inv_cnt = 1.f / (float)(L.H * L.W);
    for (int c = threadIdx.x; c < C; c += 256) {
      const float* st = stats + ((size_t)n * C + c) * 2;
      const float mean = st[0] * inv_cnt;
      const float rstd = rsqrtf(fmaxf(st[1] * inv_cnt - mean * mean, 0.f) + eps);
      const float ai = gamma[c] * rstd;
      par[c] = ai; par[C + c] = beta[c] - mean * ai; par[2 * C + c] = mean; par[3 * C + c] = rstd;
      par[4 * C + c] = drop ? drop[(size_t)n * C + c] : 1.f;
    }
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  pdl_wait();                         // (the table above reads forward-pass tensors only)
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)((size_t)L.W * C * sizeof(TA) + (gsrc ? (size_t)Wp * C * sizeof(TG) : 0) +
                                      (has_e ? (size_t)L.W * C * sizeof(TG) : 0));
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_2d(const_cast<TA*>(t_raw), &map_raw, &bar, 0, (n * L.H + h) * L.W);
    if (gsrc) tma_load_4d(const_cast<TG*>(t_g), &map_g, &bar, 0, 0, h + L.pad, n);
    if (has_e) tma_load_2d(const_cast<TG*>(t_e), &map_e, &bar, 0, (n * L.H + h) * L.W);
  }
  float a[8], b[8], mean[8], rstd[8], ds[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = par[c0 + i]; b[i] = par[C + c0 + i]; mean[i] = par[2 * C + c0 + i]; rstd[i] = par[3 * C + c0 + i]; ds[i] = par[4 * C + c0 + i];
  }
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  int hs[3]; const int nh = L.sources(h, L.H, hs);
  mbar_wait(&bar, 0);
  for (int w = pl; w < L.W; w += PL) {
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = 0.f;
    if (gsrc) {
      float t[8];
      load8<TG>(t_g + (size_t)(w + L.pad) * C + c0, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = t[i];
      int wsrc[3]; const int nw = L.sources(w, L.W, wsrc);
      if (nh > 1 || nw > 1) {          // ReflectionPad2d fold: this row's halo columns from the tile, other halo rows from memory
        for (int ih = 0; ih < nh; ++ih)
          for (int iw = (ih == 0 ? 1 : 0); iw < nw; ++iw) {
            if (ih == 0) load8<TG>(t_g + (size_t)wsrc[iw] * C + c0, t);
            else load8<TG>(gsrc + L.index(n, hs[ih], wsrc[iw], c0), t);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] += t[i];
          }
      }
    }
    if (has_e) {
      float t[8];
      load8<TG>(t_e + (size_t)w * C + c0, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += t[i];
    }
    float x[8];
    load8<TA>(t_raw + (size_t)w * C + c0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float y = fmaf(x[i], a[i], b[i]);
      float gv = g[i] * ds[i];
      if (relu && !(y > 0.f)) gv = 0.f;
      g[i] = gv;
      s1[i] += gv;
      s2[i] = fmaf(gv, (x[i] - mean[i]) * rstd[i], s2[i]);
    }
    store8<TG>(gy + (((size_t)n * L.H + h) * L.W + w) * C + c0, g);
  }
  __syncthreads();                    // every thread is done with the tiles: their memory becomes the reduction scratch
#pragma unroll
  for (int i = 0; i < 8; ++i) { part[pl * 2 * C + c0 + i] = s1[i]; part[pl * 2 * C + C + c0 + i] = s2[i]; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float t = 0.f;
    for (int q = 0; q < PL; ++q) t += part[q * 2 * C + i];
    const int c = i < C ? i : i - C, which = i < C ? 0 : 1;
    atomicAdd(&sums[((size_t)n * C + c) * 2 + which], t);
    if (dgb) atomicAdd(&dgb[which ? c : C + c], t);
  }
}

template <typename TA, typename TG>
__global__ void __launch_bounds__(256) inorm_bwd_apply_kernel(const TG* __restrict__ gy, const TA* __restrict__ raw,
                                                              const float* __restrict__ stats, const float* __restrict__ sums,
                                                              const float* __restrict__ gamma, TG* __restrict__ draw,
                                                              int H, int W, int C, float eps, int out_s2d, float* __restrict__ dgb,
                                                              int rows_per_block, int out_pad) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float par[];     // [5][C]: k0 = gamma*rstd, mean(gy), mean(gy*xhat), mean, rstd
  const int n = blockIdx.y;
  if (dgb && blockIdx.x == 0 && blockIdx.y == 0) {
    // d gamma = sum_n sum gy*xhat, d beta = sum_n sum gy: from the per-(n,c) sums of pass 1 (one block, plain stores --
    // accumulating them with atomics in pass 1 would put every block of the grid on the same 2C addresses)
    const int N = gridDim.y;
    for (int c = threadIdx.x; c < C; c += 256) {
      float dg = 0.f, db = 0.f;
      for (int i = 0; i < N; ++i) { db += sums[((size_t)i * C + c) * 2 + 0]; dg += sums[((size_t)i * C + c) * 2 + 1]; }
      dgb[c] = dg; dgb[C + c] = db;
    }
  }
  {
    const float inv_cnt = 1.f / (float)(H * W);
    for (int c = threadIdx.x; c < C; c += 256) {
      const float* st = stats + ((size_t)n * C + c) * 2;
      const float* smv = sums + ((size_t)n * C + c) * 2;
      const float mean = st[0] * inv_cnt;
      const float rstd = rsqrtf(fmaxf(st[1] * inv_cnt - mean * mean, 0.f) + eps);
      par[c] = gamma[c] * rstd; par[C + c] = smv[0] * inv_cnt; par[2 * C + c] = smv[1] * inv_cnt;
      par[3 * C + c] = mean; par[4 * C + c] = rstd;
    }
  }
  __syncthreads();
  const int CG = C >> 3, PL = 256 / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, c0 = cg * 8;
  float mean[8], rstd[8], k0[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    k0[i] = par[c0 + i]; m1[i] = par[C + c0 + i]; m2[i] = par[2 * C + c0 + i]; mean[i] = par[3 * C + c0 + i]; rstd[i] = par[4 * C + c0 + i];
  }
  constexpr int U = 2;
  for (int rr = 0; rr < rows_per_block; ++rr) {
    const int h = blockIdx.x * rows_per_block + rr;
    if (h >= H) break;
    for (int w0 = pl; w0 < W; w0 += U * PL) {
      Raw8<TG> gr[U];
      Raw8<TA> xr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int w = w0 + u * PL;
        if (w < W) {
          const size_t idx = (((size_t)n * H + h) * W + w) * C + c0;
          gr[u] = load_raw8<TG>(gy + idx);
          xr[u] = load_raw8<TA>(raw + idx);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int w = w0 + u * PL;
        if (w >= W) continue;
        const size_t idx = (((size_t)n * H + h) * W + w) * C + c0;
        float g[8], x[8];
        raw8_to_f32<TG>(gr[u], g);
        raw8_to_f32<TA>(xr[u], x);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = k0[i] * (g[i] - m1[i] - (x[i] - mean[i]) * rstd[i] * m2[i]);
        TG* dst = out_s2d ? draw + (((size_t)n * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1)) * (4 * C) + ((h & 1) * 2 + (w & 1)) * C + c0
                 : out_pad ? draw + (((size_t)n * (H + 2 * out_pad) + h + out_pad) * (W + 2 * out_pad) + w + out_pad) * C + c0
                           : draw + idx;
        store8<TG>(dst, g);
      }
    }
  }
  if (out_pad) {
    // zero halo of the padded output [n][H+2p][W+2p][C]: the 2p border pixels of this block's rows, and an equal share of
    // the p top + p bottom rows (the consumer is a zero-padded gather-GEMM over the linear pixel stream of this buffer)
    const int Wp = W + 2 * out_pad, Hp = H + 2 * out_pad;
    float z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = 0.f;
    TG* img = draw + (size_t)n * Hp * Wp * C;
    const int h_lo = blockIdx.x * rows_per_block, h_hi = min(h_lo + rows_per_block, H);
    for (int i = pl; i < (h_hi - h_lo) * 2 * out_pad; i += PL) {
      const int h = h_lo + i / (2 * out_pad), k = i % (2 * out_pad);
      const int wp = k < out_pad ? k : W + k;                       // left border [0,p), right border [W+p, W+2p)
      store8<TG>(img + ((size_t)(h + out_pad) * Wp + wp) * C + c0, z);
    }
    const int frame = 2 * out_pad * Wp;                              // pixels of the top and bottom rows
    const int per = (frame + gridDim.x - 1) / gridDim.x;
    for (int i = blockIdx.x * per + pl; i < min(frame, (int)(blockIdx.x + 1) * per); i += PL) {
      const int r = i / Wp, wp = i - r * Wp;
      const int hp = r < out_pad ? r : H + r;                        // rows [0,p) and [H+p, H+2p)
      store8<TG>(img + ((size_t)hp * Wp + wp) * C + c0, z);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// One-pass InstanceNorm backward.  A slab = (image n, 16-channel group) is owned by a thread-block CLUSTER of K CTAs
// (K = 1, 2, 4 or 8; CTA k takes rows [k*rows_per_part, ...)).  Each CTA streams its part of the slab ONCE:
//   stage    one thread issues TMA box loads {16 channels, W pixels, R rows} of raw, gsrc (interior of the halo buffer) and
//            extra for the WHOLE part up front -- every byte the CTA needs is in flight at once (a latency-bound register
//            pipeline reached 0.8-1.1 TB/s here; the bulk loads arrive at memory speed), chunk by chunk on mbarriers;
//   phase 1  as chunks land: g = (fold(gsrc) + extra) * dropout * relu-mask overwrites the gsrc tile in shared memory;
//            per-channel sum g and sum g*xhat are reduced warp -> CTA -> cluster (partials exchanged through distributed
//            shared memory, summed in rank order: deterministic);
//   phase 2  d_raw = gamma*rstd * (g - mean(g) - xhat * mean(g*xhat)) from shared memory.
// HBM/L2 traffic: read gsrc (+extra) + raw, write d_raw (+gy): nothing is read twice, no gy round trip, no atomics.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dsmem_map(const void* p, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ float dsmem_ld_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int IBF_THREADS = 512;
constexpr int IBF_MAX_CHUNKS = 32;           // TMA chunks (<= 256 pixels each) per CTA

struct IbfParams {
  HaloLayout L;
  int relu, out_s2d, rows_per_part, K, chunk_rows, chunks, chunk_stride;   // chunk_stride: pixels between chunk starts in shared memory
  int cw, tpp_log2;                                                         // channels per slab (16/32/64), log2(threads per pixel)
  float eps;
  unsigned long long* dbg;                                                  // measurement only (fnst_set_debug_buffer): 8 stamps per CTA
};

__device__ __forceinline__ void dsmem_st_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template <typename TA, typename TG>
__global__ void __launch_bounds__(IBF_THREADS, 1)
inorm_bwd_fused_kernel(const __grid_constant__ CUtensorMap map_raw, const __grid_constant__ CUtensorMap map_g,
                       const __grid_constant__ CUtensorMap map_e, const TG* __restrict__ gsrc, int has_extra,
                       const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const float* __restrict__ drop, TG* __restrict__ draw, TG* __restrict__ gy_out, float* __restrict__ sums,
                       const IbfParams P) {
  pdl_trigger();
  extern __shared__ uint8_t ibf_raw_smem[];
  uint8_t* ibf_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ibf_raw_smem) + 127) & ~uintptr_t(127));
  __shared__ float red[IBF_THREADS / 32][128];   // per warp: [sub][s1 x 8 | s2 x 8]
  __shared__ float peer_tot[8][128];             // one row per cluster rank, written by that rank (remote stores)
  __shared__ __align__(8) uint64_t full[IBF_MAX_CHUNKS];
  const HaloLayout& L = P.L;
  const int part = blockIdx.x, n = blockIdx.z, K = P.K;
  unsigned long long* tl = P.dbg ? P.dbg + 8 * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
  auto stamp = [&](int slot) {
    if (tl && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); tl[slot] = t; }
  };
  stamp(0);
  const int TPP = 1 << P.tpp_log2, PL = IBF_THREADS >> P.tpp_log2, cw = P.cw;
  const int sub = threadIdx.x & (TPP - 1), pl = threadIdx.x >> P.tpp_log2;
  const int c0 = blockIdx.y * cw + sub * 8;
  const int C = L.C, H = L.H, W = L.W;
  const int r0 = part * P.rows_per_part, r1 = min(H, r0 + P.rows_per_part);
  const int nrows = max(0, r1 - r0);
  const size_t cap = (size_t)P.chunks * P.chunk_stride;                // pixel slots of each shared-memory tile
  TA* x_s = reinterpret_cast<TA*>(ibf_smem);
  TG* g_s = reinterpret_cast<TG*>(ibf_smem + cap * cw * sizeof(TA));  // gsrc tile, overwritten by g
  TG* e_s = g_s + cap * cw;                                            // extra tile (when present)
  if (!gsrc) { e_s = g_s; }                                            // extra only: its tile is the one overwritten by g
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_raw);
    if (gsrc) tma_prefetch_desc(&map_g);
    if (has_extra) tma_prefetch_desc(&map_e);
    for (int i = 0; i < P.chunks; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  if (threadIdx.x == 0) {
    // element unit of the maps is 2 bytes: channel coordinates / boxes are scaled by the element size
    const int ca = blockIdx.y * cw * (int)(sizeof(TA) / 2), cg = blockIdx.y * cw * (int)(sizeof(TG) / 2);
    const uint32_t bytes = (uint32_t)(P.chunk_rows * W) * cw * (uint32_t)(sizeof(TA) + (gsrc ? sizeof(TG) : 0) + (has_extra ? sizeof(TG) : 0));
    for (int i = 0; i < P.chunks; ++i) {
      const int r = r0 + i * P.chunk_rows;
      const size_t so = (size_t)i * P.chunk_stride * cw;
      mbar_arrive_expect_tx(&full[i], bytes);
      tma_load_4d(x_s + so, &map_raw, &full[i], ca, 0, r, n);
      if (gsrc) tma_load_4d(g_s + so, &map_g, &full[i], cg, L.pad, r + L.pad, n);
      if (has_extra) tma_load_4d(e_s + so, &map_e, &full[i], cg, 0, r, n);
    }
  }

  // per-channel constants (their global-memory latency overlaps the bulk loads)
  float a[8], b[8], mean[8], rstd[8], ds[8];
  {
    const float inv_cnt = 1.f / (float)(H * W);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* st = stats + ((size_t)n * C + c0 + i) * 2;
      const float m = st[0] * inv_cnt;
      const float r = rsqrtf(fmaxf(st[1] * inv_cnt - m * m, 0.f) + P.eps);
      mean[i] = m; rstd[i] = r;
      a[i] = gamma[c0 + i] * r; b[i] = beta[c0 + i] - m * a[i];
      ds[i] = drop ? drop[(size_t)n * C + c0 + i] : 1.f;
    }
  }

  // ReflectionPad2d backward ("fold"): halo positions of the consumer's buffer that mirror an interior pixel add their gradient
  // to it.  Only border pixels have such sources; they are enumerated explicitly and spread over ALL threads (inside the
  // streaming loop the same few warps would meet a border column in every iteration and serialise one global-memory latency
  // per pixel).  fold_tasks(PREFETCH): touch the sources while the bulk loads are in flight; fold_tasks(APPLY): add them into
  // the staged tile (read straight from global memory: written by the data-gradient GEMM just before, now L1/L2 hits).
  auto fold_tasks = [&](const bool apply) {
    const int p2 = 2 * L.pad;
    for (int pass = 0; pass < 2; ++pass) {
      // pass 0: rows that are mirrored themselves (h in [1, pad] or [H-1-pad, H-2]): every pixel; pass 1: all other rows: 2*pad columns
      const int per_row = pass == 0 ? W : p2;
      for (int t = pl; t < nrows * per_row; t += PL) {
        const int hr = t / per_row, j = t - hr * per_row, h = r0 + hr;
        const bool row_mirrored = (h >= 1 && h <= L.pad) || (h <= H - 2 && h >= H - 1 - L.pad);
        int w;
        if (pass == 0) {
          if (!row_mirrored) continue;
          w = j;
        } else {
          if (row_mirrored) continue;
          w = j < L.pad ? 1 + j : W - 1 - L.pad + (j - L.pad);
          if (w < 1 || w > W - 2 || (j >= L.pad && w <= L.pad)) continue;      // narrow planes: the two column groups overlap
        }
        int hs[3], wsrc[3];
        const int nh = L.sources(h, H, hs), nw = L.sources(w, W, wsrc);
        if (nh == 1 && nw == 1) continue;
        if (!apply) {
          for (int ih = 0; ih < nh; ++ih)
            for (int iw = (ih == 0 ? 1 : 0); iw < nw; ++iw)
              asm volatile("prefetch.global.L1 [%0];" ::"l"(gsrc + L.index(n, hs[ih], wsrc[iw], c0)));
          continue;
        }
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int ih = 0; ih < nh; ++ih)
          for (int iw = (ih == 0 ? 1 : 0); iw < nw; ++iw) {
            float tt[8];
            load8<TG>(gsrc + L.index(n, hs[ih], wsrc[iw], c0), tt);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += tt[i];
          }
        const int chh = hr / P.chunk_rows;
        const size_t so = ((size_t)chh * P.chunk_stride + (size_t)(hr - chh * P.chunk_rows) * W + w) * cw + sub * 8;
        float g[8];
        load8<TG>(g_s + so, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] += acc[i];
        store8<TG>(g_s + so, g);
      }
    }
  };
  const bool folding = gsrc && L.reflect && L.pad > 0;
  if (folding) fold_tasks(false);
  stamp(1);
  for (int ch = 0; ch < P.chunks; ++ch) mbar_wait(&full[ch], 0);
  stamp(2);
  if (folding) {
    fold_tasks(true);
    __syncthreads();
  }
  stamp(3);

  // ---- phase 1: g = (gsrc + extra) * dropout * relu-mask; s1 = sum g, s2 = sum g*x (xhat is affine in x: corrected below) ----
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  const int npx = nrows * W;
  const int step_rows = PL / W, step_w = PL - step_rows * W;           // pixel index advances by PL per iteration
  {
    int hr = pl / W, w = pl - hr * W;
    for (int p = pl; p < npx; p += PL) {
      const int chh = hr / P.chunk_rows;
      const size_t so = ((size_t)chh * P.chunk_stride + (size_t)(hr - chh * P.chunk_rows) * W + w) * cw + sub * 8;
      float g[8], x[8];
      load8<TA>(x_s + so, x);
      if (gsrc) {
        load8<TG>(g_s + so, g);
        if (has_extra) {
          float t[8];
          load8<TG>(e_s + so, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] += t[i];
        }
      } else {
        load8<TG>(e_s + so, g);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float gv = g[i] * ds[i];
        if (P.relu && !(fmaf(x[i], a[i], b[i]) > 0.f)) gv = 0.f;
        g[i] = gv;
        s1[i] += gv;
        s2[i] = fmaf(gv, x[i], s2[i]);
      }
      store8<TG>(g_s + so, g);
      if (gy_out) store8<TG>(gy_out + (((size_t)n * H + r0 + hr) * W + w) * C + c0, g);
      hr += step_rows; w += step_w;
      if (w >= W) { w -= W; ++hr; }
    }
  }
  stamp(4);

  // ---- reduction: lanes of equal channel group -> warp -> CTA -> cluster (fixed order everywhere) ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s2[i] = (s2[i] - mean[i] * s1[i]) * rstd[i];           // sum g*xhat = rstd * (sum g*x - mean * sum g)   (per thread: exact algebra)
    for (int o = TPP; o < 32; o <<= 1) {
      s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
      s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane < TPP) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[wid][lane * 16 + i] = s1[i]; red[wid][lane * 16 + 8 + i] = s2[i]; }
  }
  __syncthreads();
  const int nred = TPP * 16;
  if ((int)threadIdx.x < nred) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < IBF_THREADS / 32; ++q) t += red[q][threadIdx.x];
    if (K > 1) {
      // push this CTA's partial into row `part` of every cluster member's table (remote stores do not stall); after the
      // cluster barrier every member sums the K rows in rank order -- identical, deterministic totals everywhere
      for (int r = 0; r < K; ++r) dsmem_st_f32(dsmem_map(&peer_tot[part][threadIdx.x], (uint32_t)r), t);
    } else {
      peer_tot[0][threadIdx.x] = t;
    }
  }
  if (K > 1) cluster_barrier(); else __syncthreads();
  float m1[8], m2[8];
  {
    const float inv_cnt = 1.f / (float)(H * W);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t1 = 0.f, t2 = 0.f;
      for (int r = 0; r < K; ++r) { t1 += peer_tot[r][sub * 16 + i]; t2 += peer_tot[r][sub * 16 + 8 + i]; }
      m1[i] = t1; m2[i] = t2;
    }
    if (part == 0 && pl == 0) {
      // per-(n,c) sums for d gamma / d beta (summed over images by fnst_affine_grads): [n][c][0] = sum g, [1] = sum g*xhat
#pragma unroll
      for (int i = 0; i < 8; ++i) { sums[((size_t)n * C + c0 + i) * 2 + 0] = m1[i]; sums[((size_t)n * C + c0 + i) * 2 + 1] = m2[i]; }
    }
    // d_raw = a*(g - mean_g - xhat*mean_gx) = A*g + B*x + D with xhat = (x - mean)*rstd
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float mg = m1[i] * inv_cnt, mgx = m2[i] * inv_cnt;
      const float Bc = -a[i] * rstd[i] * mgx;
      m1[i] = Bc;                                          // coefficient of x
      m2[i] = -a[i] * mg - Bc * mean[i];                   // constant term
    }
  }
  stamp(5);

  // ---- phase 2 ------------------------------------------------------------------------------
  {
    int hr = pl / W, w = pl - hr * W;
    for (int p = pl; p < npx; p += PL) {
      const int chh = hr / P.chunk_rows, h = r0 + hr;
      const size_t so = ((size_t)chh * P.chunk_stride + (size_t)(hr - chh * P.chunk_rows) * W + w) * cw + sub * 8;
      float g[8], x[8];
      load8<TG>(g_s + so, g);
      load8<TA>(x_s + so, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = fmaf(a[i], g[i], fmaf(m1[i], x[i], m2[i]));
      TG* dst = P.out_s2d ? draw + (((size_t)n * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1)) * (4 * C) + ((h & 1) * 2 + (w & 1)) * C + c0
                          : draw + (((size_t)n * H + h) * W + w) * C + c0;
      store8<TG>(dst, g);
      hr += step_rows; w += step_w;
      if (w >= W) { w -= W; ++hr; }
    }
  }
  stamp(6);
  // (no exit barrier: after the cluster barrier above nobody touches a partner's shared memory any more)
}

// d gamma[c] = sum_n sums[n][c][1], d beta[c] = sum_n sums[n][c][0] for a list of InstanceNorm layers in one launch
// (fixed summation order).  table[l] = {offset of the layer's [N][C][2] block in `sums`, C, offset of d gamma in `out`,
// offset of d beta in `out`}; one block per (layer, 256-channel chunk).
struct AffineGradEntry { int64_t src, dgamma, dbeta; int32_t C, pad_; };
__global__ void __launch_bounds__(256) affine_grads_kernel(const float* __restrict__ sums, const AffineGradEntry* __restrict__ table,
                                                           int N, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const AffineGradEntry e = table[blockIdx.x];
  for (int c = blockIdx.y * 256 + threadIdx.x; c < e.C; c += gridDim.y * 256) {
    float dg = 0.f, db = 0.f;
    for (int i = 0; i < N; ++i) { db += sums[e.src + ((size_t)i * e.C + c) * 2 + 0]; dg += sums[e.src + ((size_t)i * e.C + c) * 2 + 1]; }
    out[e.dgamma + c] = dg; out[e.dbeta + c] = db;
  }
}

template <typename TA, typename TG>
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const TA* __restrict__ in, const TG* __restrict__ gout,
                                                           const TG* __restrict__ extra, TG* __restrict__ gin,
                                                           int N, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H >> 1, Wo = W >> 1, CG = C >> 3;
  // one thread per (2x2 window incl. the odd tail handled below, 8 channels)
  const int Hc = (H + 1) >> 1, Wc = (W + 1) >> 1;
  const size_t total = (size_t)N * Hc * Wc * CG;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cg = i % CG; size_t r = i / CG;
    const int wo = r % Wc; r /= Wc;
    const int ho = r % Hc; const int n = r / Hc;
    const bool pooled = ho < Ho && wo < Wo;      // odd trailing row/column is not covered by any window
    // all loads of the window (pooled gradient, four inputs, four residual-branch gradients) are issued before any of them
    // is used: one round trip to memory instead of two (the residual loads used to wait for the arg-max of the inputs)
    Raw8<TG> go_r, ex_r[4];
    Raw8<TA> in_r[4];
    bool ok[4];
    if (pooled) go_r = load_raw8<TG>(gout + (((size_t)n * Ho + ho) * Wo + wo) * C + cg * 8);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int hh = 2 * ho + (q >> 1), ww = 2 * wo + (q & 1);
      ok[q] = hh < H && ww < W;
      if (ok[q]) {
        const size_t idx = (((size_t)n * H + hh) * W + ww) * C + cg * 8;
        in_r[q] = load_raw8<TA>(in + idx);
        if (extra) ex_r[q] = load_raw8<TG>(extra + idx);
      }
    }
    float go[8];
    if (pooled) raw8_to_f32<TG>(go_r, go);
    float v[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (ok[q]) raw8_to_f32<TA>(in_r[q], v[q]);
    int best[8];
    if (pooled) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int bi = 0; float bv = v[0][k];
#pragma unroll
        for (int q = 1; q < 4; ++q) if (v[q][k] > bv) { bv = v[q][k]; bi = q; }
        best[k] = bi;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!ok[q]) continue;
      const int hh = 2 * ho + (q >> 1), ww = 2 * wo + (q & 1);
      const size_t idx = (((size_t)n * H + hh) * W + ww) * C + cg * 8;
      float o[8];
      if (extra) raw8_to_f32<TG>(ex_r[q], o);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (pooled && best[k] == q) o[k] += go[k];
        if (!(v[q][k] > 0.f)) o[k] = 0.f;
      }
      store8<TG>(gin + idx, o);
    }
  }
}

template <typename TA, typename TB, typename TG>
__global__ void __launch_bounds__(256) sse_bwd_kernel(const TA* __restrict__ a, const TB* __restrict__ b, int64_t count,
                                                      int64_t period, const float* __restrict__ scale, float coef,
                                                      TG* __restrict__ da, int relu_mask) {
  pdl_trigger();
  pdl_wait();
  const float s = 2.f * coef * scale[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float av = to_f32<TA>(a[i]);
    float v = s * (av - to_f32<TB>(b[i % period]));
    if (relu_mask && !(av > 0.f)) v = 0.f;
    da[i] = from_f32<TG>(v);
  }
}

template <typename TA, typename TG>
__global__ void __launch_bounds__(256) relu_mask_kernel(const TG* __restrict__ g, const TG* __restrict__ extra,
                                                        const TA* __restrict__ act, TG* __restrict__ out, int64_t count8) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8], a[8];
    load8<TG>(g + i * 8, v);
    load8<TA>(act + i * 8, a);
    if (extra) {
      float e[8];
      load8<TG>(extra + i * 8, e);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += e[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = a[k] > 0.f ? v[k] : 0.f;
    store8<TG>(out + i * 8, v);
  }
}

// Style-loss backward factor: S[n][i][j] = scale[0] * coef * ((G - Gt)[n][i][j] + (G - Gt)[n][j][i]), written in the
// feature element type; dF[n] = F[n] * S[n]  (gradient of coef' * sum (G - Gt)^2 through G = F^T F, losses.py:6-44).
template <typename TG>
__global__ void __launch_bounds__(256) gram_diff_sym_kernel(const float* __restrict__ G, const float* __restrict__ Gt, int64_t period,
                                                            int C, int64_t total, const float* __restrict__ scale, float coef,
                                                            TG* __restrict__ S) {
  pdl_trigger();
  pdl_wait();
  const float k = scale[0] * coef;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / ((int64_t)C * C), r = i - n * (int64_t)C * C;
    const int a = r / C, b = r - (int64_t)a * C;
    const int64_t t = n * (int64_t)C * C + (int64_t)b * C + a;
    const float d = (G[i] - Gt[i % period]) + (G[t] - Gt[t % period]);
    S[i] = from_f32<TG>(k * d);
  }
}

__global__ void __launch_bounds__(256) tv_bwd_kernel(const float* __restrict__ img, int planes, int H, int W,
                                                     const float* __restrict__ scale, float coef, float* __restrict__ dimg) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)planes * H * W;
  const float s = 2.f * coef * scale[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = i % W; const int y = (i / W) % H;
    const float v = img[i];
    float d = 0.f;
    if (y > 0) d += v - img[i - W];
    if (y + 1 < H) d -= img[i + W] - v;
    if (x > 0) d += v - img[i - 1];
    if (x + 1 < W) d -= img[i + 1] - v;
    dimg[i] = s * d;
  }
}

__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ x, int N, int C, int HW, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  // grid = (chunks, C); each block reduces a chunk of one channel over all images
  const int c = blockIdx.y;
  float s = 0.f;
  const int64_t total = (int64_t)N * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = i / HW; const int p = i - (int64_t)n * HW;
    s += x[((size_t)n * C + c) * HW + p];
  }
  __shared__ float part[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(&out[c], t);
  }
}

// Rows per block of the InstanceNorm backward kernels: the per-block set-up (channel constants, partial-sum flush) is
// amortised over several rows once the grid is large enough to fill the GPU; single rows for small problems.
static int rows_per_block(int h, int n) {
  int r = 1;
  while (r < 8 && (int64_t)n * ((h + 2 * r - 1) / (2 * r)) >= 148 * 8) r *= 2;
  return r;
}

static int grid_cap(int64_t items) {
  int64_t b = (items + 255) / 256;
  const int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace fnst

using namespace fnst;

extern "C" int fnst_wgrad_simt(const fnst_conv_desc* d, int g_dtype, int device, void* stream) {
  if (int r = validate_conv_desc(d)) return r;
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ktot = (size_t)d->ntaps * d->kc;
  if (!(d->flags & FNST_DESC_PREZEROED)) FNST_CUDA(cudaMemsetAsync(d->out, 0, sizeof(float) * ktot * d->n_gemm, st));
  const int slabs = (d->out_h + WG_ROWS - 1) / WG_ROWS;
  dim3 grid(d->ntaps * ((d->kc + 63) / 64), (d->n_gemm + 63) / 64, d->out_n * slabs);
  FNST_CHECK_ARG(grid.z <= 65535, "wgrad: too many pixel slabs (%u)", grid.z);
  FNST_DISPATCH_DTYPE(d->dtype, TA, {
    FNST_DISPATCH_DTYPE(g_dtype, TG, { wgrad_simt_kernel<TA, TG><<<grid, 256, 0, st>>>(*d); });
  });
  return launch_status("wgrad_simt");
}

extern "C" int fnst_conv_first_wgrad(const float* x, int n, int h, int w, const void* g, int g_dtype, int c_out, int k,
                                     int stride, int pad, int pad_mode, float* dw, int device, void* stream) {
  FNST_CHECK_ARG(x && g && dw, "conv_first_wgrad: null pointer");
  FNST_CHECK_ARG(c_out == 64 && k % 2 == 1 && k <= 9 && (stride == 1 || stride == 2), "conv_first_wgrad: unsupported configuration");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  FNST_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 3 * k * k * c_out, st));
  const int pdim = 7 * stride + k;
  const size_t smem = sizeof(float) * (3 * pdim * pdim + 64 * 64);
  const int tiles = ((wo + 7) / 8) * ((ho + 7) / 8);
  const int tpb = 4;
  dim3 grid((tiles + tpb - 1) / tpb, n);
  FNST_DISPATCH_DTYPE(g_dtype, TG, {
    conv_first_wgrad_kernel<TG><<<grid, 256, smem, st>>>(x, h, w, reinterpret_cast<const TG*>(g), c_out, k, stride, pad,
                                                        pad_mode, ho, wo, tpb, dw);
  });
  return launch_status("conv_first_wgrad");
}

static int encode_tensor_map_plain(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                   const uint32_t* box);

extern "C" int fnst_inorm_bwd_reduce(const void* gsrc, const void* extra, const void* raw, const float* stats,
                                     const float* gamma, const float* beta, const float* drop, void* gy, float* sums,
                                     float* dgb, int n, int h, int w, int c, int act_dtype, int g_dtype, int relu, float eps,
                                     int pad, int pad_mode, int s2d, int gsrc_slack, int prezeroed, int device, void* stream) {
  FNST_CHECK_ARG((gsrc || extra) && raw && stats && gamma && beta && gy && sums, "inorm_bwd_reduce: null pointer");
  FNST_CHECK_ARG(gsrc_slack >= 0, "inorm_bwd_reduce: negative gsrc_slack");
  FNST_CHECK_ARG(c % 8 == 0 && c <= 1024 && 256 % (c / 8) == 0, "inorm_bwd_reduce: unsupported channel count %d", c);
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!prezeroed) {
    FNST_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * (size_t)n * c, st));
    if (dgb) FNST_CUDA(cudaMemsetAsync(dgb, 0, sizeof(float) * 2 * (size_t)c, st));
  }
  HaloLayout L{h, w, c, pad, pad_mode == FNST_PAD_REFLECT ? 1 : 0, s2d, gsrc_slack};
  if (!gsrc) L.pad = 0;
  {
    // TMA-staged form: 2-byte tensors, plain halo layout, the padded row fits one box, two CTAs' tiles fit one SM
    const int wp = w + 2 * L.pad;
    const size_t tile_bytes = 2 * ((size_t)w * c + (gsrc ? (size_t)wp * c : 0) + (extra ? (size_t)w * c : 0));
    const size_t scratch = sizeof(float) * (size_t)(256 / (c / 8)) * 2 * c;
    const size_t smem_tma = 20 * (size_t)c + 128 + (tile_bytes > scratch ? tile_bytes : scratch);
    if (tuning().inorm_bwd_tma && act_dtype != FNST_F32 && g_dtype != FNST_F32 && !s2d && c <= 256 && wp <= 256 && h <= 65535 &&
        smem_tma <= 110 * 1024 && (int64_t)n * h * w < ((int64_t)1 << 31)) {
      CUtensorMap m_raw, m_g, m_e;
      memset(&m_g, 0, sizeof(m_g)); memset(&m_e, 0, sizeof(m_e));
      {
        const uint64_t dims[2] = {(uint64_t)c, (uint64_t)n * h * w};
        const uint64_t str[1] = {(uint64_t)c * 2};
        const uint32_t box[2] = {(uint32_t)c, (uint32_t)w};
        if (int r = encode_tensor_map_plain(&m_raw, raw, 2, dims, str, box)) return r;
        if (extra) { if (int r = encode_tensor_map_plain(&m_e, extra, 2, dims, str, box)) return r; }
      }
      if (gsrc) {
        const uint64_t wa = wp + gsrc_slack, ha = h + 2 * L.pad + gsrc_slack;
        const uint64_t dims[4] = {(uint64_t)c, wa, ha, (uint64_t)n};
        const uint64_t str[3] = {(uint64_t)c * 2, (uint64_t)c * 2 * wa, (uint64_t)c * 2 * wa * ha};
        const uint32_t box[4] = {(uint32_t)c, (uint32_t)wp, 1, 1};
        if (int r = encode_tensor_map_plain(&m_g, gsrc, 4, dims, str, box)) return r;
      }
      FNST_DISPATCH_DTYPE(act_dtype, TA, {
        FNST_DISPATCH_DTYPE(g_dtype, TG, {
          if (sizeof(TA) == 2 && sizeof(TG) == 2) {
            auto kern = inorm_bwd_reduce_tma_kernel<TA, TG>;
            FNST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            launch_pdl(kern, dim3(h, n), dim3(256), smem_tma, st, m_raw, m_g, m_e, reinterpret_cast<const TG*>(gsrc), stats, gamma, beta,
                       drop, reinterpret_cast<TG*>(gy), sums, dgb, L, relu, eps, extra ? 1 : 0);
          }
        });
      });
      return launch_status("inorm_bwd_reduce (TMA)");
    }
  }
  const int rpb = rows_per_block(h, n);
  dim3 grid((h + rpb - 1) / rpb, n);
  const size_t smem = sizeof(float) * (5 * (size_t)c + (size_t)(256 / (c / 8)) * 2 * c);
  FNST_DISPATCH_DTYPE(act_dtype, TA, {
    FNST_DISPATCH_DTYPE(g_dtype, TG, {
      // two resident blocks per SM for the 16-bit variants (tuning knob inorm_bwd_blocks = 1 restores the unconstrained build)
      auto kern = (sizeof(TA) + sizeof(TG) <= 4 && tuning().inorm_bwd_blocks >= 2) ? inorm_bwd_reduce_kernel<TA, TG, 2>
                                                                                    : inorm_bwd_reduce_kernel<TA, TG, 1>;
      launch_pdl(kern, dim3(grid), dim3(256), smem, st,
          reinterpret_cast<const TG*>(gsrc), reinterpret_cast<const TG*>(extra), reinterpret_cast<const TA*>(raw), stats, gamma,
          beta, drop, reinterpret_cast<TG*>(gy), sums, dgb, L, relu, eps, rpb);
    });
  });
  return launch_status("inorm_bwd_reduce");
}

extern "C" int fnst_inorm_bwd_apply(const void* gy, const void* raw, const float* stats, const float* sums, const float* gamma,
                                    void* draw, float* dgb, int n, int h, int w, int c, int act_dtype, int g_dtype, float eps,
                                    int out_s2d, int out_pad, int device, void* stream) {
  FNST_CHECK_ARG(gy && raw && stats && sums && gamma && draw, "inorm_bwd_apply: null pointer");
  FNST_CHECK_ARG(out_pad >= 0 && !(out_pad && out_s2d), "inorm_bwd_apply: out_pad is for the plain NHWC output only");
  FNST_CHECK_ARG(c % 8 == 0 && c <= 2048 && 256 % (c / 8) == 0, "inorm_bwd_apply: unsupported channel count %d", c);
  FNST_CHECK_ARG(!out_s2d || (h % 2 == 0 && w % 2 == 0), "inorm_bwd_apply: space-to-depth output needs even h, w");
  FNST_DEVICE(device);
  const int rpb = rows_per_block(h, n);
  dim3 grid((h + rpb - 1) / rpb, n);
  FNST_DISPATCH_DTYPE(act_dtype, TA, {
    FNST_DISPATCH_DTYPE(g_dtype, TG, {
      launch_pdl(inorm_bwd_apply_kernel<TA, TG>, dim3(grid), dim3(256), sizeof(float) * 5 * c, (cudaStream_t)stream,
          reinterpret_cast<const TG*>(gy), reinterpret_cast<const TA*>(raw), stats, sums, gamma, reinterpret_cast<TG*>(draw),
          h, w, c, eps, out_s2d, dgb, rpb, out_pad);
    });
  });
  return launch_status("inorm_bwd_apply");
}

// Geometry of the fused kernel.  chunk_rows x w <= 256 pixels per TMA box.  For each slab width cw (64 / 32 / 16 channels) the
// smallest cluster K <= 8 whose parts (raw + gsrc + extra tiles) fit ~200 KB of shared memory; among the widths the one with
// the most CTAs (capped at 128), ties to the widest (longest contiguous runs for TMA and DRAM).  K = 0: unsupported (w > 256 or
// the plane does not fit) -> use fnst_inorm_bwd_reduce + fnst_inorm_bwd_apply.
struct IbfGeometry { int K, rows_per_part, chunk_rows, chunks, chunk_stride, cw; size_t smem; };
static IbfGeometry ibf_geometry(int n, int h, int w, int c, int act_dtype, int g_dtype, int tiles_g) {
  IbfGeometry best{0, 0, 0, 0, 0, 0, 0};
  if (w > 256 || w <= 0 || h <= 0 || n <= 0 || c % 16 != 0) return best;
  const int chunk_rows = 256 / w < h ? 256 / w : h;
  const int chunk_stride = (chunk_rows * w + 7) / 8 * 8;               // chunk starts stay 128-byte aligned in shared memory
  const size_t budget = 200 * 1024;
  long best_ctas = 0;
  for (int cw = 64; cw >= 16; cw /= 2) {
    if (c % cw != 0) continue;
    const size_t per_slot = (size_t)cw * (dtype_size(act_dtype) + (size_t)tiles_g * dtype_size(g_dtype));
    for (int k = 1; k <= 8; k *= 2) {
      int rows = (h + k - 1) / k;
      rows = (rows + chunk_rows - 1) / chunk_rows * chunk_rows;        // whole chunks
      const int chunks = rows / chunk_rows;
      const size_t bytes = (size_t)chunks * chunk_stride * per_slot;
      if (bytes > budget || chunks > IBF_MAX_CHUNKS) continue;
      long ctas = (long)k * (c / cw) * n;
      if (ctas > 128) ctas = 128;
      // ties: a cluster of 8 x ~200 KB places only one or two clusters per GPC (measured: two waves for 16 clusters), so
      // prefer K <= 4; otherwise the wider slab (visited first)
      if (ctas > best_ctas || (ctas == best_ctas && best.K > 4 && k <= 4)) {
        best = IbfGeometry{k, rows, chunk_rows, chunks, chunk_stride, cw, bytes + 128};
        best_ctas = ctas;
      }
      break;                                                           // smallest cluster that fits this width
    }
  }
  return best;
}

// tiled, un-swizzled tensor map over a tensor of 2-byte units (dims innermost first)
static int encode_tensor_map_plain(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                   const uint32_t* box) {
  PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
  FNST_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  FNST_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bdim[5]; cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    FNST_CHECK_ARG(strides_bytes[i] % 16 == 0, "TMA stride %d (%llu B) must be a multiple of 16", i, (unsigned long long)strides_bytes[i]);
    gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNST_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

extern "C" int fnst_inorm_bwd_fused_parts(int n, int h, int w, int c, int act_dtype, int g_dtype, int has_gsrc, int has_extra, int s2d) {
  if (c % 16 != 0 || s2d || (!has_gsrc && !has_extra) || n <= 0) return 0;
  return ibf_geometry(n, h, w, c, act_dtype, g_dtype, (has_gsrc ? 1 : 0) + (has_extra ? 1 : 0)).K;
}

extern "C" int fnst_inorm_bwd_fused(const void* gsrc, const void* extra, const void* raw, const float* stats,
                                    const float* gamma, const float* beta, const float* drop, void* draw, void* gy_out,
                                    float* sums, int n, int h, int w, int c, int act_dtype, int g_dtype, int relu, float eps,
                                    int pad, int pad_mode, int s2d, int out_s2d, int device, void* stream) {
  FNST_CHECK_ARG((gsrc || extra) && raw && stats && gamma && beta && draw && sums, "inorm_bwd_fused: null pointer");
  FNST_CHECK_ARG(c % 16 == 0, "inorm_bwd_fused: channel count %d must be a multiple of 16", c);
  FNST_CHECK_ARG(!s2d, "inorm_bwd_fused: a space-to-depth gradient buffer needs the two-pass operators");
  FNST_CHECK_ARG(!out_s2d || (h % 2 == 0 && w % 2 == 0), "inorm_bwd_fused: space-to-depth output needs even h, w");
  FNST_CHECK_ARG(act_dtype != FNST_F32 || g_dtype == FNST_F32, "inorm_bwd_fused: fp32 activations need fp32 gradients");
  const IbfGeometry G = ibf_geometry(n, h, w, c, act_dtype, g_dtype, (gsrc ? 1 : 0) + (extra ? 1 : 0));
  FNST_CHECK_ARG(G.K > 0, "inorm_bwd_fused: a %dx%d plane does not fit the shared memory of 8 CTAs (use the two-pass operators)", h, w);
  FNST_DEVICE(device);
  IbfParams P;
  P.L = HaloLayout{h, w, c, gsrc ? pad : 0, (gsrc && pad_mode == FNST_PAD_REFLECT) ? 1 : 0, 0};
  P.relu = relu; P.out_s2d = out_s2d; P.rows_per_part = G.rows_per_part; P.K = G.K; P.chunk_rows = G.chunk_rows; P.chunks = G.chunks;
  P.chunk_stride = G.chunk_stride; P.cw = G.cw; P.tpp_log2 = G.cw == 64 ? 3 : (G.cw == 32 ? 2 : 1);
  P.eps = eps;
  P.dbg = tuning().debug_buf;
  const uint64_t ea = dtype_size(act_dtype) / 2, eg = dtype_size(g_dtype) / 2;      // 2-byte units per element
  CUtensorMap m_raw, m_g, m_e;
  memset(&m_g, 0, sizeof(m_g)); memset(&m_e, 0, sizeof(m_e));
  {
    const uint64_t dims[4] = {(uint64_t)c * ea, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)c * ea * 2, (uint64_t)c * ea * 2 * w, (uint64_t)c * ea * 2 * w * h};
    const uint32_t box[4] = {(uint32_t)(G.cw * ea), (uint32_t)w, (uint32_t)G.chunk_rows, 1};
    if (int r = encode_tensor_map_plain(&m_raw, raw, 4, dims, str, box)) return r;
  }
  if (gsrc) {
    const uint64_t wp = w + 2 * P.L.pad, hp = h + 2 * P.L.pad;
    const uint64_t dims[4] = {(uint64_t)c * eg, wp, hp, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)c * eg * 2, (uint64_t)c * eg * 2 * wp, (uint64_t)c * eg * 2 * wp * hp};
    const uint32_t box[4] = {(uint32_t)(G.cw * eg), (uint32_t)w, (uint32_t)G.chunk_rows, 1};
    if (int r = encode_tensor_map_plain(&m_g, gsrc, 4, dims, str, box)) return r;
  }
  if (extra) {
    const uint64_t dims[4] = {(uint64_t)c * eg, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)c * eg * 2, (uint64_t)c * eg * 2 * w, (uint64_t)c * eg * 2 * w * h};
    const uint32_t box[4] = {(uint32_t)(G.cw * eg), (uint32_t)w, (uint32_t)G.chunk_rows, 1};
    if (int r = encode_tensor_map_plain(&m_e, extra, 4, dims, str, box)) return r;
  }
  dim3 grid(G.K, c / G.cw, n);
  FNST_DISPATCH_DTYPE(act_dtype, TA, {
    FNST_DISPATCH_DTYPE(g_dtype, TG, {
      auto kern = inorm_bwd_fused_kernel<TA, TG>;
      FNST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
      launch_pdl_cluster(kern, grid, dim3(IBF_THREADS), G.smem, (cudaStream_t)stream, G.K, m_raw, m_g, m_e,
          reinterpret_cast<const TG*>(gsrc), extra ? 1 : 0, stats, gamma, beta, drop, reinterpret_cast<TG*>(draw),
          reinterpret_cast<TG*>(gy_out), sums, P);
    });
  });
  return launch_status("inorm_bwd_fused");
}

extern "C" int fnst_affine_grads(const float* sums, const void* table, int layers, int n, int max_c, float* out, int device, void* stream) {
  FNST_CHECK_ARG(sums && table && out && layers > 0 && n > 0 && max_c > 0, "affine_grads: bad arguments");
  FNST_DEVICE(device);
  dim3 grid(layers, (max_c + 255) / 256);
  launch_pdl(affine_grads_kernel, grid, dim3(256), 0, (cudaStream_t)stream, sums, reinterpret_cast<const AffineGradEntry*>(table), n, out);
  return launch_status("affine_grads");
}

extern "C" int fnst_maxpool2_bwd(const void* in, const void* gout, const void* extra, void* gin, int n, int h, int w, int c,
                                 int act_dtype, int g_dtype, int device, void* stream) {
  FNST_CHECK_ARG(in && gout && gin && c % 8 == 0 && h >= 2 && w >= 2, "maxpool2_bwd: bad arguments");
  FNST_DEVICE(device);
  const int64_t items = (int64_t)n * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8);
  FNST_DISPATCH_DTYPE(act_dtype, TA, {
    FNST_DISPATCH_DTYPE(g_dtype, TG, {
      launch_pdl(maxpool2_bwd_kernel<TA, TG>, dim3(grid_cap(items)), dim3(256), 0, (cudaStream_t)stream, 
          reinterpret_cast<const TA*>(in), reinterpret_cast<const TG*>(gout), reinterpret_cast<const TG*>(extra),
          reinterpret_cast<TG*>(gin), n, h, w, c);
    });
  });
  return launch_status("maxpool2_bwd");
}

extern "C" int fnst_sse_bwd(const void* a, const void* b, int64_t count, int64_t b_period, int dtype_a, int dtype_b,
                            const float* scale, float coef, void* da, int g_dtype, int relu_mask, int device, void* stream) {
  FNST_CHECK_ARG(a && b && scale && da && count > 0 && b_period > 0, "sse_bwd: bad arguments");
  FNST_DEVICE(device);
  FNST_DISPATCH_DTYPE(dtype_a, TA, {
    FNST_DISPATCH_DTYPE(dtype_b, TB, {
      FNST_DISPATCH_DTYPE(g_dtype, TG, {
        launch_pdl(sse_bwd_kernel<TA, TB, TG>, dim3(grid_cap(count / 4)), dim3(256), 0, (cudaStream_t)stream, 
            reinterpret_cast<const TA*>(a), reinterpret_cast<const TB*>(b), count, b_period, scale, coef, reinterpret_cast<TG*>(da), relu_mask);
      });
    });
  });
  return launch_status("sse_bwd");
}

extern "C" int fnst_relu_mask(const void* g, const void* extra, const void* act, void* out, int64_t count, int act_dtype,
                              int g_dtype, int device, void* stream) {
  FNST_CHECK_ARG(g && act && out && count > 0 && count % 8 == 0, "relu_mask: bad arguments (count must be a multiple of 8)");
  FNST_DEVICE(device);
  FNST_DISPATCH_DTYPE(act_dtype, TA, {
    FNST_DISPATCH_DTYPE(g_dtype, TG, {
      launch_pdl(relu_mask_kernel<TA, TG>, dim3(grid_cap(count / 8)), dim3(256), 0, (cudaStream_t)stream, 
          reinterpret_cast<const TG*>(g), reinterpret_cast<const TG*>(extra), reinterpret_cast<const TA*>(act),
          reinterpret_cast<TG*>(out), count / 8);
    });
  });
  return launch_status("relu_mask");
}

extern "C" int fnst_gram_diff_sym(const float* g, const float* gt, int n, int c, int64_t gt_numel, const float* scale, float coef,
                                  void* s_out, int out_dtype, int device, void* stream) {
  FNST_CHECK_ARG(g && gt && scale && s_out && n > 0 && c > 0 && gt_numel > 0, "gram_diff_sym: bad arguments");
  FNST_DEVICE(device);
  const int64_t total = (int64_t)n * c * c;
  FNST_DISPATCH_DTYPE(out_dtype, TG, {
    launch_pdl(gram_diff_sym_kernel<TG>, dim3(grid_cap(total)), dim3(256), 0, (cudaStream_t)stream, g, gt, gt_numel, c, total, scale, coef,
                                                                              reinterpret_cast<TG*>(s_out));
  });
  return launch_status("gram_diff_sym");
}

extern "C" int fnst_tv_bwd(const float* img, int planes, int h, int w, const float* scale, float coef, float* dimg, int device,
                           void* stream) {
  FNST_CHECK_ARG(img && scale && dimg && planes > 0 && h > 0 && w > 0, "tv_bwd: bad arguments");
  FNST_DEVICE(device);
  launch_pdl(tv_bwd_kernel, dim3(grid_cap((int64_t)planes * h * w / 4)), dim3(256), 0, (cudaStream_t)stream, img, planes, h, w, scale, coef, dimg);
  return launch_status("tv_bwd");
}

extern "C" int fnst_channel_sum(const float* x, int n, int c, int hw, float* out, int device, void* stream) {
  FNST_CHECK_ARG(x && out && n > 0 && c > 0 && hw > 0, "channel_sum: bad arguments");
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  FNST_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * c, st));
  int chunks = (int)(((int64_t)n * hw + 256 * 16 - 1) / (256 * 16));
  if (chunks > 512) chunks = 512;
  dim3 grid(chunks, c);
  launch_pdl(channel_sum_kernel, dim3(grid), dim3(256), 0, st, x, n, c, hw, out);
  return launch_status("channel_sum");
}
