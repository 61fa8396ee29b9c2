// 3x3 convolution with 64 input channels as a ROW-STREAMING tensor-core kernel (sm_100a): VGG conv1_2 / conv2_1
// (torchvision features[2], [5]; models/vgg19_net.py:38-51) forward, and the data gradient of conv1_2 (train.py:200).
//
// Why a second kernel.  The generic gather-GEMM (conv_tc_kernel) re-loads the activation tile once per tap (9x the unique
// bytes through TMA into shared memory) and re-loads the weights per k-block.  With few output channels that is fatal: an
// M128 x N64 x K16 MMA takes ~32 tensor clocks but reads 6 KB of operands, and the TMA writes of a k-block add as much again
// -- shared memory is oversubscribed ~3x, and conv1_2 (19.3 GFLOP at 4 x 256 x 256) ran at 380 TFLOP/s.  Here
//   * the packed weights of all nine taps stay RESIDENT in shared memory (72 KB for 64 outputs, 144 KB for 128): nine TMA
//     box loads per CTA, [tap][output j][64 channels = one 128-byte swizzled row];
//   * a CTA owns runs of output rows of one 128-column strip and streams the input rows through a ring ONCE: a row lands
//     as 130 pixels x 128 bytes (K-major, 128-byte swizzle), and the operand of horizontal tap kw is the SAME buffer with the
//     descriptor's start address advanced by kw pixel rows (kw x 128 B; the swizzle is a function of the absolute
//     shared-memory address, so a row-shifted start needs nothing else) -- no im2col, no re-load; the three vertical taps are the three newest ring
//     entries.  Zero padding = TMA out-of-bounds fill (rows and columns).  (First version: un-swizzled core-matrix rows
//     filled by TMA boxes with a 16-byte inner extent -- measured ~5 clocks per 16-byte piece, 3 us per input row: TMA-bound.)
//   * one output row = 36 MMAs M128 (pixels) x N x K16 into one TMEM slot of a ring of 512/N slots (N = 64: TWO output rows per
//     128-column slot -- the two middle rows of the four-row input window feed both of them with one N = 128 MMA against the
//     stacked weight blocks, 48 MMAs per row pair instead of 72); two groups of four
//     epilogue warps take alternate output rows (tcgen05.ld -> bias / ReLU or the data-gradient form (acc + addend) * (mask > 0)
//     -> 16-byte NHWC stores), so the latency of the mask / addend loads of one row hides under the next row's MMAs.
// Work split: the (image, strip, output row) units are cut into equal CONTIGUOUS ranges, one per SM -- 2048 row units on 148
// SMs is 13.8 per CTA, not 14 whole tiles on some and 13 on others plus a ragged wave.
//   warp 0: TMA producer      warp 1: MMA issuer      warps 2..9: epilogue
#include "tc_common.cuh"

namespace fnst {

constexpr int RC_STRIP = 128;                         // output columns per strip (= MMA M)
constexpr int RC_PW = RC_STRIP + 2;                   // input pixels per row segment
constexpr int RC_ROW_BYTES = RC_PW * 128;             // 16 640 bytes land per input row (64 channels = 128 B per pixel)
constexpr int RC_ROW_STRIDE = 17 * 1024;              // ring entries start on the 1024-byte swizzle pattern
constexpr int RC_THREADS = 320;
constexpr int RC_EPI_GROUPS = 2;

template <int N> struct RcCfg {
  static constexpr int W_BYTES = 9 * N * 128;                     // [tap][j][64 channels], 128-byte swizzled rows
  static constexpr int RING = N == 64 ? 8 : 4;                    // input rows resident in shared memory
  // N = 64: a TMEM slot holds TWO output rows (columns [0,64) = row y, [64,128) = row y+1): the two middle input rows of the
  // four-row window feed both of them with ONE N = 128 MMA against the stacked weight blocks [W(kh) ; W(kh-1)] -- 48 MMAs per
  // row pair instead of 72, and an M128 x K16 MMA costs ~64-76 clocks of A-operand read whatever its N (header comment).
  static constexpr bool PAIRED = N == 64;
  static constexpr int SLOT_COLS = PAIRED ? 128 : N;
  static constexpr int SLOTS = 512 / SLOT_COLS;                   // TMEM accumulator slots (units in flight)
  static constexpr int SMEM = W_BYTES + RING * RC_ROW_STRIDE + 1024 + 512;
};

struct RowConvParams {
  int32_t n_img, in_h, in_w, out_h, out_w;
  int32_t strips, units;                 // units = n_img * strips * out_h
  int32_t base_h, base_w;                // input coordinate of tap (kh, kw) = (0, 0) relative to the output pixel
  int32_t relu, out_is_bf16, mask_dtype, dbg_mode;
  uint32_t idesc, idesc_wide;            // MMA instruction descriptors for N and (paired rows) 2N columns
  int8_t tap_of[9];                      // descriptor tap index of kernel position kh * 3 + kw
  void* out;
  const float* bias;
  const void* addend;
  const void* mask;
  unsigned long long* dbg;               // measurement only (fnst_set_debug_buffer): per-row timeline of CTA 0
};

// K-major 128-byte-swizzled operand whose first row is NOT on the 1024-byte swizzle pattern boundary (start address advanced by
// whole 128-byte rows).  Measured on B200 (tools/dbg_rowconv.py): the tensor core applies the swizzle XOR to the ABSOLUTE
// shared-memory address bits, exactly as TMA did when it wrote the tile, so the plain descriptor with the shifted start address
// reads rows kw, kw+1, ... bit-exactly; setting the descriptor's base-offset field (bits 49..51) to (addr >> 7) & 7 as well
// makes every shifted tap wrong.  The ring entries themselves must start on the 1024-byte pattern.
__device__ __forceinline__ uint64_t umma_desc_sw128_rowshift(uint32_t smem_addr) { return umma_smem_desc(smem_addr, 16, 1024); }

template <int N>
__global__ void __launch_bounds__(RC_THREADS, 1)
rowconv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ RowConvParams p) {
  using Cfg = RcCfg<N>;
  constexpr int RING = Cfg::RING, SLOTS = Cfg::SLOTS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                                  // resident weights
  uint8_t* s_in = smem + Cfg::W_BYTES;                  // input row ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_in + RING * RC_ROW_STRIDE);
  uint64_t* loaded = bars;                              // [RING]   TMA bytes of an input row have landed        (producer -> MMA)
  uint64_t* freed = bars + RING;                        // [RING]   last MMA reading the row has completed      (MMA -> producer)
  uint64_t* acc_full = bars + 2 * RING;                 // [SLOTS]  all 36 MMAs of an output row have completed (MMA -> epilogue)
  uint64_t* acc_empty = bars + 2 * RING + SLOTS;        // [SLOTS]  the epilogue has read the slot             (epilogue -> MMA)
  uint64_t* w_bar = bars + 2 * RING + 2 * SLOTS;        // weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  __shared__ float s_bias[N];                           // broadcast reads in the epilogue instead of N global loads per pixel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Paired rows: who stores what.  Data-gradient form (addend / mask loads: the epilogue is latency-bound on them) -- BOTH warp
  // groups work on every unit, group g on its row g, so the unit's two rows proceed concurrently (measured 61 vs 66 us);
  // plain form -- the groups take alternate units (23 vs 25 us).
  const bool split_rows = Cfg::PAIRED && (p.addend != nullptr || p.mask != nullptr);

  pdl_trigger();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < RING; ++s) { mbar_init(&loaded[s], 1); mbar_init(&freed[s], 1); }
    for (int s = 0; s < SLOTS; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], split_rows ? 8 : 4); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // measurement only: CTA 0 stamps (globaltimer ns) [g] TMA issue of input row g, [64+o] operands of output row o landed,
  // [128+o] its MMAs issued, [192+o] its accumulator complete (seen by the epilogue), [256+o] epilogue of the row done
  unsigned long long* tl = (p.dbg && blockIdx.x == 0) ? p.dbg : nullptr;
  auto stamp = [&](int slot) {
    if (tl && slot < 320) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tl[slot] = t; }
  };
  // this CTA's contiguous range of (image, strip, output row) units; a segment = a run of rows inside one strip
  const int u_begin = (int)((int64_t)p.units * blockIdx.x / gridDim.x);
  const int u_end = (int)((int64_t)p.units * (blockIdx.x + 1) / gridDim.x);
  auto segment = [&](int u, int& n, int& x0, int& y0, int& rows) {
    const int col = u / p.out_h;
    y0 = u - col * p.out_h;
    n = col / p.strips;
    x0 = (col - n * p.strips) * RC_STRIP;
    rows = min(u_end - u, p.out_h - y0);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && u_begin < u_end) {
      mbar_arrive_expect_tx(w_bar, Cfg::W_BYTES);
      // {64 k, N outputs} per tap, stored [kw][kh = 2, 1, 0]: the blocks of two vertically adjacent taps are consecutive rows of
      // one wider B operand (paired rows)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
          tma_load_2d(s_w + (kw * 3 + (2 - kh)) * (N * 128), &map_b, w_bar, p.tap_of[kh * 3 + kw] * 64, 0);
      uint32_t g = 0;                                    // running input-row count = ring position
      for (int u = u_begin; u < u_end;) {
        int n, x0, y0, rows;
        segment(u, n, x0, y0, rows);
        for (int r = 0; r < rows + 2; ++r, ++g) {
          const uint32_t i = g % RING;
          mbar_wait_spin(&freed[i], ((g / RING) & 1) ^ 1);
          mbar_arrive_expect_tx(&loaded[i], RC_ROW_BYTES);
          stamp(g < 64 ? (int)g : 1000);
          // box {64 ch, 130 px, 1 row, 1 image}; rows / pixels outside the image read as zero
          tma_load_4d(s_in + i * RC_ROW_STRIDE, &map_a, &loaded[i], 0, x0 + p.base_w, y0 + p.base_h + r, n);
        }
        u += rows;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && u_begin < u_end) {
      uint64_t da_kw[3];                                 // ring entry 0 seen through horizontal tap kw (start row kw of the pattern)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) da_kw[kw] = umma_desc_sw128_rowshift(smem_u32(s_in) + kw * 128);
      const uint64_t db_base = umma_smem_desc(smem_u32(s_w), 16, 1024);
      constexpr uint32_t BLK = N * 128 >> 4;             // (addr >> 4) size of one weight block [N outputs][64 channels]
      mbar_wait(w_bar, 0);
      tc_fence_after();
      uint32_t g0 = 0, waited = 0, o = 0;                // first input row of the segment; input rows already waited for; unit count
      for (int u = u_begin; u < u_end;) {
        int n, x0, y0, rows;
        segment(u, n, x0, y0, rows);
        for (int j = 0; j < rows; ++o) {
          const int nr = (Cfg::PAIRED && j + 1 < rows) ? 2 : 1;            // output rows of this unit
          const uint32_t s = o % SLOTS;
          mbar_wait_spin(&acc_empty[s], ((o / SLOTS) & 1) ^ 1);
          while (waited <= g0 + j + nr + 1) { mbar_wait_spin(&loaded[waited % RING], (waited / RING) & 1); ++waited; }
          tc_fence_after();
          stamp(o < 64 ? 64 + (int)o : 1000);
          const uint32_t d_tmem = tmem_base + s * Cfg::SLOT_COLS;
          auto row_desc = [&](int r, int kw) { return da_kw[kw] + (uint64_t)(((g0 + j + r) % RING) * (RC_ROW_STRIDE >> 4)); };
          if (nr == 2) {
            // window rows r = 0..3; r = 1, 2 feed both output rows (N = 2 x 64 against [W(r) ; W(r-1)]), r = 0 only the first
            // (W0 into columns [0,64)), r = 3 only the second (W2 into columns [64,128)).  The first MMA issued must cover all
            // 128 columns with accumulate = 0, so the wide rows go first.
#pragma unroll
            for (int r = 1; r <= 2; ++r)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const uint64_t da = row_desc(r, kw), db = db_base + (uint64_t)((kw * 3 + (2 - r)) * BLK);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_f16(d_tmem, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), p.idesc_wide, (r != 1 || kw != 0 || ks != 0) ? 1u : 0u);
              }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint64_t da0 = row_desc(0, kw), db0 = db_base + (uint64_t)((kw * 3 + 2) * BLK);      // W0
              const uint64_t da3 = row_desc(3, kw), db3 = db_base + (uint64_t)((kw * 3 + 0) * BLK);      // W2
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma_f16(d_tmem, da0 + (uint64_t)(2 * ks), db0 + (uint64_t)(2 * ks), p.idesc, 1u);
                umma_f16(d_tmem + N, da3 + (uint64_t)(2 * ks), db3 + (uint64_t)(2 * ks), p.idesc, 1u);
              }
            }
          } else {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const uint64_t da = row_desc(kh, kw), db = db_base + (uint64_t)((kw * 3 + (2 - kh)) * BLK);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)          // 16 K-elements = 32 bytes further inside the 128-byte row: +2 in the (addr >> 4) field
                  umma_f16(d_tmem, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), p.idesc, (kh | kw | ks) != 0 ? 1u : 0u);
              }
          }
          umma_commit(&acc_full[s]);
          stamp(o < 64 ? 128 + (int)o : 1000);
          // (the input rows this unit read for the last time are handed back to the producer by the epilogue, which observes
          // the same completion through acc_full: one tcgen05.commit per unit on this thread)
          j += nr;
        }
        g0 += rows + 2;
        u += rows;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: group (warp - 2) / 4 takes every second output row =====================
    const int q = warp & 3, grp = (warp - 2) >> 2;
    if (threadIdx.x - 64 < N) s_bias[threadIdx.x - 64] = p.bias ? p.bias[threadIdx.x - 64] : 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");        // the eight epilogue warps only
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    int rows_prev = 0;
    uint32_t o = 0, g0 = 0;
    for (int u = u_begin; u < u_end; g0 += rows_prev + 2) {
      int n, x0, y0, rows;
      segment(u, n, x0, y0, rows);
      rows_prev = rows;
      const int x = x0 + q * 32 + lane;
      const bool valid = x < p.out_w;
      for (int j = 0; j < rows; ++o) {
        const int nr = (Cfg::PAIRED && j + 1 < rows) ? 2 : 1;              // output rows of this unit (same walk as the MMA issuer)
        const int jj = j;
        j += nr;
        if (!split_rows && (int)(o % RC_EPI_GROUPS) != grp) continue;      // alternate units per warp group (see split_rows)
        const uint32_t s = o % SLOTS;
        // Data-gradient form: this thread's 128-byte addend and mask rows are requested BEFORE the wait for the accumulator, so
        // their memory latency (the epilogue's bound: 134 MB per launch against 20 us of MMA work) runs under the unit's MMAs
        uint4 pre_a[N == 64 ? 8 : 1], pre_m[N == 64 ? 8 : 1];
        const bool prefetch = N == 64 && split_rows && grp < nr && valid && !(p.dbg_mode & 8) && (!p.mask || p.mask_dtype != FNST_F32);
        if (N == 64 && prefetch) {
          const size_t off = (((size_t)n * p.out_h + (y0 + jj + grp)) * p.out_w + x) * N;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (p.addend) pre_a[i] = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.addend) + off)[i];
            if (p.mask) pre_m[i] = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.mask) + off)[i];
          }
        }
        mbar_wait_spin(&acc_full[s], (o / SLOTS) & 1);
        tc_fence_after();
        if (split_rows && grp >= nr) {                     // single-row unit: the second group only releases the slot
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[s]);
          continue;
        }
        if (q == 0 && lane == 0 && (!split_rows || grp == 0)) {
          // all MMAs up to this unit have completed: its first nr input rows are dead, and so are the two rows below them
          // once the segment ends
          for (int r = 0; r < nr; ++r) mbar_arrive(&freed[(g0 + jj + r) % RING]);
          if (jj + nr == rows) { mbar_arrive(&freed[(g0 + rows) % RING]); mbar_arrive(&freed[(g0 + rows + 1) % RING]); }
        }
        if (threadIdx.x == 64 || threadIdx.x == 192) stamp(o < 64 ? 192 + (int)o : 1000);
#pragma unroll
        for (int half = 0; half < (Cfg::PAIRED ? 2 : 1); ++half) {
          if (half >= nr || (split_rows && half != grp)) continue;
          const size_t off = (((size_t)n * p.out_h + (y0 + jj + half)) * p.out_w + x) * N;
#pragma unroll
          for (int cb = 0; cb < N; cb += 32) {
            uint32_t raw[32];
            tmem_ld_x32(t_lane + s * Cfg::SLOT_COLS + half * N + cb, raw);
            tmem_ld_wait();
            if ((split_rows || half == nr - 1) && cb + 32 >= N) {       // this warp has read all it needs: hand the slot back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&acc_empty[s]);
            }
            if (!valid || (p.dbg_mode & 8)) continue;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += s_bias[cb + i];
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (N == 64 && prefetch) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                float t[8];
                if (p.addend) {
                  if (p.out_is_bf16) { Raw8<__nv_bfloat16> r; r.u = pre_a[(cb + i) >> 3]; raw8_to_f32<__nv_bfloat16>(r, t); }
                  else { Raw8<__half> r; r.u = pre_a[(cb + i) >> 3]; raw8_to_f32<__half>(r, t); }
#pragma unroll
                  for (int k = 0; k < 8; ++k) v[i + k] += t[k];
                }
                if (p.mask) {
                  if (p.mask_dtype == FNST_BF16) { Raw8<__nv_bfloat16> r; r.u = pre_m[(cb + i) >> 3]; raw8_to_f32<__nv_bfloat16>(r, t); }
                  else { Raw8<__half> r; r.u = pre_m[(cb + i) >> 3]; raw8_to_f32<__half>(r, t); }
#pragma unroll
                  for (int k = 0; k < 8; ++k) v[i + k] = t[k] > 0.f ? v[i + k] : 0.f;
                }
              }
            } else if (p.addend || p.mask) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                float t[8];
                if (p.addend) {
                  load8_dyn(p.addend, p.out_is_bf16 ? FNST_BF16 : FNST_F16, off + cb + i, t);
#pragma unroll
                  for (int k = 0; k < 8; ++k) v[i + k] += t[k];
                }
                if (p.mask) {
                  load8_dyn(p.mask, p.mask_dtype, off + cb + i, t);
#pragma unroll
                  for (int k = 0; k < 8; ++k) v[i + k] = t[k] > 0.f ? v[i + k] : 0.f;
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              float t[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) t[k] = v[i + k];
              if (p.out_is_bf16) store8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + off + cb + i, t);
              else store8<__half>(reinterpret_cast<__half*>(p.out) + off + cb + i, t);
            }
          }
        }
        if (threadIdx.x == 64 || threadIdx.x == 192) stamp(o < 64 ? 256 + (int)o : 1000);
      }
      u += rows;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// 1: the descriptor is a full 3x3 tap grid over a 64-channel NHWC view with 64 or 128 outputs and a plain NHWC 16-bit
// epilogue (bias / ReLU / addend / mask, no statistics) -- the row-streaming kernel computes it.
bool rowconv_eligible(const fnst_conv_desc* d) {
  if (!tuning().conv_rowstream) return false;
  if (d->ntaps != 9 || d->kc != 64 || d->a_c != 64 || d->a_stride_w != 64) return false;
  if (!(d->n_gemm == 64 || d->n_gemm == 128) || d->c_out != d->n_gemm || d->epilogue != FNST_EPI_NHWC) return false;
  if (d->stats || d->b_image_rows || (d->flags & FNST_DESC_LINEAR)) return false;
  if (!(d->out_dtype == FNST_F16 || d->out_dtype == FNST_BF16)) return false;
  if (d->a_stride_h % 8 || d->a_stride_n % 8) return false;
  int mh = 127, mw = 127;
  for (int t = 0; t < 9; ++t) {
    if (d->tap_c0[t] != 0) return false;
    mh = d->tap_dh[t] < mh ? d->tap_dh[t] : mh;
    mw = d->tap_dw[t] < mw ? d->tap_dw[t] : mw;
  }
  int seen = 0;
  for (int t = 0; t < 9; ++t) {
    const int kh = d->tap_dh[t] - mh, kw = d->tap_dw[t] - mw;
    if (kh > 2 || kw > 2) return false;
    seen |= 1 << (kh * 3 + kw);
  }
  return seen == 0x1FF;
}

template <int N>
static int launch_rowconv(const CUtensorMap& ma, const CUtensorMap& mb, const RowConvParams& p, int sms, cudaStream_t st) {
  auto kern = rowconv_tc_kernel<N>;
  FNST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RcCfg<N>::SMEM));
  const int grid = p.units < sms ? p.units : sms;
  launch_pdl(kern, dim3(grid), dim3(RC_THREADS), RcCfg<N>::SMEM, st, ma, mb, p);
  return launch_status("rowconv_tc");
}

int rowconv_tc(const fnst_conv_desc* d, int device, cudaStream_t st) {
  RowConvParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = d->out_n; p.in_h = d->a_h; p.in_w = d->a_w; p.out_h = d->out_h; p.out_w = d->out_w;
  p.strips = (d->out_w + RC_STRIP - 1) / RC_STRIP;
  const int64_t units = (int64_t)p.n_img * p.strips * p.out_h;
  FNST_CHECK_ARG(units < (int64_t)1 << 30, "rowconv_tc: too many row units");
  p.units = (int32_t)units;
  int mh = 127, mw = 127;
  for (int t = 0; t < 9; ++t) { mh = d->tap_dh[t] < mh ? d->tap_dh[t] : mh; mw = d->tap_dw[t] < mw ? d->tap_dw[t] : mw; }
  p.base_h = d->h0 + mh; p.base_w = d->w0 + mw;
  for (int t = 0; t < 9; ++t) p.tap_of[(d->tap_dh[t] - mh) * 3 + (d->tap_dw[t] - mw)] = (int8_t)t;
  p.dbg_mode = tuning().dbg_mode;
  p.dbg = tuning().debug_buf;
  p.relu = d->relu; p.out_is_bf16 = d->out_dtype == FNST_BF16; p.mask_dtype = d->mask_dtype;
  p.idesc = umma_idesc_f16(d->dtype == FNST_BF16 ? 1 : 0, d->n_gemm, 0, 0);
  p.idesc_wide = umma_idesc_f16(d->dtype == FNST_BF16 ? 1 : 0, 2 * d->n_gemm, 0, 0);
  p.out = d->out; p.bias = d->bias; p.addend = d->addend; p.mask = d->mask;

  CUtensorMap ma, mb;
  {
    const uint64_t dims[4] = {64, (uint64_t)d->a_w, (uint64_t)d->a_h, (uint64_t)d->a_n};
    const uint64_t str[3] = {(uint64_t)d->a_stride_w * 2, (uint64_t)d->a_stride_h * 2, (uint64_t)d->a_stride_n * 2};
    const uint32_t box[4] = {64, (uint32_t)RC_PW, 1, 1};
    if (int r = encode_tensor_map_2b(&ma, d->a, 4, dims, str, box)) return r;
  }
  {
    const uint64_t dims[2] = {9 * 64, (uint64_t)d->n_gemm};
    const uint64_t str[1] = {9 * 64 * 2};
    const uint32_t box[2] = {64, (uint32_t)d->n_gemm};
    if (int r = encode_tensor_map_2b(&mb, d->b, 2, dims, str, box)) return r;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return d->n_gemm == 64 ? launch_rowconv<64>(ma, mb, p, sms, st) : launch_rowconv<128>(ma, mb, p, sms, st);
}

}  // namespace fnst
