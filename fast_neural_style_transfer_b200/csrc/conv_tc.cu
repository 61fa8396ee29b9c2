// Gather-GEMM convolution on 5th-generation tensor cores (sm_100a).
//
//   out[pixel, j] = sum_{tap t} sum_{c < kc} A[n, h+h0+dh[t], w+w0+dw[t], c0[t]+c] * B[j][t*kc + c]
//
// One persistent CTA per SM, 192 or 320 threads, warp-specialised:
//   warp 0      TMA producer: per k-block (one tap x 64 channels) one 4-D box load of the
//               activation tile {64 ch, TW, TH, 1} (implicit im2col: a spatial box shifted by the
//               tap offset; out-of-range coordinates are zero-filled by TMA = zero padding) and
//               one 2-D box load {64 k, BLOCK_N} of packed weights, 128-byte swizzled, into a
//               STAGES-deep mbarrier ring.
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128 (pixels) x N=BLOCK_N x K=16,
//               fp32 accumulators in TMEM, double-buffered (2 x BLOCK_N columns) so the epilogue
//               of tile i overlaps the main loop of tile i+1.
//   warps 2..   epilogue: tcgen05.ld (32 lanes x 32 columns), bias / ReLU, per-(image,channel)
//               sum and sum-of-squares for InstanceNorm (smem transposition -> per-warp partials ->
//               one global atomic per column per tile), vector stores as NHWC / depth-to-space /
//               NCHW fp32.  Four warps (one per TMEM lane quarter) for column tiles < 128, eight
//               (two per quarter, half of the columns each) for the wide tiles: with one tile per
//               CTA (batch <= 4) the epilogue is exposed, and it is latency-bound per warp.
//
// PAIR = true: the same kernel as a CTA pair (cluster of 2 on the two SMs of a TPC, tcgen05 cta_group::2): one
// M=256 x N=BLOCK_N MMA spans two adjacent pixel tiles; each CTA stages its own 128 A rows and only HALF of the B
// (weight) rows, so per SM the MMA reads 4 KB (A) + BLOCK_N/2 x 32 B (B) of shared memory per K=16 step instead of
// 4 KB + BLOCK_N x 32 B, and TMA writes a third less -- the single-CTA N=256 form is shared-memory-bandwidth bound
// (96 B/clk of operand reads + 96 B/clk of TMA writes against 128 B/clk).  The even CTA issues all MMAs; TMA bytes of
// both CTAs are credited to its full barrier; tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals
// to both CTAs; each CTA runs the epilogue of its own 128 TMEM lanes.
#include "tc_common.cuh"

#include <cstdlib>
#include <mutex>

namespace fnst {

PFN_tensorMapEncodeTiled tensor_map_encoder() {
  static PFN_tensorMapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(p);
  });
  return fn;
}

int encode_tensor_map_2b(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box) {
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
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNST_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

struct ConvTcParams {
  int32_t out_n, out_h, out_w;
  int32_t tiles_w, tiles_h, tw_log2, h_step;     // h_step: rows between successive tiles (TH, or 8 for ROWSUM9)
  int32_t num_m_tiles, num_n_tiles, num_tiles;
  int32_t num_kblocks, chunks_per_tap;
  int32_t h0, w0;
  int32_t epilogue, c_out, n_gemm, relu, out_is_bf16, out_is_f32, b_image_rows;
  int32_t stage_out;             // 1: epilogue stages the tile in shared memory (swizzled) and stores it with TMA (one tile per CTA)
  int32_t lin_pitch;             // > 0: pixel-stream form (FNST_DESC_LINEAR): tap (dh, dw) = pixel offset dh * lin_pitch + dw
  uint32_t idesc;
  int32_t out_dtype, mask_dtype;
  void* out;
  const float* bias;
  float* stats;
  const void* addend;
  const void* mask;
  unsigned long long* dbg;       // measurement only (fnst_set_debug_buffer)
  int32_t dbg_mode;
  int8_t tap_dh[FNST_MAX_TAPS];
  int8_t tap_dw[FNST_MAX_TAPS];
  int16_t tap_c0[FNST_MAX_TAPS];
};

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;                       // 64 two-byte elements = one 128-byte swizzle row
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;

template <int BLOCK_N, bool PAIR = false> struct TcCfg {
  static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;     // weight rows staged by this CTA
  static constexpr int B_BYTES = B_ROWS * TC_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  // column tiles of 64 / 128: a PRIVATE staging tile for the shared-memory + TMA-store epilogue (multi-tile launches cannot
  // borrow the operand ring: the next tile is loading into it), paid for with one operand stage
  static constexpr int STAGING_BYTES = (BLOCK_N == 64 || BLOCK_N == 128) ? TC_BLOCK_M * BLOCK_N * 2 : 0;
  static constexpr int STAGES_RAW = (200 * 1024 - STAGING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int CHUNK = BLOCK_N < 32 ? BLOCK_N : 32;     // epilogue column chunk
  static constexpr int EPI_WARPS = BLOCK_N >= 128 ? 8 : 4;      // epilogue warps: 1 or 2 per TMEM lane quarter
  static constexpr int EPI_THREADS = 32 * EPI_WARPS;
  static constexpr int THREADS = 64 + EPI_THREADS;              // + TMA producer warp + MMA warp
  static constexpr int COLS_PER_WARP = BLOCK_N / (EPI_WARPS / 4);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <typename TOut>
__device__ __forceinline__ void store_chunk(TOut* p, const float (&v)[32], int count) {
  // count in {16, 32}; p is 32-byte aligned
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    if (i < count) {
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = v[i + k];
      store8<TOut>(p + i, t);
    }
  }
}

template <int BLOCK_N, bool PAIR>
__global__ void __launch_bounds__((TcCfg<BLOCK_N, PAIR>::THREADS), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_out, const __grid_constant__ ConvTcParams p) {
  using Cfg = TcCfg<BLOCK_N, PAIR>;
  static_assert(!PAIR || BLOCK_N >= 64, "CTA pairs need at least 32 weight rows per CTA");
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CHUNK = Cfg::CHUNK;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  __shared__ float s_sum[4 * BLOCK_N], s_sq[4 * BLOCK_N];     // per epilogue warp: column sums of its 32 rows (plain stores, no atomics)
  // 128 x 33 floats of scratch: ROWSUM9 staging (128 T-pixels x 27 partials), or four per-warp 32 x 33 transposition
  // tiles for the InstanceNorm column sums (the two uses never occur in the same launch)
  __shared__ float s_rows[TC_BLOCK_M * 33];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PAIR: rank of this CTA inside its pair, work-unit stride and first unit (unit = pair of adjacent pixel tiles x column tile)
  const int rank = PAIR ? (int)cluster_cta_rank() : 0;
  const bool leader = rank == 0;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_units = PAIR ? ((p.num_m_tiles + 1) >> 1) * p.num_n_tiles : p.num_tiles;

  pdl_trigger();
  // measurement only (fnst_set_debug_buffer): per-CTA timeline in globaltimer nanoseconds behind the 4 x 148 main-loop slots:
  // [0] kernel entry, [1] set-up done (after griddepcontrol.wait), [2] first operand stage landed, [3] last MMA complete,
  // [4] epilogue done (first epilogue warp), [5] after the final CTA barrier
  unsigned long long* tl = p.dbg ? p.dbg + 4 * 148 + 8 * (PAIR ? (blockIdx.x >> 1) : blockIdx.x) : nullptr;
  auto stamp = [&](int slot) {
    if (tl) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tl[slot] = t; }
  };
  if (threadIdx.x == 0) stamp(0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    // tmem_empty: 4 epilogue warps per CTA; in a pair both CTAs' warps arrive on the leader's barrier
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], (PAIR ? 2 : 1) * Cfg::EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) { if (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot); else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();      // pair: the partner's barriers must exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // set-up above overlapped the predecessor's tail; global memory is touched only from here on
  if (threadIdx.x == 0) stamp(1);

  const int TW = 1 << p.tw_log2, TH = TC_BLOCK_M >> p.tw_log2;
  // unit -> (column tile, pixel tile of this CTA); a pair's phantom second tile (odd tile count) has m_tile == num_m_tiles:
  // its image index is out of range, so TMA zero-fills the loads and the epilogue stores nothing
  auto decode = [&](int unit, int& n_tile, int& m_tile) {
    n_tile = unit % p.num_n_tiles;
    m_tile = unit / p.num_n_tiles;
    if (PAIR) m_tile = 2 * m_tile + rank;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = unit0; unit < num_units; unit += unit_step) {
        int n_tile, m_tile;
        decode(unit, n_tile, m_tile);
        const int tw = m_tile % p.tiles_w; m_tile /= p.tiles_w;
        const int th = m_tile % p.tiles_h;
        const int n = m_tile / p.tiles_h;
        const int hb = th * p.h_step + p.h0, wb = tw * TW + p.w0;
        // weight rows of this CTA: the whole column tile, or its half of it in a pair
        const int brow = n * p.b_image_rows + n_tile * BLOCK_N + rank * Cfg::B_ROWS;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          const int t = kb / p.chunks_per_tap, ch = kb - t * p.chunks_per_tap;
          // pixel-stream form: the tile is 128 consecutive pixels of a linear stream and a tap is a linear shift
          const int cw = wb + p.tap_dw[t] + p.tap_dh[t] * p.lin_pitch;
          const int chh = p.lin_pitch ? 0 : hb + p.tap_dh[t];
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + TC_A_BYTES;
          if (PAIR) {
            // one arrival + the bytes of BOTH CTAs on the leader's barrier (the partner's loads are credited to it)
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_4d_pair(sa, &map_a, &full_bar[stage], p.tap_c0[t] + ch * TC_BLOCK_K, cw, chh, n);
            tma_load_2d_pair(sb, &map_b, &full_bar[stage], kb * TC_BLOCK_K, brow);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_4d(sa, &map_a, &full_bar[stage], p.tap_c0[t] + ch * TC_BLOCK_K, cw, chh, n);
            tma_load_2d(sb, &map_b, &full_bar[stage], kb * TC_BLOCK_K, brow);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && leader) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      long long dbg_c0 = 0; unsigned long long dbg_t0 = 0;
      for (int unit = unit0; unit < num_units; unit += unit_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (p.dbg && it == 0 && kb == 0) {
            dbg_c0 = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
          }
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = umma_smem_desc(sa, 16, 1024);
          const uint64_t db = umma_smem_desc(sa + TC_A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
            // advance 16 K-elements = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
            if (PAIR) umma_f16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (PAIR) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma_commit_pair(&tmem_full[as]); else umma_commit(&tmem_full[as]);
      }
      if (p.dbg && it > 0) {
        const int as = (it - 1) & 1;
        mbar_wait(&tmem_full[as], ((it - 1) >> 1) & 1);           // last accumulator complete
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        unsigned long long* o = p.dbg + 4 * (PAIR ? (blockIdx.x >> 1) : blockIdx.x);
        o[0] = (unsigned long long)(clock64() - dbg_c0); o[1] = t1 - dbg_t0;
        o[2] = (unsigned long long)it * p.num_kblocks; o[3] = (unsigned long long)it;
        tl[2] = dbg_t0; tl[3] = t1;
      }
      pdl_trigger_tail();      // all MMAs of this CTA are issued: let the next kernel launch under the epilogue
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2.. = EPI_THREADS threads) =====================
    constexpr int EPI_THREADS = Cfg::EPI_THREADS;
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int ew = warp - 2;                    // epilogue warp index; ew >> 2 selects the column half (wide tiles)
    const int et = threadIdx.x - 64;            // 0..EPI_THREADS-1
    const int col_begin = (ew >> 2) * Cfg::COLS_PER_WARP, col_end = col_begin + Cfg::COLS_PER_WARP;
    int it = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step, ++it) {
      int n_tile, m_tile;
      decode(unit, n_tile, m_tile);
      const bool tile_live = m_tile < p.num_m_tiles;          // false only for a pair's phantom second tile
      const int tw = m_tile % p.tiles_w; m_tile /= p.tiles_w;
      const int th = m_tile % p.tiles_h;
      const int n = m_tile / p.tiles_h;
      const int r = q * 32 + lane;
      const int h = th * p.h_step + (r >> p.tw_log2), w = tw * TW + (r & (TW - 1));
      const bool valid = tile_live && h < p.out_h && w < p.out_w;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N);

      if (BLOCK_N >= 64 && p.stage_out) {
        // ---- staged epilogue (one tile per CTA: the operand ring is idle once the accumulator is complete) ----------------
        // TMEM -> registers -> 16-bit tile in shared memory, [BLOCK_N/64 column blocks][128 pixel rows][128 B] in the 128-byte
        // swizzle TMA expects -> (a) InstanceNorm column sums with ONE THREAD PER COLUMN walking the 128 staged rows (no
        // transposition, no shuffles; statistics of exactly the values that are stored), (b) TMA tile stores (full 128-byte
        // lines, ragged edges clipped by the tensor map) instead of 16-byte stores at a 512-byte stride per lane.
        uint8_t* stage = Cfg::STAGING_BYTES ? smem + STAGES * Cfg::STAGE_BYTES : smem;     // private tile, or the idle operand ring (N = 256)
#pragma unroll 1
        for (int cb = col_begin; cb < col_end; cb += 32) {
          uint32_t raw[32];
          tmem_ld_x32(t_row + cb, raw);
          tmem_ld_wait();
          const int col0 = n_tile * BLOCK_N + cb;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = valid ? __uint_as_float(raw[i]) : 0.f;      // rows outside the image: zeros (statistics)
          if (p.bias && valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += p.bias[col0 + i];
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          uint8_t* rowp = stage + (cb >> 6) * (TC_BLOCK_M * 128) + r * 128;
          const int j0 = (cb & 63) >> 3;                       // first 16-byte chunk of these 32 columns inside the 128-byte row
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 u;
            if (p.out_is_bf16) {
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[8 * k + 2 * e], v[8 * k + 2 * e + 1]);
            } else {
              __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(v[8 * k + 2 * e], v[8 * k + 2 * e + 1]);
            }
            *reinterpret_cast<uint4*>(rowp + (((j0 + k) ^ (r & 7)) << 4)) = u;
          }
        }
        // generic-proxy writes -> visible to the async proxy (TMA), then all epilogue warps meet
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        if (Cfg::STAGING_BYTES) {
          // every warp has read its part of the accumulator: hand the TMEM stage back now, so that in a multi-tile launch the
          // MMAs of the tile after next start while this tile is still being stored
          tc_fence_before();
          if (lane == 0) { if (PAIR) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]); }
        }
        if (et == 0 && tile_live) {
#pragma unroll 1
          for (int b = 0; b < BLOCK_N / 64; ++b) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&map_out)), "r"(smem_u32(stage + b * (TC_BLOCK_M * 128))),
                           "r"(n_tile * BLOCK_N + b * 64), "r"(tw * TW), "r"(th * p.h_step), "r"(n) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (p.stats && tile_live) {
          for (int c = et; c < BLOCK_N; c += EPI_THREADS) {
            const uint8_t* colp = stage + (c >> 6) * (TC_BLOCK_M * 128) + ((c & 7) << 1);
            const int j = (c & 63) >> 3;
            float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
            for (int rr = 0; rr < TC_BLOCK_M; ++rr) {
              const unsigned short bits = *reinterpret_cast<const unsigned short*>(colp + rr * 128 + ((j ^ (rr & 7)) << 4));
              const float x = p.out_is_bf16 ? __uint_as_float((uint32_t)bits << 16) : __half2float(__ushort_as_half(bits));
              s1[rr & 3] += x;
              s2[rr & 3] = fmaf(x, x, s2[rr & 3]);
            }
            const int col = n_tile * BLOCK_N + c;
            atomicAdd(&p.stats[((size_t)n * p.c_out + col) * 2 + 0], (s1[0] + s1[1]) + (s1[2] + s1[3]));
            atomicAdd(&p.stats[((size_t)n * p.c_out + col) * 2 + 1], (s2[0] + s2[1]) + (s2[2] + s2[3]));
          }
        }
        if (Cfg::STAGING_BYTES) {
          // the staging tile is written again by the next tile of this CTA: nobody may pass before TMA has read it
          if (et == 0 && tile_live) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
          continue;
        }
        if (et == 0 && tile_live) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        // accumulator fully read long ago: hand the TMEM stage back (keeps the barrier protocol of the generic path)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]); }
        continue;
      }

#pragma unroll 1
      for (int cb = col_begin; cb < col_end; cb += CHUNK) {
        uint32_t raw[32];
        if (CHUNK == 32) tmem_ld_x32(t_row + cb, raw); else tmem_ld_x16(t_row + cb, raw);
        tmem_ld_wait();
        if (cb + CHUNK >= col_end) {
          // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]); }
        }
        const int col0 = n_tile * BLOCK_N + cb;           // first GEMM column of this chunk
        if (col0 >= p.n_gemm || !tile_live) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = i < CHUNK ? __uint_as_float(raw[i]) : 0.f;

        if (BLOCK_N == 32 && p.epilogue == FNST_EPI_ROWSUM9) {
          // stage the 27 horizontal partials of this T-pixel, then sum 9 rows per output pixel
#pragma unroll
          for (int i = 0; i < 27; ++i) s_rows[r * 33 + i] = v[i];
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
          const int co = p.c_out;
          const int npix = p.h_step * TW;                       // output pixels of this tile: h_step rows x TW columns
          for (int idx = et; idx < npix * co; idx += EPI_THREADS) {
            const int o = idx / npix, pp = idx - o * npix;
            const int oh = th * p.h_step + (pp >> p.tw_log2), ow = tw * TW + (pp & (TW - 1));
            if (oh < p.out_h && ow < p.out_w) {
              float acc = p.bias ? p.bias[o] : 0.f;
#pragma unroll
              for (int kh = 0; kh < 9; ++kh) acc += s_rows[(pp + TW * kh) * 33 + kh * co + o];
              reinterpret_cast<float*>(p.out)[(((size_t)n * co + o) * p.out_h + oh) * p.out_w + ow] = acc;
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
          continue;
        }

        if (p.epilogue == FNST_EPI_NCHW_F32) {
          if (valid) {
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ch = col0 + j;
              if (ch < p.c_out) {
                float y = v[j] + (p.bias ? p.bias[ch] : 0.f);
                o[(((size_t)n * p.c_out + ch) * p.out_h + h) * p.out_w + w] = y;
              }
            }
          }
          continue;
        }

        int ch0 = col0, phase_id = 0;
        if (p.epilogue == FNST_EPI_D2S) { phase_id = col0 / p.c_out; ch0 = col0 - phase_id * p.c_out; }
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < CHUNK; ++i) v[i] += p.bias[ch0 + i];
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < CHUNK; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (valid) {
          size_t pix;
          if (p.epilogue == FNST_EPI_D2S)
            pix = ((size_t)n * (2 * p.out_h) + 2 * h + (phase_id >> 1)) * (size_t)(2 * p.out_w) + 2 * w + (phase_id & 1);
          else
            pix = ((size_t)n * p.out_h + h) * (size_t)p.out_w + w;
          const size_t off = pix * p.c_out + ch0;
          if (p.addend || p.mask) {
#pragma unroll
            for (int i = 0; i < CHUNK; i += 8) {
              float t[8];
              if (p.addend) {
                load8_dyn(p.addend, p.out_dtype, off + i, t);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[i + k] += t[k];
              }
              if (p.mask) {
                load8_dyn(p.mask, p.mask_dtype, off + i, t);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[i + k] = t[k] > 0.f ? v[i + k] : 0.f;
              }
            }
          }
          if (p.out_is_f32) store_chunk<float>(reinterpret_cast<float*>(p.out) + off, v, CHUNK);
          else if (p.out_is_bf16) store_chunk<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + off, v, CHUNK);
          else store_chunk<__half>(reinterpret_cast<__half*>(p.out) + off, v, CHUNK);
        }
        if (p.stats && !(p.dbg_mode & 1)) {
          // column sums over this warp's 32 rows through a padded 16-column smem transposition (conflict-free both
          // ways): lane l sums column (l & 15) over rows 16*(l >> 4) .. +15 with four independent partial sums, one
          // shuffle joins the two row halves
          float* tr = s_rows + ew * (16 * 33);
#pragma unroll
          for (int hh = 0; hh < CHUNK; hh += 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) tr[i * 33 + lane] = valid ? v[hh + i] : 0.f;
            __syncwarp();
            const float* col = tr + (lane & 15) * 33 + (lane >> 4) * 16;
            float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float x = col[j]; cs[j & 3] += x; cq[j & 3] = fmaf(x, x, cq[j & 3]); }
            float s1 = (cs[0] + cs[1]) + (cs[2] + cs[3]), s2 = (cq[0] + cq[1]) + (cq[2] + cq[3]);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
            if (lane < 16) { s_sum[q * BLOCK_N + cb + hh + lane] = s1; s_sq[q * BLOCK_N + cb + hh + lane] = s2; }
            __syncwarp();
          }
        }
      }

      if (p.stats) {
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        for (int c = et; c < BLOCK_N; c += EPI_THREADS) {
          const int col = n_tile * BLOCK_N + c;
          if (col < p.n_gemm && tile_live && !(p.dbg_mode & 2)) {
            const int ch = p.epilogue == FNST_EPI_D2S ? col % p.c_out : col;
            if (ch < p.c_out) {
              atomicAdd(&p.stats[((size_t)n * p.c_out + ch) * 2 + 0], s_sum[c] + s_sum[BLOCK_N + c] + s_sum[2 * BLOCK_N + c] + s_sum[3 * BLOCK_N + c]);
              atomicAdd(&p.stats[((size_t)n * p.c_out + ch) * 2 + 1], s_sq[c] + s_sq[BLOCK_N + c] + s_sq[2 * BLOCK_N + c] + s_sq[3 * BLOCK_N + c]);
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      }
    }
  }

  if (Cfg::STAGING_BYTES && threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // TMA stores of the last tiles
  if (threadIdx.x == 64) stamp(4);
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();     // pair: neither CTA may leave while the other still signals / reads it
  if (threadIdx.x == 0) stamp(5);
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base); else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

int validate_conv_desc(const fnst_conv_desc* d);
bool rowconv_eligible(const fnst_conv_desc* d);                        // rowconv_tc.cu
int rowconv_tc(const fnst_conv_desc* d, int device, cudaStream_t st);

template <int BLOCK_N, bool PAIR>
static int launch_conv_tc(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const ConvTcParams& p, int num_sms,
                          cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N, PAIR>;
  auto kern = conv_tc_kernel<BLOCK_N, PAIR>;
  FNST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  if (PAIR) {
    // persistent CTA pairs: one cluster of 2 per TPC; a unit = two adjacent pixel tiles x one column tile
    const int units = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    const int pairs = units < num_sms / 2 ? units : num_sms / 2;
    launch_pdl_cluster(kern, dim3(2 * pairs), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, 2, ma, mb, mo, p);
    return launch_status("conv_tc (CTA pair)");
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  launch_pdl(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, ma, mb, mo, p);
  return launch_status("conv_tc");
}

static int device_sm_count(int device) {
  static int cached[64] = {0};
  if (device < 0 || device >= 64) return 148;
  if (!cached[device]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
    cached[device] = v;
  }
  return cached[device];
}

}  // namespace fnst

using namespace fnst;

extern "C" int fnst_device_supports_tc(int device) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int fnst_conv_tc(const fnst_conv_desc* d, int device, void* stream) {
  if (int r = validate_conv_desc(d)) return r;
  FNST_CHECK_ARG(d->dtype == FNST_F16 || d->dtype == FNST_BF16, "conv_tc: operands must be fp16 or bf16");
  FNST_CHECK_ARG(d->kc % TC_BLOCK_K == 0, "conv_tc: kc %d must be a multiple of 64", d->kc);
  FNST_CHECK_ARG(d->a_stride_w % 8 == 0 && d->a_stride_h % 8 == 0 && d->a_stride_n % 8 == 0, "conv_tc: A strides must be multiples of 8 elements");
  FNST_CHECK_ARG(d->a_c % 8 == 0, "conv_tc: a_c must be a multiple of 8");
  if (d->epilogue == FNST_EPI_D2S) FNST_CHECK_ARG(d->c_out % 32 == 0, "conv_tc: d2s needs c_out %% 32 == 0");
  if (d->epilogue == FNST_EPI_NHWC) FNST_CHECK_ARG(d->c_out == d->n_gemm && (d->n_gemm == 16 || d->n_gemm % 32 == 0), "conv_tc: NHWC epilogue needs c_out == n_gemm, a multiple of 32 (or 16)");
  if (d->epilogue == FNST_EPI_NCHW_F32) FNST_CHECK_ARG(d->c_out <= 16 && !d->stats && !d->relu, "conv_tc: NCHW epilogue supports c_out <= 16, no stats/relu");
  const bool rowsum = d->epilogue == FNST_EPI_ROWSUM9;
  const bool linear = (d->flags & FNST_DESC_LINEAR) != 0;
  if (linear) {
    // pixel-stream form: one row of a_w pixels; tap (dh, dw) is the linear pixel shift dh * pitch + dw, pitch = a_stride_h / a_stride_w
    FNST_CHECK_ARG(d->a_h == 1 && d->a_n == 1 && d->out_h == 1 && d->out_n == 1, "conv_tc: the pixel-stream form needs a_h == a_n == out_h == out_n == 1");
    FNST_CHECK_ARG(d->epilogue == FNST_EPI_NHWC && !d->stats && d->a_stride_w > 0 && d->a_stride_h % d->a_stride_w == 0,
                   "conv_tc: the pixel-stream form needs the NHWC epilogue, no statistics and a row stride that is a multiple of the pixel stride");
  }
  if (rowsum) {
    FNST_CHECK_ARG(d->n_gemm == 32 && 9 * d->c_out <= 27 && !d->stats && !d->relu, "conv_tc: ROWSUM9 needs n_gemm == 32, c_out <= 3, no stats/relu");
    for (int t = 0; t < d->ntaps; ++t) FNST_CHECK_ARG(d->tap_dh[t] == 0, "conv_tc: ROWSUM9 taps must have dh == 0");
  }
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  // 3x3 over 64 input channels with few outputs: the row-streaming kernel (resident weights, each input row staged once)
  if (rowconv_eligible(d)) return rowconv_tc(d, device, st);

  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.out_n = d->out_n; p.out_h = d->out_h; p.out_w = d->out_w;
  // pixel tile TH x TW = 128: wide tiles for wide images, never wider than needed
  // ROWSUM9: tall 4 x 32 T tiles advance by 24 output rows (8-row halo below): 75 % of the loaded rows are new
  const int tw_log2 = linear ? 7 : rowsum ? 2 : (d->out_w <= 8 ? 3 : 4);
  p.tw_log2 = tw_log2;
  const int TW = 1 << tw_log2, TH = TC_BLOCK_M >> tw_log2;
  p.h_step = rowsum ? TH - 8 : TH;
  p.tiles_w = (d->out_w + TW - 1) / TW;
  p.tiles_h = (d->out_h + p.h_step - 1) / p.h_step;
  p.num_m_tiles = p.tiles_w * p.tiles_h * d->out_n;
  const int num_sms = device_sm_count(device);

  // column tile: the widest one that still puts a tile on at least ~60 % of the SMs (measured at batch 1..16 of the
  // 3x3 256->256 conv: N=256 beats N=128 as soon as it yields >= ~90 tiles; N=64 is shared-memory-bandwidth bound)
  int block_n;
  const int64_t fill = (int64_t)num_sms * 6 / 10;
  if (d->n_gemm <= 16) block_n = 16;
  else if (d->n_gemm <= 32) block_n = 32;
  else if (d->n_gemm <= 64 || d->n_gemm % 128 != 0) block_n = 64;
  else if (d->n_gemm % 256 == 0 && (int64_t)p.num_m_tiles * (d->n_gemm / 256) >= fill) block_n = 256;
  else if ((int64_t)p.num_m_tiles * (d->n_gemm / 128) >= fill) block_n = 128;
  else block_n = 64;
  {
    const int force_n = tuning().conv_block_n;
    if (force_n && d->n_gemm % force_n == 0 && d->n_gemm >= 64) block_n = force_n;     // tuning override
  }
  FNST_CHECK_ARG(d->n_gemm % block_n == 0 || d->n_gemm < block_n, "conv_tc: n_gemm %d not tileable", d->n_gemm);
  p.num_n_tiles = (d->n_gemm + block_n - 1) / block_n;
  p.num_tiles = p.num_m_tiles * p.num_n_tiles;
  p.chunks_per_tap = d->kc / TC_BLOCK_K;
  p.num_kblocks = d->ntaps * p.chunks_per_tap;
  p.h0 = d->h0; p.w0 = d->w0;
  p.lin_pitch = linear ? (int32_t)(d->a_stride_h / d->a_stride_w) : 0;
  p.epilogue = d->epilogue; p.c_out = d->c_out; p.n_gemm = d->n_gemm; p.relu = d->relu;
  p.out_is_bf16 = d->out_dtype == FNST_BF16; p.out_is_f32 = d->out_dtype == FNST_F32;
  // CTA pairs (M = 256 MMAs across two SMs) for the wide column tiles.  Measured (tools/exp_conv_fixed_cost.py): ~7 % less
  // time per k-block but ~1 us more fixed cost per launch, so only when every pair gets at least two units of work.
  const int pair_mode = tuning().conv_pair;
  const int64_t units_pair = (int64_t)((p.num_m_tiles + 1) / 2) * ((d->n_gemm + block_n - 1) / block_n);   // 74 pairs on 148 SMs
  const bool pair = block_n >= 128 && p.num_m_tiles >= 2 && (pair_mode == 2 || (pair_mode == 1 && units_pair >= num_sms));
  p.dbg = tuning().debug_buf;
  p.dbg_mode = tuning().dbg_mode;
  p.idesc = umma_idesc_f16(d->dtype == FNST_BF16 ? 1 : 0, block_n, 0, 0, pair ? 256 : 128);
  p.out = d->out; p.bias = d->bias; p.stats = d->stats;
  p.out_dtype = d->out_dtype; p.mask_dtype = d->mask_dtype; p.b_image_rows = d->b_image_rows;
  p.addend = d->epilogue == FNST_EPI_NHWC ? d->addend : nullptr;
  p.mask = d->epilogue == FNST_EPI_NHWC ? d->mask : nullptr;
  memcpy(p.tap_dh, d->tap_dh, sizeof(p.tap_dh));
  memcpy(p.tap_dw, d->tap_dw, sizeof(p.tap_dw));
  memcpy(p.tap_c0, d->tap_c0, sizeof(p.tap_c0));

  CUtensorMap ma, mb;
  {
    const uint64_t dims[4] = {(uint64_t)d->a_c, (uint64_t)d->a_w, (uint64_t)d->a_h, (uint64_t)d->a_n};
    uint64_t str[3] = {(uint64_t)d->a_stride_w * 2, (uint64_t)d->a_stride_h * 2, (uint64_t)d->a_stride_n * 2};
    if (linear) str[1] = str[2] = (uint64_t)d->a_stride_w * 2 * (uint64_t)d->a_w;       // extent-1 dimensions: any legal stride
    const uint32_t box[4] = {TC_BLOCK_K, (uint32_t)TW, (uint32_t)TH, 1};
    if (int r = encode_tensor_map_2b(&ma, d->a, 4, dims, str, box)) return r;
  }
  {
    const uint64_t ktot = (uint64_t)d->ntaps * d->kc;
    const uint64_t dims[2] = {ktot, (uint64_t)(d->b_image_rows ? (int64_t)d->b_image_rows * d->out_n : d->n_gemm)};
    const uint64_t str[1] = {ktot * 2};
    const uint32_t box[2] = {TC_BLOCK_K, (uint32_t)(pair ? block_n / 2 : block_n)};     // a pair's CTAs load half the rows each
    if (int r = encode_tensor_map_2b(&mb, d->b, 2, dims, str, box)) return r;
  }
  // Staged epilogue (shared-memory tile + TMA store + per-column statistics): plain NHWC 16-bit outputs of full 64-column
  // blocks, one tile per CTA (the operand ring doubles as the staging buffer, so no next tile may be loading into it).
  CUtensorMap mo;
  memset(&mo, 0, sizeof(mo));
  p.stage_out = 0;
  // (column tiles of 64 / 128 own a private staging tile: any tile count, CTA pairs included)
  const bool private_stage = block_n == 64 || block_n == 128;
  if (tuning().conv_stage_out && d->epilogue == FNST_EPI_NHWC && !p.out_is_f32 && block_n >= 64 && d->c_out == d->n_gemm &&
      d->n_gemm % 64 == 0 && !p.addend && !p.mask && !p.dbg_mode &&
      ((private_stage && tuning().conv_stage_out == 1) || (!pair && p.num_tiles <= num_sms))) {
    const uint64_t dims[4] = {(uint64_t)d->c_out, (uint64_t)d->out_w, (uint64_t)d->out_h, (uint64_t)d->out_n};
    const uint64_t str[3] = {(uint64_t)d->c_out * 2, (uint64_t)d->c_out * 2 * d->out_w, (uint64_t)d->c_out * 2 * d->out_w * d->out_h};
    const uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    if (int r = encode_tensor_map_2b(&mo, d->out, 4, dims, str, box)) return r;
    p.stage_out = 1;
  }
  if (d->stats && !(d->flags & FNST_DESC_PREZEROED))
    FNST_CUDA(cudaMemsetAsync(d->stats, 0, sizeof(float) * 2 * (size_t)d->out_n * d->c_out, st));
  switch (block_n) {
    case 16: return launch_conv_tc<16, false>(ma, mb, mo, p, num_sms, st);
    case 32: return launch_conv_tc<32, false>(ma, mb, mo, p, num_sms, st);
    case 64: return launch_conv_tc<64, false>(ma, mb, mo, p, num_sms, st);
    case 128: return pair ? launch_conv_tc<128, true>(ma, mb, mo, p, num_sms, st) : launch_conv_tc<128, false>(ma, mb, mo, p, num_sms, st);
    default: return pair ? launch_conv_tc<256, true>(ma, mb, mo, p, num_sms, st) : launch_conv_tc<256, false>(ma, mb, mo, p, num_sms, st);
  }
}
