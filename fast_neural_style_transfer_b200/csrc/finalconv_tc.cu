// final_conv (32 -> 3 channels, 9x9, models/model.py:47) as a row-streaming tensor-core kernel.
//
//   T[y][x][kh*3+o] = sum_kw sum_c  in[y][x+kw][c] * W[o][c][kh][kw]        (one GEMM row block per input row y)
//   out[o][yo][x]   = bias[o] + sum_kh T[yo+kh][x][kh*3+o]                    (9-row sum in the epilogue)
//
// The generic gather-GEMM form of this layer (conv_tc_kernel, ROWSUM9 epilogue) is bound by L2 -> SM traffic: every
// tap re-loads its shifted activation window, ~13x the unique bytes.  Here a CTA owns a vertical strip of 128 output
// columns and streams the input rows through shared memory ONCE: a row is stored un-swizzled as
// [4 channel chunks][136 pixels][8 channels = 16 B], the canonical K-major no-swizzle core-matrix layout (8 pixels x
// 16 B contiguous), so the operand of tap kw is the same buffer with its start address advanced by kw x 16 B -- no
// re-load, no im2col.  Per input row: one TMA box load (8.7 KB, 16-stage ring), 18 MMAs M128 x N32 x K16 (9 taps x 2
// channel halves) into one 32-column TMEM slot of a 16-slot ring; the epilogue warps read the 9 slots yo..yo+8 with tcgen05.ld.x4 and
// write the three output planes (NCHW fp32, coalesced along x).
//   warp 0: TMA producer      warp 1: MMA issuer      warps 2..5: epilogue (one per TMEM lane quarter)
#include "tc_common.cuh"

namespace fnst {

constexpr int FC_STRIP = 128;                   // output columns per CTA strip (= MMA M)
constexpr int FC_PW = FC_STRIP + 8;             // input pixels per row segment
constexpr int FC_ROW_BYTES = 4 * FC_PW * 16;    // [4 chunks][136 px][16 B]
constexpr int FC_RING = 16;                     // rows in flight: smem stages = TMEM slots (16 x 32 columns = all of TMEM)
constexpr int FC_W_BYTES = 9 * 2 * 1024;        // packed weights: [kw][channel half][2 k-chunks][32 rows][16 B]
constexpr int FC_THREADS = 192;
constexpr int FC_SMEM = FC_RING * FC_ROW_BYTES + FC_W_BYTES + 1024 + 512;

struct FinalTcParams {
  int32_t n_img, H, W;            // output extents; the input halo buffer is (H + 8) x (W + 8) x 32 per image
  int32_t strips, row_blocks, rb_rows, num_tasks;
  uint32_t idesc;
  const float* bias;              // [3] on the device
  float* out;
  const void* wpacked;
};

// Shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous (128 B);
// LBO = byte distance between the two core matrices of a K = 16 step, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_smem_desc_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;                                  // layout type 0 = no swizzle
}

__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

// Row g of a CTA's row stream uses ring entry g & 15 for the k-th time, k = g >> 4.  Three barriers per entry:
//   loaded[i]  TMA bytes of the row have landed                   (producer -> MMA)
//   summed[i]  the row's 18 MMAs have completed (tcgen05.commit)   (MMA -> producer: smem stage free;  MMA -> epilogue: T ready)
//   drained[i] the epilogue has read the T slot for the last time  (epilogue -> MMA: TMEM slot free)
// Measured while bringing this kernel up (tools/check_finalconv.py): the loop is bound by the single MMA-issuing thread --
// ~59 clocks per M128 x N32 x K16 MMA plus its own bookkeeping -- so that thread does nothing per MMA except add a
// compile-time constant to two descriptors, ring indices are powers of two, and one commit serves both consumers.
__global__ void __launch_bounds__(FC_THREADS, 1)
finalconv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ FinalTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem + FC_RING * FC_ROW_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + FC_W_BYTES);
  uint64_t* loaded = bars;                       // [RING]
  uint64_t* summed = bars + FC_RING;             // [RING]
  uint64_t* drained = bars + 2 * FC_RING;        // [RING]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * FC_RING);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_trigger();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    for (int s = 0; s < FC_RING; ++s) { mbar_init(&loaded[s], 1); mbar_init(&summed[s], 1); mbar_init(&drained[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  pdl_wait();
  {  // packed weights -> smem (generic-proxy writes, made visible to the MMA's async proxy below)
    const uint4* src = reinterpret_cast<const uint4*>(p.wpacked);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < FC_W_BYTES / 16; i += FC_THREADS) dst[i] = src[i];
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int Hq = p.H + 8;

  // task -> (image, strip, row block); every role walks the same (task, row) sequence with a running row counter g
  auto task_rows = [&](int task, int& n, int& x0, int& y0, int& rows_out) {
    const int rb = task % p.row_blocks; task /= p.row_blocks;
    const int sx = task % p.strips;
    n = task / p.strips;
    x0 = sx * FC_STRIP; y0 = rb * p.rb_rows;
    rows_out = min(p.rb_rows, p.H - y0);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int task = blockIdx.x; task < p.num_tasks; task += gridDim.x) {
        int n, x0, y0, rows_out;
        task_rows(task, n, x0, y0, rows_out);
        const int row0 = n * Hq + y0;
        for (int r = 0; r < rows_out + 8; ++r, ++g) {
          const uint32_t i = g & (FC_RING - 1);
          mbar_wait_spin(&summed[i], ((g >> 4) & 1) ^ 1);              // previous user of this stage has been consumed
          mbar_arrive_expect_tx(&loaded[i], FC_ROW_BYTES);
          // box {8 ch, 136 px, 4 chunks, 1 row}: lands as [chunk][px][8 ch]; pixels beyond the buffer are zero-filled
          tma_load_4d(smem + i * FC_ROW_BYTES, &map_a, &loaded[i], 0, x0, 0, row0 + r);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t a_lbo = FC_PW * 16, a_sbo = 128, b_lbo = 512, b_sbo = 128;
      const uint64_t db0 = umma_smem_desc_noswizzle(smem_u32(s_w), b_lbo, b_sbo);
      const uint64_t da00 = umma_smem_desc_noswizzle(smem_u32(smem), a_lbo, a_sbo);
      const uint32_t idesc = p.idesc;
      int total_rows = 0;
      for (int task = blockIdx.x; task < p.num_tasks; task += gridDim.x) {
        int n, x0, y0, rows_out;
        task_rows(task, n, x0, y0, rows_out);
        total_rows += rows_out + 8;
      }
      for (uint32_t g = 0; g < (uint32_t)total_rows; ++g) {
        const uint32_t i = g & (FC_RING - 1), ph = (g >> 4) & 1;
        mbar_wait_spin(&drained[i], ph ^ 1);
        mbar_wait_spin(&loaded[i], ph);
        tc_fence_after();
        const uint64_t da0 = da00 + (uint64_t)(i * (FC_ROW_BYTES >> 4));
        const uint32_t d_tmem = tmem_base + i * 32;
#pragma unroll
        for (int kw = 0; kw < 9; ++kw) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16(d_tmem, da0 + (uint64_t)((2 * ks * a_lbo + kw * 16) >> 4), db0 + (uint64_t)(((kw * 2 + ks) * 1024) >> 4), idesc,
                     (kw | ks) != 0 ? 1u : 0u);
        }
        umma_commit(&summed[i]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const float bias0 = p.bias[0], bias1 = p.bias[1], bias2 = p.bias[2];
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const size_t plane = (size_t)p.H * p.W;
    uint32_t g = 0;                              // row-stream index of the first input row of the current task
    for (int task = blockIdx.x; task < p.num_tasks; task += gridDim.x) {
      int n, x0, y0, rows_out;
      task_rows(task, n, x0, y0, rows_out);
      const int x = x0 + q * 32 + lane;
      float* orow = p.out + ((size_t)n * 3 * p.H + y0) * p.W + x;
      for (int j = 0; j < rows_out; ++j) {
        const uint32_t last = g + j + 8;         // newest input row this output row needs; MMAs complete in order
        mbar_wait_spin(&summed[last & (FC_RING - 1)], (last >> 4) & 1);
        tc_fence_after();
        uint32_t v[9][4];
#pragma unroll
        for (int kh = 0; kh < 9; ++kh) tmem_ld_x4(t_lane + (((g + j + kh) & (FC_RING - 1)) * 32 + kh * 3), v[kh]);
        tmem_ld_wait();
        // input row g + j is not needed by later output rows: hand its slot back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&drained[(g + j) & (FC_RING - 1)]);
        float acc0 = bias0, acc1 = bias1, acc2 = bias2;
#pragma unroll
        for (int kh = 0; kh < 9; ++kh) {
          acc0 += __uint_as_float(v[kh][0]); acc1 += __uint_as_float(v[kh][1]); acc2 += __uint_as_float(v[kh][2]);
        }
        if (x < p.W) { orow[0] = acc0; orow[plane] = acc1; orow[2 * plane] = acc2; }
        orow += p.W;
      }
      // the last 8 input rows of the task only held partial sums of output rows beyond this block: release their slots
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (int r = rows_out; r < rows_out + 8; ++r) mbar_arrive(&drained[(g + r) & (FC_RING - 1)]);
      g += rows_out + 8;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace fnst

using namespace fnst;

extern "C" int fnst_finalconv_tc(const void* act, const void* wpacked, const float* bias3, float* out, int n, int h, int w,
                                 int dtype, int device, void* stream) {
  FNST_CHECK_ARG(act && wpacked && bias3 && out && n > 0 && h > 0 && w > 0, "finalconv_tc: bad arguments");
  FNST_CHECK_ARG(dtype == FNST_F16 || dtype == FNST_BF16, "finalconv_tc: activations must be fp16 or bf16");
  FNST_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  FinalTcParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = n; p.H = h; p.W = w;
  p.strips = (w + FC_STRIP - 1) / FC_STRIP;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  // row blocks: 128 output rows per task (8 halo rows = 6 % extra) when that already yields two waves of tasks,
  // shorter blocks for small batches (a batch-1 256x256 image would otherwise be 4 tasks on 148 SMs)
  p.rb_rows = 128;
  while (p.rb_rows > 16 && (int64_t)n * p.strips * ((h + p.rb_rows - 1) / p.rb_rows) < 2 * sms) p.rb_rows /= 2;
  p.row_blocks = (h + p.rb_rows - 1) / p.rb_rows;
  p.num_tasks = n * p.strips * p.row_blocks;
  p.idesc = umma_idesc_f16(dtype == FNST_BF16 ? 1 : 0, 32, 0, 0);
  p.bias = bias3;
  p.out = out; p.wpacked = wpacked;
  const int hq = h + 8, wq = w + 8;
  CUtensorMap ma;
  {
    // activation halo buffer [n*hq][wq][32] as (8 ch | px | 4 chunks | row): box lands as [chunk][px][8 ch]
    PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
    FNST_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled unavailable");
    FNST_CHECK_ARG((reinterpret_cast<uintptr_t>(act) & 15) == 0, "finalconv_tc: activation pointer must be 16-byte aligned");
    cuuint64_t gdim[4] = {8, (cuuint64_t)wq, 4, (cuuint64_t)n * hq};
    cuuint64_t gstr[3] = {64, 16, (cuuint64_t)wq * 64};
    cuuint32_t bdim[4] = {8, (cuuint32_t)FC_PW, 4, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(act), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FNST_CHECK_ARG(r == CUDA_SUCCESS, "finalconv_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  }
  FNST_CUDA(cudaFuncSetAttribute(finalconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM));
  const int grid = p.num_tasks < sms ? p.num_tasks : sms;
  launch_pdl(finalconv_tc_kernel, dim3(grid), dim3(FC_THREADS), FC_SMEM, st, ma, p);
  return launch_status("finalconv_tc");
}
