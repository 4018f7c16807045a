// Persistent 128x256-tile variant of the fp32-faithful ("bf16x3") tcgen05 GEMM of tc_gemm.cu, for the wide layers of
// the model head (/root/reference/dgcnn/model.py:65-72,88 + ops.py:151-160: MergedEdgeConv 1024, FC 512/256) and their
// gradients, i.e. every product whose N is a multiple of 256.
//
// Why a second kernel: with 128x128 tiles every k-slice moves (128+128) operand rows for 128x128 outputs, and at three
// MMAs per slice (hi.hi + hi.lo + lo.hi) both the shared-memory read port (8 KB per 64-cycle MMA = 128 B/clk) and the
// per-SM share of L2 bandwidth (64 KB per 768 MMA cycles = 83 B/clk against ~42 B/clk) sit at or past their limits.
// A 128x256 tile moves (128+256) rows for twice the outputs: 96 B/clk of shared memory and 62 -> 47 B/clk of L2.
// Structure (320 threads, one CTA per SM, CTAs loop over (m-tile, n-tile, k-split) work items, n fastest so that
// neighbouring CTAs share A rows in L2):
//   warp 0    TMA producer: 2-stage ring of {A hi, A lo (16 KB each), B hi, B lo (32 KB each)} = 96 KB per stage
//   warp 1    MMA issuer: tcgen05.mma M=128 N=256 K=16, accumulators double-buffered in TMEM (2 x 256 columns), so
//             the epilogue of one tile overlaps the main loop of the next
//   warps 2-9 epilogue: tcgen05.ld 32 lanes x 32 columns -> fp32 stores; optionally the per-column sum and sum of
//             squares of the tile's rows (train-mode BatchNorm statistics of slim.batch_norm, ops.py:53, taken from
//             the accumulator instead of a second pass over the output) -> colstats[m_tile][2][N]
#include <stdlib.h>

#include "tc_common.cuh"

namespace dgcnn {

constexpr int W_M = 128, W_N = 256, W_K = 64, W_THREADS = 320, W_STAGES = 2;
constexpr uint32_t W_ATILE = W_M * W_K * 2;   // 16 KB: one bf16 plane of the A tile
constexpr uint32_t W_BTILE = W_N * W_K * 2;   // 32 KB: one bf16 plane of the B tile
constexpr uint32_t W_STAGE = 2 * W_ATILE + 2 * W_BTILE;
constexpr size_t W_TR_OFF = (size_t)W_STAGES * W_STAGE;            // 8 warps x [32][17] floats (half-chunk transposes)
constexpr size_t W_CS_OFF = W_TR_OFF + 8 * 32 * 17 * 4;            // [4 sub-partitions][256 cols][2] floats
constexpr size_t W_BAR_OFF = W_CS_OFF + 4 * 256 * 2 * 4;
constexpr size_t W_SMEM = W_BAR_OFF + 128 + 1024 /*align*/;

template <bool A_K, bool B_K>
__global__ void __launch_bounds__(W_THREADS, 1)
    tc_gemm_wide_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                        int K, int kblocks_per_split, int splits, int planes, const __grid_constant__ WideOut out) {
  extern __shared__ unsigned char w_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)w_smem_raw + 1023) & ~(uintptr_t)1023);
  float* tr = reinterpret_cast<float*>(smem + W_TR_OFF);
  float* cs = reinterpret_cast<float*>(smem + W_CS_OFF);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + W_BAR_OFF);
  uint64_t* empty_bar = full_bar + W_STAGES;
  uint64_t* acc_full = empty_bar + W_STAGES;    // 2
  uint64_t* acc_empty = acc_full + 2;           // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt_n = (M + W_M - 1) / W_M, nt_n = (N + W_N - 1) / W_N;   // the last n-tile may be partial (N % 32 == 0)
  const int kb_total = (K + W_K - 1) / W_K;
  const int total = mt_n * nt_n * splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < W_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;   // stage uses so far
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int split = w / (mt_n * nt_n);
        const int rem = w - split * (mt_n * nt_n);
        const int m0 = (rem / nt_n) * W_M, n0 = (rem % nt_n) * W_N;
        const int kb_begin = split * kblocks_per_split;
        const int kb_end = min(kb_total, kb_begin + kblocks_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % W_STAGES;
          mbar_wait(&empty_bar[s], ((it / W_STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], planes == 2 ? W_STAGE : W_STAGE / 2);
          unsigned char* st = smem + (size_t)s * W_STAGE;
          const int k0 = kb * W_K;
#pragma unroll
          for (int plane = 0; plane < 2; ++plane) {
            if (plane >= planes) break;                     // single-plane (plain bf16) operands: hi only
            unsigned char* a_dst = st + plane * W_ATILE;
            unsigned char* b_dst = st + 2 * W_ATILE + plane * W_BTILE;
            if (A_K) {
              tma_load_3d(a_dst, &tmA, k0, m0, plane, &full_bar[s]);                  // box {64 k, 128 m}
            } else {
              tma_load_3d(a_dst, &tmA, m0, k0, plane, &full_bar[s]);                  // box {64 m, 64 k} x 2
              tma_load_3d(a_dst + W_ATILE / 2, &tmA, m0 + 64, k0, plane, &full_bar[s]);
            }
            if (B_K) {
              tma_load_3d(b_dst, &tmB, k0, n0, plane, &full_bar[s]);                  // box {64 k, 128 n} x 2
              tma_load_3d(b_dst + W_BTILE / 2, &tmB, k0, n0 + 128, plane, &full_bar[s]);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)                                              // box {64 n, 64 k} x 4
                tma_load_3d(b_dst + q * (W_BTILE / 4), &tmB, n0 + 64 * q, k0, plane, &full_bar[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_K ? 0u : 1u) << 15) | ((B_K ? 0u : 1u) << 16) |
                           ((uint32_t)(W_N >> 3) << 17) | ((uint32_t)(W_M >> 4) << 24);
    const uint32_t a_lbo = A_K ? 16u : 8192u, b_lbo = B_K ? 16u : 8192u;
    const uint32_t a_step = A_K ? 32u : 2048u, b_step = B_K ? 32u : 2048u;
    uint32_t it = 0, tile = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++tile) {
      const int split = w / (mt_n * nt_n);
      const int kb_begin = split * kblocks_per_split;
      const int kb_end = min(kb_total, kb_begin + kblocks_per_split);
      const uint32_t ab = tile & 1;
      mbar_wait(&acc_empty[ab], ((tile >> 1) & 1) ^ 1);
      const uint32_t acc = tmem_base + ab * W_N;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const int s = it % W_STAGES;
        mbar_wait(&full_bar[s], (it / W_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t st = smem_u32(smem + (size_t)s * W_STAGE);
          const uint32_t a_hi = st, a_lo = st + W_ATILE, b_hi = st + 2 * W_ATILE, b_lo = b_hi + W_BTILE;
#pragma unroll
          for (int ks = 0; ks < W_K / 16; ++ks) {
            const uint64_t dah = umma_desc(a_hi + ks * a_step, a_lbo, 1024);
            const uint64_t dal = umma_desc(a_lo + ks * a_step, a_lbo, 1024);
            const uint64_t dbh = umma_desc(b_hi + ks * b_step, b_lbo, 1024);
            const uint64_t dbl = umma_desc(b_lo + ks * b_step, b_lbo, 1024);
            if (planes == 2) {
              umma_bf16(acc, dal, dbh, idesc, (kb != kb_begin) || ks != 0);   // small terms first
              umma_bf16(acc, dah, dbl, idesc, 1);
              umma_bf16(acc, dah, dbh, idesc, 1);
            } else {
              umma_bf16(acc, dah, dbh, idesc, (kb != kb_begin) || ks != 0);   // plain bf16: one MMA per k-slice
            }
          }
          umma_commit(&empty_bar[s]);
          if (kb == kb_end - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
      }
      if (kb_end <= kb_begin && lane == 0) mbar_arrive(&acc_full[ab]);   // empty split: publish (garbage is never read)
    }
  } else {
    // epilogue warps 2..9 -> TMEM sub-partitions (warp % 4); the two warps of a sub-partition take four 32-column
    // chunks each (the epilogue is bound by latency and stores in flight, not by instruction issue)
    const int sub = warp & 3;
    const int eh = (warp - 2) >> 2;    // 0: warps 2..5, 1: warps 6..9
    const bool active = true;
    const int ch_begin = eh * 4, ch_end = eh * 4 + 4;
    const int et = threadIdx.x - 64;   // 0..255
    float* trw = tr + (warp - 2) * 32 * 17;
    uint32_t tile = 0;
    for (int w = blockIdx.x; active && w < total; w += gridDim.x, ++tile) {
      const int split = w / (mt_n * nt_n);
      const int rem = w - split * (mt_n * nt_n);
      const int mt = rem / nt_n;
      const int m0 = mt * W_M, n0 = (rem % nt_n) * W_N;
      const int kb_begin = split * kblocks_per_split;
      const bool empty = min(kb_total, kb_begin + kblocks_per_split) <= kb_begin;
      const uint32_t ab = tile & 1;
      float* Cout = out.C + (size_t)split * M * N;
      mbar_wait(&acc_full[ab], (tile >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int ch = ch_begin; ch < ch_end; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(sub * 32) << 16) + ab * W_N + (uint32_t)(ch * 32), v);
        if (empty) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        const int c0 = n0 + ch * 32;
        // where this 32-column chunk goes (chunks beyond N: the partial last tile; TMA zero-filled their operands)
        float* obase = Cout + c0;
        int opitch = N;
        if (out.n_groups > 0) {
          obase = nullptr;
          for (int g = 0; g < out.n_groups; ++g)
            if (c0 >= out.start[g] && c0 < out.start[g] + out.width[g]) {
              obase = out.ptr[g] + (c0 - out.start[g]);
              opitch = out.width[g];
            }
        }
        if (out.dbg_nostore || c0 >= N) obase = nullptr;
        // Transpose through shared memory in two 16-column halves: lane (c, rh) then owns column c of the rows
        // rh, rh+2, ...  Every store instruction writes two 64-byte row segments (whole 32-byte sectors; a lane
        // storing 16 bytes of its own row would fill half a sector per transaction), and the column statistics
        // fall out of the same loop.
        const int c = lane & 15, rh = lane >> 4;
        const int rbase = m0 + sub * 32;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int i = 0; i < 16; ++i) trw[lane * 17 + i] = __uint_as_float(v[h * 16 + i]);
          __syncwarp();
          float s1 = 0.0f, s2 = 0.0f;
          float* o = obase != nullptr ? obase + (size_t)(rbase + rh) * opitch + h * 16 + c : nullptr;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float z = trw[(2 * r + rh) * 17 + c];
            s1 += z;
            s2 = fmaf(z, z, s2);
            if (o != nullptr && rbase + 2 * r + rh < M) o[(size_t)(2 * r) * opitch] = z;
          }
          if (out.colstats != nullptr) {
            s1 += __shfl_xor_sync(FULL, s1, 16);
            s2 += __shfl_xor_sync(FULL, s2, 16);
            if (rh == 0) {
              cs[(sub * 256 + ch * 32 + h * 16 + c) * 2] = s1;
              cs[(sub * 256 + ch * 32 + h * 16 + c) * 2 + 1] = s2;
            }
          }
          __syncwarp();
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);   // accumulator drained: the MMA warp may start tile + 2
      if (out.colstats != nullptr) {
        asm volatile("bar.sync 2, 256;" ::: "memory");
        float* o = out.colstats + (size_t)mt * 2 * N + n0;
        {
          const int c = et;   // one column per epilogue thread
          float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
          for (int wv = 0; wv < 4; ++wv) {
            s1 += cs[(wv * 256 + c) * 2];
            s2 += cs[(wv * 256 + c) * 2 + 1];
          }
          o[c] = s1;
          o[N + c] = s2;
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// number of k-splits for the wide kernel: fill the machine when there are fewer tiles than SMs (weight gradients)
int tc_wide_splits(int M, int N, int K) {
  const int64_t tiles = (int64_t)cdiv(M, W_M) * cdiv(N, W_N);
  const int kb = cdiv(K, W_K);
  const int sms = num_sms();
  if (tiles >= sms || kb < 16) return 1;
  int s = (int)(sms / tiles);
  if (s > kb / 8) s = kb / 8;
  return s < 1 ? 1 : s;
}

// N need not be a multiple of the 256-column tile: the last tile's missing columns are zero-filled by TMA and never stored
bool tc_wide_ok(int M, int N, int K) { return N >= 256 && (N % 32) == 0 && M >= 128; }

// C (or groups / split partials) = op(A).op(B) with the wide kernel.  colstats may be null.
int tc_gemm_wide_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, bool a_k, bool b_k, int M, int N, int K,
                        int splits, int planes, const WideOut& out, cudaStream_t st) {
  const int kb = cdiv(K, W_K);
  const int kper = cdiv(kb, splits);
  const int total = cdiv(M, W_M) * cdiv(N, W_N) * splits;
  const int grid = total < num_sms() ? total : num_sms();
#define DG_WIDE(AK_, BK_)                                                                                     \
  do {                                                                                                         \
    static bool done_ = false;                                                                                 \
    if (!done_) {                                                                                              \
      cudaFuncSetAttribute(tc_gemm_wide_kernel<AK_, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                           (int)W_SMEM);                                                                       \
      done_ = true;                                                                                            \
    }                                                                                                          \
    tc_gemm_wide_kernel<AK_, BK_><<<grid, W_THREADS, W_SMEM, st>>>(tmA, tmB, M, N, K, kper, splits, planes, out); \
  } while (0)
  if (a_k && b_k) DG_WIDE(true, true);
  else if (a_k && !b_k) DG_WIDE(true, false);
  else if (!a_k && !b_k) DG_WIDE(false, false);
  else DG_WIDE(false, true);
#undef DG_WIDE
  count_launch();
  DG_CUDA_LAUNCH_CHECK("tc_gemm_wide_kernel");
  return DGCNN_OK;
}

}  // namespace dgcnn
