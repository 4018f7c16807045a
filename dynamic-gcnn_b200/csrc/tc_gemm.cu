// tcgen05 / TMEM / TMA GEMM with fp32-faithful arithmetic ("bf16x3"):
//   every fp32 operand x is pre-split into two bf16 planes  hi = bf16(x), lo = bf16(x - hi)   (dgcnn_split_bf16)
//   C = A.B is accumulated in fp32 TMEM as  hi_a.hi_b + hi_a.lo_b + lo_a.hi_b   (3 tcgen05.mma per k-slice);
//   the dropped lo.lo term and the lo rounding leave a relative error of ~2^-17 per product, i.e. fp32-class
//   results (the logits parity bound is 1e-3) at 3/2 of the tensor time of one TF32 pass and 2^-6 of its error.
// Used for the 1x1 convolutions (slim.conv2d kernel_size=1: /root/reference/dgcnn/ops.py:47-54,62-70,151-160,
// model.py:65-72) and their gradients.  One kernel covers the three operand layouts that occur, with no
// transposed copies:   forward  C = X.W      A K-major,  B MN-major (W is [K,N])
//                      dX       C = g.W^T    A K-major,  B K-major
//                      dW       C = X^T.g    A MN-major, B MN-major, split over the point dimension
// Structure (one 128x128 output tile per CTA, 192 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor into a 3-stage ring of 128B-swizzled tiles, mbarrier tx-count
//   warp 1   : MMA issuer    -- allocates 128 TMEM columns, one elected lane issues tcgen05.mma (M=128,N=128,K=16),
//                               tcgen05.commit releases ring slots / publishes the accumulator
//   warps 2-5: epilogue      -- tcgen05.ld 32 lanes x 32 columns, fp32 stores (each warp owns its TMEM sub-partition)
#include <stdlib.h>

#include "tc_common.cuh"

namespace dgcnn {

constexpr int GB_M = 128, GB_N = 128, GB_K = 64, G_THREADS = 192;
constexpr uint32_t TILE_BYTES = GB_M * GB_K * 2;          // 16 KB: one bf16 plane of one operand tile
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;          // A_hi, A_lo, B_hi, B_lo
constexpr size_t g_smem(int stages) { return (size_t)stages * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/; }

// A_K / B_K: operand is K-major (contraction index contiguous in global memory) or MN-major.
// Column groups of the output written to separate contiguous buffers (gradients of a multi-source concat operand:
// each source gets its own [M, width] tensor instead of a strided slice).  Widths are multiples of 32.
struct OutGroups {
  int n;
  int start[32];
  int width[32];
  float* ptr[32];
};

// G_STAGES = 3: deep TMA ring, one CTA per SM (long K).  G_STAGES = 1: 64 KB per CTA so that three CTAs share an SM and
// overlap each other's load / MMA / epilogue phases (short K, where the epilogue dominates).
template <bool A_K, bool B_K, int G_STAGES>
__global__ void __launch_bounds__(G_THREADS, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   float* __restrict__ C, int M, int N, int K, int kblocks_per_split, int planes,
                   const __grid_constant__ OutGroups og) {
  extern __shared__ unsigned char g_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)g_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)G_STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G_STAGES;
  uint64_t* accum_bar = empty_bar + G_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GB_M, n0 = blockIdx.y * GB_N;
  const int kb_total = (K + GB_K - 1) / GB_K;
  const int kb_begin = blockIdx.z * kblocks_per_split;
  const int kb_end = min(kb_total, kb_begin + kblocks_per_split);
  const int nkb = kb_end - kb_begin;
  float* Cout = C + (size_t)blockIdx.z * M * N;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 128 columns x 128 lanes of fp32 accumulator
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nkb > 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % G_STAGES;
        const uint32_t ph = (i / G_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], planes == 2 ? STAGE_BYTES : STAGE_BYTES / 2);
        unsigned char* st = smem + (size_t)s * STAGE_BYTES;
        const int k0 = (kb_begin + i) * GB_K;
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          if (plane >= planes) break;                       // single-plane (plain bf16) operands: hi only
          unsigned char* a_dst = st + plane * TILE_BYTES;
          unsigned char* b_dst = st + (2 + plane) * TILE_BYTES;
          if (A_K) {
            tma_load_3d(a_dst, &tmA, k0, m0, plane, &full_bar[s]);               // box {64 k, 128 m}
          } else {
            tma_load_3d(a_dst, &tmA, m0, k0, plane, &full_bar[s]);               // box {64 m, 64 k} x 2
            tma_load_3d(a_dst + TILE_BYTES / 2, &tmA, m0 + 64, k0, plane, &full_bar[s]);
          }
          if (B_K) {
            tma_load_3d(b_dst, &tmB, k0, n0, plane, &full_bar[s]);
          } else {
            tma_load_3d(b_dst, &tmB, n0, k0, plane, &full_bar[s]);
            tma_load_3d(b_dst + TILE_BYTES / 2, &tmB, n0 + 64, k0, plane, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_K ? 0u : 1u) << 15) | ((B_K ? 0u : 1u) << 16) |
                           ((uint32_t)(GB_N >> 3) << 17) | ((uint32_t)(GB_M >> 4) << 24);
    // K-major SW128: 8-row groups 1024 B apart, +32 B per 16-element k-slice
    // MN-major SW128: 64-wide MN atoms 8192 B apart (LBO), 8-k groups 1024 B apart (SBO), +2048 B per k-slice
    const uint32_t a_lbo = A_K ? 16u : 8192u, b_lbo = B_K ? 16u : 8192u;
    const uint32_t a_step = A_K ? 32u : 2048u, b_step = B_K ? 32u : 2048u;
    for (int i = 0; i < nkb; ++i) {
      const int s = i % G_STAGES;
      const uint32_t ph = (i / G_STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t a_hi = st, a_lo = st + TILE_BYTES, b_hi = st + 2 * TILE_BYTES, b_lo = st + 3 * TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < GB_K / 16; ++ks) {
          const uint64_t dah = umma_desc(a_hi + ks * a_step, a_lbo, 1024);
          const uint64_t dal = umma_desc(a_lo + ks * a_step, a_lbo, 1024);
          const uint64_t dbh = umma_desc(b_hi + ks * b_step, b_lbo, 1024);
          const uint64_t dbl = umma_desc(b_lo + ks * b_step, b_lbo, 1024);
          if (planes == 2) {
            umma_bf16(tmem_base, dal, dbh, idesc, (i | ks) != 0);   // small terms first
            umma_bf16(tmem_base, dah, dbl, idesc, 1);
            umma_bf16(tmem_base, dah, dbh, idesc, 1);
          } else {
            umma_bf16(tmem_base, dah, dbh, idesc, (i | ks) != 0);   // plain bf16: one MMA per k-slice
          }
        }
        umma_commit(&empty_bar[s]);                       // ring slot free once these MMAs retire
        if (i == nkb - 1) umma_commit(accum_bar);        // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // epilogue warps 2..5 -> TMEM sub-partitions (warp % 4)
    const int sub = warp & 3;
    const int row = m0 + sub * 32 + lane;
    if (nkb > 0) {
      mbar_wait(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
#pragma unroll 1
    for (int ch = 0; ch < GB_N / 32; ++ch) {
      uint32_t v[32];
      if (nkb > 0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(ch * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (row < M) {
        const int c0 = n0 + ch * 32;
        float* o = Cout + (size_t)row * N + c0;
        if (og.n > 0) {
          o = nullptr;
          for (int g = 0; g < og.n; ++g)
            if (c0 >= og.start[g] && c0 < og.start[g] + og.width[g])
              o = og.ptr[g] + (size_t)row * og.width[g] + (c0 - og.start[g]);
          if (o == nullptr) continue;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(o + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                            __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        } else if ((N & 3) == 0 && c0 + 31 < N) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(o + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                            __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < N) o[i] = __uint_as_float(v[i]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// fp32 [rows, cols] (row pitch ldx) -> bf16 planes hi / lo written at row pitch ldo (so that several sources can
// fill column slices of one [2][rows][ldo] operand: the reference's tf.concat, model.py:60-63,83-85, by construction)
__global__ void split_bf16_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ldx,
                                  __nv_bfloat16* __restrict__ planes, int64_t ldo, int64_t plane_elems) {
  const int cv = cols >> 2;  // float4 groups per row
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cv) return;
  const int64_t r = i / cv;
  const int c = (int)(i - r * cv) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
  const float f[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2bfloat16_rn(f[j]);
    l[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h[j]));
  }
  *reinterpret_cast<uint2*>(planes + r * ldo + c) = *reinterpret_cast<uint2*>(h);
  if (plane_elems) *reinterpret_cast<uint2*>(planes + plane_elems + r * ldo + c) = *reinterpret_cast<uint2*>(l);
}

EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;  // immutable after first successful lookup
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_plane_map(CUtensorMap* tm, const void* planes, int64_t rows, int64_t cols, uint32_t box_rows, int n_planes) {
  return make_plane_map_ld(tm, planes, rows, cols, cols, rows * cols, box_rows, n_planes);
}

// column slice of a wider operand: row pitch ld elements, second plane plane_elems elements after the first
int make_plane_map_ld(CUtensorMap* tm, const void* planes, int64_t rows, int64_t cols, int64_t ld, int64_t plane_elems,
                      uint32_t box_rows, int n_planes) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return set_err(DGCNN_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(n_planes == 1 ? 1 : 2)};
  if (plane_elems <= 0) plane_elems = rows * ld;      // unused stride of a single-plane operand (must still be valid)
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(planes), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(DGCNN_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DGCNN_OK;
}

// kmajor: box {64 cols(k), 128 rows};  mn-major: box {64 cols(mn), 64 rows(k)}
static int make_map(CUtensorMap* tm, const void* planes, int64_t rows, int64_t cols, bool kmajor, int n_planes) {
  return make_plane_map(tm, planes, rows, cols, kmajor ? 128u : 64u, n_planes);
}

static int tc_splits(int M, int N, int K) {
  const int64_t tiles = (int64_t)cdiv(M, GB_M) * cdiv(N, GB_N);
  const int kb = cdiv(K, GB_K);
  const int sms = num_sms();
  if (tiles >= sms || kb < 16) return 1;
  int s = (int)((sms + tiles - 1) / tiles);
  if (s > kb / 8) s = kb / 8;
  return s < 1 ? 1 : s;
}

// C = sum over splits of part[z], in a fixed order: block = 32 consecutive outputs x 8 split lanes (warp w adds
// splits w, w+8, ...: coalesced 128-byte reads), the 8 partial sums are then added in warp order; 8x the parallelism
// of one thread per output, which matters because the weight-gradient shapes have few outputs and ~100 splits
__global__ void __launch_bounds__(256)
    tc_splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ C, int64_t MN, int splits) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t o = (int64_t)blockIdx.x * 32 + lane;
  float s = 0.0f;
  if (o < MN)
    for (int z = w; z < splits; z += 8) s += part[(size_t)z * MN + o];
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && o < MN) {
    float t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][lane];
    C[o] = t;
  }
}

int launch_split_bf16(const float* x, int64_t rows, int cols, int64_t ldx, void* planes, int64_t ldo,
                      int64_t plane_elems, cudaStream_t st) {
  const int64_t blocks = (rows * (cols >> 2) + 255) / 256;
  DG_REQUIRE(blocks > 0 && blocks < (1ll << 31), DGCNN_ERR_UNSUPPORTED, "split_bf16: bad element count");
  split_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, rows, cols, ldx, (__nv_bfloat16*)planes, ldo, plane_elems);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("split_bf16_kernel");
  return DGCNN_OK;
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" int dgcnn_split_bf16(const float* x, int64_t rows, int cols, int64_t ldx, void* planes, int64_t ldo,
                                int64_t plane_elems, dgcnn_stream_t stream) {
  DG_REQUIRE(x && planes, DGCNN_ERR_INVALID, "split_bf16: null pointer");
  DG_REQUIRE(rows > 0 && cols > 0 && (cols & 3) == 0, DGCNN_ERR_INVALID,
             "split_bf16: rows=%lld cols=%d (cols must be a positive multiple of 4)", (long long)rows, cols);
  DG_REQUIRE(ldx >= cols && ldo >= cols && (ldx & 3) == 0 && (ldo & 3) == 0 && (plane_elems & 3) == 0 && plane_elems >= 0,
             DGCNN_ERR_INVALID, "split_bf16: pitches must be multiples of 4 and >= cols");
  DG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)planes & 7) == 0, DGCNN_ERR_INVALID, "split_bf16: alignment");
  return launch_split_bf16(x, rows, cols, ldx, planes, ldo, plane_elems, (cudaStream_t)stream);
}

extern "C" size_t dgcnn_tc_gemm_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int s = tc_wide_ok(M, N, K) ? tc_wide_splits(M, N, K) : tc_splits(M, N, K);
  return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

static int tc_gemm_impl(const void* a_planes, const void* b_planes, float* C, int M, int N, int K, int transA,
                        int transB, int planes, void* ws, size_t ws_bytes, const OutGroups& og, float* colstats,
                        dgcnn_stream_t stream, int64_t a_ld = 0, int64_t a_plane_elems = 0);

extern "C" int dgcnn_tc_gemm(const void* a_planes, const void* b_planes, float* C, int M, int N, int K, int transA,
                             int transB, int planes, void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  DG_REQUIRE(C, DGCNN_ERR_INVALID, "tc_gemm: null output");
  OutGroups og;
  og.n = 0;
  return tc_gemm_impl(a_planes, b_planes, C, M, N, K, transA, transB, planes, ws, ws_bytes, og, nullptr, stream);
}

extern "C" int dgcnn_tc_gemm_a_slice(const void* a_planes, int64_t a_ld, int64_t a_plane_elems, const void* b_planes,
                                     float* C, int M, int N, int K, int transA, int transB, int planes, void* ws,
                                     size_t ws_bytes, dgcnn_stream_t stream) {
  DG_REQUIRE(C, DGCNN_ERR_INVALID, "tc_gemm_a_slice: null output");
  const int64_t a_cols = transA ? M : K;
  DG_REQUIRE(a_ld >= a_cols && (a_ld & 7) == 0 && a_plane_elems >= 0 && (a_plane_elems & 7) == 0 &&
                 (a_plane_elems > 0 || planes == 1), DGCNN_ERR_INVALID,
             "tc_gemm_a_slice: bad pitch %lld / plane distance %lld", (long long)a_ld, (long long)a_plane_elems);
  OutGroups og;
  og.n = 0;
  return tc_gemm_impl(a_planes, b_planes, C, M, N, K, transA, transB, planes, ws, ws_bytes, og, nullptr, stream, a_ld,
                      a_plane_elems);
}

extern "C" int dgcnn_tc_gemm_stats_supported(int M, int N, int K) {
  return (tc_wide_ok(M, N, K) && (N % 256) == 0 && tc_wide_splits(M, N, K) == 1) ? 1 : 0;
}

extern "C" int dgcnn_tc_gemm_stats(const void* a_planes, const void* b_planes, float* C, int M, int N, int K, int transA,
                                   int transB, int planes, float* colstats, dgcnn_stream_t stream) {
  DG_REQUIRE(C && colstats, DGCNN_ERR_INVALID, "tc_gemm_stats: null output");
  DG_REQUIRE(dgcnn_tc_gemm_stats_supported(M, N, K), DGCNN_ERR_UNSUPPORTED,
             "tc_gemm_stats: needs N %% 256 == 0, M >= 128 and no k-split (M=%d N=%d K=%d)", M, N, K);
  OutGroups og;
  og.n = 0;
  return tc_gemm_impl(a_planes, b_planes, C, M, N, K, transA, transB, planes, nullptr, 0, og, colstats, stream);
}

extern "C" int dgcnn_tc_gemm_grouped(const void* a_planes, const void* b_planes, int M, int N, int K, int transA,
                                     int transB, int planes, int n_groups, const int* starts, const int* widths,
                                     float* const* outs, dgcnn_stream_t stream) {
  DG_REQUIRE(n_groups >= 1 && n_groups <= 32 && starts && widths && outs, DGCNN_ERR_INVALID,
             "tc_gemm_grouped: need 1..32 output groups");
  DG_REQUIRE((tc_wide_ok(M, N, K) ? tc_wide_splits(M, N, K) : tc_splits(M, N, K)) == 1, DGCNN_ERR_UNSUPPORTED,
             "tc_gemm_grouped: shape would need split-K");
  OutGroups og;
  og.n = n_groups;
  int covered = 0;
  for (int g = 0; g < n_groups; ++g) {
    DG_REQUIRE(outs[g] && widths[g] > 0 && (widths[g] & 31) == 0 && (starts[g] & 31) == 0 && starts[g] >= 0 &&
                   starts[g] + widths[g] <= N && ((uintptr_t)outs[g] & 15) == 0,
               DGCNN_ERR_INVALID, "tc_gemm_grouped: group %d must be 32-column aligned inside [0,N) and 16-byte aligned", g);
    og.start[g] = starts[g];
    og.width[g] = widths[g];
    og.ptr[g] = outs[g];
    covered += widths[g];
  }
  DG_REQUIRE(covered <= N, DGCNN_ERR_INVALID, "tc_gemm_grouped: groups overlap");
  return tc_gemm_impl(a_planes, b_planes, outs[0], M, N, K, transA, transB, planes, nullptr, 0, og, nullptr, stream);
}

static int tc_gemm_impl(const void* a_planes, const void* b_planes, float* C, int M, int N, int K, int transA,
                        int transB, int planes, void* ws, size_t ws_bytes, const OutGroups& og, float* colstats,
                        dgcnn_stream_t stream, int64_t a_ld, int64_t a_plane_elems) {
  cudaStream_t st = (cudaStream_t)stream;
  DG_REQUIRE(planes == 1 || planes == 2, DGCNN_ERR_INVALID, "tc_gemm: planes must be 1 (bf16) or 2 (bf16 hi/lo), got %d", planes);
  DG_REQUIRE(a_planes && b_planes && C, DGCNN_ERR_INVALID, "tc_gemm: null pointer");
  DG_REQUIRE(M > 0 && N > 0 && K > 0, DGCNN_ERR_INVALID, "tc_gemm: bad shape M=%d N=%d K=%d", M, N, K);
  DG_REQUIRE((M & 7) == 0 && (N & 7) == 0 && (K & 7) == 0, DGCNN_ERR_UNSUPPORTED,
             "tc_gemm: M, N, K must be multiples of 8 (TMA row pitch), got %d %d %d", M, N, K);
  DG_REQUIRE(((uintptr_t)a_planes & 15) == 0 && ((uintptr_t)b_planes & 15) == 0 && ((uintptr_t)C & 15) == 0,
             DGCNN_ERR_INVALID, "tc_gemm: buffers must be 16-byte aligned");
  // op(A) is [M,K]: stored [M,K] (K-major) or, transA, [K,M] (MN-major).  op(B) is [K,N]: stored [K,N] (MN-major) or,
  // transB, [N,K] (K-major).
  const bool a_k = !transA, b_k = transB != 0;
  CUtensorMap tmA, tmB;
  int rc = a_ld > 0 ? make_plane_map_ld(&tmA, a_planes, a_k ? M : K, a_k ? K : M, a_ld, a_plane_elems, a_k ? 128u : 64u, planes)
                    : make_map(&tmA, a_planes, a_k ? M : K, a_k ? K : M, a_k, planes);
  if (rc) return rc;
  rc = make_map(&tmB, b_planes, b_k ? N : K, b_k ? K : N, b_k, planes);
  if (rc) return rc;
  if (tc_wide_ok(M, N, K)) {
    // 128x256 persistent kernel (tc_gemm_wide.cu)
    const int wsplits = tc_wide_splits(M, N, K);
    WideOut wo;
    wo.C = C;
    wo.colstats = colstats;
    wo.n_groups = og.n;
    wo.dbg_nostore = 0;
    for (int g = 0; g < og.n; ++g) {
      wo.start[g] = og.start[g];
      wo.width[g] = og.width[g];
      wo.ptr[g] = og.ptr[g];
    }
    if (wsplits > 1) {
      const size_t need = (size_t)wsplits * M * N * sizeof(float);
      DG_REQUIRE(ws && ws_bytes >= need, DGCNN_ERR_WORKSPACE, "tc_gemm: workspace %zu < %zu bytes", ws_bytes, need);
      wo.C = reinterpret_cast<float*>(ws);
    }
    rc = tc_gemm_wide_launch(tmA, tmB, a_k, b_k, M, N, K, wsplits, planes, wo, st);
    if (rc) return rc;
    if (wsplits > 1) {
      const int64_t MN = (int64_t)M * N;
      tc_splitk_reduce_kernel<<<cdiv(MN, 32), 256, 0, st>>>(wo.C, C, MN, wsplits);
      count_launch();
      DG_CUDA_LAUNCH_CHECK("tc_splitk_reduce_kernel");
    }
    return DGCNN_OK;
  }
  DG_REQUIRE(colstats == nullptr, DGCNN_ERR_UNSUPPORTED, "tc_gemm: column statistics need the wide kernel");
  const int splits = tc_splits(M, N, K);
  float* out = C;
  if (splits > 1) {
    const size_t need = (size_t)splits * M * N * sizeof(float);
    DG_REQUIRE(ws && ws_bytes >= need, DGCNN_ERR_WORKSPACE, "tc_gemm: workspace %zu < %zu bytes", ws_bytes, need);
    out = reinterpret_cast<float*>(ws);
  }
  const int kb = cdiv(K, GB_K);
  const int kper = cdiv(kb, splits);
  dim3 grid(cdiv(M, GB_M), cdiv(N, GB_N), splits);
  // deep ring when the k-loop is long, or when the grid leaves SMs idle anyway (split-K weight gradients: one CTA
  // per SM at most, so the single-stage variant's 3-CTAs-per-SM overlap cannot happen and every k-block would
  // expose a full TMA round trip)
  const int64_t ctas = (int64_t)grid.x * grid.y * grid.z;
  const int stages = (kper >= 64 || (ctas <= num_sms() && kper >= 3)) ? 3 : 1;
#define DG_TC_LAUNCH(AK_, BK_, ST_)                                                                          \
  do {                                                                                                       \
    static bool done_ = false;                                                                               \
    if (!done_) {                                                                                            \
      cudaFuncSetAttribute(tc_gemm_kernel<AK_, BK_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                           (int)g_smem(ST_));                                                                \
      done_ = true;                                                                                          \
    }                                                                                                        \
    tc_gemm_kernel<AK_, BK_, ST_><<<grid, G_THREADS, g_smem(ST_), st>>>(tmA, tmB, out, M, N, K, kper, planes, og); \
  } while (0)
#define DG_TC_LAUNCH_ST(AK_, BK_)          \
  do {                                     \
    if (stages == 3) DG_TC_LAUNCH(AK_, BK_, 3); \
    else DG_TC_LAUNCH(AK_, BK_, 1);        \
  } while (0)
  if (a_k && b_k) DG_TC_LAUNCH_ST(true, true);
  else if (a_k && !b_k) DG_TC_LAUNCH_ST(true, false);
  else if (!a_k && !b_k) DG_TC_LAUNCH_ST(false, false);
  else DG_TC_LAUNCH_ST(false, true);
#undef DG_TC_LAUNCH_ST
#undef DG_TC_LAUNCH
  count_launch();
  DG_CUDA_LAUNCH_CHECK("tc_gemm_kernel");
  if (splits > 1) {
    const int64_t MN = (int64_t)M * N;
    tc_splitk_reduce_kernel<<<cdiv(MN, 32), 256, 0, st>>>(out, C, MN, splits);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("tc_splitk_reduce_kernel");
  }
  return DGCNN_OK;
}
