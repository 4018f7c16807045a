// Per-point GEMM = the 1x1 convolutions of the EdgeConv path (slim.conv2d kernel_size=1,
// /root/reference/dgcnn/ops.py:47-54,62-70) and their gradients.  fp32 SIMT, every output element is a
// sequential-in-k fmaf chain (per k-split), so results are run-to-run deterministic.
//   C[M,N] = op(A)[M,K] . op(B)[K,N]     transA: A stored [K,M]   transB: B stored [N,K]
// Weight-gradient shapes (M,N small, K = B*N points) are split over K across the grid and reduced in a
// fixed order by a second kernel.
#include "common.cuh"

namespace dgcnn {

constexpr int BM = 128, BN = 64, BK = 16, GEMM_THREADS = 256, PADA = 4, PADB = 4;

// Global -> register staging of one (BM x BK) A tile and one (BK x BN) B tile.  VEC: 16-byte loads along the
// contiguous dimension (needs that dimension and the base pointers 4-float aligned); otherwise scalar, fully guarded.
// Register prefetch: the next k-tile is fetched while the current one is multiplied (one __syncthreads pair per step).
template <bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
    sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ Cout, int M, int N,
                 int K, int kper) {
  __shared__ __align__(16) float As[BK][BM + PADA];
  __shared__ __align__(16) float Bs[BK][BN + PADB];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kper;
  const int kend = min(K, kbeg + kper);
  float* Cp = Cout + (size_t)blockIdx.z * M * N;

  constexpr int A_PER = (BM * BK) / GEMM_THREADS;   // 8 floats per thread
  constexpr int B_PER = (BN * BK) / GEMM_THREADS;   // 4 floats per thread
  float ra[A_PER], rb[B_PER];

  auto fetch = [&](int k0) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < A_PER / 4; ++i) {
        const int e = tid + i * GEMM_THREADS;            // float4 index
        int m, k;
        if (TA) { m = (e % (BM / 4)) * 4; k = e / (BM / 4); } else { k = (e % (BK / 4)) * 4; m = e / (BK / 4); }
        const int gm = m0 + m, gk = k0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (TA) { if (gk < kend && gm < M) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)gk * M + gm)); }
        else    { if (gm < M && gk < kend) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)gm * K + gk)); }
        ra[i * 4 + 0] = v.x; ra[i * 4 + 1] = v.y; ra[i * 4 + 2] = v.z; ra[i * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < B_PER / 4; ++i) {
        const int e = tid + i * GEMM_THREADS;
        int n, k;
        if (TB) { k = (e % (BK / 4)) * 4; n = e / (BK / 4); } else { n = (e % (BN / 4)) * 4; k = e / (BN / 4); }
        const int gn = n0 + n, gk = k0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (TB) { if (gn < N && gk < kend) v = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)gn * K + gk)); }
        else    { if (gk < kend && gn < N) v = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)gk * N + gn)); }
        rb[i * 4 + 0] = v.x; rb[i * 4 + 1] = v.y; rb[i * 4 + 2] = v.z; rb[i * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        const int e = tid + i * GEMM_THREADS;
        int m, k;
        if (TA) { m = e & (BM - 1); k = e / BM; } else { k = e & (BK - 1); m = e / BK; }
        const int gm = m0 + m, gk = k0 + k;
        ra[i] = (gm < M && gk < kend) ? (TA ? __ldg(A + (size_t)gk * M + gm) : __ldg(A + (size_t)gm * K + gk)) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < B_PER; ++i) {
        const int e = tid + i * GEMM_THREADS;
        int n, k;
        if (TB) { k = e & (BK - 1); n = e / BK; } else { n = e & (BN - 1); k = e / BN; }
        const int gn = n0 + n, gk = k0 + k;
        rb[i] = (gn < N && gk < kend) ? (TB ? __ldg(Bm + (size_t)gn * K + gk) : __ldg(Bm + (size_t)gk * N + gn)) : 0.f;
      }
    }
  };
  auto stash = [&]() {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < A_PER / 4; ++i) {
        const int e = tid + i * GEMM_THREADS;
        if (TA) {
          const int m = (e % (BM / 4)) * 4, k = e / (BM / 4);
          *reinterpret_cast<float4*>(&As[k][m]) = make_float4(ra[i * 4], ra[i * 4 + 1], ra[i * 4 + 2], ra[i * 4 + 3]);
        } else {
          const int k = (e % (BK / 4)) * 4, m = e / (BK / 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) As[k + j][m] = ra[i * 4 + j];
        }
      }
#pragma unroll
      for (int i = 0; i < B_PER / 4; ++i) {
        const int e = tid + i * GEMM_THREADS;
        if (TB) {
          const int k = (e % (BK / 4)) * 4, n = e / (BK / 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) Bs[k + j][n] = rb[i * 4 + j];
        } else {
          const int n = (e % (BN / 4)) * 4, k = e / (BN / 4);
          *reinterpret_cast<float4*>(&Bs[k][n]) = make_float4(rb[i * 4], rb[i * 4 + 1], rb[i * 4 + 2], rb[i * 4 + 3]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        const int e = tid + i * GEMM_THREADS;
        if (TA) As[e / BM][e & (BM - 1)] = ra[i]; else As[e & (BK - 1)][e / BK] = ra[i];
      }
#pragma unroll
      for (int i = 0; i < B_PER; ++i) {
        const int e = tid + i * GEMM_THREADS;
        if (TB) Bs[e & (BK - 1)][e / BK] = rb[i]; else Bs[e / BN][e & (BN - 1)] = rb[i];
      }
    }
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK);          // overlaps with the FMAs below
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
    const int gn = n0 + tx * 4;
    float* o = Cp + (size_t)gm * N + gn;
    if ((N & 3) == 0 && gn + 3 < N) {
      *reinterpret_cast<float4*>(o) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < N) o[j] = acc[i][j];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ C, int64_t MN, int splits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float s = 0.0f;
  for (int z = 0; z < splits; ++z) s += part[(size_t)z * MN + i];
  C[i] = s;
}

// ---- skinny shapes: the 2-class `Final` conv (model.py:94-101: [P,256] x [256,2]) and the weight gradients whose
// output is a few hundred numbers but whose contraction runs over all B*N points (Final, and EdgeConv0's 3-channel
// input).  The 128x64 tile kernel wastes >90 % of its lanes on them; these two are plain streaming kernels that keep
// the sequential-in-k fmaf order per output (per k-chunk for the gradient, chunks then summed in a fixed order).
// C[M,N] = A[M,K] . B[K,N], N <= 4: one warp per row, lanes own interleaved k (coalesced 16-byte loads), a fixed
// shuffle tree adds the 32 partial dot products.  Deterministic; NOT the sequential-in-k order of sgemm_kernel -- used
// only for the class-score layer, which no kNN graph depends on.  The lane's slice of B (K <= 512: at most 4 chunks of
// 4 k's x 4 columns) is kept in registers across rows.
__global__ void __launch_bounds__(256)
    sgemm_skinny_n_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, int M, int N,
                          int K) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  if (vec && K <= 512) {
    float bw[4][4][4];   // [chunk][k in chunk][n]
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int k = q * 128 + lane * 4 + i;
          bw[q][i][n] = (k < K && n < N) ? __ldg(Bm + (size_t)k * N + n) : 0.f;
        }
    for (int row = warp; row < M; row += nwarps) {
      const float* ar = A + (size_t)row * K;
      float a[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = q * 128 + lane * 4;
        if (k < K) *reinterpret_cast<float4*>(a[q]) = __ldg(reinterpret_cast<const float4*>(ar + k));
        else a[q][0] = a[q][1] = a[q][2] = a[q][3] = 0.f;
      }
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int n = 0; n < 4; ++n) acc[n] = __fmaf_rn(a[q][i], bw[q][i][n], acc[n]);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(FULL, acc[n], o);
      }
      if (lane == 0) {
        for (int n = 0; n < N; ++n) C[(size_t)row * N + n] = acc[n];
      }
    }
    return;
  }
  for (int row = warp; row < M; row += nwarps) {
    const float* ar = A + (size_t)row * K;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 4; k < K; k += 128) {
      float a[4];
      if (vec && k + 3 < K) {
        *reinterpret_cast<float4*>(a) = __ldg(reinterpret_cast<const float4*>(ar + k));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = k + i < K ? __ldg(ar + k + i) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (k + i < K) {
#pragma unroll
          for (int n = 0; n < 4; ++n)
            if (n < N) acc[n] = __fmaf_rn(a[i], __ldg(Bm + (size_t)(k + i) * N + n), acc[n]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(FULL, acc[n], o);
    }
    if (lane == 0) {
      for (int n = 0; n < N; ++n) C[(size_t)row * N + n] = acc[n];
    }
  }
}

// Weight gradient with one narrow operand (width <= 4) and one wide operand (width 64 / 128 / 256), contraction over
// K rows:  out[w][s] = sum_k Wd[k][w] * Sd[k][s].  A thread owns 4 consecutive wide columns (16-byte loads of the Wd
// rows, two rows in flight), the narrow row is a broadcast load; the block's row groups are summed in a fixed order and
// the block's partial is written in C's own [M][N] layout: part[block][m*N + n].
__global__ void __launch_bounds__(256)
    sgemm_skinny_dw_kernel(const float* __restrict__ Wd, const float* __restrict__ Sd, float* __restrict__ part, int wide,
                           int narrow, int K, int kper, int wide_is_a) {
  __shared__ float red[256][17];
  const int tpr = wide >> 2;                    // threads per row
  const int groups = 256 / tpr;                 // row groups per block
  const int wv = threadIdx.x % tpr, rg = threadIdx.x / tpr;
  const int kbeg = blockIdx.x * kper, kend = min(K, kbeg + kper);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  auto step = [&](int k) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(Wd + (size_t)k * wide + wv * 4));
    float sv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sv[j] = j < narrow ? __ldg(Sd + (size_t)k * narrow + j) : 0.f;
    const float xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(xx[i], sv[j], acc[i][j]);
  };
  int k = kbeg + rg;
  for (; k + groups < kend; k += 2 * groups) {
    step(k);
    step(k + groups);
  }
  if (k < kend) step(k);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.x][i * 4 + j] = acc[i][j];
  __syncthreads();
  // thread t < wide * narrow: wide column t / narrow, narrow column t % narrow
  for (int t = threadIdx.x; t < wide * narrow; t += 256) {
    const int w = t / narrow, j = t - w * narrow;
    float sum = 0.f;
    for (int g = 0; g < groups; ++g) sum += red[g * tpr + (w >> 2)][(w & 3) * 4 + j];
    const int m = wide_is_a ? w : j, n = wide_is_a ? j : w;
    const int N = wide_is_a ? narrow : wide;
    part[(size_t)blockIdx.x * wide * narrow + m * N + n] = sum;
  }
}

// C = sum over blocks of part[b][M*N], fixed order: 32 consecutive outputs x 8 partial lanes per block
__global__ void __launch_bounds__(256)
    skinny_dw_reduce_kernel(const float* __restrict__ part, float* __restrict__ C, int MN, int nb) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f;
  if (o < MN) {
    int b = w;
    for (; b + 8 < nb; b += 16) {
      s0 += part[(size_t)b * MN + o];
      s1 += part[(size_t)(b + 8) * MN + o];
    }
    if (b < nb) s0 += part[(size_t)b * MN + o];
  }
  red[w][lane] = s0 + s1;
  __syncthreads();
  if (w == 0 && o < MN) {
    float t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][lane];
    C[o] = t;
  }
}

static inline bool skinny_n(int M, int N, int K, int transA, int transB) {
  return !transA && !transB && N <= 4 && M >= 4096;
}
static inline bool skinny_wide_ok(int w) { return w == 64 || w == 128 || w == 256; }
static inline bool skinny_dw(int M, int N, int K, int transA, int transB) {
  return transA && !transB && K >= 8192 && ((M <= 4 && skinny_wide_ok(N)) || (N <= 4 && skinny_wide_ok(M)));
}
// (the wide operand must be 16-byte aligned for the float4 loads: checked at launch)
static inline int skinny_dw_blocks(int K) {
  int b = 2 * num_sms();
  const int maxb = cdiv(K, 64);
  return b < maxb ? b : maxb;
}

static int pick_splits(int M, int N, int K) {
  const int64_t tiles = (int64_t)cdiv(M, BM) * cdiv(N, BN);
  const int sms = num_sms();
  if (tiles >= sms || K < 512) return 1;
  int s = (int)((2 * (int64_t)sms + tiles - 1) / tiles);
  const int smax = K / 128;
  if (s > smax) s = smax;
  if (s > 1024) s = 1024;
  return s < 1 ? 1 : s;
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" size_t dgcnn_gemm_workspace_bytes(int M, int N, int K, int transA, int transB) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (skinny_n(M, N, K, transA, transB)) return 0;
  if (skinny_dw(M, N, K, transA, transB)) return (size_t)skinny_dw_blocks(K) * M * N * sizeof(float);
  const int s = pick_splits(M, N, K);
  return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

extern "C" int dgcnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, int transA, int transB,
                          void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DG_REQUIRE(A && B && C, DGCNN_ERR_INVALID, "gemm: null pointer");
  DG_REQUIRE(M > 0 && N > 0 && K > 0, DGCNN_ERR_INVALID, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  DG_REQUIRE(((uintptr_t)C & 15) == 0, DGCNN_ERR_INVALID, "gemm: C must be 16-byte aligned");
  if (skinny_n(M, N, K, transA, transB)) {
    const int blocks = cdiv(M, 8) < 3 * num_sms() ? cdiv(M, 8) : 3 * num_sms();   // >= 14 rows per warp: B stays in registers
    sgemm_skinny_n_kernel<<<blocks, 256, 0, st>>>(A, B, C, M, N, K);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("sgemm_skinny_n_kernel");
    return DGCNN_OK;
  }
  if (skinny_dw(M, N, K, transA, transB) && (((uintptr_t)A | (uintptr_t)B) & 15) == 0) {
    const int nb = skinny_dw_blocks(K);
    const size_t need = (size_t)nb * M * N * sizeof(float);
    DG_REQUIRE(ws && ws_bytes >= need, DGCNN_ERR_WORKSPACE, "gemm: workspace %zu < %zu bytes", ws_bytes, need);
    const int kper = cdiv(K, nb);
    const bool wide_is_a = N <= 4 && skinny_wide_ok(M);
    if (wide_is_a)
      sgemm_skinny_dw_kernel<<<nb, 256, 0, st>>>(A, B, reinterpret_cast<float*>(ws), M, N, K, kper, 1);
    else
      sgemm_skinny_dw_kernel<<<nb, 256, 0, st>>>(B, A, reinterpret_cast<float*>(ws), N, M, K, kper, 0);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("sgemm_skinny_dw_kernel");
    skinny_dw_reduce_kernel<<<cdiv((int64_t)M * N, 32), 256, 0, st>>>(reinterpret_cast<float*>(ws), C, M * N, nb);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("skinny_dw_reduce_kernel");
    return DGCNN_OK;
  }
  const int splits = pick_splits(M, N, K);
  float* out = C;
  if (splits > 1) {
    const size_t need = (size_t)splits * M * N * sizeof(float);
    DG_REQUIRE(ws && ws_bytes >= need, DGCNN_ERR_WORKSPACE, "gemm: workspace %zu < %zu bytes", ws_bytes, need);
    DG_REQUIRE(((uintptr_t)ws & 15) == 0, DGCNN_ERR_INVALID, "gemm: workspace must be 16-byte aligned");
    out = reinterpret_cast<float*>(ws);
  }
  int kper = cdiv(K, splits);
  kper = cdiv(kper, BK) * BK;
  dim3 grid(cdiv(N, BN), cdiv(M, BM), splits);
  DG_REQUIRE(grid.y <= 65535, DGCNN_ERR_UNSUPPORTED, "gemm: M=%d too large for grid.y", M);
  // 16-byte loads need the contiguous dimension of each operand (and kper for k-contiguous ones) 4-float aligned
  const bool a_ok = transA ? (M & 3) == 0 : ((K & 3) == 0 && (kper & 3) == 0);
  const bool b_ok = transB ? ((K & 3) == 0 && (kper & 3) == 0) : (N & 3) == 0;
  const bool vec = a_ok && b_ok && (((uintptr_t)A | (uintptr_t)B) & 15) == 0;
#define DG_LAUNCH_SGEMM(TA_, TB_)                                                                     \
  do {                                                                                                \
    if (vec) sgemm_kernel<TA_, TB_, true><<<grid, GEMM_THREADS, 0, st>>>(A, B, out, M, N, K, kper);   \
    else sgemm_kernel<TA_, TB_, false><<<grid, GEMM_THREADS, 0, st>>>(A, B, out, M, N, K, kper);      \
  } while (0)
  if (transA && transB) DG_LAUNCH_SGEMM(true, true);
  else if (transA) DG_LAUNCH_SGEMM(true, false);
  else if (transB) DG_LAUNCH_SGEMM(false, true);
  else DG_LAUNCH_SGEMM(false, false);
#undef DG_LAUNCH_SGEMM
  count_launch();
  DG_CUDA_LAUNCH_CHECK("sgemm_kernel");
  if (splits > 1) {
    const int64_t MN = (int64_t)M * N;
    splitk_reduce_kernel<<<cdiv(MN, 256), 256, 0, st>>>(out, C, MN, splits);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("splitk_reduce_kernel");
  }
  return DGCNN_OK;
}
