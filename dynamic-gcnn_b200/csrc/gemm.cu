// Per-point GEMM = the 1x1 convolutions of the EdgeConv path (slim.conv2d kernel_size=1,
// /root/reference/dgcnn/ops.py:47-54,62-70) and their gradients.  fp32 SIMT, every output element is a
// sequential-in-k fmaf chain (per k-split), so results are run-to-run deterministic.
//   C[M,N] = op(A)[M,K] . op(B)[K,N]     transA: A stored [K,M]   transB: B stored [N,K]
// Weight-gradient shapes (M,N small, K = B*N points) are split over K across the grid and reduced in a
// fixed order by a second kernel.
#include "common.cuh"

namespace dgcnn {

constexpr int BM = 128, BN = 64, BK = 16, GEMM_THREADS = 256, PADA = 4, PADB = 4;

template <bool TA, bool TB>
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

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / GEMM_THREADS; ++i) {
      const int e = tid + i * GEMM_THREADS;
      int m, k;
      if (TA) {
        m = e & (BM - 1);
        k = e / BM;
      } else {
        k = e & (BK - 1);
        m = e / BK;
      }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.0f;
      if (gm < M && gk < kend) v = TA ? __ldg(A + (size_t)gk * M + gm) : __ldg(A + (size_t)gm * K + gk);
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / GEMM_THREADS; ++i) {
      const int e = tid + i * GEMM_THREADS;
      int n, k;
      if (TB) {
        k = e & (BK - 1);
        n = e / BK;
      } else {
        n = e & (BN - 1);
        k = e / BN;
      }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.0f;
      if (gn < N && gk < kend) v = TB ? __ldg(Bm + (size_t)gn * K + gk) : __ldg(Bm + (size_t)gk * N + gn);
      Bs[k][n] = v;
    }
    __syncthreads();
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
  (void)transA;
  (void)transB;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int s = pick_splits(M, N, K);
  return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

extern "C" int dgcnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, int transA, int transB,
                          void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DG_REQUIRE(A && B && C, DGCNN_ERR_INVALID, "gemm: null pointer");
  DG_REQUIRE(M > 0 && N > 0 && K > 0, DGCNN_ERR_INVALID, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  DG_REQUIRE(((uintptr_t)C & 15) == 0, DGCNN_ERR_INVALID, "gemm: C must be 16-byte aligned");
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
  if (transA && transB)
    sgemm_kernel<true, true><<<grid, GEMM_THREADS, 0, st>>>(A, B, out, M, N, K, kper);
  else if (transA)
    sgemm_kernel<true, false><<<grid, GEMM_THREADS, 0, st>>>(A, B, out, M, N, K, kper);
  else if (transB)
    sgemm_kernel<false, true><<<grid, GEMM_THREADS, 0, st>>>(A, B, out, M, N, K, kper);
  else
    sgemm_kernel<false, false><<<grid, GEMM_THREADS, 0, st>>>(A, B, out, M, N, K, kper);
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
