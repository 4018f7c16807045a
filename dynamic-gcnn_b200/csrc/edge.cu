// EdgeConv gather path.
//   edge_feature(+bwd): /root/reference/dgcnn/ops.py:21-40 materialised (API parity for edges()).
//   edgeconv_*        : ops.py:45-57 fused -- gather -> conv0 -> BN(train) -> ReLU -> max_k / mean_k with
//                       z_ij = u_i + v_{idx(i,j)} (uv = x.[Wa-Wb | Wb], formed by dgcnn_gemm), so neither the
//                       [B,N,k,2C] edge tensor nor the [B,N,k,F] activation ever exists in HBM.
// Work split of the gather passes: one warp per point, lanes over channels (lane, lane+32 of a 64-channel
// chunk; chunk = blockIdx.y), neighbours unrolled for memory-level parallelism.  v rows are 256 B and the
// whole uv table (25 MB at B=24,N=2048,F=64) is L2-resident, so these passes run at L2 gather rate.
#include <cuda_bf16.h>

#include "common.cuh"

namespace dgcnn {

constexpr int EC_THREADS = 256;          // 8 warps
constexpr int EC_WARPS = EC_THREADS / 32;
constexpr int STAT_BLOCKS_PER_SM = 8;

// ------------------------------------------------------------------------------------------ edges()
__global__ void edge_feature_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                    float* __restrict__ out, int N, int C, int k, int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int C2 = 2 * C;
  const int ch = (int)(e % C2);
  const int64_t edge = e / C2;      // (b*N + i)*k + j
  const int64_t pt = edge / k;      // b*N + i
  const int64_t b = pt / N;
  const float* xi = x + pt * C;
  if (ch < C) {
    out[e] = xi[ch];                                          // ops.py:35-37 central
  } else {
    const int64_t nb = b * N + idx[edge];                     // ops.py:30-34 flat gather
    out[e] = __fsub_rn(x[nb * C + (ch - C)], xi[ch - C]);     // ops.py:39 neighbours - central
  }
}

__global__ void edge_feature_bwd_kernel(const float* __restrict__ g, const int32_t* __restrict__ idx,
                                        float* __restrict__ gx, int N, int C, int k, int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over P*k*C
  if (e >= total) return;
  const int ch = (int)(e % C);
  const int64_t edge = e / C;
  const int64_t pt = edge / k;
  const int64_t b = pt / N;
  const float gc = g[edge * 2 * C + ch];
  const float gn = g[edge * 2 * C + C + ch];
  atomicAdd(gx + pt * C + ch, gc - gn);
  atomicAdd(gx + (b * N + idx[edge]) * C + ch, gn);
}

// ------------------------------------------------------------------------------- fused EdgeConv passes
struct EcArgs {
  const float* uv;      // [P, 2F]   u | v
  const int32_t* idx;   // [P, k]    neighbour index inside the cloud
  int P, N, F, k;
};

// Load this point's k neighbour rows (global point index) into lanes; broadcast later by shuffle.
__device__ __forceinline__ void load_nbrs(const EcArgs& a, int p, int lane, int (&rows)[2]) {
  const int base = (p / a.N) * a.N;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int j = s * 32 + lane;
    rows[s] = (j < a.k) ? base + a.idx[(int64_t)p * a.k + j] : 0;
  }
}
// offset (in floats) of neighbour j's v row inside uv; j is warp-uniform
__device__ __forceinline__ int64_t nbr_off(const EcArgs& a, const int (&rows)[2], int j) {
  const int r = (j < 32) ? __shfl_sync(FULL, rows[0], j) : __shfl_sync(FULL, rows[1], j - 32);
  return (int64_t)r * (2 * a.F) + a.F;
}

// A lane owns the channel PAIR (f0, f0+1), f0 = 64*chunk + 2*lane: one 8-byte load per gathered row (and one 8-byte
// vector atomic per scattered row) instead of two 4-byte ones.  `pair` = both channels exist and F is even (8-byte
// alignment of every row start); otherwise element-wise with guards.
struct EcLane {
  int f0, f1;
  bool ok0, ok1, pair;
  __device__ __forceinline__ EcLane(int F, int chunk, int lane) {
    f0 = chunk * 64 + 2 * lane;
    f1 = f0 + 1;
    ok0 = f0 < F;
    ok1 = f1 < F;
    pair = ok1 && ((F & 1) == 0);
  }
  __device__ __forceinline__ void ld(const float* __restrict__ base, float& x0, float& x1) const {
    if (pair) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(base + f0));
      x0 = t.x;
      x1 = t.y;
    } else {
      x0 = ok0 ? __ldg(base + f0) : 0.f;
      x1 = ok1 ? __ldg(base + f1) : 0.f;
    }
  }
  __device__ __forceinline__ void st(float* __restrict__ base, float x0, float x1) const {
    if (pair) {
      *reinterpret_cast<float2*>(base + f0) = make_float2(x0, x1);
    } else {
      if (ok0) base[f0] = x0;
      if (ok1) base[f1] = x1;
    }
  }
  __device__ __forceinline__ void red(float* __restrict__ base, float x0, float x1) const {
    if (pair) {
      atomicAdd(reinterpret_cast<float2*>(base + f0), make_float2(x0, x1));
    } else {
      if (ok0) atomicAdd(base + f0, x0);
      if (ok1) atomicAdd(base + f1, x1);
    }
  }
};

// pass 1 forward: zmax, tie count, per-block partial sum / sum of squares
__global__ void __launch_bounds__(EC_THREADS)
    ec_fwd_stats_kernel(EcArgs a, float* __restrict__ zmax, float* __restrict__ cnt, double* __restrict__ acc) {
  __shared__ float red[2][EC_WARPS][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const EcLane L(a.F, blockIdx.y, lane);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (int p = blockIdx.x * EC_WARPS + warp; p < a.P; p += gridDim.x * EC_WARPS) {
    int rows[2];
    load_nbrs(a, p, lane, rows);
    float u0, u1;
    L.ld(a.uv + (int64_t)p * 2 * a.F, u0, u1);
    float m0 = -INFINITY, m1 = -INFINITY, c0 = 0.f, c1 = 0.f;
#pragma unroll 4
    for (int j = 0; j < a.k; ++j) {
      float v0, v1;
      L.ld(a.uv + nbr_off(a, rows, j), v0, v1);
      const float z0 = u0 + v0;
      const float z1 = u1 + v1;
      s0 += z0; q0 = fmaf(z0, z0, q0);
      s1 += z1; q1 = fmaf(z1, z1, q1);
      if (z0 > m0) { m0 = z0; c0 = 1.f; } else if (z0 == m0) c0 += 1.f;
      if (z1 > m1) { m1 = z1; c1 = 1.f; } else if (z1 == m1) c1 += 1.f;
    }
    L.st(zmax + (int64_t)p * a.F, m0, m1);
    L.st(cnt + (int64_t)p * a.F, c0, c1);
  }
  red[0][warp][2 * lane] = s0; red[0][warp][2 * lane + 1] = s1;
  red[1][warp][2 * lane] = q0; red[1][warp][2 * lane + 1] = q1;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < EC_WARPS; ++w) t += red[which][w][c];
    const int f = blockIdx.y * 64 + c;
    if (f < a.F) atomicAdd(&acc[which * a.F + f], (double)t);
  }
}

// Per-channel totals are accumulated by the statistics kernels with fp64 atomics into acc[2][C] (sum, sum of
// squares / cross term); fp64 makes the summation order irrelevant at fp32 resolution.  These kernels turn the
// totals into what the apply passes need.
__global__ void finalize_stats_kernel(const double* __restrict__ acc, int C, double count, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc[c] / count;
  double var = acc[C + c] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void finalize_sums_kernel(const double* __restrict__ acc, int C, float* __restrict__ s1,
                                     float* __restrict__ s2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  s1[c] = (float)acc[c];
  s2[c] = (float)acc[C + c];
}

// pass 2 forward: out_max, out_mean
__global__ void __launch_bounds__(EC_THREADS)
    ec_fwd_apply_kernel(EcArgs a, const float* __restrict__ zmax, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ beta, float* __restrict__ omax,
                        float* __restrict__ omean, int opitch, __nv_bfloat16* __restrict__ sink, int sink_ld,
                        size_t sink_plane) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const EcLane L(a.F, blockIdx.y, lane);
  const int f0 = L.f0, f1 = L.f1;
  const bool ok0 = L.ok0, ok1 = L.ok1;
  const float mu0 = ok0 ? mean[f0] : 0.f, mu1 = ok1 ? mean[f1] : 0.f;
  const float r0 = ok0 ? rstd[f0] : 0.f, r1 = ok1 ? rstd[f1] : 0.f;
  const float b0 = ok0 ? beta[f0] : 0.f, b1 = ok1 ? beta[f1] : 0.f;
  const float invk = 1.0f / (float)a.k;
  for (int p = blockIdx.x * EC_WARPS + warp; p < a.P; p += gridDim.x * EC_WARPS) {
    int rows[2];
    load_nbrs(a, p, lane, rows);
    float u0, u1;
    L.ld(a.uv + (int64_t)p * 2 * a.F, u0, u1);
    float y0 = 0.f, y1 = 0.f;
#pragma unroll 4
    for (int j = 0; j < a.k; ++j) {
      float v0, v1;
      L.ld(a.uv + nbr_off(a, rows, j), v0, v1);
      const float z0 = u0 + v0;
      const float z1 = u1 + v1;
      y0 += fmaxf(fmaf(z0 - mu0, r0, b0), 0.f);
      y1 += fmaxf(fmaf(z1 - mu1, r1, b1), 0.f);
    }
    const int64_t o = (int64_t)p * a.F, oo = (int64_t)p * opitch;
    // BN(+)ReLU are monotone, so the max commutes with them
    float zm0, zm1;
    L.ld(zmax + o, zm0, zm1);
    const float vx0 = fmaxf(fmaf(zm0 - mu0, r0, b0), 0.f), vx1 = fmaxf(fmaf(zm1 - mu1, r1, b1), 0.f);
    const float vm0 = y0 * invk, vm1 = y1 * invk;
    L.st(omax + oo, vx0, vx1);
    L.st(omean + oo, vm0, vm1);
    // optional plane sink: the same values as bf16 hi / lo planes at columns [0,F) (max) and [F,2F) (mean) of a
    // tensor-core operand (the consumer's concat operand), so that no separate split pass reads them again
    if (sink != nullptr && L.pair) {
      auto put = [&](int col, float x0, float x1) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
        const float2 hf = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
        const size_t e = (size_t)p * sink_ld + col;
        *reinterpret_cast<__nv_bfloat162*>(sink + e) = h;
        *reinterpret_cast<__nv_bfloat162*>(sink + sink_plane + e) = l;
      };
      put(f0, vx0, vx1);
      put(a.F + f0, vm0, vm1);
    }
  }
}

// backward pass 1: s1 = sum g_pre, s2 = sum g_pre * zhat (per-block partials)
// g_y_ij  = g_mean_i/k + [z_ij == zmax_i] g_max_i / cnt_i      (tf reduce_mean / reduce_max grads; ties share)
// g_pre_ij = g_y_ij * [pre_ij > 0]                             (ReluGrad)
template <bool APPLY>
__global__ void __launch_bounds__(EC_THREADS)
    ec_bwd_kernel(EcArgs a, const float* __restrict__ zmax, const float* __restrict__ cnt,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ beta,
                  const float* __restrict__ gmax, const float* __restrict__ gmean, const float* __restrict__ gboth,
                  const float* __restrict__ s1, const float* __restrict__ s2, double* __restrict__ acc,
                  float* __restrict__ guv) {
  __shared__ float red[2][EC_WARPS][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const EcLane L(a.F, blockIdx.y, lane);
  const int f0 = L.f0, f1 = L.f1;
  const bool ok0 = L.ok0, ok1 = L.ok1;
  const float mu0 = ok0 ? mean[f0] : 0.f, mu1 = ok1 ? mean[f1] : 0.f;
  const float r0 = ok0 ? rstd[f0] : 0.f, r1 = ok1 ? rstd[f1] : 0.f;
  const float b0 = ok0 ? beta[f0] : 0.f, b1 = ok1 ? beta[f1] : 0.f;
  const float invk = 1.0f / (float)a.k;
  const float invE = 1.0f / ((float)a.P * (float)a.k);
  float m10 = 0.f, m11 = 0.f, m20 = 0.f, m21 = 0.f;
  if (APPLY) {
    m10 = ok0 ? s1[f0] * invE : 0.f; m11 = ok1 ? s1[f1] * invE : 0.f;
    m20 = ok0 ? s2[f0] * invE : 0.f; m21 = ok1 ? s2[f1] * invE : 0.f;
  }
  float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (int p = blockIdx.x * EC_WARPS + warp; p < a.P; p += gridDim.x * EC_WARPS) {
    int rows[2];
    load_nbrs(a, p, lane, rows);
    const int64_t o = (int64_t)p * a.F;
    float u0, u1, zm0, zm1, cn0, cn1;
    L.ld(a.uv + (int64_t)p * 2 * a.F, u0, u1);
    L.ld(zmax + o, zm0, zm1);
    L.ld(cnt + o, cn0, cn1);
    // the gradients of max / mean arrive from up to two consumers: separate [P,F] tensors and/or one packed
    // [P,2F] = (max | mean) tensor (the conv1 operand); summed here instead of in a separate pass
    float gM0 = 0.f, gM1 = 0.f, gA0 = 0.f, gA1 = 0.f;
    if (gmax) L.ld(gmax + o, gM0, gM1);
    if (gmean) L.ld(gmean + o, gA0, gA1);
    if (gboth) {
      const float* gb = gboth + (int64_t)p * 2 * a.F;
      float t0, t1;
      L.ld(gb, t0, t1);
      gM0 += t0; gM1 += t1;
      L.ld(gb + a.F, t0, t1);
      gA0 += t0; gA1 += t1;
    }
    const float gm0 = gA0 * invk, gm1 = gA1 * invk;
    const float gx0 = ok0 ? gM0 / cn0 : 0.f, gx1 = ok1 ? gM1 / cn1 : 0.f;
    float gu0 = 0.f, gu1 = 0.f;
    if (!APPLY && guv != nullptr)     // statistics pass: clear this point's v half for the scatter-add of the apply pass
      L.st(guv + (int64_t)p * 2 * a.F + a.F, 0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < a.k; ++j) {
      const int64_t off = nbr_off(a, rows, j);
      float v0, v1;
      L.ld(a.uv + off, v0, v1);
      const float z0 = u0 + v0;
      const float z1 = u1 + v1;
      const float zh0 = (z0 - mu0) * r0, zh1 = (z1 - mu1) * r1;
      const bool act0 = fmaf(z0 - mu0, r0, b0) > 0.f, act1 = fmaf(z1 - mu1, r1, b1) > 0.f;
      const float gp0 = act0 ? gm0 + (z0 == zm0 ? gx0 : 0.f) : 0.f;
      const float gp1 = act1 ? gm1 + (z1 == zm1 ? gx1 : 0.f) : 0.f;
      if (!APPLY) {
        a0 += gp0; q0 = fmaf(gp0, zh0, q0);
        a1 += gp1; q1 = fmaf(gp1, zh1, q1);
      } else {
        const float gz0 = r0 * (gp0 - m10 - zh0 * m20);
        const float gz1 = r1 * (gp1 - m11 - zh1 * m21);
        gu0 += gz0; gu1 += gz1;
        L.red(guv + off, gz0, gz1);                // scatter-add into the v half (tf.gather grad)
      }
    }
    if (APPLY) L.st(guv + (int64_t)p * 2 * a.F, gu0, gu1);
  }
  if (!APPLY) {
    red[0][warp][2 * lane] = a0; red[0][warp][2 * lane + 1] = a1;
    red[1][warp][2 * lane] = q0; red[1][warp][2 * lane + 1] = q1;
    __syncthreads();
    if (threadIdx.x < 128) {
      const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < EC_WARPS; ++w) t += red[which][w][c];
      const int f = blockIdx.y * 64 + c;
      if (f < a.F) atomicAdd(&acc[which * a.F + f], (double)t);
    }
  }
}

// zero only the v half of g_uv [P,2F] (u half is fully overwritten by bwd_apply)
__global__ void zero_vhalf_kernel(float* __restrict__ guv, int64_t P, int F) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P * F) return;
  guv[(e / F) * 2 * F + F + (e % F)] = 0.f;
}

int stats_acc_reset(void* ws, int C, cudaStream_t st) {
  if (cudaMemsetAsync(ws, 0, (size_t)2 * C * sizeof(double), st) != cudaSuccess)
    return set_err(DGCNN_ERR_CUDA, "stats: memset failed");
  return DGCNN_OK;
}
int launch_finalize_stats(const double* acc, int C, double count, float eps, float* mean, float* rstd, cudaStream_t st) {
  finalize_stats_kernel<<<cdiv(C, 128), 128, 0, st>>>(acc, C, count, eps, mean, rstd);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("finalize_stats_kernel");
  return DGCNN_OK;
}
int launch_finalize_sums(const double* acc, int C, float* s1, float* s2, cudaStream_t st) {
  finalize_sums_kernel<<<cdiv(C, 128), 128, 0, st>>>(acc, C, s1, s2);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("finalize_sums_kernel");
  return DGCNN_OK;
}

static inline int stat_blocks(int P) {
  int nb = num_sms() * STAT_BLOCKS_PER_SM;
  const int need = cdiv(P, EC_WARPS);
  return nb < need ? nb : need;
}

static int ec_check(const float* uv, const int32_t* idx, int B, int N, int F, int k) {
  DG_REQUIRE(uv && idx, DGCNN_ERR_INVALID, "edgeconv: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && F > 0, DGCNN_ERR_INVALID, "edgeconv: bad shape B=%d N=%d F=%d", B, N, F);
  DG_REQUIRE(k >= 1 && k <= N, DGCNN_ERR_INVALID, "edgeconv: need 1 <= k <= N (k=%d N=%d)", k, N);
  DG_REQUIRE(k <= DGCNN_KNN_MAX_K, DGCNN_ERR_UNSUPPORTED, "edgeconv: k=%d > %d", k, DGCNN_KNN_MAX_K);
  DG_REQUIRE((int64_t)B * N < (1ll << 31) / 2, DGCNN_ERR_UNSUPPORTED, "edgeconv: B*N too large");
  return DGCNN_OK;
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" int dgcnn_edge_feature(const float* x, const int32_t* idx, float* out, int B, int N, int C, int k,
                                  dgcnn_stream_t stream) {
  DG_REQUIRE(x && idx && out, DGCNN_ERR_INVALID, "edge_feature: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && C > 0 && k > 0, DGCNN_ERR_INVALID, "edge_feature: bad shape");
  const int64_t total = (int64_t)B * N * k * 2 * C;
  const int64_t blocks = (total + 255) / 256;
  DG_REQUIRE(blocks < (1ll << 31), DGCNN_ERR_UNSUPPORTED, "edge_feature: too many elements");
  edge_feature_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, idx, out, N, C, k, total);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("edge_feature_kernel");
  return DGCNN_OK;
}

extern "C" int dgcnn_edge_feature_bwd(const float* g_out, const int32_t* idx, float* gx, int B, int N, int C, int k,
                                      dgcnn_stream_t stream) {
  DG_REQUIRE(g_out && idx && gx, DGCNN_ERR_INVALID, "edge_feature_bwd: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && C > 0 && k > 0, DGCNN_ERR_INVALID, "edge_feature_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(gx, 0, (size_t)B * N * C * sizeof(float), st) != cudaSuccess)
    return set_err(DGCNN_ERR_CUDA, "edge_feature_bwd: memset failed");
  const int64_t total = (int64_t)B * N * k * C;
  const int64_t blocks = (total + 255) / 256;
  DG_REQUIRE(blocks < (1ll << 31), DGCNN_ERR_UNSUPPORTED, "edge_feature_bwd: too many elements");
  edge_feature_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(g_out, idx, gx, N, C, k, total);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("edge_feature_bwd_kernel");
  return DGCNN_OK;
}

extern "C" size_t dgcnn_edgeconv_workspace_bytes(int F) {
  if (F <= 0) return 0;
  return (size_t)2 * F * sizeof(double);
}

extern "C" int dgcnn_edgeconv_fwd_stats(const float* uv, const int32_t* idx, int B, int N, int F, int k, float* zmax,
                                        float* cnt, float* mean, float* rstd, void* ws, size_t ws_bytes,
                                        dgcnn_stream_t stream) {
  int rc = ec_check(uv, idx, B, N, F, k);
  if (rc) return rc;
  DG_REQUIRE(zmax && cnt && mean && rstd && ws, DGCNN_ERR_INVALID, "edgeconv_fwd_stats: null pointer");
  DG_REQUIRE(ws_bytes >= dgcnn_edgeconv_workspace_bytes(F), DGCNN_ERR_WORKSPACE, "edgeconv_fwd_stats: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  EcArgs a{uv, idx, B * N, N, F, k};
  const int nb = stat_blocks(a.P);
  dim3 grid(nb, cdiv(F, 64));
  DG_REQUIRE(((uintptr_t)ws & 7) == 0, DGCNN_ERR_INVALID, "edgeconv_fwd_stats: workspace must be 8-byte aligned");
  rc = stats_acc_reset(ws, F, st);
  if (rc) return rc;
  ec_fwd_stats_kernel<<<grid, EC_THREADS, 0, st>>>(a, zmax, cnt, (double*)ws);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("ec_fwd_stats_kernel");
  return launch_finalize_stats((const double*)ws, F, (double)a.P * (double)k, 1e-3f, mean, rstd, st);
}

static int ec_fwd_apply_impl(const float* uv, const int32_t* idx, int B, int N, int F, int k, const float* zmax,
                             const float* mean, const float* rstd, const float* beta, float* out_max, float* out_mean,
                             int pitch, void* sink, int sink_ld, int64_t sink_plane, dgcnn_stream_t stream) {
  int rc = ec_check(uv, idx, B, N, F, k);
  if (rc) return rc;
  DG_REQUIRE(zmax && mean && rstd && beta && out_max && out_mean, DGCNN_ERR_INVALID,
             "edgeconv_fwd_apply: null pointer");
  EcArgs a{uv, idx, B * N, N, F, k};
  dim3 grid(stat_blocks(a.P), cdiv(F, 64));
  ec_fwd_apply_kernel<<<grid, EC_THREADS, 0, (cudaStream_t)stream>>>(a, zmax, mean, rstd, beta, out_max, out_mean,
                                                                     pitch, (__nv_bfloat16*)sink, sink_ld,
                                                                     (size_t)sink_plane);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("ec_fwd_apply_kernel");
  return DGCNN_OK;
}

extern "C" int dgcnn_edgeconv_fwd_apply(const float* uv, const int32_t* idx, int B, int N, int F, int k,
                                        const float* zmax, const float* mean, const float* rstd, const float* beta,
                                        float* out_max, float* out_mean, dgcnn_stream_t stream) {
  return ec_fwd_apply_impl(uv, idx, B, N, F, k, zmax, mean, rstd, beta, out_max, out_mean, F, nullptr, 0, 0, stream);
}

extern "C" int dgcnn_edgeconv_fwd_apply_packed(const float* uv, const int32_t* idx, int B, int N, int F, int k,
                                               const float* zmax, const float* mean, const float* rstd,
                                               const float* beta, float* out_both, dgcnn_stream_t stream) {
  DG_REQUIRE(out_both, DGCNN_ERR_INVALID, "edgeconv_fwd_apply_packed: null pointer");
  return ec_fwd_apply_impl(uv, idx, B, N, F, k, zmax, mean, rstd, beta, out_both, out_both + F, 2 * F, nullptr, 0, 0,
                           stream);
}

extern "C" int dgcnn_edgeconv_fwd_apply_packed_sink(const float* uv, const int32_t* idx, int B, int N, int F, int k,
                                                    const float* zmax, const float* mean, const float* rstd,
                                                    const float* beta, float* out_both, void* sink_planes, int sink_ld,
                                                    int64_t sink_plane_elems, dgcnn_stream_t stream) {
  DG_REQUIRE(out_both, DGCNN_ERR_INVALID, "edgeconv_fwd_apply_packed_sink: null pointer");
  DG_REQUIRE(!sink_planes || (sink_ld >= 2 * F && sink_plane_elems > 0 && (F & 1) == 0 && (sink_ld & 1) == 0 &&
                              (sink_plane_elems & 1) == 0 && ((uintptr_t)sink_planes & 3) == 0),
             DGCNN_ERR_INVALID, "edgeconv_fwd_apply_packed_sink: bad sink geometry (F, pitch, plane distance must be even)");
  return ec_fwd_apply_impl(uv, idx, B, N, F, k, zmax, mean, rstd, beta, out_both, out_both + F, 2 * F, sink_planes, sink_ld,
                           sink_plane_elems, stream);
}

extern "C" int dgcnn_edgeconv_bwd_stats(const float* uv, const int32_t* idx, int B, int N, int F, int k,
                                        const float* zmax, const float* cnt, const float* mean, const float* rstd,
                                        const float* beta, const float* g_max, const float* g_mean, float* s1,
                                        float* s2, void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  DG_REQUIRE(g_max && g_mean, DGCNN_ERR_INVALID, "edgeconv_bwd_stats: null pointer");
  return dgcnn_edgeconv_bwd_stats_packed(uv, idx, B, N, F, k, zmax, cnt, mean, rstd, beta, g_max, g_mean, nullptr, s1,
                                         s2, ws, ws_bytes, stream);
}

extern "C" int dgcnn_edgeconv_bwd_stats_packed(const float* uv, const int32_t* idx, int B, int N, int F, int k,
                                               const float* zmax, const float* cnt, const float* mean,
                                               const float* rstd, const float* beta, const float* g_max,
                                               const float* g_mean, const float* g_both, float* s1, float* s2,
                                               void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  return dgcnn_edgeconv_bwd_stats_packed_z(uv, idx, B, N, F, k, zmax, cnt, mean, rstd, beta, g_max, g_mean, g_both, s1, s2,
                                           nullptr, ws, ws_bytes, stream);
}

extern "C" int dgcnn_edgeconv_bwd_stats_packed_z(const float* uv, const int32_t* idx, int B, int N, int F, int k,
                                                 const float* zmax, const float* cnt, const float* mean,
                                                 const float* rstd, const float* beta, const float* g_max,
                                                 const float* g_mean, const float* g_both, float* s1, float* s2,
                                                 float* g_uv_clear, void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  int rc = ec_check(uv, idx, B, N, F, k);
  if (rc) return rc;
  DG_REQUIRE(zmax && cnt && mean && rstd && beta && s1 && s2 && ws, DGCNN_ERR_INVALID,
             "edgeconv_bwd_stats: null pointer");
  DG_REQUIRE(ws_bytes >= dgcnn_edgeconv_workspace_bytes(F), DGCNN_ERR_WORKSPACE, "edgeconv_bwd_stats: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  EcArgs a{uv, idx, B * N, N, F, k};
  const int nb = stat_blocks(a.P);
  dim3 grid(nb, cdiv(F, 64));
  DG_REQUIRE(((uintptr_t)ws & 7) == 0, DGCNN_ERR_INVALID, "edgeconv_bwd_stats: workspace must be 8-byte aligned");
  rc = stats_acc_reset(ws, F, st);
  if (rc) return rc;
  ec_bwd_kernel<false><<<grid, EC_THREADS, 0, st>>>(a, zmax, cnt, mean, rstd, beta, g_max, g_mean, g_both, nullptr,
                                                     nullptr, (double*)ws, g_uv_clear);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("ec_bwd_kernel<stats>");
  return launch_finalize_sums((const double*)ws, F, s1, s2, st);
}

extern "C" int dgcnn_edgeconv_bwd_apply(const float* uv, const int32_t* idx, int B, int N, int F, int k,
                                        const float* zmax, const float* cnt, const float* mean, const float* rstd,
                                        const float* beta, const float* g_max, const float* g_mean, const float* s1,
                                        const float* s2, float* g_uv, dgcnn_stream_t stream) {
  DG_REQUIRE(g_max && g_mean, DGCNN_ERR_INVALID, "edgeconv_bwd_apply: null pointer");
  return dgcnn_edgeconv_bwd_apply_packed(uv, idx, B, N, F, k, zmax, cnt, mean, rstd, beta, g_max, g_mean, nullptr, s1,
                                         s2, g_uv, stream);
}

extern "C" int dgcnn_edgeconv_bwd_apply_packed(const float* uv, const int32_t* idx, int B, int N, int F, int k,
                                               const float* zmax, const float* cnt, const float* mean,
                                               const float* rstd, const float* beta, const float* g_max,
                                               const float* g_mean, const float* g_both, const float* s1,
                                               const float* s2, float* g_uv, dgcnn_stream_t stream) {
  return dgcnn_edgeconv_bwd_apply_packed_z(uv, idx, B, N, F, k, zmax, cnt, mean, rstd, beta, g_max, g_mean, g_both, s1, s2,
                                           g_uv, 0, stream);
}

extern "C" int dgcnn_edgeconv_bwd_apply_packed_z(const float* uv, const int32_t* idx, int B, int N, int F, int k,
                                                 const float* zmax, const float* cnt, const float* mean,
                                                 const float* rstd, const float* beta, const float* g_max,
                                                 const float* g_mean, const float* g_both, const float* s1,
                                                 const float* s2, float* g_uv, int v_half_cleared,
                                                 dgcnn_stream_t stream) {
  int rc = ec_check(uv, idx, B, N, F, k);
  if (rc) return rc;
  DG_REQUIRE(zmax && cnt && mean && rstd && beta && s1 && s2 && g_uv, DGCNN_ERR_INVALID,
             "edgeconv_bwd_apply: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  EcArgs a{uv, idx, B * N, N, F, k};
  if (!v_half_cleared) {
    const int64_t PF = (int64_t)a.P * F;
    zero_vhalf_kernel<<<cdiv(PF, 256), 256, 0, st>>>(g_uv, a.P, F);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("zero_vhalf_kernel");
  }
  dim3 grid(stat_blocks(a.P), cdiv(F, 64));
  ec_bwd_kernel<true><<<grid, EC_THREADS, 0, st>>>(a, zmax, cnt, mean, rstd, beta, g_max, g_mean, g_both, s1, s2,
                                                    nullptr, g_uv);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("ec_bwd_kernel<apply>");
  return DGCNN_OK;
}
