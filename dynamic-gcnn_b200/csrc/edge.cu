// EdgeConv gather path.
//   edge_feature(+bwd): /root/reference/dgcnn/ops.py:21-40 materialised (API parity for edges()).
//   edgeconv_*        : ops.py:45-57 fused -- gather -> conv0 -> BN(train) -> ReLU -> max_k / mean_k with
//                       z_ij = u_i + v_{idx(i,j)} (uv = x.[Wa-Wb | Wb], formed by dgcnn_gemm), so neither the
//                       [B,N,k,2C] edge tensor nor the [B,N,k,F] activation ever exists in HBM.
// Three gather passes per layer (two forward, one backward), each bound by the L2 gather rate of the v rows
// (E*F*sizeof(T) bytes per pass; the uv table is L2 resident):
//   one warp per point; a HALF-warp covers one 64-channel v row with 16-byte lanes (4 channels per lane), so one load
//   instruction fetches two neighbours; all ceil(k/2) loads of a point are issued back to back into registers before
//   any arithmetic, and the next point's index list is prefetched meanwhile.  Everything that needs several looks at
//   the k values of a point (max, tie count, ReLU mask, BN backward) happens on those registers:
//     fwd_stats : sum_j v, sum_j v^2 per point -> BN batch statistics of z (pivot-shifted, fp64 accumulation)
//     fwd_apply : max_j z, mean_j relu(bn(z)), #active edges  -> (max | mean) [P,2F] (+ bf16 operand planes)
//     bwd_stats : NO gather: sum g_pre and sum g_pre*zhat follow from per-point quantities (the forward outputs, the
//                 active-edge counts and the incoming gradients), one streaming pass over [P,F]
//     bwd_apply : recompute z, max, ties, mask from the gathered rows; g_z scattered with 16-byte vector atomics
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace dgcnn {

constexpr int EC_THREADS = 256;          // 8 warps
constexpr int EC_WARPS = EC_THREADS / 32;
#ifndef EC_MIN_BLOCKS_V
#define EC_MIN_BLOCKS_V 3
#define EC_MIN_BLOCKS_BWD_V 3
#define EC_GROUP_V 5
#endif
constexpr int EC_MIN_BLOCKS = EC_MIN_BLOCKS_V;         // register cap of the gather kernels: 3 x 8 warps per SM
constexpr int EC_MIN_BLOCKS_BWD = EC_MIN_BLOCKS_BWD_V;     // the backward pass keeps more state (no spills at 128 registers)
constexpr int EC_GROUP = EC_GROUP_V;              // gathered rows in flight per half-warp

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
template <typename T>
struct EcArgs {
  const T* uv;          // [P, 2F]   u | v   (fp32, or bf16 for the reduced-precision variant)
  const int32_t* idx;   // [P, k]    neighbour index inside the cloud
  int P, N, F, k;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4(const __nv_bfloat16* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xffff0000u));
}
__device__ __forceinline__ float4 ldg4(const __half* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float ldg1(const __half* p) {
  return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(p))));
}
__device__ __forceinline__ float ldg1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldg1(const __nv_bfloat16* p) {
  return __uint_as_float((unsigned)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}

__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator*(float a, float4 b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 splat4(float a) { return make_float4(a, a, a, a); }
__device__ __forceinline__ float4 xor16(float4 a) {
  return make_float4(__shfl_xor_sync(FULL, a.x, 16), __shfl_xor_sync(FULL, a.y, 16), __shfl_xor_sync(FULL, a.z, 16),
                     __shfl_xor_sync(FULL, a.w, 16));
}

// A lane owns the channel QUAD [f0, f0+4), f0 = 64*chunk + 4*(lane & 15); the two half-warps own the same channels and
// split the neighbours.  vec (F % 4 == 0): every row start is 16-byte aligned (8 bytes for bf16) and the quad is whole
// or empty; otherwise element-wise accesses with guards.
template <bool VEC>
struct EcQuad {
  int f0, nv;   // first channel of the quad, number of valid channels in it (VEC: 0 or 4)
  int fl;       // load offset: f0, or 0 for an empty quad (VEC loads are unconditional; their result is never used)
  static constexpr bool vec = VEC;
  __device__ __forceinline__ EcQuad(int F, int chunk, int lane16) {
    f0 = chunk * 64 + 4 * lane16;
    const int r = F - f0;
    nv = r < 0 ? 0 : (r > 4 ? 4 : r);
    fl = nv ? f0 : 0;
  }
  template <typename T>
  __device__ __forceinline__ float4 ld(const T* __restrict__ row) const {
    if constexpr (VEC) {
      return ldg4(row + fl);
    } else {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (nv > 0) o.x = ldg1(row + f0);
      if (nv > 1) o.y = ldg1(row + f0 + 1);
      if (nv > 2) o.z = ldg1(row + f0 + 2);
      if (nv > 3) o.w = ldg1(row + f0 + 3);
      return o;
    }
  }
  // row `r` of a table whose rows are `pitch` elements apart (one 32x32->64 multiply-add per address)
  template <typename T>
  __device__ __forceinline__ float4 ld_row(const T* __restrict__ base, int r, int pitch) const {
    if constexpr (VEC) {
      const char* p = reinterpret_cast<const char*>(base + fl) + (uint64_t)(uint32_t)r * (uint32_t)(pitch * (int)sizeof(T));
      return ldg4(reinterpret_cast<const T*>(p));
    } else {
      return ld(base + (int64_t)r * pitch);
    }
  }
  __device__ __forceinline__ void st(float* __restrict__ row, float4 v) const {
    if constexpr (VEC) {
      if (nv) *reinterpret_cast<float4*>(row + f0) = v;
    } else {
      if (nv > 0) row[f0] = v.x;
      if (nv > 1) row[f0 + 1] = v.y;
      if (nv > 2) row[f0 + 2] = v.z;
      if (nv > 3) row[f0 + 3] = v.w;
    }
  }
  __device__ __forceinline__ void red_row(float* __restrict__ base, int r, int pitch, float4 v) const {
    if constexpr (VEC) {
      char* p = reinterpret_cast<char*>(base + fl) + (uint64_t)(uint32_t)r * (uint32_t)(pitch * 4);
      if (nv) atomicAdd(reinterpret_cast<float4*>(p), v);          // one 16-byte RED per lane
    } else {
      float* row = base + (int64_t)r * pitch;
      if (nv > 0) atomicAdd(row + f0, v.x);
      if (nv > 1) atomicAdd(row + f0 + 1, v.y);
      if (nv > 2) atomicAdd(row + f0 + 2, v.z);
      if (nv > 3) atomicAdd(row + f0 + 3, v.w);
    }
  }
  __device__ __forceinline__ uchar4 ld_u8(const uint8_t* __restrict__ row) const {
    if constexpr (VEC) {
      return __ldg(reinterpret_cast<const uchar4*>(row + fl));
    } else {
      uchar4 o = make_uchar4(0, 0, 0, 0);
      if (nv > 0) o.x = __ldg(row + f0);
      if (nv > 1) o.y = __ldg(row + f0 + 1);
      if (nv > 2) o.z = __ldg(row + f0 + 2);
      if (nv > 3) o.w = __ldg(row + f0 + 3);
      return o;
    }
  }
  __device__ __forceinline__ void st_u8(uint8_t* __restrict__ row, uchar4 v) const {
    if constexpr (VEC) {
      if (nv) *reinterpret_cast<uchar4*>(row + f0) = v;
    } else {
      if (nv > 0) row[f0] = v.x;
      if (nv > 1) row[f0 + 1] = v.y;
      if (nv > 2) row[f0 + 2] = v.z;
      if (nv > 3) row[f0 + 3] = v.w;
    }
  }
};

// Walks the points of one warp (p, p + stride, ...) keeping the first row of p's cloud without a division per point.
struct CloudWalk {
  int p, base, sq, sr, N;
  __device__ __forceinline__ CloudWalk(int p0, int stride, int N_) : p(p0), N(N_) {
    base = (p0 / N_) * N_;
    sq = (stride / N_) * N_;
    sr = stride % N_;
  }
  __device__ __forceinline__ void next(int stride) {
    p += stride;
    base += sq;
    if (p - base >= N) base += N;      // (p - old base) < N and sr < N: at most one more cloud
    (void)sr;
  }
};

// this point's neighbour rows (global point index), one per lane (r0: j = lane, r1: j = 32 + lane)
template <typename T, int NB>
__device__ __forceinline__ void load_rows(const EcArgs<T>& a, int p, int base, int lane, int& r0, int& r1) {
  const int32_t* ip = a.idx + (int64_t)p * a.k;
  r0 = lane < a.k ? base + __ldg(ip + lane) : 0;
  r1 = 0;
  if (NB > 16) r1 = (32 + lane) < a.k ? base + __ldg(ip + 32 + lane) : 0;
}

// slot t of half-warp h is neighbour j = 2t + h.  The slots are walked in groups of G: the G loads of a group are issued
// back to back, then consumed by fn(v, row) -- (G x 512) bytes in flight per warp, few live registers, many warps.
// EXACT: k == 2 NB, every slot of both half-warps is live and nothing needs a predicate.
template <typename T, int NB, int G, bool EXACT, bool VEC, typename Fn>
__device__ __forceinline__ void for_each_nbr(const EcArgs<T>& a, const EcQuad<VEC>& q, int r0, int r1, int half, int kk,
                                             Fn&& fn) {
  const T* vbase = a.uv + a.F;
  const int pitch = 2 * a.F;
#pragma unroll
  for (int g0 = 0; g0 < NB; g0 += G) {
    float4 v[G];
    int rr[G];
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int t = g0 + i;
      rr[i] = 0;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < NB && (EXACT || 2 * t < a.k)) {              // warp-uniform
        const int j = 2 * t + half;
        rr[i] = (t < 16) ? __shfl_sync(FULL, r0, j & 31) : __shfl_sync(FULL, r1, j & 31);
        if (EXACT || t < kk) v[i] = q.ld_row(vbase, rr[i], pitch);
      }
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int t = g0 + i;
      if (t < NB && (EXACT || t < kk)) fn(v[i], rr[i]);
    }
  }
}

// per-warp double partials [4] of both halves -> shared -> one fp64 atomic per channel and block
__device__ __forceinline__ void block_acc_2x4(double (&s)[4], double (&qq)[4], double (*red)[EC_WARPS][64], int lane,
                                              int warp, int chunk, int F, double* __restrict__ acc) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    s[c] += __shfl_xor_sync(FULL, s[c], 16);
    qq[c] += __shfl_xor_sync(FULL, qq[c], 16);
  }
  if (lane < 16) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      red[0][warp][4 * lane + c] = s[c];
      red[1][warp][4 * lane + c] = qq[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < EC_WARPS; ++w) t += red[which][w][c];
    const int f = chunk * 64 + c;
    if (f < F) atomicAdd(&acc[which * F + f], t);
  }
}

// forward pass 1: BN batch statistics of z_ij = u_i + v_j over all edges.  Per point only sum_j v and sum_j v^2 are
// needed (sum_j z = k u + sum v ; sum_j z^2 = k u^2 + 2 u sum v + sum v^2).  Both u and v are shifted by the first point's
// values (pivot) before squaring and every per-point total goes into an fp64 accumulator, so E[z^2] - E[z]^2 does not
// cancel for channels whose mean is large against their spread.
template <typename T, int NB, int G, bool EXACT, bool VEC>
__global__ void __launch_bounds__(EC_THREADS, EC_MIN_BLOCKS)
    ec_fwd_stats_kernel(EcArgs<T> a, double* __restrict__ acc, float* __restrict__ pivot) {
  __shared__ double red[2][EC_WARPS][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4;
  const EcQuad<VEC> q(a.F, blockIdx.y, lane & 15);
  const float4 cu = q.ld(a.uv), cv = q.ld(a.uv + a.F);
  if (blockIdx.x == 0 && warp == 0 && half == 0) q.st(pivot, cu + cv);
  const int kk = (a.k - half + 1) >> 1;                  // slots of this half-warp
  const float kf = half == 0 ? (float)a.k : 0.f;        // the u terms are counted once (by half 0)
  double s[4] = {0.0, 0.0, 0.0, 0.0}, qq[4] = {0.0, 0.0, 0.0, 0.0};
  const int stride = gridDim.x * EC_WARPS;
  CloudWalk cw(blockIdx.x * EC_WARPS + warp, stride, a.N);
  int r0 = 0, r1 = 0;
  if (cw.p < a.P) load_rows<T, NB>(a, cw.p, cw.base, lane, r0, r1);
  while (cw.p < a.P) {
    const float4 u = q.ld(a.uv + (int64_t)cw.p * 2 * a.F);
    float4 sv = splat4(0.f), qv = splat4(0.f);
    for_each_nbr<T, NB, G, EXACT, VEC>(a, q, r0, r1, half, kk, [&](float4 v, int) {
      const float4 d = v - cv;
      sv = sv + d;
      qv = fma4(d, d, qv);
    });
    cw.next(stride);
    if (cw.p < a.P) load_rows<T, NB>(a, cw.p, cw.base, lane, r0, r1);
    const float4 ud = u - cu;
    const float4 ps = fma4(splat4(kf), ud, sv);
    const float4 pq = fma4(2.f * ud, sv, fma4(splat4(kf), ud * ud, qv));
    s[0] += (double)ps.x; s[1] += (double)ps.y; s[2] += (double)ps.z; s[3] += (double)ps.w;
    qq[0] += (double)pq.x; qq[1] += (double)pq.y; qq[2] += (double)pq.z; qq[3] += (double)pq.w;
  }
  block_acc_2x4(s, qq, red, lane, warp, blockIdx.y, a.F, acc);
}

// Per-channel totals are accumulated by the statistics kernels with fp64 atomics into acc[2][C] (sum, sum of
// squares / cross term); fp64 makes the summation order irrelevant at fp32 resolution.  These kernels turn the
// totals into what the apply passes need.  pivot (optional): the totals are of (z - pivot).
__global__ void finalize_stats_kernel(const double* __restrict__ acc, int C, double count, float eps,
                                      const float* __restrict__ pivot, float* __restrict__ mean,
                                      float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc[c] / count;
  double var = acc[C + c] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)(m + (pivot ? (double)pivot[c] : 0.0));
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void finalize_sums_kernel(const double* __restrict__ acc, int C, float* __restrict__ s1,
                                     float* __restrict__ s2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  s1[c] = (float)acc[c];
  s2[c] = (float)acc[C + c];
}

// bf16 hi / lo planes of 4 consecutive values (the tensor-core operand format of tc_gemm*.cu)
__device__ __forceinline__ void put_planes4(__nv_bfloat16* __restrict__ hi, size_t plane, float4 x) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(x.x - f0.x, x.y - f0.y), l1 = __floats2bfloat162_rn(x.z - f1.x, x.w - f1.y);
  uint2 hv, lv;
  hv.x = *reinterpret_cast<const unsigned*>(&h0); hv.y = *reinterpret_cast<const unsigned*>(&h1);
  lv.x = *reinterpret_cast<const unsigned*>(&l0); lv.y = *reinterpret_cast<const unsigned*>(&l1);
  *reinterpret_cast<uint2*>(hi) = hv;
  if (plane) *reinterpret_cast<uint2*>(hi + plane) = lv;
}

// forward pass 2: y_ij = relu(r z_ij + b'), b' = beta - mean r (BN folded into one FMA; r = rstd > 0, no gamma)
//   both[p] = ( relu(r max_j z + b') | mean_j y_ij )   -- BN(+)ReLU are monotone, so the max commutes with them
//   zmax[p] = max_j z_ij, npos[p] = #{j : y_ij > 0}   (optional, both or neither: what the backward passes need)
template <typename T, int NB, int G, bool EXACT, bool VEC>
__global__ void __launch_bounds__(EC_THREADS, EC_MIN_BLOCKS)
    ec_fwd_apply_kernel(EcArgs<T> a, const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ beta, float* __restrict__ both, float* __restrict__ zmax,
                        uint8_t* __restrict__ npos, __nv_bfloat16* __restrict__ sink, int sink_ld, size_t sink_plane) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4;
  const EcQuad<VEC> q(a.F, blockIdx.y, lane & 15);
  const float4 r = q.ld(rstd);
  const float4 bp = fma4(splat4(-1.f) * q.ld(mean), r, q.ld(beta));
  const int kk = (a.k - half + 1) >> 1;
  const float invk = 1.0f / (float)a.k;
  const int stride = gridDim.x * EC_WARPS;
  CloudWalk cw(blockIdx.x * EC_WARPS + warp, stride, a.N);
  int r0 = 0, r1 = 0;
  if (cw.p < a.P) load_rows<T, NB>(a, cw.p, cw.base, lane, r0, r1);
  while (cw.p < a.P) {
    const int p = cw.p;
    const float4 u = q.ld(a.uv + (int64_t)p * 2 * a.F);
    float4 m = splat4(-INFINITY), ys = splat4(0.f);
    int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
    for_each_nbr<T, NB, G, EXACT, VEC>(a, q, r0, r1, half, kk, [&](float4 v, int) {
      const float4 z = u + v;
      m = max4(m, z);
      const float4 y = fma4(z, r, bp);
      ys = ys + max4(y, splat4(0.f));
      n0 += y.x > 0.f; n1 += y.y > 0.f; n2 += y.z > 0.f; n3 += y.w > 0.f;
    });
    cw.next(stride);
    if (cw.p < a.P) load_rows<T, NB>(a, cw.p, cw.base, lane, r0, r1);
    m = max4(m, xor16(m));
    ys = ys + xor16(ys);
    const float4 omax = max4(fma4(m, r, bp), splat4(0.f));
    const float4 omean = invk * ys;
    float* o = both + (int64_t)p * 2 * a.F;
    if (half == 0) q.st(o, omax); else q.st(o + a.F, omean);
    if (npos != nullptr) {
      n0 += __shfl_xor_sync(FULL, n0, 16); n1 += __shfl_xor_sync(FULL, n1, 16);
      n2 += __shfl_xor_sync(FULL, n2, 16); n3 += __shfl_xor_sync(FULL, n3, 16);
      if (half == 0) q.st_u8(npos + (int64_t)p * a.F, make_uchar4(n0, n1, n2, n3));
      else q.st(zmax + (int64_t)p * a.F, m);
    }
    // optional plane sink: the same values as bf16 hi / lo planes at columns [0,F) (max) and [F,2F) (mean) of a
    // tensor-core operand (the consumer's concat operand), so that no separate split pass reads them again
    if (sink != nullptr && q.nv) {
      __nv_bfloat16* d = sink + (size_t)p * sink_ld + q.f0;
      if (half == 0) put_planes4(d, sink_plane, omax); else put_planes4(d + a.F, sink_plane, omean);
    }
  }
}

// backward statistics without a gather.  With g_y_ij = g_mean_i/k + [z_ij == zmax_i] g_max_i / cnt_i (tf reduce_mean /
// reduce_max gradients, ties share) and g_pre_ij = g_y_ij [y_ij > 0] (ReluGrad), zhat_ij = y_ij - beta:
//   s1 = sum g_pre        = sum_i  g_mean_i/k * npos_i                      + [omax_i > 0] g_max_i
//   s2 = sum g_pre * zhat = sum_i  g_mean_i/k * (k omean_i - beta npos_i)   + [omax_i > 0] g_max_i (omax_i - beta)
// (every tie of the max has the same y, so the tie shares add up to g_max).  Also clears the v half of g_uv for the
// scatter-add of the apply pass.  Thread = (row, channel quad); 16 rows x 16 quads per block step.
template <bool VEC>
__global__ void __launch_bounds__(EC_THREADS)
    ec_bwd_stats_kernel(const float* __restrict__ both, const uint8_t* __restrict__ npos, const float* __restrict__ beta,
                        const float* __restrict__ gmax, const float* __restrict__ gmean, const float* __restrict__ gboth,
                        int P, int F, int k, float* __restrict__ guv_clear, double* __restrict__ acc) {
  __shared__ double red[2][16][64];
  const int l16 = threadIdx.x & 15, rsub = threadIdx.x >> 4;
  const EcQuad<VEC> q(F, blockIdx.y, l16);
  const float4 be = q.ld(beta);
  const float invk = 1.0f / (float)k, kf = (float)k;
  double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
  for (int p = blockIdx.x * 16 + rsub; p < P; p += gridDim.x * 16) {
    const int64_t o = (int64_t)p * F, o2 = (int64_t)p * 2 * F;
    float4 gM = splat4(0.f), gA = splat4(0.f);
    if (gmax) gM = q.ld(gmax + o);
    if (gmean) gA = q.ld(gmean + o);
    if (gboth) {
      gM = gM + q.ld(gboth + o2);
      gA = gA + q.ld(gboth + o2 + F);
    }
    const float4 omax = q.ld(both + o2), omean = q.ld(both + o2 + F);
    const uchar4 nb = q.ld_u8(npos + o);
    const float4 np = make_float4((float)nb.x, (float)nb.y, (float)nb.z, (float)nb.w);
    const float4 gm = invk * gA;
    const float4 gx = make_float4(omax.x > 0.f ? gM.x : 0.f, omax.y > 0.f ? gM.y : 0.f, omax.z > 0.f ? gM.z : 0.f,
                                  omax.w > 0.f ? gM.w : 0.f);
    const float4 t1 = fma4(gm, np, gx);
    const float4 t2 = fma4(gm, kf * omean - be * np, gx * (omax - be));
    s1[0] += (double)t1.x; s1[1] += (double)t1.y; s1[2] += (double)t1.z; s1[3] += (double)t1.w;
    s2[0] += (double)t2.x; s2[1] += (double)t2.y; s2[2] += (double)t2.z; s2[3] += (double)t2.w;
    if (guv_clear != nullptr) q.st(guv_clear + o2 + F, splat4(0.f));
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    red[0][rsub][4 * l16 + c] = s1[c];
    red[1][rsub][4 * l16 + c] = s2[c];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 16; ++w) t += red[which][w][c];
    const int f = blockIdx.y * 64 + c;
    if (f < F) atomicAdd(&acc[which * F + f], t);
  }
}

// backward gather / scatter pass (the only one).  z and the ReLU mask are recomputed from the gathered rows, the max
// comes from the forward pass (zmax):
//   g_pre_ij = [y_ij > 0] (g_mean_i/k + [z_ij == zmax_i] g_max_i / cnt_i)            (ties share the max gradient)
//   g_z_ij = r (g_pre_ij - s1/E - zhat_ij s2/E) = r g_pre_ij - c0 - c1 z_ij ,  c1 = r^2 s2/E, c0 = r s1/E - c1 mean
//   g_u_i  = sum_j g_z_ij (stored) ;  g_v_n += g_z_ij for n = idx(i,j) (16-byte vector atomics: tf.gather's scatter-add)
// The L2 atomic units set the floor of this pass (E*F*4 bytes of fp32 adds, ~5 TB/s: profiles/scripts/scatter_bench.cu).
// The tie count cnt_i is taken as 1 while the edges stream through; the ties are counted on the way and points that
// turn out to have an exact tie in some channel (duplicate points) get a correction pass over their k edges.
template <typename T, int NB, int G, bool EXACT, bool VEC>
__global__ void __launch_bounds__(EC_THREADS, EC_MIN_BLOCKS_BWD)
    ec_bwd_apply_kernel(EcArgs<T> a, const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ beta, const float* __restrict__ zmax, const float* __restrict__ gmax,
                        const float* __restrict__ gmean, const float* __restrict__ gboth, const float* __restrict__ s1,
                        const float* __restrict__ s2, float* __restrict__ guv) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4;
  const EcQuad<VEC> q(a.F, blockIdx.y, lane & 15);
  const float4 r = q.ld(rstd), mu = q.ld(mean);
  const float4 bp = fma4(splat4(-1.f) * mu, r, q.ld(beta));
  const float invE = 1.0f / ((float)a.P * (float)a.k);
  const float4 c1 = r * r * (invE * q.ld(s2));
  const float4 nc0 = c1 * mu - r * (invE * q.ld(s1));     // -c0
  const float4 nc1 = splat4(-1.f) * c1;
  const int kk = (a.k - half + 1) >> 1;
  const float invk = 1.0f / (float)a.k;
  const int stride = gridDim.x * EC_WARPS;
  float* vg = guv + a.F;
  const int pitch = 2 * a.F;
  CloudWalk cw(blockIdx.x * EC_WARPS + warp, stride, a.N);
  int r0 = 0, r1 = 0;
  if (cw.p < a.P) load_rows<T, NB>(a, cw.p, cw.base, lane, r0, r1);
  while (cw.p < a.P) {
    const int p = cw.p;
    const int64_t o = (int64_t)p * a.F, o2 = (int64_t)p * 2 * a.F;
    const float4 u = q.ld(a.uv + o2);
    const float4 m = q.ld(zmax + o);
    // the gradients of max / mean arrive from up to two consumers: separate [P,F] tensors and/or one packed
    // [P,2F] = (max | mean) tensor (the conv1 operand); summed here instead of in a separate pass
    float4 gM = splat4(0.f), gA = splat4(0.f);
    if (gmax) gM = q.ld(gmax + o);
    if (gmean) gA = q.ld(gmean + o);
    if (gboth) {
      gM = gM + q.ld(gboth + o2);
      gA = gA + q.ld(gboth + o2 + a.F);
    }
    const float4 ym = fma4(m, r, bp);
    const float4 gm = invk * gA;
    // an inactive maximum passes nothing on (then every edge of the channel is inactive)
    gM = make_float4(ym.x > 0.f ? gM.x : 0.f, ym.y > 0.f ? gM.y : 0.f, ym.z > 0.f ? gM.z : 0.f, ym.w > 0.f ? gM.w : 0.f);
    const float4 gt = gm + gM;
    float4 gu = splat4(0.f);
    int t0 = 0, t1 = 0, t2 = 0, t3 = 0;                    // edges attaining the maximum, per channel
    for_each_nbr<T, NB, G, EXACT, VEC>(a, q, r0, r1, half, kk, [&](float4 v, int row) {
      const float4 z = u + v;
      const float4 y = fma4(z, r, bp);
      const bool e0 = z.x == m.x, e1 = z.y == m.y, e2 = z.z == m.z, e3 = z.w == m.w;
      float4 gp;
      gp.x = y.x > 0.f ? (e0 ? gt.x : gm.x) : 0.f;
      gp.y = y.y > 0.f ? (e1 ? gt.y : gm.y) : 0.f;
      gp.z = y.z > 0.f ? (e2 ? gt.z : gm.z) : 0.f;
      gp.w = y.w > 0.f ? (e3 ? gt.w : gm.w) : 0.f;
      t0 += e0; t1 += e1; t2 += e2; t3 += e3;
      const float4 gz = fma4(nc1, z, fma4(r, gp, nc0));
      gu = gu + gz;
      q.red_row(vg, row, pitch, gz);
    });
    t0 += __shfl_xor_sync(FULL, t0, 16); t1 += __shfl_xor_sync(FULL, t1, 16);
    t2 += __shfl_xor_sync(FULL, t2, 16); t3 += __shfl_xor_sync(FULL, t3, 16);
    const bool tied = (t0 > 1 && gM.x != 0.f) || (t1 > 1 && gM.y != 0.f) || (t2 > 1 && gM.z != 0.f) || (t3 > 1 && gM.w != 0.f);
    if (__any_sync(FULL, tied && q.nv > 0)) {
      // exact ties: each of the t edges was given g_max instead of g_max / t; take the excess back
      const float4 ex = make_float4(t0 > 1 ? gM.x * (1.f / (float)t0 - 1.f) : 0.f, t1 > 1 ? gM.y * (1.f / (float)t1 - 1.f) : 0.f,
                                    t2 > 1 ? gM.z * (1.f / (float)t2 - 1.f) : 0.f, t3 > 1 ? gM.w * (1.f / (float)t3 - 1.f) : 0.f);
      for_each_nbr<T, NB, G, EXACT, VEC>(a, q, r0, r1, half, kk, [&](float4 v, int row) {
        const float4 z = u + v;
        float4 d;
        d.x = z.x == m.x ? r.x * ex.x : 0.f;
        d.y = z.y == m.y ? r.y * ex.y : 0.f;
        d.z = z.z == m.z ? r.z * ex.z : 0.f;
        d.w = z.w == m.w ? r.w * ex.w : 0.f;
        if (d.x != 0.f || d.y != 0.f || d.z != 0.f || d.w != 0.f) {
          gu = gu + d;
          q.red_row(vg, row, pitch, d);
        }
      });
    }
    cw.next(stride);
    if (cw.p < a.P) load_rows<T, NB>(a, cw.p, cw.base, lane, r0, r1);
    gu = gu + xor16(gu);
    if (half == 0) q.st(guv + o2, gu);
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
int launch_finalize_stats(const double* acc, int C, double count, float eps, float* mean, float* rstd, cudaStream_t st,
                          const float* pivot) {
  finalize_stats_kernel<<<cdiv(C, 128), 128, 0, st>>>(acc, C, count, eps, pivot, mean, rstd);
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

// persistent grid: as many blocks as are resident at once (every block then does the same number of points)
template <typename K>
static int resident_blocks(K kernel, int want) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, EC_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int nb = num_sms() * per_sm;
  return nb < want ? nb : (want < 1 ? 1 : want);
}

static int ec_check(const void* uv, const int32_t* idx, int B, int N, int F, int k, int dtype) {
  DG_REQUIRE(uv && idx, DGCNN_ERR_INVALID, "edgeconv: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && F > 0, DGCNN_ERR_INVALID, "edgeconv: bad shape B=%d N=%d F=%d", B, N, F);
  DG_REQUIRE(k >= 1 && k <= N, DGCNN_ERR_INVALID, "edgeconv: need 1 <= k <= N (k=%d N=%d)", k, N);
  DG_REQUIRE(k <= DGCNN_KNN_MAX_K, DGCNN_ERR_UNSUPPORTED, "edgeconv: k=%d > %d", k, DGCNN_KNN_MAX_K);
  DG_REQUIRE((int64_t)B * N < (1ll << 31) / 2, DGCNN_ERR_UNSUPPORTED, "edgeconv: B*N too large");
  DG_REQUIRE(dtype == DGCNN_F32 || dtype == DGCNN_BF16 || dtype == DGCNN_F16, DGCNN_ERR_INVALID,
             "edgeconv: uv_dtype must be DGCNN_F32, DGCNN_BF16 or DGCNN_F16");
  DG_REQUIRE(((uintptr_t)uv & 15) == 0, DGCNN_ERR_INVALID, "edgeconv: uv must be 16-byte aligned");
  return DGCNN_OK;
}

// kernel variants: EXACT (k == 2 NB: 20, 40, 64 -- no predicates) needs VEC; otherwise runtime k <= 2 NB, with or
// without 16-byte rows
#define EC_VARIANT(nb, exact, vec, ...)                                  \
  {                                                                      \
    constexpr int NB = nb;                                               \
    constexpr bool EXACT = exact, VEC = vec;                             \
    __VA_ARGS__;                                                         \
  }
#define EC_DISPATCH(k, F, ...)                                           \
  do {                                                                   \
    const bool vec_ = ((F) & 3) == 0;                                    \
    if (vec_ && (k) == 20) EC_VARIANT(10, true, true, __VA_ARGS__)       \
    else if (vec_ && (k) == 40) EC_VARIANT(20, true, true, __VA_ARGS__)  \
    else if (vec_ && (k) == 64) EC_VARIANT(32, true, true, __VA_ARGS__)  \
    else if ((k) <= 20) {                                                \
      if (vec_) EC_VARIANT(10, false, true, __VA_ARGS__)                 \
      else EC_VARIANT(10, false, false, __VA_ARGS__)                     \
    } else if ((k) <= 40) {                                              \
      if (vec_) EC_VARIANT(20, false, true, __VA_ARGS__)                 \
      else EC_VARIANT(20, false, false, __VA_ARGS__)                     \
    } else {                                                             \
      if (vec_) EC_VARIANT(32, false, true, __VA_ARGS__)                 \
      else EC_VARIANT(32, false, false, __VA_ARGS__)                     \
    }                                                                    \
  } while (0)

template <typename T>
static int fwd_stats_t(const void* uv, const int32_t* idx, int B, int N, int F, int k, float* mean, float* rstd, void* ws,
                       cudaStream_t st) {
  EcArgs<T> a{(const T*)uv, idx, B * N, N, F, k};
  double* acc = (double*)ws;
  float* pivot = (float*)(acc + 2 * F);
  int rc = stats_acc_reset(ws, F, st);
  if (rc) return rc;
  EC_DISPATCH(k, F, {
    dim3 grid(resident_blocks(ec_fwd_stats_kernel<T, NB, EC_GROUP, EXACT, VEC>, cdiv(a.P, EC_WARPS)), cdiv(F, 64));
    ec_fwd_stats_kernel<T, NB, EC_GROUP, EXACT, VEC><<<grid, EC_THREADS, 0, st>>>(a, acc, pivot);
  });
  count_launch();
  DG_CUDA_LAUNCH_CHECK("ec_fwd_stats_kernel");
  return launch_finalize_stats(acc, F, (double)a.P * (double)k, 1e-3f, mean, rstd, st, pivot);
}

template <typename T>
static int fwd_apply_t(const void* uv, const int32_t* idx, int B, int N, int F, int k, const float* mean,
                       const float* rstd, const float* beta, float* both, float* zmax, uint8_t* npos, void* sink, int sink_ld,
                       int64_t sink_plane, cudaStream_t st) {
  EcArgs<T> a{(const T*)uv, idx, B * N, N, F, k};
  EC_DISPATCH(k, F, {
    dim3 grid(resident_blocks(ec_fwd_apply_kernel<T, NB, EC_GROUP, EXACT, VEC>, cdiv(a.P, EC_WARPS)), cdiv(F, 64));
    ec_fwd_apply_kernel<T, NB, EC_GROUP, EXACT, VEC><<<grid, EC_THREADS, 0, st>>>(a, mean, rstd, beta, both, zmax, npos, (__nv_bfloat16*)sink,
                                                            sink_ld, (size_t)sink_plane);
  });
  count_launch();
  DG_CUDA_LAUNCH_CHECK("ec_fwd_apply_kernel");
  return DGCNN_OK;
}

template <typename T>
static int bwd_apply_t(const void* uv, const int32_t* idx, int B, int N, int F, int k, const float* mean,
                       const float* rstd, const float* beta, const float* zmax, const float* gmax, const float* gmean, const float* gboth,
                       const float* s1, const float* s2, float* guv, cudaStream_t st) {
  EcArgs<T> a{(const T*)uv, idx, B * N, N, F, k};
  EC_DISPATCH(k, F, {
    dim3 grid(resident_blocks(ec_bwd_apply_kernel<T, NB, EC_GROUP, EXACT, VEC>, cdiv(a.P, EC_WARPS)), cdiv(F, 64));
    ec_bwd_apply_kernel<T, NB, EC_GROUP, EXACT, VEC><<<grid, EC_THREADS, 0, st>>>(a, mean, rstd, beta, zmax, gmax, gmean, gboth, s1, s2, guv);
  });
  count_launch();
  DG_CUDA_LAUNCH_CHECK("ec_bwd_apply_kernel");
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
  return (size_t)2 * F * sizeof(double) + (size_t)F * sizeof(float);   // fp64 totals [2][F] + pivot [F]
}

extern "C" int dgcnn_edgeconv_fwd_stats(const void* uv, int uv_dtype, const int32_t* idx, int B, int N, int F, int k,
                                        float* mean, float* rstd, void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  int rc = ec_check(uv, idx, B, N, F, k, uv_dtype);
  if (rc) return rc;
  DG_REQUIRE(mean && rstd && ws, DGCNN_ERR_INVALID, "edgeconv_fwd_stats: null pointer");
  DG_REQUIRE(ws_bytes >= dgcnn_edgeconv_workspace_bytes(F), DGCNN_ERR_WORKSPACE, "edgeconv_fwd_stats: workspace");
  DG_REQUIRE(((uintptr_t)ws & 7) == 0, DGCNN_ERR_INVALID, "edgeconv_fwd_stats: workspace must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (uv_dtype == DGCNN_BF16) return fwd_stats_t<__nv_bfloat16>(uv, idx, B, N, F, k, mean, rstd, ws, st);
  if (uv_dtype == DGCNN_F16) return fwd_stats_t<__half>(uv, idx, B, N, F, k, mean, rstd, ws, st);
  return fwd_stats_t<float>(uv, idx, B, N, F, k, mean, rstd, ws, st);
}

extern "C" int dgcnn_edgeconv_fwd_apply(const void* uv, int uv_dtype, const int32_t* idx, int B, int N, int F, int k,
                                        const float* mean, const float* rstd, const float* beta, float* out_both,
                                        float* zmax, uint8_t* npos, void* sink_planes, int sink_ld,
                                        int64_t sink_plane_elems, dgcnn_stream_t stream) {
  int rc = ec_check(uv, idx, B, N, F, k, uv_dtype);
  if (rc) return rc;
  DG_REQUIRE(mean && rstd && beta && out_both, DGCNN_ERR_INVALID, "edgeconv_fwd_apply: null pointer");
  DG_REQUIRE(((uintptr_t)out_both & 15) == 0, DGCNN_ERR_INVALID, "edgeconv_fwd_apply: out_both must be 16-byte aligned");
  DG_REQUIRE((zmax == nullptr) == (npos == nullptr), DGCNN_ERR_INVALID, "edgeconv_fwd_apply: zmax and npos go together");
  DG_REQUIRE(((uintptr_t)zmax & 15) == 0 && ((uintptr_t)npos & 3) == 0, DGCNN_ERR_INVALID,
             "edgeconv_fwd_apply: zmax must be 16-byte, npos 4-byte aligned");
  DG_REQUIRE(!sink_planes || (sink_ld >= 2 * F && sink_plane_elems >= 0 && (F & 3) == 0 && (sink_ld & 3) == 0 &&
                              (sink_plane_elems & 3) == 0 && ((uintptr_t)sink_planes & 7) == 0),
             DGCNN_ERR_INVALID, "edgeconv_fwd_apply: bad sink geometry (F, pitch, plane distance must be multiples of 4)");
  cudaStream_t st = (cudaStream_t)stream;
  if (uv_dtype == DGCNN_BF16)
    return fwd_apply_t<__nv_bfloat16>(uv, idx, B, N, F, k, mean, rstd, beta, out_both, zmax, npos, sink_planes, sink_ld,
                                      sink_plane_elems, st);
  if (uv_dtype == DGCNN_F16)
    return fwd_apply_t<__half>(uv, idx, B, N, F, k, mean, rstd, beta, out_both, zmax, npos, sink_planes, sink_ld,
                               sink_plane_elems, st);
  return fwd_apply_t<float>(uv, idx, B, N, F, k, mean, rstd, beta, out_both, zmax, npos, sink_planes, sink_ld,
                            sink_plane_elems, st);
}

extern "C" int dgcnn_edgeconv_bwd_stats(const float* out_both, const uint8_t* npos, const float* beta, const float* g_max,
                                        const float* g_mean, const float* g_both, int B, int N, int F, int k, float* s1,
                                        float* s2, float* g_uv_clear, void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  DG_REQUIRE(out_both && npos && beta && s1 && s2 && ws, DGCNN_ERR_INVALID, "edgeconv_bwd_stats: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && F > 0 && k >= 1 && k <= DGCNN_KNN_MAX_K, DGCNN_ERR_INVALID,
             "edgeconv_bwd_stats: bad shape B=%d N=%d F=%d k=%d", B, N, F, k);
  DG_REQUIRE((int64_t)B * N < (1ll << 31) / 2, DGCNN_ERR_UNSUPPORTED, "edgeconv: B*N too large");
  DG_REQUIRE(ws_bytes >= dgcnn_edgeconv_workspace_bytes(F), DGCNN_ERR_WORKSPACE, "edgeconv_bwd_stats: workspace");
  DG_REQUIRE(((uintptr_t)ws & 7) == 0, DGCNN_ERR_INVALID, "edgeconv_bwd_stats: workspace must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = stats_acc_reset(ws, F, st);
  if (rc) return rc;
  const int P = B * N;
  int nb = num_sms() * 4;
  const int need = cdiv(P, 16);
  if (nb > need) nb = need;
  dim3 grid(nb, cdiv(F, 64));
  if ((F & 3) == 0)
    ec_bwd_stats_kernel<true><<<grid, EC_THREADS, 0, st>>>(out_both, npos, beta, g_max, g_mean, g_both, P, F, k,
                                                           g_uv_clear, (double*)ws);
  else
    ec_bwd_stats_kernel<false><<<grid, EC_THREADS, 0, st>>>(out_both, npos, beta, g_max, g_mean, g_both, P, F, k,
                                                            g_uv_clear, (double*)ws);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("ec_bwd_stats_kernel");
  return launch_finalize_sums((const double*)ws, F, s1, s2, st);
}

extern "C" int dgcnn_edgeconv_bwd_apply(const void* uv, int uv_dtype, const int32_t* idx, int B, int N, int F, int k,
                                        const float* mean, const float* rstd, const float* beta, const float* zmax,
                                        const float* g_max, const float* g_mean, const float* g_both, const float* s1,
                                        const float* s2, float* g_uv, int v_half_cleared, dgcnn_stream_t stream) {
  int rc = ec_check(uv, idx, B, N, F, k, uv_dtype);
  if (rc) return rc;
  DG_REQUIRE(mean && rstd && beta && zmax && s1 && s2 && g_uv, DGCNN_ERR_INVALID, "edgeconv_bwd_apply: null pointer");
  DG_REQUIRE(((uintptr_t)g_uv & 15) == 0, DGCNN_ERR_INVALID, "edgeconv_bwd_apply: g_uv must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (!v_half_cleared) {
    const int64_t PF = (int64_t)B * N * F;
    zero_vhalf_kernel<<<cdiv(PF, 256), 256, 0, st>>>(g_uv, (int64_t)B * N, F);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("zero_vhalf_kernel");
  }
  if (uv_dtype == DGCNN_BF16)
    return bwd_apply_t<__nv_bfloat16>(uv, idx, B, N, F, k, mean, rstd, beta, zmax, g_max, g_mean, g_both, s1, s2, g_uv, st);
  if (uv_dtype == DGCNN_F16)
    return bwd_apply_t<__half>(uv, idx, B, N, F, k, mean, rstd, beta, zmax, g_max, g_mean, g_both, s1, s2, g_uv, st);
  return bwd_apply_t<float>(uv, idx, B, N, F, k, mean, rstd, beta, zmax, g_max, g_mean, g_both, s1, s2, g_uv, st);
}
