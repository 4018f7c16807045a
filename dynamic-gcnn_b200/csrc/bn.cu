// Train-mode BatchNorm (+residual)(+ReLU) on per-point [rows, C] tensors, forward and backward, and the
// TF-form Adam update.  Replaces slim.batch_norm / tf.nn.relu after the 1x1 convs of
// /root/reference/dgcnn/ops.py:53,68,131,134 and tf.train.AdamOptimizer (trainval.py:17,80).
// slim.batch_norm defaults: is_training=True (always -- SURVEY.md section 0), center=True, scale=False,
// epsilon=1e-3, biased batch variance over every non-channel axis.
#include <cuda_bf16.h>

#include "common.cuh"

namespace dgcnn {

constexpr int BN_THREADS = 256;

// up to two "plane sinks" of an apply pass: the output is also written as bf16 hi / lo planes into column slices of
// tensor-core operands (row pitch ld, second plane `plane` elements later), so that no split pass reads it again
struct BnSinks {
  int n;
  __nv_bfloat16* p[2];
  int ld[2];
  size_t plane[2];
};
// Max-pool over the rows of each group fused around a layer (model.py:76-77: the global max over the N points of a cloud
// of the MergedEdgeConv output).  Forward: the apply pass leaves max / #argmax per (group, channel).  Backward: the pool's
// gradient  [y == max] g_pool / cnt  is added to the layer's incoming gradient on the fly, y re-evaluated from z.
struct BnPool {
  const float* pmax;    // [groups][C] or null (no pool)
  const float* pcnt;    // [groups][C]
  const float* pgrad;   // [groups][C]
  int rows;             // rows per group
};
constexpr int BN_BLOCKS_PER_SM = 8;

static inline int bn_max_blocks() { return num_sms() * BN_BLOCKS_PER_SM; }
static inline int bn_blocks(int64_t rows) {
  int64_t need = (rows + 63) / 64;
  int nb = bn_max_blocks();
  return (int)(need < nb ? (need < 1 ? 1 : need) : nb);
}

// MODE 0: partial (sum z, sum z^2).   MODE 1: partial (sum g_pre, sum g_pre*zhat)
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
    bn_colsum_kernel(const float* __restrict__ z, const float* __restrict__ out, const float* __restrict__ gout,
                     const float* __restrict__ mean, const float* __restrict__ rstd, int relu, int64_t rows, int C,
                     double* __restrict__ acc, const float* __restrict__ gbias, int grows,
                     const float* __restrict__ beta, float* __restrict__ pivot) {
  __shared__ float red[2][4][64];
  const int cl = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cl;
  const bool ok = c < C;
  const int64_t rpb = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t rbeg = (int64_t)blockIdx.x * rpb;
  const int64_t rend = rbeg + rpb < rows ? rbeg + rpb : rows;
  float a = 0.f, q = 0.f;
  float mu = 0.f, rs = 0.f, be = 0.f;
  // MODE 0: sums of (z - pivot), pivot = the first row: E[z^2] - E[z]^2 then does not cancel for channels whose mean is
  // large against their spread (the totals are fp64, but every thread's partial is fp32)
  float pv = 0.f;
  if (MODE == 0 && ok) {
    pv = z[c] + (gbias ? gbias[c] : 0.f);
    if (blockIdx.x == 0 && rg == 0) pivot[c] = pv;
  }
  if (MODE == 1 && ok) {
    mu = mean[c];
    rs = rstd[c];
    if (beta) be = beta[c];
  }
  if (ok) {
    for (int64_t r = rbeg + rg; r < rend; r += 4) {
      const int64_t o = r * C + c;
      if (MODE == 0) {
        float v = z[o];
        if (gbias) v += gbias[(r / grows) * C + c];
        v -= pv;
        a += v;
        q = fmaf(v, v, q);
      } else {
        float gp = gout[o];
        float zv = z[o];
        if (gbias) zv += gbias[(r / grows) * C + c];
        // ReLU mask: from the saved output, or (out == null, no residual) by re-evaluating the forward expression
        if (relu && !((out ? out[o] : fmaf(zv - mu, rs, be)) > 0.f)) gp = 0.f;
        a += gp;
        q = fmaf(gp, (zv - mu) * rs, q);
      }
    }
  }
  red[0][rg][cl] = a;
  red[1][rg][cl] = q;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6;
    const float t = red[which][0][cl] + red[which][1][cl] + red[which][2][cl] + red[which][3][cl];
    if (ok) atomicAdd(&acc[which * C + c], (double)t);
  }
}

// The same reduction with 16-byte loads (C % 4 == 0, 16-byte aligned buffers): a thread owns 4 consecutive channels,
// a block covers CB = min(C, 256) channels x (256 / (CB/4)) row lanes, U rows in flight per thread (all of a batch's
// loads are issued before the first is consumed).  Row indices are 32-bit (rows < 2^31 is checked by the callers'
// 2^32-element limit): the per-row group / pool divisions are 32-bit ones.
template <int MODE, int U>
__global__ void __launch_bounds__(BN_THREADS, U == 2 ? 3 : 2)
    bn_colsum_vec_kernel(const float* __restrict__ z, const float* __restrict__ out, const float* __restrict__ gout,
                         const float* __restrict__ mean, const float* __restrict__ rstd, int relu, int rows, int C,
                         int cb, double* __restrict__ acc, const float* __restrict__ gbias, int grows,
                         const float* __restrict__ beta, const BnPool pool, float* __restrict__ pivot) {
  __shared__ float red[2][BN_THREADS][4];
  const int tpr = cb >> 2;                        // threads per row
  const int cv = threadIdx.x % tpr, rl = threadIdx.x / tpr, lanes = BN_THREADS / tpr;
  const int c = blockIdx.y * cb + cv * 4;
  const bool ok = c < C;
  const int rpb = (rows + gridDim.x - 1) / gridDim.x;
  const int rbeg = blockIdx.x * rpb;
  const int rend = rbeg + rpb < rows ? rbeg + rpb : rows;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  float mu[4] = {0.f, 0.f, 0.f, 0.f}, rs[4] = {0.f, 0.f, 0.f, 0.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
  float pv[4] = {0.f, 0.f, 0.f, 0.f};            // MODE 0: the first row as pivot (see bn_colsum_kernel)
  if (MODE == 0 && ok) {
    *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(z + c);
    if (gbias) {
      const float4 g0 = *reinterpret_cast<const float4*>(gbias + c);
      pv[0] += g0.x; pv[1] += g0.y; pv[2] += g0.z; pv[3] += g0.w;
    }
    if (blockIdx.x == 0 && rl == 0) *reinterpret_cast<float4*>(pivot + c) = *reinterpret_cast<float4*>(pv);
  }
  if (MODE == 1 && ok) {
    *reinterpret_cast<float4*>(mu) = *reinterpret_cast<const float4*>(mean + c);
    *reinterpret_cast<float4*>(rs) = *reinterpret_cast<const float4*>(rstd + c);
    if (beta) *reinterpret_cast<float4*>(be) = *reinterpret_cast<const float4*>(beta + c);
  }
  const bool use_out = MODE == 1 && relu && out;
  if (ok) {
    for (int rb = rbeg + rl; rb < rend; rb += U * lanes) {
      float zv[U][4], gp[U][4], ov[U][4], gb[U][4], pm[U][4];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {                  // every load of the batch first
        const int r = rb + u * lanes;
        live[u] = r < rend;
        const int rr = live[u] ? r : rend - 1;       // clamped duplicate, skipped below
        const size_t o = (size_t)rr * C + c;
        *reinterpret_cast<float4*>(zv[u]) = *reinterpret_cast<const float4*>(z + o);
        if (MODE == 1) {
          *reinterpret_cast<float4*>(gp[u]) = *reinterpret_cast<const float4*>(gout + o);
          if (use_out) *reinterpret_cast<float4*>(ov[u]) = *reinterpret_cast<const float4*>(out + o);
        }
        if (gbias) *reinterpret_cast<float4*>(gb[u]) = __ldg(reinterpret_cast<const float4*>(gbias + (size_t)(rr / grows) * C + c));
        if (MODE == 1 && pool.pmax)
          *reinterpret_cast<float4*>(pm[u]) = __ldg(reinterpret_cast<const float4*>(pool.pmax + (size_t)(rr / pool.rows) * C + c));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = gbias ? zv[u][i] + gb[u][i] : zv[u][i];
          if (MODE == 0) {
            const float d = v - pv[i];
            a[i] += d;
            q[i] = fmaf(d, d, q[i]);
          } else {
            float g = gp[u][i];
            const float t = use_out ? ov[u][i] : fmaf(v - mu[i], rs[i], be[i]);   // the forward output before the ReLU clamp
            // the pooled copy of this output: its gradient goes to the arg-max rows (ties share); a hit is rare, so the
            // count and the pooled gradient are fetched only then
            if (pool.pmax && fmaxf(t, 0.f) == pm[u][i]) {
              const size_t po = (size_t)((rb + u * lanes) / pool.rows) * C + c + i;
              g += __ldg(pool.pgrad + po) / __ldg(pool.pcnt + po);
            }
            if (relu && !(t > 0.f)) g = 0.f;
            a[i] += g;
            q[i] = fmaf(g, (v - mu[i]) * rs[i], q[i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[0][threadIdx.x][i] = a[i];
    red[1][threadIdx.x][i] = q[i];
  }
  __syncthreads();
  // thread t < 2*cb: which = t / cb, channel t % cb; sum over the row lanes in a fixed order
  for (int t = threadIdx.x; t < 2 * cb; t += BN_THREADS) {
    const int which = t / cb, ch = t - which * cb;
    const int ccv = ch >> 2, ci = ch & 3;
    float sum = 0.f;
    for (int l = 0; l < lanes; ++l) sum += red[which][l * tpr + ccv][ci];
    const int cc = blockIdx.y * cb + ch;
    if (cc < C) atomicAdd(&acc[which * C + cc], (double)sum);
  }
}

template <int MODE>
static void launch_colsum(const float* z, const float* out, const float* gout, const float* mean, const float* rstd,
                          int relu, int64_t rows, int C, double* acc, const float* gbias, int grows, const float* beta,
                          cudaStream_t st, const BnPool pool = BnPool{nullptr, nullptr, nullptr, 1}, float* pivot = nullptr) {
  const bool vec = (C & 3) == 0 && C >= 16 &&
                   (((uintptr_t)z | (uintptr_t)out | (uintptr_t)gout | (uintptr_t)gbias | (uintptr_t)mean |
                     (uintptr_t)rstd | (uintptr_t)beta) & 15) == 0;
  int nb = bn_blocks(rows);
  // every block ends with 2*cb fp64 atomics onto the same 2*C addresses: with narrow tensors (C <= 128) that tail, not
  // the streaming, sets the time -- fewer, longer blocks there
  if (C <= 128 && nb > 2 * num_sms()) nb = 2 * num_sms();
  if (vec && rows < (1ll << 31)) {
    int cb = 256;
    while (cb > C) cb >>= 1;                        // 16 .. 256, a power of two <= C
    const int slabs = cdiv(C, cb);
    if (MODE == 1) {
      // backward statistics (two streams in, z and g): ONE wave of 2 blocks per SM with four rows in flight per thread
      // beats many short blocks with two (profiles/r02_bn_bwd_sweep.txt: -17 us at C = 1024 / 512, -9 us at 256, -5 us
      // at 64 on P = 49152 rows): the per-block reduction + fp64 atomics tail is paid 296 times instead of 3072
      const int one_wave = max(1, 2 * num_sms() / slabs);
      dim3 grid(nb < one_wave ? nb : one_wave, slabs);
      bn_colsum_vec_kernel<MODE, 4><<<grid, BN_THREADS, 0, st>>>(z, out, gout, mean, rstd, relu, (int)rows, C, cb, acc,
                                                                 gbias, grows, beta, pool, pivot);
    } else {
      dim3 grid(nb, slabs);
      bn_colsum_vec_kernel<MODE, 2><<<grid, BN_THREADS, 0, st>>>(z, out, gout, mean, rstd, relu, (int)rows, C, cb, acc,
                                                                 gbias, grows, beta, pool, pivot);
    }
  } else {
    dim3 grid(nb, cdiv(C, 64));
    bn_colsum_kernel<MODE><<<grid, BN_THREADS, 0, st>>>(z, out, gout, mean, rstd, relu, rows, C, acc, gbias, grows, beta,
                                                        pivot);
  }
}

// Element-wise passes.  VEC = 4: one thread handles 4 consecutive channels of one row (float4 traffic, one 32-bit
// division per 4 elements); VEC = 1 is the fallback for C % 4 != 0.
template <int VEC>
__global__ void __launch_bounds__(256)
    bn_act_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ beta, const float* __restrict__ res, int relu, uint32_t nvec, int C,
                      float* __restrict__ out, const float* __restrict__ gbias, int grows, const BnSinks sinks) {
  const uint32_t cv = (uint32_t)C / VEC;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
    const uint32_t r = v / cv;
    const uint32_t c = (v - r * cv) * VEC;
    const size_t e = (size_t)r * C + c;
    float zz[VEC], rr[VEC], gb[VEC], y[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(zz) = *reinterpret_cast<const float4*>(z + e);
      if (res) *reinterpret_cast<float4*>(rr) = *reinterpret_cast<const float4*>(res + e);
      if (gbias) *reinterpret_cast<float4*>(gb) = *reinterpret_cast<const float4*>(gbias + (size_t)(r / grows) * C + c);
    } else {
      zz[0] = z[e];
      if (res) rr[0] = res[e];
      if (gbias) gb[0] = gbias[(size_t)(r / grows) * C + c];
    }
    float mu[VEC], rs[VEC], be[VEC];
    if (VEC == 4) {   // per-channel constants as 16-byte loads too (they are L1-resident)
      *reinterpret_cast<float4*>(mu) = __ldg(reinterpret_cast<const float4*>(mean + c));
      *reinterpret_cast<float4*>(rs) = __ldg(reinterpret_cast<const float4*>(rstd + c));
      *reinterpret_cast<float4*>(be) = __ldg(reinterpret_cast<const float4*>(beta + c));
    } else {
      mu[0] = mean[c];
      rs[0] = rstd[c];
      be[0] = beta[c];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float zv = zz[i];
      if (gbias) zv += gb[i];
      float t = fmaf(zv - mu[i], rs[i], be[i]);
      if (res) t += rr[i];
      y[i] = relu ? fmaxf(t, 0.f) : t;
    }
    if (VEC == 4) {
      if (out) *reinterpret_cast<float4*>(out + e) = *reinterpret_cast<float4*>(y);   // null: only the planes are wanted
      if (sinks.n > 0) {
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          h[i] = __float2bfloat16_rn(y[i]);
          l[i] = __float2bfloat16_rn(y[i] - __bfloat162float(h[i]));
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q < sinks.n) {
            const size_t se = (size_t)r * sinks.ld[q] + c;
            *reinterpret_cast<uint2*>(sinks.p[q] + se) = *reinterpret_cast<uint2*>(h);
            if (sinks.plane[q]) *reinterpret_cast<uint2*>(sinks.p[q] + sinks.plane[q] + se) = *reinterpret_cast<uint2*>(l);
          }
        }
      }
    } else {
      out[e] = y[0];
    }
  }
}

// Apply pass of a layer whose output is also max-pooled over the rows of each group (MergedEdgeConv, model.py:65-81):
// y = relu((z - mean) rstd + beta) goes to the plane sinks (and to `out` if given) while every thread keeps the running
// (max, #argmax) of its 4 channels over the rows of its slice of one group; slices are combined with a 64-bit CAS on
// packed (value bits : count) words -- y >= 0, so the unsigned order of the bits is the order of the values.
// grid = (row slices per group, groups); a thread owns channel quad c4 = threadIdx.x, + 256, ...
__global__ void __launch_bounds__(256)
    bn_apply_pool_kernel(const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ beta, int C, int pool_rows, int slice_rows, float* __restrict__ out,
                         const BnSinks sinks, unsigned long long* __restrict__ packed) {
  const int g = blockIdx.y;
  const int r_lo = blockIdx.x * slice_rows, r_hi = min(pool_rows, r_lo + slice_rows);
  for (int c4 = threadIdx.x; c4 < (C >> 2); c4 += blockDim.x) {
    const int c = c4 * 4;
    float mu[4], rs[4], be[4];
    *reinterpret_cast<float4*>(mu) = __ldg(reinterpret_cast<const float4*>(mean + c));
    *reinterpret_cast<float4*>(rs) = __ldg(reinterpret_cast<const float4*>(rstd + c));
    *reinterpret_cast<float4*>(be) = __ldg(reinterpret_cast<const float4*>(beta + c));
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned n[4] = {0u, 0u, 0u, 0u};
    constexpr int U = 4;                                   // rows in flight per thread
    for (int rb = r_lo; rb < r_hi; rb += U) {
      float zz[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rl = rb + u < r_hi ? rb + u : r_hi - 1;     // clamp: the duplicate row is skipped below
        *reinterpret_cast<float4*>(zz[u]) =
            *reinterpret_cast<const float4*>(z + ((size_t)g * pool_rows + rl) * C + c);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (rb + u >= r_hi) break;
        const size_t r = (size_t)g * pool_rows + rb + u;
        const size_t e = r * C + c;
        float y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          y[i] = fmaxf(fmaf(zz[u][i] - mu[i], rs[i], be[i]), 0.f);
          if (y[i] > m[i]) { m[i] = y[i]; n[i] = 1u; } else if (y[i] == m[i]) ++n[i];
        }
        if (out) *reinterpret_cast<float4*>(out + e) = *reinterpret_cast<float4*>(y);
        if (sinks.n > 0) {
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            h[i] = __float2bfloat16_rn(y[i]);
            l[i] = __float2bfloat16_rn(y[i] - __bfloat162float(h[i]));
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q < sinks.n) {
              const size_t se = r * sinks.ld[q] + c;
              *reinterpret_cast<uint2*>(sinks.p[q] + se) = *reinterpret_cast<uint2*>(h);
              if (sinks.plane[q]) *reinterpret_cast<uint2*>(sinks.p[q] + sinks.plane[q] + se) = *reinterpret_cast<uint2*>(l);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (n[i] == 0u) continue;
      unsigned long long* w = packed + (size_t)g * C + c + i;
      const unsigned long long mine_bits = (unsigned long long)__float_as_uint(m[i]) << 32;
      unsigned long long old = *w;
      while (true) {
        const unsigned long long ov = old & 0xffffffff00000000ull;
        unsigned long long nw;
        if (ov > mine_bits) break;                                      // a larger maximum is already there
        if (ov == mine_bits) nw = old + n[i];                            // same maximum: counts add
        else nw = mine_bits | n[i];                                      // ours is larger
        const unsigned long long seen = atomicCAS(w, old, nw);
        if (seen == old) break;
        old = seen;
      }
    }
  }
}

__global__ void bn_pool_unpack_kernel(const unsigned long long* __restrict__ packed, int n, float* __restrict__ pmax,
                                      float* __restrict__ pcnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long w = packed[i];
  pmax[i] = __uint_as_float((unsigned)(w >> 32));
  pcnt[i] = (float)(unsigned)(w & 0xffffffffull);
}

template <int VEC>
__global__ void __launch_bounds__(256)
    bn_act_bwd_kernel(const float* __restrict__ z, const float* __restrict__ out, const float* __restrict__ gout,
                      const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ s1,
                      const float* __restrict__ s2, int relu, uint32_t nvec, int C, float inv_rows,
                      float* __restrict__ gz, float* __restrict__ gpre, const float* __restrict__ gbias, int grows,
                      __nv_bfloat16* __restrict__ gz_planes, size_t plane_elems, const float* __restrict__ beta,
                      const BnPool pool) {
  const uint32_t cv = (uint32_t)C / VEC;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
    const uint32_t r = v / cv;
    const uint32_t c = (v - r * cv) * VEC;
    const size_t e = (size_t)r * C + c;
    float zz[VEC], oo[VEC], gg[VEC], gb[VEC], gzv[VEC], gpv[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(zz) = *reinterpret_cast<const float4*>(z + e);
      *reinterpret_cast<float4*>(gg) = *reinterpret_cast<const float4*>(gout + e);
      if (relu && out) *reinterpret_cast<float4*>(oo) = *reinterpret_cast<const float4*>(out + e);
      if (gbias) *reinterpret_cast<float4*>(gb) = *reinterpret_cast<const float4*>(gbias + (size_t)(r / grows) * C + c);
    } else {
      zz[0] = z[e];
      gg[0] = gout[e];
      if (relu && out) oo[0] = out[e];
      if (gbias) gb[0] = gbias[(size_t)(r / grows) * C + c];
    }
    float mu[VEC], rsv[VEC], be[VEC], a1[VEC], a2[VEC];
    if (VEC == 4) {   // per-channel constants as 16-byte loads too (they are L1-resident)
      *reinterpret_cast<float4*>(mu) = __ldg(reinterpret_cast<const float4*>(mean + c));
      *reinterpret_cast<float4*>(rsv) = __ldg(reinterpret_cast<const float4*>(rstd + c));
      *reinterpret_cast<float4*>(a1) = __ldg(reinterpret_cast<const float4*>(s1 + c));
      *reinterpret_cast<float4*>(a2) = __ldg(reinterpret_cast<const float4*>(s2 + c));
      if (beta) *reinterpret_cast<float4*>(be) = __ldg(reinterpret_cast<const float4*>(beta + c));
    } else {
      mu[0] = mean[c];
      rsv[0] = rstd[c];
      a1[0] = s1[c];
      a2[0] = s2[c];
      if (beta) be[0] = beta[c];
    }
    float pm[VEC];
    size_t po = 0;
    if (VEC == 4 && pool.pmax) {
      po = (size_t)(r / pool.rows) * C + c;
      *reinterpret_cast<float4*>(pm) = __ldg(reinterpret_cast<const float4*>(pool.pmax + po));
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float gp = gg[i];
      const float rs = rsv[i];
      float zv = zz[i];
      if (gbias) zv += gb[i];
      const float t = (relu || pool.pmax) ? (out ? oo[i] : fmaf(zv - mu[i], rs, be[i])) : 1.f;
      if (VEC == 4 && pool.pmax && fmaxf(t, 0.f) == pm[i]) gp += __ldg(pool.pgrad + po + i) / __ldg(pool.pcnt + po + i);
      if (relu && !(t > 0.f)) gp = 0.f;
      const float zh = (zv - mu[i]) * rs;
      gzv[i] = rs * (gp - a1[i] * inv_rows - zh * (a2[i] * inv_rows));
      gpv[i] = gp;
    }
    if (VEC == 4) {
      if (gz) *reinterpret_cast<float4*>(gz + e) = *reinterpret_cast<float4*>(gzv);
      if (gpre) *reinterpret_cast<float4*>(gpre + e) = *reinterpret_cast<float4*>(gpv);
      if (gz_planes) {   // the gradient as a tcgen05 operand: bf16 hi / lo planes (tc_gemm.cu), no fp32 round trip
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          h[i] = __float2bfloat16_rn(gzv[i]);
          l[i] = __float2bfloat16_rn(gzv[i] - __bfloat162float(h[i]));
        }
        *reinterpret_cast<uint2*>(gz_planes + e) = *reinterpret_cast<uint2*>(h);
        if (plane_elems) *reinterpret_cast<uint2*>(gz_planes + plane_elems + e) = *reinterpret_cast<uint2*>(l);
      }
    } else {
      if (gz) gz[e] = gzv[0];
      if (gpre) gpre[e] = gpv[0];
    }
  }
}

__global__ void adam_tf_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                               float* __restrict__ v, int64_t n, float lr_t, float b1, float b2, float eps,
                               float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// Global max over the points of each cloud (gen_nn_ops.max_pool_v2 with ksize [1,N,1,1], model.py:77) and its
// gradient (MaxPoolGrad routes to the arg-max; exact ties share the gradient like tf.reduce_max / torch.amax).
// x [G, rows, C] -> out [G, C], cnt [G, C] = number of points attaining the maximum.
__global__ void __launch_bounds__(256)
    group_max_fwd_kernel(const float* __restrict__ x, int rows, int C, float* __restrict__ out, float* __restrict__ cnt) {
  // block = 64 channels (16 float4 lanes) x 16 row groups; 4 independent 16-byte loads in flight per thread
  __shared__ float rm[16][64];
  __shared__ float rc[16][64];
  const int cv = threadIdx.x & 15, rg = threadIdx.x >> 4;
  const int c0 = blockIdx.x * 64 + cv * 4;
  const int g = blockIdx.y;
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, n[4] = {0.f, 0.f, 0.f, 0.f};
  const float* xg = x + (size_t)g * rows * C;
  auto upd = [&](const float (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (v[i] > m[i]) { m[i] = v[i]; n[i] = 1.f; } else if (v[i] == m[i]) n[i] += 1.f;
    }
  };
  if ((C & 3) == 0 && c0 + 3 < C && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    int r = rg;
    for (; r + 48 < rows; r += 64) {
      float v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        *reinterpret_cast<float4*>(v[u]) = *reinterpret_cast<const float4*>(xg + (size_t)(r + 16 * u) * C + c0);
#pragma unroll
      for (int u = 0; u < 4; ++u) upd(v[u]);
    }
    for (; r < rows; r += 16) {
      float v[4];
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(xg + (size_t)r * C + c0);
      upd(v);
    }
  } else {
    for (int r = rg; r < rows; r += 16) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = c0 + i < C ? xg[(size_t)r * C + c0 + i] : -INFINITY;
      upd(v);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    rm[rg][cv * 4 + i] = m[i];
    rc[rg][cv * 4 + i] = n[i];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c = blockIdx.x * 64 + threadIdx.x;
    float mm = rm[0][threadIdx.x], nn = rc[0][threadIdx.x];
    for (int i = 1; i < 16; ++i) {
      const float mi = rm[i][threadIdx.x], ni = rc[i][threadIdx.x];
      if (mi > mm) { mm = mi; nn = ni; } else if (mi == mm) nn += ni;
    }
    if (c < C) {
      out[(size_t)g * C + c] = mm;
      cnt[(size_t)g * C + c] = nn;
    }
  }
}

__global__ void __launch_bounds__(256)
    group_max_bwd_kernel(const float* __restrict__ x, const float* __restrict__ out, const float* __restrict__ cnt,
                         const float* __restrict__ gout, int rows, int C, uint32_t nvec, float* __restrict__ gx) {
  const uint32_t cv = (uint32_t)C / 4;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
    const uint32_t r = v / cv;
    const uint32_t c = (v - r * cv) * 4;
    const size_t go = (size_t)(r / rows) * C + c;
    const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)r * C + c);
    const float4 mv = *reinterpret_cast<const float4*>(out + go);
    const float4 nv = *reinterpret_cast<const float4*>(cnt + go);
    const float4 gv = *reinterpret_cast<const float4*>(gout + go);
    float4 o;
    o.x = xv.x == mv.x ? gv.x / nv.x : 0.f;
    o.y = xv.y == mv.y ? gv.y / nv.y : 0.f;
    o.z = xv.z == mv.z ? gv.z / nv.z : 0.f;
    o.w = xv.w == mv.w ? gv.w / nv.w : 0.f;
    *reinterpret_cast<float4*>(gx + (size_t)r * C + c) = o;
  }
}

// BatchNorm statistics from the per-tile column partials the wide GEMM epilogue leaves behind (tc_gemm_wide.cu):
// colstats [tiles][2][C] (sum, sum of squares over the 128 rows of each tile) -> mean, rstd.  With a per-group bias
// (rows of group g get gbias[g] added before the statistics, group_rows % 128 == 0):
//   sum (z+b) = sum z + group_rows * sum_g b_g ;  sum (z+b)^2 = sum z^2 + 2 sum_g b_g (sum_{tiles of g} sum z) + group_rows sum_g b_g^2
__global__ void __launch_bounds__(256)
    bn_tile_stats_kernel(const float* __restrict__ colstats, int tiles, int C, double rows, const float* __restrict__ gbias,
                         int tiles_per_group, double group_rows, float eps, float* __restrict__ mean,
                         float* __restrict__ rstd) {
  // block = 8 columns x 32 tile lanes; 8 consecutive columns of one tile are one 32-byte sector
  __shared__ double r1[32][8], r2[32][8];
  const int cl = threadIdx.x & 7, tl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  double a = 0.0, q = 0.0;
  if (c < C) {
    for (int t = tl; t < tiles; t += 32) {
      const double s1 = (double)colstats[((size_t)t * 2) * C + c];
      const double s2 = (double)colstats[((size_t)t * 2 + 1) * C + c];
      a += s1;
      q += s2;
      if (gbias) {
        const double b = (double)gbias[(size_t)(t / tiles_per_group) * C + c];
        q += 2.0 * b * s1;
        if (t % tiles_per_group == 0) {
          a += group_rows * b;
          q += group_rows * b * b;
        }
      }
    }
  }
  r1[tl][cl] = a;
  r2[tl][cl] = q;
  __syncthreads();
  if (tl == 0 && c < C) {
    for (int i = 1; i < 32; ++i) {
      a += r1[i][cl];
      q += r2[i][cl];
    }
    const double m = a / rows;
    double var = q / rows - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// gx += gradient of the group max, in place: only the arg-max positions are touched (everything else of the incoming
// gradient buffer is left as it is), so the dense [G,rows,C] pooling gradient and its accumulation pass never exist.
__global__ void __launch_bounds__(256)
    group_max_bwd_add_kernel(const float* __restrict__ x, const float* __restrict__ out, const float* __restrict__ cnt,
                             const float* __restrict__ gout, int rows, int C, uint32_t nvec, float* __restrict__ gx) {
  const uint32_t cv = (uint32_t)C / 4;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
    const uint32_t r = v / cv;
    const uint32_t c = (v - r * cv) * 4;
    const size_t go = (size_t)(r / rows) * C + c;
    const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)r * C + c);
    const float4 mv = *reinterpret_cast<const float4*>(out + go);
    if (xv.x == mv.x || xv.y == mv.y || xv.z == mv.z || xv.w == mv.w) {
      const float4 nv = *reinterpret_cast<const float4*>(cnt + go);
      const float4 gv = *reinterpret_cast<const float4*>(gout + go);
      float4 o = *reinterpret_cast<float4*>(gx + (size_t)r * C + c);
      if (xv.x == mv.x) o.x += gv.x / nv.x;
      if (xv.y == mv.y) o.y += gv.y / nv.y;
      if (xv.z == mv.z) o.z += gv.z / nv.z;
      if (xv.w == mv.w) o.w += gv.w / nv.w;
      *reinterpret_cast<float4*>(gx + (size_t)r * C + c) = o;
    }
  }
}

static inline int ew_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace dgcnn

using namespace dgcnn;

extern "C" size_t dgcnn_bn_workspace_bytes(int C) {
  if (C <= 0) return 0;
  return (size_t)2 * C * sizeof(double) + (size_t)C * sizeof(float);
}

extern "C" int dgcnn_bn_act_fwd(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                                int relu, float* out, float* mean, float* rstd, void* ws, size_t ws_bytes,
                                dgcnn_stream_t stream) {
  return dgcnn_bn_act_fwd_gb(z, rows, C, beta, residual, nullptr, 0, relu, out, mean, rstd, ws, ws_bytes, stream);
}

static int bn_make_sinks(BnSinks* sk, int n_sinks, void* const* sink_planes, const int* sink_lds,
                         const int64_t* sink_plane_elems, int C) {
  sk->n = 0;
  DG_REQUIRE(n_sinks >= 0 && n_sinks <= 2, DGCNN_ERR_INVALID, "bn sinks: at most two");
  for (int i = 0; i < n_sinks; ++i) {
    DG_REQUIRE(sink_planes && sink_lds && sink_plane_elems && sink_planes[i] && sink_lds[i] >= C &&
                   (sink_lds[i] & 3) == 0 && (sink_plane_elems[i] & 3) == 0 && ((uintptr_t)sink_planes[i] & 7) == 0,
               DGCNN_ERR_INVALID, "bn sinks: bad geometry of sink %d", i);
    sk->p[i] = (__nv_bfloat16*)sink_planes[i];
    sk->ld[i] = sink_lds[i];
    sk->plane[i] = (size_t)sink_plane_elems[i];
  }
  sk->n = n_sinks;
  return DGCNN_OK;
}

static int bn_act_fwd_impl(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                           const float* group_bias, int group_rows, int relu, float* out, float* mean, float* rstd,
                           void* ws, size_t ws_bytes, const BnSinks& sinks, dgcnn_stream_t stream);

extern "C" int dgcnn_bn_act_fwd_gb(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                                   const float* group_bias, int group_rows, int relu, float* out, float* mean,
                                   float* rstd, void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  BnSinks sk;
  sk.n = 0;
  return bn_act_fwd_impl(z, rows, C, beta, residual, group_bias, group_rows, relu, out, mean, rstd, ws, ws_bytes, sk,
                         stream);
}

extern "C" int dgcnn_bn_act_fwd_sinks(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                                      const float* group_bias, int group_rows, int relu, float* out, float* mean,
                                      float* rstd, void* ws, size_t ws_bytes, int n_sinks, void* const* sink_planes,
                                      const int* sink_lds, const int64_t* sink_plane_elems, dgcnn_stream_t stream) {
  BnSinks sk;
  int rc = bn_make_sinks(&sk, n_sinks, sink_planes, sink_lds, sink_plane_elems, C);
  if (rc) return rc;
  return bn_act_fwd_impl(z, rows, C, beta, residual, group_bias, group_rows, relu, out, mean, rstd, ws, ws_bytes, sk,
                         stream);
}

static int bn_act_fwd_impl(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                           const float* group_bias, int group_rows, int relu, float* out, float* mean, float* rstd,
                           void* ws, size_t ws_bytes, const BnSinks& sinks, dgcnn_stream_t stream) {
  DG_REQUIRE(!group_bias || (group_rows > 0 && rows % group_rows == 0), DGCNN_ERR_INVALID,
             "bn_act_fwd: rows=%lld is not a multiple of group_rows=%d", (long long)rows, group_rows);
  DG_REQUIRE(z && beta && out && mean && rstd && ws, DGCNN_ERR_INVALID, "bn_act_fwd: null pointer");
  DG_REQUIRE(rows > 0 && C > 0, DGCNN_ERR_INVALID, "bn_act_fwd: bad shape rows=%lld C=%d", (long long)rows, C);
  DG_REQUIRE(ws_bytes >= dgcnn_bn_workspace_bytes(C), DGCNN_ERR_WORKSPACE, "bn_act_fwd: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  DG_REQUIRE(((uintptr_t)ws & 7) == 0, DGCNN_ERR_INVALID, "bn_act_fwd: workspace must be 8-byte aligned");
  int rc = stats_acc_reset(ws, C, st);
  if (rc) return rc;
  float* pivot = reinterpret_cast<float*>(reinterpret_cast<double*>(ws) + (size_t)2 * C);   // the workspace's float[C] tail
  launch_colsum<0>(z, nullptr, nullptr, nullptr, nullptr, 0, rows, C, (double*)ws, group_bias, group_rows, nullptr, st,
                   BnPool{nullptr, nullptr, nullptr, 1}, pivot);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("bn_colsum_kernel<0>");
  rc = launch_finalize_stats((const double*)ws, C, (double)rows, 1e-3f, mean, rstd, st, pivot);
  if (rc) return rc;
  const int64_t total = rows * C;
  DG_REQUIRE(total < (1ll << 32), DGCNN_ERR_UNSUPPORTED, "bn_act_fwd: more than 2^32 elements");
  const bool vec = (C & 3) == 0 && (((uintptr_t)z | (uintptr_t)out | (uintptr_t)residual | (uintptr_t)group_bias |
                                     (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta) & 15) == 0;
  DG_REQUIRE(vec || sinks.n == 0, DGCNN_ERR_INVALID, "bn_act_fwd: plane sinks need C %% 4 == 0 and 16-byte aligned buffers");
  if (vec)
    bn_act_fwd_kernel<4><<<ew_blocks(total / 4), 256, 0, st>>>(z, mean, rstd, beta, residual, relu, (uint32_t)(total / 4), C,
                                                               out, group_bias, group_rows, sinks);
  else
    bn_act_fwd_kernel<1><<<ew_blocks(total), 256, 0, st>>>(z, mean, rstd, beta, residual, relu, (uint32_t)total, C, out,
                                                           group_bias, group_rows, sinks);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("bn_act_fwd_kernel");
  return DGCNN_OK;
}

extern "C" int dgcnn_bn_stats_from_tiles(const float* colstats, int tiles, int C, int64_t rows, const float* group_bias,
                                         int group_rows, float* mean, float* rstd, dgcnn_stream_t stream) {
  DG_REQUIRE(colstats && mean && rstd, DGCNN_ERR_INVALID, "bn_stats_from_tiles: null pointer");
  DG_REQUIRE(tiles > 0 && C > 0 && rows > 0 && (int64_t)tiles * 128 >= rows, DGCNN_ERR_INVALID,
             "bn_stats_from_tiles: bad shape tiles=%d C=%d rows=%lld", tiles, C, (long long)rows);
  DG_REQUIRE(!group_bias || (group_rows > 0 && group_rows % 128 == 0 && rows % group_rows == 0), DGCNN_ERR_INVALID,
             "bn_stats_from_tiles: group_rows=%d must be a multiple of 128 dividing rows", group_rows);
  bn_tile_stats_kernel<<<cdiv(C, 8), 256, 0, (cudaStream_t)stream>>>(colstats, tiles, C, (double)rows, group_bias,
                                                                      group_bias ? group_rows / 128 : 1,
                                                                      (double)group_rows, 1e-3f, mean, rstd);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("bn_tile_stats_kernel");
  return DGCNN_OK;
}

extern "C" int dgcnn_bn_apply_fwd(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                                  const float* group_bias, int group_rows, int relu, const float* mean,
                                  const float* rstd, float* out, dgcnn_stream_t stream) {
  return dgcnn_bn_apply_fwd_sinks(z, rows, C, beta, residual, group_bias, group_rows, relu, mean, rstd, out, 0, nullptr,
                                  nullptr, nullptr, stream);
}

extern "C" int dgcnn_bn_apply_fwd_sinks(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                                        const float* group_bias, int group_rows, int relu, const float* mean,
                                        const float* rstd, float* out, int n_sinks, void* const* sink_planes,
                                        const int* sink_lds, const int64_t* sink_plane_elems, dgcnn_stream_t stream) {
  BnSinks sinks;
  {
    int rcs = bn_make_sinks(&sinks, n_sinks, sink_planes, sink_lds, sink_plane_elems, C);
    if (rcs) return rcs;
  }
  DG_REQUIRE(z && beta && mean && rstd && (out || sinks.n > 0), DGCNN_ERR_INVALID, "bn_apply_fwd: null pointer");
  DG_REQUIRE(rows > 0 && C > 0, DGCNN_ERR_INVALID, "bn_apply_fwd: bad shape rows=%lld C=%d", (long long)rows, C);
  DG_REQUIRE(!group_bias || (group_rows > 0 && rows % group_rows == 0), DGCNN_ERR_INVALID,
             "bn_apply_fwd: rows=%lld is not a multiple of group_rows=%d", (long long)rows, group_rows);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = rows * C;
  DG_REQUIRE(total < (1ll << 32), DGCNN_ERR_UNSUPPORTED, "bn_apply_fwd: more than 2^32 elements");
  const bool vec = (C & 3) == 0 && (((uintptr_t)z | (uintptr_t)out | (uintptr_t)residual | (uintptr_t)group_bias |
                                     (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta) & 15) == 0;
  DG_REQUIRE(vec || (sinks.n == 0 && out), DGCNN_ERR_INVALID,
             "bn_apply_fwd: plane sinks need C %% 4 == 0 and 16-byte aligned buffers");
  if (vec)
    bn_act_fwd_kernel<4><<<ew_blocks(total / 4), 256, 0, st>>>(z, mean, rstd, beta, residual, relu, (uint32_t)(total / 4), C,
                                                               out, group_bias, group_rows, sinks);
  else
    bn_act_fwd_kernel<1><<<ew_blocks(total), 256, 0, st>>>(z, mean, rstd, beta, residual, relu, (uint32_t)total, C, out,
                                                           group_bias, group_rows, sinks);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("bn_act_fwd_kernel");
  return DGCNN_OK;
}

extern "C" size_t dgcnn_bn_pool_workspace_bytes(int groups, int C) {
  return groups > 0 && C > 0 ? (size_t)groups * C * sizeof(unsigned long long) : 0;
}

extern "C" int dgcnn_bn_apply_fwd_pool(const float* z, int64_t rows, int C, const float* beta, const float* mean,
                                       const float* rstd, float* out, int pool_rows, float* pool_max, float* pool_cnt,
                                       void* ws, size_t ws_bytes, int n_sinks, void* const* sink_planes,
                                       const int* sink_lds, const int64_t* sink_plane_elems, dgcnn_stream_t stream) {
  BnSinks sinks;
  {
    int rcs = bn_make_sinks(&sinks, n_sinks, sink_planes, sink_lds, sink_plane_elems, C);
    if (rcs) return rcs;
  }
  DG_REQUIRE(z && beta && mean && rstd && pool_max && pool_cnt && ws, DGCNN_ERR_INVALID, "bn_apply_fwd_pool: null pointer");
  DG_REQUIRE(out || sinks.n > 0, DGCNN_ERR_INVALID, "bn_apply_fwd_pool: the output needs at least one destination");
  DG_REQUIRE(rows > 0 && C >= 4 && (C & 3) == 0 && pool_rows > 0 && rows % pool_rows == 0, DGCNN_ERR_INVALID,
             "bn_apply_fwd_pool: bad shape rows=%lld C=%d pool_rows=%d", (long long)rows, C, pool_rows);
  DG_REQUIRE((((uintptr_t)z | (uintptr_t)out | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta | (uintptr_t)pool_max |
               (uintptr_t)pool_cnt) & 15) == 0 && ((uintptr_t)ws & 7) == 0, DGCNN_ERR_INVALID, "bn_apply_fwd_pool: alignment");
  const int groups = (int)(rows / pool_rows);
  DG_REQUIRE(groups <= 65535, DGCNN_ERR_UNSUPPORTED, "bn_apply_fwd_pool: more than 65535 groups");
  const size_t need = dgcnn_bn_pool_workspace_bytes(groups, C);
  DG_REQUIRE(ws_bytes >= need, DGCNN_ERR_WORKSPACE, "bn_apply_fwd_pool: workspace %zu < %zu bytes", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(ws, 0, need, st) != cudaSuccess) return set_err(DGCNN_ERR_CUDA, "bn_apply_fwd_pool: memset failed");
  // enough slices per group to fill the machine several times over, at least 16 rows each
  int slices = cdiv((int64_t)num_sms() * 8, groups);
  if (slices > cdiv(pool_rows, 16)) slices = cdiv(pool_rows, 16);
  if (slices < 1) slices = 1;
  const int slice_rows = cdiv(pool_rows, slices);
  dim3 grid(cdiv(pool_rows, slice_rows), groups);
  bn_apply_pool_kernel<<<grid, 256, 0, st>>>(z, mean, rstd, beta, C, pool_rows, slice_rows, out, sinks,
                                             (unsigned long long*)ws);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("bn_apply_pool_kernel");
  bn_pool_unpack_kernel<<<cdiv((int64_t)groups * C, 256), 256, 0, st>>>((const unsigned long long*)ws, groups * C,
                                                                          pool_max, pool_cnt);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("bn_pool_unpack_kernel");
  return DGCNN_OK;
}

static int bn_act_bwd_impl(const float* z, const float* out, const float* g_out, int64_t rows, int C, const float* mean,
                           const float* rstd, const float* group_bias, int group_rows, int relu, float* g_z,
                           void* g_z_planes, float* g_beta, float* g_pre, const float* beta, void* ws, size_t ws_bytes,
                           dgcnn_stream_t stream, int n_planes = 2, const BnPool pool = BnPool{nullptr, nullptr, nullptr, 1});

extern "C" int dgcnn_bn_act_bwd_planes(const float* z, const float* out, const float* beta, const float* g_out,
                                       int64_t rows, int C, const float* mean, const float* rstd,
                                       const float* group_bias, int group_rows, int relu, float* g_z, void* g_z_planes,
                                       int n_planes, float* g_beta, const float* pool_max, const float* pool_cnt,
                                       const float* pool_grad, int pool_rows, void* ws, size_t ws_bytes,
                                       dgcnn_stream_t stream) {
  DG_REQUIRE(!relu || out || beta, DGCNN_ERR_INVALID, "bn_act_bwd_planes: relu backward needs out or beta");
  DG_REQUIRE(!pool_max || (pool_cnt && pool_grad), DGCNN_ERR_INVALID, "bn_act_bwd_planes: pool_max needs pool_cnt and pool_grad");
  DG_REQUIRE(n_planes == 1 || n_planes == 2, DGCNN_ERR_INVALID, "bn_act_bwd_planes: n_planes must be 1 or 2");
  DG_REQUIRE(g_z_planes && (C & 3) == 0 && ((uintptr_t)g_z_planes & 7) == 0, DGCNN_ERR_INVALID,
             "bn_act_bwd_planes: needs a plane buffer and C %% 4 == 0");
  return bn_act_bwd_impl(z, out, g_out, rows, C, mean, rstd, group_bias, group_rows, relu, g_z, g_z_planes, g_beta,
                         nullptr, out ? nullptr : beta, ws, ws_bytes, stream, n_planes,
                         BnPool{pool_max, pool_cnt, pool_grad, pool_rows > 0 ? pool_rows : 1});
}

extern "C" int dgcnn_bn_act_bwd(const float* z, const float* out, const float* g_out, int64_t rows, int C,
                                const float* mean, const float* rstd, int relu, float* g_z, float* g_beta,
                                float* g_pre, void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  return dgcnn_bn_act_bwd_gb(z, out, g_out, rows, C, mean, rstd, nullptr, 0, relu, g_z, g_beta, g_pre, ws, ws_bytes,
                             stream);
}

extern "C" int dgcnn_bn_act_bwd_gb(const float* z, const float* out, const float* g_out, int64_t rows, int C,
                                   const float* mean, const float* rstd, const float* group_bias, int group_rows,
                                   int relu, float* g_z, float* g_beta, float* g_pre, void* ws, size_t ws_bytes,
                                   dgcnn_stream_t stream) {
  DG_REQUIRE(g_z, DGCNN_ERR_INVALID, "bn_act_bwd: null pointer");
  return bn_act_bwd_impl(z, out, g_out, rows, C, mean, rstd, group_bias, group_rows, relu, g_z, nullptr, g_beta, g_pre,
                         nullptr, ws, ws_bytes, stream);
}

static int bn_act_bwd_impl(const float* z, const float* out, const float* g_out, int64_t rows, int C, const float* mean,
                           const float* rstd, const float* group_bias, int group_rows, int relu, float* g_z,
                           void* g_z_planes, float* g_beta, float* g_pre, const float* beta, void* ws, size_t ws_bytes,
                           dgcnn_stream_t stream, int n_planes, const BnPool pool) {
  DG_REQUIRE(!group_bias || (group_rows > 0 && rows % group_rows == 0), DGCNN_ERR_INVALID,
             "bn_act_bwd: rows=%lld is not a multiple of group_rows=%d", (long long)rows, group_rows);
  DG_REQUIRE(z && g_out && mean && rstd && (g_z || g_z_planes) && g_beta && ws, DGCNN_ERR_INVALID,
             "bn_act_bwd: null pointer");
  DG_REQUIRE(!relu || out || beta, DGCNN_ERR_INVALID, "bn_act_bwd: relu backward needs the forward output");
  DG_REQUIRE(rows > 0 && C > 0, DGCNN_ERR_INVALID, "bn_act_bwd: bad shape rows=%lld C=%d", (long long)rows, C);
  DG_REQUIRE(ws_bytes >= dgcnn_bn_workspace_bytes(C), DGCNN_ERR_WORKSPACE, "bn_act_bwd: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  DG_REQUIRE(((uintptr_t)ws & 7) == 0, DGCNN_ERR_INVALID, "bn_act_bwd: workspace must be 8-byte aligned");
  double* acc = (double*)ws;
  float* s2 = reinterpret_cast<float*>(acc + (size_t)2 * C);
  int rc = stats_acc_reset(ws, C, st);
  if (rc) return rc;
  launch_colsum<1>(z, out, g_out, mean, rstd, relu, rows, C, acc, group_bias, group_rows, beta, st, pool);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("bn_colsum_kernel<1>");
  rc = launch_finalize_sums(acc, C, g_beta, s2, st);
  if (rc) return rc;
  const int64_t total = rows * C;
  DG_REQUIRE(total < (1ll << 32), DGCNN_ERR_UNSUPPORTED, "bn_act_bwd: more than 2^32 elements");
  const bool vec = (C & 3) == 0 && (((uintptr_t)z | (uintptr_t)out | (uintptr_t)g_out | (uintptr_t)g_z | (uintptr_t)g_pre |
                                     (uintptr_t)group_bias | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta |
                                     (uintptr_t)g_beta | (uintptr_t)s2) & 15) == 0;
  DG_REQUIRE(vec || !g_z_planes, DGCNN_ERR_INVALID, "bn_act_bwd: plane output needs 16-byte aligned buffers, C %% 4 == 0");
  DG_REQUIRE(!pool.pmax || (vec && C >= 16 && beta && !out && relu && pool.rows > 0 && rows % pool.rows == 0 &&
                            (((uintptr_t)pool.pmax | (uintptr_t)pool.pcnt | (uintptr_t)pool.pgrad) & 15) == 0),
             DGCNN_ERR_INVALID, "bn_act_bwd: a fused pool needs C %% 4 == 0, relu, beta (mask from z) and rows %% pool_rows == 0");
  if (vec)
    bn_act_bwd_kernel<4><<<ew_blocks(total / 4), 256, 0, st>>>(z, out, g_out, mean, rstd, g_beta, s2, relu,
                                                               (uint32_t)(total / 4), C, 1.0f / (float)rows, g_z, g_pre,
                                                               group_bias, group_rows, (__nv_bfloat16*)g_z_planes,
                                                               n_planes == 2 ? (size_t)total : 0, beta, pool);
  else
    bn_act_bwd_kernel<1><<<ew_blocks(total), 256, 0, st>>>(z, out, g_out, mean, rstd, g_beta, s2, relu, (uint32_t)total, C,
                                                           1.0f / (float)rows, g_z, g_pre, group_bias, group_rows, nullptr,
                                                           0, beta, pool);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("bn_act_bwd_kernel");
  return DGCNN_OK;
}

extern "C" int dgcnn_adam_tf_step(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1,
                                  float b2, float eps, float grad_scale, dgcnn_stream_t stream) {
  DG_REQUIRE(p && g && m && v, DGCNN_ERR_INVALID, "adam_tf_step: null pointer");
  DG_REQUIRE(n > 0, DGCNN_ERR_INVALID, "adam_tf_step: n=%lld", (long long)n);
  adam_tf_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr_t, b1, b2, eps, grad_scale);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("adam_tf_kernel");
  return DGCNN_OK;
}

extern "C" int dgcnn_group_max_fwd(const float* x, int groups, int rows, int C, float* out, float* cnt,
                                   dgcnn_stream_t stream) {
  DG_REQUIRE(x && out && cnt, DGCNN_ERR_INVALID, "group_max_fwd: null pointer");
  DG_REQUIRE(groups > 0 && rows > 0 && C > 0 && groups <= 65535, DGCNN_ERR_INVALID, "group_max_fwd: bad shape");
  dim3 grid(cdiv(C, 64), groups);
  group_max_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, C, out, cnt);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("group_max_fwd_kernel");
  return DGCNN_OK;
}

extern "C" int dgcnn_group_max_bwd(const float* x, const float* out, const float* cnt, const float* g_out, int groups,
                                   int rows, int C, float* g_x, dgcnn_stream_t stream) {
  DG_REQUIRE(x && out && cnt && g_out && g_x, DGCNN_ERR_INVALID, "group_max_bwd: null pointer");
  DG_REQUIRE(groups > 0 && rows > 0 && C > 0 && (C & 3) == 0, DGCNN_ERR_INVALID, "group_max_bwd: bad shape (C %% 4)");
  const int64_t total = (int64_t)groups * rows * C;
  DG_REQUIRE(total < (1ll << 32), DGCNN_ERR_UNSUPPORTED, "group_max_bwd: more than 2^32 elements");
  group_max_bwd_kernel<<<ew_blocks(total / 4), 256, 0, (cudaStream_t)stream>>>(x, out, cnt, g_out, rows, C,
                                                                               (uint32_t)(total / 4), g_x);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("group_max_bwd_kernel");
  return DGCNN_OK;
}

extern "C" int dgcnn_group_max_bwd_add(const float* x, const float* out, const float* cnt, const float* g_out, int groups,
                                       int rows, int C, float* g_x_inout, dgcnn_stream_t stream) {
  DG_REQUIRE(x && out && cnt && g_out && g_x_inout, DGCNN_ERR_INVALID, "group_max_bwd_add: null pointer");
  DG_REQUIRE(groups > 0 && rows > 0 && C > 0 && (C & 3) == 0, DGCNN_ERR_INVALID, "group_max_bwd_add: bad shape (C %% 4)");
  const int64_t total = (int64_t)groups * rows * C;
  DG_REQUIRE(total < (1ll << 32), DGCNN_ERR_UNSUPPORTED, "group_max_bwd_add: more than 2^32 elements");
  group_max_bwd_add_kernel<<<ew_blocks(total / 4), 256, 0, (cudaStream_t)stream>>>(x, out, cnt, g_out, rows, C,
                                                                                   (uint32_t)(total / 4), g_x_inout);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("group_max_bwd_add_kernel");
  return DGCNN_OK;
}

// ---- loss head: softmax + sparse cross-entropy (x weight) + accuracy + d loss / d logits in one pass ----------------
// /root/reference/dgcnn/trainval.py:39-52: softmax, argmax == label accuracy, mean over all points of the (weighted)
// per-point cross-entropy.  One thread per point; the gradient of the MEAN loss is written directly, so the
// backward pass has nothing left to compute.  acc = {loss sum, correct count} (fp64), out = {loss, accuracy}.
namespace dgcnn {
__global__ void __launch_bounds__(256)
    softmax_xent_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                        const float* __restrict__ weights, int64_t P, int K, float* __restrict__ grad,
                        double* __restrict__ acc, unsigned int* __restrict__ counter, float* __restrict__ out) {
  __shared__ double red[2][8];
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double loss = 0.0, correct = 0.0;
  if (p < P) {
    const float* l = logits + p * K;
    float m = l[0];
    int am = 0;
    for (int c = 1; c < K; ++c)
      if (l[c] > m) { m = l[c]; am = c; }          // first maximum, like argmax
    float se = 0.f;
    for (int c = 0; c < K; ++c) se += expf(l[c] - m);
    const int y = (int)labels[p];
    const float w = weights ? weights[p] : 1.0f;
    const float lse = m + logf(se);
    loss = (double)((lse - l[y]) * w);
    correct = am == y ? 1.0 : 0.0;
    const float gs = w / (float)P;
    const float inv = 1.0f / se;
    for (int c = 0; c < K; ++c) grad[p * K + c] = (expf(l[c] - m) * inv - (c == y ? 1.0f : 0.0f)) * gs;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(FULL, loss, o);
    correct += __shfl_xor_sync(FULL, correct, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = loss;
    red[1][threadIdx.x >> 5] = correct;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      loss += red[0][w];
      correct += red[1][w];
    }
    atomicAdd(&acc[0], loss);
    atomicAdd(&acc[1], correct);
    __threadfence();
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {   // last block: publish the means
      __threadfence();
      out[0] = (float)(atomicAdd(&acc[0], 0.0) / (double)P);
      out[1] = (float)(atomicAdd(&acc[1], 0.0) / (double)P);
    }
  }
}
}  // namespace dgcnn

extern "C" size_t dgcnn_softmax_xent_workspace_bytes(void) { return 32; }

extern "C" int dgcnn_softmax_xent(const float* logits, const int64_t* labels, const float* weights, int64_t P, int K,
                                  float* grad, float* loss_acc, void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  DG_REQUIRE(logits && labels && grad && loss_acc && ws, DGCNN_ERR_INVALID, "softmax_xent: null pointer");
  DG_REQUIRE(P > 0 && K > 0 && K <= 4096, DGCNN_ERR_INVALID, "softmax_xent: bad shape P=%lld K=%d", (long long)P, K);
  DG_REQUIRE(ws_bytes >= 32 && ((uintptr_t)ws & 7) == 0, DGCNN_ERR_WORKSPACE, "softmax_xent: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(ws, 0, 32, st) != cudaSuccess) return set_err(DGCNN_ERR_CUDA, "softmax_xent: memset failed");
  const int64_t blocks = (P + 255) / 256;
  DG_REQUIRE(blocks < (1ll << 31), DGCNN_ERR_UNSUPPORTED, "softmax_xent: too many points");
  softmax_xent_kernel<<<(unsigned)blocks, 256, 0, st>>>(logits, labels, weights, P, K, grad, (double*)ws,
                                                        reinterpret_cast<unsigned int*>((char*)ws + 16), loss_acc);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("softmax_xent_kernel");
  return DGCNN_OK;
}
