// Warp-level selection primitives shared by the SIMT and the tcgen05 kNN kernels.
#pragma once
#include "common.cuh"

namespace dgcnn {

// ---------------------------------------------------------------------------------------------
// Per-row selection state, owned by one warp.
//   * an ascending list of the best 32*KS (distance, index) pairs: slot p lives in lane p&31, register p>>5.
//     Order is lexicographic (d, j): equal distances keep the lower index first = tf.nn.top_k's tie rule
//     (ops.py:18).  Empty slots hold (+inf, INT_MAX).
//   * a 64-entry shared-memory queue of candidates that passed the threshold filter.  Whenever 32 are
//     queued the warp sorts them with a 32-lane bitonic network and bitonic-merges them into the list
//     (one ~200-instruction dependent chain per 32 candidates instead of one per candidate).
// The threshold (td, tj) = list entry k-1 only tightens at drains; stale entries are merely merged and drop off.
constexpr int QCAP = 64;

__device__ __forceinline__ bool lex_less(float ad, int aj, float bd, int bj) {
  return (ad < bd) || (ad == bd && aj < bj);
}

// ascending bitonic sort of one (d, j) pair per lane
__device__ __forceinline__ void warp_sort32(float& d, int& j, int lane) {
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (int jm = k2 >> 1; jm > 0; jm >>= 1) {
      const float od = __shfl_xor_sync(FULL, d, jm);
      const int oj = __shfl_xor_sync(FULL, j, jm);
      const bool up = (lane & k2) == 0;
      const bool lower = (lane & jm) == 0;
      const bool other_less = lex_less(od, oj, d, j);
      const bool take = (lower == up) ? other_less : !other_less;
      if (take) {
        d = od;
        j = oj;
      }
    }
  }
}

// d,j hold a bitonic sequence across the 32 lanes -> ascending
__device__ __forceinline__ void warp_bitonic_merge32(float& d, int& j, int lane) {
#pragma unroll
  for (int jm = 16; jm > 0; jm >>= 1) {
    const float od = __shfl_xor_sync(FULL, d, jm);
    const int oj = __shfl_xor_sync(FULL, j, jm);
    const bool lower = (lane & jm) == 0;
    const bool other_less = lex_less(od, oj, d, j);
    if (lower ? other_less : !other_less) {
      d = od;
      j = oj;
    }
  }
}

template <int KS>
struct RowSel {
  float d[KS];
  int j[KS];
  float td;  // admission threshold = entry k-1 of the list
  int tj;
  int cnt;   // queued candidates (warp-uniform)

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      d[s] = __int_as_float(0x7f800000);
      j[s] = 0x7fffffff;
    }
    td = __int_as_float(0x7f800000);
    tj = 0x7fffffff;
    cnt = 0;
  }

  // merge one candidate per lane (bd, bj; +inf pads) into the list and refresh the threshold
  __device__ __forceinline__ void merge_batch(float bd, int bj, int k, int lane) {
    warp_sort32(bd, bj, lane);
    const float rd = __shfl_sync(FULL, bd, 31 - lane);  // reversed batch
    const int rj = __shfl_sync(FULL, bj, 31 - lane);
    if (KS == 1) {
      if (lex_less(rd, rj, d[0], j[0])) {
        d[0] = rd;
        j[0] = rj;
      }
      warp_bitonic_merge32(d[0], j[0], lane);
    } else {
      // 64 smallest of list(64) U batch(32): C = [L0, min(L1, rev(batch))] is bitonic; merge network over 64
      if (lex_less(rd, rj, d[KS - 1], j[KS - 1])) {
        d[KS - 1] = rd;
        j[KS - 1] = rj;
      }
      if (lex_less(d[KS - 1], j[KS - 1], d[0], j[0])) {
        const float t = d[0];
        d[0] = d[KS - 1];
        d[KS - 1] = t;
        const int u = j[0];
        j[0] = j[KS - 1];
        j[KS - 1] = u;
      }
      warp_bitonic_merge32(d[0], j[0], lane);
      warp_bitonic_merge32(d[KS - 1], j[KS - 1], lane);
    }
    const int src = (k - 1) & 31;
    float a = __shfl_sync(FULL, d[0], src);
    int bjj = __shfl_sync(FULL, j[0], src);
    if (KS == 2) {
      const float a1 = __shfl_sync(FULL, d[KS - 1], src);
      const int b1 = __shfl_sync(FULL, j[KS - 1], src);
      if (k > 32) {
        a = a1;
        bjj = b1;
      }
    }
    td = a;
    tj = bjj;
  }

  // offer 4 candidates per lane; qd/qj = this row's queue (QCAP entries)
  __device__ __forceinline__ void offer4(const float (&dv)[4], const int (&cj)[4], int N, int k,
                                         float* __restrict__ qd, int* __restrict__ qj, int lane) {
    bool p[4];
    bool anyp = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      p[q] = (cj[q] < N) && lex_less(dv[q], cj[q], td, tj);
      anyp |= p[q];
    }
    if (__ballot_sync(FULL, anyp) == 0) return;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned m = __ballot_sync(FULL, p[q]);
      if (m == 0) continue;  // warp-uniform
      if (p[q]) {
        const int pos = cnt + __popc(m & lt);
        qd[pos] = dv[q];
        qj[pos] = cj[q];
      }
      cnt += __popc(m);
      if (cnt >= 32) {
        __syncwarp();
        cnt -= 32;
        const float bd = qd[cnt + lane];
        const int bj = qj[cnt + lane];
        __syncwarp();
        merge_batch(bd, bj, k, lane);
      }
    }
  }

  // drain what is left in the queue
  __device__ __forceinline__ void finish(int k, const float* __restrict__ qd, const int* __restrict__ qj, int lane) {
    if (cnt > 0) {
      __syncwarp();
      const float bd = lane < cnt ? qd[lane] : __int_as_float(0x7f800000);
      const int bj = lane < cnt ? qj[lane] : 0x7fffffff;
      cnt = 0;
      __syncwarp();
      merge_batch(bd, bj, k, lane);
    }
  }
};

}  // namespace dgcnn
