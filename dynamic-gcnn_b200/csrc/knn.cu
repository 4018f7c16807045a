// k_nn hot path: pairwise squared distance + per-row k-smallest, bit-exact with oracle/knn_oracle.c.
// Replaces the TF kernels behind /root/reference/dgcnn/ops.py:8-19 (BatchMatMul, Square, Sum, Add, Sub,
// Neg, TopKV2).  fp32 SIMT formulation: every p_ij is ONE sequential-in-c fmaf chain (the oracle's order),
// the [B,N,N] matrix lives only in registers, selection is a warp-resident sorted list.
#include "common.cuh"
#include "knn_select.cuh"

namespace dgcnn {

constexpr int TM = 64;    // query rows per CTA (8 per warp)
constexpr int TN = 128;   // candidate columns per tile (4 per lane)
constexpr int CK = 16;    // channels per pipeline stage
constexpr int KNN_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// prep: x [B,N,C] -> xT [B,C,Npad] (channel-major, zero padded) and s [B,Npad] (squared norms).
// s follows ops.py:14: square (rounded) then sum, sequential in c, no FMA.
__global__ void knn_prep_kernel(const float* __restrict__ x, float* __restrict__ xT, float* __restrict__ s,
                                int N, int Npad, int C) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Npad) return;
  float* xTb = xT + (size_t)b * C * Npad;
  float acc = 0.0f;
  if (n < N) {
    const float* r = x + ((size_t)b * N + n) * C;
    for (int c = 0; c < C; ++c) {
      float v = r[c];
      xTb[(size_t)c * Npad + n] = v;
      acc = __fadd_rn(acc, __fmul_rn(v, v));
    }
  } else {
    for (int c = 0; c < C; ++c) xTb[(size_t)c * Npad + n] = 0.0f;
  }
  s[(size_t)b * Npad + n] = acc;
}

// ---------------------------------------------------------------------------------------------
// Shared-memory selection state of the tile kernel (one row = one warp-owned record):
//   queue  [TM][TQ]      candidates that passed the filter, (d, j) split in two planes
//   list   [TM][32*KS]   ascending best-so-far list
//   tau    [TM]          admission threshold (list entry k-1), count [TM] queued candidates
//   dst    [TM][DLD]     the tile's distances, staged so that the scan can re-map threads: the FMA micro-tile is
//                        8 rows x 4 columns per lane, the scan gives every row to 4 lanes (columns q, q+4, ...);
//                        row pitch 132 floats makes both the float4 writes and the strided reads conflict-free
// Keeping this state out of registers lets the drain (sort + merge, ~200 instructions) exist ONCE in the
// instruction stream: a first version that kept the lists in registers had to unroll it per row and ran
// instruction-cache bound (ncu: stall_no_instruction 7 of 14.9 cycles per issue).
constexpr int TQ = 64;   // 31 left over + at most 32 appended per 32-column chunk
constexpr int DLD = TN + 4;

template <int KS>
struct KnnSel {
  float* qd;
  int* qj;
  float* ld;
  int* lj;
  float* taud;
  int* tauj;
  int* qcnt;
  float* dst;
  static constexpr int LW = 32 * KS;
  static constexpr size_t bytes() { return (size_t)TM * (DLD * 4 + TQ * 8 + LW * 8 + 12); }
  __device__ __forceinline__ void carve(unsigned char* base) {
    dst = reinterpret_cast<float*>(base);
    qd = reinterpret_cast<float*>(base + (size_t)TM * DLD * 4);
    qj = reinterpret_cast<int*>(qd + TM * TQ);
    ld = reinterpret_cast<float*>(qj + TM * TQ);
    lj = reinterpret_cast<int*>(ld + TM * LW);
    taud = reinterpret_cast<float*>(lj + TM * LW);
    tauj = reinterpret_cast<int*>(taud + TM);
    qcnt = tauj + TM;
  }
  __device__ __forceinline__ void init_row(int row, int lane) {
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      ld[row * LW + s * 32 + lane] = __int_as_float(0x7f800000);
      lj[row * LW + s * 32 + lane] = 0x7fffffff;
    }
    if (lane == 0) {
      taud[row] = __int_as_float(0x7f800000);
      tauj[row] = 0x7fffffff;
      qcnt[row] = 0;
    }
  }
  // merge queue entries [base, base+n) (n <= 32) of `row` into its list; refresh tau
  __device__ __noinline__ void drain(int row, int base, int n, int k, int lane) {
    RowSel<KS> R;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      R.d[s] = ld[row * LW + s * 32 + lane];
      R.j[s] = lj[row * LW + s * 32 + lane];
    }
    const float bd = lane < n ? qd[row * TQ + base + lane] : __int_as_float(0x7f800000);
    const int bj = lane < n ? qj[row * TQ + base + lane] : 0x7fffffff;
    R.merge_batch(bd, bj, k, lane);
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      ld[row * LW + s * 32 + lane] = R.d[s];
      lj[row * LW + s * 32 + lane] = R.j[s];
    }
    // the hint bound (if any) stays in force until the list itself holds k entries below it
    const float otd = taud[row];
    const int otj = tauj[row];
    __syncwarp();
    if (lane == 0 && lex_less(R.td, R.tj, otd, otj)) {
      taud[row] = R.td;
      tauj[row] = R.tj;
    }
  }
  // Warm start: distances from `row` to k distinct hinted columns bound the k-th smallest distance from above.
  // Same operation order as the tile loop, so the hinted columns themselves always pass the filter.
  __device__ __forceinline__ void hint_bound(int rowl, const float* __restrict__ xb, const float* __restrict__ sb,
                                             const int32_t* __restrict__ hrow, int row, int N, int C, int k, int lane) {
    float m = -__int_as_float(0x7f800000);
    const float* xi = xb + (size_t)row * C;
    const float si = sb[row];
    for (int h = lane; h < k; h += 32) {
      int j = hrow[h];
      j = j < 0 ? 0 : (j >= N ? N - 1 : j);
      const float* xj = xb + (size_t)j * C;
      float pacc = 0.0f;
#pragma unroll 8
      for (int c = 0; c < C; ++c) pacc = __fmaf_rn(__ldg(xi + c), __ldg(xj + c), pacc);
      const float dd = __fadd_rn(__fsub_rn(__fadd_rn(si, sb[j]), __fmul_rn(2.0f, pacc)), 0.0f);
      m = fmaxf(m, dd);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    if (lane == 0) {
      taud[rowl] = m;
      tauj[rowl] = 0x7fffffff;
    }
  }
};

// Tiled distance kernel.  CTA = 64 query rows of one cloud x all candidate columns, streamed in
// 128-column tiles and 16-channel stages through a 2-deep cp.async ring.  Thread micro-tile:
// 8 rows (the warp's rows, smem-broadcast) x 4 columns (lane*4..+3).  The same warp that computes a
// row also selects for it, so selection needs no CTA barrier.
template <int KS, bool WRITE_D>
__global__ void __launch_bounds__(KNN_THREADS, 2)
    knn_tile_kernel(const float* __restrict__ xT, const float* __restrict__ s, const float* __restrict__ x,
                    const int32_t* __restrict__ hint, const int32_t* __restrict__ rowflags, int N, int Npad, int C,
                    int k, int32_t* __restrict__ idx, float* __restrict__ D) {
  __shared__ __align__(16) float As[2][CK][TM];
  __shared__ __align__(16) float Bs[2][CK][TN];
  __shared__ __align__(16) float sBs[2][TN];
  __shared__ float sAs[TM];
  extern __shared__ __align__(16) unsigned char knn_dyn_smem[];  // selection state (selection variant only)

  const int b = blockIdx.y;
  const int r0 = blockIdx.x * TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xTb = xT + (size_t)b * C * Npad;
  const float* sb = s + (size_t)b * Npad;
  const int T = Npad / TN;
  const int Q = (C + CK - 1) / CK;
  const int S = T * Q;

  if (!WRITE_D && rowflags != nullptr) {
    // fallback mode behind the tensor-core filter: only CTAs that own an uncertified row do any work
    int f = 0;
    if (tid < TM && r0 + tid < N) f = rowflags[(size_t)b * N + r0 + tid];
    if (!__syncthreads_or(f)) return;
  }
  if (tid < TM) sAs[tid] = sb[r0 + tid];

  KnnSel<KS> sel;
  if (!WRITE_D) {
    sel.carve(knn_dyn_smem);
#pragma unroll 1
    for (int r = 0; r < 8; ++r) sel.init_row(warp * 8 + r, lane);
    __syncwarp();
    if (hint != nullptr) {
#pragma unroll 1
      for (int r = 0; r < 8; ++r) {
        const int row = r0 + warp * 8 + r;
        if (row < N)
          sel.hint_bound(warp * 8 + r, x + (size_t)b * N * C, sb, hint + ((size_t)b * N + row) * k, row, N, C, k, lane);
      }
      __syncwarp();
    }
  }

  auto load_stage = [&](int st) {
    const int t = st / Q, q = st - t * Q, buf = st & 1;
    const int c0 = q * CK;
    const int ck = min(CK, C - c0);
    for (int e = tid; e < ck * (TM / 4); e += KNN_THREADS) {
      const int cc = e / (TM / 4), v = e % (TM / 4);
      cp_async16(&As[buf][cc][v * 4], xTb + (size_t)(c0 + cc) * Npad + r0 + v * 4);
    }
    for (int e = tid; e < ck * (TN / 4); e += KNN_THREADS) {
      const int cc = e / (TN / 4), v = e % (TN / 4);
      cp_async16(&Bs[buf][cc][v * 4], xTb + (size_t)(c0 + cc) * Npad + t * TN + v * 4);
    }
    if (q == 0 && tid < TN / 4) cp_async16(&sBs[t & 1][tid * 4], sb + t * TN + tid * 4);
  };

  float acc[8][4];
  load_stage(0);
  cp_async_commit();
  for (int st = 0; st < S; ++st) {
    if (st + 1 < S) load_stage(st + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int t = st / Q, q = st - t * Q, buf = st & 1;
    const int ck = min(CK, C - q * CK);
    if (q == 0) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
    }
#pragma unroll 4
    for (int cc = 0; cc < ck; ++cc) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][cc][warp * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][cc][warp * 8 + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][cc][lane * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = __fmaf_rn(a[r], bb[c], acc[r][c]);
    }
    if (q == Q - 1) {
      const float4 sj4 = *reinterpret_cast<const float4*>(&sBs[t & 1][lane * 4]);
      const float sj[4] = {sj4.x, sj4.y, sj4.z, sj4.w};
      const int col0 = t * TN + lane * 4;
      // ops.py:16: (s_i + s_j) - 2*p ; +0 canonicalises -0.  acc now holds distances.
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float si = sAs[warp * 8 + r];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          acc[r][c] = __fadd_rn(__fsub_rn(__fadd_rn(si, sj[c]), __fmul_rn(2.0f, acc[r][c])), 0.0f);
      }
      if (WRITE_D) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int row = r0 + warp * 8 + r;
          if (row < N) {
            float* drow = D + ((size_t)b * N + row) * N;
            if ((N & 3) == 0 && col0 + 3 < N) {
              *reinterpret_cast<float4*>(drow + col0) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (col0 + c < N) drow[col0 + c] = acc[r][c];
            }
          }
        }
      } else {
        if (col0 + 3 >= N) {  // ragged last tile: columns >= N become NaN and fail every comparison
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (col0 + c >= N) acc[r][c] = __int_as_float(0x7fc00000);
        }
        // stage the warp's 8x128 distances, then scan them with a different thread mapping: lane = (row, quarter),
        // each lane filters ITS OWN elements against its row's threshold (one compare per element, no collectives);
        // survivors go to the row's queue through a shared counter; the warp drains a row (sort + merge) only when
        // 32 candidates are queued.  Chunks of 32 columns bound the queue at 31 + 32 entries.
#pragma unroll
        for (int r = 0; r < 8; ++r)
          *reinterpret_cast<float4*>(&sel.dst[(warp * 8 + r) * DLD + lane * 4]) =
              make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        __syncwarp();
        const int srow = warp * 8 + (lane >> 2);   // the row this lane scans
        const int sq = lane & 3;
        const float* drow = sel.dst + srow * DLD + sq;
#pragma unroll 1
        for (int ch = 0; ch < TN / 32; ++ch) {
          const float td = sel.taud[srow];
          const int tj = sel.tauj[srow];
          float dv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = drow[ch * 32 + 4 * i];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (dv[i] <= td) {                                   // superset test; NaN (ragged columns) fails
              const int col = t * TN + ch * 32 + 4 * i + sq;
              if (lex_less(dv[i], col, td, tj)) {
                const int pos = atomicAdd(&sel.qcnt[srow], 1);
                sel.qd[srow * TQ + pos] = dv[i];
                sel.qj[srow * TQ + pos] = col;
              }
            }
          }
          __syncwarp();
          const int cnt = sel.qcnt[srow];
          unsigned need = __ballot_sync(FULL, cnt >= 32 && sq == 0);
          while (need) {                                         // warp-uniform
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const int c = __shfl_sync(FULL, cnt, src) - 32;
            const int row = warp * 8 + (src >> 2);
            sel.drain(row, c, 32, k, lane);
            if (lane == 0) sel.qcnt[row] = c;
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
  }
  if (!WRITE_D) {
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
      const int rowl = warp * 8 + r;
      const int c = sel.qcnt[rowl];
      if (c > 0) sel.drain(rowl, 0, c, k, lane);
      __syncwarp();
      const int row = r0 + rowl;
      if (row < N) {
        int32_t* o = idx + ((size_t)b * N + row) * k;
        for (int pos = lane; pos < k; pos += 32) o[pos] = sel.lj[rowl * KnnSel<KS>::LW + pos];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row top-k on a materialised matrix (ops.py:18 alone): one warp per row, coalesced 128 B reads.
template <int KS>
__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ D, int64_t rows, int N, int k,
                                                        int32_t* __restrict__ idx) {
  __shared__ float qd_s[8][QCAP];
  __shared__ int qj_s[8][QCAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float* dr = D + row * (int64_t)N;
  RowSel<KS> R;
  R.init();
  for (int c0 = 0; c0 < N; c0 += 128) {
    float dv[4];
    int cj[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cj[q] = c0 + q * 32 + lane;
      dv[q] = (cj[q] < N) ? __fadd_rn(__ldg(dr + cj[q]), 0.0f) : 0.0f;
    }
    R.offer4(dv, cj, N, k, qd_s[warp], qj_s[warp], lane);
  }
  R.finish(k, qd_s[warp], qj_s[warp], lane);
  int32_t* o = idx + row * (int64_t)k;
#pragma unroll
  for (int sl = 0; sl < KS; ++sl) {
    const int pos = sl * 32 + lane;
    if (pos < k) o[pos] = R.j[sl];
  }
}


static inline int npad_of(int N) { return ((N + TN - 1) / TN) * TN; }

static int knn_common(const float* x, int B, int N, int C, void* ws, size_t ws_bytes, cudaStream_t st,
                      float** xT, float** s, int* Npad) {
  DG_REQUIRE(x && ws, DGCNN_ERR_INVALID, "knn: null pointer");
  DG_REQUIRE(B > 0 && N > 0 && C > 0, DGCNN_ERR_INVALID, "knn: bad shape B=%d N=%d C=%d", B, N, C);
  DG_REQUIRE(B <= 65535, DGCNN_ERR_UNSUPPORTED, "knn: B=%d > 65535", B);
  DG_REQUIRE(((uintptr_t)ws & 15) == 0, DGCNN_ERR_INVALID, "knn: workspace must be 16-byte aligned");
  DG_REQUIRE(ws_bytes >= dgcnn_knn_workspace_bytes(B, N, C), DGCNN_ERR_WORKSPACE,
             "knn: workspace %zu < %zu bytes", ws_bytes, dgcnn_knn_workspace_bytes(B, N, C));
  *Npad = npad_of(N);
  *xT = reinterpret_cast<float*>(ws);
  *s = *xT + (size_t)B * C * (*Npad);
  dim3 g(cdiv(*Npad, 256), B);
  knn_prep_kernel<<<g, 256, 0, st>>>(x, *xT, *s, N, *Npad, C);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_prep_kernel");
  return DGCNN_OK;
}

}  // namespace dgcnn

using namespace dgcnn;

namespace dgcnn {
// knn_tc.cu: tensor-core filter + exact refinement
bool knn_tc_eligible(int B, int N, int C, int k);
size_t knn_tc_bytes(int B, int N, int C, int k_max);
int knn_tc_run(const float* x, int32_t* idx, int B, int N, int C, int k, int filter_mode, void* ws, cudaStream_t st);
static inline size_t knn_base_bytes(int B, int N, int C) {
  const size_t Npad = npad_of(N);
  return ((((size_t)B * C * Npad + (size_t)B * Npad) * sizeof(float)) + 255) & ~(size_t)255;
}
}  // namespace dgcnn

extern "C" size_t dgcnn_knn_workspace_bytes(int B, int N, int C) {
  if (B <= 0 || N <= 0 || C <= 0) return 0;
  size_t n = knn_base_bytes(B, N, C);
  if (knn_tc_eligible(B, N, C, 1)) n += knn_tc_bytes(B, N, C, 48);
  return n;
}

extern "C" int dgcnn_pairwise_distance(const float* x, float* D, int B, int N, int C, void* ws, size_t ws_bytes,
                                       dgcnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DG_REQUIRE(D, DGCNN_ERR_INVALID, "pairwise_distance: null output");
  DG_REQUIRE(((uintptr_t)D & 15) == 0, DGCNN_ERR_INVALID, "pairwise_distance: D must be 16-byte aligned");
  float *xT = nullptr, *s = nullptr;
  int Npad = 0;
  int rc = knn_common(x, B, N, C, ws, ws_bytes, st, &xT, &s, &Npad);
  if (rc) return rc;
  dim3 g(cdiv(N, TM), B);
  knn_tile_kernel<1, true><<<g, KNN_THREADS, 0, st>>>(xT, s, x, nullptr, nullptr, N, Npad, C, 1, nullptr, D);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tile_kernel<D>");
  return DGCNN_OK;
}

extern "C" int dgcnn_knn(const float* x, int32_t* idx, int B, int N, int C, int k, void* ws, size_t ws_bytes,
                         dgcnn_stream_t stream) {
  return dgcnn_knn_hinted(x, nullptr, idx, B, N, C, k, ws, ws_bytes, stream);
}

extern "C" int dgcnn_knn_hinted(const float* x, const int32_t* hint, int32_t* idx, int B, int N, int C, int k,
                                void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  return dgcnn_knn_mode(x, hint, idx, B, N, C, k, DGCNN_KNN_AUTO, ws, ws_bytes, stream);
}

extern "C" int dgcnn_knn_mode(const float* x, const int32_t* hint, int32_t* idx, int B, int N, int C, int k,
                              int filter_mode, void* ws, size_t ws_bytes, dgcnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DG_REQUIRE(filter_mode >= DGCNN_KNN_AUTO && filter_mode <= DGCNN_KNN_FINE, DGCNN_ERR_INVALID,
             "knn: filter_mode must be DGCNN_KNN_AUTO, _COARSE or _FINE");
  DG_REQUIRE(idx, DGCNN_ERR_INVALID, "knn: null output");
  DG_REQUIRE(k >= 1 && k <= N, DGCNN_ERR_INVALID, "knn: need 1 <= k <= N (k=%d, N=%d)", k, N);
  DG_REQUIRE(k <= DGCNN_KNN_MAX_K, DGCNN_ERR_UNSUPPORTED, "knn: k=%d > %d", k, DGCNN_KNN_MAX_K);
  if (knn_tc_eligible(B, N, C, k)) {
    // tensor-core filter + exact refinement (knn_tc.cu); the warm start is not needed on this path
    DG_REQUIRE(x && ws, DGCNN_ERR_INVALID, "knn: null pointer");
    DG_REQUIRE(B > 0 && B <= 65535, DGCNN_ERR_UNSUPPORTED, "knn: B=%d outside [1, 65535]", B);
    DG_REQUIRE(((uintptr_t)ws & 255) == 0 && ((uintptr_t)x & 15) == 0, DGCNN_ERR_INVALID,
               "knn: workspace must be 256-byte and x 16-byte aligned");
    DG_REQUIRE(ws_bytes >= dgcnn_knn_workspace_bytes(B, N, C), DGCNN_ERR_WORKSPACE, "knn: workspace %zu < %zu bytes",
               ws_bytes, dgcnn_knn_workspace_bytes(B, N, C));
    return knn_tc_run(x, idx, B, N, C, k, filter_mode, reinterpret_cast<unsigned char*>(ws) + knn_base_bytes(B, N, C), st);
  }
  float *xT = nullptr, *s = nullptr;
  int Npad = 0;
  int rc = knn_common(x, B, N, C, ws, ws_bytes, st, &xT, &s, &Npad);
  if (rc) return rc;
  const int32_t* rowflags = nullptr;
  dim3 g(cdiv(N, TM), B);
  static bool attr_done = false;  // raise the dynamic-smem cap once (idempotent, benign if raced)
  if (!attr_done) {
    cudaFuncSetAttribute(knn_tile_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)KnnSel<1>::bytes());
    cudaFuncSetAttribute(knn_tile_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)KnnSel<2>::bytes());
    attr_done = true;
  }
  if (k <= 32)
    knn_tile_kernel<1, false><<<g, KNN_THREADS, KnnSel<1>::bytes(), st>>>(xT, s, x, hint, rowflags, N, Npad, C, k, idx, nullptr);
  else
    knn_tile_kernel<2, false><<<g, KNN_THREADS, KnnSel<2>::bytes(), st>>>(xT, s, x, hint, rowflags, N, Npad, C, k, idx, nullptr);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tile_kernel");
  return DGCNN_OK;
}

extern "C" int dgcnn_topk_rows(const float* D, int32_t* idx, int64_t rows, int N, int k, dgcnn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DG_REQUIRE(D && idx, DGCNN_ERR_INVALID, "topk_rows: null pointer");
  DG_REQUIRE(rows > 0 && N > 0, DGCNN_ERR_INVALID, "topk_rows: bad shape rows=%lld N=%d", (long long)rows, N);
  DG_REQUIRE(k >= 1 && k <= N, DGCNN_ERR_INVALID, "topk_rows: need 1 <= k <= N (k=%d, N=%d)", k, N);
  DG_REQUIRE(k <= DGCNN_KNN_MAX_K, DGCNN_ERR_UNSUPPORTED, "topk_rows: k=%d > %d", k, DGCNN_KNN_MAX_K);
  const int wpb = 8;
  const int64_t blocks = (rows + wpb - 1) / wpb;
  DG_REQUIRE(blocks < (1ll << 31), DGCNN_ERR_UNSUPPORTED, "topk_rows: too many rows");
  if (k <= 32)
    topk_rows_kernel<1><<<(unsigned)blocks, wpb * 32, 0, st>>>(D, rows, N, k, idx);
  else
    topk_rows_kernel<2><<<(unsigned)blocks, wpb * 32, 0, st>>>(D, rows, N, k, idx);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("topk_rows_kernel");
  return DGCNN_OK;
}
