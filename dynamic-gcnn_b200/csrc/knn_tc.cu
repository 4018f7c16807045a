// k_nn on the 5th-gen tensor cores, bit-exact.  /root/reference/dgcnn/ops.py:8-19 for the layers whose input is a
// 64-channel feature map and whose previous graph is known (every EdgeConv layer after the first, ops.py:91-96).
//
//   K1 knn_tc_filter_kernel : G~ = X.X^T by tcgen05 (bf16 hi/lo split, 3 MMAs per k-slice, fp32 accumulate in
//      TMEM, double-buffered), D~ = (s_i+s_j) - 2G~.  |D~ - D_exact| <= delta_ij = EPS*(s_i+s_j) (bound below), so
//      every column gets an interval [D~-delta, D~+delta] around the oracle's fp32 value.  One epilogue thread per
//      query row scans the accumulator straight out of TMEM: a column is a candidate iff its LOWER bound is <= the
//      row's threshold U (an upper bound on the k-th smallest exact distance: first the largest exact distance to
//      the k hinted neighbours, later the k-th smallest UPPER bound seen).  Candidates go to the row's queue; the
//      warp sorts+merges 32 at a time into the row's list of the 32*KS smallest upper bounds.
//      Certificate: if the list is full and (largest kept upper bound - 2*delta_max) does not clear U, an excluded
//      column could still belong to the exact top-k (massive near-ties, e.g. voxel lattices) -> the row is flagged.
//   K2 knn_tc_refine_kernel : exact fp32 distances (the oracle's fmaf chain) of the <= 32*KS candidates, warp sort by
//      (distance, index), first k written.  Since candidates are a superset of the exact top-k, the result equals
//      the oracle bit for bit, tie order included.
//   K3 : flagged rows are recomputed by the exact SIMT kernel (knn.cu), which skips CTAs without flags.
//
// Error bound.  With s = fl-sum of squares (>= 0.999 ||x||^2) and |x_i.x_j| <= sqrt(s_i s_j) <= (s_i+s_j)/2:
//   operand split   : x = hi + lo + e, |e| <= 2^-16|x|; dropped lo.lo <= 2^-16|x||y|     -> 3 * 2^-16
//   accumulation    : <= 200 fp32 adds (possibly truncating) of exact bf16 products         -> 200 * 2^-23 < 2^-15
//   oracle fmaf chain vs real dot product (C <= 64)                                        -> 64 * 2^-24 = 2^-18
//   final (s_i+s_j) - 2p roundings                                                         -> < 2^-22 (s_i+s_j)
// => |D~ - D| <= 2 * (3*2^-16 + 2^-15 + 2^-18) * (s_i+s_j)/2 < 2^-13 (s_i+s_j).  EPS = 2^-11 keeps a 4x margin.
#include "knn_select.cuh"
#include "tc_common.cuh"

namespace dgcnn {

constexpr int KT_ROWS = 128;      // query rows per CTA = TMEM lanes
constexpr int KT_COLS = 128;      // candidate columns per tile = accumulator columns
constexpr int KT_SCAN_WARPS = 8;  // two warps per TMEM sub-partition, each scanning one half of the tile's columns
constexpr int KT_THREADS = 64 + 32 * KT_SCAN_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2..9 scan
constexpr int KT_QC = 41;         // queue pitch (odd: conflict-free when every lane appends); 31 + 8 entries max
constexpr int KT_CHK = 8;         // columns between drain checks
constexpr int KT_LW = 32;         // list width per (row, column half)
constexpr int KT_RH = 2 * KT_ROWS;  // (row, half) records
constexpr float KT_EPS = 1.0f / 2048.0f;
// Fast test: lower bound <= U  <=>  a(1-EPS) - 2g <= U  <=>  g - 0.5(1-EPS)s_j >= 0.5((1-EPS)s_i - U).  Evaluated with a
// slightly smaller (1-EPS) factor and a relative slack so that fp32 reassociation can only ADMIT more, never fewer.
constexpr float KT_HALF1ME = 0.5f * (1.0f - 1.0f / 2048.0f) * (1.0f - 1.0f / 1048576.0f);
constexpr uint32_t KT_TILE = KT_ROWS * 64 * 2;  // one bf16 plane tile, 16 KB

// byte offsets into the (1024-aligned) dynamic shared memory
constexpr size_t KT_A_OFF = 0;                                       // A hi, lo
constexpr size_t KT_B_OFF = KT_A_OFF + 2 * KT_TILE;                  // B hi, lo (single stage: the MMA of a tile is
                                                                     // far shorter than its scan; TMEM is double-buffered)
constexpr size_t KT_QD_OFF = KT_B_OFF + 2 * KT_TILE;
constexpr size_t KT_QJ_OFF = KT_QD_OFF + (size_t)KT_RH * KT_QC * 4;
constexpr size_t KT_LD_OFF = KT_QJ_OFF + (size_t)KT_RH * KT_QC * 4;
constexpr size_t KT_LJ_OFF = KT_LD_OFF + (size_t)KT_RH * KT_LW * 4;
constexpr size_t KT_TAUD_OFF = KT_LJ_OFF + (size_t)KT_RH * KT_LW * 4;
constexpr size_t KT_TAUJ_OFF = KT_TAUD_OFF + KT_RH * 4;
constexpr size_t KT_SB_OFF = KT_TAUJ_OFF + KT_RH * 4;                // 2 x 128 column norms
constexpr size_t KT_HB_OFF = KT_SB_OFF + 2 * KT_COLS * 4;                // 2 x 128 pre-scaled column norms
constexpr size_t KT_BAR_OFF = KT_HB_OFF + 2 * KT_COLS * 4;
constexpr size_t KT_SMEM = KT_BAR_OFF + 128;

extern __shared__ __align__(1024) unsigned char kt_smem[];

// warp-cooperative: merge queue entries [base, base+n) of record `rh` into its list (ascending by (upper bound, index))
__device__ __noinline__ void kt_drain(int rh, int base, int n, int k, int lane) {
  float* qd = reinterpret_cast<float*>(kt_smem + KT_QD_OFF);
  int* qj = reinterpret_cast<int*>(kt_smem + KT_QJ_OFF);
  float* ld = reinterpret_cast<float*>(kt_smem + KT_LD_OFF);
  int* lj = reinterpret_cast<int*>(kt_smem + KT_LJ_OFF);
  float* taud = reinterpret_cast<float*>(kt_smem + KT_TAUD_OFF);
  int* tauj = reinterpret_cast<int*>(kt_smem + KT_TAUJ_OFF);
  RowSel<1> R;
  R.d[0] = ld[rh * KT_LW + lane];
  R.j[0] = lj[rh * KT_LW + lane];
  const float bd = lane < n ? qd[rh * KT_QC + base + lane] : __int_as_float(0x7f800000);
  const int bj = lane < n ? qj[rh * KT_QC + base + lane] : 0x7fffffff;
  R.merge_batch(bd, bj, k, lane);
  ld[rh * KT_LW + lane] = R.d[0];
  lj[rh * KT_LW + lane] = R.j[0];
  const float otd = taud[rh];
  const int otj = tauj[rh];
  __syncwarp();
  if (lane == 0 && lex_less(R.td, R.tj, otd, otj)) {
    taud[rh] = R.td;
    tauj[rh] = R.tj;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(KT_THREADS, 1)
    knn_tc_filter_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ x,
                         const float* __restrict__ s, const float* __restrict__ smax,
                         const float* __restrict__ ubound, int N, int Npad, int C, int k,
                         int32_t* __restrict__ cand, int32_t* __restrict__ flags) {
  unsigned char* sm = kt_smem;
  float* qd = reinterpret_cast<float*>(kt_smem + KT_QD_OFF);
  int* qj = reinterpret_cast<int*>(kt_smem + KT_QJ_OFF);
  float* ld = reinterpret_cast<float*>(kt_smem + KT_LD_OFF);
  int* lj = reinterpret_cast<int*>(kt_smem + KT_LJ_OFF);
  float* taud = reinterpret_cast<float*>(kt_smem + KT_TAUD_OFF);
  int* tauj = reinterpret_cast<int*>(kt_smem + KT_TAUJ_OFF);
  float* sB = reinterpret_cast<float*>(kt_smem + KT_SB_OFF);
  float* hB = reinterpret_cast<float*>(kt_smem + KT_HB_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(kt_smem + KT_BAR_OFF);
  uint64_t* a_full = bars;            // 1
  uint64_t* b_full = bars + 1;        // 2
  uint64_t* b_empty = bars + 3;       // 2
  uint64_t* acc_full = bars + 5;      // 2
  uint64_t* acc_empty = bars + 7;     // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * KT_ROWS;
  const int T = (N + KT_COLS - 1) / KT_COLS;
  const int64_t cloud0 = (int64_t)b * N;  // first global row of this cloud

  if (threadIdx.x == 0) {
    if (smem_u32(kt_smem) & 1023u) __trap();  // 128B-swizzled operand tiles need a 1024-aligned base
    mbar_init(a_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], KT_SCAN_WARPS);  // one arrival per scan warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // two 128-column fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(a_full, 2 * KT_TILE);
      tma_load_3d(sm + KT_A_OFF, &tmX, 0, (int)(cloud0 + r0), 0, a_full);
      tma_load_3d(sm + KT_A_OFF + KT_TILE, &tmX, 0, (int)(cloud0 + r0), 1, a_full);
      for (int t = 0; t < T; ++t) {
        mbar_wait(&b_empty[0], (t & 1) ^ 1);
        mbar_expect_tx(&b_full[0], 2 * KT_TILE);
        unsigned char* dst = sm + KT_B_OFF;
        tma_load_3d(dst, &tmX, 0, (int)(cloud0 + (int64_t)t * KT_COLS), 0, &b_full[0]);
        tma_load_3d(dst + KT_TILE, &tmX, 0, (int)(cloud0 + (int64_t)t * KT_COLS), 1, &b_full[0]);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KT_COLS >> 3) << 17) |
                           ((uint32_t)(KT_ROWS >> 4) << 24);
    const int nks = (C + 15) / 16;  // k-slices of 16 channels (zero-filled beyond C)
    mbar_wait(a_full, 0);
    for (int t = 0; t < T; ++t) {
      const int st = t & 1;
      const uint32_t ph = (t >> 1) & 1;
      mbar_wait(&b_full[0], t & 1);
      mbar_wait(&acc_empty[st], ph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(sm + KT_A_OFF), a_lo = a_hi + KT_TILE;
        const uint32_t b_hi = smem_u32(sm + KT_B_OFF), b_lo = b_hi + KT_TILE;
        const uint32_t acc = tmem_base + (uint32_t)(st * KT_COLS);
        for (int ks = 0; ks < nks; ++ks) {
          const uint64_t dah = umma_desc(a_hi + ks * 32, 16, 1024), dal = umma_desc(a_lo + ks * 32, 16, 1024);
          const uint64_t dbh = umma_desc(b_hi + ks * 32, 16, 1024), dbl = umma_desc(b_lo + ks * 32, 16, 1024);
          umma_bf16(acc, dal, dbh, idesc, ks != 0);
          umma_bf16(acc, dah, dbl, idesc, 1);
          umma_bf16(acc, dah, dbh, idesc, 1);
        }
        umma_commit(&b_empty[0]);
        umma_commit(&acc_full[st]);
      }
      __syncwarp();
    }
  } else {
    // ---------------- scan warps: thread <-> (query row = TMEM lane, column half) ----------------
    const int sub = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rowl = sub * 32 + lane;            // row inside the CTA
    const int rh = half * KT_ROWS + rowl;        // this thread's record
    const int row = r0 + rowl;                   // row inside the cloud
    const bool valid = row < N;
    const int e = threadIdx.x - 64;              // 0..255 among scan threads
    const float* sb = s + (size_t)b * Npad;
    const float si = valid ? sb[row] : 0.0f;
    const float ninf = -__int_as_float(0x7f800000);
#pragma unroll 1
    for (int i = 0; i < KT_LW; ++i) {
      ld[rh * KT_LW + i] = __int_as_float(0x7f800000);
      lj[rh * KT_LW + i] = 0x7fffffff;
    }
    // warm start (knn_hint_bound_kernel): an upper bound on the k-th smallest exact distance of this row
    const float u0 = valid ? ubound[cloud0 + row] : ninf;   // rows beyond N never admit anything
    taud[rh] = u0;
    tauj[rh] = 0x7fffffff;
    __syncwarp();
    const int other = (half ^ 1) * KT_ROWS + rowl;   // the record of the thread scanning the other column half
    int cnt = 0;
    for (int t = 0; t < T; ++t) {
      const int st = t & 1;
      // column norms of this tile -> smem (+inf for ragged columns: they then fail every test), then the accumulator
      if (e < KT_COLS) {
        const float sj = (t * KT_COLS + e < N) ? sb[t * KT_COLS + e] : __int_as_float(0x7f800000);
        sB[st * KT_COLS + e] = sj;
        hB[st * KT_COLS + e] = KT_HALF1ME * sj;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&acc_full[st], (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float* sBt = sB + st * KT_COLS + half * 64;
      const float* hBt = hB + st * KT_COLS + half * 64;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(st * KT_COLS + half * 64 + ch * 32), v);
#pragma unroll
        for (int hf = 0; hf < 32 / KT_CHK; ++hf) {
          const float uthr = fminf(taud[rh], taud[other]);
          // per-row constant of the fast test (conservative: decreased by a relative + absolute slack)
          const float ci0 = KT_HALF1ME * si - 0.5f * uthr;
          const float ci = ci0 - (fabsf(ci0) + fabsf(si)) * (1.0f / 524288.0f);
          const float4 h0 = *reinterpret_cast<const float4*>(hBt + ch * 32 + hf * KT_CHK);
          const float4 h1 = *reinterpret_cast<const float4*>(hBt + ch * 32 + hf * KT_CHK + 4);
          const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
          unsigned mask = 0;
#pragma unroll
          for (int i = 0; i < KT_CHK; ++i)
            mask |= (__uint_as_float(v[hf * KT_CHK + i]) - hh[i] >= ci) ? (1u << i) : 0u;
          // exact interval test + append, one candidate per lane per iteration (iterations = max popcount in the warp)
          while (__any_sync(FULL, mask != 0)) {
            if (mask) {
              const int i = __ffs(mask) - 1;
              mask &= mask - 1;
              float g = __uint_as_float(v[hf * KT_CHK]);
#pragma unroll
              for (int q = 1; q < KT_CHK; ++q) g = (i == q) ? __uint_as_float(v[hf * KT_CHK + q]) : g;
              const float a = si + sBt[ch * 32 + hf * KT_CHK + i];
              const float dt = fmaf(-2.0f, g, a);
              if (fmaf(-KT_EPS, a, dt) <= uthr) {
                qd[rh * KT_QC + cnt] = fmaf(KT_EPS, a, dt);   // upper bound of the interval
                qj[rh * KT_QC + cnt] = t * KT_COLS + half * 64 + ch * 32 + hf * KT_CHK + i;
                ++cnt;
              }
            }
          }
          unsigned need = __ballot_sync(FULL, cnt >= 32);
          while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const int c = __shfl_sync(FULL, cnt, src) - 32;
            kt_drain(half * KT_ROWS + sub * 32 + src, c, 32, k, lane);
            if (lane == src) cnt = c;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[st]);
    }
    // flush the queues, certify, emit candidates
    __syncwarp();
    for (int src = 0; src < 32; ++src) {
      const int c = __shfl_sync(FULL, cnt, src);
      if (c > 0) kt_drain(half * KT_ROWS + sub * 32 + src, 0, c, k, lane);
    }
    __syncwarp();
    asm volatile("bar.sync 1, 256;" ::: "memory");   // both halves of every row are final
    // merge the two half-lists of each row into the 32 smallest upper bounds overall (warp-cooperative, 16 rows
    // per warp), certify, emit.  Every excluded column has an upper bound >= the merged list's largest entry.
#pragma unroll 1
    for (int q = 0; q < 16; ++q) {
      const int rl = sub * 32 + half * 16 + q;
      const int grow = r0 + rl;
      if (grow >= N) break;                       // warp-uniform
      RowSel<1> R;
      R.d[0] = ld[rl * KT_LW + lane];
      R.j[0] = lj[rl * KT_LW + lane];
      R.merge_batch(ld[(KT_ROWS + rl) * KT_LW + lane], lj[(KT_ROWS + rl) * KT_LW + lane], k, lane);
      const float uthr = fminf(fminf(taud[rl], taud[KT_ROWS + rl]), R.td);
      const float w = __shfl_sync(FULL, R.d[0], 31);                      // largest kept upper bound (+inf if not full)
      const float dmax2 = 2.0f * KT_EPS * (sb[grow] + smax[b]);
      cand[(cloud0 + grow) * KT_LW + lane] = R.j[0];
      if (lane == 0 && (w - dmax2 <= uthr)) flags[cloud0 + grow] = 1;    // flags are zeroed by the host wrapper
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// exact fp32 distances of the candidates, sort by (distance, index), write the first k
template <int KS>
__global__ void __launch_bounds__(256)
    knn_tc_refine_kernel(const float* __restrict__ x, const float* __restrict__ s, const int32_t* __restrict__ cand,
                         int N, int Npad, int C, int k, int64_t P, int32_t* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const int b = (int)(p / N);
  const int row = (int)(p - (int64_t)b * N);
  const float* sb = s + (size_t)b * Npad;
  const float* xi = x + p * C;
  const float si = sb[row];
  RowSel<KS> R;
  R.init();
#pragma unroll
  for (int q = 0; q < KS; ++q) {
    const int j = cand[p * (32 * KS) + q * 32 + lane];
    float d = __int_as_float(0x7f800000);
    if (j >= 0 && j < N) {
      const float* xj = x + ((int64_t)b * N + j) * C;
      float acc = 0.0f;
      for (int c = 0; c < C; c += 4) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(xi + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(xj + c));
        acc = __fmaf_rn(a4.x, b4.x, acc);
        acc = __fmaf_rn(a4.y, b4.y, acc);
        acc = __fmaf_rn(a4.z, b4.z, acc);
        acc = __fmaf_rn(a4.w, b4.w, acc);
      }
      d = __fadd_rn(__fsub_rn(__fadd_rn(si, sb[j]), __fmul_rn(2.0f, acc)), 0.0f);
    }
    R.merge_batch(d, (j >= 0 && j < N) ? j : 0x7fffffff, k, lane);
  }
  int32_t* o = idx + p * k;
#pragma unroll
  for (int q = 0; q < KS; ++q) {
    const int pos = q * 32 + lane;
    if (pos < k) o[pos] = R.j[q];
  }
}

// Warm start: for every row an upper bound on its k-th smallest exact distance = the largest distance to k distinct
// hinted columns.  Coalesced (warp per row, lanes over channels) with a shuffle-tree dot product, so the value can
// differ from the oracle's fmaf chain by rounding: inflate by 2^-16 (s_i + s_max) (>= 60x the possible difference).
__global__ void __launch_bounds__(256)
    knn_hint_bound_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ smax,
                          const int32_t* __restrict__ hint, int N, int Npad, int C, int k, int64_t P,
                          float* __restrict__ ubound) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const int b = (int)(p / N);
  const float* sb = s + (size_t)b * Npad;
  const float* xb = x + (int64_t)b * N * C;
  const float* xi = x + p * C;
  const float a0 = lane < C ? xi[lane] : 0.0f, a1 = lane + 32 < C ? xi[lane + 32] : 0.0f;
  const int hj = lane < k ? hint[p * k + lane] : 0;     // k <= 24 on this path
  const float si = sb[p - (int64_t)b * N];
  float m = -__int_as_float(0x7f800000);
#pragma unroll 4
  for (int q = 0; q < k; ++q) {
    int j = __shfl_sync(FULL, hj, q);
    j = j < 0 ? 0 : (j >= N ? N - 1 : j);
    const float* xj = xb + (int64_t)j * C;
    float part = a0 * (lane < C ? __ldg(xj + lane) : 0.0f) + a1 * (lane + 32 < C ? __ldg(xj + lane + 32) : 0.0f);
    part = warp_sum(part);
    m = fmaxf(m, (si + sb[j]) - 2.0f * part);
  }
  if (lane == 0) ubound[p] = m + (1.0f / 65536.0f) * (si + smax[b]);
}

// K3: rows the filter could not certify (distance ties beyond its resolution) are recomputed exactly, one warp per
// flagged row: all N distances by the oracle's fmaf chain (lanes over columns), selection as in topk_rows_kernel.
__global__ void __launch_bounds__(256)
    knn_row_fallback_kernel(const float* __restrict__ x, const float* __restrict__ s, const int32_t* __restrict__ flags,
                            int N, int Npad, int C, int k, int64_t P, int32_t* __restrict__ idx) {
  __shared__ float qd_s[8][QCAP];
  __shared__ int qj_s[8][QCAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p = (int64_t)blockIdx.x * 8 + warp;
  if (p >= P || flags[p] == 0) return;
  const int b = (int)(p / N);
  const float* sb = s + (size_t)b * Npad;
  const float* xb = x + (int64_t)b * N * C;
  const float* xi = x + p * C;
  const float si = sb[p - (int64_t)b * N];
  RowSel<1> R;
  R.init();
  for (int c0 = 0; c0 < N; c0 += 128) {
    float dv[4];
    int cj[4];
    const float* xj[4];
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cj[q] = c0 + q * 32 + lane;
      xj[q] = xb + (int64_t)(cj[q] < N ? cj[q] : N - 1) * C;
    }
    for (int c = 0; c < C; c += 4) {       // four independent fmaf chains, each in the oracle's channel order
      const float4 a4 = __ldg(reinterpret_cast<const float4*>(xi + c));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(xj[q] + c));
        acc[q] = __fmaf_rn(a4.x, b4.x, acc[q]);
        acc[q] = __fmaf_rn(a4.y, b4.y, acc[q]);
        acc[q] = __fmaf_rn(a4.z, b4.z, acc[q]);
        acc[q] = __fmaf_rn(a4.w, b4.w, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dv[q] = __fadd_rn(__fsub_rn(__fadd_rn(si, sb[cj[q] < N ? cj[q] : N - 1]), __fmul_rn(2.0f, acc[q])), 0.0f);
    R.offer4(dv, cj, N, k, qd_s[warp], qj_s[warp], lane);
  }
  R.finish(k, qd_s[warp], qj_s[warp], lane);
  if (lane < k) idx[p * k + lane] = R.j[0];
}

__global__ void cloud_max_kernel(const float* __restrict__ s, int N, int Npad, float* __restrict__ smax) {
  // one block per cloud; s >= 0 so the int ordering of the bit patterns is the float ordering
  __shared__ float red[32];
  const int b = blockIdx.x;
  float m = 0.0f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) m = fmaxf(m, s[(size_t)b * Npad + n]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    if (threadIdx.x == 0) smax[b] = m;
  }
}

// defined in tc_gemm.cu
int launch_split_bf16(const float* x, int64_t rows, int cols, int64_t ldx, void* planes, int64_t ldo,
                      int64_t plane_elems, cudaStream_t st);

// Called by dgcnn_knn_hinted (knn.cu) after its prep kernel.  Layout of `extra`: planes | cand | flags | smax.
int knn_tc_run(const float* x, const float* s, const int32_t* hint, int32_t* idx, int B, int N, int Npad, int C, int k,
               void* extra, int32_t** flags_out, cudaStream_t st) {
  const int64_t P = (int64_t)B * N;
  unsigned char* base = reinterpret_cast<unsigned char*>(extra);
  void* planes = base;
  size_t off = ((size_t)2 * P * C * 2 + 255) & ~(size_t)255;
  int32_t* cand = reinterpret_cast<int32_t*>(base + off);
  off += ((size_t)P * 64 * 4 + 255) & ~(size_t)255;
  int32_t* flags = reinterpret_cast<int32_t*>(base + off);
  off += ((size_t)P * 4 + 255) & ~(size_t)255;
  float* ubound = reinterpret_cast<float*>(base + off);
  off += ((size_t)P * 4 + 255) & ~(size_t)255;
  float* smax = reinterpret_cast<float*>(base + off);
  *flags_out = flags;

  int rc = launch_split_bf16(x, P, C, C, planes, C, P * C, st);
  if (rc) return rc;
  cloud_max_kernel<<<B, 256, 0, st>>>(s, N, Npad, smax);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("cloud_max_kernel");
  knn_hint_bound_kernel<<<cdiv(P, 8), 256, 0, st>>>(x, s, smax, hint, N, Npad, C, k, P, ubound);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_hint_bound_kernel");
  CUtensorMap tm;
  rc = make_plane_map(&tm, planes, P, C, 128);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(knn_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KT_SMEM);
    attr_done = true;
  }
  if (cudaMemsetAsync(flags, 0, (size_t)P * sizeof(int32_t), st) != cudaSuccess)
    return set_err(DGCNN_ERR_CUDA, "knn_tc: memset failed");
  dim3 grid(cdiv(N, KT_ROWS), B);
  knn_tc_filter_kernel<<<grid, KT_THREADS, KT_SMEM, st>>>(tm, x, s, smax, ubound, N, Npad, C, k, cand, flags);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tc_filter_kernel");
  knn_tc_refine_kernel<1><<<cdiv(P, 8), 256, 0, st>>>(x, s, cand, N, Npad, C, k, P, idx);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tc_refine_kernel");
  knn_row_fallback_kernel<<<cdiv(P, 8), 256, 0, st>>>(x, s, flags, N, Npad, C, k, P, idx);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_row_fallback_kernel");
  return DGCNN_OK;
}

size_t knn_tc_extra_bytes(int B, int N, int C) {
  const size_t P = (size_t)B * N;
  return (((size_t)2 * P * C * 2 + 255) & ~(size_t)255) + ((P * 64 * 4 + 255) & ~(size_t)255) +
         2 * ((P * 4 + 255) & ~(size_t)255) + 256 + (size_t)B * 4;
}

}  // namespace dgcnn
