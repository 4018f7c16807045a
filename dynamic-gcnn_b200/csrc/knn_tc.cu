// k_nn on the 5th-gen tensor cores, bit-exact.  /root/reference/dgcnn/ops.py:8-19 (tf.matmul + top_k) for clouds
// with C <= 64 channels: every EdgeConv layer of the model (ops.py:91-96), the xyz layer included.
//
// The [B,N,N] matrix never exists.  One CTA owns 128 query rows of one cloud (= the 128 TMEM lanes) and sweeps the
// cloud's columns TWICE in 128-column tiles; each sweep recomputes the tile on the tensor cores (the MMA of a tile
// is far cheaper than any way of keeping 128 x N distances on chip):
//
//   operands   y = x - mean(cloud)  (distances are translation invariant; centring shrinks the norms the error
//              bound scales with), split into bf16 planes hi = bf16(y), lo = bf16(y - hi); per column the three
//              bf16 parts q of -0.5*n_j (n = |y|^2) ride in an extra 16-channel k-slice against a tile of ones, so
//              the accumulator holds   a_ij = y_i.y_j - 0.5 n_j   and   D~_ij = n_i - 2 a_ij :
//              the nearest columns of a row are simply its LARGEST accumulator entries, no per-element arithmetic.
//   sweep 1    coarse a (hi.hi only, or all three products when C <= 16): each scan thread keeps the maximum of
//              every group of W columns -> G <= 128 group maxima per row in shared memory.  The k-th largest
//              group maximum a_k (bisection) certifies k DISTINCT columns with D~ <= n_i - 2 a_k, i.e. an upper
//              bound U_i = n_i - 2 a_k + e1_i on the row's exact k-th smallest distance that is only a few ranks
//              loose (k-th largest of G random groups ~ rank G ln(G/(G-k)) of the row).
//   sweep 2    fine a (hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM): a column is a candidate iff the lower end of
//              its error interval clears U_i:  a_ij >= a_k - (e1_i + e2_i)/2.  One compare per element, the index of
//              a survivor goes to the thread's list (<= cap per (row, column half)); a row whose list overflows
//              (massive ties, e.g. duplicate points) is flagged.
//   refine     (knn_tc_refine_kernel) exact fp32 distances of the candidates by the oracle's fmaf chain, warp sort
//              by (distance, index), first k written.  The candidates are a superset of every column whose exact
//              distance is <= the exact k-th smallest, so the result equals the oracle bit for bit, tie order included.
//   fallback   flagged rows are recomputed exactly, one warp per row (knn_row_fallback_kernel).
//
// Error bounds (C <= 64; s = oracle norms of x, n = fp32 norms of y, both >= 0):
//   oracle vs real distance : fmaf chain 2^-18 * 2|x_i||x_j| / 2, the two norm sums 2^-18 (s_i+s_j), final roundings
//                             -> |D_oracle - D_real| < 2^-17 (s_i+s_j)                       ; budget  2^-16
//   centring                : y = fl(x - mu) moves D_real by < 2^-22 (n_i+n_j)
//   fine product            : operand split 3*2^-18 |y_i||y_j|, <= 13*16 (possibly truncating) fp32 adds of exact bf16
//                             products 2^-15 |y_i||y_j|, q parts exact, n_j itself 2^-18 n_j
//                             -> |D~ - D_real| < 2^-13 (n_i+n_j)                             ; budget  2^-11 (4x)
//   coarse product          : |y - hi| <= 2^-9 |y| per operand -> 2^-8 |y_i||y_j| * 2 / 2 ... < 2^-8 (n_i+n_j)/1 in D
//                             plus the fine terms                                          ; budget  0.005 (1.25x of
//                             a worst-case bound that no realistic row approaches)
//   per row i the j-dependence is removed with the cloud maxima: e_i = EPS (n_i + n_max) + 2^-16 (s_i + s_max).
#include <stdlib.h>

#include "knn_select.cuh"
#include "tc_common.cuh"

namespace dgcnn {

constexpr int K2_ROWS = 128;        // query rows per CTA = TMEM lanes
constexpr int K2_COLS = 128;        // candidate columns per tile = accumulator columns
constexpr int K2_THREADS = 64 + 256;  // warp 0 TMA, warp 1 MMA, warps 2..9 scan (2 per TMEM sub-partition)
constexpr int K2_GMAX = 128;        // group maxima per row
constexpr int K2_CAPMAX = 64;       // list entries per (row, column half)
constexpr int K2_SLACK = 8;         // writes past the cap land here (one clamp per 8 columns)
constexpr int K2_BISECT = 12;
constexpr float K2_EPS_FINE = 1.0f / 2048.0f;
constexpr float K2_EPS_COARSE = 0.005f;
constexpr float K2_EPS_ORACLE = 1.0f / 65536.0f;
constexpr float K2_PAD_NORM = 1.0e30f;   // squared norm given to padding columns: they never pass a test
constexpr uint32_t K2_TILE = K2_ROWS * 64 * 2;   // one bf16 plane tile, 128B rows, 16 KB
constexpr uint32_t K2_QTILE = K2_ROWS * 16 * 2;  // the -0.5 n_j k-slice, 32B rows, 4 KB
constexpr uint32_t K2_STAGE = 2 * K2_TILE + K2_QTILE;

// byte offsets into the (1024-aligned) dynamic shared memory
constexpr size_t K2_A_OFF = 0;                                // A hi, lo
constexpr size_t K2_ONES_OFF = K2_A_OFF + 2 * K2_TILE;        // 128 x 16 ones
constexpr size_t K2_B_OFF = K2_ONES_OFF + K2_QTILE;           // 2 stages of (B hi, B lo, q)
constexpr size_t K2_GM_OFF = K2_B_OFF + 2 * K2_STAGE;         // group maxima (sweep 1) / candidate lists (sweep 2)
constexpr size_t K2_GM_BYTES = (size_t)K2_GMAX * K2_ROWS * 4;
constexpr size_t K2_THR_OFF = K2_GM_OFF + K2_GM_BYTES;        // per-row admission threshold
constexpr size_t K2_BAR_OFF = K2_THR_OFF + K2_ROWS * 4;
constexpr size_t K2_SMEM = K2_BAR_OFF + 128;
static_assert((size_t)(K2_CAPMAX + K2_SLACK) * 256 * 2 <= K2_GM_BYTES, "candidate lists alias the group maxima");

extern __shared__ __align__(1024) unsigned char k2_smem[];

// K-major tile of 32-byte rows, 32B swizzle: 8-row groups 256 B apart
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// 32 lanes x 64 consecutive fp32 accumulator columns -> registers
__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, float (&v)[64]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct K2Args {
  const float* s;      // [B][Npad] oracle norms of x (0 on padding)
  const float* n;      // [B][Npad] norms of the centred points (K2_PAD_NORM on padding)
  const float* smax;   // [B][2]: max s, max n over the cloud's real points
  uint16_t* cand;      // [B*N][2][cap]
  uint8_t* ccnt;       // [B*N][2]   (255 = overflow)
  int32_t* flags;      // [B*N]      (zeroed by the host wrapper)
  int N, Npad, C, k, cap;
  int wg;              // columns per group (8, 16, 32 or 64)
  int tiles_per_group; // > 1 only with wg == 64
  int fine_first;      // sweep 1 at full precision
};

// group maxima of this thread's 64 columns: GP = 64 / W groups
template <int W>
__device__ __forceinline__ void k2_group_max(const float (&v)[64], float (&m)[64 / W]) {
#pragma unroll
  for (int g = 0; g < 64 / W; ++g) {
    float a = v[g * W];
#pragma unroll
    for (int i = 1; i < W; ++i) a = fmaxf(a, v[g * W + i]);
    m[g] = a;
  }
}

__global__ void __launch_bounds__(K2_THREADS, 1)
    knn_tc_filter_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQ,
                         const K2Args A) {
  unsigned char* sm = k2_smem;
  float* gm = reinterpret_cast<float*>(sm + K2_GM_OFF);
  uint16_t* lists = reinterpret_cast<uint16_t*>(sm + K2_GM_OFF);
  float* thr = reinterpret_cast<float*>(sm + K2_THR_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + K2_BAR_OFF);
  uint64_t* a_full = bars;            // 1
  uint64_t* b_full = bars + 1;        // 2
  uint64_t* b_empty = bars + 3;       // 2
  uint64_t* acc_full = bars + 5;      // 2
  uint64_t* acc_empty = bars + 7;     // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * K2_ROWS;
  const int N = A.N, Npad = A.Npad;
  const int T = Npad / K2_COLS;
  const int grow0 = b * Npad;           // first row of this cloud in the padded operand arrays
  const int nks = (A.C + 15) >> 4;      // 16-channel k-slices that hold data
  const bool fine1 = A.fine_first != 0 || nks == 1;

  if (threadIdx.x == 0) {
    if (smem_u32(sm) & 1023u) __trap();  // 128B-swizzled operand tiles need a 1024-aligned base
    mbar_init(a_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);  // one arrival per scan warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the all-ones A tile of the norm k-slice (every element 1.0 => independent of the swizzle pattern)
  for (int i = threadIdx.x; i < (int)(K2_QTILE / 16); i += K2_THREADS)
    reinterpret_cast<uint4*>(sm + K2_ONES_OFF)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
      mbar_expect_tx(a_full, 2 * K2_TILE);
      tma_load_3d(sm + K2_A_OFF, &tmX, 0, grow0 + r0, 0, a_full);
      tma_load_3d(sm + K2_A_OFF + K2_TILE, &tmX, 0, grow0 + r0, 1, a_full);
      for (int step = 0; step < 2 * T; ++step) {
        const int st = step & 1;
        const int t = step < T ? step : step - T;
        const bool lo = step >= T || fine1;
        mbar_wait(&b_empty[st], ((step >> 1) & 1) ^ 1);
        mbar_expect_tx(&b_full[st], K2_TILE + K2_QTILE + (lo ? K2_TILE : 0u));
        unsigned char* dst = sm + K2_B_OFF + (size_t)st * K2_STAGE;
        tma_load_3d(dst, &tmX, 0, grow0 + t * K2_COLS, 0, &b_full[st]);
        if (lo) tma_load_3d(dst + K2_TILE, &tmX, 0, grow0 + t * K2_COLS, 1, &b_full[st]);
        tma_load_2d(dst + 2 * K2_TILE, &tmQ, 0, grow0 + t * K2_COLS, &b_full[st]);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(K2_COLS >> 3) << 17) |
                           ((uint32_t)(K2_ROWS >> 4) << 24);
    mbar_wait(a_full, 0);
    for (int step = 0; step < 2 * T; ++step) {
      const int st = step & 1;
      const uint32_t ph = (step >> 1) & 1;
      const bool fine = step >= T || fine1;
      mbar_wait(&b_full[st], ph);
      mbar_wait(&acc_empty[st], ph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(sm + K2_A_OFF), a_lo = a_hi + K2_TILE;
        const uint32_t b_hi = smem_u32(sm + K2_B_OFF + (size_t)st * K2_STAGE), b_lo = b_hi + K2_TILE;
        const uint32_t acc = tmem_base + (uint32_t)(st * K2_COLS);
        // -0.5 n_j first (it also clears the accumulator), then the small products, then hi.hi
        umma_bf16(acc, umma_desc_sw32(smem_u32(sm + K2_ONES_OFF)), umma_desc_sw32(b_hi + 2 * K2_TILE), idesc, 0);
        for (int ks = 0; ks < nks; ++ks) {
          const uint64_t dah = umma_desc(a_hi + ks * 32, 16, 1024), dbh = umma_desc(b_hi + ks * 32, 16, 1024);
          if (fine) {
            const uint64_t dal = umma_desc(a_lo + ks * 32, 16, 1024), dbl = umma_desc(b_lo + ks * 32, 16, 1024);
            umma_bf16(acc, dal, dbh, idesc, 1);
            umma_bf16(acc, dah, dbl, idesc, 1);
          }
          umma_bf16(acc, dah, dbh, idesc, 1);
        }
        umma_commit(&b_empty[st]);
        umma_commit(&acc_full[st]);
      }
      __syncwarp();
    }
  } else {
    // ---------------- scan warps: thread <-> (query row = TMEM lane, column half) ----------------
    const int sub = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rowl = sub * 32 + lane;            // row inside the CTA
    const int e = half * K2_ROWS + rowl;         // 0..255 among scan threads
    const int row = r0 + rowl;                   // row inside the cloud
    const uint32_t tlane = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(half * 64);
    const int gpt = 64 / A.wg;                   // groups per thread per tile
    const int tpg = A.tiles_per_group;
    const int G = 2 * gpt * ((T + tpg - 1) / tpg);
    float v[64];

    // ---- sweep 1: group maxima ----
    float run = -__int_as_float(0x7f800000);
    for (int t = 0; t < T; ++t) {
      const int st = t & 1;
      mbar_wait(&acc_full[st], (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tmem_ld_32x64(tlane + (uint32_t)(st * K2_COLS), v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[st]);   // the tile is in registers: the MMA warp may overwrite it
      float* g0 = gm + (size_t)((t / tpg) * 2 + half) * gpt * K2_ROWS + rowl;
      if (A.wg == 8) {
        float m[8];
        k2_group_max<8>(v, m);
#pragma unroll
        for (int g = 0; g < 8; ++g) g0[g * K2_ROWS] = m[g];
      } else if (A.wg == 16) {
        float m[4];
        k2_group_max<16>(v, m);
#pragma unroll
        for (int g = 0; g < 4; ++g) g0[g * K2_ROWS] = m[g];
      } else if (A.wg == 32) {
        float m[2];
        k2_group_max<32>(v, m);
        g0[0] = m[0];
        g0[K2_ROWS] = m[1];
      } else {
        float m[1];
        k2_group_max<64>(v, m);
        run = fmaxf(run, m[0]);
        if ((t + 1) % tpg == 0 || t == T - 1) {
          g0[0] = run;
          run = -__int_as_float(0x7f800000);
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ---- k-th largest group maximum -> admission threshold (one thread per row) ----
    if (half == 0) {
      const float sN = A.smax[2 * b], nN = A.smax[2 * b + 1];
      const bool valid = row < N;
      const float si = valid ? A.s[(size_t)b * Npad + row] : 0.0f;
      const float ni = valid ? A.n[(size_t)b * Npad + row] : 0.0f;
      const float eo = K2_EPS_ORACLE * (si + sN);
      const float e1 = (fine1 ? K2_EPS_FINE : K2_EPS_COARSE) * (ni + nN) + eo;
      const float e2 = K2_EPS_FINE * (ni + nN) + eo;
      float lo = __int_as_float(0x7f800000), hi = -lo;
      int nreal = 0;
      for (int g = 0; g < G; ++g) {
        const float m = gm[g * K2_ROWS + rowl];
        if (m > -1.0e29f) {   // groups made of padding columns only never count
          lo = fminf(lo, m);
          hi = fmaxf(hi, m);
          ++nreal;
        }
      }
      float tval = -__int_as_float(0x7f800000);   // admit everything
      if (nreal >= A.k) {
        // invariant: at least k group maxima are >= lo
        for (int it = 0; it < K2_BISECT; ++it) {
          const float mid = 0.5f * (lo + hi);
          if (!(mid > lo && mid < hi)) break;
          int c = 0;
          for (int g = 0; g < G; ++g) c += (gm[g * K2_ROWS + rowl] >= mid) ? 1 : 0;
          if (c >= A.k) lo = mid; else hi = mid;
        }
        tval = lo - 0.5f * (e1 + e2);
        tval -= (fabsf(lo) + e1 + e2) * (1.0f / 1048576.0f);   // fp32 evaluation slack, always towards admitting more
      }
      thr[rowl] = valid ? tval : __int_as_float(0x7f800000);   // rows beyond N admit nothing
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // thresholds visible; group maxima dead (lists alias them)

    // ---- sweep 2: collect the columns that clear the threshold ----
    const float tv = thr[rowl];
    const int cap = A.cap;
    uint16_t* mylist = lists + e;                 // slot s at mylist[s * 256]
    int cnt = 0;
    for (int t = 0; t < T; ++t) {
      const int step = T + t;
      const int st = step & 1;
      mbar_wait(&acc_full[st], (step >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tmem_ld_32x64(tlane + (uint32_t)(st * K2_COLS), v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[st]);
      const int col0 = t * K2_COLS + half * 64;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint16_t* w0 = mylist + min(cnt, cap) * 256;
        uint16_t* w = w0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (v[c8 * 8 + i] >= tv) {
            *w = (uint16_t)(col0 + c8 * 8 + i);
            w += 256;
          }
        }
        cnt += (int)(w - w0) >> 8;
      }
    }
    // ---- emit ----
    if (row < N) {
      const size_t rec = ((size_t)b * N + row) * 2 + half;
      const bool over = cnt > cap;
      A.ccnt[rec] = over ? (uint8_t)255 : (uint8_t)cnt;
      if (over) A.flags[(size_t)b * N + row] = 1;
      uint16_t* o = A.cand + rec * cap;
      for (int s8 = 0; s8 < cap; s8 += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          pk[i] = (uint32_t)mylist[(s8 + 2 * i) * 256] | ((uint32_t)mylist[(s8 + 2 * i + 1) * 256] << 16);
        *reinterpret_cast<uint4*>(o + s8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// exact fp32 distances of the candidates, sort by (distance, index), write the first k.  One warp per row; the
// candidate rows are staged through shared memory (coalesced 16-byte cp.async) so that each lane can run the oracle's
// sequential fmaf chain over ITS candidate without 32-way scattered global loads.
template <int KS>
__global__ void __launch_bounds__(256)
    knn_tc_refine_kernel(const float* __restrict__ x, const float* __restrict__ s, const uint16_t* __restrict__ cand,
                         const uint8_t* __restrict__ ccnt, int N, int Npad, int C, int k, int cap, int64_t P,
                         int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) float rf_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (p >= P) return;
  const int n0 = ccnt[2 * p], n1 = ccnt[2 * p + 1];
  if (n0 == 255 || n1 == 255) return;   // overflowed: knn_row_fallback_kernel owns this row
  const int nc = n0 + n1;
  const int b = (int)(p / N);
  const int row = (int)(p - (int64_t)b * N);
  const float* sb = s + (size_t)b * Npad;
  const float* xb = x + (int64_t)b * N * C;
  const float* xi = x + p * C;
  const float si = sb[row];
  const bool staged = (C & 3) == 0 && C >= 16;
  const int pitch = C + 4;
  float* stg = rf_smem + (size_t)warp * 33 * pitch;
  const int cpr = C >> 2;   // 16-byte chunks per row
  if (staged) {
    for (int c4 = lane; c4 < cpr; c4 += 32) cp_async16(stg + 32 * pitch + c4 * 4, xi + c4 * 4);
  }
  RowSel<KS> R;
  R.init();
  for (int base = 0; base < nc; base += 32) {
    const int e = base + lane;
    int j = -1;
    if (e < nc) j = e < n0 ? cand[(2 * p) * cap + e] : cand[(2 * p + 1) * cap + (e - n0)];
    if (j >= N) j = -1;
    float d = __int_as_float(0x7f800000);
    if (staged) {
      const int total = 32 * cpr;
      for (int q = lane; q < total; q += 32) {
        const int r = q / cpr, c4 = q - r * cpr;
        const int jr = __shfl_sync(FULL, j, r);
        if (jr >= 0) cp_async16(stg + r * pitch + c4 * 4, xb + (int64_t)jr * C + c4 * 4);
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      if (j >= 0) {
        float acc = 0.0f;
        const float* a = stg + 32 * pitch;
        const float* bb = stg + lane * pitch;
        for (int c = 0; c < C; c += 4) {
          const float4 a4 = *reinterpret_cast<const float4*>(a + c);
          const float4 b4 = *reinterpret_cast<const float4*>(bb + c);
          acc = __fmaf_rn(a4.x, b4.x, acc);
          acc = __fmaf_rn(a4.y, b4.y, acc);
          acc = __fmaf_rn(a4.z, b4.z, acc);
          acc = __fmaf_rn(a4.w, b4.w, acc);
        }
        d = __fadd_rn(__fsub_rn(__fadd_rn(si, sb[j]), __fmul_rn(2.0f, acc)), 0.0f);
      }
      __syncwarp();
    } else if (j >= 0) {
      const float* xj = xb + (int64_t)j * C;
      float acc = 0.0f;
      for (int c = 0; c < C; ++c) acc = __fmaf_rn(__ldg(xi + c), __ldg(xj + c), acc);
      d = __fadd_rn(__fsub_rn(__fadd_rn(si, sb[j]), __fmul_rn(2.0f, acc)), 0.0f);
    }
    R.merge_batch(d, j >= 0 ? j : 0x7fffffff, k, lane);
  }
  int32_t* o = idx + p * k;
#pragma unroll
  for (int q = 0; q < KS; ++q) {
    const int pos = q * 32 + lane;
    if (pos < k) o[pos] = R.j[q];
  }
}

// Flagged rows (candidate list overflow: distance ties beyond the filter's resolution) are recomputed exactly, one
// warp per row: all N distances by the oracle's fmaf chain (lanes over columns), selection as in topk_rows_kernel.
template <int KS>
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
  RowSel<KS> R;
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
    for (int c = 0; c < C; ++c) {       // four independent fmaf chains, each in the oracle's channel order
      const float a = __ldg(xi + c);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = __fmaf_rn(a, __ldg(xj[q] + c), acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dv[q] = __fadd_rn(__fsub_rn(__fadd_rn(si, sb[cj[q] < N ? cj[q] : N - 1]), __fmul_rn(2.0f, acc[q])), 0.0f);
    R.offer4(dv, cj, N, k, qd_s[warp], qj_s[warp], lane);
  }
  R.finish(k, qd_s[warp], qj_s[warp], lane);
  int32_t* o = idx + p * k;
#pragma unroll
  for (int q = 0; q < KS; ++q) {
    const int pos = q * 32 + lane;
    if (pos < k) o[pos] = R.j[q];
  }
}

// per-cloud channel means (the centring origin; any value is valid, it only tightens the error budget)
__global__ void __launch_bounds__(256) knn_tc_mean_kernel(const float* __restrict__ x, int N, int C, float* __restrict__ mean) {
  __shared__ float red[256];
  const int b = blockIdx.x;
  const float* xb = x + (size_t)b * N * C;
  // thread t accumulates channel t % C of the points t / C, t / C + 256 / C, ...   (256 / C point lanes; C <= 64)
  const int lanes = 256 / C;
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  float acc = 0.0f;
  if (pl < lanes)
    for (int n = pl; n < N; n += lanes) acc += xb[(size_t)n * C + c];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.0f;
    for (int q = 0; q < lanes; ++q) t += red[q * C + threadIdx.x];
    mean[(size_t)b * C + threadIdx.x] = t / (float)N;
  }
}

// x [B,N,C] -> s (oracle norms, ops.py:14: square rounded, then summed sequentially), n (norms of the centred points),
// bf16 planes [2][B*Npad][Cp] of the centred points (zero padded rows / channels), q [B*Npad][16] = bf16 parts of -0.5 n.
__global__ void __launch_bounds__(128)
    knn_tc_prep_kernel(const float* __restrict__ x, const float* __restrict__ mean, int N, int Npad, int C, int Cp,
                       float* __restrict__ s, float* __restrict__ nrm, __nv_bfloat16* __restrict__ planes,
                       __nv_bfloat16* __restrict__ q, int64_t plane_elems) {
  extern __shared__ float pt[];   // [128][C + 1]
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * 128;
  const int tid = threadIdx.x;
  const int pitch = C + 1;
  const int nvalid = min(128, N - n0);   // may be <= 0 for pure padding tiles
  const float* src = x + ((size_t)b * N + n0) * C;
  for (int e = tid; e < 128 * C; e += 128) {
    const int r = e / C, c = e - r * C;
    pt[r * pitch + c] = r < nvalid ? src[e] : 0.0f;
  }
  __syncthreads();
  {
    const float* mu = mean + (size_t)b * C;
    float so = 0.0f, sc = 0.0f;
    const bool valid = tid < nvalid;
    for (int c = 0; c < C; ++c) {
      const float v = pt[tid * pitch + c];
      so = __fadd_rn(so, __fmul_rn(v, v));
      const float y = valid ? __fsub_rn(v, mu[c]) : 0.0f;
      pt[tid * pitch + c] = y;
      sc = __fmaf_rn(y, y, sc);
    }
    const size_t g = (size_t)b * Npad + n0 + tid;
    s[g] = valid ? so : 0.0f;
    const float nn = valid ? sc : K2_PAD_NORM;
    nrm[g] = nn;
    const float h = -0.5f * nn;
    const __nv_bfloat16 q1 = __float2bfloat16_rn(h);
    const float r1 = h - __bfloat162float(q1);
    const __nv_bfloat16 q2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 q3 = __float2bfloat16_rn(r1 - __bfloat162float(q2));
    uint4 w0 = make_uint4(0u, 0u, 0u, 0u);
    w0.x = (uint32_t)__bfloat16_as_ushort(q1) | ((uint32_t)__bfloat16_as_ushort(q2) << 16);
    w0.y = (uint32_t)__bfloat16_as_ushort(q3);
    uint4* qo = reinterpret_cast<uint4*>(q + g * 16);
    qo[0] = w0;
    qo[1] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int hp = Cp >> 1;  // bf16 pairs per row
  __nv_bfloat16* ph = planes + ((size_t)b * Npad + n0) * Cp;
  __nv_bfloat16* pl = ph + plane_elems;
  for (int e = tid; e < 128 * hp; e += 128) {
    const int r = e / hp, c = (e - r * hp) * 2;
    const float v0 = c < C ? pt[r * pitch + c] : 0.0f;
    const float v1 = c + 1 < C ? pt[r * pitch + c + 1] : 0.0f;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
    *reinterpret_cast<uint32_t*>(ph + (size_t)r * Cp + c) =
        (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    *reinterpret_cast<uint32_t*>(pl + (size_t)r * Cp + c) =
        (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
}

// per cloud: max oracle norm and max centred norm over the real points
__global__ void __launch_bounds__(256)
    knn_tc_cloud_max_kernel(const float* __restrict__ s, const float* __restrict__ nrm, int N, int Npad,
                            float* __restrict__ smax) {
  __shared__ float red[2][8];
  const int b = blockIdx.x;
  float m0 = 0.0f, m1 = 0.0f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    m0 = fmaxf(m0, s[(size_t)b * Npad + n]);
    m1 = fmaxf(m1, nrm[(size_t)b * Npad + n]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 = fmaxf(m0, __shfl_xor_sync(FULL, m0, o));
    m1 = fmaxf(m1, __shfl_xor_sync(FULL, m1, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = m0;
    red[1][threadIdx.x >> 5] = m1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      m0 = fmaxf(m0, red[0][w]);
      m1 = fmaxf(m1, red[1][w]);
    }
    smax[2 * b] = m0;
    smax[2 * b + 1] = m1;
  }
}

static int make_q_map(CUtensorMap* tm, const void* q, int64_t rows) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return set_err(DGCNN_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {16, (cuuint64_t)rows};
  cuuint64_t strides[1] = {32};
  cuuint32_t box[2] = {16, (cuuint32_t)K2_COLS};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(q), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(DGCNN_ERR_CUDA, "tensor map (q): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DGCNN_OK;
}

static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }
static inline int k2_cp(int C) { return ((C + 7) / 8) * 8; }
static inline int k2_cap(int k) { return k <= 24 ? 32 : 64; }

bool knn_tc_eligible(int B, int N, int C, int k) {
  return C >= 1 && C <= 64 && k <= 48 && N >= 256 && N <= 65536 && (int64_t)B * (((N + 127) / 128) * 128) < (1ll << 31);
}

// scratch layout: s | n | mean | smax | planes | q | cand | ccnt | flags
size_t knn_tc_bytes(int B, int N, int C, int k_max) {
  const size_t Npad = ((size_t)N + 127) / 128 * 128;
  const size_t P = (size_t)B * N, Pp = (size_t)B * Npad;
  return 2 * al256(Pp * 4) + al256((size_t)B * C * 4) + al256((size_t)B * 8) + al256(2 * Pp * k2_cp(C) * 2) +
         al256(Pp * 32) + al256(P * 2 * k2_cap(k_max) * 2) + al256(P * 2) + al256(P * 4);
}

int knn_tc_run(const float* x, int32_t* idx, int B, int N, int C, int k, void* ws, cudaStream_t st) {
  const int Npad = ((N + 127) / 128) * 128;
  const int Cp = k2_cp(C);
  const int cap = k2_cap(k);
  const int64_t P = (int64_t)B * N, Pp = (int64_t)B * Npad;
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  size_t off = 0;
  float* s = reinterpret_cast<float*>(base + off); off += al256((size_t)Pp * 4);
  float* nrm = reinterpret_cast<float*>(base + off); off += al256((size_t)Pp * 4);
  float* mean = reinterpret_cast<float*>(base + off); off += al256((size_t)B * C * 4);
  float* smax = reinterpret_cast<float*>(base + off); off += al256((size_t)B * 8);
  __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256((size_t)2 * Pp * Cp * 2);
  __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256((size_t)Pp * 32);
  uint16_t* cand = reinterpret_cast<uint16_t*>(base + off); off += al256((size_t)P * 2 * cap * 2);
  uint8_t* ccnt = reinterpret_cast<uint8_t*>(base + off); off += al256((size_t)P * 2);
  int32_t* flags = reinterpret_cast<int32_t*>(base + off);

  knn_tc_mean_kernel<<<B, 256, 0, st>>>(x, N, C, mean);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tc_mean_kernel");
  dim3 gp(Npad / 128, B);
  knn_tc_prep_kernel<<<gp, 128, (size_t)128 * (C + 1) * 4, st>>>(x, mean, N, Npad, C, Cp, s, nrm, planes, q, Pp * Cp);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tc_prep_kernel");
  knn_tc_cloud_max_kernel<<<B, 256, 0, st>>>(s, nrm, N, Npad, smax);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tc_cloud_max_kernel");

  CUtensorMap tmX, tmQ;
  int rc = make_plane_map(&tmX, planes, Pp, Cp, 128);
  if (rc) return rc;
  rc = make_q_map(&tmQ, q, Pp);
  if (rc) return rc;
  static bool attr_done = false;
  static int fine_first = 0;
  if (!attr_done) {
    cudaFuncSetAttribute(knn_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K2_SMEM);
    cudaFuncSetAttribute(knn_tc_refine_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 33 * 68 * 4);
    cudaFuncSetAttribute(knn_tc_refine_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 33 * 68 * 4);
    const char* e = getenv("DGCNN_KNN_FINE_FIRST");   // tuning aid: sweep 1 at full precision
    fine_first = e ? atoi(e) : 0;
    attr_done = true;
  }
  if (cudaMemsetAsync(flags, 0, (size_t)P * sizeof(int32_t), st) != cudaSuccess)
    return set_err(DGCNN_ERR_CUDA, "knn_tc: memset failed");
  K2Args a;
  a.s = s; a.n = nrm; a.smax = smax; a.cand = cand; a.ccnt = ccnt; a.flags = flags;
  a.N = N; a.Npad = Npad; a.C = C; a.k = k; a.cap = cap;
  // group width: as many groups as fit (<= 128 per row), at least 8 columns each
  const int T = Npad / K2_COLS;
  int wg = 8;
  while (wg < 64 && Npad / wg > K2_GMAX) wg *= 2;
  a.wg = wg;
  a.tiles_per_group = 1;
  if (wg == 64) a.tiles_per_group = (2 * T + K2_GMAX - 1) / K2_GMAX;
  a.fine_first = fine_first;
  dim3 grid(Npad / K2_ROWS, B);
  knn_tc_filter_kernel<<<grid, K2_THREADS, K2_SMEM, st>>>(tmX, tmQ, a);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tc_filter_kernel");
  const bool staged = (C & 3) == 0 && C >= 16;
  const size_t rsm = staged ? (size_t)8 * 33 * (C + 4) * 4 : 0;
  if (k <= 32) {
    knn_tc_refine_kernel<1><<<cdiv(P, 8), 256, rsm, st>>>(x, s, cand, ccnt, N, Npad, C, k, cap, P, idx);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("knn_tc_refine_kernel");
    knn_row_fallback_kernel<1><<<cdiv(P, 8), 256, 0, st>>>(x, s, flags, N, Npad, C, k, P, idx);
  } else {
    knn_tc_refine_kernel<2><<<cdiv(P, 8), 256, rsm, st>>>(x, s, cand, ccnt, N, Npad, C, k, cap, P, idx);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("knn_tc_refine_kernel");
    knn_row_fallback_kernel<2><<<cdiv(P, 8), 256, 0, st>>>(x, s, flags, N, Npad, C, k, P, idx);
  }
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_row_fallback_kernel");
  return DGCNN_OK;
}

}  // namespace dgcnn
