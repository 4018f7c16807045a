// k_nn on the 5th-gen tensor cores, bit-exact.  /root/reference/dgcnn/ops.py:8-19 (tf.matmul + top_k) for clouds
// with C <= 64 channels: every EdgeConv layer of the model (ops.py:91-96), the xyz layer included.
//
// The [B,N,N] matrix never exists.  One CTA owns 128 query rows of one cloud (= the 128 TMEM lanes) and sweeps the
// cloud's columns TWICE in 128-column tiles; each sweep recomputes the tile on the tensor cores (the MMA of a tile
// is far cheaper than any way of keeping 128 x N distances on chip):
//
//   operands   y = (x - origin) * 2^e per cloud (distances are translation invariant and scale by 4^e; the origin is
//              the mid-range point, e puts max|y| in [2^13, 2^14)), rounded ONCE to fp16: h = fp16(y).  One
//              tcgen05.mma kind::f16 pass gives g_ij = h_i.h_j with |g - y_i.y_j| <= 2^-10 |y_i||y_j| (11-bit
//              significands, exact products, fp32 accumulation in TMEM).
//              Per column, the three bf16 parts of -0.5*c_j ride in an extra 16-channel k-slice against a tile of
//              ones, so the accumulator holds  a_ij = g_ij - 0.5 c_j  directly: the nearest columns of a row are its
//              LARGEST accumulator entries, and the scan needs no per-element arithmetic.  c_j = n_j + sig_j in
//              sweep 1 and n_j - sig_j in sweep 2 (n = |y|^2, sig = the point's share of the error budget), i.e.
//              sweep 1 sees UPPER bounds  n_i + sig_i - 2 a1_ij  of the exact distance and sweep 2 LOWER bounds
//              n_i - sig_i - 2 a2_ij; the column-dependent part of the budget costs nothing.
//   sweep 1    each scan thread keeps the maximum of every group of W columns -> G <= 128 group maxima per row in
//              shared memory.  The k-th largest group maximum a_k (bisection) certifies k DISTINCT columns whose exact
//              distance is <= U_i = n_i + sig_i - 2 a_k: an upper bound on the row's k-th smallest distance that is
//              only a few ranks loose (k-th largest of G random groups ~ rank G ln(G/(G-k)) of the row).
//   sweep 2    a column is a candidate iff its lower bound clears U_i:  a2_ij >= a_k - sig_i.  One compare per
//              element; the index of a survivor goes to the thread's list (<= cap per (row, column quarter)); a row
//              whose list overflows (massive ties, e.g. duplicate points) is appended to the fallback queue.
//   refine     (knn_tc_refine_kernel) exact fp32 distances of the candidates by the oracle's fmaf chain, warp sort
//              by (distance, index), first k written.  The candidates are a superset of every column whose exact
//              distance is <= the exact k-th smallest, so the result equals the oracle bit for bit, tie order included.
//   fallback   queued rows are recomputed exactly, one warp per row (knn_row_fallback_kernel).
//
// Error budget, in the scaled units (C <= 64; S = 4^e * oracle norm of x, n = fp32 norm of y):
//   oracle vs real distance : fmaf chain 2^-18 |x_i||x_j|, the two sequential norm sums 2^-18 (s_i+s_j), final
//                             roundings  ->  |D_oracle - D_real| < 2^-17 (S_i+S_j)             ; budget 2^-16 S
//   centring                : y = fl(x - origin) moves D_real by < 2^-22 (n_i+n_j)
//   fp16 product            : |h - y| <= max(2^-11 |y|, 2^-25)  ->  2 |g - y_i.y_j| <= 2^-10 (n_i+n_j) (1 + 2^-11)
//                             + 2^-24 sqrt(C) (sqrt n_i + sqrt n_j)   (the absolute term: fp16 subnormals)
//   accumulation            : <= 80 fp32 adds (possibly truncating) of exact products, |sum| <= (n_i+n_j)  -> 2^-16
//   norms                   : n itself 2^-18 n; the bf16 parts of c are exact to 2^-24 c
//   => coarse mode: sig = 1.125 * 2^-10 n + 2^-16 S + 2^-21 sqrt(n): the dominant fp16 term is a theorem, the 12.5 %
//      on top of it is ~6x the sum of the remaining terms.
//   fine mode (clouds of more than 4096 points): operands h = hi + lo (two fp16 roundings, |y - hi - lo| <= 2^-22 |y|) and three MMAs per
//      k-slice (lo.hi + hi.lo + hi.hi): the product term drops to 2^-20 (n_i+n_j) and the accumulation dominates:
//      <= 13*16 adds, each off by <= 2^-23 of a partial sum that never exceeds n_i/2 + n_j  ->  2^-14.3 in D units
//      if EVERY add truncated in the same direction.  Budget 2^-13 n per point (2.5x that worst case).  Dense clouds
//      (N = 16384: neighbour distances ~1 % of the norms) need this mode; it costs tensor time only, and the sweeps
//      are bound by reading the accumulators out of TMEM (64 B/clk/SM), not by the MMAs.
#include <stdlib.h>

#include <cuda_fp16.h>

#include "knn_select.cuh"
#include "tc_common.cuh"

namespace dgcnn {

constexpr int K2_ROWS = 128;        // query rows per CTA = TMEM lanes
constexpr int K2_COLS = 128;        // candidate columns per tile = accumulator columns
constexpr int K2_SCANW = 16;        // scan warps: 4 per TMEM sub-partition, one 32-column quarter of the tile each
constexpr int K2_SCAN = 32 * K2_SCANW;
constexpr int K2_THREADS = 64 + K2_SCAN;  // warp 0 TMA, warp 1 MMA
constexpr int K2_GMAX = 128;        // group maxima per row (32 per scan thread)
constexpr int K2_CAPMAX = 48;       // list entries per (row, column quarter)
constexpr int K2_SLACK = 8;         // writes past the cap land here (one clamp per 8 columns)
constexpr int K2_BISECT = 10;
constexpr int K2_NST = 3;           // B-tile ring depth (TMA latency >> MMA time of a tile)
constexpr int K2_NACC = 4;          // TMEM accumulator ring: 4 x 128 columns
constexpr float K2_EPS = 1.125f / 1024.0f;          // one fp16 product per pair
constexpr float K2_EPS_FINE = 1.0f / 8192.0f;        // fp16 hi/lo operands, three products per pair
constexpr float K2_EPS_ORACLE = 1.0f / 65536.0f;
constexpr float K2_EPS_ABS = 1.0f / 2097152.0f;
constexpr float K2_PAD_NORM = 1.0e36f;   // squared norm given to padding columns: they never pass a test
constexpr uint32_t K2_TILE = K2_ROWS * 64 * 2;   // one fp16 tile, 128B rows, 16 KB
constexpr uint32_t K2_QTILE = K2_ROWS * 16 * 2;  // the -0.5 c_j k-slice, 32B rows, 4 KB
constexpr uint32_t K2_STAGE = 2 * K2_TILE + K2_QTILE;   // B hi, B lo (fine mode only), norm slice

// byte offsets into the (1024-aligned) dynamic shared memory
constexpr size_t K2_A_OFF = 0;                                // A hi, A lo
constexpr size_t K2_ONES_OFF = K2_A_OFF + 2 * K2_TILE;        // 128 x 16 ones
constexpr size_t K2_B_OFF = K2_ONES_OFF + K2_QTILE;           // K2_NST stages of (B tile, q slice)
constexpr size_t K2_GM_OFF = K2_B_OFF + (size_t)K2_NST * K2_STAGE;   // group maxima (sweep 1) / candidate lists (sweep 2)
constexpr size_t K2_GM_BYTES = (size_t)K2_GMAX * K2_ROWS * 4;
constexpr size_t K2_PART_OFF = K2_GM_OFF + K2_GM_BYTES;       // bisection exchange: 2 x [4][128] ints
constexpr size_t K2_BAR_OFF = K2_PART_OFF + 2 * 4 * K2_ROWS * 4;
constexpr size_t K2_SMEM = K2_BAR_OFF + 256;
static_assert((size_t)(K2_CAPMAX + K2_SLACK) * K2_SCAN * 2 <= K2_GM_BYTES, "candidate lists alias the group maxima");
static_assert(K2_SMEM <= 227 * 1024, "shared memory budget");

extern __shared__ __align__(1024) unsigned char k2_smem[];

#ifdef K2_DEBUG
__device__ long long k2_dbg[512 * 16];
__device__ __forceinline__ long long k2_now() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define K2_STAMP(slot) do { k2_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = k2_now(); } while (0)
#else
#define K2_STAMP(slot) do { } while (0)
#endif

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

// 32 lanes x 32 consecutive fp32 accumulator columns -> registers
__device__ __forceinline__ void tmem_ld_f32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct K2Args {
  const float* sig;    // [B][Npad] the row's share of the error budget (scaled units)
  uint16_t* cand;      // [B*N][4][cap]
  uint8_t* ccnt;       // [B*N][4]   (255 = overflow)
  int32_t* fbq;        // fallback queue: fbq[0] = count, fbq[1..] = global row ids
  int N, Npad, C, k, cap;
  int wg;              // columns per group (8, 16 or 32)
  int tiles_per_group; // > 1 only with wg == 32
  int fine;            // operands split hi + lo, three MMAs per k-slice (error budget 2^-15 instead of 2^-10)
  int dbg;             // K2_DEBUG experiments
};

__global__ void __launch_bounds__(K2_THREADS, 1)
    knn_tc_filter_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQ,
                         const K2Args A) {
  unsigned char* sm = k2_smem;
  float* gm = reinterpret_cast<float*>(sm + K2_GM_OFF);
  uint16_t* lists = reinterpret_cast<uint16_t*>(sm + K2_GM_OFF);
  int* part = reinterpret_cast<int*>(sm + K2_PART_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + K2_BAR_OFF);
  uint64_t* a_full = bars;                        // 1
  uint64_t* b_full = bars + 1;                    // K2_NST
  uint64_t* b_empty = b_full + K2_NST;            // K2_NST
  uint64_t* acc_full = b_empty + K2_NST;          // K2_NACC
  uint64_t* acc_empty = acc_full + K2_NACC;       // K2_NACC
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + K2_NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * K2_ROWS;
  const int N = A.N, Npad = A.Npad;
  const int T = Npad / K2_COLS;
  const int grow0 = b * Npad;           // first row of this cloud in the padded operand arrays
  const int nks = (A.C + 15) >> 4;      // 16-channel k-slices that hold data
  if (threadIdx.x == 64) K2_STAMP(0);

  if (threadIdx.x == 0) {
    if (smem_u32(sm) & 1023u) __trap();  // 128B-swizzled operand tiles need a 1024-aligned base
    mbar_init(a_full, 1);
    for (int i = 0; i < K2_NST; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < K2_NACC; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], K2_SCANW);  // one arrival per scan warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the all-ones A tile of the norm k-slice (bf16 1.0 everywhere => independent of the swizzle pattern)
  for (int i = threadIdx.x; i < (int)(K2_QTILE / 16); i += K2_THREADS)
    reinterpret_cast<uint4*>(sm + K2_ONES_OFF)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {  // all 512 columns: four 128-column fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 64) K2_STAMP(1);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(a_full, A.fine ? 2 * K2_TILE : K2_TILE);
      tma_load_3d(sm + K2_A_OFF, &tmX, 0, grow0 + r0, 0, a_full);
      if (A.fine) tma_load_3d(sm + K2_A_OFF + K2_TILE, &tmX, 0, grow0 + r0, 1, a_full);
      for (int step = 0; step < 2 * T; ++step) {
        const int st = step % K2_NST;
        const int sweep = step >= T;
        const int t = sweep ? step - T : step;
        mbar_wait(&b_empty[st], ((step / K2_NST) & 1) ^ 1);
        unsigned char* dst = sm + K2_B_OFF + (size_t)st * K2_STAGE;
        mbar_expect_tx(&b_full[st], (A.fine ? 2 * K2_TILE : K2_TILE) + K2_QTILE);
        tma_load_3d(dst, &tmX, 0, grow0 + t * K2_COLS, 0, &b_full[st]);
        if (A.fine) tma_load_3d(dst + K2_TILE, &tmX, 0, grow0 + t * K2_COLS, 1, &b_full[st]);
        tma_load_3d(dst + 2 * K2_TILE, &tmQ, 0, grow0 + t * K2_COLS, sweep, &b_full[st]);
        if (step == T - 1) K2_STAMP(8);
      }
      K2_STAMP(9);
    }
  } else if (warp == 1) {
    // instruction descriptors: D = f32, K-major A and B, N = 128, M = 128; operands fp16 (data) / bf16 (norm slice)
    const uint32_t idesc_dim = (1u << 4) | ((uint32_t)(K2_COLS >> 3) << 17) | ((uint32_t)(K2_ROWS >> 4) << 24);
    const uint32_t idesc_f16 = idesc_dim;
    const uint32_t idesc_bf16 = idesc_dim | (1u << 7) | (1u << 10);
    mbar_wait(a_full, 0);
    if (lane == 0) K2_STAMP(10);
    for (int step = 0; step < 2 * T; ++step) {
      const int st = step % K2_NST;
      const int ab = step % K2_NACC;
      if (lane == 0 && step == T) K2_STAMP(11);
      if (lane == 0 && step == 1) K2_STAMP(13);
      mbar_wait(&b_full[st], (step / K2_NST) & 1);
      mbar_wait(&acc_empty[ab], ((step / K2_NACC) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_h = smem_u32(sm + K2_A_OFF), a_l = a_h + K2_TILE;
        const uint32_t b_h = smem_u32(sm + K2_B_OFF + (size_t)st * K2_STAGE), b_l = b_h + K2_TILE;
        const uint32_t acc = tmem_base + (uint32_t)(ab * K2_COLS);
        // -0.5 c_j first (it also clears the accumulator), then the products (small terms before hi.hi)
        umma_bf16(acc, umma_desc_sw32(smem_u32(sm + K2_ONES_OFF)), umma_desc_sw32(b_h + 2 * K2_TILE), idesc_bf16, 0);
        for (int ks = 0; ks < nks; ++ks) {
          const uint64_t dah = umma_desc(a_h + ks * 32, 16, 1024), dbh = umma_desc(b_h + ks * 32, 16, 1024);
          if (A.fine) {
            umma_bf16(acc, umma_desc(a_l + ks * 32, 16, 1024), dbh, idesc_f16, 1);
            umma_bf16(acc, dah, umma_desc(b_l + ks * 32, 16, 1024), idesc_f16, 1);
          }
          umma_bf16(acc, dah, dbh, idesc_f16, 1);
        }
        umma_commit(&b_empty[st]);
        umma_commit(&acc_full[ab]);
      }
      __syncwarp();
    }
    if (lane == 0) K2_STAMP(12);
  } else {
    // ---------------- scan warps: thread <-> (query row = TMEM lane, 32-column quarter of the tile) ----------------
    const int sub = warp & 3;                    // TMEM sub-partition of this warp
    const int qd = (warp - 2) >> 2;              // column quarter
    const int rowl = sub * 32 + lane;            // row inside the CTA
    const int e = qd * K2_ROWS + rowl;           // 0..511 among scan threads
    const int row = r0 + rowl;                   // row inside the cloud
    const uint32_t tlane = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(qd * 32);
    const int gpt = 32 / A.wg;                   // groups per thread per tile
    const int tpg = A.tiles_per_group;
    const int gq = gpt * ((T + tpg - 1) / tpg);  // groups owned by this thread (<= 32)
    const float ninf = -__int_as_float(0x7f800000);
    float v[32];

    // ---- sweep 1: group maxima; group li of this thread lives at gm[(qd*32 + li)*128 + rowl] ----
    float run = ninf;
    float* gbase = gm + (size_t)qd * 32 * K2_ROWS + rowl;
    for (int t = 0; t < T; ++t) {
      const int ab = t % K2_NACC;
      mbar_wait(&acc_full[ab], (t / K2_NACC) & 1);
      if (threadIdx.x == 64 && t == 0) K2_STAMP(2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tmem_ld_f32x32(tlane + (uint32_t)(ab * K2_COLS), v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);   // the tile is in registers: the MMA warp may overwrite it
      if (A.wg == 8) {
        float* g0 = gbase + (size_t)t * 4 * K2_ROWS;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float m = v[g * 8];
#pragma unroll
          for (int i = 1; i < 8; ++i) m = fmaxf(m, v[g * 8 + i]);
          g0[g * K2_ROWS] = m;
        }
      } else if (A.wg == 16) {
        float* g0 = gbase + (size_t)t * 2 * K2_ROWS;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float m = v[g * 16];
#pragma unroll
          for (int i = 1; i < 16; ++i) m = fmaxf(m, v[g * 16 + i]);
          g0[g * K2_ROWS] = m;
        }
      } else {
        float m = v[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) m = fmaxf(m, v[i]);
        run = fmaxf(run, m);
        if ((t + 1) % tpg == 0 || t == T - 1) {
          gbase[(size_t)(t / tpg) * K2_ROWS] = run;
          run = ninf;
        }
      }
    }
    if (threadIdx.x == 64) K2_STAMP(3);
    // ---- k-th largest group maximum of the row: 4 threads per row bisect together, each on its own <= 32 groups ----
    float mv[32];
    float lo = -ninf, hi = ninf;
    int nreal = 0;
#pragma unroll
    for (int g = 0; g < 32; ++g) {
      mv[g] = g < gq ? gbase[(size_t)g * K2_ROWS] : ninf;
      if (mv[g] > -1.0e35f) {   // groups made of padding columns only never count
        lo = fminf(lo, mv[g]);
        hi = fmaxf(hi, mv[g]);
        ++nreal;
      }
    }
    // the four threads of a row sit in the four warps of one TMEM sub-partition: they meet at named barrier 1 + sub
    float* partf = reinterpret_cast<float*>(part);
#define K2_ROWSYNC() asm volatile("bar.sync %0, 128;" ::"r"(1 + sub) : "memory")
    partf[e] = lo;
    partf[4 * K2_ROWS + e] = hi;
    K2_ROWSYNC();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      lo = fminf(lo, partf[q * K2_ROWS + rowl]);
      hi = fmaxf(hi, partf[4 * K2_ROWS + q * K2_ROWS + rowl]);
    }
    K2_ROWSYNC();
    part[4 * K2_ROWS + e] = nreal;   // second buffer: iteration 0 below writes the first one
    K2_ROWSYNC();
    nreal = part[4 * K2_ROWS + rowl] + part[5 * K2_ROWS + rowl] + part[6 * K2_ROWS + rowl] + part[7 * K2_ROWS + rowl];
    // invariant: at least k group maxima are >= lo.  All four threads of a row take identical decisions.
    for (int it = 0; it < K2_BISECT; ++it) {
      const float mid = 0.5f * lo + 0.5f * hi;
      int c = 0;
#pragma unroll
      for (int g = 0; g < 32; ++g)
        asm("{\n.reg .pred p;\nsetp.ge.f32 p, %1, %2;\n@p add.s32 %0, %0, 1;\n}" : "+r"(c) : "f"(mv[g]), "f"(mid));
      int* pb = part + (it & 1) * 4 * K2_ROWS;
      pb[e] = c;
      K2_ROWSYNC();
      c = pb[rowl] + pb[K2_ROWS + rowl] + pb[2 * K2_ROWS + rowl] + pb[3 * K2_ROWS + rowl];
      if (mid > lo && mid < hi) {
        if (c >= A.k) lo = mid; else hi = mid;
      }
    }
#undef K2_ROWSYNC
    float tv = ninf;   // admit everything (the row then overflows into the fallback queue)
    if (nreal >= A.k) {
      const float sg = A.sig[(size_t)grow0 + min(row, Npad - 1)];
      tv = lo - sg;
      tv -= (fabsf(lo) + sg) * (1.0f / 1048576.0f);   // fp32 evaluation slack, always towards admitting more
    }
    if (row >= N) tv = -ninf;                          // rows beyond N admit nothing
    asm volatile("bar.sync 5, 512;" ::: "memory");      // group maxima are dead: the lists alias them

    // ---- sweep 2: collect the columns that clear the threshold ----
    // Two instructions per element (compare + predicated bit set) build a 32-bit mask of the tile's survivors; the
    // rare survivors are then appended to the thread's list (shared memory, 32-bit addressing).
    if (threadIdx.x == 64) K2_STAMP(4);
    const int cap = A.cap;
    uint16_t* mylist = lists + e;                 // slot s at mylist[s * K2_SCAN]
    const uint32_t list_addr = smem_u32(mylist);
    int cnt = 0;
    for (int t = 0; t < T; ++t) {
      const int step = T + t;
      const int ab = step % K2_NACC;
      mbar_wait(&acc_full[ab], (step / K2_NACC) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tmem_ld_f32x32(tlane + (uint32_t)(ab * K2_COLS), v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      uint32_t m = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (v[i] >= tv) m |= 1u << i;
      const int col0 = t * K2_COLS + qd * 32;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t slot = (uint32_t)min(cnt, cap);    // overflow lands in the scratch slot `cap`
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr + slot * (uint32_t)(K2_SCAN * 2)), "h"((uint16_t)(col0 + i)) : "memory");
        ++cnt;
      }
    }
    if (threadIdx.x == 64) K2_STAMP(5);
    // ---- emit ----
    if (row < N) {
      const size_t rec = ((size_t)b * N + row) * 4 + qd;
      const bool over = cnt > cap;
      A.ccnt[rec] = over ? (uint8_t)255 : (uint8_t)cnt;
      if (over) {
        const int pos = atomicAdd(A.fbq, 1);
        A.fbq[1 + pos] = b * N + row;   // a row may be queued by several of its quarters: the fallback is idempotent
      }
      uint16_t* o = A.cand + rec * cap;
      for (int s8 = 0; s8 < cap; s8 += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          pk[i] = (uint32_t)mylist[(s8 + 2 * i) * K2_SCAN] | ((uint32_t)mylist[(s8 + 2 * i + 1) * K2_SCAN] << 16);
        *reinterpret_cast<uint4*>(o + s8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  if (threadIdx.x == 64) K2_STAMP(6);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (threadIdx.x == 64) K2_STAMP(7);
}

// exact fp32 distances of the candidates, sort by (distance, index), write the first k.  One warp per row; the
// candidate rows are staged through shared memory (coalesced 16-byte cp.async) so that each lane can run the oracle's
// sequential fmaf chain over ITS candidate without 32-way scattered global loads.
// CT > 0: the channel count as a compile-time constant (multiple of 16; the model's 64-channel layers).
template <int KS, int CT>
__global__ void __launch_bounds__(256)
    knn_tc_refine_kernel(const float* __restrict__ x, const float* __restrict__ s, const uint16_t* __restrict__ cand,
                         const uint8_t* __restrict__ ccnt, int N, int Npad, int C_rt, int k, int cap, int64_t P,
                         int32_t* __restrict__ idx) {
  const int C = CT > 0 ? CT : C_rt;
  extern __shared__ __align__(16) float rf_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (p >= P) return;
  const uchar4 cc = *reinterpret_cast<const uchar4*>(ccnt + 4 * p);
  if (cc.x == 255 || cc.y == 255 || cc.z == 255 || cc.w == 255) return;   // overflowed: the fallback owns this row
  const int e1 = cc.x, e2 = e1 + cc.y, e3 = e2 + cc.z, nc = e3 + cc.w;
  const int b = (int)(p / N);
  const int row = (int)(p - (int64_t)b * N);
  const float* sb = s + (size_t)b * Npad;
  const float* xb = x + (int64_t)b * N * C;
  const float* xi = x + p * C;
  const float si = sb[row];
  const bool staged = (C & 3) == 0 && C >= 16;
  // CT > 0: the candidate rows are staged in two halves of CT/2 channels ([32][CT/2 + 4] floats + the query row:
  // half the shared memory per warp = twice the resident warps; the fmaf chain simply continues across the halves)
  constexpr int HC = CT > 0 ? CT / 2 : 0;
  const int pitch = CT > 0 ? HC + 4 : C + 4;
  float* stg = CT > 0 ? rf_smem + (size_t)warp * (CT + 32 * (HC + 4)) + CT : rf_smem + (size_t)warp * 33 * pitch;
  float* xq = CT > 0 ? stg - CT : stg + 32 * pitch;   // the query row
  const int cpr = C >> 2;   // 16-byte chunks per row
  if (staged) {
    for (int c4 = lane; c4 < cpr; c4 += 32) cp_async16(xq + c4 * 4, xi + c4 * 4);
  }
  const uint16_t* cl = cand + (size_t)4 * p * cap;
  RowSel<KS> R;
  R.init();
  for (int base = 0; base < nc; base += 32) {
    const int e = base + lane;
    int j = -1;
    if (e < nc) {
      const int q = (e >= e1) + (e >= e2) + (e >= e3);
      const int st = q == 0 ? 0 : (q == 1 ? e1 : (q == 2 ? e2 : e3));
      j = cl[q * cap + (e - st)];
    }
    if (j >= N) j = -1;
    float d = __int_as_float(0x7f800000);
    if (staged) {
      if (CT > 0) {
        constexpr int CPH = HC >= 4 ? HC / 4 : 1;                      // 16-byte chunks per half row
        constexpr int RPP = 32 / CPH > 0 ? 32 / CPH : 1;               // rows per pass
        constexpr int LPR = 32 / RPP;                                  // lanes per row
        const int sub = lane / LPR, c4l = lane - sub * LPR;
        float acc = 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll 4
          for (int r = 0; r < 32; r += RPP) {
            const int jr = __shfl_sync(FULL, j, r + sub);
            if (jr >= 0) {
#pragma unroll
              for (int c4 = c4l; c4 < CPH; c4 += LPR)
                cp_async16(stg + (r + sub) * pitch + c4 * 4, xb + (int64_t)jr * CT + h * HC + c4 * 4);
            }
          }
          cp_async_commit();
          cp_async_wait<0>();
          __syncwarp();
          if (j >= 0) {
            const float* a = xq + h * HC;
            const float* bb = stg + lane * pitch;
#pragma unroll 4
            for (int c = 0; c < HC; c += 4) {
              const float4 a4 = *reinterpret_cast<const float4*>(a + c);
              const float4 b4 = *reinterpret_cast<const float4*>(bb + c);
              acc = __fmaf_rn(a4.x, b4.x, acc);
              acc = __fmaf_rn(a4.y, b4.y, acc);
              acc = __fmaf_rn(a4.z, b4.z, acc);
              acc = __fmaf_rn(a4.w, b4.w, acc);
            }
          }
          __syncwarp();
        }
        if (j >= 0) d = __fadd_rn(__fsub_rn(__fadd_rn(si, sb[j]), __fmul_rn(2.0f, acc)), 0.0f);
      } else {
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
          const float* a = xq;
          const float* bb = stg + lane * pitch;
#pragma unroll 4
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
      }
    } else if (j >= 0) {
      const float* xj = xb + (int64_t)j * C;
      float acc = 0.0f;
      for (int c = 0; c < C; ++c) acc = __fmaf_rn(__ldg(xi + c), __ldg(xj + c), acc);
      d = __fadd_rn(__fsub_rn(__fadd_rn(si, sb[j]), __fmul_rn(2.0f, acc)), 0.0f);
    }
    int jj = j >= 0 ? j : 0x7fffffff;
    if (base == 0) {      // the list is still empty: the sorted batch IS the list (saves the merge network)
      warp_sort32(d, jj, lane);
      R.d[0] = d;
      R.j[0] = jj;
    } else {
      R.merge_batch(d, jj, k, lane);
    }
  }
  if (staged) {
    cp_async_commit();
    cp_async_wait<0>();
  }
  int32_t* o = idx + p * k;
#pragma unroll
  for (int q = 0; q < KS; ++q) {
    const int pos = q * 32 + lane;
    if (pos < k) o[pos] = R.j[q];
  }
}

// Queued rows (candidate list overflow: distance ties beyond the filter's resolution) are recomputed exactly, one
// warp per row: all N distances by the oracle's fmaf chain (lanes over columns), selection as in topk_rows_kernel.
template <int KS>
__global__ void __launch_bounds__(256)
    knn_row_fallback_kernel(const float* __restrict__ x, const float* __restrict__ s, const int32_t* __restrict__ fbq,
                            int N, int Npad, int C, int k, int32_t* __restrict__ idx) {
  // one BLOCK per queued row: the 8 warps scan interleaved 128-column chunks, then warp 0 merges the 8 lists
  __shared__ float qd_s[8][QCAP];
  __shared__ int qj_s[8][QCAP];
  __shared__ float md_s[8][32 * KS];
  __shared__ int mj_s[8][32 * KS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nq = fbq[0];
  for (int it = blockIdx.x; it < nq; it += gridDim.x) {
    const int64_t p = fbq[1 + it];
    const int b = (int)(p / N);
    const float* sb = s + (size_t)b * Npad;
    const float* xb = x + (int64_t)b * N * C;
    const float* xi = x + p * C;
    const float si = sb[p - (int64_t)b * N];
    RowSel<KS> R;
    R.init();
    for (int c0 = warp * 128; c0 < N; c0 += 8 * 128) {
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
#pragma unroll
    for (int q = 0; q < KS; ++q) {
      md_s[warp][q * 32 + lane] = R.d[q];
      mj_s[warp][q * 32 + lane] = R.j[q];
    }
    __syncthreads();
    if (warp == 0) {
      for (int w = 1; w < 8; ++w) {
#pragma unroll
        for (int q = 0; q < KS; ++q) R.merge_batch(md_s[w][q * 32 + lane], mj_s[w][q * 32 + lane], k, lane);
      }
      int32_t* o = idx + p * k;
#pragma unroll
      for (int q = 0; q < KS; ++q) {
        const int pos = q * 32 + lane;
        if (pos < k) o[pos] = R.j[q];
      }
    }
    __syncthreads();
  }
}

// per-cloud, per-channel min / max of x, in K2_CHUNKS partial results (no atomics, no zeroing)
constexpr int K2_CHUNKS = 16;
__global__ void __launch_bounds__(256)
    knn_tc_range_kernel(const float* __restrict__ x, int N, int C, float* __restrict__ part, int32_t* __restrict__ fbq) {
  __shared__ float rmin[256], rmax[256];
  const int b = blockIdx.y, ch = blockIdx.x;
  if (b == 0 && ch == 0 && threadIdx.x == 0) fbq[0] = 0;   // the fallback queue of this call starts empty
  const int per = (N + K2_CHUNKS - 1) / K2_CHUNKS;
  const int n_lo = ch * per, n_hi = min(N, n_lo + per);
  const float* xb = x + (size_t)b * N * C;
  // thread t: channel t % C of the points n_lo + t / C, + 256 / C, ...   (C <= 64)
  const int lanes = 256 / C;
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  float mn = __int_as_float(0x7f800000), mx = -mn;
  if (pl < lanes)
    for (int n = n_lo + pl; n < n_hi; n += lanes) {
      const float v = xb[(size_t)n * C + c];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  rmin[threadIdx.x] = mn;
  rmax[threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x < C) {
    for (int q = 1; q < lanes; ++q) {
      mn = fminf(mn, rmin[q * C + threadIdx.x]);
      mx = fmaxf(mx, rmax[q * C + threadIdx.x]);
    }
    float* o = part + (((size_t)b * K2_CHUNKS + ch) * 2) * C;
    o[threadIdx.x] = mn;
    o[C + threadIdx.x] = mx;
  }
}

// x [B,N,C] -> s (oracle norms, ops.py:14: square rounded, then summed sequentially), sig (error budget share),
// fp16 operand h [B*Npad][Cp] of the centred + scaled points (zero padded rows / channels),
// q [2][B*Npad][16] = bf16 parts of -0.5 (n + sig) (sweep 1) and -0.5 (n - sig) (sweep 2).
__global__ void __launch_bounds__(128)
    knn_tc_prep_kernel(const float* __restrict__ x, const float* __restrict__ part, int N, int Npad, int C, int Cp,
                       float* __restrict__ s, float* __restrict__ sig, __half* __restrict__ h,
                       __nv_bfloat16* __restrict__ q, int64_t q_plane, int64_t h_plane, int fine) {
  extern __shared__ float pt[];   // [128][C + 1] | origin [C] | scale, scale^2
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * 128;
  const int tid = threadIdx.x;
  const int pitch = C + 1;
  float* org = pt + 128 * pitch;
  float* scl = org + C;
  __shared__ float ext_s[64];
  const int nvalid = min(128, N - n0);   // may be <= 0 for pure padding tiles
  const float* src = x + ((size_t)b * N + n0) * C;
  for (int e = tid; e < 128 * C; e += 128) {
    const int r = e / C, c = e - r * C;
    pt[r * pitch + c] = r < nvalid ? src[e] : 0.0f;
  }
  if (tid < C) {
    float mn = __int_as_float(0x7f800000), mx = -mn;
    for (int ch = 0; ch < K2_CHUNKS; ++ch) {
      const float* o = part + (((size_t)b * K2_CHUNKS + ch) * 2) * C;
      mn = fminf(mn, o[tid]);
      mx = fmaxf(mx, o[C + tid]);
    }
    const float mid = 0.5f * mn + 0.5f * mx;
    org[tid] = mid;
    ext_s[tid] = fmaxf(mx - mid, mid - mn);
  }
  __syncthreads();
  if (tid == 0) {
    float ext = 0.0f;
    for (int c = 0; c < C; ++c) ext = fmaxf(ext, ext_s[c]);
    int ex = 0;
    if (ext > 0.0f && ext < 3.0e38f) {
      int xe;
      frexpf(ext, &xe);          // ext = f * 2^xe, f in [0.5, 1)
      ex = 13 - xe;              // ext * 2^ex in [2^12, 2^13): |y| stays below 2^14 with room for the rounding of x - mid
      ex = max(-60, min(60, ex));
    }
    scl[0] = ldexpf(1.0f, ex);
    scl[1] = ldexpf(1.0f, 2 * ex);
  }
  __syncthreads();
  {
    const float sc = scl[0], sc2 = scl[1];
    float so = 0.0f, nn = 0.0f;
    const bool valid = tid < nvalid;
    for (int c = 0; c < C; ++c) {
      const float v = pt[tid * pitch + c];
      so = __fadd_rn(so, __fmul_rn(v, v));
      const float y = valid ? __fsub_rn(v, org[c]) * sc : 0.0f;
      pt[tid * pitch + c] = y;
      nn = __fmaf_rn(y, y, nn);
    }
    const size_t g = (size_t)b * Npad + n0 + tid;
    s[g] = valid ? so : 0.0f;
    const float sg = (fine ? K2_EPS_FINE : K2_EPS) * nn + K2_EPS_ORACLE * (so * sc2) + K2_EPS_ABS * sqrtf(nn);
    sig[g] = valid ? sg : 0.0f;
#pragma unroll
    for (int sw = 0; sw < 2; ++sw) {
      const float cj = valid ? (sw == 0 ? nn + sg : nn - sg) : K2_PAD_NORM;
      const float hv = -0.5f * cj;
      const __nv_bfloat16 q1 = __float2bfloat16_rn(hv);
      const float r1 = hv - __bfloat162float(q1);
      const __nv_bfloat16 q2 = __float2bfloat16_rn(r1);
      const __nv_bfloat16 q3 = __float2bfloat16_rn(r1 - __bfloat162float(q2));
      uint4 w0 = make_uint4(0u, 0u, 0u, 0u);
      w0.x = (uint32_t)__bfloat16_as_ushort(q1) | ((uint32_t)__bfloat16_as_ushort(q2) << 16);
      w0.y = (uint32_t)__bfloat16_as_ushort(q3);
      uint4* qo = reinterpret_cast<uint4*>(q + (size_t)sw * q_plane + g * 16);
      qo[0] = w0;
      qo[1] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();
  const int hp = Cp >> 1;  // fp16 pairs per row
  __half* ph = h + ((size_t)b * Npad + n0) * Cp;
  for (int e = tid; e < 128 * hp; e += 128) {
    const int r = e / hp, c = (e - r * hp) * 2;
    const float v0 = c < C ? pt[r * pitch + c] : 0.0f;
    const float v1 = c + 1 < C ? pt[r * pitch + c + 1] : 0.0f;
    const __half2 hh = __floats2half2_rn(v0, v1);
    *reinterpret_cast<__half2*>(ph + (size_t)r * Cp + c) = hh;
    if (fine) {   // second plane: what the first rounding left over
      const float2 hf = __half22float2(hh);
      *reinterpret_cast<__half2*>(ph + h_plane + (size_t)r * Cp + c) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    }
  }
}

static int make_h_map(CUtensorMap* tm, const void* h, int64_t rows, int64_t cols) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return set_err(DGCNN_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * (cuuint64_t)cols * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)K2_COLS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(h), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(DGCNN_ERR_CUDA, "tensor map (h): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DGCNN_OK;
}

static int make_q_map(CUtensorMap* tm, const void* q, int64_t rows) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return set_err(DGCNN_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {16, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {32, (cuuint64_t)rows * 32};
  cuuint32_t box[3] = {16, (cuuint32_t)K2_COLS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(q), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(DGCNN_ERR_CUDA, "tensor map (q): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DGCNN_OK;
}

static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }
static inline int k2_cp(int C) { return ((C + 7) / 8) * 8; }
static inline int k2_cap(int k) { return k <= 24 ? 32 : 48; }

bool knn_tc_eligible(int B, int N, int C, int k) {
  return C >= 1 && C <= 64 && k <= 48 && N >= 256 && N <= 65536 && (int64_t)B * (((N + 127) / 128) * 128) < (1ll << 31);
}

// scratch layout: s | sig | range partials | h | q | cand | ccnt | fallback queue
size_t knn_tc_bytes(int B, int N, int C, int k_max) {
  const size_t Npad = ((size_t)N + 127) / 128 * 128;
  const size_t P = (size_t)B * N, Pp = (size_t)B * Npad;
  return 2 * al256(Pp * 4) + al256((size_t)B * K2_CHUNKS * 2 * C * 4) + al256(2 * Pp * k2_cp(C) * 2) + al256(2 * Pp * 32) +
         al256(P * 4 * k2_cap(k_max) * 2) + al256(P * 4) + al256((4 * P + 1) * 4);
}

int knn_tc_run(const float* x, int32_t* idx, int B, int N, int C, int k, int filter_mode, void* ws, cudaStream_t st) {
  const int Npad = ((N + 127) / 128) * 128;
  const int Cp = k2_cp(C);
  const int cap = k2_cap(k);
  const int64_t P = (int64_t)B * N, Pp = (int64_t)B * Npad;
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  size_t off = 0;
  float* s = reinterpret_cast<float*>(base + off); off += al256((size_t)Pp * 4);
  float* sig = reinterpret_cast<float*>(base + off); off += al256((size_t)Pp * 4);
  float* part = reinterpret_cast<float*>(base + off); off += al256((size_t)B * K2_CHUNKS * 2 * C * 4);
  __half* h = reinterpret_cast<__half*>(base + off); off += al256((size_t)2 * Pp * Cp * 2);
  __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256((size_t)2 * Pp * 32);
  uint16_t* cand = reinterpret_cast<uint16_t*>(base + off); off += al256((size_t)P * 4 * cap * 2);
  uint8_t* ccnt = reinterpret_cast<uint8_t*>(base + off); off += al256((size_t)P * 4);
  int32_t* fbq = reinterpret_cast<int32_t*>(base + off);

  dim3 gr(K2_CHUNKS, B);
  knn_tc_range_kernel<<<gr, 256, 0, st>>>(x, N, C, part, fbq);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tc_range_kernel");
  dim3 gp(Npad / 128, B);
  // Precision mode of the filter.  Coarse (one fp16 product) is enough while the k-th neighbour distance is more than
  // a few percent of the squared norms; dense clouds (many points per cloud) need the hi/lo split or most rows would
  // overflow their candidate lists into the slow exact fallback.  The result is identical either way.
  // filter_mode (dgcnn_knn_mode's argument): DGCNN_KNN_AUTO applies the rule, COARSE / FINE force a mode (A/B timing).
  const int fine = filter_mode >= 0 ? (filter_mode != 0) : (N > 4096 ? 1 : 0);
  knn_tc_prep_kernel<<<gp, 128, (size_t)(128 * (C + 1) + C + 2) * 4, st>>>(x, part, N, Npad, C, Cp, s, sig, h, q, Pp * 16,
                                                                          Pp * Cp, fine);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tc_prep_kernel");

  CUtensorMap tmX, tmQ;
  int rc = make_h_map(&tmX, h, Pp, Cp);
  if (rc) return rc;
  rc = make_q_map(&tmQ, q, Pp);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(knn_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K2_SMEM);
    cudaFuncSetAttribute(knn_tc_refine_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 33 * 68 * 4);
    cudaFuncSetAttribute(knn_tc_refine_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 33 * 68 * 4);
    cudaFuncSetAttribute(knn_tc_refine_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 33 * 68 * 4);
    cudaFuncSetAttribute(knn_tc_refine_kernel<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 33 * 68 * 4);
    attr_done = true;
  }
  K2Args a;
  a.sig = sig; a.cand = cand; a.ccnt = ccnt; a.fbq = fbq;
  a.N = N; a.Npad = Npad; a.C = C; a.k = k; a.cap = cap;
  // group width: as many groups as fit (<= 128 per row = 32 per scan thread), at least 8 columns each
  const int T = Npad / K2_COLS;
  int wg = 8;
  while (wg < 32 && Npad / wg > K2_GMAX) wg *= 2;
  a.wg = wg;
  a.tiles_per_group = wg == 32 ? (T + 31) / 32 : 1;
  a.fine = fine;
  a.dbg = 0;
  dim3 grid(Npad / K2_ROWS, B);
  knn_tc_filter_kernel<<<grid, K2_THREADS, K2_SMEM, st>>>(tmX, tmQ, a);
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_tc_filter_kernel");
  const bool staged = (C & 3) == 0 && C >= 16;
  const size_t rsm = !staged ? 0 : (C == 64 ? (size_t)8 * (64 + 32 * 36) * 4 : (size_t)8 * 33 * (C + 4) * 4);
  const int fb_grid = 4 * num_sms();
  if (k <= 32) {
    if (C == 64)
      knn_tc_refine_kernel<1, 64><<<cdiv(P, 8), 256, rsm, st>>>(x, s, cand, ccnt, N, Npad, C, k, cap, P, idx);
    else
      knn_tc_refine_kernel<1, 0><<<cdiv(P, 8), 256, rsm, st>>>(x, s, cand, ccnt, N, Npad, C, k, cap, P, idx);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("knn_tc_refine_kernel");
    knn_row_fallback_kernel<1><<<fb_grid, 256, 0, st>>>(x, s, fbq, N, Npad, C, k, idx);
  } else {
    if (C == 64)
      knn_tc_refine_kernel<2, 64><<<cdiv(P, 8), 256, rsm, st>>>(x, s, cand, ccnt, N, Npad, C, k, cap, P, idx);
    else
      knn_tc_refine_kernel<2, 0><<<cdiv(P, 8), 256, rsm, st>>>(x, s, cand, ccnt, N, Npad, C, k, cap, P, idx);
    count_launch();
    DG_CUDA_LAUNCH_CHECK("knn_tc_refine_kernel");
    knn_row_fallback_kernel<2><<<fb_grid, 256, 0, st>>>(x, s, fbq, N, Npad, C, k, idx);
  }
  count_launch();
  DG_CUDA_LAUNCH_CHECK("knn_row_fallback_kernel");
  return DGCNN_OK;
}

}  // namespace dgcnn

#ifdef K2_DEBUG
extern "C" int dgcnn_knn_debug_stamps(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, dgcnn::k2_dbg, sizeof(long long) * 512 * 16) == cudaSuccess ? 0 : -4;
}
#endif
