/*
 * oracle/knn_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Plain-C CPU restatement of the reference's k_nn():
 *   /root/reference/dgcnn/ops.py:8-19
 *     inner_prod = matmul(M, M^T)                       (ops.py:13)
 *     squared    = reduce_sum(square(M), axis=-1)       (ops.py:14)
 *     nn_dist    = squared + squared^T - 2 * inner_prod (ops.py:16)
 *     _, idx     = top_k(-nn_dist, k)                   (ops.py:18)
 *
 * PIN: the reference ships no tests / golden vectors, and TensorFlow 1.x cannot be installed
 * here -- but the reference's own k_nn() / edges() (ops.py:8-40, unmodified) run in this
 * container on top of oracle/tf1_shim, and tests/golden/ref_knn_edges.npz holds what they
 * return on clouds whose distances are exact in fp32 (duplicates and lattice ties included):
 * this file must reproduce those indices bit for bit (tests/test_oracle_vs_reference.py), and
 * on the reference's own feature-space inputs it may differ from them only where two
 * distances tie to rounding.  What stays unpinned is the accumulation order inside TF's
 * matmul kernel (unknowable without TF); this file FIXES an order and calls it "the answer":
 *   s_i  = sequential-in-c   s = fl(s + fl(x_c * x_c))   (square is its own TF op => rounded
 *          before the sum; no FMA)
 *   p_ij = sequential-in-c   p = fmaf(x_ic, x_jc, p), p0 = +0   (what a one-thread-per-output
 *          SGEMM inner loop does)
 *   D_ij = fl( fl(s_i + s_j) - fl(2 * p_ij) )            (precedence of ops.py:16; 2*p exact)
 *   top_k: k smallest D per row, ascending, ties -> lower index first
 *          (tf.nn.top_k sorted=True contract), int32 indices, self included.
 * Integer-lattice known-answer tests (tests/test_oracle_knn.py) pin the tie rule
 * independently of the accumulation order (all arithmetic exact there).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; fmaf is single-rounding
 * either via -mfma or glibc's exact software fmaf).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* squared norms, ops.py:14 */
void oracle_sqnorm(const float* x, int64_t P, int C, float* s) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < P; ++p) {
    const float* r = x + p * C;
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) {
      float sq = r[c] * r[c]; /* -ffp-contract=off: stays a rounded product */
      acc = acc + sq;
    }
    s[p] = acc;
  }
}

static inline float dist_ij(const float* xi, const float* xj, int C, float si, float sj) {
  float p = 0.0f;
  for (int c = 0; c < C; ++c) p = fmaf(xi[c], xj[c], p);
  float a = si + sj;
  float b = 2.0f * p;
  float d = a - b;
  return d + 0.0f; /* canonicalise -0 */
}

/* full matrix, ops.py:11-16.  D is [B,N,N] */
void oracle_pairwise_distance(const float* x, int B, int N, int C, float* D) {
  float* s = (float*)malloc(sizeof(float) * (size_t)B * N);
  oracle_sqnorm(x, (int64_t)B * N, C, s);
#pragma omp parallel for schedule(static)
  for (int64_t bi = 0; bi < (int64_t)B * N; ++bi) {
    int b = (int)(bi / N);
    const float* xb = x + (size_t)b * N * C;
    const float* sb = s + (size_t)b * N;
    const float* xi = x + (size_t)bi * C;
    float* row = D + (size_t)bi * N;
    for (int j = 0; j < N; ++j) row[j] = dist_ij(xi, xb + (size_t)j * C, C, s[bi], sb[j]);
  }
  free(s);
}

/* insertion of (d,j) into an ascending (d, idx) list of length k; scanning j in
 * increasing order makes "strictly smaller d" the only way to displace an entry,
 * which is exactly the lower-index-first tie rule. */
static inline void topk_row(const float* row, int N, int k, int32_t* out, float* dtmp) {
  int n = 0;
  for (int j = 0; j < N; ++j) {
    float d = row[j];
    if (n == k && !(d < dtmp[k - 1])) continue;
    int pos = (n < k) ? n : k - 1;
    while (pos > 0 && d < dtmp[pos - 1]) {
      dtmp[pos] = dtmp[pos - 1];
      out[pos] = out[pos - 1];
      --pos;
    }
    dtmp[pos] = d;
    out[pos] = j;
    if (n < k) ++n;
  }
}

/* ops.py:18 on an already materialised matrix: rows x N -> rows x k */
void oracle_topk_rows(const float* D, int64_t rows, int N, int k, int32_t* idx) {
#pragma omp parallel
  {
    float* dtmp = (float*)malloc(sizeof(float) * (size_t)k);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < rows; ++r) topk_row(D + (size_t)r * N, N, k, idx + (size_t)r * k, dtmp);
    free(dtmp);
  }
}

/* ops.py:8-19 end to end without materialising [B,N,N] (row at a time) */
void oracle_knn(const float* x, int B, int N, int C, int k, int32_t* idx) {
  float* s = (float*)malloc(sizeof(float) * (size_t)B * N);
  oracle_sqnorm(x, (int64_t)B * N, C, s);
#pragma omp parallel
  {
    float* row = (float*)malloc(sizeof(float) * (size_t)N);
    float* dtmp = (float*)malloc(sizeof(float) * (size_t)k);
#pragma omp for schedule(static)
    for (int64_t bi = 0; bi < (int64_t)B * N; ++bi) {
      int b = (int)(bi / N);
      const float* xb = x + (size_t)b * N * C;
      const float* sb = s + (size_t)b * N;
      const float* xi = x + (size_t)bi * C;
      for (int j = 0; j < N; ++j) row[j] = dist_ij(xi, xb + (size_t)j * C, C, s[bi], sb[j]);
      topk_row(row, N, k, idx + (size_t)bi * k, dtmp);
    }
    free(row);
    free(dtmp);
  }
  free(s);
}
