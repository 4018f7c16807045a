/*
 * dgcnn_b200.h -- C ABI of the B200-native EdgeConv hot path (libdgcnn_b200.so, sm_100a).
 *
 * The reference (DeepLearnPhysics/dynamic-gcnn) has no native code and no FFI: its boundary
 * is the Python function API of dgcnn/ops.py + dgcnn/model.py, whose arithmetic is executed
 * by TensorFlow-1.x kernels.  Each entry point below replaces the TF kernels behind one
 * piece of that API (file:line cited per function).  The Python mirror
 * (dynamic-gcnn_b200/dgcnn/ops.py) binds these with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers + sizes, all DEVICE pointers unless stated; fp32 / int32, row-major,
 *     channels-last ([B,N,C]) exactly like the reference's tensors (ops.py:9).
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), re-entrant,
 *     and allocates nothing: the caller owns all buffers including `ws` scratch, sized by the
 *     matching *_workspace_bytes().
 *   - return 0 on success, negative DGCNN_ERR_* otherwise; text via dgcnn_last_error()
 *     (thread-local).  Bad shapes never abort the process.
 *   - no CPU fallback: without a CUDA device every compute entry returns DGCNN_ERR_CUDA.
 */
#ifndef DGCNN_B200_H_
#define DGCNN_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGCNN_OK 0
#define DGCNN_ERR_INVALID (-1)     /* bad shape / null pointer / misaligned buffer */
#define DGCNN_ERR_UNSUPPORTED (-2) /* valid request outside the compiled envelope (e.g. k > 64) */
#define DGCNN_ERR_WORKSPACE (-3)   /* ws_bytes too small */
#define DGCNN_ERR_CUDA (-4)        /* launch / runtime failure */

#define DGCNN_KNN_MAX_K 64

typedef void* dgcnn_stream_t; /* cudaStream_t */

int dgcnn_abi_version(void);
const char* dgcnn_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t dgcnn_launch_count(void);

/* One scratch-size query for every op of the path (dgcnn_workspace_bytes(op, B, N, C, k, F)): the bytes
 * of `ws` the named op needs for clouds [B,N,C], k neighbours and F output channels -- a dispatcher over the per-op
 * *_workspace_bytes() below (P = B*N rows).  Pure function of its arguments; 0 for an unknown op or an empty shape.  */
#define DGCNN_OP_KNN 0          /* dgcnn_knn / dgcnn_knn_hinted / dgcnn_knn_mode / dgcnn_pairwise_distance (k unused) */
#define DGCNN_OP_EDGECONV 1     /* dgcnn_edgeconv_{fwd,bwd}_{stats,apply} with F filters                            */
#define DGCNN_OP_CONV_FWD 2     /* dgcnn_gemm, [P,C] . [C,F]: uv (F = 2*filters) / conv1 forward                      */
#define DGCNN_OP_CONV_DW 3      /* dgcnn_gemm, [C,P] . [P,F]: the weight gradient of the same conv (split over points) */
#define DGCNN_OP_BN 4           /* dgcnn_bn_act_fwd / bwd and friends on [P,F]                                        */
#define DGCNN_OP_SOFTMAX_XENT 5 /* dgcnn_softmax_xent                                                                 */
size_t dgcnn_workspace_bytes(int op, int B, int N, int C, int k, int F);

/* ---- k_nn: dgcnn/ops.py:8-19 -------------------------------------------------------------
 * D[b,i,j] = fl(fl(s_i + s_j) - 2*p_ij), s = sum_c fl(x_c^2) (sequential, no FMA),
 * p = sequential-in-c fmaf chain (oracle/knn_oracle.c fixes this order).                   */
size_t dgcnn_knn_workspace_bytes(int B, int N, int C);
/* ops.py:11-16 ("pairwise_distance"): x [B,N,C] -> D [B,N,N] */
int dgcnn_pairwise_distance(const float* x, float* D, int B, int N, int C, void* ws, size_t ws_bytes,
                            dgcnn_stream_t stream);
/* ops.py:8-19 fused ("knn"): the [B,N,N] matrix never leaves the SM.  idx [B,N,k] int32, nearest
 * first, self included, equal distances -> lower index first (tf.nn.top_k sorted=True).     */
int dgcnn_knn(const float* x, int32_t* idx, int B, int N, int C, int k, void* ws, size_t ws_bytes,
              dgcnn_stream_t stream);
/* Same result as dgcnn_knn, bit for bit, with a warm start: hint [B,N,k] int32 holds, per row, k DISTINCT
 * in-range column indices (typically the previous EdgeConv layer's neighbours of the same clouds: the reference
 * recomputes the graph per layer, ops.py:91-96).  Their exact distances bound the k-th smallest from above.  Only the
 * SIMT path (clouds of fewer than 256 points, or more than 64 channels) uses it; the tensor-core path derives a tighter
 * bound from its own first sweep and ignores the hint.  hint == NULL is dgcnn_knn.  A hint row with duplicate indices
 * violates the precondition (the bound would be invalid).                                                      */
int dgcnn_knn_hinted(const float* x, const int32_t* hint, int32_t* idx, int B, int N, int C, int k, void* ws,
                     size_t ws_bytes, dgcnn_stream_t stream);
/* Same again with the precision mode of the tensor-core distance filter chosen by the caller instead of by the
 * library's rule (AUTO: fine for clouds of more than 4096 points).  The result is identical in every mode; COARSE / FINE
 * exist for A/B timing.  The library itself keeps no such switch: there is no environment variable or other global
 * mutable state behind any entry point.                                                                         */
#define DGCNN_KNN_AUTO (-1)
#define DGCNN_KNN_COARSE 0
#define DGCNN_KNN_FINE 1
int dgcnn_knn_mode(const float* x, const int32_t* hint, int32_t* idx, int B, int N, int C, int k, int filter_mode,
                   void* ws, size_t ws_bytes, dgcnn_stream_t stream);
/* ops.py:18 on a materialised matrix: D [rows,N] -> idx [rows,k] (k smallest, same tie rule) */
int dgcnn_topk_rows(const float* D, int32_t* idx, int64_t rows, int N, int k, dgcnn_stream_t stream);

/* ---- edges (= get_edge_feature): dgcnn/ops.py:21-40 ----------------------------------------
 * out[b,i,j,:] = concat(x[b,i,:], x[b,idx[b,i,j],:] - x[b,i,:])  -> [B,N,k,2C]              */
int dgcnn_edge_feature(const float* x, const int32_t* idx, float* out, int B, int N, int C, int k,
                       dgcnn_stream_t stream);
/* gradient of the above wrt x (tf.gather grad = scatter-add): g_out [B,N,k,2C] -> gx [B,N,C] */
int dgcnn_edge_feature_bwd(const float* g_out, const int32_t* idx, float* gx, int B, int N, int C, int k,
                           dgcnn_stream_t stream);

/* ---- 1x1 conv as a per-point GEMM: slim.conv2d kernel_size=1, ops.py:47-54,62-70 -----------
 * C[M,N] = op(A)[M,K] . op(B)[K,N], fp32, each output one sequential-in-k fmaf chain per
 * k-split.  transA: A is stored [K,M]; transB: B is stored [N,K].                            */
size_t dgcnn_gemm_workspace_bytes(int M, int N, int K, int transA, int transB);
int dgcnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, int transA, int transB,
               void* ws, size_t ws_bytes, dgcnn_stream_t stream);

/* ---- the same GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) ---------------------------------------
 * planes = 2, fp32-faithful ("bf16x3"): operands are pre-split fp32 -> two bf16 planes, planes[0] = bf16(x),
 *   planes[1] = bf16(x - planes[0]), stored back to back ([2][rows][cols], same row-major layout as the fp32
 *   source); the kernel accumulates hi.hi + hi.lo + lo.hi in fp32 (relative error ~2^-17 per product).
 * planes = 1, plain bf16 (the reduced-precision variant, BASELINE.json configs[2]): only planes[0] exists
 *   ([1][rows][cols]); one MMA per k-slice, fp32 accumulation.  Both operands of a call use the same mode.
 * M, N, K multiples of 8; all pointers 16-byte aligned.  transA / transB as in dgcnn_gemm; no transposed copies.  */
/* x [rows, cols] fp32 at row pitch ldx -> planes: hi at planes[r*ldo + c], lo at planes[plane_elems + r*ldo + c]
 * (plane_elems == 0: hi only, the plain-bf16 operand).  With ldo > cols several sources fill column slices of one
 * operand (concat by construction).  The same convention (plane distance 0 = hi only) holds for every "sink" below. */
int dgcnn_split_bf16(const float* x, int64_t rows, int cols, int64_t ldx, void* planes, int64_t ldo,
                     int64_t plane_elems, dgcnn_stream_t stream);
size_t dgcnn_tc_gemm_workspace_bytes(int M, int N, int K);
int dgcnn_tc_gemm(const void* a_planes, const void* b_planes, float* C, int M, int N, int K, int transA, int transB,
                  int planes, void* ws, size_t ws_bytes, dgcnn_stream_t stream);

/* Same product with op(A) taken from a column slice of a wider pair of planes: a_planes points at the first element
 * of the slice, rows are a_ld elements apart and the lo plane starts a_plane_elems elements after the hi plane (the
 * operand a producer already filled for another layer is reused, e.g. for a weight gradient X^T.g).              */
int dgcnn_tc_gemm_a_slice(const void* a_planes, int64_t a_ld, int64_t a_plane_elems, const void* b_planes, float* C,
                          int M, int N, int K, int transA, int transB, int planes, void* ws, size_t ws_bytes,
                          dgcnn_stream_t stream);

/* Same product plus, from the accumulator, the per-column sum and sum of squares of every 128-row tile:
 * colstats [ceil(M/128)][2][N] fp32 (train-mode BatchNorm statistics of the layer output, slim.batch_norm at
 * ops.py:53,158 / model.py:71, without a second pass over C).  Only for shapes the 128x256 persistent kernel takes
 * without a k-split (dgcnn_tc_gemm_stats_supported(M,N,K) == 1: N % 256 == 0, M >= 128).                       */
int dgcnn_tc_gemm_stats_supported(int M, int N, int K);
int dgcnn_tc_gemm_stats(const void* a_planes, const void* b_planes, float* C, int M, int N, int K, int transA,
                        int transB, int planes, float* colstats, dgcnn_stream_t stream);

/* Same product, but column ranges [starts[g], starts[g]+widths[g]) of the result go to separate contiguous
 * [M, widths[g]] buffers outs[g] (host arrays of n_groups <= 32 entries; 32-column aligned ranges).  Used for the
 * gradient of a multi-source (concatenated) operand: every source receives its own dense gradient tensor.   */
int dgcnn_tc_gemm_grouped(const void* a_planes, const void* b_planes, int M, int N, int K, int transA, int transB,
                          int planes, int n_groups, const int* starts, const int* widths, float* const* outs,
                          dgcnn_stream_t stream);

/* ---- fused EdgeConv core: ops.py:45-57 (edges -> conv0 -> BN(train) -> ReLU -> max_k, mean_k) ; ops.py:58 concat
 * Uses [x_i, x_j - x_i].[Wa;Wb] = x_i.(Wa-Wb) + x_j.Wb: the caller first forms
 * uv[P, 2F] = x[P,C] . [Wa-Wb | Wb] (P = B*N), so that z_ij = u_i + v_{idx(i,j)} and the [B,N,k,2C] / [B,N,k,F]
 * tensors are never materialised.  uv is fp32 (uv_dtype DGCNN_F32) or, for the reduced-precision variant, fp16 / bf16
 * (DGCNN_F16 / DGCNN_BF16: half the gather bytes; all arithmetic and every other tensor stay fp32); 16-byte aligned.
 * Three gather passes per layer (two forward, one backward):
 *   fwd_stats : BN batch statistics of z over all B*N*k edges: mean[F], rstd[F] = 1/sqrt(var_biased + 1e-3)
 *   fwd_apply : out_both[P,2F] = ( relu(bn(max_j z_ij)) | mean_j relu(bn(z_ij)) )  -- ops.py:58's concat in place;
 *               zmax[P,F] = max_j z_ij and npos[P,F] (uint8) = #{j : relu input > 0}: what the backward passes need
 *               (both NULL when no backward follows);
 *               sink_planes (may be NULL): the same values also as bf16 hi / lo planes in columns [0,2F) of a
 *               tensor-core operand (row pitch sink_ld elements, lo plane sink_plane_elems later; 0 = hi plane only):
 *               the consumer's concat operand (model.py:83-85) is filled by the producer
 *   bwd_stats : s1[F] = sum g_pre (= d beta), s2[F] = sum g_pre*zhat, from per-point quantities only (no gather): the
 *               forward outputs, npos, and the incoming gradients, which may arrive as separate [P,F] tensors (g_max,
 *               g_mean; either may be NULL) and/or as one packed [P,2F] tensor g_both (may be NULL); they are summed
 *               on the fly.  g_uv_clear (may be NULL): the v half of that [P,2F] buffer is zeroed for bwd_apply.
 *   bwd_apply : the only backward gather pass.  g_uv[P,2F]: u half written, v half scatter-added with 16-byte vector
 *               atomics (tf.gather's gradient; zeroed here unless v_half_cleared != 0).  The max gradient is shared
 *               equally among exact ties (tf.reduce_max's gradient).                                             */
#define DGCNN_F32 0
#define DGCNN_BF16 1
#define DGCNN_F16 2
size_t dgcnn_edgeconv_workspace_bytes(int F);
int dgcnn_edgeconv_fwd_stats(const void* uv, int uv_dtype, const int32_t* idx, int B, int N, int F, int k, float* mean,
                             float* rstd, void* ws, size_t ws_bytes, dgcnn_stream_t stream);
int dgcnn_edgeconv_fwd_apply(const void* uv, int uv_dtype, const int32_t* idx, int B, int N, int F, int k,
                             const float* mean, const float* rstd, const float* beta, float* out_both, float* zmax,
                             uint8_t* npos, void* sink_planes, int sink_ld, int64_t sink_plane_elems,
                             dgcnn_stream_t stream);
int dgcnn_edgeconv_bwd_stats(const float* out_both, const uint8_t* npos, const float* beta, const float* g_max,
                             const float* g_mean, const float* g_both, int B, int N, int F, int k, float* s1, float* s2,
                             float* g_uv_clear, void* ws, size_t ws_bytes, dgcnn_stream_t stream);
int dgcnn_edgeconv_bwd_apply(const void* uv, int uv_dtype, const int32_t* idx, int B, int N, int F, int k,
                             const float* mean, const float* rstd, const float* beta, const float* zmax,
                             const float* g_max, const float* g_mean, const float* g_both, const float* s1,
                             const float* s2, float* g_uv, int v_half_cleared, dgcnn_stream_t stream);

/* ---- train-mode BatchNorm (+residual) (+ReLU) on a [rows,C] per-point tensor ---------------
 * slim.batch_norm defaults (is_training=True, center=True, scale=False, eps=1e-3) after a 1x1 conv:
 * ops.py:53,68 ; residual add + relu: ops.py:134.
 *   fwd: out = act((z-mean)*rstd + beta [+ residual]); mean/rstd [C] are outputs (saved for bwd)
 *   bwd: g_pre = g_out * (relu ? out>0 : 1)  (optional output, = gradient of the residual)
 *        g_beta = sum g_pre ; g_z = rstd*(g_pre - mean(g_pre) - zhat*mean(g_pre*zhat))      */
size_t dgcnn_bn_workspace_bytes(int C);
int dgcnn_bn_act_fwd(const float* z, int64_t rows, int C, const float* beta, const float* residual, int relu,
                     float* out, float* mean, float* rstd, void* ws, size_t ws_bytes, dgcnn_stream_t stream);
int dgcnn_bn_act_bwd(const float* z, const float* out, const float* g_out, int64_t rows, int C,
                     const float* mean, const float* rstd, int relu, float* g_z, float* g_beta, float* g_pre,
                     void* ws, size_t ws_bytes, dgcnn_stream_t stream);

/* Same with a per-group additive term: row r gets group_bias[(r / group_rows), :] added to z before the statistics
 * (the model head's global max-pooled feature, tiled over the N points of each cloud by model.py:80-81, enters the
 * first FC layer as one [B, C] term instead of a [B*N, 1024] operand).  The gradient of the bias is the per-group
 * row sum of g_z (done by the caller).                                                                        */
int dgcnn_bn_act_fwd_gb(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                        const float* group_bias, int group_rows, int relu, float* out, float* mean, float* rstd,
                        void* ws, size_t ws_bytes, dgcnn_stream_t stream);
int dgcnn_bn_act_bwd_gb(const float* z, const float* out, const float* g_out, int64_t rows, int C, const float* mean,
                        const float* rstd, const float* group_bias, int group_rows, int relu, float* g_z,
                        float* g_beta, float* g_pre, void* ws, size_t ws_bytes, dgcnn_stream_t stream);

/* The same layer in pieces, for a 1x1 conv that ran on dgcnn_tc_gemm_stats:
 *   stats_from_tiles : colstats [tiles][2][C] (per 128-row tile: column sum, column sum of squares of z) -> mean, rstd,
 *                      including the per-group bias analytically (group_rows % 128 == 0)
 *   apply_fwd        : out = act((z [+ bias_g] - mean) * rstd + beta [+ residual]) with given statistics
 *   bwd_planes       : dgcnn_bn_act_bwd_gb whose g_z leaves as bf16 planes [n_planes][rows][C] -- the operand format of
 *                      the weight / input gradient GEMMs -- and, only if g_z != NULL, also as fp32.  out may be NULL
 *                      (layers without residual): the ReLU mask is then re-evaluated from z, mean, rstd, beta      */
int dgcnn_bn_stats_from_tiles(const float* colstats, int tiles, int C, int64_t rows, const float* group_bias,
                              int group_rows, float* mean, float* rstd, dgcnn_stream_t stream);
int dgcnn_bn_apply_fwd(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                       const float* group_bias, int group_rows, int relu, const float* mean, const float* rstd,
                       float* out, dgcnn_stream_t stream);
/* ..._sinks: the same forward passes with up to two plane sinks (host arrays of n_sinks entries: device pointer of the
 * first element of the column slice, row pitch, distance of the lo plane in elements): the output is also written as
 * bf16 hi / lo planes into the operands of the layers that consume it.  apply_fwd_sinks: out may be NULL when there is
 * at least one sink (every consumer reads the planes: the fp32 copy is not written at all).                          */
int dgcnn_bn_act_fwd_sinks(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                           const float* group_bias, int group_rows, int relu, float* out, float* mean, float* rstd,
                           void* ws, size_t ws_bytes, int n_sinks, void* const* sink_planes, const int* sink_lds,
                           const int64_t* sink_plane_elems, dgcnn_stream_t stream);
int dgcnn_bn_apply_fwd_sinks(const float* z, int64_t rows, int C, const float* beta, const float* residual,
                             const float* group_bias, int group_rows, int relu, const float* mean, const float* rstd,
                             float* out, int n_sinks, void* const* sink_planes, const int* sink_lds,
                             const int64_t* sink_plane_elems, dgcnn_stream_t stream);
int dgcnn_bn_act_bwd_planes(const float* z, const float* out, const float* beta, const float* g_out, int64_t rows, int C,
                            const float* mean, const float* rstd, const float* group_bias, int group_rows, int relu,
                            float* g_z, void* g_z_planes, int n_planes, float* g_beta, const float* pool_max,
                            const float* pool_cnt, const float* pool_grad, int pool_rows, void* ws, size_t ws_bytes,
                            dgcnn_stream_t stream);

/* A layer whose output is ALSO max-pooled over the rows of each group (MergedEdgeConv followed by max_pool_v2 over the
 * N points of a cloud, model.py:65-81), without a pass of its own for the pool:
 *   apply_fwd_pool : out = relu((z - mean) * rstd + beta) to the plane sinks and, if out != NULL, as fp32 (when every
 *                    consumer reads the planes the fp32 copy is never written); pool_max / pool_cnt [rows/pool_rows, C]
 *                    = maximum of each group and the number of rows attaining it.  ws: dgcnn_bn_pool_workspace_bytes.
 *   bwd_planes     : with pool_max / pool_cnt / pool_grad given (NULL = no pool), the pool's gradient
 *                    [out == pool_max] * pool_grad / pool_cnt (ties share, like tf.reduce_max / MaxPoolGrad) is added to
 *                    g_out on the fly; out is re-evaluated from z (pass out = NULL and beta).                          */
size_t dgcnn_bn_pool_workspace_bytes(int groups, int C);
int dgcnn_bn_apply_fwd_pool(const float* z, int64_t rows, int C, const float* beta, const float* mean, const float* rstd,
                            float* out, int pool_rows, float* pool_max, float* pool_cnt, void* ws, size_t ws_bytes,
                            int n_sinks, void* const* sink_planes, const int* sink_lds, const int64_t* sink_plane_elems,
                            dgcnn_stream_t stream);

/* ---- global max over the points of each cloud: gen_nn_ops.max_pool_v2 ksize [1,N,1,1], model.py:77 ----------
 * x [groups, rows, C] -> out [groups, C] and cnt [groups, C] (# points attaining the max); the gradient goes to the
 * arg-max, shared equally among exact ties.                                                                  */
int dgcnn_group_max_fwd(const float* x, int groups, int rows, int C, float* out, float* cnt, dgcnn_stream_t stream);
int dgcnn_group_max_bwd(const float* x, const float* out, const float* cnt, const float* g_out, int groups, int rows,
                        int C, float* g_x, dgcnn_stream_t stream);

/* g_x_inout [groups, rows, C] += that gradient, touching only the arg-max positions (fuses the accumulation with the
 * gradient the tensor receives from its other consumer: model.py:83-85 feeds the pooled tensor to the concat too). */
int dgcnn_group_max_bwd_add(const float* x, const float* out, const float* cnt, const float* g_out, int groups, int rows,
                            int C, float* g_x_inout, dgcnn_stream_t stream);

/* ---- loss head: trainval.py:39-52 in one pass -------------------------------------------------------------------
 * logits [P,K], labels int64 [P], optional weights [P]:
 *   loss_acc[0] = mean_p( xent(softmax(logits_p), label_p) * weight_p ), loss_acc[1] = mean_p( argmax_p == label_p ),
 *   grad [P,K] = d loss_acc[0] / d logits.  Labels outside [0,K) are the caller's error (unchecked, like TF on GPU). */
size_t dgcnn_softmax_xent_workspace_bytes(void);
int dgcnn_softmax_xent(const float* logits, const int64_t* labels, const float* weights, int64_t P, int K, float* grad,
                       float* loss_acc, void* ws, size_t ws_bytes, dgcnn_stream_t stream);

/* ---- tf.train.AdamOptimizer update on a flat buffer: trainval.py:17,80 --------------------
 * g' = g*grad_scale; m = b1*m+(1-b1)g'; v = b2*v+(1-b2)g'^2; p -= lr_t*m/(sqrt(v)+eps),
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed by the caller (epsilon outside the bias correction). */
int dgcnn_adam_tf_step(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1,
                       float b2, float eps, float grad_scale, dgcnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DGCNN_B200_H_ */
