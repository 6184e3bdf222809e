/*
 * gkg_abi.h -- C ABI of libgkg_b200.so: the B200 (sm_100a) graph hot path of GKGNet.
 *
 * Drop-in boundary.  The reference (jin-s13/GKGNet) is pure Python; its hot path is the
 * functional seam
 *     DenseDilatedKnnGraph.forward(x, y=None, relative_pos=None) -> edge_index
 *         (mmcls/models/backbones/vig_model/torch_edge.py:164-176)
 *     MRConv2d.forward(x, edge_index, y=None)
 *         (mmcls/models/backbones/vig_model/torch_vertex.py:47-62)
 * A maintainer binds these entry points with ctypes; the host-side
 * mirror of the reference classes lives in gkgnet_b200/.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless
 *     the function name ends in _host;
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, a negative GKG_E* code otherwise; gkg_last_error() gives a
 *     thread-local message; no exceptions cross the ABI;
 *   - the caller owns every buffer (outputs and workspace included);
 *   - node features are addressed as element (b, n, c) at
 *         base + b*stride_b + n*stride_n + c        (element strides; channel stride 1)
 *     i.e. channels-last / token-major.  A contiguous NCHW tensor must be permuted by
 *     the caller (the Python mirror does it); `c = g*D + d` for channel group g of G.
 *   - "problem" p = b*G + g, matching the reference's (B*G, D, N, 1) regrouping
 *     (torch_vertex.py:197-202).
 */
#ifndef GKG_ABI_H_
#define GKG_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GKG_ABI_VERSION 2

/* feature dtypes */
#define GKG_F32 0
#define GKG_BF16 1

/* kNN algorithms */
#define GKG_KNN_AUTO 0       /* tcgen05 path when the shape allows it, else exact */
#define GKG_KNN_EXACT_FP32 1 /* CUDA-core fp32 brute force, reference association order */
#define GKG_KNN_TCGEN05 2    /* fp16 GEMM on tcgen05/TMEM (raw 16-bit rows, or the fp16x3 split of fp32 rows) + fused top-k + exact re-rank */

/* error codes */
#define GKG_OK 0
#define GKG_EINVAL (-1)   /* bad argument / unsupported shape */
#define GKG_ECUDA (-2)    /* CUDA runtime error (message in gkg_last_error) */
#define GKG_EWORKSPACE (-3) /* workspace too small */

typedef void* gkg_stream_t; /* cudaStream_t */

int gkg_abi_version(void);
const char* gkg_last_error(void);

/* Bytes of scratch gkg_knn_graph needs for this shape, dtype and algorithm (normalised keys, norms,
 * tensor-core operands, candidate lists).  self_keys != 0 when y == NULL (keys are the queries). */
size_t gkg_knn_workspace_bytes(int B, int G, int N, int M, int D, int k, int dilation,
                               int self_keys, int dtype, int algo);

/*
 * Dilated (group) kNN graph.  Replaces DenseDilatedKnnGraph.forward
 * (torch_edge.py:164-176 -> xy_dense_knn_matrix :89-106 / dense_knn_matrix :54-86 ->
 * DenseDilated :139-149): L2-normalise queries and keys over the D channels of each
 * group, rank keys by (|x|^2 - 2 x.y) + |y|^2 (+ relative_pos[n, m]), keep the
 * k*dilation nearest in ascending order and emit ranks 0, d, 2d, ...
 *
 *   x        queries, (B, N, G*D) via strides, dtype `dtype`
 *   y        keys, (B, M, G*D) via strides, or NULL -> keys are the queries (M == N)
 *   relpos   fp32 (N, M) row-major bias shared by all problems, or NULL
 *   idx_out  int32 (B*G, N, k): neighbour ids (edge_index[0]; edge_index[1][p,n,:] == n)
 *   limits   D <= 640 channels per group, k*dilation <= 36 (tensor-core path) / 64 (exact path)
 *
 * Ties: the smaller key id wins (torch.topk leaves tie order unspecified).
 *
 * Optional separable form of the bias (a hint that only accelerates the tensor-core path; the
 * dense `relpos` stays authoritative and is what exact re-ranking reads):
 *     relpos[n, m] == sep_a[n % sep_grid_w][m % sep_kw] + sep_b[n / sep_grid_w][m / sep_kw]
 * sep_a is fp32 (sep_grid_w, sep_kw), sep_b fp32 (N / sep_grid_w, M / sep_kw); pass NULL / 0
 * when unknown.  The analytic table of the reference (pos_embed.py:21-29 +
 * torch_vertex.py:309-315) always has this form; the caller must have verified it to within 5e-7 (absolute) of
 * `relpos` -- the residual is added to the error bound of the certified ordering.
 */
int gkg_knn_graph(const void* x, int64_t x_stride_b, int64_t x_stride_n,
                  const void* y, int64_t y_stride_b, int64_t y_stride_n,
                  const float* relpos, const float* relpos_sep_a, const float* relpos_sep_b,
                  int sep_grid_w, int sep_kw, int32_t* idx_out,
                  int B, int G, int N, int M, int D, int k, int dilation, int dtype,
                  int algo, void* workspace, size_t workspace_bytes, gkg_stream_t stream);

/*
 * The two phases of gkg_knn_graph, exposed separately so a caller can time them.  gkg_knn_prepare fills the
 * workspace (tensor-core operands, normalised fp32 keys, norms); gkg_knn_select ranks and writes idx_out (it
 * reads the query features again only for the few rows it re-ranks exactly).  Same arguments and workspace
 * as gkg_knn_graph; gkg_knn_graph == prepare followed by select.
 */
int gkg_knn_prepare(const void* x, int64_t x_stride_b, int64_t x_stride_n,
                    const void* y, int64_t y_stride_b, int64_t y_stride_n,
                    int B, int G, int N, int M, int D, int k, int dilation, int dtype,
                    int algo, void* workspace, size_t workspace_bytes, gkg_stream_t stream);
int gkg_knn_select(const void* x, int64_t x_stride_b, int64_t x_stride_n,
                   const float* relpos, const float* relpos_sep_a, const float* relpos_sep_b,
                   int sep_grid_w, int sep_kw, int32_t* idx_out,
                   int B, int G, int N, int M, int D, int k, int dilation, int self_keys, int dtype,
                   int algo, void* workspace, size_t workspace_bytes, gkg_stream_t stream);

/*
 * Testing only: gkg_knn_graph with per-call debug options (no process state is involved).
 *   debug_flags  1 = re-rank every row with the exact formula, 3 = send every row through the brute-force
 *                fix-up kernel, 0 = as gkg_knn_graph
 *   dbg_dist     fp32 (B*G, N, M) device buffer receiving the approximate distances (minus |xh|^2) as the
 *                tensor-core kernel ranks them, or NULL
 *   stats_out    3 HOST words: rows taken by the fix-up kernel, rows re-ranked exactly, bits of the largest
 *                |approximate - exact| distance seen by the re-rank; NULL = none (non-NULL synchronises)
 *   skip, ga     tuning overrides of the threshold sweep (-1 / 0 = automatic)
 */
int gkg_knn_graph_debug(const void* x, int64_t x_stride_b, int64_t x_stride_n,
                        const void* y, int64_t y_stride_b, int64_t y_stride_n,
                        const float* relpos, const float* relpos_sep_a, const float* relpos_sep_b,
                        int sep_grid_w, int sep_kw, int32_t* idx_out,
                        int B, int G, int N, int M, int D, int k, int dilation, int dtype,
                        int algo, void* workspace, size_t workspace_bytes, gkg_stream_t stream,
                        int debug_flags, float* dbg_dist, unsigned int* stats_out, int skip, int ga);

/*
 * Max-relative aggregation, forward.  Replaces the gather / subtract / max / interleave
 * part of MRConv2d.forward (torch_vertex.py:49-61; batched_index_select torch_nn.py:84-105):
 *     out[b, n, 2c]   = x[b, n, c]
 *     out[b, n, 2c+1] = max_j y[b, idx[b*G+g, n, j], c] - x[b, n, c],   g = c / D
 *   y == NULL gathers from x (M == N).  out is contiguous (B, N, 2*G*D), same dtype.
 *   argmax (uint8 (B, N, G*D), nullable) records the winning j for the backward pass.
 */
int gkg_mr_aggregate_fwd(const void* x, int64_t x_stride_b, int64_t x_stride_n,
                         const void* y, int64_t y_stride_b, int64_t y_stride_n,
                         const int32_t* idx, void* out, uint8_t* argmax,
                         int B, int G, int N, int M, int D, int k, int dtype,
                         gkg_stream_t stream);

/*
 * Max-relative aggregation, backward (what autograd derives for torch_vertex.py:49-61:
 * max-backward routes to the arg-max neighbour, index_put_(accumulate) into y).
 *   grad_out      (B, N, 2*G*D) contiguous, dtype
 *   grad_x        (B, N, G*D) contiguous, dtype:  g[2c] - g[2c+1]
 *   grad_y_accum  fp32 (B, M, G*D) contiguous, MUST be zero-filled by the caller;
 *                 receives sum over (n, j == argmax) of g[2c+1]   (atomic adds)
 */
int gkg_mr_aggregate_bwd(const void* grad_out, const int32_t* idx, const uint8_t* argmax,
                         void* grad_x, float* grad_y_accum,
                         int B, int G, int N, int M, int D, int k, int dtype,
                         gkg_stream_t stream);

/*
 * Deterministic form of gkg_mr_aggregate_bwd (bitwise reproducible whatever the order of the atomics): the scattered
 * sum runs in 64-bit fixed point.
 *   grad_y_fixed  int64 (B, M, G*D), zero-filled by the caller: sum of round(g * scale)
 *   scale         DEVICE pointer to one fp32 power of two, e.g. 2^(40 - ceil(log2(max |grad_out|))): ~40 bits below the
 *                 largest gradient, so the rounding is far below fp32 resolution and 2^23 terms cannot overflow
 * gkg_fixed_to_float converts: out[i] = in[i] / scale.
 */
int gkg_mr_aggregate_bwd_det(const void* grad_out, const int32_t* idx, const uint8_t* argmax,
                             void* grad_x, long long* grad_y_fixed, const float* scale,
                             int B, int G, int N, int M, int D, int k, int dtype, gkg_stream_t stream);
int gkg_fixed_to_float(const long long* in, const float* scale, float* out, long long n, gkg_stream_t stream);

/*
 * Grouped 1x1 FC of the max-relative convolution with norm and activation folded in (inference form).
 * Replaces MRConv2d.nn = BasicConv([2C, 2C]) = Conv2d(2C, 2C, 1, groups=4, bias) -> norm -> act
 * (torch_nn.py:57-81; torch_vertex.py:45,61) in eval mode, where the batch norm is an affine map:
 *     out[r, o] = act(sum_i (scale[o] * W[o, i]) * in[r, (o / CG) * CG + i] + shift[o]),   CG = C2 / 4
 * with scale[o] = gamma / sqrt(running_var + eps) (1 when there is no norm).
 *   in, out   bf16 (rows, C2) contiguous (token-major: rows = B * N, C2 = 2C interleaved channels)
 *   w_op      bf16 SCALED weights in tensor-core operand order: 4 groups x [NP/8][KP/8][8][8] with
 *             element [q][n/8][k/8][n%8][k%8] = scale[q*CG + n] * W[q*CG + n][k] (zero for n, k >= CG),
 *             NP = KP = ceil16(CG)
 *   shift     fp32 (C2): (conv_bias - running_mean) * scale + beta
 *   act       0 none, 1 relu, 2 gelu (erf form, nn.GELU())
 * gkg_grouped_fc_supported(C2) != 0 for every width a kernel exists for (2C a multiple of 32, CG <= 320 and beyond).
 * Narrow groups (CG <= 96) keep four accumulators in tensor memory and take the weight operand laid out as above;
 * wider groups (stages 3 - 4: CG = 200, 320) run one (conv group, column pass) at a time and take
 *     w_op  [4][passes][NT/8][KP/8][8][8],  element [q][p][n/8][k/8][n%8][k%8] = scale * W[q*CG + p*NT + n][k]
 * with NT = gkg_grouped_fc_pass_width(C2) output channels per pass (0 = narrow layout), passes = ceil(ceil16(CG) / NT).
 * The data gradient of the convolution is the same call with the per-group transposed weights.
 * Launches with at least two 128-row tiles per SM and CG <= 80 copy the rows in through a 3-D TMA tensor map
 * (cuTensorMapEncodeTiled, resolved with cudaGetDriverEntryPoint: no link-time libcuda dependency; if the driver does not
 * provide it the per-thread cp.async kernel runs instead -- same results bit for bit).  The map is built from `in`, `rows`
 * and C2 at call time and baked into the launch, so a captured CUDA graph must be replayed with the same buffers (as for
 * every other pointer argument).  When CG is not a multiple of 16 the K padding of conv group q multiplies the first 8
 * channels of group q+1 of the same row by zero weights: non-finite values there reach group q of that row.
 */
int gkg_grouped_fc_supported(int C2);
int gkg_grouped_fc_pass_width(int C2);
int gkg_grouped_fc_fwd(const void* in, const void* w_op, const float* shift, void* out, long long rows,
                       int C2, int act, gkg_stream_t stream);

/* Packs a (2C, 2C/4, 1, 1) fp32 conv weight into the operand order gkg_grouped_fc_fwd takes for this width (see above):
 * w_op[..] = bf16(scale[o] * W[o][i]) (scale fp32 (2C,) or NULL), per-group transposed when `transpose` != 0 (the data
 * gradient).  w_op holds 4 * rows * KP bf16 with rows = passes * NT (wide) or KP (narrow). */
int gkg_grouped_fc_pack_weights(const float* weight, const float* scale, void* w_op, int C2, int transpose,
                                gkg_stream_t stream);

/*
 * Weight gradient of the grouped 1x1 FC (what autograd derives for Conv2d(2C, 2C, 1, groups=4), torch_nn.py:61):
 *     grad_w[q][o][i] += sum_r grad_out[r, q*CG + o] * in[r, q*CG + i]
 *   grad_out, in  bf16 (rows, C2) contiguous;  grad_w  fp32 (4, CG, CG) == the (2C, 2C/4, 1, 1) weight layout,
 *   MUST be zero-filled by the caller (the row range is split over CTAs, partial sums are added atomically).
 * tcgen05 kernel: both operands are read as they lie in memory (channel-contiguous rows are the MN-major operand form).
 */
int gkg_grouped_fc_wgrad(const void* grad_out, const void* in, float* grad_w, long long rows, int C2,
                         gkg_stream_t stream);

/*
 * Key pooling of the dynamic graph convolution.  Replaces `y = F.avg_pool2d(x, r, r)` of
 * DyGraphConv2dMultiGroup.forward / DyGraphConv2d.forward (torch_vertex.py:194-196, :221-223) on the
 * token-major layout: window r x r, stride r, no padding, floor mode (rows / columns past
 * floor(H/r)*r are dropped), fp32 accumulation in row-major window order, one division by r*r.
 *   x        (B, H*W, C) via strides, dtype
 *   y        (B, (H/r)*(W/r), C) contiguous, dtype
 * Backward (what autograd derives for avg_pool2d): grad_x[b, ih, iw, c] = grad_y[b, ih/r, iw/r, c] / r^2,
 * zero in the dropped rows / columns.  grad_y and grad_x are contiguous, dtype.
 */
int gkg_pool_keys_fwd(const void* x, int64_t x_stride_b, int64_t x_stride_n, void* y,
                      int B, int H, int W, int C, int r, int dtype, gkg_stream_t stream);
int gkg_pool_keys_bwd(const void* grad_y, void* grad_x,
                      int B, int H, int W, int C, int r, int dtype, gkg_stream_t stream);

/*
 * Neighbour gather / neighbour sum for the graph-convolution variants that need the gathered rows (EdgeConv2d,
 * GraphSAGE, GraphAtten, GINConv2d: torch_vertex.py:16-150).  Replaces batched_index_select (torch_nn.py:84-105) on
 * the token-major layout:
 *     gather: out[b, n, j, c] = y[b, idx[b*G + c/D, n, j], c]        out (B, N, k, G*D) contiguous, dtype
 *     sum:    out[b, n, c]    = sum_j y[b, idx[b*G + c/D, n, j], c]  out (B, N, G*D) contiguous, dtype (fp32 accumulation)
 * Backward (index_put_(accumulate=True) of autograd): grad_y_accum fp32 (B, M, G*D), zero-filled by the caller, receives
 * the scattered gradients (atomic adds).
 */
int gkg_neighbor_gather_fwd(const void* y, int64_t y_stride_b, int64_t y_stride_n, const int32_t* idx, void* out,
                            int B, int G, int N, int M, int D, int k, int dtype, gkg_stream_t stream);
int gkg_neighbor_gather_bwd(const void* grad, const int32_t* idx, float* grad_y_accum,
                            int B, int G, int N, int M, int D, int k, int dtype, gkg_stream_t stream);
int gkg_neighbor_sum_fwd(const void* y, int64_t y_stride_b, int64_t y_stride_n, const int32_t* idx, void* out,
                         int B, int G, int N, int M, int D, int k, int dtype, gkg_stream_t stream);
int gkg_neighbor_sum_bwd(const void* grad, const int32_t* idx, float* grad_y_accum,
                         int B, int G, int N, int M, int D, int k, int dtype, gkg_stream_t stream);

/*
 * Label-query head.  Replaces LabelQueryHead.get_score (mmcls/models/heads/label_query_head.py:49-57: fc1 on all
 * label embeddings -- a (B, n, n) product -- masked to its diagonal, plus fc2(gap)):
 *     score[b, i] = W1[i] . L[b, i] + b1[i] + W2[i] . gap[b] + b2[i]
 *   L fp32 (B, n, C) contiguous, gap fp32 (B, C), W1 / W2 fp32 (n, C), b1 / b2 fp32 (n), score fp32 (B, n).
 * gkg_label_score_bwd: what autograd derives for it -- dL (B, n, C), dgap (B, C), dW1 / dW2 (n, C), db1 / db2 (n) from
 * dscore (B, n); every output is overwritten.
 */
int gkg_label_score_fwd(const float* L, const float* gap, const float* W1, const float* b1, const float* W2,
                        const float* b2, float* score, int B, int n, int C, gkg_stream_t stream);
int gkg_label_score_bwd(const float* dscore, const float* L, const float* gap, const float* W1, const float* W2,
                        float* dL, float* dgap, float* dW1, float* dW2, float* db1, float* db2,
                        int B, int n, int C, gkg_stream_t stream);

/*
 * The two losses of LabelQueryHead.forward_train (label_query_head.py:70-85) on `total` (score, target) entries:
 *   sums[0] = sum of AsymmetricLoss terms (losses/asymmetric_loss.py:9-72: sigmoid, clip, gamma_pos / gamma_neg, eps)
 *   sums[1] = sum of BCE-with-logits terms against the smoothed target t (1 - 2 smooth) + smooth
 *             (label_smooth_loss.py:122-126, 168-175)
 *   d_asl, d_bce  fp32 (total): derivatives of the two sums with respect to every score.
 * The caller divides by the batch and applies the x10 weight of the reference.
 */
int gkg_multilabel_loss(const float* score, const float* target, float* sums, float* d_asl, float* d_bce,
                        long long total, float gamma_pos, float gamma_neg, float clip, float eps, float smooth,
                        gkg_stream_t stream);

/*
 * Training-mode batch-norm reductions on the token-major (rows, C) layout (rows = B*H*W of a channels-last tensor).
 * Replace the statistics / backward-reduce halves of `norm_layer('batch')` = (Sync)BatchNorm2d in training
 * (torch_nn.py:32-42; used by BasicConv torch_nn.py:61-65, Grapher.fc1 / fc2 torch_vertex.py:290-306, FFN
 * gkgnet.py:46-72); the elementwise halves are the caller's (ATen batch_norm_elemt / batch_norm_backward_elemt).
 *   gkg_bn_stats:  mean[c], invstd[c] = rsqrt(biased var + eps) over the rows; when running_mean / running_var are
 *     given they are updated like nn.BatchNorm2d (momentum, unbiased variance).  Deterministic (no atomics).
 *   gkg_bn_backward_reduce:  sum_dy[c] = sum_r dy, sum_dy_xmu[c] = sum_r dy * (x - mean[c]); optionally
 *     grad_weight = sum_dy_xmu * invstd and grad_bias = sum_dy.
 *   x, grad_out  (rows, C) contiguous, dtype; C a multiple of 8 (bf16) / 4 (fp32); everything else fp32 (C).
 *   ws: gkg_bn_workspace_bytes(rows, C) bytes of scratch (per-block partials).
 */
size_t gkg_bn_workspace_bytes(long long rows, int C);
int gkg_bn_stats(const void* x, long long rows, int C, int dtype, float eps, float momentum,
                 float* mean, float* invstd, float* running_mean, float* running_var,
                 void* ws, size_t ws_bytes, gkg_stream_t stream);
int gkg_bn_backward_reduce(const void* grad_out, const void* x, const float* mean, const float* invstd,
                           long long rows, int C, int dtype, float* sum_dy, float* sum_dy_xmu,
                           float* grad_weight, float* grad_bias, void* ws, size_t ws_bytes, gkg_stream_t stream);

/*
 * norm -> GELU pairs (BasicConv torch_nn.py:61-65, FFN.fc1 gkgnet.py:52-58, Stem gkgnet.py:82-90) in training:
 *   gkg_bn_act_forward:   y = gelu(z), z = (x - mean) * invstd * gamma + beta, with the statistics of gkg_bn_stats;
 *                         dact (nullable, (rows, C), dtype) receives gelu'(z) for the backward
 *   gkg_bn_act_backward:  everything autograd derives for that pair from grad_out = dL/dy: grad_x (rows, C),
 *                         grad_weight = dL/dgamma, grad_bias = dL/dbeta.  With dact the two passes are pure streaming
 *                         kernels; without it the derivative is re-evaluated from x (nothing but x is saved).
 * act: 2 = GELU (erf form); x, y, grad_out, grad_x (rows, C) contiguous, dtype; ws as above.
 */
int gkg_bn_act_forward(const void* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                       long long rows, int C, int dtype, int act, void* y, void* dact, gkg_stream_t stream);
int gkg_bn_act_backward(const void* grad_out, const void* x, const void* dact, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, long long rows, int C, int dtype, int act,
                        void* grad_x, float* grad_weight, float* grad_bias, void* ws, size_t ws_bytes,
                        gkg_stream_t stream);
/* The same backward in its two halves, for synchronised statistics (SyncBatchNorm, the reference's default norm,
 * torch_nn.py:8): the caller all-reduces `sums` (fp32 [2][C]: sum g, sum g (x - mean)) over the data-parallel ranks
 * between the calls and passes the total row count they cover. */
int gkg_bn_act_backward_reduce(const void* grad_out, const void* x, const void* dact, const float* mean,
                               const float* invstd, const float* gamma, const float* beta, long long rows, int C,
                               int dtype, int act, float* sums, float* grad_weight, float* grad_bias,
                               void* ws, size_t ws_bytes, gkg_stream_t stream);
int gkg_bn_act_backward_elemt(const void* grad_out, const void* x, const void* dact, const float* mean,
                              const float* invstd, const float* gamma, const float* beta, const float* sums,
                              long long rows, long long total_rows, int C, int dtype, int act, void* grad_x,
                              gkg_stream_t stream);

/*
 * Column sums of a (rows, C) contiguous activation into fp32 out (C): the bias gradient autograd derives for a
 * token-major 1x1 convolution (Grapher.fc1 / fc2 torch_vertex.py:290-306, FFN gkgnet.py:46-72).  Same row-range
 * partials as the batch-norm reductions (deterministic); ws: gkg_bn_workspace_bytes(rows, C).
 */
int gkg_column_sum(const void* x, long long rows, int C, int dtype, float* out, void* ws, size_t ws_bytes,
                   gkg_stream_t stream);

/* Number of kernels this library has launched since load (for bench accounting). */
uint64_t gkg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GKG_ABI_H_ */
