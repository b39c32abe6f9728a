/*
 * rank_b200.h — C ABI of librank_b200.so: the B200 (sm_100a) CTR hot path.
 *
 * The reference (reallinshengxiang/Implementation-of-Rank-Algorithm-for-Mainstream-
 * Recommender-Systems) has no FFI of its own: its hot path is the body of each
 * nn.Module.forward plus the autograd backward of it.  Every entry point below replaces one
 * such body (cited as file:line relative to the reference root, algorithm/...).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the comment says "host";
 *   - tables are fp32 row-major [rows, dim]; indices are int64 (torch.long), lengths int64;
 *   - all work is enqueued on `stream`; the library never synchronises, allocates or frees
 *     device memory, so a whole training step is CUDA-graph capturable;
 *   - return value: 0 = ok, >0 = cudaError_t of the failing runtime call, <0 = argument
 *     error; rk_last_error() gives a thread-local message;
 *   - out-of-range indices are clamped to row 0 and raise bit 0 of *err_flag (the reference
 *     raises IndexError on CPU / device-asserts on CUDA); err_flag may be NULL.
 */
#ifndef RANK_B200_H
#define RANK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RK_ABI_VERSION 2
#define RK_MAX_FIELDS 24
#define RK_MAX_TABLES 32
#define RK_MAX_LAYERS 8

typedef void* rk_stream_t; /* cudaStream_t */

/* One sparse field: a table and the index column that reads it (nn.Embedding + lookup,
 * e.g. DeepFM/deepfm.py:90-98,123-132; DCN/dcn.py:130-137,163-167). */
typedef struct rk_field {
    const float*   weight;  /* [rows, dim] */
    const int64_t* idx;     /* [n] */
    int64_t        rows;    /* V */
    int32_t        dim;     /* D */
    int32_t        out_off; /* first column of this field in the concatenated row */
} rk_field_t;

/* One table's share of the embedding-gradient reduction: per-occurrence gradient rows live at
 * g + occurrence*ld (+0..dim-1); the dense gradient [rows, dim] is written to dw. */
typedef struct rk_grad_table {
    const float* g;
    int64_t      ld;
    float*       dw;     /* [rows, dim], must be zero-filled by the caller */
    int32_t      dim;
    int32_t      field;  /* which field's (sorted) occurrences feed this table */
} rk_grad_table_t;

int         rk_version(void);
const char* rk_last_error(void);
int         rk_device_sm_count(void);
/* Measurement aid (bench.py): occupy `stream` for `us` microseconds (<= 100 ms) so the host can
 * queue a step behind it and CUDA events around each call then see device time only. */
int         rk_debug_spin(int us, rk_stream_t stream);
long long   rk_launch_count(void);   /* kernels launched through this library so far (process-wide) */

/* ---- sparse embedding-gradient reduction (autograd of nn.Embedding: embedding_dense_backward,
 *      reached from loss.backward() e.g. DIN/din.py:346) -------------------------------------
 * rk_plan_build sorts all occurrences of F index columns by (field, row), stably, so that the
 * later reduction sums each row's gradients in occurrence order: deterministic, no atomics.
 *   idx[f]  device pointer to n[f] int64 indices, rows[f] = table height      (host arrays)
 *   sorted_keys[n_total], perm[n_total] : outputs (uint32); perm holds the occurrence number
 *   inside its field.  ws: scratch of rk_plan_workspace_bytes(n_total).
 *   Sequence fields may drop padded positions, whose gradient is identically zero, from the
 *   reduction: live_mode[f] = RK_LIVE_PREFIX keeps t < seq_len[f][b] of every sample b (n[f] =
 *   B*seq_T[f]); RK_LIVE_PREFIX_OR_EMPTY also keeps every t of samples with length 0 (DIN's
 *   softmax mode gives those uniform weights).  seq_len/seq_T/live_mode may be NULL (all live). */
#define RK_LIVE_ALL 0
#define RK_LIVE_PREFIX 1
#define RK_LIVE_PREFIX_OR_EMPTY 2
size_t rk_plan_workspace_bytes(int64_t n_total);
int rk_plan_build(const int64_t* const* idx, const int64_t* n, const int64_t* rows, int F,
                  const int64_t* const* seq_len, const int32_t* seq_T, const int32_t* live_mode,
                  uint32_t* sorted_keys, uint32_t* perm, void* ws, size_t ws_bytes,
                  int32_t* err_flag, rk_stream_t stream);
/* ws for the reduction: rk_reduce_workspace_bytes(sum over tables of ceil(n/16)*dim floats). */
size_t rk_reduce_workspace_bytes(const int64_t* n, int F, const rk_grad_table_t* tables,
                                 int n_tables);
int rk_embgrad_segment_reduce(const uint32_t* sorted_keys, const uint32_t* perm,
                              const int64_t* n, const int64_t* rows, int F,
                              const rk_grad_table_t* tables, int n_tables, void* ws,
                              size_t ws_bytes, rk_stream_t stream);

/* One-launch reduction for per-sample index columns (n <= RK_DIRECT_MAX_N occurrences per table; every
 * per-sample field at batch <= 8192), replacing rk_plan_build + a memset + rk_embgrad_segment_reduce for
 * those tables (same reference code: embedding_dense_backward under loss.backward(), e.g. DCN/dcn.py:166).
 * Output-partitioned: every CTA owns 256 rows of one table, scans the whole index column, orders the
 * occurrences that fall into its rows by (row, occurrence) and sums their gradient rows in that order
 * (runs longer than 32 occurrences: fixed interleaved slots + a fixed tree).  dw is written completely —
 * zeros for rows no occurrence touches — so it need NOT be pre-zeroed.  Tables of one launch that share
 * idx, rows and n (DeepFM / FwFM: the [rows,1] and [rows,D] tables of a field) share the scan and the order.
 * Deterministic, no float atomics.  Out-of-range indices count as row 0 and raise *err_flag. */
#define RK_DIRECT_MAX_N 8192
typedef struct rk_direct_table {
    const int64_t* idx;   /* [n] index column */
    const float*   g;     /* gradient row of occurrence o at g + o*ld (+0..dim-1) */
    int64_t        ld;
    float*         dw;    /* [rows, dim], written completely */
    int64_t        rows;
    int64_t        n;
    int32_t        dim;
    int32_t        reserved;
} rk_direct_table_t;
int rk_embgrad_direct_reduce(const rk_direct_table_t* tables, int n_tables, int32_t* err_flag,
                             rk_stream_t stream);

/* ---- gather + concat (the per-field lookup loops + torch.cat, DCN/dcn.py:163-169,
 *      DeepCrossing/deepcrossing.py:148-155) ------------------------------------------------
 * out[b, 0:n_dense] = dense[b, :]; out[b, off_f:off_f+dim_f] = W_f[idx_f[b], :]. */
int rk_gather_concat_fwd(const rk_field_t* fields, int F, const float* dense, int n_dense,
                         int64_t B, float* out, int ld_out, int32_t* err_flag,
                         rk_stream_t stream);

/* ---- DeepFM FM part (DeepFM/deepfm.py:121-142) ------------------------------------------
 * second[f] are the F second-order tables (dim D each, idx shared with first[f], dim 1).
 * deep_input[B, F*D] = cat_f e_f; fm_first[B] = sum_f w_f; fm_second[B] = 0.5*sum_d((sum_f e)^2
 * - sum_f e^2). */
int rk_deepfm_fwd(const rk_field_t* second, const float* const* first_weight, int F, int64_t B,
                  float* deep_input, float* fm_first, float* fm_second, int32_t* err_flag,
                  rk_stream_t stream);
/* Per-occurrence second-order gradients g_rows[B, F*D] = g_deep + g_second*(S - e_f);
 * g_deep or g_second may be NULL (treated as zero). */
int rk_deepfm_bwd(const float* deep_input, const float* g_deep, const float* g_second, int F,
                  int D, int64_t B, float* g_rows, rk_stream_t stream);

/* ---- FwFM (FwFM/fwfm.py:87-139; "next" row of the scope table) -----------------------------
 * second[f]: the F embedding tables (dim D each); first_weight[f]: the `linear` tables [rows,1],
 * same idx; field_weight[P], P = F(F-1)/2 in the reference's pair order (i outer, j inner,
 * fwfm.py:126-135); bias[1].  emb[B, F*D] = cat_f e_f (kept for the backward);
 * y[B] = sigmoid(sum_f w_f + sum_p r_p <e_i, e_j> + bias)  (fwfm.py:137-139). */
int rk_fwfm_fwd(const rk_field_t* second, const float* const* first_weight,
                const float* field_weight, const float* bias, int F, int64_t B, float* emb,
                float* y, int32_t* err_flag, rk_stream_t stream);
/* Rows of `partials` the backward may use (one per CTA). */
int rk_fwfm_bwd_ctas(void);
/* g_z[B] = g_y * y * (1-y) (the per-occurrence gradient of every first-order table);
 * g_rows[B, F*D]: per-occurrence gradients of the embedding rows;
 * g_pair[P+1] = [d field_weight (P), d bias]; partials: scratch [rk_fwfm_bwd_ctas()][P+1].
 * Sums over the batch have a fixed order (no float atomics). */
int rk_fwfm_bwd(const float* emb, const float* y, const float* g_y, const float* field_weight,
                int F, int D, int64_t B, float* g_rows, float* g_z, float* partials,
                float* g_pair, rk_stream_t stream);

/* ---- DCN CrossNet (cross_layer + loop, DCN/dcn.py:25-50,169-173) -------------------------
 * fwd: x0 = [dense | gathered rows] -> concat_all[B,d]; x_{l+1} = x0*(x_l.w_l) + b_l + x_l
 * -> cross_vec[B,d].  w, b: [L, d]. */
int rk_crossnet_fwd(const rk_field_t* fields, int F, const float* dense, int n_dense,
                    const float* w, const float* b, int L, int64_t B, float* concat_all,
                    float* cross_vec, int32_t* err_flag, rk_stream_t stream);
/* bwd: g_x0[B,d] = g_concat_all + d(cross_vec)/d(x0) . g_cross_vec (either may be NULL). */
int rk_crossnet_bwd(const float* concat_all, const float* w, const float* b, int L, int d,
                    int64_t B, const float* g_concat_all, const float* g_cross_vec,
                    float* g_x0, rk_stream_t stream);

/* Stand-alone cross_layer(x0, xl, index) (DCN/dcn.py:25-50): out = x0*(xl.w) + b + xl with
 * w, b of d floats; bwd gives g_x0 = g*(xl.w), g_xl = g + w*(g.x0). */
int rk_cross_layer_fwd(const float* x0, const float* xl, const float* w, const float* b, int d,
                       int64_t B, float* out, rk_stream_t stream);
int rk_cross_layer_bwd(const float* x0, const float* xl, const float* w, int d, int64_t B,
                       const float* g_out, float* g_x0, float* g_xl, rk_stream_t stream);

/* ---- DIN (DIN.forward DIN/din.py:294-323, din_attention :42-84) ----------------------------
 * One launch does every gather, the local activation unit on the history positions t < len,
 * the masked raw / scaled-softmax pooling, the concat and the per-sample L2 norm.
 *   concat_all[B,width] = [dense | cat rows | target row q | attention output]
 *   norm[B]   = || concat_all[b, l2_from:] ||_2      (l2_reg = lambda * mean(norm), DIN/din.py:322)
 *   att_w[B,T] = the attention weights w_t (raw masked scores or softmax), saved for backward
 *   relu_masks[B,T,3] = sign bits of the two hidden layers (64 + 32) for positions t < len
 * mlp: the per-call weights of att_net (DIN/din.py:61-67) packed as rk_din_mlp_floats(D) floats:
 *   [W1^T 4D x 64][b1 64][W2^T 64 x 32][b2 32][w3 32][b3 1, pad 3][W1 64 x 4D][W2 32 x 64]. */
typedef struct rk_din_args {
    const rk_field_t*   cat;          /* category fields; out_off = column in concat_all */
    int32_t             n_cat;
    int32_t             n_dense;
    const float* const* dense_cols;   /* host array of n_dense device pointers; value (b,c) at
                                         dense_cols[c][b * dense_stride] (the reference passes a
                                         dict of [B] tensors, DIN/din.py:296) */
    int64_t             dense_stride;
    rk_field_t          target;       /* feedid table, idx [B]; out_off = column of q */
    rk_field_t          history;      /* his_read_comment_7d_seq table, idx [B,T] */
    const int64_t*      hist_len;     /* [B] */
    int32_t             T;
    int32_t             att_off;      /* column of the attention output */
    int32_t             width;        /* row width of concat_all */
    int32_t             l2_from;      /* first column under the L2 norm */
    int32_t             use_softmax;
    int32_t             precision;    /* RK_DIN_FP32: fp32 SIMT MLP (1e-5 parity);
                                         RK_DIN_BF16_TENSOR: the activation-unit MLP on tcgen05 with
                                         bf16 operands / fp32 TMEM accumulators (D = 16, 2e-2 bar) */
    const float*        mlp;
    int64_t             B;
    void*               mlp_tiles;    /* RK_DIN_BF16_TENSOR only, may be NULL: rk_din_tile_bytes() bytes of
                                         scratch (128-byte aligned) that rk_din_fwd fills with the split-bf16
                                         operand tiles of `mlp` for both directions (one small prologue launch);
                                         every CTA then fetches its weights with one bulk (TMA) copy instead of
                                         converting them itself.  rk_din_bwd reads what the forward wrote: pass
                                         the same buffer, untouched in between. */
} rk_din_args_t;
#define RK_DIN_FP32 0
#define RK_DIN_BF16_TENSOR 1
int rk_din_tile_bytes(void);

int rk_din_mlp_floats(int D);
int rk_din_fwd(const rk_din_args_t* args, float* concat_all, float* norm, float* att_w,
               uint32_t* relu_masks, int32_t* err_flag, rk_stream_t stream);
/* g_row[B,width]: per-occurrence gradients of every concat column (tower gradient g_concat + L2
 * term through g_norm[B] + the attention's d/dq on the target columns); g_hist[B,T,D]: per-
 * occurrence gradients of the history rows, written only where the position is live (t < len,
 * or every t of a len == 0 sample in softmax mode). */
int rk_din_bwd(const rk_din_args_t* args, const float* concat_all, const float* norm,
               const float* att_w, const uint32_t* relu_masks, const float* g_concat,
               const float* g_norm, float* g_row, float* g_hist, int32_t* err_flag,
               rk_stream_t stream);

/* ---- AFM pairwise-interaction attention pooling (AFM/afm.py:92-115, attention net :84-88) ---
 * fields: F tables of equal dim D (4..32, multiple of 4), 2 <= F <= 16; w1 [A,D], b1 [A],
 * w2 [A] (attention.2.weight is [1,A]), b2 [1]; A <= 128.  out[B,D] = sum_p softmax_p(s) v_p.
 * bwd: g_rows[B,F*D] per-occurrence embedding gradients; the gradients of the four registered
 * attention tensors are written to ONE contiguous buffer g_w1|g_b1|g_w2|g_b2; partials is scratch
 * of n_ctas * (A*D + 2A + 1) floats with n_ctas = rk_afm_bwd_ctas(B, F). */
int rk_afm_bwd_ctas(int64_t B, int F);
int rk_afm_fwd(const rk_field_t* fields, int F, const float* w1, const float* b1, const float* w2,
               const float* b2, int A, int64_t B, float* out, int32_t* err_flag, rk_stream_t stream);
int rk_afm_bwd(const rk_field_t* fields, int F, const float* w1, const float* b1, const float* w2,
               const float* b2, int A, int64_t B, const float* g_out, float* g_rows, float* g_w1,
               float* g_b1, float* g_w2, float* g_b2, float* partials, int n_ctas,
               int32_t* err_flag, rk_stream_t stream);
/* The same forward with the attention MLP on the tensor cores (tcgen05.mma, split-bf16 operands
 * hi.hi + lo.hi + hi.lo, fp32 accumulation in TMEM); same arguments and outputs as rk_afm_fwd, plus
 * `tiles`: NULL, or rk_afm_tile_bytes() bytes of scratch (128-byte aligned) that the forward fills
 * once (one small prologue launch) with the weights' operand tiles for BOTH directions; every CTA
 * then fetches them by bulk (TMA) copy instead of converting them itself.  rk_afm_tc_bwd reads
 * what the forward wrote: pass the same buffer, untouched in between (or NULL). */
int rk_afm_tile_bytes(void);
int rk_afm_tc_fwd(const rk_field_t* fields, int F, const float* w1, const float* b1, const float* w2,
                  const float* b2, int A, int64_t B, float* out, void* tiles, int32_t* err_flag,
                  rk_stream_t stream);
/* The backward on the tensor cores: same arguments and outputs as rk_afm_bwd, with
 * n_ctas = rk_afm_tc_bwd_ctas(B, F).  The hidden layer only enters through its 0/1 ReLU mask
 * (exact in bf16), the other operands are split-bf16; the weight-gradient accumulator lives in
 * TMEM for the life of a CTA and the per-CTA partials are added in a fixed order.  A ReLU
 * decision whose pre-activation lies inside the error band of the split product is recomputed
 * in fp32, so the masks are the fp32 kernel's masks and the results meet the fp32 parity bar
 * (1e-5); these two entry points are what the AFM module uses by default. */
int rk_afm_tc_bwd_ctas(int64_t B, int F);
int rk_afm_tc_bwd(const rk_field_t* fields, int F, const float* w1, const float* b1, const float* w2,
                  const float* b2, int A, int64_t B, const float* g_out, float* g_rows, float* g_w1,
                  float* g_b1, float* g_w2, float* g_b2, float* partials, int n_ctas,
                  const void* tiles, int32_t* err_flag, rk_stream_t stream);

/* ---- fused tower layers: Dice (DIN/din.py:26-36) and the BatchNorm1d that
 *      follows it in the DIN tower (DIN/din.py:272-285), training mode ------------------------
 * x, z: [B, units] row-major.  xh = batchnorm(x; eps1, no affine), p = sigmoid(xh),
 * y = alpha*(1-p)*x + p*x, z = gamma * batchnorm(y; eps2) + beta; gamma = beta = NULL: z = y.
 * Running statistics (may be NULL) are updated in place as nn.BatchNorm1d does (momentum,
 * unbiased variance, num_batches += 1).  stats[4, units] = mean1 | rstd1 | mean2 | rstd2 is kept
 * for the backward.  1 <= B <= rk_dice_bn_max_batch() (a CTA holds the whole batch of its
 * columns in registers; every batch statistic is a fixed-order block reduction). */
int rk_dice_bn_max_batch(void);
int rk_dice_bn_fwd(const float* x, int64_t B, int units, const float* alpha, float eps1,
                   const float* gamma, const float* beta, float eps2, float momentum1,
                   float* running_mean1, float* running_var1, int64_t* num_batches1,
                   float momentum2, float* running_mean2, float* running_var2,
                   int64_t* num_batches2, float* z, float* stats, rk_stream_t stream);
/* g_x[B, units]; g_alpha, g_gamma, g_beta: [units] (complete sums over the batch). */
int rk_dice_bn_bwd(const float* x, const float* g_z, int64_t B, int units, const float* alpha,
                   const float* gamma, const float* stats, float* g_x, float* g_alpha,
                   float* g_gamma, float* g_beta, rk_stream_t stream);

/* BatchNorm1d (affine, training mode) fused with the ReLU / LeakyReLU that follows it in the DeepFM
 * and BST towers (DeepFM/deepfm.py:100-110, BST/bst.py:203-214): z = act(gamma * batchnorm(x) + beta),
 * act(u) = u > 0 ? u : slope * u (slope 0 = ReLU, 0.01 = the BST tower, 1 = batch norm alone).
 * stats[2, units] = mean | rstd.  Same batch limit and determinism as rk_dice_bn_fwd. */
int rk_bn_act_fwd(const float* x, int64_t B, int units, const float* gamma, const float* beta,
                  float eps, float momentum, float* running_mean, float* running_var,
                  int64_t* num_batches, float slope, float* z, float* stats, rk_stream_t stream);
int rk_bn_act_bwd(const float* x, const float* g_z, int64_t B, int units, const float* gamma,
                  const float* beta, float slope, const float* stats, float* g_x, float* g_gamma,
                  float* g_beta, rk_stream_t stream);

/* ---- row-wise Adam on the touched rows of a table (opt-in: the reference
 *      uses dense optim.Adam, e.g. DeepFM/deepfm.py:226) ---------------------------------------
 * rows[n]: distinct row numbers, grads[n, D]: their summed gradient rows (a coalesced sparse
 * gradient: RowShardedEmbedding, nn.Embedding(sparse=True)).  torch.optim.SparseAdam's update:
 *   m += (g - m)(1 - beta1); v += (g*g - v)(1 - beta2);
 *   w -= lr * sqrt(1 - beta2^step) / (1 - beta1^step) * m / (sqrt(v) + eps)
 * on those rows of weight / exp_avg / exp_avg_sq ([V, D] each) only; step >= 1 is the global step. */
int rk_rowwise_adam(float* weight, float* exp_avg, float* exp_avg_sq, const int64_t* rows,
                    const float* grads, int64_t n, int D, int64_t V, float lr, float beta1,
                    float beta2, float eps, int64_t step, int32_t* err_flag, rk_stream_t stream);
/* The same update for a touched-rows gradient whose live count is on the device (no host sync
 * between the backward and the optimizer): rows[capacity], grads[capacity, D], only the first
 * *count (<= capacity) entries are applied.  This is what the replicated tables' opt-in sparse
 * gradients feed (rk_plan_compact_fields + rk_embgrad_segment_reduce on the rank keys). */
int rk_rowwise_adam_touched(float* weight, float* exp_avg, float* exp_avg_sq, const int64_t* rows,
                            const float* grads, const int64_t* count, int64_t capacity, int D,
                            int64_t V, float lr, float beta1, float beta2, float eps, int64_t step,
                            int32_t* err_flag, rk_stream_t stream);

/* ---- row-sharded table (BASELINE config 5 "scaled": BST feedid table of 1e8 rows block-
 *      partitioned by row over the ranks; the reference itself is single-process) ------------
 * Bookkeeping around the two all-to-alls (indices out / rows back, mirrored for gradients):
 *   rk_shard_owner : owner[i] = idx[i] / rows_per_rank (out-of-range indices -> row 0 + err flag)
 *   rk_shard_route : given the stable owner-sorted order of owner[] from rk_plan_build
 *                    (sorted_owner, perm), write send_local[i] = owner-local row of the i-th sorted
 *                    request, inv[p] = sorted slot of original position p, counts[w] = requests
 *                    for rank w (W <= 64)
 *   rk_plan_compact: for an all-live single-field plan, rank_keys[i] = number of distinct rows
 *                    before sorted position i, uniq_rows[r] = the r-th distinct row, *n_uniq =
 *                    their count; rk_embgrad_segment_reduce run on rank_keys (rows = *n_uniq)
 *                    then yields the compact [n_uniq, D] gradient of a sparse update. */
int rk_shard_owner(const int64_t* idx, int64_t n, int64_t rows_total, int64_t rows_per_rank,
                   int64_t* owner, int32_t* err_flag, rk_stream_t stream);
int rk_shard_route(const int64_t* idx, const uint32_t* sorted_owner, const uint32_t* perm, int64_t n,
                   int64_t rows_total, int64_t rows_per_rank, int W, int64_t* send_local, int64_t* inv,
                   int64_t* counts, rk_stream_t stream);
int rk_plan_compact(const uint32_t* sorted_keys, int64_t n, int64_t rows, uint32_t* rank_keys,
                    int64_t* uniq_rows, int64_t* n_uniq, rk_stream_t stream);
/* The multi-field form, for the touched-rows gradients of replicated tables:
 * sorted_keys is a whole rk_plan_build output over F fields (field f: n[f] keys, rows[f] table
 * rows).  Field f's slice of rank_keys receives keys in the layout rk_embgrad_segment_reduce
 * expects for tables of cap[f] rows (cap[f] >= the number of distinct rows; dead positions are
 * parked on cap[f]), uniq_rows its distinct rows (slice start = sum of n[g], g < f) and n_uniq[f]
 * their count.  One launch, one CTA per field, no host synchronisation. */
int rk_plan_compact_fields(const uint32_t* sorted_keys, const int64_t* n, const int64_t* rows,
                           const int64_t* cap, int F, uint32_t* rank_keys, int64_t* uniq_rows,
                           int64_t* n_uniq, rk_stream_t stream);

/* ---- DeepCrossing residual units (residual_unit + loop, DeepCrossing/deepcrossing.py:25-42,
 *      148-159), fused with the gather + concat ----------------------------------------------
 * units: n_units packed per-call weight blocks of rk_resunit_pack_floats(d, H) floats each,
 *   [W1^T d x Hp][b1 Hp][W2^T Hp x dp][b2 dp][W2 d x Hp][W1 Hp x dp], Hp = roundup(H,8),
 *   dp = roundup(d,4), zero padded (W1 [H,d], W2 [d,H] as nn.Linear stores them).
 * nets[(n_units+1), B, d]: nets[0] = the concat row, nets[i+1] = output of unit i (the last one
 * is the result; all are kept for the backward).  bwd: g_x0[B,d] = d(loss)/d(nets[0]). */
int rk_resunit_pack_floats(int d, int H);
int rk_resunits_fwd(const rk_field_t* fields, int F, const float* dense, int n_dense,
                    const float* units, int n_units, int H, int64_t B, float* nets,
                    int32_t* err_flag, rk_stream_t stream);
int rk_resunits_bwd(const float* nets, const float* units, int n_units, int H, int d, int64_t B,
                    const float* g_out, float* g_x0, rk_stream_t stream);

/* ---- BST transformer block (BSTTransformer.forward BST/bst.py:66-91; gather :224; pooling
 *      :238-241), d_model = 16, nhead in {1,2,4,8,16} ------------------------------------------
 * Input rows x[b,t,:] come either from a table through idx[B,T] (first block: the feedid
 * embedding) or from x_in[B,T,16] (a previous block's output).  Keys t >= seq_len[b] are masked
 * with -inf (a length-0 sample yields NaN, as in the reference).  y_out[B,T,16] and/or the
 * pooled row sum_t y[b,t,:] (divided by seq_len[b] when pool_mean) written to
 * pool_out[b*pool_ld + 0..15] are produced.
 * Dropout (BST/bst.py:86 on w_o's output, :62 inside the FFN, :90 on the FFN output): dropout_p > 0
 * applies inverted dropout at the three sites with keep-bits that are a pure function of
 * (rng[0] = seed, rng[1] = offset, row b*T+t, site) — Philox4x32-10, 16-bit uniforms, an element is
 * dropped when its uniform < round(p*65536).  The backward regenerates the same bits from the same
 * rng values; the caller advances the offset between forwards.  The reference's own draws
 * (torch generators) cannot be replayed, so parity is defined at dropout_p = 0. */
typedef struct rk_bst_block {
    const float* pos;                                  /* position_embedding.weight [max_len,16] */
    const float *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo; /* w_q/w_k/w_v/w_o .weight [16,16], .bias [16] */
    const float *ln1_g, *ln1_b;                        /* norm1 */
    const float *w1, *b1, *w2, *b2;                    /* ffn.0, ffn.3 */
    const float *ln2_g, *ln2_b;                        /* norm2 */
    const uint64_t* rng;                               /* device [seed, offset]; may be NULL when dropout_p == 0 */
    float dropout_p;                                   /* 0 in eval mode */
    int32_t precision;                                 /* RK_BST_FP32: fp32 SIMT block (1e-5 parity); RK_BST_BF16_TENSOR: projections / FFN / weight gradients on tcgen05 (2e-2 bar) */
} rk_bst_block_t;
#define RK_BST_FP32 0
#define RK_BST_BF16_TENSOR 1

int rk_bst_grad_floats(int T);            /* T*16 + 6*256 + 10*16 */
int rk_bst_bwd_ctas(int64_t B, int T, int precision);
int rk_bst_block_fwd(const rk_bst_block_t* blk, int nhead, const float* table, const int64_t* idx,
                     int64_t table_rows, const float* x_in, const int64_t* seq_len, int64_t B, int T,
                     float* y_out, float* pool_out, int pool_ld, int pool_mean, int32_t* err_flag,
                     rk_stream_t stream);
/* g_y[B,T,16] and/or g_pool (row b at g_pool + b*g_pool_ld) -> g_x[B,T,16] and the gradients of the
 * block's registered tensors in g_params, laid out
 *   [pos T*16][wq 256][bq 16][wk][bk][wv][bv][wo][bo][ln1_g][ln1_b][w1][b1][w2][b2][ln2_g][ln2_b];
 * partials: scratch of rk_bst_bwd_ctas(B,T,blk->precision) * rk_bst_grad_floats(T) floats. */
int rk_bst_block_bwd(const rk_bst_block_t* blk, int nhead, const float* table, const int64_t* idx,
                     int64_t table_rows, const float* x_in, const int64_t* seq_len, int64_t B, int T,
                     const float* g_y, const float* g_pool, int g_pool_ld, int pool_mean, float* g_x,
                     float* g_params, float* partials, int n_ctas, int32_t* err_flag,
                     rk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RANK_B200_H */
