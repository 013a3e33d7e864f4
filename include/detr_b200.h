/*
 * detr_b200.h -- C ABI of the B200-native DETR training hot path (libdetr_b200.so).
 *
 * The reference (anenbergb/DETR-object-detection) has no FFI layer: its hot path is reached through
 * Python class identity.  Every entry point below names the reference interface it
 * replaces (file:line under /root/reference).  Conventions:
 *   - plain pointers and sizes only, no torch types; all pointers are DEVICE pointers unless marked host;
 *   - every launcher enqueues on `stream` (a cudaStream_t passed as void*), never synchronises, never
 *     allocates, is re-entrant and graph-capturable;
 *   - return 0 on success, non-zero on a configuration/launch error (text via detr_b200_last_error);
 *   - data-dependent faults (degenerate boxes, NaN costs, infeasible assignment) are reported by OR-ing
 *     bits into a caller-owned device int32 `status` word, checked by the host at its next sync point.
 *
 * Packed ground truth ("CSR targets"): the reference passes List[Tensor] per image
 * (detr/matcher.py:44-46, detr/loss.py:217); here the lists are concatenated:
 *   gt_labels int64[sumM], gt_boxes float[sumM*4] (XYXY, normalised), gt_off int32[B+1] (prefix of M_b),
 *   match_off int32[B+1] (prefix of n_b = min(Q, M_b)).
 * A "problem" is one (image b, decoder layer l) pair, p = b*L + l.
 *   cost matrix of p : float[Q*M_b] row-major (query, gt) at  cost + Q*(L*gt_off[b] + l*M_b)
 *   matches of p     : n_b pairs at  idx_q/idx_gt + L*match_off[b] + l*n_b   (idx_q ascending)
 */
#ifndef DETR_B200_H
#define DETR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DETR_B200_ABI_VERSION 6

/* status bits (device-side, sticky) -- mirror the reference's failure modes */
#define DETR_ST_DEGENERATE_BOX 1 /* AssertionError at detr/utils.py:87-88 */
#define DETR_ST_INVALID_COST 2   /* SciPy ValueError "matrix contains invalid numeric entries" */
#define DETR_ST_INFEASIBLE 4     /* SciPy ValueError "cost matrix is infeasible" */
#define DETR_ST_BAD_LABEL 8      /* label outside [0,K) (PyTorch would raise an index error) */

/* ---- library ---------------------------------------------------------------------------------- */
int detr_b200_abi_version(void);
/* copies the calling thread's last error text into buf (NUL terminated); returns its length */
int detr_b200_last_error(char* buf, int n);
/* 0 if `device` is an sm_100 part this library was built for, non-zero otherwise */
int detr_b200_check_device(int device);
/* bytes of dynamic shared memory the matcher needs for (Q, max_M); <0 if it must use the global path */
int64_t detr_matcher_smem_bytes(int Q, int max_m, int elem_size);

/* ---- HungarianMatcher (detr/matcher.py:40-99) ------------------------------------------------- */
/* Cost matrices only: C = w_bbox*L1(cxcywh) + w_class*(-softmax(logits)[:,label]) + w_giou*(-GIoU)
 * (detr/matcher.py:66-93, detr/utils.py:57-97).  logits: element strides for (image, layer, query),
 * class stride 1; boxes likewise with 4 contiguous floats (cx,cy,w,h). */
int detr_cost_matrix_f32(const float* logits, int64_t lg_sb, int64_t lg_sl, int64_t lg_sq,
                         const float* boxes, int64_t bx_sb, int64_t bx_sl, int64_t bx_sq,
                         const int64_t* gt_labels, const float* gt_boxes, const int32_t* gt_off,
                         int B, int L, int Q, int K, int max_m,
                         float w_class, float w_bbox, float w_giou,
                         float* cost_out, int32_t* status, void* stream);

/* Fused matcher: cost matrix kept in shared memory + assignment, one CTA per problem.
 * Replaces the per-image loop + `.cpu()` + scipy call of detr/matcher.py:69-97 for all images and all
 * decoder layers (detr/loss.py:213-217) in ONE launch.  cost_out may be NULL (not exported).  order_ws: optional
 * int32[B] scratch; when given and the launch has more problems than one wave of CTAs, a tiny pre-kernel ranks the images by
 * descending GT count and the largest assignment problems are started first (results are unaffected). */
int detr_hungarian_match_f32(const float* logits, int64_t lg_sb, int64_t lg_sl, int64_t lg_sq,
                             const float* boxes, int64_t bx_sb, int64_t bx_sl, int64_t bx_sq,
                             const int64_t* gt_labels, const float* gt_boxes, const int32_t* gt_off,
                             const int32_t* match_off, int B, int L, int Q, int K, int max_m,
                             float w_class, float w_bbox, float w_giou,
                             float* cost_out, int64_t* idx_q, int64_t* idx_gt, int32_t* status,
                             int32_t* order_ws, void* stream);

/* Batched rectangular linear-sum assignment on caller-provided cost matrices: the drop-in for
 * scipy.optimize.linear_sum_assignment at detr/matcher.py:94 (bit-exact indices).
 * Problem p: nr[p] x nc[p] row-major at cost + cost_off[p]; writes min(nr,nc) pairs at out_off[p]
 * (rows ascending).  max_nr/max_nc: host-known maxima (size the launch). */
int detr_lsap_f32(const float* cost, const int64_t* cost_off, const int32_t* nr, const int32_t* nc,
                  int n_problems, int max_nr, int max_nc, const int64_t* out_off, int64_t* rows_out,
                  int64_t* cols_out, int32_t* status, void* stream);
int detr_lsap_f64(const double* cost, const int64_t* cost_off, const int32_t* nr, const int32_t* nc,
                  int n_problems, int max_nr, int max_nc, const int64_t* out_off, int64_t* rows_out,
                  int64_t* cols_out, int32_t* status, void* stream);

/* ---- SetCriterion (detr/loss.py:57-164, 198-231) ---------------------------------------------- */
/* Forward, all layers at once.  Outputs:
 *   losses float[L*5] = {loss_label_ce, cardinality_error, loss_l1_bbox, loss_giou, class_error} per layer
 *   (already multiplied by w_ce/w_l1/w_giou as detr/loss.py:91,152-162 does);
 * saved for backward: lse float[B*L*Q], wsum float[L], and the assignment expanded to two dense per-query arrays --
 * tgt int32[B*L*Q] (target class, K-1 = "no object") and tbox float[B*L*Q*4] (matched target box XYXY, NaN x1 =
 * unmatched; 16-byte aligned) -- so that neither the loss kernel nor the backward walks indices or offsets;
 * partials float[B*L*8] is scratch.  class_weight float[K] is SetCriterion.empty_weight (detr/loss.py:53-55).
 * num_boxes: optional device scalar (the all-reduced normaliser); NULL -> max(sumM,1) local
 * as detr/loss.py:142 does. */
int detr_criterion_fwd_f32(const float* logits, int64_t lg_sb, int64_t lg_sl, int64_t lg_sq,
                           const float* boxes, int64_t bx_sb, int64_t bx_sl, int64_t bx_sq,
                           const int64_t* gt_labels, const float* gt_boxes, const int32_t* gt_off,
                           const int32_t* match_off, const int64_t* idx_q, const int64_t* idx_gt,
                           const float* class_weight, const float* num_boxes,
                           int B, int L, int Q, int K, float w_ce, float w_l1, float w_giou,
                           float* partials, float* lse, int32_t* tgt, float* tbox, float* wsum, float* losses,
                           int32_t* status, void* stream);

/* Backward of the 3 differentiable losses per layer.  grad_losses float[L*5] (same layout as `losses`;
 * columns 1 and 4 ignored).  Writes dense grad_logits float[B*L*Q*K] and grad_boxes float[B*L*Q*4]
 * (both contiguous, (B,L,Q,.) order).  lse / tgt / tbox / wsum are the forward call's outputs; gt_off is read only
 * for the local normaliser max(sum M, 1) when num_boxes is NULL. */
int detr_criterion_bwd_f32(const float* grad_losses,
                           const float* logits, int64_t lg_sb, int64_t lg_sl, int64_t lg_sq,
                           const float* boxes, int64_t bx_sb, int64_t bx_sl, int64_t bx_sq,
                           const int32_t* gt_off, const float* class_weight, const float* num_boxes,
                           const float* lse, const int32_t* tgt, const float* tbox, const float* wsum,
                           int B, int L, int Q, int K, float w_ce, float w_l1, float w_giou,
                           float* grad_logits, float* grad_boxes, void* stream);

/* ---- ScaledDotProductAttention core (detr/model.py:317-352) ------------------------------------ */
/* O = dropout(softmax(Q K^T / sqrt(32) + masks)) V for head_dim 32, bf16 in / bf16 out, fp32 softmax.
 * q (B,L,nh*32), k/v (B,S,nh*32), o (B,L,nh*32): bf16, channel stride 1, element strides (batch, row) given;
 * head h is channels [32h, 32h+32) -- i.e. the layout nn.Linear produces, so the view/transpose/contiguous of
 * detr/model.py:317-319,352 disappear.  lse float[B*nh*L] (natural log, saved for backward).
 * key_padding_mask: (B,S) bytes, non-zero = ignore (detr/model.py:326-330), row stride kpm_sb, may be NULL;
 * attention_mask: (L,S) bytes contiguous, non-zero = ignore (detr/model.py:332-334), may be NULL.
 * dropout_p is quantised to k/32768 (0.1 -> 0.100006; in-kernel counter-based mask; 0 disables, as in eval()).  The mask seed is
 * `seed + *seed_ptr` (seed_ptr: optional DEVICE uint64, so that CUDA-graph replays draw fresh masks).
 * workspace float[detr_attention_fwd_workspace_floats(B,nh,L,S)]: partial results of the (batch, head, query tile) items
 * that the persistent kernel splits between two CTAs (merged by a second small launch). */
int64_t detr_attention_fwd_workspace_floats(int B, int nh, int L, int S);
int detr_attention_fwd_bf16(const void* q, int64_t q_sb, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sl,
                            const void* v, int64_t v_sb, int64_t v_sl, void* o, int64_t o_sb, int64_t o_sl, void* o_lo,
                            float* lse, float* workspace, const uint8_t* key_padding_mask, int64_t kpm_sb,
                            const uint8_t* attention_mask, int B, int nh, int L, int S, float dropout_p,
                            uint64_t seed, const uint64_t* seed_ptr, void* stream);

/* `o_lo` (optional, forward output / backward input, bf16 with o's strides): the part of the fp32 output lost when o is
 * rounded to bf16.  The backward's delta = rowsum(dO * (o + o_lo)) is then exact to ~2^-17: an error of delta is common to all
 * keys of a row, so with nearly uniform attention (dS = P (dP - delta) is a difference of almost equal numbers) it does not
 * average out in dQ / dK the way the reference's per-element bf16 rounding does.
 * Backward of the call above: dq (B,L,C), dk/dv (B,S,C) bf16 from d_o (B,L,C) bf16, the forward's inputs, output
 * `o` and `lse`.  Scratch: delta float[B*nh*L] (rowsum(dO o O)) and dq_partial
 * float[detr_attention_bwd_workspace_floats(B,nh,L,S)] (one fp32 dQ partial per 128-key tile).  Same masks /
 * dropout_p / seed as forward.  Three launches: delta, the fused dK+dV+dQ-partial kernel (CTA per key tile, every
 * score tile recomputed once), the fixed-order dQ reduction; deterministic, no atomics. */
int64_t detr_attention_bwd_workspace_floats(int B, int nh, int L, int S);
int detr_attention_bwd_bf16(const void* q, int64_t q_sb, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sl,
                            const void* v, int64_t v_sb, int64_t v_sl, const void* o, int64_t o_sb, int64_t o_sl, const void* o_lo,
                            const void* d_o, int64_t do_sb, int64_t do_sl, const float* lse, float* delta,
                            float* dq_partial, void* dq, int64_t dq_sb, int64_t dq_sl, void* dk, int64_t dk_sb, int64_t dk_sl,
                            void* dv, int64_t dv_sb, int64_t dv_sl, const uint8_t* key_padding_mask, int64_t kpm_sb,
                            const uint8_t* attention_mask, int B, int nh, int L, int S, float dropout_p,
                            uint64_t seed, const uint64_t* seed_ptr, void* stream);

/* ---- transformer block row operations ------------------------------------------------------------ */
/* Bias gradient of nn.Linear (backward of detr/model.py:312-314,354,405-411): out[n] = sum_m g[m][n], g bf16 (M,N)
 * row stride ld.  partial float[detr_colsum_chunks(M,N) * N] is scratch; counters: 64 uint32 that are zero on entry
 * (the kernel leaves them zero).  One launch, deterministic. */
int detr_colsum_chunks(int M, int N);
int detr_colsum_bf16(const void* g, int64_t ld, int M, int N, float* partial, float* out, uint32_t* counters, void* stream);

/* Fused pre-LN prologue (detr/model.py:221-222,173-174,177-178,224,182,209,148): y = LayerNorm(x)*gamma+beta and
 * y2 = y + addend (the positional / query embedding added to the attention's query and key inputs only) in one pass,
 * written in the consumer GEMM's dtype.  dtype codes: 0 float32, 1 bfloat16.  x (rows,C) row stride x_ld; addend row
 * of (batch b, row r) at addend + b*add_sb + r*add_sr (so a (Q,C) embedding broadcasts with add_sb = 0); y / y2
 * (rows,C) contiguous, either may be NULL; mean/rstd float[rows] saved for backward.  C % 32 == 0, C <= 1024. */
int detr_layernorm_grid(int rows);
int detr_layernorm_fwd(const void* x, int x_dtype, int64_t x_ld, const float* gamma, const float* beta,
                       const void* addend, int add_dtype, int64_t add_sb, int64_t add_sr, int rows_per_batch,
                       void* y, void* y2, int out_dtype, float* mean, float* rstd, int rows, int C, float eps, void* stream);
/* Backward: dy / dy2 are the gradients of y / y2 (dtype g_dtype, contiguous rows, either may be NULL); dres (x's dtype,
 * contiguous rows, may be NULL) is the gradient that reaches x through the block's residual add (x + f(LN(x)),
 * detr/model.py:223-224,176-182) and is added to dx in the same pass; dx has x's dtype, contiguous; dgamma/dbeta float[C];
 * partial float[detr_layernorm_grid(rows)*2*C] scratch; counters as colsum. */
int detr_layernorm_bwd(const void* dy, const void* dy2, int g_dtype, const void* dres, const void* x, int x_dtype, int64_t x_ld,
                       const float* gamma, const float* mean, const float* rstd, void* dx, float* partial,
                       float* dgamma, float* dbeta, uint32_t* counters, int rows, int C, void* stream);

/* The same backward pass with a third product: dx is also the gradient that reaches the PREVIOUS block's tail
 * (x = res + dropout(z), detr/model.py:223-224,354-355,410), so its masked bf16 form dz = mask(dx)/(1-p) (the operand of that
 * tail's input- and weight-gradient GEMMs) and dbias[n] = sum_m dz[m][n] are written here instead of by a separate
 * detr_epilogue_bwd(mode 0) launch.  C = 256; partial float[detr_layernorm_grid(rows)*3*C]; (dropout_p, seed, seed_ptr) are the
 * TAIL's forward values. */
int detr_layernorm_bwd_tail(const void* dy, const void* dy2, int g_dtype, const void* dres, const void* x, int x_dtype, int64_t x_ld,
                            const float* gamma, const float* mean, const float* rstd, void* dx, float* partial,
                            float* dgamma, float* dbeta, uint32_t* counters, int rows, int C, void* dz, float* dbias,
                            float dropout_p, uint64_t seed, const uint64_t* seed_ptr, void* stream);

/* Both calls above launch two kernels: the row pass (dx, dz, per-CTA partial sums) and the fold of the partials into the
 * PARAMETER gradients dgamma / dbeta (/ dbias).  Called with dgamma == NULL they launch only the row pass; the caller then
 * folds with this entry point -- on another stream if it likes: nothing on the critical path of the backward pass reads
 * parameter gradients.  dbias: NULL unless the row pass was detr_layernorm_bwd_tail. */
int detr_layernorm_bwd_fold(const float* partial, int rows, int C, float* dgamma, float* dbeta, float* dbias, void* stream);

/* Block epilogues of the pre-LN layers, one pass each (detr/model.py:223-224, 176-182, FFN 405-411).
 * mode 0: out(x's dtype) = x + dropout(y)  -- residual add after the attention output / second FFN projection;
 * mode 1: out(bf16) = dropout(gelu_tanh(y)) -- between the FFN projections (x unused).
 * y bf16 (M,N) contiguous = the producing Linear's output; x_dtype / g_dtype: 0 float32, 1 bfloat16.  The dropout mask
 * is counter-based (15 bits per element, p quantised to k/32768, seed + *seed_ptr as in the attention kernels) and is
 * regenerated by the backward call, which also returns the producing Linear's bias gradient:
 *   dy(bf16) = mask(g)/(1-p) [* gelu'(y) in mode 1],  db[n] = sum_m dy[m][n]
 * partial float[detr_epilogue_chunks(M,N) * N] is scratch; counters as for detr_colsum_bf16.  N % 8 == 0. */
int detr_epilogue_fwd(int mode, const void* x, int x_dtype, const void* y, void* out, int M, int N, float dropout_p,
                      uint64_t seed, const uint64_t* seed_ptr, void* stream);
int detr_epilogue_chunks(int M, int N);
int detr_epilogue_bwd(int mode, const void* g, int g_dtype, const void* y, void* dy, float* partial, float* db,
                      uint32_t* counters, int M, int N, float dropout_p, uint64_t seed, const uint64_t* seed_ptr, void* stream);

/* Backward entry of the prediction heads (detr/model.py:92-93 `class_embedding(.)`, `bbox_embedding(.).sigmoid()`): the criterion's
 * fp32 gradients d_logits [rows][K], d_boxes [rows][4] become the bf16 operands of the heads' gradient GEMMs, in one pass:
 * dl16 [rows][ld_l] = bf16(d_logits) zero-padded to ld_l columns; dz16 [rows][ld_z] = bf16(d_boxes * b * (1 - b)) (b = `boxes`, the
 * sigmoid output) zero-padded to ld_z columns. */
int detr_heads_grad_prep(const float* d_logits, int K, const float* d_boxes, const float* boxes, void* dl16, int ld_l,
                         void* dz16, int ld_z, int rows, void* stream);

/* ---- tcgen05 GEMMs of the transformer's nn.Linear layers (detr/model.py:312-314,354 attention projections;
 *      :405-411 FFN) with the surrounding row work fused in (csrc/gemm.cu) -------------------------------------- */
/* C[M][N] = epilogue(A[M][K] . B^T): a, b bf16 row-major with row strides lda / ldb (elements); b_kn = 0: b is [N][K]
 * (an nn.Linear weight), b_kn = 1: b is [K][N] (C = A . B, the input-gradient form dX = dY . W).  N % 32 == 0, K % 64 == 0;
 * the fp32 outputs of epilogues 0 and 4 with b_kn = 0 also take N % 4 == 0 (the prediction heads' 92 / 4 columns): bias must
 * then be readable up to the next multiple of 64 entries.
 * epilogue 0: out = acc + bias                                   (bias fp32 [N] or NULL; out_dtype 0 fp32 / 1 bf16)
 *          1: aux = bf16(acc + bias); out = dropout(gelu_tanh(aux))          (detr/model.py:405-408; out bf16)
 *          2: out = res + dropout(acc + bias)                    (detr/model.py:223-224,354-355,410; res has out's dtype)
 *          3: out = acc * dropout_mask/(1-p) * gelu_tanh'(aux)   (backward of 1; no bias)
 *          4: out = sigmoid(acc + bias)                          (box head, detr/model.py:93; out fp32, b_kn = 0)
 * dropout as detr_epilogue_*: counter-based, regenerated by the backward launches from (seed + *seed_ptr). */
int detr_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, int b_kn, int M, int N, int K, int epilogue,
                   const float* bias, void* out, int out_dtype, int64_t ldo, void* aux, int64_t ld_aux, const void* res,
                   int64_t ld_res, float dropout_p, uint64_t seed, const uint64_t* seed_ptr, void* stream);
/* out[M][N] (bf16) = epilogue((LayerNorm(x) [+ addend]) . w[N][256]^T): the pre-LN LayerNorm and the "+ positional /
 * query embedding" of detr/model.py:221-224,173-182 run as the PROLOGUE of the projections that consume them.  x fp32 /
 * bf16 (x_dtype 0 / 1) [M][256] row stride ldx; output columns < n_pos_end (a multiple of 128) are computed from
 * LN(x) + addend, the others from LN(x) (q|k|v in one launch: n_pos_end = 512).  addend fp32, row of flattened row m at
 * (m / rows_per_batch) * add_sb + (m % rows_per_batch) * add_sr, in one of two forms: a dense (M, 256) block
 * (add_sb == rows_per_batch * add_sr) or a broadcast of rows_per_batch >= 8 rows over the batch (add_sb == 0).  epilogue 0 or 1 as above.  Optional outputs for the
 * backward pass: a_plain / a_pos bf16 [M][256] (the two normalised operands), mean / rstd float[M]. */
int detr_gemm_ln_bf16(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, float eps,
                      const float* addend, int64_t add_sb, int64_t add_sr, int rows_per_batch, int n_pos_end,
                      const void* w, int64_t ldw, int M, int N, int epilogue, const float* bias, void* out, int64_t ldo,
                      void* aux, int64_t ld_aux, void* a_plain, void* a_pos, float* mean, float* rstd, float dropout_p,
                      uint64_t seed, const uint64_t* seed_ptr, void* stream);
/* Host-only: the (rows per row block, column groups per row block) detr_gemm_ln_bf16 uses for an (M, N) problem on `sms` SMs. */
int detr_gemm_ln_partition(int M, int N, int gelu, int sms, int* rows_per_cta, int* groups);

/* Weight and bias gradients of nn.Linear: dw[N][K] (fp32, contiguous) = dy[M][N]^T . x[M][K], db[N] (optional) = column
 * sums of dy; dy, x bf16 row-major.  Rows n >= n_switch (a multiple of 128) of dw are computed from x1 instead of x0
 * (fused q|k|v projection: q/k rows from LN(x)+pos, v rows from LN(x)); x1 may be NULL.  Split over M with fp32
 * partials folded in a fixed order (deterministic); workspace: detr_gemm_wgrad_workspace_floats(M, N, K) floats. */
int64_t detr_gemm_wgrad_workspace_floats(int M, int N, int K);
int detr_gemm_wgrad_bf16(const void* dy, int64_t ld_dy, const void* x0, int64_t ld_x0, const void* x1, int64_t ld_x1, int n_switch,
                         int M, int N, int K, float* dw, float* db, float* workspace, void* stream);

/* ---- caller-side glue: frozen-BatchNorm fold of the backbone weights (detr/model.py:427-438) ----------------- */
/* dst[o][i][hw] = (out dtype)(src[o][i][hw] * scale[o]) for up to 64 (O, I, H*W) tensors in ONE launch; element strides
 * are given per tensor for (o, i, hw) on both sides (hw must be flattenable: stride_h == W * stride_w).  Used forward
 * (fp32 conv weights -> bf16 folded weights) and backward (bf16 weight gradients -> fp32 parameter gradients).
 * dtype codes: 0 float32, 1 bfloat16.  The table is read on the host during the call. */
#define DETR_FOLD_MAX_TENSORS 64
typedef struct DetrFoldTable {
    const void* src[DETR_FOLD_MAX_TENSORS];
    void* dst[DETR_FOLD_MAX_TENSORS];
    const float* scale[DETR_FOLD_MAX_TENSORS];
    int O[DETR_FOLD_MAX_TENSORS], I[DETR_FOLD_MAX_TENSORS], HW[DETR_FOLD_MAX_TENSORS];
    int src_stride[DETR_FOLD_MAX_TENSORS][3];
    int dst_stride[DETR_FOLD_MAX_TENSORS][3];
    int n;
} DetrFoldTable;
int detr_scale_cast_multi(const DetrFoldTable* table, int in_dtype, int out_dtype, void* stream);

/* Channels-last (B,H,W,C physical) bf16 max pooling, kernel 3 / stride 2 / padding 1 -- the ResNet stem's `maxpool`
 * (torchvision, called from detr/model.py:437).  idx (B,Ho,Wo,C) uint8 receives the window position (0..8) of the
 * first maximum (ATen's tie rule) for the backward gather.  C % 8 == 0; Ho = detr_maxpool3x3s2_out(H). */
int detr_maxpool3x3s2_out(int n);
int detr_maxpool3x3s2_fwd_bf16(const void* x, void* y, uint8_t* idx, int B, int H, int W, int C, void* stream);
int detr_maxpool3x3s2_bwd_bf16(const void* dy, const uint8_t* idx, void* dx, int B, int H, int W, int C, void* stream);
/* out = (a + b) * (x > 0) on dense bf16 buffers of n elements (n % 8 == 0): residual-gradient accumulation of a ResNet
 * bottleneck fused with the previous block's ReLU backward (harness glue). */
int detr_add_relu_mask_bf16(const void* a, const void* b, const void* x, void* out, long long n, void* stream);
/* Stem input of the ResNet harness in one pass: zero-pad by 3, 2x2 space-to-depth, cast to bf16, pad channels.
 * x fp32 (B,Cin,H,W) with element strides (sb,sc,sh,sw); out bf16 dense (B,(H+6)/2,(W+6)/2,Cout) i.e. channels_last,
 * out[b][i][j][c*4+r*2+s] = x[b][c][2i+r-3][2j+s-3].  H, W even; Cout % 8 == 0, Cout >= 4*Cin. */
int detr_stem_s2d_bf16(const float* x, long long sb, long long sc, long long sh, long long sw, int B, int Cin, int H, int W,
                       void* out, int Cout, void* stream);



/* Sine positional encoding + padding mask of DETR.forward (detr/position_encoding.py:5-97, detr/model.py:96-114) in one
 * launch, token-major: pos float[B][H'*W'][2F] (channel order of the reference: F y-channels then F x-channels, sin/cos
 * interleaved), mask uint8[B][H'*W'] (1 = the reference's bottom-right padding corner), heights/widths DEVICE int32[B]
 * image sizes before padding; scale = backbone stride.  mask may be NULL. */
int detr_positional_encoding_f32(const int32_t* heights, const int32_t* widths, int B, int embed_h, int embed_w, int scale,
                                 int num_pos_feats, float temperature, float* pos, uint8_t* mask, void* stream);

/* Caller-side glue of the benchmark harness: clip_grad_norm_(max_norm) + AdamW.step (detr/train.py:265-267) on FLAT fp32
 * buffers.  detr_sumsq_f32: out[0] = sum g^2 (partial float[detr_sumsq_grid(n)] scratch, counter: one zeroed uint32,
 * left zero).  detr_adamw_clip_f32 updates p, m, v in place with g scaled by grad_div * min(1, max_norm /
 * (grad_div * sqrt(*sumsq) + 1e-6)) (sumsq NULL or max_norm <= 0: no clipping); `step` (1-based step count, for the bias
 * correction) and `lr` are DEVICE floats, so a captured CUDA graph follows the step count and the LR schedule.
 * Fault gating: detr_sumsq_f32 adds 1 to *step (optional) only when the sum is finite, and detr_adamw_clip_f32 leaves p, m, v
 * untouched when *sumsq is NaN / inf -- a batch whose losses were poisoned by the device status word (degenerate box, NaN
 * cost) is skipped instead of destroying the weights; the reference raises at that point (detr/utils.py:87-88). */
int detr_sumsq_grid(long long n);
int detr_sumsq_f32(const float* g, long long n, float* partial, float* out, uint32_t* counter, float* step, void* stream);
int detr_adamw_clip_f32(float* p, const float* g, float* m, float* v, long long n, const float* lr, float beta1, float beta2, float eps,
                        float weight_decay, const float* step, const float* sumsq, float max_norm, float grad_div, void* stream);






#ifdef __cplusplus
}
#endif
#endif /* DETR_B200_H */
