/*
 * unirec_b200 - C ABI of the B200 (sm_100a) kernels behind UniRec's nested Q-Former
 * encode-and-rank path.
 *
 * The reference (ulab-uiuc/UniRec) is pure PyTorch: it has no FFI of its own, so the "binding a
 * maintainer would add" is a ctypes stub.  Every entry point takes plain device
 * pointers, sizes and a CUDA stream handle (cudaStream_t passed as void*; NULL = legacy default
 * stream) - no torch types cross this boundary.  Pointers are BORROWED for the duration of the call;
 * the library allocates nothing it returns, never synchronises the stream and never falls back to
 * the CPU.  Every function returns 0 on success or a non-zero UNIREC_ERR_* code, in which case
 * unirec_last_error() describes the failure (thread-local string).
 *
 * Conventions: bf16 = 16-bit bfloat16 storage; all matrices are row-major with an explicit row
 * stride ("ld", in elements); rows must be 16-byte aligned.  `out_fp32` selects fp32 (1) or bf16 (0)
 * output storage.  Reference file:line citations are relative to the reference repository root.
 */
#ifndef UNIREC_B200_H_
#define UNIREC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNIREC_B200_ABI_VERSION 3

#define UNIREC_OK 0
#define UNIREC_ERR_BAD_ARG 1
#define UNIREC_ERR_CUDA 2
#define UNIREC_ERR_TENSORMAP 3
#define UNIREC_ERR_NO_DEVICE 4

/* GEMM epilogues */
#define UNIREC_EPI_BIAS 0          /* out = A W^T + b                                  */
#define UNIREC_EPI_BIAS_GELU 1     /* out = gelu_erf(A W^T + b)   (models/qformer.py:359-361) */
#define UNIREC_EPI_BIAS_RESIDUAL 2 /* out = A W^T + b + residual  (pre-LayerNorm sum, :286-288, :372-374) */

int unirec_abi_version(void);
const char* unirec_last_error(void);

/* Number of kernels this library has launched in this process (all entry points), for bench.py's
 * "gpu_launches" claim. */
int64_t unirec_launch_count(void);

/* nn.Linear on the path: out[M,N] = epilogue(A[M,K] @ W[N,K]^T + bias[N]).
 * Replaces every nn.Linear of the path: Q/K/V projections models/qformer.py:185-198, attention output
 * dense :286, FFN up :359 and down :372, heads models/qformer_utils.py:50,53 and
 * training/user_qformer_training.py:38-43.  tcgen05/TMEM/TMA kernel.
 * A, W bf16; bias fp32 or NULL; residual bf16 (epilogue 2), residual row = row % res_row_mod when
 * res_row_mod > 0 (batch-invariant residual).  K % 64 == 0, N % 8 == 0.
 * block_n: 0 = auto; 128 or 256 = single-CTA kernel with that tile width; 2 = CTA-pair kernel
 * (tcgen05 cta_group::2, 256 x 256 tile per SM pair, TMA-store epilogue; needs bf16 out, N % 256 == 0,
 * res_row_mod == 0).  max_ctas: 0 = one per SM. */
int unirec_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                       const void* residual, int64_t ldr, int64_t res_row_mod,
                       void* out, int64_t ldo, int out_fp32,
                       int64_t M, int64_t N, int64_t K, int epilogue, int block_n, int max_ctas, void* stream);

/* nn.Linear with the LayerNorms around it folded in (models/qformer.py:285-289, 371-375: h = LayerNorm(dropout(dense(x)) + input);
 * the LayerNorm output h is never written).  CTA-pair kernel: bf16 output, N % 256 == 0, K % 64 == 0, bias required.
 *   stats_out    fp32 [M, unirec_linear_ln_stats_parts(N), 2]: the call WRITES, per row, one (sum, sum of squares) partial of
 *                the bf16 values it stores per column piece of an epilogue warp (no atomics: consumers add the partials in
 *                index order, so results do not depend on the launch geometry) - this GEMM is the producer of a LayerNorm
 *                input `pre`;
 *   ln_in_stats  fp32 [M, ln_parts, 2] + ln_in_c fp32 [N]: A is such a `pre`; W must already be scaled by the LayerNorm's
 *                gamma, bias = b + W beta, ln_in_c[n] = sum_k W'[n, k]: out = rstd (A W'^T) - mu rstd c + bias;
 *   ln_res_stats fp32 [M, ln_parts, 2] + ln_res_gamma / ln_res_beta fp32 [N] (epilogue UNIREC_EPI_BIAS_RESIDUAL): the
 *                residual tensor is such a `pre` and enters as LayerNorm(residual) = (x - mu) rstd gamma + beta.
 * Any of the three groups may be NULL (plain behaviour of unirec_linear_bf16).  ln_hidden = width the statistics cover,
 * ln_parts = partials per row in ln_in_stats / ln_res_stats (unirec_linear_ln_stats_parts(ln_hidden) when this entry point
 * wrote them). */
int64_t unirec_linear_ln_stats_parts(int64_t N);
int unirec_linear_ln_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* residual,
                          int64_t ldr, void* out, int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue,
                          const float* ln_in_stats, const float* ln_in_c, const float* ln_res_stats,
                          const float* ln_res_gamma, const float* ln_res_beta, float* stats_out, int ln_parts,
                          float ln_eps, int64_t ln_hidden, void* stream);

/* nn.LayerNorm over the last dim (models/qformer.py:104, :288, :374; user head
 * training/user_qformer_training.py:41): out = LN(x [+ residual]) * gamma + beta.
 * x fp32 (x_fp32=1) or bf16; x row = row % in_row_mod when in_row_mod > 0; residual bf16 or NULL. */
int unirec_layernorm(const void* x, int x_fp32, int64_t ldx, int64_t in_row_mod,
                     const void* residual, int64_t ldres,
                     const float* gamma, const float* beta, float eps,
                     void* out, int out_fp32, int64_t ldo, int64_t rows, int64_t H, void* stream);

/* Fused small-query multi-head attention, head_dim 64 (models/qformer.py:161-167, 205, 244-268).
 * q [batch*nq rows] (q_batch_rows = nq, or 0 when the same queries serve every batch element),
 * k, v [batch*kv_batch_rows rows]; head h occupies columns [h*64, (h+1)*64) of each row;
 * key_mask [batch, nk] fp32 (1 = attend, 0 = masked) or NULL; out [batch*nq, ldo] head-merged. */
int unirec_attention(const void* q, int64_t ldq, int64_t q_batch_rows,
                     const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t kv_batch_rows,
                     const float* key_mask, void* out, int64_t ldo,
                     int64_t batch, int64_t num_heads, int64_t nq, int64_t nk, int64_t head_dim,
                     float scale, void* stream);

/* fp32 -> bf16 (callers hand fp32 field embeddings, models/qformer_utils.py:37). n % 8 == 0. */
int unirec_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream);

/* out[b,:] = mean_t x[b,t,:]  (models/qformer_utils.py:50; training/user_qformer_training.py:60). */
int unirec_mean_tokens(const void* x, int64_t ldx, int64_t B, int64_t T, int64_t H,
                       void* out, int64_t ldo, int out_fp32, void* stream);

/* out[b,f,:] = sum_t Wp[f,t] * rec[b,t,:] + bp[f]  (field_projection, models/qformer_utils.py:54). */
int unirec_field_projection(const void* rec, const float* Wp, const float* bp, void* out, int out_fp32,
                            int64_t B, int64_t T, int64_t F, int64_t E, void* stream);

/* User-sequence builder (models/user_sequence_encoder.py:128-140 + padding of
 * training/user_qformer_training.py:153-161): seq[b, h*Q+q, :] = table[history[b,h], q, :] (+ ctx[b,h,:])
 * + PE[h*Q+q, :] for h < lengths[b], 0 otherwise; mask[b,s] = s < lengths[b]*Q.
 * table bf16 [num_items, Q, D]; history int64 [B, Hmax]; lengths int32 [B]; ctx bf16 [B,Hmax,D] or NULL;
 * pe_table fp32 [Hmax*Q, D] from unirec_positional_encoding, or NULL (PE evaluated in the kernel: slower, same values);
 * seq bf16 [B, Hmax*Q, D]; mask fp32 [B, Hmax*Q].  A history id outside [0, num_items) contributes a zero token row
 * (the slot keeps ctx + PE and stays attended) - the same as unirec_linear_gather_bf16; never an out-of-bounds read. */
int unirec_build_user_sequence(const void* table, int64_t num_items, const int64_t* history,
                               const int32_t* lengths, const void* ctx, const float* pe_table, void* seq, float* mask,
                               int64_t B, int64_t Hmax, int64_t Q, int64_t D, void* stream);

/* pe_table[p, 2i] = sin(p * w_i), pe_table[p, 2i+1] = cos(p * w_i), w_i = exp(-(2i) ln(10000) / D)
 * (PositionalEncoding buffer, models/user_sequence_encoder.py:20-24); fp32 [S, D], D even. */
int unirec_positional_encoding(float* pe_table, int64_t S, int64_t D, void* stream);

/* inv[r] = 1 / max(||x_r||_2, eps)  (F.normalize, training/train_item_individual_token_joint.py:405-406,412). */
int unirec_inv_l2_norm(const void* x, int x_fp32, int64_t ldx, float* inv, int64_t rows, int64_t D,
                       float eps, void* stream);

/* Fused cosine scoring + top-k (training/train_item_individual_token_joint.py:405-415 at scale):
 * scores[b,n] = <users[b], cands[n]> * user_inv[b] * cand_inv[n]; returns for each user the k best
 * candidates in descending score order without materialising [B, N].
 * users bf16 [B, D]; cands bf16 [N, D]; user_inv fp32 [B]; cand_inv fp32 [N]; index_base is added to
 * every returned index (row-sharded candidate pools); out_scores fp32 [B, k]; out_idx int64 [B, k].
 * workspace: unirec_score_topk_workspace_bytes(B, N, k) bytes of device memory. */
int64_t unirec_score_topk_workspace_bytes(int64_t B, int64_t N, int64_t k);
int unirec_score_topk(const void* users, int64_t ldu, const float* user_inv,
                      const void* cands, int64_t ldc, const float* cand_inv,
                      int64_t B, int64_t N, int64_t D, int64_t k, int64_t index_base,
                      float* out_scores, int64_t* out_idx, void* workspace, int64_t workspace_bytes,
                      void* stream);

/* Merge per-shard top-k lists (after the NCCL all-gather of config 5): in_scores/in_idx
 * [G, B, k] (each list descending) -> out [B, k] descending; ties broken by smaller index. */
int unirec_topk_merge(const float* in_scores, const int64_t* in_idx, int64_t G, int64_t B, int64_t k,
                      float* out_scores, int64_t* out_idx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward pass of the item Q-Former training step (training/item_qformer_training.py:129,
 * loss.backward(); dropout disabled).  Gradients of activations are bf16, of parameters fp32.
 * ------------------------------------------------------------------------------------------- */

/* out[M,N] (+)= sum_k A(m,k) B(n,k), bf16 operands, fp32 accumulation (tcgen05).  Per operand:
 *   a_mn = 0: A stored [M, K] row-major;  a_mn = 1: A stored [K, M] row-major
 *   b_mn = 0: B stored [N, K] row-major;  b_mn = 1: B stored [K, N] row-major
 * dgrad of y = x W^T:  dx = gemm(A = dy, a_mn 0, B = W [N,K_in], b_mn 1);  wgrad: dW = gemm(A = dy [rows,N], a_mn 1,
 * B = x [rows,K_in], b_mn 1, accumulate 1).  accumulate = 1 needs fp32 out and adds with fp32 atomics; the
 * contraction is then split over ksplit CTAs per tile (0 = auto).  K % 8 == 0, N % 8 == 0, 16-byte aligned rows. */
int unirec_gemm_general(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                        void* out, int64_t ldo, int out_fp32, int accumulate,
                        int64_t M, int64_t N, int64_t K, int ksplit, void* stream);

/* out = gelu_erf(z) / dz = da * gelu_erf'(z), bf16, n % 8 == 0 (models/qformer.py:360 with the pre-activation kept). */
int unirec_gelu_forward(const void* z, void* out, int64_t n, void* stream);
int unirec_gelu_backward(const void* z, const void* da, void* dz, int64_t n, void* stream);

/* out[n] += sum_rows x[row, n]  (bias gradients); x bf16 [rows, N] row stride ld, out fp32 [N] (atomic accumulate). */
int unirec_colsum(const void* x, int64_t ld, int64_t rows, int64_t N, float* out, void* stream);

/* LayerNorm backward over the last dim H <= 1024: x = the LayerNorm INPUT (bf16), dy (+ optional dy2) = gradient of
 * its output, dx bf16, dgamma / dbeta fp32 [H] atomically accumulated (models/qformer.py:104, :288, :374). */
int unirec_layernorm_backward(const void* x, int64_t ldx, const void* dy, int64_t lddy, const void* dy2, int64_t lddy2,
                              const float* gamma, float eps, void* dx, int64_t lddx, float* dgamma, float* dbeta,
                              int64_t rows, int64_t H, void* stream);

/* unirec_layernorm_backward with the two things that always follow it in the reference's blocks
 * "LayerNorm(dropout(dense(x)) + input)" (models/qformer.py:285-289, :371-375) fused in:
 *   dx_drop (may be NULL) = dx o mask(thr16, seed, site) * scale - the gradient of dense(x) behind its dropout
 *                           (thr16 = 0: a plain copy of dx; the mask is applied to the bf16-rounded dx, like
 *                           unirec_dropout_backward on the stored tensor);
 *   dbias (may be NULL, fp32 [H], atomically accumulated) += column sums of dx_drop (of dx when dx_drop is NULL) - the
 *                           gradient of that dense layer's bias. */
int unirec_layernorm_backward_fused(const void* x, int64_t ldx, const void* dy, int64_t lddy, const void* dy2,
                                    int64_t lddy2, const float* gamma, float eps, void* dx, int64_t lddx, float* dgamma,
                                    float* dbeta, int64_t rows, int64_t H, uint32_t thr16, uint64_t seed, uint32_t site,
                                    const uint64_t* seed_offset, void* dx_drop, int64_t lddrop, float* dbias,
                                    void* stream);

/* Backward of unirec_attention for nq <= 64 and nk <= 64 (item self- and cross-attention): dq/dk/dv bf16 with the
 * layouts of q/k/v (row strides lddq/lddk/lddv); probabilities are recomputed from q, k and key_mask. */
int unirec_attention_backward(const void* q, int64_t ldq, int64_t q_batch_rows,
                              const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t kv_batch_rows,
                              const float* key_mask, const void* dout, int64_t lddo,
                              void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                              int64_t batch, int64_t num_heads, int64_t nq, int64_t nk, int64_t head_dim,
                              float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Train-mode dropout (nn.Dropout at models/qformer.py:107 embeddings, :258 attention probabilities, :287 attention
 * output dense, :373 FFN output dense; p = 0.2 item / 0.1 user, models/qformer_utils.py:19,25).  The reference draws
 * masks from torch's global RNG; here a mask bit is a pure function of (seed, site, element) through Philox4x32-10 so
 * that forward, backward and the CPU oracle agree bit for bit (definition: unirec_b200/csrc/dropout.cuh, restated in
 * oracle/dropout_masks.py).  thr16 = round(p * 65536) (0 = off); kept elements are scaled by 65536 / (65536 - thr16).
 * seed_offset: NULL, or a DEVICE pointer to one uint64 that the kernels add to `seed` when they run - a training step
 * captured once into a CUDA graph (unirec_b200/training.py::TrainStepGraph) increments that word inside the graph and
 * so draws fresh masks on every replay although `seed` itself is frozen in the captured launch parameters.
 * ------------------------------------------------------------------------------------------- */

/* unirec_attention with dropout on the probabilities (after the softmax, before the product with V). */
int unirec_attention_dropout(const void* q, int64_t ldq, int64_t q_batch_rows,
                             const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t kv_batch_rows,
                             const float* key_mask, void* out, int64_t ldo,
                             int64_t batch, int64_t num_heads, int64_t nq, int64_t nk, int64_t head_dim,
                             float scale, uint32_t thr16, uint64_t seed, uint32_t site,
                             const uint64_t* seed_offset, void* stream);

/* unirec_attention_backward for a forward pass run with unirec_attention_dropout(thr16, seed, site). */
int unirec_attention_dropout_backward(const void* q, int64_t ldq, int64_t q_batch_rows,
                                      const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t kv_batch_rows,
                                      const float* key_mask, const void* dout, int64_t lddo,
                                      void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                      int64_t batch, int64_t num_heads, int64_t nq, int64_t nk, int64_t head_dim,
                                      float scale, uint32_t thr16, uint64_t seed, uint32_t site,
                             const uint64_t* seed_offset, void* stream);

/* out[r,:] = dropout(x[r % x_row_mod or r,:]) + residual[r,:]  (bf16; residual NULL = no residual; H % 8 == 0):
 * the pre-LayerNorm sum "dropout(dense(x)) + input_tensor" of models/qformer.py:287-288 / :373-374, and the dropped
 * query embeddings of :107 (x_row_mod = number of query tokens, residual NULL). */
int unirec_dropout_add(const void* x, int64_t ldx, int64_t x_row_mod, const void* residual, int64_t ldres,
                       void* out, int64_t ldo, int64_t rows, int64_t H,
                       uint32_t thr16, uint64_t seed, uint32_t site, const uint64_t* seed_offset, void* stream);

/* dx = dy o mask * scale  (gradient of the dropped branch). */
int unirec_dropout_backward(const void* dy, int64_t lddy, void* dx, int64_t lddx, int64_t rows, int64_t H,
                            uint32_t thr16, uint64_t seed, uint32_t site, const uint64_t* seed_offset, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Per-user candidate-LIST scoring - the ranking of the reference's joint trainer:
 * InfoNCELoss.forward (training/train_item_individual_token_joint.py:331-352) and
 * MRREvaluator._compute_batch_mrr (:405-418).  Every user has its own list: one positive row and up to C negatives,
 * either PADDED (cands = [B, C, D] rows b * C + c, mask [B, C] bytes, 0 = padding; the training collate :300-323) or
 * RAGGED (cands = all lists concatenated, offsets [B + 1] row offsets, C = longest list; the validation collate
 * :381-390).  users / pos / cands share one dtype (fp32 = 1: float, else bf16), rows 16-byte aligned, D % 8 == 0.
 * List entry 0 is the positive, entry 1 + c the c-th negative.  F.normalize semantics: x / max(||x||, eps).
 * ------------------------------------------------------------------------------------------- */

/* sims[b, e] = cosine(user b, entry e), -inf for padding; inv_norm[b, e] = 1 / max(||entry||, eps) (may be NULL);
 * both fp32 [B, 1 + C]. */
int unirec_list_scores(const void* users, int64_t ldu, const void* pos, int64_t ldp, const void* cands, int64_t ldc,
                       int fp32, const uint8_t* mask, const int64_t* offsets, int64_t B, int64_t C, int64_t D, float eps,
                       float* sims, float* inv_norm, void* stream);

/* loss[b] = -sims[b,0] / T + logsumexp_e(sims[b,e] / T) over the valid entries (:347-350);
 * rank[b] = 1 + number of valid negatives scoring above the positive (the positive's 1-based position in the
 * descending order of :413-415; MRR = mean(1 / rank)).  Either output may be NULL. */
int unirec_infonce_rank(const float* sims, int64_t B, int64_t C, float temperature, float* loss, int32_t* rank,
                        void* stream);

/* Gradient of sum_b dloss[b] * loss[b]: d_user fp32 [B, D] is ACCUMULATED (zero it first); d_list (may be NULL) fp32
 * [B, 1 + C, D] receives the gradient of every list entry (zero rows for padding). */
int unirec_list_scores_backward(const void* users, int64_t ldu, const void* pos, int64_t ldp, const void* cands,
                                int64_t ldc, int fp32, const uint8_t* mask, const int64_t* offsets, int64_t B, int64_t C,
                                int64_t D, float eps, const float* sims, const float* inv_norm, const float* dloss,
                                float temperature, float* d_user, float* d_list, void* stream);

/* Token injection of the joint model (:160-171): text_embeds[b, s, :] = tokens[b, slot, :] wherever
 * input_ids[b, s] == token_ids[slot]; tokens [B, num_slots, Hd] (num_slots = history items x query tokens per item),
 * text_embeds rows [B * S] with row stride ld_text; fp32 or bf16 on either side (converted like the reference's
 * indexed assignment). */
int unirec_inject_tokens(const int64_t* input_ids, int64_t B, int64_t S, const int64_t* token_ids, int64_t num_slots,
                         const void* tokens, int tokens_fp32, void* text_embeds, int text_fp32, int64_t ld_text,
                         int64_t Hd, void* stream);

/* Backward of unirec_inject_tokens (the reference's overwrite is a differentiable index assignment - the only path by
 * which the joint trainer's loss reaches the item Q-Former, training/train_item_individual_token_joint.py:160-171):
 * d_text [B * S rows, stride ld_text] holds the upstream gradient and is edited IN PLACE (overwritten positions -> 0);
 * d_tokens fp32 [B, num_slots, Hd] is ACCUMULATED (zero it first) with the upstream rows of the positions of each slot. */
int unirec_inject_tokens_backward(const int64_t* input_ids, int64_t B, int64_t S, const int64_t* token_ids,
                                  int64_t num_slots, void* d_text, int text_fp32, int64_t ld_text, float* d_tokens,
                                  int64_t Hd, void* stream);

/* User cross-attention with the K/V projection fused in (models/qformer.py:185-188 key / value
 * projections of the encoder states + :205, :244-268 attention): no K or V is written to memory.
 *   out[u, q, h*64:(h+1)*64] = softmax_k(Q[u,q,h,:] . (x[u,k,:] Wk_h^T) * scale + mask[u,k]) (x[u,k,:] Wv_h^T) + bv_h
 * x bf16 [users * S, K] (row stride ldx); w_packed bf16 [2 H, K], H = num_heads * 64, packed per head PAIR j:
 * rows [256 j, 256 j + 128) = Wk[128 j : 128 j + 128], rows [256 j + 128, 256 j + 256) = Wv[128 j : 128 j + 128];
 * q bf16: projected queries [users * 64, H] (q_batch_rows = 64) or one shared set [64, H] (q_batch_rows = 0);
 * key_mask fp32 [users, S] (1 attend / 0 masked) or NULL; v_bias fp32 [H] or NULL (the key bias cancels in the softmax);
 * out bf16 [users * 64, H].  Needs S % 64 == 0, 64 queries per user, head_dim 64, an even number of heads, K % 64 == 0.
 * workspace: unirec_kv_attention_workspace_bytes(users, num_heads) bytes of device memory, 16-byte aligned. */
int64_t unirec_kv_attention_workspace_bytes(int64_t users, int64_t num_heads);
int unirec_kv_attention_fused(const void* x, int64_t ldx, const void* w_packed, int64_t ldw, const void* q, int64_t ldq,
                              int64_t q_batch_rows, const float* key_mask, const float* v_bias, void* out, int64_t ldo,
                              void* workspace, int64_t workspace_bytes, int64_t users, int64_t S, int64_t num_heads,
                              int64_t K, float scale, void* stream);

/* Event-context encoders in front of the user-sequence builder (models/mwne.py:504-566 TimestampEncoder,
 * :569-610 GeoCoordinateEncoder; summed per event at models/user_sequence_encoder.py:125-131).  First half of both MLPs:
 *   out[e, 0:H]  = gelu(W1t f_time(timestamps[e]) + b1t)   f_time = 9 features (secular + 4 sin/cos pairs), W1t [H, 9]
 *   out[e, H:2H] = gelu(W1g f_geo(coords[e]) + b1g)        f_geo = unit-sphere xyz of (lat, lon) degrees, W1g [H, 3]
 * H = `hidden` = 2 x embedding_dim; out bf16 [n, >= 2H] with row stride ldo.  The second Linear of both encoders and their
 * sum is then ONE unirec_linear_bf16 call with W = [W2t | W2g] ([D, 2H]) and bias b2t + b2g.  timestamps: fp32, or int64
 * converted like torch's .float() (ts_int64 = 1); either input may be NULL (its half is written as zeros).
 * feats (may be NULL): fp32 [n, 12], the raw features (9 time + 3 geo), for tests. */
int unirec_context_hidden(const void* timestamps, int ts_int64, const float* coords, const float* w1t, const float* b1t,
                          const float* w1g, const float* b1g, int64_t n, int64_t hidden, void* out, int64_t ldo,
                          float* feats, void* stream);

/* ImprovedMathematicalEncoder.forward (models/mwne.py:134-183): out[i, :] = [interleaved (cos, sin)(x f_k) o fourier_w |
 * (x, sign x) o raw_scale | x * extra_w], times `scale` when given (eval-mode MathematicallyAwareNormalizer, :55-62:
 * scale = clamp(target_std / (running_std + 1e-8), 0.1, 10)).  numbers fp32 [n]; freqs [F]; fourier_w [2F]; raw_scale
 * [2] or NULL (include_raw = False); extra_w [D - 2F - raw] or NULL; out fp32 / bf16 [n, D]. */
int unirec_mwne_encode(const float* numbers, int64_t n, const float* freqs, int64_t F, const float* fourier_w,
                       const float* raw_scale, const float* extra_w, const float* scale, int64_t D, void* out,
                       int out_fp32, void* stream);

/* Cross-attention K/V projection of the user Q-Former straight from the item-token table: the user
 * sequence of models/user_sequence_encoder.py:128-140 + training/user_qformer_training.py:153-161 is never materialised.
 *   out[m, :] = A[m, :] W^T + bias + posbias[m % period, :],   m = (user * slots_per_user + slot) * 32 + token
 *   A[m, :]   = table[ids[user, slot] * 32 + token, :]   if slot < lengths[user]   (an id outside the table reads zeros)
 *             = pad_table[slot * 32 + token, :]           otherwise
 * table = item tokens viewed 2-D [table_rows = items * 32, K]; posbias bf16 [>= period + 128, N] = PE W^T with its first
 * 128 rows repeated after row `period` (a 128-row tile may run over a user boundary); pad_table bf16 [period, K] = -PE, so
 * that padding rows come out as `bias` alone like the reference's zero-padded sequence.  period = slots_per_user * 32,
 * M a multiple of period, N % 256 == 0, K % 64 == 0.  tcgen05 CTA-pair kernel, A tiles loaded as four 32-row TMA boxes. */
int unirec_linear_gather_bf16(const void* table, int64_t ld_table, int64_t table_rows, const int64_t* ids,
                              const int32_t* lengths, int64_t slots_per_user, const void* pad_table, int64_t ld_pad,
                              int64_t pad_rows, const void* W, int64_t ldw, const float* bias, const void* posbias,
                              int64_t ld_pos, int64_t pos_rows, int64_t period, void* out, int64_t ldo, int64_t M,
                              int64_t N, int64_t K, void* stream);

/* Reconstruction-quality metrics (evaluation/evaluate_item_qformer.py:66-95), one pass, no synchronisation:
 * over the rows (item, field) with mask != 0:  acc[0] += ||rec - orig||^2,  acc[1] += cosine(rec, orig),  acc[2] += 1.
 * rec [rows, E] fp32 (rec_fp32 = 1) or bf16, orig fp32 [rows, E], mask fp32 [rows], acc = 3 doubles on the device
 * (accumulated: zero them before the first batch).  masked MSE of a batch (:74-75) = acc[0] / acc[2]. */
int unirec_reconstruction_metrics(const void* rec, int rec_fp32, const float* orig, const float* mask, int64_t rows,
                                  int64_t E, float eps, double* acc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIREC_B200_H_ */
