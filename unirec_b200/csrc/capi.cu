// extern "C" surface of libunirec_b200.so (declared in include/unirec_b200.h): thin argument
// marshalling over the kernel launchers; no torch types, borrowed device pointers, caller's stream.
#include "common.cuh"
#include "umma_pipe.cuh"
#include "../../include/unirec_b200.h"

#include <atomic>

namespace unirec {
const char* get_last_error();
int gemm_bf16_ln(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* residual,
                 long long ldr, void* out, long long ldo, long long M, long long N, long long K, int epilogue,
                 const LnFold& ln, cudaStream_t stream);
long long gemm_cg2_stats_parts(long long N);
int gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* residual,
              long long ldr, int res_row_mod, void* out, long long ldo, int out_fp32, long long M, long long N,
              long long K, int epilogue, int block_n, int max_ctas, cudaStream_t stream);
int layernorm(const void* x, int x_fp32, long long ldx, int in_row_mod, const void* residual, long long ldres,
              const float* gamma, const float* beta, float eps, void* out, int out_fp32, long long ldo,
              long long rows, long long H, cudaStream_t stream);
int cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream);
int mean_tokens(const void* x, long long ldx, long long B, long long T, long long H, void* out, long long ldo,
                int out_fp32, cudaStream_t stream);
int field_projection(const void* rec, const float* Wp, const float* bp, void* out, int out_fp32, long long B,
                     long long T, long long F, long long E, cudaStream_t stream);
int build_user_sequence(const void* table, long long num_items, const long long* history, const int* lengths,
                        const void* ctx, const float* pe, void* seq, float* mask, long long B, long long Hmax, long long Q,
                        long long D, cudaStream_t stream);
int positional_encoding(float* pe, long long S, long long D, cudaStream_t stream);
int inv_l2_norm(const void* x, int x_fp32, long long ldx, float* inv, long long rows, long long D, float eps,
                cudaStream_t stream);
int attention(const void* q, long long ldq, long long q_batch_rows, const void* k, long long ldk, const void* v,
              long long ldv, long long kv_batch_rows, const float* key_mask, void* out, long long ldo, long long batch,
              long long num_heads, long long nq, long long nk, long long head_dim, float scale, unsigned drop_thr16,
              unsigned long long drop_seed, unsigned drop_site, const unsigned long long* drop_seed_offset,
              cudaStream_t stream);
int dropout_add(const void* x, long long ldx, long long x_row_mod, const void* res, long long ldres, void* out,
                long long ldo, long long rows, long long H, unsigned thr16, unsigned long long seed, unsigned site,
                const unsigned long long* seed_offset, cudaStream_t stream);
int dropout_backward(const void* dy, long long lddy, void* dx, long long lddx, long long rows, long long H,
                     unsigned thr16, unsigned long long seed, unsigned site, const unsigned long long* seed_offset,
                     cudaStream_t stream);

int gemm_bf16_general(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, void* out,
                      long long ldo, int out_fp32, int accumulate, long long M, long long N, long long K, int ksplit,
                      cudaStream_t stream);
int gelu_forward(const void* z, void* out, long long n, cudaStream_t stream);
int gelu_backward(const void* z, const void* da, void* dz, long long n, cudaStream_t stream);
int colsum(const void* x, long long ld, long long rows, long long N, float* out, cudaStream_t stream);
int layernorm_backward(const void* x, long long ldx, const void* dy, long long lddy, const void* dy2, long long lddy2,
                       const float* gamma, float eps, void* dx, long long lddx, float* dgamma, float* dbeta,
                       long long rows, long long H, unsigned drop_thr16, unsigned long long drop_seed, unsigned drop_site,
                       const unsigned long long* drop_seed_offset, void* dx_drop, long long lddrop, float* dbias,
                       cudaStream_t stream);
int attention_backward(const void* q, long long ldq, long long q_batch_rows, const void* k, long long ldk, const void* v,
                       long long ldv, long long kv_batch_rows, const float* key_mask, const void* dout, long long lddo,
                       void* dq, long long lddq, void* dk, long long lddk, void* dv, long long lddv, long long batch,
                       long long num_heads, long long nq, long long nk, long long head_dim, float scale,
                       unsigned drop_thr16, unsigned long long drop_seed, unsigned drop_site,
                       const unsigned long long* drop_seed_offset, cudaStream_t stream);

int list_scores(const void* users, long long ldu, const void* pos, long long ldp, const void* cands, long long ldc,
                int fp32, const unsigned char* mask, const long long* offsets, long long B, long long C, long long D,
                float eps, float* sims, float* inv_norm, cudaStream_t stream);
int infonce_rank(const float* sims, long long B, long long C, float temperature, float* loss, int* rank,
                 cudaStream_t stream);
int list_scores_backward(const void* users, long long ldu, const void* pos, long long ldp, const void* cands,
                         long long ldc, int fp32, const unsigned char* mask, const long long* offsets, long long B,
                         long long C, long long D, float eps, const float* sims, const float* inv_norm,
                         const float* dloss, float temperature, float* d_user, float* d_list, cudaStream_t stream);
int inject_tokens(const long long* input_ids, long long B, long long S, const long long* token_ids, long long num_slots,
                  const void* tokens, int tokens_fp32, void* text_embeds, int text_fp32, long long ld_text, long long Hd,
                  cudaStream_t stream);

int inject_tokens_backward(const long long* input_ids, long long B, long long S, const long long* token_ids,
                           long long num_slots, void* d_text, int text_fp32, long long ld_text, float* d_tokens,
                           long long Hd, cudaStream_t stream);

int context_hidden(const void* timestamps, int ts_int64, const float* coords, const float* w1t, const float* b1t,
                   const float* w1g, const float* b1g, long long n, long long hidden, void* out, long long ldo,
                   float* feats, cudaStream_t stream);
int mwne_encode(const float* numbers, long long n, const float* freqs, long long F, const float* fourier_w,
                const float* raw_scale, const float* extra_w, const float* scale, long long D, void* out, int out_fp32,
                cudaStream_t stream);

long long kv_attention_workspace_bytes(long long users, long long num_heads);
int kv_attention_fused(const void* x, long long ldx, const void* w_packed, long long ldw, const void* q, long long ldq,
                       long long q_batch_rows, const float* key_mask, const float* v_bias, void* out, long long ldo,
                       void* workspace, long long workspace_bytes, long long users, long long S, long long num_heads,
                       long long K, float scale, cudaStream_t stream);

int gemm_bf16_cg2_gather(const void* table, long long ld_table, long long table_rows, const long long* ids,
                         const int* lengths, long long slots_per_user, const void* pad_table, long long ld_pad,
                         long long pad_rows, const void* W, long long ldw, const float* bias, const void* posbias,
                         long long ld_pos, long long pos_rows, long long period, void* out, long long ldo, long long M,
                         long long N, long long K, cudaStream_t stream);
int reconstruction_metrics(const void* rec, int rec_fp32, const float* orig, const float* mask, long long rows, long long E,
                           float eps, double* acc, cudaStream_t stream);

std::atomic<long long> g_launch_count{0};
}  // namespace unirec

using namespace unirec;

#define COUNTED(expr)                                                                  \
    do {                                                                               \
        int rc__ = (expr);                                                             \
        if (rc__ == UNIREC_OK) g_launch_count.fetch_add(1, std::memory_order_relaxed); \
        return rc__;                                                                   \
    } while (0)

extern "C" {

int unirec_abi_version(void) { return UNIREC_B200_ABI_VERSION; }
const char* unirec_last_error(void) { return get_last_error(); }
int64_t unirec_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

int unirec_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                       const void* residual, int64_t ldr, int64_t res_row_mod, void* out, int64_t ldo, int out_fp32,
                       int64_t M, int64_t N, int64_t K, int epilogue, int block_n, int max_ctas, void* stream) {
    COUNTED(gemm_bf16(A, lda, W, ldw, bias, residual, ldr, static_cast<int>(res_row_mod), out, ldo, out_fp32, M, N, K,
                      epilogue, block_n, max_ctas, static_cast<cudaStream_t>(stream)));
}

int64_t unirec_linear_ln_stats_parts(int64_t N) { return gemm_cg2_stats_parts(N); }

int unirec_linear_ln_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* residual,
                          int64_t ldr, void* out, int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue,
                          const float* ln_in_stats, const float* ln_in_c, const float* ln_res_stats,
                          const float* ln_res_gamma, const float* ln_res_beta, float* stats_out, int ln_parts,
                          float ln_eps, int64_t ln_hidden, void* stream) {
    LnFold ln;
    ln.parts = ln_parts;
    ln.in_stats = ln_in_stats; ln.in_c = ln_in_c; ln.res_stats = ln_res_stats; ln.res_gamma = ln_res_gamma;
    ln.res_beta = ln_res_beta; ln.stats_out = stats_out; ln.eps = ln_eps; ln.hidden = ln_hidden;
    COUNTED(gemm_bf16_ln(A, lda, W, ldw, bias, residual, ldr, out, ldo, M, N, K, epilogue, ln,
                         static_cast<cudaStream_t>(stream)));
}

int unirec_layernorm(const void* x, int x_fp32, int64_t ldx, int64_t in_row_mod, const void* residual, int64_t ldres,
                     const float* gamma, const float* beta, float eps, void* out, int out_fp32, int64_t ldo,
                     int64_t rows, int64_t H, void* stream) {
    COUNTED(layernorm(x, x_fp32, ldx, static_cast<int>(in_row_mod), residual, ldres, gamma, beta, eps, out, out_fp32,
                      ldo, rows, H, static_cast<cudaStream_t>(stream)));
}

int unirec_attention(const void* q, int64_t ldq, int64_t q_batch_rows, const void* k, int64_t ldk, const void* v,
                     int64_t ldv, int64_t kv_batch_rows, const float* key_mask, void* out, int64_t ldo, int64_t batch,
                     int64_t num_heads, int64_t nq, int64_t nk, int64_t head_dim, float scale, void* stream) {
    COUNTED(attention(q, ldq, q_batch_rows, k, ldk, v, ldv, kv_batch_rows, key_mask, out, ldo, batch, num_heads, nq, nk,
                      head_dim, scale, 0u, 0ull, 0u, nullptr, static_cast<cudaStream_t>(stream)));
}

int unirec_attention_dropout(const void* q, int64_t ldq, int64_t q_batch_rows, const void* k, int64_t ldk, const void* v,
                             int64_t ldv, int64_t kv_batch_rows, const float* key_mask, void* out, int64_t ldo,
                             int64_t batch, int64_t num_heads, int64_t nq, int64_t nk, int64_t head_dim, float scale,
                             uint32_t thr16, uint64_t seed, uint32_t site, const uint64_t* seed_offset, void* stream) {
    COUNTED(attention(q, ldq, q_batch_rows, k, ldk, v, ldv, kv_batch_rows, key_mask, out, ldo, batch, num_heads, nq, nk,
                      head_dim, scale, thr16, seed, site, reinterpret_cast<const unsigned long long*>(seed_offset),
                      static_cast<cudaStream_t>(stream)));
}

int unirec_dropout_add(const void* x, int64_t ldx, int64_t x_row_mod, const void* residual, int64_t ldres, void* out,
                       int64_t ldo, int64_t rows, int64_t H, uint32_t thr16, uint64_t seed, uint32_t site,
                       const uint64_t* seed_offset, void* stream) {
    COUNTED(dropout_add(x, ldx, x_row_mod, residual, ldres, out, ldo, rows, H, thr16, seed, site,
                        reinterpret_cast<const unsigned long long*>(seed_offset), static_cast<cudaStream_t>(stream)));
}

int unirec_dropout_backward(const void* dy, int64_t lddy, void* dx, int64_t lddx, int64_t rows, int64_t H, uint32_t thr16,
                            uint64_t seed, uint32_t site, const uint64_t* seed_offset, void* stream) {
    COUNTED(dropout_backward(dy, lddy, dx, lddx, rows, H, thr16, seed, site,
                             reinterpret_cast<const unsigned long long*>(seed_offset), static_cast<cudaStream_t>(stream)));
}

int unirec_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
    COUNTED(cast_f32_to_bf16(in, out, n, static_cast<cudaStream_t>(stream)));
}

int unirec_mean_tokens(const void* x, int64_t ldx, int64_t B, int64_t T, int64_t H, void* out, int64_t ldo,
                       int out_fp32, void* stream) {
    COUNTED(mean_tokens(x, ldx, B, T, H, out, ldo, out_fp32, static_cast<cudaStream_t>(stream)));
}

int unirec_field_projection(const void* rec, const float* Wp, const float* bp, void* out, int out_fp32, int64_t B,
                            int64_t T, int64_t F, int64_t E, void* stream) {
    COUNTED(field_projection(rec, Wp, bp, out, out_fp32, B, T, F, E, static_cast<cudaStream_t>(stream)));
}

int unirec_build_user_sequence(const void* table, int64_t num_items, const int64_t* history, const int32_t* lengths,
                               const void* ctx, const float* pe_table, void* seq, float* mask, int64_t B, int64_t Hmax,
                               int64_t Q, int64_t D, void* stream) {
    COUNTED(build_user_sequence(table, num_items, reinterpret_cast<const long long*>(history), lengths, ctx, pe_table, seq,
                                mask, B, Hmax, Q, D, static_cast<cudaStream_t>(stream)));
}

int unirec_positional_encoding(float* pe_table, int64_t S, int64_t D, void* stream) {
    COUNTED(positional_encoding(pe_table, S, D, static_cast<cudaStream_t>(stream)));
}

int unirec_inv_l2_norm(const void* x, int x_fp32, int64_t ldx, float* inv, int64_t rows, int64_t D, float eps,
                       void* stream) {
    COUNTED(inv_l2_norm(x, x_fp32, ldx, inv, rows, D, eps, static_cast<cudaStream_t>(stream)));
}

int unirec_gemm_general(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* out,
                        int64_t ldo, int out_fp32, int accumulate, int64_t M, int64_t N, int64_t K, int ksplit,
                        void* stream) {
    COUNTED(gemm_bf16_general(A, lda, a_mn, B, ldb, b_mn, out, ldo, out_fp32, accumulate, M, N, K, ksplit,
                              static_cast<cudaStream_t>(stream)));
}

int unirec_gelu_forward(const void* z, void* out, int64_t n, void* stream) {
    COUNTED(gelu_forward(z, out, n, static_cast<cudaStream_t>(stream)));
}

int unirec_gelu_backward(const void* z, const void* da, void* dz, int64_t n, void* stream) {
    COUNTED(gelu_backward(z, da, dz, n, static_cast<cudaStream_t>(stream)));
}

int unirec_colsum(const void* x, int64_t ld, int64_t rows, int64_t N, float* out, void* stream) {
    COUNTED(colsum(x, ld, rows, N, out, static_cast<cudaStream_t>(stream)));
}

int unirec_layernorm_backward(const void* x, int64_t ldx, const void* dy, int64_t lddy, const void* dy2, int64_t lddy2,
                              const float* gamma, float eps, void* dx, int64_t lddx, float* dgamma, float* dbeta,
                              int64_t rows, int64_t H, void* stream) {
    COUNTED(layernorm_backward(x, ldx, dy, lddy, dy2, lddy2, gamma, eps, dx, lddx, dgamma, dbeta, rows, H, 0u, 0ull, 0u,
                               nullptr, nullptr, 0, nullptr, static_cast<cudaStream_t>(stream)));
}

int unirec_layernorm_backward_fused(const void* x, int64_t ldx, const void* dy, int64_t lddy, const void* dy2,
                                    int64_t lddy2, const float* gamma, float eps, void* dx, int64_t lddx, float* dgamma,
                                    float* dbeta, int64_t rows, int64_t H, uint32_t thr16, uint64_t seed, uint32_t site,
                                    const uint64_t* seed_offset, void* dx_drop, int64_t lddrop, float* dbias,
                                    void* stream) {
    COUNTED(layernorm_backward(x, ldx, dy, lddy, dy2, lddy2, gamma, eps, dx, lddx, dgamma, dbeta, rows, H, thr16, seed,
                               site, reinterpret_cast<const unsigned long long*>(seed_offset), dx_drop, lddrop, dbias,
                               static_cast<cudaStream_t>(stream)));
}

int unirec_attention_backward(const void* q, int64_t ldq, int64_t q_batch_rows, const void* k, int64_t ldk, const void* v,
                              int64_t ldv, int64_t kv_batch_rows, const float* key_mask, const void* dout, int64_t lddo,
                              void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t batch,
                              int64_t num_heads, int64_t nq, int64_t nk, int64_t head_dim, float scale, void* stream) {
    COUNTED(attention_backward(q, ldq, q_batch_rows, k, ldk, v, ldv, kv_batch_rows, key_mask, dout, lddo, dq, lddq, dk, lddk,
                               dv, lddv, batch, num_heads, nq, nk, head_dim, scale, 0u, 0ull, 0u, nullptr,
                               static_cast<cudaStream_t>(stream)));
}

int unirec_attention_dropout_backward(const void* q, int64_t ldq, int64_t q_batch_rows, const void* k, int64_t ldk,
                                      const void* v, int64_t ldv, int64_t kv_batch_rows, const float* key_mask,
                                      const void* dout, int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk,
                                      void* dv, int64_t lddv, int64_t batch, int64_t num_heads, int64_t nq, int64_t nk,
                                      int64_t head_dim, float scale, uint32_t thr16, uint64_t seed, uint32_t site,
                                      const uint64_t* seed_offset, void* stream) {
    COUNTED(attention_backward(q, ldq, q_batch_rows, k, ldk, v, ldv, kv_batch_rows, key_mask, dout, lddo, dq, lddq, dk, lddk,
                               dv, lddv, batch, num_heads, nq, nk, head_dim, scale, thr16, seed, site,
                               reinterpret_cast<const unsigned long long*>(seed_offset), static_cast<cudaStream_t>(stream)));
}

int unirec_list_scores(const void* users, int64_t ldu, const void* pos, int64_t ldp, const void* cands, int64_t ldc,
                       int fp32, const uint8_t* mask, const int64_t* offsets, int64_t B, int64_t C, int64_t D, float eps,
                       float* sims, float* inv_norm, void* stream) {
    COUNTED(list_scores(users, ldu, pos, ldp, cands, ldc, fp32, mask, reinterpret_cast<const long long*>(offsets), B, C, D,
                        eps, sims, inv_norm, static_cast<cudaStream_t>(stream)));
}

int unirec_infonce_rank(const float* sims, int64_t B, int64_t C, float temperature, float* loss, int32_t* rank,
                        void* stream) {
    COUNTED(infonce_rank(sims, B, C, temperature, loss, rank, static_cast<cudaStream_t>(stream)));
}

int unirec_list_scores_backward(const void* users, int64_t ldu, const void* pos, int64_t ldp, const void* cands,
                                int64_t ldc, int fp32, const uint8_t* mask, const int64_t* offsets, int64_t B, int64_t C,
                                int64_t D, float eps, const float* sims, const float* inv_norm, const float* dloss,
                                float temperature, float* d_user, float* d_list, void* stream) {
    COUNTED(list_scores_backward(users, ldu, pos, ldp, cands, ldc, fp32, mask, reinterpret_cast<const long long*>(offsets),
                                 B, C, D, eps, sims, inv_norm, dloss, temperature, d_user, d_list,
                                 static_cast<cudaStream_t>(stream)));
}

int unirec_inject_tokens(const int64_t* input_ids, int64_t B, int64_t S, const int64_t* token_ids, int64_t num_slots,
                         const void* tokens, int tokens_fp32, void* text_embeds, int text_fp32, int64_t ld_text,
                         int64_t Hd, void* stream) {
    COUNTED(inject_tokens(reinterpret_cast<const long long*>(input_ids), B, S, reinterpret_cast<const long long*>(token_ids),
                          num_slots, tokens, tokens_fp32, text_embeds, text_fp32, ld_text, Hd,
                          static_cast<cudaStream_t>(stream)));
}

int unirec_inject_tokens_backward(const int64_t* input_ids, int64_t B, int64_t S, const int64_t* token_ids,
                                  int64_t num_slots, void* d_text, int text_fp32, int64_t ld_text, float* d_tokens,
                                  int64_t Hd, void* stream) {
    COUNTED(inject_tokens_backward(reinterpret_cast<const long long*>(input_ids), B, S,
                                   reinterpret_cast<const long long*>(token_ids), num_slots, d_text, text_fp32, ld_text,
                                   d_tokens, Hd, static_cast<cudaStream_t>(stream)));
}

int unirec_context_hidden(const void* timestamps, int ts_int64, const float* coords, const float* w1t, const float* b1t,
                          const float* w1g, const float* b1g, int64_t n, int64_t hidden, void* out, int64_t ldo,
                          float* feats, void* stream) {
    COUNTED(context_hidden(timestamps, ts_int64, coords, w1t, b1t, w1g, b1g, n, hidden, out, ldo, feats,
                           static_cast<cudaStream_t>(stream)));
}

int unirec_mwne_encode(const float* numbers, int64_t n, const float* freqs, int64_t F, const float* fourier_w,
                       const float* raw_scale, const float* extra_w, const float* scale, int64_t D, void* out,
                       int out_fp32, void* stream) {
    COUNTED(mwne_encode(numbers, n, freqs, F, fourier_w, raw_scale, extra_w, scale, D, out, out_fp32,
                        static_cast<cudaStream_t>(stream)));
}

int64_t unirec_kv_attention_workspace_bytes(int64_t users, int64_t num_heads) {
    return kv_attention_workspace_bytes(users, num_heads);
}

int unirec_kv_attention_fused(const void* x, int64_t ldx, const void* w_packed, int64_t ldw, const void* q, int64_t ldq,
                              int64_t q_batch_rows, const float* key_mask, const float* v_bias, void* out, int64_t ldo,
                              void* workspace, int64_t workspace_bytes, int64_t users, int64_t S, int64_t num_heads,
                              int64_t K, float scale, void* stream) {
    COUNTED(kv_attention_fused(x, ldx, w_packed, ldw, q, ldq, q_batch_rows, key_mask, v_bias, out, ldo, workspace,
                               workspace_bytes, users, S, num_heads, K, scale, static_cast<cudaStream_t>(stream)));
}

int unirec_reconstruction_metrics(const void* rec, int rec_fp32, const float* orig, const float* mask, int64_t rows,
                                  int64_t E, float eps, double* acc, void* stream) {
    COUNTED(reconstruction_metrics(rec, rec_fp32, orig, mask, rows, E, eps, acc, static_cast<cudaStream_t>(stream)));
}

int unirec_linear_gather_bf16(const void* table, int64_t ld_table, int64_t table_rows, const int64_t* ids,
                              const int32_t* lengths, int64_t slots_per_user, const void* pad_table, int64_t ld_pad,
                              int64_t pad_rows, const void* W, int64_t ldw, const float* bias, const void* posbias,
                              int64_t ld_pos, int64_t pos_rows, int64_t period, void* out, int64_t ldo, int64_t M,
                              int64_t N, int64_t K, void* stream) {
    COUNTED(gemm_bf16_cg2_gather(table, ld_table, table_rows, reinterpret_cast<const long long*>(ids), lengths,
                                 slots_per_user, pad_table, ld_pad, pad_rows, W, ldw, bias, posbias, ld_pos, pos_rows, period,
                                 out, ldo, M, N, K, static_cast<cudaStream_t>(stream)));
}

}  // extern "C"
