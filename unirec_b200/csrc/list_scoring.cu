// Per-user candidate-LIST scoring: the ranking the reference's joint trainer actually performs.
//   InfoNCELoss.forward            training/train_item_individual_token_joint.py:331-352
//   MRREvaluator._compute_batch_mrr                                              :405-418
//   token injection of JointQwen3WithQFormer.forward                              :160-171
// Every user has its OWN short candidate list (one positive + up to ~100 negatives), either padded to
// [B, C, D] with a validity mask (the training collate) or ragged (the validation collate: a Python list of
// [n_b, D] tensors -> rows concatenated, CSR offsets).  The reference normalises all three operands, takes the
// dot products in a bmm / a Python loop per user and builds the loss per user in another Python loop.
//
// Here: one streaming pass over the candidate rows (the only large operand; HBM-bound, B*(1+C)*D elements read
// once, 16-byte loads, one warp per candidate row with four rows in flight) produces cosine similarities and the
// rows' inverse norms; a second tiny kernel (one warp per user) turns a user's similarities into the InfoNCE loss
// (-s_pos/T + logsumexp over the positive and the valid negatives) and the 1-based rank of the positive; the
// backward kernel re-streams the rows once and accumulates d loss / d user (the only operand that carries a
// gradient in the reference - item embeddings are precomputed data), optionally d loss / d candidates.
// F.normalize semantics: x / max(||x||, eps), eps = 1e-12.
#include "common.cuh"

namespace unirec {

constexpr int LS_THREADS = 256;
constexpr int LS_WARPS = LS_THREADS / 32;
constexpr int LS_ROWS_PER_WARP = 4;
constexpr int LS_ROWS_PER_CTA = LS_WARPS * LS_ROWS_PER_WARP;     // 32 list entries (entry 0 = the positive)

struct ListParams {
    const void* users; long long ldu;
    const void* pos; long long ldp;              // [B, D]
    const void* cands; long long ldc;            // padded: row (b * C + c); ragged: row (offsets[b] + c)
    const unsigned char* mask;                   // padded only, [B, C], 0 = padding; nullptr = all valid
    const long long* offsets;                    // ragged only, [B + 1]
    int C;                                       // padded: list length; ragged: longest list
    int D;
    float eps;
};

template <bool FP32>
UNIREC_DEVICE void load8(const void* row, int vi, float (&f)[8]) {
    if constexpr (FP32) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(row) + 2 * vi);
        const float4 b = __ldg(reinterpret_cast<const float4*>(row) + 2 * vi + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(row) + vi);
        f[0] = bf16_lo(a.x); f[1] = bf16_hi(a.x); f[2] = bf16_lo(a.y); f[3] = bf16_hi(a.y);
        f[4] = bf16_lo(a.z); f[5] = bf16_hi(a.z); f[6] = bf16_lo(a.w); f[7] = bf16_hi(a.w);
    }
}

// Row pointer of list entry e (0 = positive, e >= 1 = negative e - 1) of user b, or nullptr if the entry is padding.
template <bool FP32>
UNIREC_DEVICE const void* list_row(const ListParams& p, long long b, int e) {
    const size_t es = FP32 ? 4 : 2;
    if (e == 0) return reinterpret_cast<const unsigned char*>(p.pos) + static_cast<size_t>(b * p.ldp) * es;
    const int c = e - 1;
    long long row;
    if (p.offsets != nullptr) {
        const long long lo = __ldg(p.offsets + b), hi = __ldg(p.offsets + b + 1);
        if (c >= hi - lo) return nullptr;
        row = lo + c;
    } else {
        if (c >= p.C) return nullptr;
        if (p.mask != nullptr && __ldg(p.mask + b * p.C + c) == 0) return nullptr;
        row = b * p.C + c;
    }
    return reinterpret_cast<const unsigned char*>(p.cands) + static_cast<size_t>(row * p.ldc) * es;
}

// Stage the user's vector in shared memory as fp32 and return 1 / max(||u||, eps) (and ||u|| clamp flag through inv).
template <bool FP32>
UNIREC_DEVICE float stage_user(const ListParams& p, long long b, float* s_u, float* s_red) {
    const size_t es = FP32 ? 4 : 2;
    const void* urow = reinterpret_cast<const unsigned char*>(p.users) + static_cast<size_t>(b * p.ldu) * es;
    float ss = 0.f;
    for (int vi = threadIdx.x; vi < p.D / 8; vi += LS_THREADS) {
        float f[8];
        load8<FP32>(urow, vi, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s_u[vi * 8 + j] = f[j]; ss += f[j] * f[j]; }
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < LS_WARPS; ++w) tot += s_red[w];
    return 1.0f / fmaxf(sqrtf(tot), p.eps);
}

// sims[b, e] = cos(user_b, entry e) (-inf for padding), inv_norm[b, e] = 1 / max(||entry||, eps) (0 for padding).
template <bool FP32>
__global__ void __launch_bounds__(LS_THREADS)
list_scores_kernel(const ListParams p, float* __restrict__ sims, float* __restrict__ inv_norm) {
    extern __shared__ float s_u[];
    __shared__ float s_red[LS_WARPS];
    const long long b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float inv_u = stage_user<FP32>(p, b, s_u, s_red);
    const int e0 = blockIdx.x * LS_ROWS_PER_CTA + warp * LS_ROWS_PER_WARP;
    const void* rows[LS_ROWS_PER_WARP];
#pragma unroll
    for (int r = 0; r < LS_ROWS_PER_WARP; ++r) rows[r] = (e0 + r <= p.C) ? list_row<FP32>(p, b, e0 + r) : nullptr;
    float dot[LS_ROWS_PER_WARP] = {0.f, 0.f, 0.f, 0.f}, ss[LS_ROWS_PER_WARP] = {0.f, 0.f, 0.f, 0.f};
    for (int vi = lane; vi < p.D / 8; vi += 32) {
        float f[LS_ROWS_PER_WARP][8];
#pragma unroll
        for (int r = 0; r < LS_ROWS_PER_WARP; ++r) {
            if (rows[r] != nullptr) load8<FP32>(rows[r], vi, f[r]);
        }
        const float4 u0 = *reinterpret_cast<const float4*>(s_u + vi * 8);
        const float4 u1 = *reinterpret_cast<const float4*>(s_u + vi * 8 + 4);
        const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
        for (int r = 0; r < LS_ROWS_PER_WARP; ++r) {
            if (rows[r] != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { dot[r] += u[j] * f[r][j]; ss[r] += f[r][j] * f[r][j]; }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < LS_ROWS_PER_WARP; ++r) {
        const int e = e0 + r;
        if (e > p.C) continue;
        const float d = warp_sum(dot[r]), s2 = warp_sum(ss[r]);
        if (lane == 0) {
            const bool valid = rows[r] != nullptr;
            const float inv_c = valid ? 1.0f / fmaxf(sqrtf(s2), p.eps) : 0.f;
            sims[b * (p.C + 1) + e] = valid ? d * inv_u * inv_c : -INFINITY;
            if (inv_norm != nullptr) inv_norm[b * (p.C + 1) + e] = inv_c;
        }
    }
}

// One warp per user: loss = -s_0/T + logsumexp_e(s_e/T) over valid entries, rank = 1 + #{valid negatives with s_e > s_0}.
__global__ void __launch_bounds__(LS_THREADS)
infonce_rank_kernel(const float* __restrict__ sims, int B, int C, float inv_temperature, float* __restrict__ loss,
                    int* __restrict__ rank) {
    const int b = (blockIdx.x * LS_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* s = sims + static_cast<long long>(b) * (C + 1);
    const float s0 = s[0];
    float m = -INFINITY;
    for (int e = lane; e <= C; e += 32) m = fmaxf(m, s[e]);
    m = warp_max(m) * inv_temperature;
    float sum = 0.f;
    int above = 0;
    for (int e = lane; e <= C; e += 32) {
        const float v = s[e];
        if (v != -INFINITY) sum += __expf(v * inv_temperature - m);
        if (e > 0 && v > s0) ++above;
    }
    sum = warp_sum(sum);
    above = static_cast<int>(warp_sum(static_cast<float>(above)) + 0.5f);
    if (lane == 0) {
        if (loss != nullptr) loss[b] = m + logf(sum) - s0 * inv_temperature;
        if (rank != nullptr) rank[b] = 1 + above;
    }
}

// Backward of mean/weighted InfoNCE: g_e = dloss[b] * (softmax_e - [e == 0]) / T is the gradient of the loss with respect
// to the similarity s_e;  s_e = u^ . c^  with x^ = x / max(||x||, eps):
//   d s / d u = (c^ - s u^) / ||u||,   d s / d c = (u^ - s c^) / ||c||     (no projection term where the norm is clamped)
// d_user[b, :] (fp32, zero-initialised by the caller) += sum_e g_e d s_e / d u; d_list (optional, fp32, [B, 1 + C, D],
// zero rows for padding) = g_e d s_e / d c.
template <bool FP32>
__global__ void __launch_bounds__(LS_THREADS)
list_scores_backward_kernel(const ListParams p, const float* __restrict__ sims, const float* __restrict__ inv_norm,
                            const float* __restrict__ dloss, float inv_temperature, float* __restrict__ d_user,
                            float* __restrict__ d_list) {
    extern __shared__ float smem_f[];
    float* s_u = smem_f;                 // user vector
    float* s_acc = smem_f + p.D;         // this CTA's share of sum_e g_e * inv_c * c
    __shared__ float s_red[LS_WARPS];
    __shared__ float s_coef[LS_WARPS];   // per warp: sum_e g_e * s_e
    const long long b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float inv_u = stage_user<FP32>(p, b, s_u, s_red);
    for (int i = threadIdx.x; i < p.D; i += LS_THREADS) s_acc[i] = 0.f;
    __syncthreads();
    // softmax statistics of the user's list (every warp recomputes them: 1 + C <= a few hundred values)
    const float* s = sims + b * (p.C + 1);
    float m = -INFINITY;
    for (int e = lane; e <= p.C; e += 32) m = fmaxf(m, s[e]);
    m = warp_max(m) * inv_temperature;
    float z = 0.f;
    for (int e = lane; e <= p.C; e += 32) {
        const float v = s[e];
        if (v != -INFINITY) z += __expf(v * inv_temperature - m);
    }
    z = warp_sum(z);
    const float gscale = dloss[b] * inv_temperature;
    const bool u_clamped = inv_u >= 1.0f / p.eps;     // ||u|| <= eps: u^ = u / eps, no projection term

    const int e0 = blockIdx.x * LS_ROWS_PER_CTA + warp * LS_ROWS_PER_WARP;
    float coef = 0.f;
#pragma unroll 1
    for (int r = 0; r < LS_ROWS_PER_WARP; ++r) {
        const int e = e0 + r;
        if (e > p.C) break;
        const void* row = list_row<FP32>(p, b, e);
        float* drow = d_list != nullptr ? d_list + (b * (p.C + 1) + e) * static_cast<long long>(p.D) : nullptr;
        if (row == nullptr) {
            if (drow != nullptr)
                for (int i = lane; i < p.D; i += 32) drow[i] = 0.f;
            continue;
        }
        const float se = s[e];
        const float inv_c = inv_norm[b * (p.C + 1) + e];
        const float g = gscale * (__expf(se * inv_temperature - m) / z - (e == 0 ? 1.f : 0.f));
        coef += g * se;
        const bool c_clamped = inv_c >= 1.0f / p.eps;
        const float wc = g * inv_c;                 // weight of c in d_user (through c^ = c * inv_c)
        for (int vi = lane; vi < p.D / 8; vi += 32) {
            float f[8];
            load8<FP32>(row, vi, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&s_acc[vi * 8 + j], wc * f[j]);
                if (drow != nullptr) {
                    const float uh = s_u[vi * 8 + j] * inv_u, ch = f[j] * inv_c;
                    drow[vi * 8 + j] = g * (uh - (c_clamped ? 0.f : se * ch)) * inv_c;
                }
            }
        }
    }
    if (lane == 0) s_coef[warp] = coef;
    __syncthreads();
    float ctot = 0.f;
#pragma unroll
    for (int w = 0; w < LS_WARPS; ++w) ctot += s_coef[w];
    if (u_clamped) ctot = 0.f;
    // d_user += (sum_e g_e c^_e - (sum_e g_e s_e) u^) / ||u||_clamped
    for (int i = threadIdx.x; i < p.D; i += LS_THREADS) {
        const float v = (s_acc[i] - ctot * s_u[i] * inv_u) * inv_u;
        if (v != 0.f) atomicAdd(d_user + b * static_cast<long long>(p.D) + i, v);
    }
}

// text_embeds[b, s, :] = tokens[b, i, j, :] wherever input_ids[b, s] == token_ids[i * Q + j]  (:160-171).  One warp per
// sequence position; token_ids holds num_hist * Q distinct placeholder ids.
template <bool SRC_FP32, bool DST_FP32>
__global__ void __launch_bounds__(LS_THREADS)
inject_tokens_kernel(const long long* __restrict__ input_ids, long long rows, int S, const long long* __restrict__ token_ids,
                     int num_slots, const void* __restrict__ tokens, void* __restrict__ text_embeds, long long ld_text,
                     int Hd) {
    const long long row = (static_cast<long long>(blockIdx.x) * LS_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long id = __ldg(input_ids + row);
    int slot = -1;
    for (int k0 = 0; k0 < num_slots; k0 += 32) {
        const int k = k0 + lane;
        const bool hit = k < num_slots && __ldg(token_ids + k) == id;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal != 0u) slot = k0 + (31 - __clz(bal));       // the reference's loops let the LAST matching (i, j) win
    }
    if (slot < 0) return;
    const long long b = row / S;
    const size_t ss = SRC_FP32 ? 4 : 2, ds = DST_FP32 ? 4 : 2;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(tokens) +
                               (static_cast<size_t>(b) * num_slots + slot) * Hd * ss;
    unsigned char* dst = reinterpret_cast<unsigned char*>(text_embeds) + static_cast<size_t>(row) * ld_text * ds;
    for (int vi = lane; vi < Hd / 8; vi += 32) {
        float f[8];
        load8<SRC_FP32>(src, vi, f);
        if constexpr (DST_FP32) {
            reinterpret_cast<float4*>(dst)[2 * vi] = make_float4(f[0], f[1], f[2], f[3]);
            reinterpret_cast<float4*>(dst)[2 * vi + 1] = make_float4(f[4], f[5], f[6], f[7]);
        } else {
            reinterpret_cast<uint4*>(dst)[vi] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                                           pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        }
    }
}

// ------------------------------------------------------------------------------------------- host launchers
static int check_list_args(const char* what, const void* users, const void* pos, const void* cands, long long ldu,
                           long long ldp, long long ldc, long long B, long long C, long long D, int fp32,
                           const unsigned char* mask, const long long* offsets) {
    if (users == nullptr || pos == nullptr || (cands == nullptr && C > 0) || B <= 0 || C < 0 || D <= 0) {
        set_last_error("%s: null pointer or empty shape (B=%lld C=%lld D=%lld)", what, B, C, D);
        return UNIREC_ERR_BAD_ARG;
    }
    const long long al = fp32 ? 4 : 8;      // 16-byte rows
    if (D % 8 != 0 || ldu % al != 0 || ldp % al != 0 || ldc % al != 0 || D > 8192 || B > 65535) {
        set_last_error("%s: D %% 8 != 0, D > 8192, B > 65535 or rows not 16-byte aligned (D=%lld)", what, D);
        return UNIREC_ERR_BAD_ARG;
    }
    if (mask != nullptr && offsets != nullptr) {
        set_last_error("%s: pass either a padding mask (padded lists) or offsets (ragged lists), not both", what);
        return UNIREC_ERR_BAD_ARG;
    }
    return UNIREC_OK;
}

static ListParams make_list_params(const void* users, long long ldu, const void* pos, long long ldp, const void* cands,
                                   long long ldc, const unsigned char* mask, const long long* offsets, long long C,
                                   long long D, float eps) {
    ListParams p;
    p.users = users; p.ldu = ldu; p.pos = pos; p.ldp = ldp; p.cands = cands; p.ldc = ldc; p.mask = mask;
    p.offsets = offsets; p.C = static_cast<int>(C); p.D = static_cast<int>(D); p.eps = eps;
    return p;
}

int list_scores(const void* users, long long ldu, const void* pos, long long ldp, const void* cands, long long ldc,
                int fp32, const unsigned char* mask, const long long* offsets, long long B, long long C, long long D,
                float eps, float* sims, float* inv_norm, cudaStream_t stream) {
    int rc = check_list_args("list_scores", users, pos, cands, ldu, ldp, ldc, B, C, D, fp32, mask, offsets);
    if (rc != UNIREC_OK) return rc;
    if (sims == nullptr) { set_last_error("list_scores: sims is null"); return UNIREC_ERR_BAD_ARG; }
    const ListParams p = make_list_params(users, ldu, pos, ldp, cands, ldc, mask, offsets, C, D, eps);
    const dim3 grid(static_cast<unsigned>((C + 1 + LS_ROWS_PER_CTA - 1) / LS_ROWS_PER_CTA), static_cast<unsigned>(B));
    const size_t smem = static_cast<size_t>(D) * sizeof(float);
    if (fp32) list_scores_kernel<true><<<grid, LS_THREADS, smem, stream>>>(p, sims, inv_norm);
    else list_scores_kernel<false><<<grid, LS_THREADS, smem, stream>>>(p, sims, inv_norm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("list_scores launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

int infonce_rank(const float* sims, long long B, long long C, float temperature, float* loss, int* rank,
                 cudaStream_t stream) {
    if (sims == nullptr || B <= 0 || C < 0 || !(temperature > 0.f) || (loss == nullptr && rank == nullptr)) {
        set_last_error("infonce_rank: null pointer, empty shape or temperature <= 0");
        return UNIREC_ERR_BAD_ARG;
    }
    const long long blocks = (B * 32 + LS_THREADS - 1) / LS_THREADS;
    infonce_rank_kernel<<<static_cast<unsigned>(blocks), LS_THREADS, 0, stream>>>(sims, static_cast<int>(B),
                                                                                   static_cast<int>(C), 1.0f / temperature,
                                                                                   loss, rank);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("infonce_rank launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

int list_scores_backward(const void* users, long long ldu, const void* pos, long long ldp, const void* cands,
                         long long ldc, int fp32, const unsigned char* mask, const long long* offsets, long long B,
                         long long C, long long D, float eps, const float* sims, const float* inv_norm,
                         const float* dloss, float temperature, float* d_user, float* d_list, cudaStream_t stream) {
    int rc = check_list_args("list_scores_backward", users, pos, cands, ldu, ldp, ldc, B, C, D, fp32, mask, offsets);
    if (rc != UNIREC_OK) return rc;
    if (sims == nullptr || inv_norm == nullptr || dloss == nullptr || d_user == nullptr || !(temperature > 0.f)) {
        set_last_error("list_scores_backward: null pointer or temperature <= 0");
        return UNIREC_ERR_BAD_ARG;
    }
    const ListParams p = make_list_params(users, ldu, pos, ldp, cands, ldc, mask, offsets, C, D, eps);
    const dim3 grid(static_cast<unsigned>((C + 1 + LS_ROWS_PER_CTA - 1) / LS_ROWS_PER_CTA), static_cast<unsigned>(B));
    const size_t smem = 2 * static_cast<size_t>(D) * sizeof(float);
    if (smem > 48 * 1024) {
        static bool attr_set[2] = {false, false};
        if (!attr_set[fp32 ? 1 : 0]) {
            cudaError_t e = fp32 ? cudaFuncSetAttribute(list_scores_backward_kernel<true>,
                                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)
                                 : cudaFuncSetAttribute(list_scores_backward_kernel<false>,
                                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (e != cudaSuccess) {
                set_last_error("list_scores_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return UNIREC_ERR_CUDA;
            }
            attr_set[fp32 ? 1 : 0] = true;
        }
    }
    if (fp32)
        list_scores_backward_kernel<true><<<grid, LS_THREADS, smem, stream>>>(p, sims, inv_norm, dloss, 1.0f / temperature,
                                                                             d_user, d_list);
    else
        list_scores_backward_kernel<false><<<grid, LS_THREADS, smem, stream>>>(p, sims, inv_norm, dloss,
                                                                              1.0f / temperature, d_user, d_list);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("list_scores_backward launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

int inject_tokens(const long long* input_ids, long long B, long long S, const long long* token_ids, long long num_slots,
                  const void* tokens, int tokens_fp32, void* text_embeds, int text_fp32, long long ld_text, long long Hd,
                  cudaStream_t stream) {
    if (input_ids == nullptr || token_ids == nullptr || tokens == nullptr || text_embeds == nullptr || B <= 0 || S <= 0 ||
        num_slots <= 0 || Hd <= 0 || Hd % 8 != 0 || ld_text % (text_fp32 ? 4 : 8) != 0) {
        set_last_error("inject_tokens: null pointer, empty shape, Hd %% 8 != 0 or unaligned rows (Hd=%lld)", Hd);
        return UNIREC_ERR_BAD_ARG;
    }
    const long long rows = B * S;
    const unsigned blocks = static_cast<unsigned>((rows * 32 + LS_THREADS - 1) / LS_THREADS);
    const int s = static_cast<int>(S), ns = static_cast<int>(num_slots), hd = static_cast<int>(Hd);
    if (tokens_fp32 && text_fp32)
        inject_tokens_kernel<true, true><<<blocks, LS_THREADS, 0, stream>>>(input_ids, rows, s, token_ids, ns, tokens,
                                                                            text_embeds, ld_text, hd);
    else if (tokens_fp32)
        inject_tokens_kernel<true, false><<<blocks, LS_THREADS, 0, stream>>>(input_ids, rows, s, token_ids, ns, tokens,
                                                                             text_embeds, ld_text, hd);
    else if (text_fp32)
        inject_tokens_kernel<false, true><<<blocks, LS_THREADS, 0, stream>>>(input_ids, rows, s, token_ids, ns, tokens,
                                                                             text_embeds, ld_text, hd);
    else
        inject_tokens_kernel<false, false><<<blocks, LS_THREADS, 0, stream>>>(input_ids, rows, s, token_ids, ns, tokens,
                                                                              text_embeds, ld_text, hd);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("inject_tokens launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

}  // namespace unirec
