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
// once, 16-byte loads, one warp per candidate row with 4 fp32 / 8 bf16 rows in flight) produces cosine similarities and the
// rows' inverse norms; a second tiny kernel (one warp per user) turns a user's similarities into the InfoNCE loss
// (-s_pos/T + logsumexp over the positive and the valid negatives) and the 1-based rank of the positive; the
// backward kernel re-streams the rows once and accumulates d loss / d user (the only operand that carries a
// gradient in the reference - item embeddings are precomputed data), optionally d loss / d candidates.
// F.normalize semantics: x / max(||x||, eps), eps = 1e-12.
#include "common.cuh"

namespace unirec {

constexpr int LS_THREADS = 256;
constexpr int LS_WARPS = LS_THREADS / 32;
// list entries one warp keeps in flight in the forward pass: 4 fp32 rows or 8 bf16 rows (the same bytes per lane;
// ncu-less first measurement: bf16 with 4 rows reached 58 % of HBM peak where fp32 reached 92 %)
template <bool FP32> constexpr int ls_rows_per_warp() { return FP32 ? 4 : 8; }
constexpr int LSB_THREADS = 128;         // backward: one 16-byte chunk of the row per thread, 16 list entries per CTA
constexpr int LSB_ROWS = 16;

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

// 8 consecutive elements of a row: raw 16-byte loads first (so that several rows' loads are in flight without holding
// their unpacked values in registers), conversion to fp32 at the point of use.
template <bool FP32> struct Raw8 { uint4 a; uint4 b; };
template <> struct Raw8<false> { uint4 a; };

template <bool FP32>
UNIREC_DEVICE Raw8<FP32> load_raw8(const void* row, int vi) {
    Raw8<FP32> q;
    if constexpr (FP32) {
        q.a = __ldg(reinterpret_cast<const uint4*>(row) + 2 * vi);
        q.b = __ldg(reinterpret_cast<const uint4*>(row) + 2 * vi + 1);
    } else {
        q.a = __ldg(reinterpret_cast<const uint4*>(row) + vi);
    }
    return q;
}

template <bool FP32>
UNIREC_DEVICE void unpack8(const Raw8<FP32>& q, float (&f)[8]) {
    if constexpr (FP32) {
        f[0] = __uint_as_float(q.a.x); f[1] = __uint_as_float(q.a.y); f[2] = __uint_as_float(q.a.z);
        f[3] = __uint_as_float(q.a.w); f[4] = __uint_as_float(q.b.x); f[5] = __uint_as_float(q.b.y);
        f[6] = __uint_as_float(q.b.z); f[7] = __uint_as_float(q.b.w);
    } else {
        f[0] = bf16_lo(q.a.x); f[1] = bf16_hi(q.a.x); f[2] = bf16_lo(q.a.y); f[3] = bf16_hi(q.a.y);
        f[4] = bf16_lo(q.a.z); f[5] = bf16_hi(q.a.z); f[6] = bf16_lo(q.a.w); f[7] = bf16_hi(q.a.w);
    }
}

template <bool FP32>
UNIREC_DEVICE void load8(const void* row, int vi, float (&f)[8]) {
    unpack8<FP32>(load_raw8<FP32>(row, vi), f);
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
__global__ void __launch_bounds__(LS_THREADS, 2)
list_scores_kernel(const ListParams p, float* __restrict__ sims, float* __restrict__ inv_norm) {
    constexpr int ROWS = ls_rows_per_warp<FP32>();
    extern __shared__ float s_u[];
    __shared__ float s_red[LS_WARPS];
    const long long b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float inv_u = stage_user<FP32>(p, b, s_u, s_red);
    const int e0 = (blockIdx.x * LS_WARPS + warp) * ROWS;
    const void* rows[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) rows[r] = (e0 + r <= p.C) ? list_row<FP32>(p, b, e0 + r) : nullptr;
    float dot[ROWS], ss[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { dot[r] = 0.f; ss[r] = 0.f; }
    for (int vi = lane; vi < p.D / 8; vi += 32) {
        Raw8<FP32> q[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (rows[r] != nullptr) q[r] = load_raw8<FP32>(rows[r], vi);
        }
        const float4 u0 = *reinterpret_cast<const float4*>(s_u + vi * 8);
        const float4 u1 = *reinterpret_cast<const float4*>(s_u + vi * 8 + 4);
        const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (rows[r] != nullptr) {
                float f[8];
                unpack8<FP32>(q[r], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) { dot[r] += u[j] * f[j]; ss[r] += f[j] * f[j]; }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
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
// A CTA owns LSB_ROWS list entries of one user; every thread owns 16-byte chunks of the D axis and walks the CTA's rows
// with its partial sum_e g_e c^_e in registers (the first version let one warp own a row and summed in shared memory
// with atomics: 18 % of HBM peak), then adds its share to d_user with one global atomic per element.
template <bool FP32>
__global__ void __launch_bounds__(LSB_THREADS)
list_scores_backward_kernel(const ListParams p, const float* __restrict__ sims, const float* __restrict__ inv_norm,
                            const float* __restrict__ dloss, float inv_temperature, float* __restrict__ d_user,
                            float* __restrict__ d_list) {
    __shared__ const void* s_row[LSB_ROWS];
    __shared__ float s_g[LSB_ROWS], s_se[LSB_ROWS], s_invc[LSB_ROWS];
    __shared__ float s_red[LSB_THREADS / 32];
    __shared__ float s_ctot;
    const long long b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t es = FP32 ? 4 : 2;
    const void* urow = reinterpret_cast<const unsigned char*>(p.users) + static_cast<size_t>(b * p.ldu) * es;
    // ||u||: every thread sums the squares of its chunks
    float ss = 0.f;
    for (int vi = threadIdx.x; vi < p.D / 8; vi += LSB_THREADS) {
        float f[8];
        load8<FP32>(urow, vi, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
    ss = warp_sum(ss);
    if (lane == 0) s_red[warp] = ss;
    // warp 0: softmax statistics of the user's list and the per-row coefficients of this CTA's entries
    const int e0 = blockIdx.x * LSB_ROWS;
    if (warp == 0) {
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
        float gs = 0.f;
        if (lane < LSB_ROWS) {
            const int e = e0 + lane;
            const void* row = e <= p.C ? list_row<FP32>(p, b, e) : nullptr;
            float g = 0.f, se = 0.f, ic = 0.f;
            if (row != nullptr) {
                se = s[e];
                ic = inv_norm[b * (p.C + 1) + e];
                g = dloss[b] * inv_temperature * (__expf(se * inv_temperature - m) / z - (e == 0 ? 1.f : 0.f));
            }
            s_row[lane] = row; s_g[lane] = g; s_se[lane] = se; s_invc[lane] = ic;
            gs = g * se;
        }
        gs = warp_sum(gs);
        if (lane == 0) s_ctot = gs;
    }
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < LSB_THREADS / 32; ++w) tot += s_red[w];
    const float inv_u = 1.0f / fmaxf(sqrtf(tot), p.eps);
    const bool u_clamped = inv_u >= 1.0f / p.eps;      // ||u|| <= eps: u^ = u / eps, no projection term
    const float ctot = u_clamped ? 0.f : s_ctot;
    const int n_rows = min(LSB_ROWS, p.C + 1 - e0);

    for (int vi = threadIdx.x; vi < p.D / 8; vi += LSB_THREADS) {
        float u[8], acc[8];
        load8<FP32>(urow, vi, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) { u[j] *= inv_u; acc[j] = 0.f; }
#pragma unroll 4
        for (int r = 0; r < n_rows; ++r) {
            const void* row = s_row[r];
            float* drow = d_list != nullptr ? d_list + ((b * (p.C + 1) + e0 + r) * static_cast<long long>(p.D) + vi * 8)
                                            : nullptr;
            if (row == nullptr) {
                if (drow != nullptr) {
                    reinterpret_cast<float4*>(drow)[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    reinterpret_cast<float4*>(drow)[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                continue;
            }
            float f[8];
            load8<FP32>(row, vi, f);
            const float g = s_g[r], se = s_se[r], ic = s_invc[r];
            const float wc = g * ic;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += wc * f[j];
            if (drow != nullptr) {
                const float proj = (ic >= 1.0f / p.eps) ? 0.f : se * ic;     // clamped ||c||: no projection term
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = wc * (u[j] - proj * f[j]);
                reinterpret_cast<float4*>(drow)[0] = make_float4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<float4*>(drow)[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
        // d_user += (sum_e g_e c^_e - (sum_e g_e s_e) u^) / ||u||_clamped
        float* du = d_user + b * static_cast<long long>(p.D) + vi * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float v = (acc[j] - ctot * u[j]) * inv_u;
            if (v != 0.f) atomicAdd(du + j, v);
        }
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


// Backward of the injection: the gradient of an overwritten position flows to the token that replaced it and NOT to the
// text embedding it replaced.  d_text (in place) = upstream gradient with the overwritten rows zeroed; d_tokens fp32
// [B, num_slots, Hd] (zeroed by the caller) += the upstream rows of every position holding that slot's placeholder.
template <bool TEXT_FP32>
__global__ void __launch_bounds__(LS_THREADS)
inject_tokens_backward_kernel(const long long* __restrict__ input_ids, long long rows, int S,
                              const long long* __restrict__ token_ids, int num_slots, void* __restrict__ d_text,
                              long long ld_text, float* __restrict__ d_tokens, int Hd) {
    const long long row = (static_cast<long long>(blockIdx.x) * LS_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long id = __ldg(input_ids + row);
    int slot = -1;
    for (int k0 = 0; k0 < num_slots; k0 += 32) {
        const int k = k0 + lane;
        const bool hit = k < num_slots && __ldg(token_ids + k) == id;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal != 0u) slot = k0 + (31 - __clz(bal));       // same winner as the forward kernel
    }
    if (slot < 0) return;
    const long long b = row / S;
    unsigned char* g = reinterpret_cast<unsigned char*>(d_text) + static_cast<size_t>(row) * ld_text * (TEXT_FP32 ? 4 : 2);
    float* dt = d_tokens + (static_cast<size_t>(b) * num_slots + slot) * Hd;
    for (int vi = lane; vi < Hd / 8; vi += 32) {
        float f[8];
        load8<TEXT_FP32>(g, vi, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dt + vi * 8 + j, f[j]);
        if constexpr (TEXT_FP32) {
            reinterpret_cast<float4*>(g)[2 * vi] = make_float4(0.f, 0.f, 0.f, 0.f);
            reinterpret_cast<float4*>(g)[2 * vi + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            reinterpret_cast<uint4*>(g)[vi] = make_uint4(0u, 0u, 0u, 0u);
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
    const int rows_per_cta = LS_WARPS * (fp32 ? ls_rows_per_warp<true>() : ls_rows_per_warp<false>());
    const dim3 grid(static_cast<unsigned>((C + 1 + rows_per_cta - 1) / rows_per_cta), static_cast<unsigned>(B));
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
    const dim3 grid(static_cast<unsigned>((C + 1 + LSB_ROWS - 1) / LSB_ROWS), static_cast<unsigned>(B));
    if (fp32)
        list_scores_backward_kernel<true><<<grid, LSB_THREADS, 0, stream>>>(p, sims, inv_norm, dloss, 1.0f / temperature,
                                                                           d_user, d_list);
    else
        list_scores_backward_kernel<false><<<grid, LSB_THREADS, 0, stream>>>(p, sims, inv_norm, dloss, 1.0f / temperature,
                                                                            d_user, d_list);
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

int inject_tokens_backward(const long long* input_ids, long long B, long long S, const long long* token_ids,
                           long long num_slots, void* d_text, int text_fp32, long long ld_text, float* d_tokens,
                           long long Hd, cudaStream_t stream) {
    if (input_ids == nullptr || token_ids == nullptr || d_text == nullptr || d_tokens == nullptr || B <= 0 || S <= 0 ||
        num_slots <= 0 || Hd <= 0 || Hd % 8 != 0 || ld_text % (text_fp32 ? 4 : 8) != 0) {
        set_last_error("inject_tokens_backward: null pointer, empty shape, Hd %% 8 != 0 or unaligned rows (Hd=%lld)", Hd);
        return UNIREC_ERR_BAD_ARG;
    }
    const long long rows = B * S;
    const unsigned blocks = static_cast<unsigned>((rows * 32 + LS_THREADS - 1) / LS_THREADS);
    const int s = static_cast<int>(S), ns = static_cast<int>(num_slots), hd = static_cast<int>(Hd);
    if (text_fp32)
        inject_tokens_backward_kernel<true><<<blocks, LS_THREADS, 0, stream>>>(input_ids, rows, s, token_ids, ns, d_text,
                                                                               ld_text, d_tokens, hd);
    else
        inject_tokens_backward_kernel<false><<<blocks, LS_THREADS, 0, stream>>>(input_ids, rows, s, token_ids, ns, d_text,
                                                                                ld_text, d_tokens, hd);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("inject_tokens_backward launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

}  // namespace unirec
