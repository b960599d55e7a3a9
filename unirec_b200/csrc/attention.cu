// Fused small-query attention for the Q-Former path (sm_100a), head_dim = 64.
//
//   ctx[b, q, h*64:(h+1)*64] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] / 8 + mask[b,k] ) @ V[b,k,h,:]
//
// Replaces models/qformer.py:161-167 (head split), :205 (QK^T), :244-250 (scale, additive mask,
// softmax), :264-268 (PV, head merge by permute+contiguous).  Used for
//   item self-attention   32 queries x 32 keys          (KT = 32, one tile)
//   item cross-attention  32 queries x 14 field keys    (KT = 16, one tile)
//   user self-attention   64 queries x 64 keys          (KT = 64, one tile)
//   user cross-attention  64 queries x <= 1600 keys     (KT = 64, key-tiled online softmax)
// Persistent CTAs: each CTA walks a strided list of (batch, head) work items and, inside an item, the key
// tiles, as ONE flat sequence of K/V tile loads through a 3-deep cp.async ring - the loads of the next two
// tiles (which may already belong to the next work item, together with its Q tile) are in flight while the
// current tile is computed, so there is no load bubble between work items (an item-side work item is a
// single 4-12 KB tile) and one __syncthreads per tile.  Each warp owns 16 query rows.  Q, K and V tiles are
// staged in shared memory with cp.async (16-byte chunks, XOR swizzle), the two small matmuls run on mma.sync
// m16n8k16 (bf16 in, fp32 accumulate) from ldmatrix fragments, softmax statistics are fp32 with
// quad shuffles, and the context is written head-merged ([B*Q, H] row-major) through shared memory
// so that global stores are 128-byte row segments.  The kernel is HBM-bound by design: algorithmic
// bytes = (Q + K + V + O) tiles, each read/written exactly once.
//
// Mask semantics: key_mask[b,k] == 0 adds a large negative constant (the
// reference adds finfo.min, models/qformer.py:927-933 via HF invert_attention_mask) so masked keys get
// probability exactly 0 whenever any key is unmasked, and a row whose keys are ALL masked
// degenerates to uniform attention over all nk keys (not NaN, not zero) - exactly the fp32
// behaviour of the reference, where score + finfo.min == finfo.min for every key.
#include "common.cuh"
#include "dropout.cuh"
#include "umma_pipe.cuh"

#include <cstdlib>

namespace unirec {

struct AttnParams {
    const float* key_mask;   // [B, nk] (1 attend / 0 masked) or nullptr
    int num_heads, nq, nk;
    int q_broadcast;         // 1: the same queries serve every batch element
    int reverse;             // 1: walk the work items from the last to the first (see stream_reverse(), common.cuh)
    float scale_log2;        // softmax scale * log2(e)
    DropoutParams drop;      // train-mode dropout of the probabilities (models/qformer.py:258); DROP kernels only
};

constexpr float kMaskedLog2 = -1.0e30f;  // stands in for finfo.min (see header comment)

constexpr int ATT_STAGES = 3;

// Shared-memory plan (all tiles are 128-byte rows written by TMA with the 128-byte swizzle, 1 KB aligned):
//   sQ [STAGES][nq_pad]  Q tile of the work item (by item index), one 16-row TMA box per warp
//   sK, sV [STAGES][KT]  key / value tiles (by flat tile index)
//   sO [2][nq_pad]       context staging for the per-warp TMA stores (two in flight)
//   sM [STAGES][KT]      additive mask (log2 domain), full[STAGES] mbarriers
template <int KT, bool DROP>
__global__ void __launch_bounds__(128)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                 const AttnParams p, int num_items) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nq_pad = nwarps * 16;

    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATT_STAGES * nq_pad * 128;
    uint8_t* sV = sK + ATT_STAGES * KT * 128;
    uint8_t* sO = sV + ATT_STAGES * KT * 128;
    float* sM = reinterpret_cast<float*>(sO + 2 * nq_pad * 128);
    uint64_t* full = reinterpret_cast<uint64_t*>(sM + ATT_STAGES * KT);

    const int ntiles = (p.nk + KT - 1) / KT;                    // key tiles per work item
    const int my_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int total = my_items * ntiles;                        // flat tile count of this CTA
    // no additive mask at all (self-attention, full last tile): the select in the softmax is skipped
    const bool plain = (p.key_mask == nullptr) && (p.nk % KT == 0);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        tma_prefetch_desc(&tmap_o);
        for (int i = 0; i < ATT_STAGES; ++i) mbar_init(&full[i], nwarps);
        fence_mbar_init();
    }
    __syncthreads();

    // Loads of flat tile g: warp 0 fetches the K and V tiles, every warp its own 16 query rows when the tile is the
    // first of a work item; all of them complete on full[g % STAGES] (one arrival per warp).
    auto issue = [&](int g) {
        if (g >= total) return;
        const int it = g / ntiles, tile = g - it * ntiles;
        const int w0 = blockIdx.x + it * gridDim.x;
        const int w = p.reverse ? num_items - 1 - w0 : w0;
        const int b = w / p.num_heads, h = w - b * p.num_heads;
        const int buf = g % ATT_STAGES;
        if (!plain) {
            const float* mbase = p.key_mask ? p.key_mask + static_cast<long long>(b) * p.nk : nullptr;
            for (int i = threadIdx.x; i < KT; i += blockDim.x) {
                const int key = tile * KT + i;
                float m = -INFINITY;  // padding beyond nk: excluded
                if (key < p.nk) m = (mbase != nullptr && mbase[key] == 0.f) ? kMaskedLog2 : 0.f;
                sM[buf * KT + i] = m;
            }
        }
        if (lane == 0) {
            uint32_t bytes = (tile == 0) ? 16u * 128u : 0u;
            if (warp == 0) bytes += 2u * KT * 128u;
            mbar_arrive_expect_tx(&full[buf], bytes);
            if (tile == 0)
                tma_load_3d(&tmap_q, &full[buf], sQ + ((it % ATT_STAGES) * nq_pad + warp * 16) * 128, h * 64, warp * 16,
                            p.q_broadcast ? 0 : b);
            if (warp == 0) {
                tma_load_3d(&tmap_k, &full[buf], sK + buf * KT * 128, h * 64, tile * KT, b);
                tma_load_3d(&tmap_v, &full[buf], sV + buf * KT * 128, h * 64, tile * KT, b);
            }
        }
    };

#pragma unroll
    for (int g = 0; g < ATT_STAGES - 1; ++g) issue(g);

    const int g4 = lane >> 2, t = lane & 3;
    uint32_t qf[4][4];
    float o[8][4];
    float m_run[2], l_run[2];

    for (int g = 0; g < total; ++g) {
        const int it = g / ntiles, tile = g - it * ntiles;
        const int buf = g % ATT_STAGES;
        __syncthreads();                   // everyone is done with flat tile g-1 (its buffers are refilled below)
        issue(g + ATT_STAGES - 1);
        mbar_wait(&full[buf], (g / ATT_STAGES) & 1);   // flat tile g has landed

        const uint32_t sQi = smem_u32(sQ + (it % ATT_STAGES) * nq_pad * 128);
        if (tile == 0) {
            // Q fragments (A operand, 16 rows x 64 dims = 4 k-steps) stay in registers for the whole item
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = kk * 2 + (lane >> 4);
                ldmatrix_x4(sQi + swz128(r, c), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
            m_run[0] = m_run[1] = -INFINITY;
            l_run[0] = l_run[1] = 0.f;
        }

        // ---- S = Q K^T  (16 x KT per warp)
        float s[KT / 8][4];
#pragma unroll
        for (int j = 0; j < KT / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
        const uint32_t kaddr = smem_u32(sK + buf * KT * 128);
#pragma unroll
        for (int jj = 0; jj < KT / 16; ++jj) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t b0, b1, b2, b3;
                const int r = jj * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = kk * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(kaddr + swz128(r, c), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * jj], qf[kk], b0, b1);
                mma_bf16_16816(s[2 * jj + 1], qf[kk], b2, b3);
            }
        }

        // ---- scale + mask, online softmax (log2 domain)
        float mx[2] = {-INFINITY, -INFINITY};
        if (plain) {
#pragma unroll
            for (int j = 0; j < KT / 8; ++j) {
                s[j][0] *= p.scale_log2; s[j][1] *= p.scale_log2; s[j][2] *= p.scale_log2; s[j][3] *= p.scale_log2;
                mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
            }
        } else {
            const float* mt = sM + buf * KT;
#pragma unroll
            for (int j = 0; j < KT / 8; ++j) {
                const float2 m01 = *reinterpret_cast<const float2*>(mt + 8 * j + 2 * t);
                // masked keys: the score is absorbed by the huge constant exactly as in the fp32 reference
                s[j][0] = (m01.x == 0.f) ? s[j][0] * p.scale_log2 : m01.x;
                s[j][1] = (m01.y == 0.f) ? s[j][1] * p.scale_log2 : m01.y;
                s[j][2] = (m01.x == 0.f) ? s[j][2] * p.scale_log2 : m01.x;
                s[j][3] = (m01.y == 0.f) ? s[j][3] * p.scale_log2 : m01.y;
                mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
            }
        }
        float alpha[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);   // finite: key 0 of tile 0 always exists
            alpha[r] = ex2_approx(m_run[r] - m_new);      // first tile: ex2(-inf) = 0
            m_run[r] = m_new;
            l_run[r] *= alpha[r];
        }
#pragma unroll
        for (int j = 0; j < KT / 8; ++j) {
            s[j][0] = ex2_approx(s[j][0] - m_run[0]);
            s[j][1] = ex2_approx(s[j][1] - m_run[0]);
            s[j][2] = ex2_approx(s[j][2] - m_run[1]);
            s[j][3] = ex2_approx(s[j][3] - m_run[1]);
            l_run[0] += s[j][0] + s[j][1];
            l_run[1] += s[j][2] + s[j][3];
        }
        if (DROP) {
            // the row sum above is that of the un-dropped probabilities (softmax first, dropout second, :250-258)
            const DropoutParams drop = dropout_resolve(p.drop);
            const unsigned long long row0 =
                static_cast<unsigned long long>(p.reverse ? num_items - 1 - static_cast<int>(blockIdx.x + it * gridDim.x)
                                                          : static_cast<int>(blockIdx.x + it * gridDim.x)) * p.nq + warp * 16 + g4;
#pragma unroll
            for (int kb = 0; kb < (KT + 31) / 32; ++kb) {
                const uint32_t group = static_cast<uint32_t>(((tile * KT) >> 5) + kb) * 4 + t;
                const uint32_t k0 = dropout_keep8(drop, row0, group);
                const uint32_t k1 = dropout_keep8(drop, row0 + 8, group);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int j = kb * 4 + nt;
                    if (j < KT / 8) {
                        s[j][0] = ((k0 >> (2 * nt)) & 1u) ? s[j][0] * drop.scale : 0.f;
                        s[j][1] = ((k0 >> (2 * nt + 1)) & 1u) ? s[j][1] * drop.scale : 0.f;
                        s[j][2] = ((k1 >> (2 * nt)) & 1u) ? s[j][2] * drop.scale : 0.f;
                        s[j][3] = ((k1 >> (2 * nt + 1)) & 1u) ? s[j][3] * drop.scale : 0.f;
                    }
                }
            }
        }
        if (tile > 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
                o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
            }
        }

        // ---- O += P V
        const uint32_t vaddr = smem_u32(sV + buf * KT * 128);
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
            uint32_t a[4];
            a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
            a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
            a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                uint32_t b0, b1, b2, b3;
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = jj * 2 + (lane >> 4);
                ldmatrix_x4_trans(vaddr + swz128(r, c), b0, b1, b2, b3);
                mma_bf16_16816(o[2 * jj], a, b0, b1);
                mma_bf16_16816(o[2 * jj + 1], a, b2, b3);
            }
        }

        if (tile == ntiles - 1) {
            // ---- normalise, stage this warp's 16 context rows (swizzled like a TMA tile) and hand them to the TMA
            //      store engine: head-merged [B*nq, H] output, rows beyond nq are clipped by the tensor map
            const int w0 = blockIdx.x + it * gridDim.x;
            const int w = p.reverse ? num_items - 1 - w0 : w0;
            const int b = w / p.num_heads, h = w - b * p.num_heads;
            float inv[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float l = l_run[r];
                l += __shfl_xor_sync(0xffffffffu, l, 1);
                l += __shfl_xor_sync(0xffffffffu, l, 2);
                inv[r] = 1.0f / l;
            }
            uint8_t* so = sO + ((it & 1) * nq_pad + warp * 16) * 128;
            if (lane == 0) tma_store_wait_read<1>();    // the store issued from this slot two items ago has read it
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                *reinterpret_cast<uint32_t*>(so + swz128(g4, j) + 4 * t) = pack_bf16(o[j][0] * inv[0], o[j][1] * inv[0]);
                *reinterpret_cast<uint32_t*>(so + swz128(g4 + 8, j) + 4 * t) = pack_bf16(o[j][2] * inv[1], o[j][3] * inv[1]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&tmap_o, so, h * 64, warp * 16, b);
                tma_store_commit();
            }
        }
    }
    if (lane == 0) tma_store_wait_all();    // global writes complete before the CTA exits
}

template <class K>
static int attention_ctas_per_sm(K kern, int threads, size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

constexpr int ATTENTION_PP_DEFAULT = 0;
// tcgen05 kernel for long key sequences (attention_tc.cu)
bool attention_tc_supported(long long num_heads, long long nq, long long nk, long long head_dim, long long ldk,
                            long long ldv, long long kv_batch_rows);
int attention_tc(const void* q, long long ldq, long long q_batch_rows, const void* k, long long ldk, const void* v,
                 long long ldv, long long kv_batch_rows, const float* key_mask, void* out, long long ldo, long long batch,
                 long long num_heads, long long nq, long long nk, float scale, cudaStream_t stream);

// two softmax groups on alternate key tiles (attention_pp.cu), more than 128 keys
bool attention_pp_supported(long long num_heads, long long nq, long long nk, long long head_dim, long long ldk,
                            long long ldv, long long kv_batch_rows);
int attention_pp(const void* q, long long ldq, long long q_batch_rows, const void* k, long long ldk, const void* v,
                 long long ldv, long long kv_batch_rows, const float* key_mask, void* out, long long ldo, long long batch,
                 long long num_heads, long long nq, long long nk, float scale, cudaStream_t stream);

static bool attention_pp_enabled() {
    // UNIREC_ATTENTION_PP=1 / 0: long key sequences on the two-group kernel / on attention_tc_kernel (A/B measurements)
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("UNIREC_ATTENTION_PP");
        on = e != nullptr ? (e[0] != '0' ? 1 : 0) : ATTENTION_PP_DEFAULT;
    }
    return on == 1;
}

static bool attention_tc_enabled() {
    // UNIREC_ATTENTION_TC=0 forces the mma.sync kernel for every shape (A/B measurements and tests)
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("UNIREC_ATTENTION_TC");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

int attention(const void* q, long long ldq, long long q_batch_rows, const void* k, long long ldk, const void* v,
              long long ldv, long long kv_batch_rows, const float* key_mask, void* out, long long ldo, long long batch,
              long long num_heads, long long nq, long long nk, long long head_dim, float scale, unsigned drop_thr16,
              unsigned long long drop_seed, unsigned drop_site, const unsigned long long* drop_seed_offset,
              cudaStream_t stream) {
    if (q == nullptr || k == nullptr || v == nullptr || out == nullptr || batch <= 0 || num_heads <= 0 || nq <= 0 ||
        nk <= 0) {
        set_last_error("attention: null pointer or empty shape");
        return UNIREC_ERR_BAD_ARG;
    }
    if (head_dim != 64 || nq > 64 || ldq % 8 != 0 || ldk % 8 != 0 || ldv % 8 != 0 || ldo % 8 != 0 ||
        batch * num_heads > 2147483647LL) {
        set_last_error("attention: supports head_dim 64, <= 64 queries, 16-byte aligned rows (head_dim=%lld nq=%lld)",
                       head_dim, nq);
        return UNIREC_ERR_BAD_ARG;
    }
    if (drop_thr16 >= 65536u) {
        set_last_error("attention: dropout probability must be < 1");
        return UNIREC_ERR_BAD_ARG;
    }
    if (drop_thr16 == 0 && attention_tc_enabled() && attention_pp_enabled() &&
        attention_pp_supported(num_heads, nq, nk, head_dim, ldk, ldv, kv_batch_rows))
        return attention_pp(q, ldq, q_batch_rows, k, ldk, v, ldv, kv_batch_rows, key_mask, out, ldo, batch, num_heads, nq,
                            nk, scale, stream);
    if (drop_thr16 == 0 && attention_tc_enabled() && attention_tc_supported(num_heads, nq, nk, head_dim, ldk, ldv, kv_batch_rows))
        return attention_tc(q, ldq, q_batch_rows, k, ldk, v, ldv, kv_batch_rows, key_mask, out, ldo, batch, num_heads, nq,
                            nk, scale, stream);
    if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
        (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_last_error("attention: q / k / v / out must be 16-byte aligned");
        return UNIREC_ERR_BAD_ARG;
    }
    AttnParams p;
    p.key_mask = key_mask;
    p.num_heads = static_cast<int>(num_heads); p.nq = static_cast<int>(nq); p.nk = static_cast<int>(nk);
    p.q_broadcast = q_batch_rows == 0 ? 1 : 0;
    p.reverse = stream_reverse() ? 1 : 0;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.drop.thr16 = drop_thr16; p.drop.seed = drop_seed; p.drop.site = drop_site; p.drop.seed_offset = drop_seed_offset;
    p.drop.scale = 65536.0f / (65536.0f - static_cast<float>(drop_thr16));
    const int nwarps = static_cast<int>((nq + 15) / 16);
    const int threads = nwarps * 32;
    const int num_items = static_cast<int>(batch * num_heads);
    const int kt = nk <= 16 ? 16 : (nk <= 32 ? 32 : 64);
    const size_t smem = 1024 + static_cast<size_t>(ATT_STAGES) * (static_cast<size_t>(nwarps) * 16 * 128 +
                                                                  2 * static_cast<size_t>(kt) * 128 + kt * sizeof(float)) +
                        2 * static_cast<size_t>(nwarps) * 16 * 128 + ATT_STAGES * sizeof(uint64_t);
    const long long hd = num_heads * 64;
    CUtensorMap tq, tk, tv, to;
    int rc = make_tmap_bf16_3d(&tq, q, q_batch_rows == 0 ? 1 : batch, nq, hd, ldq, q_batch_rows * ldq, 16);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_3d(&tk, k, batch, nk, hd, ldk, kv_batch_rows * ldk, kt);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_3d(&tv, v, batch, nk, hd, ldv, kv_batch_rows * ldv, kt);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_3d(&to, out, batch, nq, hd, ldo, nq * ldo, 16);
    if (rc != UNIREC_OK) return rc;
    const int sms = num_sms() > 0 ? num_sms() : 148;
#define UNIREC_ATT(KT_, DROP_)                                                                                     \
    do {                                                                                                           \
        static bool attr_set = false;                                                                              \
        if (!attr_set) {                                                                                           \
            if (cudaFuncSetAttribute(attention_kernel<KT_, DROP_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     100 * 1024) != cudaSuccess) {                                                 \
                set_last_error("attention: cudaFuncSetAttribute failed");                                          \
                return UNIREC_ERR_CUDA;                                                                            \
            }                                                                                                      \
            attr_set = true;                                                                                       \
        }                                                                                                          \
        long long grid = static_cast<long long>(sms) * attention_ctas_per_sm(attention_kernel<KT_, DROP_>, threads, smem); \
        if (grid > num_items) grid = num_items;                                                                    \
        attention_kernel<KT_, DROP_><<<static_cast<unsigned>(grid), threads, smem, stream>>>(tq, tk, tv, to, p, num_items); \
    } while (0)
    if (drop_thr16 == 0) {
        if (kt == 16) UNIREC_ATT(16, false);
        else if (kt == 32) UNIREC_ATT(32, false);
        else UNIREC_ATT(64, false);
    } else {
        if (kt == 16) UNIREC_ATT(16, true);
        else if (kt == 32) UNIREC_ATT(32, true);
        else UNIREC_ATT(64, true);
    }
#undef UNIREC_ATT
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("attention launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

}  // namespace unirec
