// User cross-attention with the K/V projection fused in (sm_100a) - SURVEY.md K2 "ideally fused into K5", section 8f-2:
//
//     ctx[u, q, h*64:(h+1)*64] = softmax_k( Q[u,q,h,:] . (X[u,k,:] Wk_h^T) / 8 + mask[u,k] ) @ (X[u,k,:] Wv_h^T) + bv_h
//
// The reference projects the whole user sequence to K and V (models/qformer.py:185-188: `self.key(encoder_hidden_states)`,
// `self.value(...)`, 74 % of the user model's FLOPs) and then attends (:205, :244-268).  The materialised path does the
// same with two kernels and a 13.4 GB K/V buffer per 512 users that is written to and re-read from HBM.  Here ONE kernel
// per layer produces a K/V tile on the tensor cores and consumes it on the spot - no K, no V ever reaches HBM:
//
//   * main loop = the CTA-pair tcgen05 GEMM of gemm_cg2.cu, unchanged: a cluster of two CTAs owns a 256-key x 256-column
//     tile, 5-stage TMA ring, UMMA 256 x 256 x 16 (cta_group::2), two TMEM accumulator stages.  The layer's K/V weight is
//     packed so that column block j holds [K of heads 2j, 2j+1 | V of heads 2j, 2j+1]: after the main loop each CTA has,
//     for ITS 128 keys, the complete K and V rows of one head pair in TMEM.
//   * epilogue (8 warps per CTA): tcgen05.ld -> bf16 -> the shared C tile (128-byte swizzled rows, as the GEMM stages it
//     for its TMA store) - and instead of storing it, the same warps run flash attention over it: warp w owns head
//     (w / 4) of the pair and 16 of the 64 queries, Q fragments live in registers, S = Q K^T and O += P V on mma.sync
//     m16n8k16 from ldmatrix fragments of the shared tile, online softmax in the log2 domain (fp32 statistics).
//     A CTA walks the keys of a group of users in order, so the running (max, sum, O) of a (user, head) stays in
//     registers across tiles and is written once per (user, head, CTA of the pair): 16 KB instead of the 410 KB of K/V
//     rows it stands for.  `kv_attention_combine_kernel` merges the two CTAs' partials, normalises and adds bv.
//
// Algebra used: the key bias shifts every score of a query row by the same amount (q . bk) - softmax-invariant, dropped;
// the value bias commutes with the convex combination (+ bv after normalisation).  Mask semantics as in attention.cu:
// masked keys score -1e30 (log2 domain), an all-masked row is uniform over all keys.
//
// Work decomposition: an item = (group of G users whose G * S rows are a whole number of 256-row tiles, head pair j);
// the persistent cluster grid walks items; rows beyond the last user are zero-filled by TMA and skipped.
#include "common.cuh"
#include "cg2_ptx.cuh"
#include "umma_pipe.cuh"

#include <cstdlib>
#include <cstring>

namespace unirec {

constexpr int KA_TILE = 256;
constexpr int KA_BLOCK_K = 64;
constexpr int KA_STAGES = 5;
constexpr int KA_A_BYTES = 128 * KA_BLOCK_K * 2;
constexpr int KA_B_BYTES = 128 * KA_BLOCK_K * 2;
constexpr int KA_STAGE_BYTES = KA_A_BYTES + KA_B_BYTES;
constexpr int KA_SLAB_BYTES = 128 * 64 * 2;                 // 128 keys x 64 dims, 128-byte rows
constexpr int KA_SLABS = 4;                                 // K head a, K head b, V head a, V head b
constexpr int KA_THREADS = 384;
constexpr int KA_EPI_WARPS = 8;
constexpr int KA_BARRIER_BYTES = 512;
constexpr int KA_SMEM_BYTES = KA_STAGES * KA_STAGE_BYTES + KA_SLABS * KA_SLAB_BYTES + 1024 + KA_BARRIER_BYTES;
constexpr int KA_TMEM_COLS = 512;
constexpr int KA_NQ = 64;                                   // queries per user (UserQFormer.num_query_tokens)
constexpr float KA_MASKED_LOG2 = -1.0e30f;
static_assert(KA_SMEM_BYTES <= 232448, "shared memory budget exceeded");

struct KvAttnParams {
    int K;                       // contraction length (encoder width)
    int M;                       // valid rows = users * S
    int S;                       // keys per user (multiple of 64)
    int users_per_item;          // G
    int tiles_per_item;          // G * S / 256
    int num_items;               // ceil(users / G) * n_blocks
    int n_blocks;                // head pairs = num_heads / 2
    int num_heads;
    const __nv_bfloat16* q;      // projected queries [users * 64 (or 64), ldq]
    long long ldq;
    int q_batch_rows;            // 64, or 0 when one set of queries serves every user
    const float* key_mask;       // [users, S] (1 attend / 0 masked) or nullptr
    float scale_log2;
    float* o_part;               // [(user * heads + head) * 2 + cta][64][64] unnormalised context
    float* ml_part;              // [(user * heads + head) * 2 + cta][2][64] running max (log2 domain), running sum
    int debug;                   // timing experiments only (UNIREC_KV_DEBUG): bit 0 skips the S MMAs, bit 1 the PV MMAs
};

UNIREC_DEVICE void epi_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(KA_EPI_WARPS * 32) : "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(KA_THREADS, 1)
kv_attention_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                          const KvAttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + KA_STAGES * KA_A_BYTES;
    uint8_t* smem_c = smem + KA_STAGES * KA_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + KA_SLABS * KA_SLAB_BYTES);
    uint64_t* full_bar = bars;                          // [STAGES]  used in the leader
    uint64_t* empty_bar = bars + KA_STAGES;             // [STAGES]  one per CTA
    uint64_t* tmem_full_bar = bars + 2 * KA_STAGES;     // [2]       one per CTA
    uint64_t* tmem_empty_bar = bars + 2 * KA_STAGES + 2;   // [2]    used in the leader
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * KA_STAGES + 4);

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < KA_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 2 * KA_EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp_idx == 2) {
        tmem_alloc_cg2(tmem_ptr_smem, KA_TMEM_COLS);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_kb = p.K / KA_BLOCK_K;
    const int item_rows = p.users_per_item * p.S;

    if (warp_idx == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int item = cluster_id; item < p.num_items; item += num_clusters) {
            const int grp = item / p.n_blocks, j = item - grp * p.n_blocks;
            const int n_coord = j * KA_TILE + static_cast<int>(cta_rank) * 128;
            for (int t = 0; t < p.tiles_per_item; ++t) {
                const int m_coord = grp * item_rows + t * KA_TILE + static_cast<int>(cta_rank) * 128;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (lane == 0) {
                        const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * KA_STAGE_BYTES);
                        tma_load_2d_cg2(&tmap_x, full_leader, smem_a + stage * KA_A_BYTES, kb * KA_BLOCK_K, m_coord,
                                        kCacheEvictNormal);
                        tma_load_2d_cg2(&tmap_w, full_leader, smem_b + stage * KA_B_BYTES, kb * KA_BLOCK_K, n_coord,
                                        kCacheEvictNormal);
                    }
                    __syncwarp();
                    if (++stage == KA_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== UMMA issuer (leader CTA only) =====================
        if (is_leader) {
            constexpr uint32_t idesc = umma_idesc_bf16(KA_TILE, KA_TILE);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t iter = 0;
            for (int item = cluster_id; item < p.num_items; item += num_clusters) {
                for (int t = 0; t < p.tiles_per_item; ++t, ++iter) {
                    const uint32_t as = iter & 1u;
                    const uint32_t aphase = (iter >> 1) & 1u;
                    mbar_wait_cluster(&tmem_empty_bar[as], aphase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * KA_TILE;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint32_t a_addr = smem_u32(smem_a + stage * KA_A_BYTES);
                            const uint32_t b_addr = smem_u32(smem_b + stage * KA_B_BYTES);
#pragma unroll
                            for (int k = 0; k < KA_BLOCK_K / 16; ++k) {
                                const uint64_t da = umma_smem_desc_sw128(a_addr + k * 32);
                                const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32);
                                umma_bf16_ss_cg2(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                            umma_commit_cg2_mc(&empty_bar[stage], 0x3);
                            if (kb == num_kb - 1) umma_commit_cg2_mc(&tmem_full_bar[as], 0x3);
                        }
                        __syncwarp();
                        if (++stage == KA_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue: drain the K/V tile, then attend over it (both CTAs) =====================
        const int ew = warp_idx - 4;
        const int quad = warp_idx & 3;           // TMEM lane quadrant this warp may access
        const int half = ew >> 2;                // drain: column half (0 = K slabs, 1 = V slabs); attention: head of the pair
        const int r = quad * 32 + lane;          // drain: key row inside this CTA's 128-key tile
        const int qr = (ew & 3) * 16;            // attention: first of this warp's 16 query rows
        const int g4 = lane >> 2, t4 = lane & 3;
        const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
        const uint32_t k_slab = smem_u32(smem_c + half * KA_SLAB_BYTES);
        const uint32_t v_slab = smem_u32(smem_c + (2 + half) * KA_SLAB_BYTES);

        uint32_t qf[4][4];
        float o[8][4];
        float m_run[2], l_run[2];
        uint32_t iter = 0;

        for (int item = cluster_id; item < p.num_items; item += num_clusters) {
            const int grp = item / p.n_blocks, j = item - grp * p.n_blocks;
            const int head = 2 * j + half;
            int cur_u = -1;

            auto flush = [&](int u) {
                // unnormalised partial of (user u, head, this CTA): O rows, running max, running sum
                const long long base = (static_cast<long long>(u) * p.num_heads + head) * 2 + cta_rank;
                float* op = p.o_part + base * (KA_NQ * 64);
#pragma unroll
                for (int jn = 0; jn < 8; ++jn) {
                    *reinterpret_cast<float2*>(op + (qr + g4) * 64 + 8 * jn + 2 * t4) = make_float2(o[jn][0], o[jn][1]);
                    *reinterpret_cast<float2*>(op + (qr + g4 + 8) * 64 + 8 * jn + 2 * t4) = make_float2(o[jn][2], o[jn][3]);
                }
                float l0 = l_run[0], l1 = l_run[1];
                l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
                l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
                l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
                l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
                if (t4 == 0) {
                    float* ml = p.ml_part + base * (2 * KA_NQ);
                    ml[qr + g4] = m_run[0];
                    ml[qr + g4 + 8] = m_run[1];
                    ml[KA_NQ + qr + g4] = l0;
                    ml[KA_NQ + qr + g4 + 8] = l1;
                }
            };

            for (int t = 0; t < p.tiles_per_item; ++t, ++iter) {
                const int m_coord = grp * item_rows + t * KA_TILE + static_cast<int>(cta_rank) * 128;
                const uint32_t as = iter & 1u;
                const uint32_t aphase = (iter >> 1) & 1u;
                mbar_wait(&tmem_full_bar[as], aphase);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + as * KA_TILE + (static_cast<uint32_t>(quad * 32) << 16);
                // ---- drain: this warp's 32 keys x 128 columns (its column half) -> bf16 -> two 64-column slabs
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    uint8_t* slab_smem = smem_c + (half * 2 + s) * KA_SLAB_BYTES;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_acc + half * 128 + s * 64 + c * 32, v);
                        tmem_ld_wait();
                        if (s == 1 && c == 1) {
                            // every TMEM read of this accumulator stage is in registers: hand it back to the issuer
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(tmem_empty_leader0 + as * 8);
                        }
#pragma unroll
                        for (int jn = 0; jn < 4; ++jn)
                            *reinterpret_cast<uint4*>(slab_smem + swz128(r, c * 4 + jn)) = make_uint4(
                                pack_bf16(__uint_as_float(v[8 * jn]), __uint_as_float(v[8 * jn + 1])),
                                pack_bf16(__uint_as_float(v[8 * jn + 2]), __uint_as_float(v[8 * jn + 3])),
                                pack_bf16(__uint_as_float(v[8 * jn + 4]), __uint_as_float(v[8 * jn + 5])),
                                pack_bf16(__uint_as_float(v[8 * jn + 6]), __uint_as_float(v[8 * jn + 7])));
                    }
                }
                epi_bar_sync(1);                 // the whole K/V tile of this CTA is in shared memory

                // ---- attention over this CTA's 128 keys, as two 64-key halves (a half never straddles users: S % 64 == 0)
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    const int m0 = m_coord + 64 * hf;
                    if (m0 >= p.M) continue;
                    const int u = m0 / p.S;
                    const int p0 = m0 - u * p.S;
                    if (u != cur_u) {
                        if (cur_u >= 0) flush(cur_u);
                        cur_u = u;
                        // Q fragments (A operand of m16n8k16: rows g4 / g4 + 8, columns 2 t4 (+ 8) of each 16-wide k step)
                        const __nv_bfloat16* qp = p.q + (static_cast<long long>(u) * p.q_batch_rows + qr + g4) * p.ldq +
                                                  head * 64 + 2 * t4;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            qf[kk][0] = __ldg(reinterpret_cast<const uint32_t*>(qp + kk * 16));
                            qf[kk][1] = __ldg(reinterpret_cast<const uint32_t*>(qp + 8 * p.ldq + kk * 16));
                            qf[kk][2] = __ldg(reinterpret_cast<const uint32_t*>(qp + kk * 16 + 8));
                            qf[kk][3] = __ldg(reinterpret_cast<const uint32_t*>(qp + 8 * p.ldq + kk * 16 + 8));
                        }
#pragma unroll
                        for (int jn = 0; jn < 8; ++jn) { o[jn][0] = o[jn][1] = o[jn][2] = o[jn][3] = 0.f; }
                        m_run[0] = m_run[1] = -INFINITY;
                        l_run[0] = l_run[1] = 0.f;
                    }
                    // key mask of the 64 keys as two ballot words (bit = attend)
                    uint32_t mlo = 0xffffffffu, mhi = 0xffffffffu;
                    if (p.key_mask != nullptr) {
                        const float* mp = p.key_mask + static_cast<long long>(u) * p.S + p0;
                        mlo = __ballot_sync(0xffffffffu, __ldg(mp + lane) != 0.f);
                        mhi = __ballot_sync(0xffffffffu, __ldg(mp + 32 + lane) != 0.f);
                    }

                    // S = Q K^T (16 x 64 per warp)
                    float s[8][4];
#pragma unroll
                    for (int jn = 0; jn < 8; ++jn) { s[jn][0] = s[jn][1] = s[jn][2] = s[jn][3] = 0.f; }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            uint32_t b0, b1, b2, b3;
                            const int rr = hf * 64 + jj * 16 + (lane & 7) + (lane >> 4) * 8;
                            const int cc = kk * 2 + ((lane >> 3) & 1);
                            ldmatrix_x4(k_slab + swz128(rr, cc), b0, b1, b2, b3);
                            mma_bf16_16816(s[2 * jj], qf[kk], b0, b1);
                            mma_bf16_16816(s[2 * jj + 1], qf[kk], b2, b3);
                        }
                    }
                    // scale + mask, online softmax (log2 domain)
                    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
                    for (int jn = 0; jn < 8; ++jn) {
                        const int key = 8 * jn + 2 * t4;                       // and key + 1
                        const uint32_t w = (jn < 4) ? mlo : mhi;
                        const bool a0 = (w >> (key & 31)) & 1u, a1 = (w >> ((key + 1) & 31)) & 1u;
                        s[jn][0] = a0 ? s[jn][0] * p.scale_log2 : KA_MASKED_LOG2;
                        s[jn][1] = a1 ? s[jn][1] * p.scale_log2 : KA_MASKED_LOG2;
                        s[jn][2] = a0 ? s[jn][2] * p.scale_log2 : KA_MASKED_LOG2;
                        s[jn][3] = a1 ? s[jn][3] * p.scale_log2 : KA_MASKED_LOG2;
                        mx[0] = fmaxf(mx[0], fmaxf(s[jn][0], s[jn][1]));
                        mx[1] = fmaxf(mx[1], fmaxf(s[jn][2], s[jn][3]));
                    }
                    float alpha[2];
#pragma unroll
                    for (int rw = 0; rw < 2; ++rw) {
                        mx[rw] = fmaxf(mx[rw], __shfl_xor_sync(0xffffffffu, mx[rw], 1));
                        mx[rw] = fmaxf(mx[rw], __shfl_xor_sync(0xffffffffu, mx[rw], 2));
                        const float m_new = fmaxf(m_run[rw], mx[rw]);          // finite: -1e30 at worst
                        alpha[rw] = ex2_approx(m_run[rw] - m_new);             // first half of a user: ex2(-inf) = 0
                        m_run[rw] = m_new;
                        l_run[rw] *= alpha[rw];
                    }
#pragma unroll
                    for (int jn = 0; jn < 8; ++jn) {
                        s[jn][0] = ex2_approx(s[jn][0] - m_run[0]);
                        s[jn][1] = ex2_approx(s[jn][1] - m_run[0]);
                        s[jn][2] = ex2_approx(s[jn][2] - m_run[1]);
                        s[jn][3] = ex2_approx(s[jn][3] - m_run[1]);
                        l_run[0] += s[jn][0] + s[jn][1];
                        l_run[1] += s[jn][2] + s[jn][3];
                        o[jn][0] *= alpha[0]; o[jn][1] *= alpha[0];
                        o[jn][2] *= alpha[1]; o[jn][3] *= alpha[1];
                    }
                    // O += P V
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        uint32_t a[4];
                        a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
                        a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
                        a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
                        a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            uint32_t b0, b1, b2, b3;
                            const int rr = hf * 64 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                            const int cc = jj * 2 + (lane >> 4);
                            ldmatrix_x4_trans(v_slab + swz128(rr, cc), b0, b1, b2, b3);
                            mma_bf16_16816(o[2 * jj], a, b0, b1);
                            mma_bf16_16816(o[2 * jj + 1], a, b2, b3);
                        }
                    }
                }
                epi_bar_sync(2);                 // every warp is done with the tile: the next drain may overwrite it
            }
            if (cur_u >= 0) flush(cur_u);
        }
    }

    // ---- teardown: nobody may exit (or free TMEM) while the pair still uses this CTA's smem / barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, KA_TMEM_COLS);
    }
}

// =====================================================================================================================
// Variant 2: the attention itself on tcgen05 too (`kv_attention_umma_kernel`).
//
// ncu on the mma.sync epilogue above (profiles/r02_b_ncu_stalls_kvattn.txt): 34.5 % of all stall samples are "math pipe
// throttle" on HMMA - the legacy tensor path shares the tensor cores with the UMMA main loop and is an order of magnitude
// slower per flop, so the 6 % of attention FLOPs cost more time than the 94 % of projection FLOPs they ride on (tensor pipe
// 59 % active, 30.1 ms per layer instead of the 20.3 ms of the bare projection).  Here S = Q K^T and O = P V are UMMAs as
// well.  The projection stays a cta_group::2 tile; the attention MMAs are cta_group::1 - each CTA of the pair attends over
// ITS 128 keys with its own tensor core, shared memory and TMEM (tools/probe_mixed_cta_group.cu shows that the two kinds
// coexist in one kernel on this part).
//
//   * TMEM: no extra columns.  A drained projection accumulator stage (256 columns) is idle until the projection of the
//     tile after next starts, so the attention of the tile lives in it: S of the two 64-key halves in columns 0..127, O of
//     the two halves in 128..255; the stage is handed back to the projection when the O rows have been read.
//   * a 64-key half never straddles two users (S % 64 == 0), so every half is attended INDEPENDENTLY with that user's
//     queries: thread (row = (head, query), half) owns the running (max, sum, O[64]) of its split of the user's keys in
//     registers - no exchange between threads; four partials per (user, head) (2 CTAs x 2 halves) are merged by
//     `kv_attention_combine_kernel`.
//   * heads: the two heads of the pair are stacked in the M dimension (rows 0-63 head a, 64-127 head b) WITHOUT the
//     block-diagonal zero padding of attention_tc.cu: each S / PV is two MMAs of contraction 64 that write only their
//     head's 64 TMEM lanes (tcgen05.mma disable-output-lane mask), so the stacked queries are 16 KB per user and O is 64
//     columns per half.
//   * P (bf16) is written over the K slabs (dead once S has been accumulated) as the K-major A operand of PV; V is used
//     in place as the MN-major B operand.
//   * two issuer threads: warp 1 of the leader runs the projection main loop exactly as in gemm_cg2.cu; warp 3 of EACH CTA
//     issues the S / PV MMAs of its CTA's current tile as the epilogue warps hand them over.
// =====================================================================================================================
constexpr int KU_STAGES = 4;
constexpr int KU_THREADS = 384;                             // producer, two issuers, TMEM allocator, 8 epilogue warps
constexpr int KU_Q_BYTES = 128 * 64 * 2;                     // stacked queries of one user (rows 0-63 head a, 64-127 head b)
constexpr int KU_SMEM_BYTES = KU_STAGES * KA_STAGE_BYTES + KA_SLABS * KA_SLAB_BYTES + 2 * KU_Q_BYTES + 1024 + KA_BARRIER_BYTES;
static_assert(KU_SMEM_BYTES <= 232448, "shared memory budget exceeded");

// MN-major SW128 B operand: [K rows of 128 B][64 MN elements]; 8-row groups 1024 B apart (attention_tc.cu)
UNIREC_DEVICE uint64_t ku_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>(16384u >> 4) << 16;           // LBO: unused for a single 64-element MN block
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// cta_group::1 MMA that leaves the TMEM lanes of `skip_upper ? 64..127 : 0..63` untouched
UNIREC_DEVICE void umma_bf16_ss_lanes(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                      bool skip_upper) {
    const uint32_t lo = skip_upper ? 0u : 0xffffffffu, hi = skip_upper ? 0xffffffffu : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(lo), "r"(lo), "r"(hi), "r"(hi)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(KU_THREADS, 1)
kv_attention_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const KvAttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + KU_STAGES * KA_A_BYTES;
    uint8_t* smem_c = smem + KU_STAGES * KA_STAGE_BYTES;        // slabs: K_a, K_b (later P_0, P_1), V_a, V_b
    uint8_t* smem_q = smem_c + KA_SLABS * KA_SLAB_BYTES;        // [2] stacked queries, buffer = user & 1
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_q + 2 * KU_Q_BYTES);
    uint64_t* full_bar = bars;                          // [STAGES]  used in the leader
    uint64_t* empty_bar = bars + KU_STAGES;             // [STAGES]  one per CTA
    uint64_t* tmem_full_bar = bars + 2 * KU_STAGES;     // [2]       one per CTA
    uint64_t* tmem_empty_bar = bars + 2 * KU_STAGES + 2;   // [2]    used in the leader
    uint64_t* kv_ready = bars + 2 * KU_STAGES + 4;      // epilogue warps -> this CTA's issuer: K/V slabs (+ queries) written
    uint64_t* s_full = bars + 2 * KU_STAGES + 5;        // issuer (commit): S of both halves accumulated
    uint64_t* p_ready = bars + 2 * KU_STAGES + 6;       // [2] the four warps of a half -> issuer: P written
    uint64_t* pv_done = bars + 2 * KU_STAGES + 8;       // [2] issuer (commit): O of the half accumulated
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * KU_STAGES + 10);

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < KU_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 2 * KA_EPI_WARPS);
            mbar_init(&p_ready[i], KA_EPI_WARPS / 2);
            mbar_init(&pv_done[i], 1);
        }
        mbar_init(kv_ready, KA_EPI_WARPS);
        mbar_init(s_full, 1);
        fence_mbar_init();
    }
    if (warp_idx == 2) {
        tmem_alloc_cg2(tmem_ptr_smem, KA_TMEM_COLS);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_kb = p.K / KA_BLOCK_K;
    const int item_rows = p.users_per_item * p.S;

    if (warp_idx == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int item = cluster_id; item < p.num_items; item += num_clusters) {
            const int grp = item / p.n_blocks, j = item - grp * p.n_blocks;
            const int n_coord = j * KA_TILE + static_cast<int>(cta_rank) * 128;
            for (int t = 0; t < p.tiles_per_item; ++t) {
                const int m_coord = grp * item_rows + t * KA_TILE + static_cast<int>(cta_rank) * 128;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (lane == 0) {
                        const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * KA_STAGE_BYTES);
                        tma_load_2d_cg2(&tmap_x, full_leader, smem_a + stage * KA_A_BYTES, kb * KA_BLOCK_K, m_coord,
                                        kCacheEvictNormal);
                        tma_load_2d_cg2(&tmap_w, full_leader, smem_b + stage * KA_B_BYTES, kb * KA_BLOCK_K, n_coord,
                                        kCacheEvictNormal);
                    }
                    __syncwarp();
                    if (++stage == KU_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== projection UMMA issuer (leader CTA only): the main loop of gemm_cg2.cu =====================
        if (is_leader) {
            constexpr uint32_t idesc = umma_idesc_bf16(KA_TILE, KA_TILE);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t iter = 0;
            for (int item = cluster_id; item < p.num_items; item += num_clusters) {
                for (int t = 0; t < p.tiles_per_item; ++t, ++iter) {
                    const uint32_t as = iter & 1u;
                    const uint32_t aphase = (iter >> 1) & 1u;
                    // the stage comes back when the attention of the tile before last has read its O rows (both CTAs)
                    mbar_wait_cluster(&tmem_empty_bar[as], aphase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * KA_TILE;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint32_t a_addr = smem_u32(smem_a + stage * KA_A_BYTES);
                            const uint32_t b_addr = smem_u32(smem_b + stage * KA_B_BYTES);
#pragma unroll
                            for (int k = 0; k < KA_BLOCK_K / 16; ++k)
                                umma_bf16_ss_cg2(tmem_d, umma_smem_desc_sw128(a_addr + k * 32),
                                                 umma_smem_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                            umma_commit_cg2_mc(&empty_bar[stage], 0x3);
                            if (kb == num_kb - 1) umma_commit_cg2_mc(&tmem_full_bar[as], 0x3);
                        }
                        __syncwarp();
                        if (++stage == KU_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp_idx == 3) {
        // ===================== attention UMMA issuer (both CTAs, cta_group::1): S and PV of this CTA's tiles =================
        // A thread of its own: the first version polled these barriers from the projection issuer between k-blocks and the
        // polling (mbarrier test / try_wait round trips) throttled the main loop to 41 % tensor-pipe activity (profiles/r02_f).
        // The tensor core executes what the two issuers hand it in arrival order; every dependency is an mbarrier.
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64);                       // both operands K-major
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);         // B (= V) MN-major
        const uint32_t q_addr = smem_u32(smem_q);
        const uint32_t c_addr = smem_u32(smem_c);
        uint32_t iter = 0;
        for (int item = cluster_id; item < p.num_items; item += num_clusters) {
            const int grp = item / p.n_blocks;
            for (int t = 0; t < p.tiles_per_item; ++t, ++iter) {
                const uint32_t ph = iter & 1u;
                const uint32_t tmem_att = tmem_base + (iter & 1u) * KA_TILE;
                const int m_cta = grp * item_rows + t * KA_TILE + static_cast<int>(cta_rank) * 128;
                // user whose keys each 64-key half holds (-1: rows beyond the last user, nothing to attend)
                const int u0 = m_cta < p.M ? m_cta / p.S : -1;
                const int u1 = m_cta + 64 < p.M ? (m_cta + 64) / p.S : -1;
                mbar_wait(kv_ready, ph);
                tc_fence_after();
                if (lane == 0) {
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        const int u = h == 0 ? u0 : u1;
                        if (u < 0 || (p.debug & 1)) continue;
                        const uint32_t a_addr = q_addr + (u & 1) * KU_Q_BYTES;
#pragma unroll 1
                        for (int hd = 0; hd < 2; ++hd) {                                 // head a -> lanes 0-63, head b -> 64-127
                            const uint32_t b_addr = c_addr + hd * KA_SLAB_BYTES + h * (64 * 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_ss_lanes(tmem_att + 64 * h, umma_smem_desc_sw128(a_addr + k * 32),
                                                   umma_smem_desc_sw128(b_addr + k * 32), idesc_s, k != 0 ? 1u : 0u, hd == 0);
                        }
                    }
                    umma_commit(s_full);
                }
                __syncwarp();
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(&p_ready[h], ph);
                    tc_fence_after();
                    if (lane == 0) {
                        if ((h == 0 ? u0 : u1) >= 0 && !(p.debug & 2)) {
                            const uint32_t a_addr = c_addr + h * KA_SLAB_BYTES;          // P_h, written over K slab h
#pragma unroll 1
                            for (int hd = 0; hd < 2; ++hd) {
                                const uint32_t b_addr = c_addr + (2 + hd) * KA_SLAB_BYTES + h * (64 * 128);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_ss_lanes(tmem_att + 128 + 64 * h, umma_smem_desc_sw128(a_addr + k * 32),
                                                       ku_desc_mn_sw128(b_addr + k * (16 * 128)), idesc_pv, k != 0 ? 1u : 0u,
                                                       hd == 0);
                            }
                        }
                        umma_commit(&pv_done[h]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue warps 4..11: drain the K/V tile, softmax, O accumulation (both CTAs) ==============
        const int ew = warp_idx - 4;
        const int quad = warp_idx & 3;           // TMEM lane quadrant this warp may access
        const int half = ew >> 2;                // drain: column half (0 = K slabs, 1 = V slabs); attention: 64-key half
        const int r = quad * 32 + lane;          // drain: key row of this CTA's tile; attention: row (head = r >> 6, query = r & 63)
        const uint32_t lane_field = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
        uint8_t* p_slab = smem_c + half * KA_SLAB_BYTES;            // P of this half (over K slab `half`)

        float o_acc[64];
        float m_run = -INFINITY, l_run = 0.f;
        int cur_u = -1, cur_head = 0;
        int q_have0 = -1, q_have1 = -1;          // user whose queries are in each buffer (valid for the current item)
        uint32_t iter = 0;

        auto flush = [&]() {
            // unnormalised partial of (user, head, split = 2 * cta + half): 64 O values of this thread's query row, max, sum
            const long long base = (static_cast<long long>(cur_u) * p.num_heads + cur_head) * 4 + cta_rank * 2 + half;
            float4* op = reinterpret_cast<float4*>(p.o_part + base * (KA_NQ * 64) + (r & 63) * 64);
#pragma unroll
            for (int i = 0; i < 16; ++i) op[i] = make_float4(o_acc[4 * i], o_acc[4 * i + 1], o_acc[4 * i + 2], o_acc[4 * i + 3]);
            float* ml = p.ml_part + base * (2 * KA_NQ);
            ml[r & 63] = m_run;
            ml[KA_NQ + (r & 63)] = l_run;
        };

        for (int item = cluster_id; item < p.num_items; item += num_clusters) {
            const int grp = item / p.n_blocks, j = item - grp * p.n_blocks;
            q_have0 = q_have1 = -1;
            for (int t = 0; t < p.tiles_per_item; ++t, ++iter) {
                const int m_coord = grp * item_rows + t * KA_TILE + static_cast<int>(cta_rank) * 128;
                const uint32_t as = iter & 1u;
                const uint32_t aphase = (iter >> 1) & 1u;
                const uint32_t ph = iter & 1u;
                mbar_wait(&tmem_full_bar[as], aphase);
                tc_fence_after();
                const uint32_t tmem_stage = tmem_base + as * KA_TILE;
                // ---- drain: this warp's 32 keys x 128 columns (its column half) -> bf16 -> two 64-column slabs
#pragma unroll 1
                for (int s2 = 0; s2 < 2; ++s2) {
                    uint8_t* slab_smem = smem_c + (half * 2 + s2) * KA_SLAB_BYTES;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[16];
                        tmem_ld_32x16(tmem_stage + lane_field + half * 128 + s2 * 64 + c * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int jn = 0; jn < 2; ++jn)
                            *reinterpret_cast<uint4*>(slab_smem + swz128(r, c * 2 + jn)) = make_uint4(
                                pack_bf16(__uint_as_float(v[8 * jn]), __uint_as_float(v[8 * jn + 1])),
                                pack_bf16(__uint_as_float(v[8 * jn + 2]), __uint_as_float(v[8 * jn + 3])),
                                pack_bf16(__uint_as_float(v[8 * jn + 4]), __uint_as_float(v[8 * jn + 5])),
                                pack_bf16(__uint_as_float(v[8 * jn + 6]), __uint_as_float(v[8 * jn + 7])));
                    }
                }
                // ---- stacked queries of the users of this tile's halves (thread = row r, four 16-byte chunks `half`)
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const int m0 = m_coord + 64 * h;
                    if (m0 >= p.M) continue;
                    const int u = m0 / p.S;
                    int& have = (u & 1) ? q_have1 : q_have0;
                    if (have == u) continue;
                    have = u;
                    const __nv_bfloat16* qp = p.q + (static_cast<long long>(u) * p.q_batch_rows + (r & 63)) * p.ldq +
                                              (2 * j + (r >> 6)) * 64;
                    uint8_t* qb = smem_q + (u & 1) * KU_Q_BYTES;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(qb + swz128(r, half * 4 + c)) =
                            __ldg(reinterpret_cast<const uint4*>(qp) + half * 4 + c);
                }
                fence_proxy_async_smem();        // K / V slabs and queries -> visible to the tensor core
                tc_fence_before();               // the drain's TMEM reads are ordered before the S MMAs that overwrite the stage
                __syncwarp();
                if (lane == 0) mbar_arrive(kv_ready);

                // ---- this thread's half: user, key mask
                const int m0 = m_coord + 64 * half;
                const bool valid = m0 < p.M;
                const int u = valid ? m0 / p.S : -1;
                uint32_t mlo = 0xffffffffu, mhi = 0xffffffffu;
                if (valid && p.key_mask != nullptr) {
                    const float* mp = p.key_mask + static_cast<long long>(u) * p.S + (m0 - u * p.S);
                    mlo = __ballot_sync(0xffffffffu, __ldg(mp + lane) != 0.f);
                    mhi = __ballot_sync(0xffffffffu, __ldg(mp + 32 + lane) != 0.f);
                }
                if (valid && (u != cur_u || 2 * j + (r >> 6) != cur_head)) {
                    if (cur_u >= 0) flush();
                    cur_u = u;
                    cur_head = 2 * j + (r >> 6);
                    m_run = -INFINITY;
                    l_run = 0.f;
#pragma unroll
                    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
                }

                mbar_wait(s_full, ph);
                tc_fence_after();
                float alpha = 1.0f;
                if (valid) {
                    const uint32_t tmem_s = tmem_stage + lane_field + 64 * half;
                    // pass 1: row maximum of the 64 scaled (or masked) scores (16 columns at a time: register budget)
                    float mx = -INFINITY;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t sv[16];
                        tmem_ld_32x16(tmem_s + 16 * c, sv);
                        tmem_ld_wait();
                        const uint32_t w = (c < 2 ? mlo : mhi) >> (16 * (c & 1));
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            mx = fmaxf(mx, ((w >> i) & 1u) ? __uint_as_float(sv[i]) * p.scale_log2 : KA_MASKED_LOG2);
                    }
                    const float m_new = fmaxf(m_run, mx);                 // finite: -1e30 at worst
                    alpha = ex2_approx(m_run - m_new);                    // first half of a user: ex2(-inf) = 0
                    m_run = m_new;
                    // pass 2: p = 2^(x - m) -> bf16 -> P row in shared memory (K-major A operand of PV), row sum
                    float psum = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t sv[16];
                        tmem_ld_32x16(tmem_s + 16 * c, sv);
                        tmem_ld_wait();
                        const uint32_t w = (c < 2 ? mlo : mhi) >> (16 * (c & 1));
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const float x0 = ((w >> i) & 1u) ? __uint_as_float(sv[i]) * p.scale_log2 : KA_MASKED_LOG2;
                            const float x1 = ((w >> (i + 1)) & 1u) ? __uint_as_float(sv[i + 1]) * p.scale_log2 : KA_MASKED_LOG2;
                            const float p0 = ex2_approx(x0 - m_new), p1 = ex2_approx(x1 - m_new);
                            psum += p0 + p1;
                            pk[i >> 1] = pack_bf16(p0, p1);
                        }
                        *reinterpret_cast<uint4*>(p_slab + swz128(r, 2 * c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(p_slab + swz128(r, 2 * c + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                    l_run = l_run * alpha + psum;
                }
                fence_proxy_async_smem();        // P -> visible to the tensor core
                tc_fence_before();               // the S reads are ordered before the PV MMAs (and the next tile's projection)
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_ready[half]);

                mbar_wait(&pv_done[half], ph);
                tc_fence_after();
                if (valid) {
                    // O of this half on top of the rescaled running rows
                    const uint32_t tmem_o = tmem_stage + lane_field + 128 + 64 * half;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t ov[16];
                        tmem_ld_32x16(tmem_o + 16 * c, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o_acc[16 * c + i] = fmaf(o_acc[16 * c + i], alpha, __uint_as_float(ov[i]));
                    }
                }
                // both PVs of the tile are complete: P / V slabs may be overwritten by the next drain, and every TMEM read
                // of this stage is in registers: hand the stage back to the projection
                mbar_wait(&pv_done[half ^ 1], ph);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty_leader0 + as * 8);
            }
        }
        if (cur_u >= 0) flush();
    }

    // ---- teardown: nobody may exit (or free TMEM) while the pair still uses this CTA's smem / barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, KA_TMEM_COLS);
    }
}

// ctx[u, q, head * 64 + d] = (sum over the PARTS partials of (user, head), re-referenced to their common maximum) / sum + bv
// PARTS = 2: one partial per CTA of the pair (mma.sync kernel); PARTS = 4: per CTA and 64-key half (UMMA kernel).
template <int PARTS>
__global__ void __launch_bounds__(256)
kv_attention_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part,
                            const float* __restrict__ v_bias, __nv_bfloat16* __restrict__ out, long long ldo,
                            int num_heads) {
    const long long uh = blockIdx.x;                       // user * heads + head
    const int head = static_cast<int>(uh % num_heads);
    const long long u = uh / num_heads;
    const int q = threadIdx.x >> 2, d0 = (threadIdx.x & 3) * 16;
    float mm[PARTS], ll[PARTS];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < PARTS; ++i) {
        const float* ml = ml_part + (uh * PARTS + i) * (2 * KA_NQ);
        mm[i] = ml[q];
        ll[i] = ml[KA_NQ + q];
        if (ll[i] > 0.f) m = fmaxf(m, mm[i]);             // a split that saw none of the user's keys left (0, 0) behind
    }
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    float denom = 0.f;
#pragma unroll
    for (int i = 0; i < PARTS; ++i) {
        if (!(ll[i] > 0.f)) continue;
        const float w = ex2_approx(mm[i] - m);
        denom += ll[i] * w;
        const float* o = o_part + (uh * PARTS + i) * (KA_NQ * 64) + q * 64 + d0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(o) + c);
            acc[4 * c] += x.x * w; acc[4 * c + 1] += x.y * w; acc[4 * c + 2] += x.z * w; acc[4 * c + 3] += x.w * w;
        }
    }
    const float inv = 1.0f / denom;
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float b0 = v_bias != nullptr ? __ldg(v_bias + head * 64 + d0 + 2 * i) : 0.f;
        const float b1 = v_bias != nullptr ? __ldg(v_bias + head * 64 + d0 + 2 * i + 1) : 0.f;
        pk[i] = pack_bf16(acc[2 * i] * inv + b0, acc[2 * i + 1] * inv + b1);
    }
    __nv_bfloat16* op = out + (u * KA_NQ + q) * ldo + head * 64 + d0;
    reinterpret_cast<uint4*>(op)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    reinterpret_cast<uint4*>(op)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

constexpr int KA_MAX_PARTS = 4;

long long kv_attention_workspace_bytes(long long users, long long num_heads) {
    if (users <= 0 || num_heads <= 0) return -1;
    return users * num_heads * KA_MAX_PARTS * (KA_NQ * 64 + 2 * KA_NQ) * static_cast<long long>(sizeof(float));
}

static long long gcd_ll(long long a, long long b) { return b == 0 ? a : gcd_ll(b, a % b); }

// x [users * S, K] bf16 (row stride ldx): the user sequences; w_packed [2 H, K] bf16: the layer's key / value weights
// packed per head pair j as rows [256 j, 256 j + 128) = Wk[128 j : 128 j + 128], rows [256 j + 128, 256 j + 256) =
// Wv[128 j : 128 j + 128]; q [users * 64 (q_batch_rows = 64) or 64 (q_batch_rows = 0), H] bf16 projected queries;
// key_mask fp32 [users, S] or NULL; v_bias fp32 [H] or NULL; out bf16 [users * 64, H] (row stride ldo).
int kv_attention_fused(const void* x, long long ldx, const void* w_packed, long long ldw, const void* q, long long ldq,
                       long long q_batch_rows, const float* key_mask, const float* v_bias, void* out, long long ldo,
                       void* workspace, long long workspace_bytes, long long users, long long S, long long num_heads,
                       long long K, float scale, cudaStream_t stream) {
    if (x == nullptr || w_packed == nullptr || q == nullptr || out == nullptr || workspace == nullptr || users <= 0 ||
        S <= 0 || num_heads <= 0 || K <= 0) {
        set_last_error("kv_attention_fused: null pointer or empty shape");
        return UNIREC_ERR_BAD_ARG;
    }
    if (S % 64 != 0 || num_heads % 2 != 0 || K % KA_BLOCK_K != 0 || ldx % 8 != 0 || ldw % 8 != 0 || ldq % 8 != 0 ||
        ldo % 8 != 0 || (q_batch_rows != 0 && q_batch_rows != KA_NQ) || users * S >= 2147483647LL - 512 ||
        (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w_packed) & 15) ||
        (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_last_error("kv_attention_fused: needs S %% 64 == 0, an even head count, K %% 64 == 0, 64 queries per user, "
                       "16-byte aligned rows (S=%lld heads=%lld K=%lld)", S, num_heads, K);
        return UNIREC_ERR_BAD_ARG;
    }
    if (workspace_bytes < kv_attention_workspace_bytes(users, num_heads) || (reinterpret_cast<uintptr_t>(workspace) & 15)) {
        set_last_error("kv_attention_fused: workspace too small or unaligned (need %lld bytes)",
                       kv_attention_workspace_bytes(users, num_heads));
        return UNIREC_ERR_BAD_ARG;
    }
    const long long H = num_heads * 64;
    KvAttnParams p;
    p.K = static_cast<int>(K);
    p.M = static_cast<int>(users * S);
    p.S = static_cast<int>(S);
    p.users_per_item = static_cast<int>(KA_TILE / gcd_ll(S, KA_TILE));
    p.tiles_per_item = static_cast<int>(static_cast<long long>(p.users_per_item) * S / KA_TILE);
    p.n_blocks = static_cast<int>(num_heads / 2);
    const long long groups = (users + p.users_per_item - 1) / p.users_per_item;
    if (groups * p.n_blocks > 2147483647LL) {
        set_last_error("kv_attention_fused: too many work items");
        return UNIREC_ERR_BAD_ARG;
    }
    p.num_items = static_cast<int>(groups * p.n_blocks);
    p.num_heads = static_cast<int>(num_heads);
    p.q = reinterpret_cast<const __nv_bfloat16*>(q);
    p.ldq = ldq;
    p.q_batch_rows = static_cast<int>(q_batch_rows);
    p.key_mask = key_mask;
    p.scale_log2 = scale * 1.4426950408889634f;
    const char* dbg = getenv("UNIREC_KV_DEBUG");
    p.debug = dbg != nullptr ? atoi(dbg) : 0;
    // variant: "mma_sync" (two partials per (user, head)) or "umma" (S / PV on tcgen05, four partials); UNIREC_KV_ATTENTION_IMPL
    // read per call (tests switch between the two); the mma.sync variant is the faster one in the step (30.1 vs 33.4 ms per
    // layer and 4096 users, profiles/r02_b / r02_g) and therefore the default of this optional path
    const char* env = getenv("UNIREC_KV_ATTENTION_IMPL");
    const int impl = (env != nullptr && strcmp(env, "umma") == 0) ? 1 : 0;
    const int parts = impl == 1 ? 4 : 2;
    p.o_part = reinterpret_cast<float*>(workspace);
    p.ml_part = p.o_part + users * num_heads * parts * (KA_NQ * 64);

    CUtensorMap tx, tw;
    int rc = make_tmap_bf16_2d(&tx, x, users * S, K, ldx, 128);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tw, w_packed, 2 * H, K, ldw, 128);
    if (rc != UNIREC_OK) return rc;

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kv_attention_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             KA_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(kv_attention_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KU_SMEM_BYTES);
        if (e != cudaSuccess) {
            set_last_error("kv_attention_fused: cudaFuncSetAttribute(smem=%d): %s", KA_SMEM_BYTES, cudaGetErrorString(e));
            return UNIREC_ERR_CUDA;
        }
        attr_set = true;
    }
    // partials a split never writes (it saw none of the user's keys) must read as "empty": sum == 0
    cudaError_t e = cudaMemsetAsync(p.ml_part, 0, static_cast<size_t>(users * num_heads * parts * 2 * KA_NQ) * sizeof(float),
                                    stream);
    if (e != cudaSuccess) { set_last_error("kv_attention_fused: memset: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    int clusters = num_sms() / 2;
    if (clusters > p.num_items) clusters = p.num_items;
    if (clusters < 1) clusters = 1;
    if (impl == 1) kv_attention_umma_kernel<<<2 * clusters, KU_THREADS, KU_SMEM_BYTES, stream>>>(tx, tw, p);
    else kv_attention_fused_kernel<<<2 * clusters, KA_THREADS, KA_SMEM_BYTES, stream>>>(tx, tw, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("kv_attention_fused launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    if (impl == 1)
        kv_attention_combine_kernel<4><<<static_cast<unsigned>(users * num_heads), 256, 0, stream>>>(
            p.o_part, p.ml_part, v_bias, reinterpret_cast<__nv_bfloat16*>(out), ldo, p.num_heads);
    else
        kv_attention_combine_kernel<2><<<static_cast<unsigned>(users * num_heads), 256, 0, stream>>>(
            p.o_part, p.ml_part, v_bias, reinterpret_cast<__nv_bfloat16*>(out), ldo, p.num_heads);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("kv_attention_combine launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

}  // namespace unirec
