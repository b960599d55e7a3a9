// Long-key cross-attention on the 5th-generation tensor cores (sm_100a), head_dim = 64, <= 64 queries.
//
//   ctx[b, q, h*64:(h+1)*64] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] * scale + mask[b,k] ) @ V[b,k,h,:]
//
// Used for the user Q-Former's cross-attention (64 queries x up to 1600 keys per user and head;
// models/qformer.py:185-188, 205, 244-268 with encoder_hidden_states = the user sequence).  At 64 flop per K/V
// byte this shape is compute-limited on mma.sync (attention.cu reaches 43 % of HBM bandwidth); here both small
// matmuls run on tcgen05 with accumulators in TMEM and the CUDA cores only do the softmax.
//
// One work item = (batch element, PAIR of adjacent heads).  Two heads are stacked so that every UMMA has M = 128
// (TMEM lane = accumulator row = one softmax thread):
//     S[128 x 128 keys] = Q2[128 x 128] . Kt[128 keys x 128]^T,   Q2 = [[Q_h, 0], [0, Q_h+1]]  (block diagonal)
//     O[128 x 128]     += P[128 x 128 keys] . Vt[128 keys x 128]   (useful blocks: rows 0-63 x cols 0-63 = head h,
//                                                                    rows 64-127 x cols 64-127 = head h+1)
// where Kt / Vt are 128 consecutive keys x the 128 contiguous columns of the two heads (one 256-byte segment per
// key and matrix, fetched by TMA as two 64-column slabs, 128-byte swizzle).  Kt is the K-major B operand of the
// first MMA; Vt is used in place as the MN-major B operand of the second (no transpose anywhere).
//
// Warp roles (320 threads, one persistent CTA per SM):
//     warps 0-7  softmax: two warps per TMEM lane quadrant, each thread owns one row x half of the tile's keys
//                (row maxima / sums exchanged through shared memory between the two).
//                tcgen05.ld S -> scale + additive mask -> row max -> ex2 -> row sum,
//                P (bf16) -> tensor memory (tcgen05.st; PTMEM, default) or shared memory (K-major, swizzled, in place over
//                the K tile) for the second MMA.  The running max is only raised when a tile exceeds it by more than 2^8
//                (exact: O and the row sum are accumulated against the same reference value), so O in TMEM is almost
//                never rescaled.  Without QTMA they also build the block-diagonal Q2 of the NEXT item.
//     warp 8     TMA producer: K tiles through a 3-stage ring, V tiles through a 3-stage (QTMA) / 2-stage ring, and with
//                QTMA (default) the item's stacked queries [Q_h ; Q_h+1] (two 64 x 64 boxes -> one 128 x 64 slab)
//     warp 9     TMEM allocator + UMMA issuer; issue order S(g+1), PV(g) so that the softmax of tile g overlaps
//                the first MMA of tile g+1 (two S accumulators in TMEM) - across work items too.  With QTMA S(g) is two
//                MMAs of contraction 64, each writing only its head's 64 TMEM lanes; with PTMEM PV(g) takes P from TMEM
// Round-2 history of this kernel (512 users x 1600 keys, no mask, alone): 0.635 ms -> 0.608 (P in TMEM + packed f32x2
// softmax arithmetic) -> 0.581 ms (stacked queries by TMA, third V stage) = 93 % of the measured HBM peak; DESIGN.md 3.2.
// Mask semantics are those of attention.cu: key_mask == 0 adds -1e30 (log2 domain), keys beyond nk are excluded
// (-inf), a row whose keys are all masked comes out uniform over the nk keys.
#include "common.cuh"
#include "umma_pipe.cuh"

#include <cstdlib>

namespace unirec {

constexpr int AT_KT = 128;                       // keys per tile
constexpr int AT_SLAB = 128 * 64 * 2;            // [128 rows][64 bf16] = 16 KB
constexpr int AT_TILE = 2 * AT_SLAB;             // 32 KB: Q2, a K tile, a V tile, P
constexpr int AT_THREADS = 320;                  // 8 softmax warps + TMA producer + UMMA issuer
constexpr int AT_KSTAGES = 3;                    // K ring depth (a K slot is reused as that tile's P)
constexpr int AT_SMEM_BYTES = 2 * AT_TILE /*Q2 x2*/ + AT_KSTAGES * AT_TILE /*K|P x3*/ + 2 * AT_TILE /*V x2*/ +
                              2 * 2 * AT_KT * 2 /*row-max exchange*/ + 1024 /*align*/ + 256 /*barriers*/;
static_assert(AT_SMEM_BYTES <= 232448, "shared memory budget exceeded");
constexpr float AT_MASKED = -1.2676506002282294e30f;   // -2^100: stands in for finfo.min, exact in bf16 (row-max exchange)
constexpr float AT_LAZY = 8.0f;                  // raise the running max only when exceeded by 2^8
constexpr bool AT_PTMEM_DEFAULT = true;
constexpr bool AT_QTMA_DEFAULT = true;           // stacked queries by TMA + three V stages (attention_tc_kernel<.., QTMA>)

struct AttnTcParams {
    const __nv_bfloat16* q; long long ldq; long long q_batch_rows;
    const float* key_mask;                       // [B, nk] or nullptr
    __nv_bfloat16* out; long long ldo;
    int num_heads, nq, nk, kv_batch_rows;
    float scale_log2;
    int num_items;                               // batch * num_heads / 2
};

// K-major SW128 descriptor is umma_smem_desc_sw128 (common.cuh).  MN-major SW128: the tile is
// [K rows of 128 B][64 MN elements], 8-row groups 1024 B apart (SBO), 64-element MN blocks lbo_bytes apart (LBO).
UNIREC_DEVICE uint64_t umma_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(uint32_t m, uint32_t n) {
    return umma_idesc_bf16(m, n) | (1u << 16);   // B operand MN-major
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M = 128 rows = TMEM lanes, K along the columns, two bf16 per 32-bit
// column) is read from tensor memory - P goes from the softmax registers to the second MMA without touching shared memory
UNIREC_DEVICE void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// cta_group::1 MMA that leaves the TMEM lanes 64..127 (skip_upper) or 0..63 untouched (disable-output-lane mask; checked on
// the hardware by tools/probe_mixed_cta_group.cu, used by kv_attention_fused.cu and attention_pp.cu)
UNIREC_DEVICE void umma_bf16_ss_lanes64(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
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

UNIREC_DEVICE void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// PTMEM = true: P(g) is written to tensor memory (columns 384..511: two 64-column buffers of packed bf16 pairs) and PV(g)
// takes it from there as its A operand; the K slot is released as soon as S(g) has been accumulated.  PTMEM = false: P(g) is
// written in place over K(g) in shared memory (the round-1 kernel).
// QTMA = true: the queries of a work item are the two heads' 64 x 64 blocks STACKED into one 128 x 64 slab (16 KB, loaded by
// the TMA producer as two boxes) and S(g) is two MMAs of contraction 64 that each write only their head's 64 TMEM lanes
// (disable-output-lane mask) - no block-diagonal zero padding, no Q2 code in the softmax warps, and the 32 KB saved buy a
// third V stage.  QTMA = false: the softmax warps build the block-diagonal 128 x 128 Q2 (the round-1 kernel).
template <bool PTMEM, bool QTMA>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_q, const AttnTcParams p) {
    constexpr int QBUF = QTMA ? AT_SLAB : AT_TILE;       // bytes per query buffer
    constexpr int VST = QTMA ? 3 : 2;                    // V ring depth
    extern __shared__ uint8_t smem_raw[];
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ2 = smem;                          // [2][2 slabs]
    uint8_t* sK = smem + 2 * QBUF;                // [3][2 slabs]  K(g), later P(g) unless PTMEM
    uint8_t* sV = sK + AT_KSTAGES * AT_TILE;      // [VST][2 slabs]   (2 QBUF + 3 + VST tiles = 7 tiles either way)
    // smem + 7 tiles: [2][2][128] bf16 row-max exchange between the two softmax warps of a row quadrant (1 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * AT_TILE + 2 * 2 * AT_KT * 2);
    uint64_t* k_full = bars;            // [3]
    uint64_t* k_empty = bars + 3;       // [3]  issuer (commit after PV: the slot held K, then P)
    uint64_t* v_full = bars + 6;        // [3]
    uint64_t* v_empty = bars + 9;       // [3]
    uint64_t* q_ready = bars + 12;      // [2]  softmax warps (or the TMA producer, QTMA) -> issuer
    uint64_t* q_free = bars + 14;       // [2]  issuer (commit) -> softmax warps (or the TMA producer)
    uint64_t* s_full = bars + 16;       // [2]  issuer (commit) -> softmax warps
    uint64_t* s_free = bars + 18;       // [2]  softmax warps -> issuer
    uint64_t* p_ready = bars + 20;      // softmax warps -> issuer
    uint64_t* pv_done = bars + 21;      // issuer (commit) -> softmax warps
    uint64_t* o_free = bars + 22;       // softmax warps -> issuer (O of the finished item has been read)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 23);

    const int T = (p.nk + AT_KT - 1) / AT_KT;
    const int my_items = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int G = my_items * T;
    const int pairs = p.num_heads >> 1;

    if (warp_idx == 8 && lane == 0) {
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        if constexpr (QTMA) tma_prefetch_desc(&tmap_q);
        for (int i = 0; i < AT_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < 3; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_ready[i], QTMA ? 1 : 8); mbar_init(&q_free[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 8);
        }
        mbar_init(p_ready, 8);
        mbar_init(pv_done, 1);
        mbar_init(o_free, 8);
        fence_mbar_init();
    }
    if (warp_idx == 9) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_o = tmem_base + 256;

    if (warp_idx == 8) {
        // ===================== TMA producer =====================
        for (int g = 0; g < G; ++g) {
            const int it = g / T, t = g - it * T;
            const int w = blockIdx.x + it * gridDim.x;
            const int b = w / pairs, hp = w - b * pairs;
            const int slot = g % VST;
            const uint32_t ph = (g / VST) & 1;
            const int kslot = g % AT_KSTAGES;
            const uint32_t kph = (g / AT_KSTAGES) & 1;
            const int row = b * p.kv_batch_rows + t * AT_KT;
            const int col = hp * 128;
            if constexpr (QTMA) {
                if (t == 0) {
                    // stacked queries of the item: rows 0-63 = head 2 hp, rows 64-127 = head 2 hp + 1 (rows beyond nq belong to
                    // the next batch element or are zero-filled: they land in accumulator rows that are never stored)
                    mbar_wait(&q_free[it & 1], ((it >> 1) & 1) ^ 1);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&q_ready[it & 1], AT_SLAB);
                        tma_load_2d(&tmap_q, &q_ready[it & 1], sQ2 + (it & 1) * AT_SLAB, col, b * static_cast<int>(p.q_batch_rows));
                        tma_load_2d(&tmap_q, &q_ready[it & 1], sQ2 + (it & 1) * AT_SLAB + 64 * 128, col + 64, b * static_cast<int>(p.q_batch_rows));
                    }
                    __syncwarp();
                }
            }
            mbar_wait(&k_empty[kslot], kph ^ 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&k_full[kslot], AT_TILE);
                tma_load_2d(&tmap_k, &k_full[kslot], sK + kslot * AT_TILE, col, row);
                tma_load_2d(&tmap_k, &k_full[kslot], sK + kslot * AT_TILE + AT_SLAB, col + 64, row);
            }
            __syncwarp();
            mbar_wait(&v_empty[slot], ph ^ 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&v_full[slot], AT_TILE);
                tma_load_2d(&tmap_v, &v_full[slot], sV + slot * AT_TILE, col, row);
                tma_load_2d(&tmap_v, &v_full[slot], sV + slot * AT_TILE + AT_SLAB, col + 64, row);
            }
            __syncwarp();
        }
    } else if (warp_idx == 9) {
        // ===================== UMMA issuer =====================
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
        constexpr uint32_t idesc_pv = umma_idesc_bf16_bmn(128, 128);
        auto issue_s = [&](int g) {
            const int it = g / T, t = g - it * T;
            const int slot = g & 1;
            const uint32_t ph = (g >> 1) & 1;
            const int kslot = g % AT_KSTAGES;
            if (t == 0) mbar_wait(&q_ready[it & 1], (it >> 1) & 1);
            mbar_wait(&k_full[kslot], (g / AT_KSTAGES) & 1);
            mbar_wait(&s_free[slot], ph ^ 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(sQ2 + (it & 1) * QBUF);
                const uint32_t b_addr = smem_u32(sK + kslot * AT_TILE);
                if constexpr (QTMA) {
#pragma unroll
                    for (int hd = 0; hd < 2; ++hd)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_bf16_ss_lanes64(tmem_base + slot * 128, umma_smem_desc_sw128(a_addr + ks * 32),
                                                 umma_smem_desc_sw128(b_addr + hd * AT_SLAB + ks * 32), idesc_s,
                                                 ks != 0 ? 1u : 0u, hd == 0);
                } else {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t off = (ks >> 2) * AT_SLAB + (ks & 3) * 32;
                        umma_bf16_ss(tmem_base + slot * 128, umma_smem_desc_sw128(a_addr + off),
                                     umma_smem_desc_sw128(b_addr + off), idesc_s, ks != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&s_full[slot]);
                if constexpr (PTMEM) umma_commit(&k_empty[kslot]);      // K(g) is dead once S(g) has been accumulated
                if (t == T - 1) umma_commit(&q_free[it & 1]);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int g) {
            const int it = g / T, t = g - it * T;
            const int slot = g & 1;
            const int vslot = g % VST;
            mbar_wait(&v_full[vslot], (g / VST) & 1);
            mbar_wait(p_ready, g & 1);
            if (t == 0 && it > 0) mbar_wait(o_free, (it - 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(sK + (g % AT_KSTAGES) * AT_TILE);   // P(g), written over K(g)
                const uint32_t b_addr = smem_u32(sV + vslot * AT_TILE);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t a_off = (ks >> 2) * AT_SLAB + (ks & 3) * 32;     // 16 keys = 32 B along K
                    const uint32_t b_off = ks * 16 * 128;                            // 16 key rows of 128 B
                    if constexpr (PTMEM)
                        umma_bf16_ts(tmem_o, tmem_base + 384 + slot * 64 + ks * 8,   // 16 keys = 8 packed columns
                                     umma_smem_desc_mn_sw128(b_addr + b_off, AT_SLAB), idesc_pv, (t | ks) != 0 ? 1u : 0u);
                    else
                        umma_bf16_ss(tmem_o, umma_smem_desc_sw128(a_addr + a_off),
                                     umma_smem_desc_mn_sw128(b_addr + b_off, AT_SLAB), idesc_pv, (t | ks) != 0 ? 1u : 0u);
                }
                umma_commit(&v_empty[vslot]);
                if constexpr (!PTMEM) umma_commit(&k_empty[g % AT_KSTAGES]);
                umma_commit(pv_done);
            }
            __syncwarp();
        };
        if (G > 0) issue_s(0);
        for (int g = 0; g < G; ++g) {
            if (g + 1 < G) issue_s(g + 1);
            issue_pv(g);
        }
    } else {
        // ===================== softmax warps 0-7: thread = (accumulator row r, column half hf) =====================
        // Two warps share each TMEM lane quadrant (a warp may only touch lanes 32*(warp%4)..+31): warp w and w+4 own
        // the same 32 rows and split the tile's 128 keys in halves, so a tile's 16384 exponentials are spread over
        // eight warps (two per scheduler) instead of four - the softmax, not HBM, was the tile period before.
        const int quad = warp_idx & 3, hf = warp_idx >> 2;
        const int r = quad * 32 + lane;            // 0..127
        const int hsel = r >> 6;                   // 0: head 2*hp, 1: head 2*hp + 1
        const int qrow = r & 63;
        const uint32_t lane_field = static_cast<uint32_t>(quad * 32) << 16;
        const int pair_bar = 1 + quad;             // named barrier of the two warps that share these rows (64 threads)
        // exchange area of this row quadrant (256 B, touched by its two warps only): row maxima [2 slots][2 halves][32]
        // bf16 during the tiles, row sums [2 halves][32] fp32 at the end of a work item
        __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(smem + 7 * AT_TILE + quad * 256);
        float* sSum = reinterpret_cast<float*>(smem + 7 * AT_TILE + quad * 256);

        // Q2 rows: slab `hsel` holds this row's 64 query values (written by the hf = 0 thread of the row), the other
        // slab is zero (written by the hf = 1 thread)
        auto load_q = [&](int it_, uint4 (&qv)[8]) {
            const int w = blockIdx.x + it_ * gridDim.x;
            const int b_ = w / pairs, hp_ = w - b_ * pairs;
            const __nv_bfloat16* src = p.q + (static_cast<long long>(b_) * p.q_batch_rows + qrow) * p.ldq +
                                       (2 * hp_ + hsel) * 64;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                qv[c] = (hf == 0 && qrow < p.nq) ? __ldg(reinterpret_cast<const uint4*>(src) + c) : make_uint4(0, 0, 0, 0);
        };
        auto store_q = [&](int it_, const uint4 (&qv)[8]) {
            uint8_t* base = sQ2 + (it_ & 1) * AT_TILE + (hf == 0 ? hsel : (hsel ^ 1)) * AT_SLAB;
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(base + swz128(r, c)) = qv[c];
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&q_ready[it_ & 1]);
        };

        uint4 qv[8];
        if constexpr (!QTMA) {
            if (my_items > 0) {
                load_q(0, qv);
                store_q(0, qv);
            }
        }
        float m_used = -INFINITY, l_part = 0.f;
        // key_mask values of keys (64*hf + lane) and (+32) of tile (it_, t_), fetched one tile AHEAD (ncu: issued at the
        // tile's start, the load's latency sat on the critical path of every tile); 1.0 = attend.  (Keeping the mask row
        // pointers per item instead of this per-tile division measured 4-6 % SLOWER, profiles/r02_s_*: left as it is.)
        auto mask_fetch = [&](int it_, int t_, float& a, float& bq) {
            a = 1.0f; bq = 1.0f;
            if (p.key_mask == nullptr) return;
            const int b_ = (blockIdx.x + it_ * gridDim.x) / pairs;
            const int key = t_ * AT_KT + 64 * hf + lane;
            const float* mrow = p.key_mask + static_cast<long long>(b_) * p.nk;
            if (key < p.nk) a = __ldg(mrow + key);
            if (key + 32 < p.nk) bq = __ldg(mrow + key + 32);
        };
        float mrawA = 1.0f, mrawB = 1.0f;
        if (G > 0) mask_fetch(0, 0, mrawA, mrawB);
        int it = 0, t = 0, b = 0, hp = 0;
        for (int g = 0; g < G; ++g, ++t) {
            if (t == T) { t = 0; ++it; }
            if (t == 0) {
                const int w = blockIdx.x + it * gridDim.x;
                b = w / pairs;
                hp = w - b * pairs;
                m_used = -INFINITY;
                l_part = 0.f;
                if constexpr (!QTMA) {
                    if (it + 1 < my_items) load_q(it + 1, qv);    // in flight during this tile's softmax
                }
            }
            const int slot = g & 1;
            const uint32_t ph = (g >> 1) & 1;
            // ---- this half-tile's key mask as two 32-bit words (bit j = key 64*hf + j attends); no shared memory
            const int n_exist = min(max(p.nk - t * AT_KT - 64 * hf, 0), 64);      // keys of this half inside [0, nk)
            const uint32_t att_lo = __ballot_sync(0xffffffffu, lane < n_exist && mrawA != 0.f);
            const uint32_t att_hi = __ballot_sync(0xffffffffu, lane + 32 < n_exist && mrawB != 0.f);
            const bool plain = (att_lo & att_hi) == 0xffffffffu;
            if (g + 1 < G) {
                if (t + 1 == T) mask_fetch(it + 1, 0, mrawA, mrawB);
                else mask_fetch(it, t + 1, mrawA, mrawB);
            }

            mbar_wait(&s_full[slot], ph);
            tc_fence_after();
            const uint32_t tmem_s = tmem_base + slot * 128 + 64 * hf + lane_field;

            // ---- this thread's 64 scores into registers, then hand the accumulator back to the issuer
            uint32_t sv[2][32];
            tmem_ld_32x32(tmem_s, sv[0]);
            tmem_ld_32x32(tmem_s + 32, sv[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[slot]);

            // ---- x = scale * s (+ mask); row maximum.  Plain tiles keep raw s and fold the scale into the ex2 FFMA.
            float mx = -INFINITY;
            float xs = p.scale_log2;           // multiplier applied inside the exponent
            if (plain) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sv[c][j]));
                mx *= p.scale_log2;            // scale > 0: max commutes with the scaling
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t att = c == 0 ? att_lo : att_hi;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = ((att >> j) & 1u) ? __uint_as_float(sv[c][j]) * p.scale_log2
                                                          : ((c * 32 + j < n_exist) ? AT_MASKED : -INFINITY);
                        sv[c][j] = __float_as_uint(x);
                        mx = fmaxf(mx, x);
                    }
                }
                xs = 1.0f;
            }
            // ---- row maximum over both halves: exchange bf16-rounded maxima with the partner warp (both sides use
            //      the rounded values, so they agree exactly; any common reference is valid for the lazy scheme)
            {
                const __nv_bfloat16 mine = __float2bfloat16_rn(mx);
                sX[(slot * 2 + hf) * 32 + lane] = mine;
                named_bar_sync(pair_bar, 64);
                mx = fmaxf(__bfloat162float(mine), __bfloat162float(sX[(slot * 2 + (hf ^ 1)) * 32 + lane]));
            }
            // ---- lazy running max: rescale only when this tile exceeds the reference by more than 2^8
            const bool raise = mx > m_used + AT_LAZY;      // first tile: m_used = -inf -> true (mx is finite or AT_MASKED)
            float alpha = 1.0f;
            if (raise) {
                alpha = ex2_approx(m_used - mx);           // first tile: 0
                m_used = mx;
                l_part *= alpha;
            }
            // PV(g-1) only has to be complete before O is rescaled (rare: lazy max).  P(g) goes into K(g)'s slot, which
            // the tensor core finished reading when s_full[g] completed, so the exponentials below never wait for it.
            bool pv_seen = (g == 0);
            if (t > 0 && __any_sync(0xffffffffu, raise)) {
                mbar_wait(pv_done, (g - 1) & 1);
                tc_fence_after();
                pv_seen = true;
                const uint32_t tmem_orow = tmem_o + lane_field + hsel * 64 + hf * 32;   // this thread's 32 O columns
                uint32_t v[32];
                tmem_ld_32x32(tmem_orow, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
                tmem_st_32x32(tmem_orow, v);
                tmem_st_wait();
            }
            // ---- p = 2^(x - m_used) -> bf16 -> P tile in shared memory (slab hf = this thread's 64 keys); row sum
            const float neg_m = -m_used;
            const unsigned long long xs2 = pack_f32x2(xs, xs), nm2 = pack_f32x2(neg_m, neg_m);
            uint8_t* prow = sK + (g % AT_KSTAGES) * AT_TILE + hf * AT_SLAB;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                // two elements per FMA-pipe instruction (fma.rn.f32x2 / add.rn.f32x2): the exponent x * scale - m and the row
                // sum cost 32 + 32 instead of 64 + 64 issue slots per 64 keys next to the 64 MUFU.EX2
                unsigned long long psum2 = pack_f32x2(0.f, 0.f);
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float x0, x1;
                    unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sv[c][j]), __uint_as_float(sv[c][j + 1])), xs2, nm2), x0, x1);
                    const float p0 = ex2_approx(x0);
                    const float p1 = ex2_approx(x1);
                    psum2 = add_f32x2(psum2, pack_f32x2(p0, p1));
                    pk[j >> 1] = pack_bf16(p0, p1);
                }
                {
                    float s0, s1;
                    unpack_f32x2(psum2, s0, s1);
                    l_part += s0 + s1;
                }
                if constexpr (PTMEM) {
                    // this thread's keys 64 hf + 32 c .. + 31 of the row: 16 packed columns of the P buffer
                    tmem_st_32x16(tmem_base + 384 + slot * 64 + hf * 32 + c * 16 + lane_field, pk);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(prow + swz128(r, c * 4 + j)) =
                            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
            }
            if constexpr (PTMEM) tmem_st_wait();
            // every phase of pv_done is consumed in order, and phase g-1 before p_ready(g) is signalled: PV(g) cannot be
            // issued earlier, so the barrier is never more than one phase ahead of this thread's parity bookkeeping
            if (!pv_seen) mbar_wait(pv_done, (g - 1) & 1);
            tc_fence_before();               // orders the O rescale / the P stores (tcgen05.st) before the issuer's next MMA
            if constexpr (!PTMEM) fence_proxy_async_smem();        // P writes in shared memory -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);

            if constexpr (!QTMA) {
                if (t == 0 && it + 1 < my_items) {
                    // Q2 of the next item (its buffer was last read by item it-1, long finished)
                    if (it >= 1) mbar_wait(&q_free[(it + 1) & 1], ((it - 1) >> 1) & 1);
                    store_q(it + 1, qv);
                }
            }
            if (t == T - 1) {
                // ---- finalize: row sum over both halves, then O / l -> bf16 -> global (32 of the row's 64 columns)
                named_bar_sync(pair_bar, 64);              // partner has read this tile's maxima: the area is idle
                sSum[hf * 32 + lane] = l_part;
                named_bar_sync(pair_bar, 64);
                const float inv = 1.0f / (l_part + sSum[(hf ^ 1) * 32 + lane]);
                named_bar_sync(pair_bar, 64);              // partner has read my sum before the area holds maxima again
                mbar_wait(pv_done, g & 1);
                tc_fence_after();
                const uint32_t tmem_orow = tmem_o + lane_field + hsel * 64 + hf * 32;
                __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.nq + qrow) * p.ldo + (2 * hp + hsel) * 64 +
                                     hf * 32;
                uint32_t v[32];
                tmem_ld_32x32(tmem_orow, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_free);
                if (qrow < p.nq) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(dst + j * 8) = make_uint4(
                            pack_bf16(__uint_as_float(v[8 * j]) * inv, __uint_as_float(v[8 * j + 1]) * inv),
                            pack_bf16(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv),
                            pack_bf16(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv),
                            pack_bf16(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 9) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

bool attention_tc_supported(long long num_heads, long long nq, long long nk, long long head_dim, long long ldk,
                            long long ldv, long long kv_batch_rows) {
    return head_dim == 64 && nq <= 64 && (num_heads % 2) == 0 && nk > 64 && ldk % 8 == 0 && ldv % 8 == 0 &&
           kv_batch_rows >= nk;
}

int attention_tc(const void* q, long long ldq, long long q_batch_rows, const void* k, long long ldk, const void* v,
                 long long ldv, long long kv_batch_rows, const float* key_mask, void* out, long long ldo, long long batch,
                 long long num_heads, long long nq, long long nk, float scale, cudaStream_t stream) {
    if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
        (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || ldq % 8 != 0 || ldo % 8 != 0 ||
        batch * kv_batch_rows > 2147483647LL) {
        set_last_error("attention (tcgen05): pointers must be 16-byte aligned, row strides multiples of 8");
        return UNIREC_ERR_BAD_ARG;
    }
    AttnTcParams p;
    p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.ldq = ldq; p.q_batch_rows = q_batch_rows;
    p.key_mask = key_mask;
    p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo;
    p.num_heads = static_cast<int>(num_heads); p.nq = static_cast<int>(nq); p.nk = static_cast<int>(nk);
    p.kv_batch_rows = static_cast<int>(kv_batch_rows);
    p.scale_log2 = scale * 1.4426950408889634f;
    p.num_items = static_cast<int>(batch * (num_heads / 2));
    CUtensorMap tk, tv;
    int rc = make_tmap_bf16_2d(&tk, k, batch * kv_batch_rows, num_heads * 64, ldk, AT_KT);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tv, v, batch * kv_batch_rows, num_heads * 64, ldv, AT_KT);
    if (rc != UNIREC_OK) return rc;
    // UNIREC_ATTENTION_PTMEM = 1 / 0: P as a tensor-memory operand of the second MMA / P in shared memory over the K tile;
    // UNIREC_ATTENTION_QTMA = 1 / 0: stacked queries loaded by TMA + lane-masked S MMAs + three V stages / Q2 built by the
    // softmax warps
    static const bool ptmem = [] {
        const char* e = getenv("UNIREC_ATTENTION_PTMEM");
        return e != nullptr ? e[0] != '0' : AT_PTMEM_DEFAULT;
    }();
    static const bool qtma = [] {
        const char* e = getenv("UNIREC_ATTENTION_QTMA");
        return e != nullptr ? e[0] != '0' : AT_QTMA_DEFAULT;
    }();
    CUtensorMap tq;
    const long long q_rows = q_batch_rows > 0 ? (batch - 1) * q_batch_rows + nq : nq;
    rc = make_tmap_bf16_2d(&tq, q, q_rows, num_heads * 64, ldq, 64);
    if (rc != UNIREC_OK) return rc;
    auto kern = ptmem ? (qtma ? attention_tc_kernel<true, true> : attention_tc_kernel<true, false>)
                      : (qtma ? attention_tc_kernel<false, true> : attention_tc_kernel<false, false>);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
        if (e != cudaSuccess) {
            set_last_error("attention (tcgen05): cudaFuncSetAttribute(%d): %s", AT_SMEM_BYTES, cudaGetErrorString(e));
            return UNIREC_ERR_CUDA;
        }
        attr_set = true;
    }
    int grid = num_sms() > 0 ? num_sms() : 148;
    if (grid > p.num_items) grid = p.num_items;
    kern<<<grid, AT_THREADS, AT_SMEM_BYTES, stream>>>(tk, tv, tq, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("attention (tcgen05) launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

}  // namespace unirec
