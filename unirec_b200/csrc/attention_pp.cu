// Long-key cross-attention on tcgen05, TWO softmax groups working on alternate key tiles ("ping-pong") - the successor of
// attention_tc.cu for more than 128 keys (models/qformer.py:185-188, 205, 244-268 with encoder_hidden_states = the user
// sequence: 64 queries x 1600 keys per user and head, head_dim 64).
//
// Why: ncu on attention_tc_kernel (profiles/r02_n_*) shows its eight softmax warps - two per scheduler, in lockstep on the
// same tile - issue an instruction in only ~1/3 of the cycles: a tile's chain (TMEM load -> row max -> exchange -> 2^x ->
// P store -> fence -> arrive) is a latency chain, not a throughput limit (moving half of the exponentials from MUFU to the
// FMA pipe changed nothing, profiles/r02_m_*), and its ~3600 cycles per 128-key tile make the kernel SM-clock-bound: 79-85 %
// of the HBM peak alone at 1.9 GHz, 54-67 % inside the power-capped step at 1.35-1.45 GHz.  More warps on the SAME tile do
// not shorten the chain (16 warps, four per row: measured slower).  Here sixteen softmax warps form two groups that own the
// even and the odd tiles of the stream, so a scheduler always holds warps in two different phases of two different tiles.
//
// One work item = (batch element, pair of adjacent heads), as in attention_tc.cu:
//     S_g[128 x 128 keys] = [Q_h ; Q_h+1] (stacked, 128 x 64) . Kt_g : two MMAs of contraction 64, each writing only its
//                           head's 64 TMEM lanes (disable-output-lane mask) - no block-diagonal zero padding, Q is 16 KB
//     O_grp[128 x 128]   += P_g[128 x 128 keys] . Vt_g   (useful blocks: rows 0-63 x cols 0-63, rows 64-127 x cols 64-127)
// Group grp = g & 1 keeps its own running max / row sum and its own O accumulator in TMEM (S_0 S_1 O_0 O_1 = 512 columns);
// the two partial results of an item are merged when its last PV has completed (both groups, 16 columns per thread).
//
// Warp roles (640 threads, one persistent CTA per SM; setmaxnreg moves registers from the control warps to the softmax warps):
//     warps 0-7 / 8-15  softmax group 0 / 1: two warps per TMEM lane quadrant, thread = (row, half of the tile's keys)
//     warp 16           TMA producer: Q (two 64 x 64 boxes per item, single buffer), K tiles through a 4-stage ring (P_g is
//                       written in place over K_g, so a slot is busy from the load to the end of PV_g - with two tiles in
//                       the softmax at any time a 3-stage ring leaves no slot to prefetch into)
//     warp 18           TMA producer of the V tiles (2 stages)
//     warp 17           TMEM allocator + UMMA issuer: polls the S stream and the PV stream, issues whichever is ready
// Mask semantics are those of attention.cu / attention_tc.cu: key_mask == 0 adds -2^100 (log2 domain), keys beyond nk are
// excluded (-inf), a row whose keys are all masked comes out uniform over the nk keys.
#include "common.cuh"
#include "umma_pipe.cuh"

#include <cstdlib>

namespace unirec {

constexpr int PP_KT = 128;                       // keys per tile
constexpr int PP_SLAB = 128 * 64 * 2;            // [128 rows][64 bf16] = 16 KB
constexpr int PP_TILE = 2 * PP_SLAB;             // 32 KB: a K tile (later P), a V tile
constexpr int PP_KSTAGES = 4;
constexpr int PP_THREADS = 20 * 32;              // 16 softmax warps + TMA producer + UMMA issuer + 2 idle warps (registers
                                                 // are allocated per group of four warps; the fifth group gives its own back)
constexpr int PP_XCHG_BYTES = 2 * 4 * 256;       // row-max exchange: [group][quadrant][2 slots][2 halves][32] bf16
constexpr int PP_MERGE_M = 2 * 2 * 128 * 4;      // running max      [item parity][group][row] fp32
constexpr int PP_MERGE_L = 2 * 2 * 2 * 128 * 4;  // partial row sums [item parity][group][half][row] fp32
constexpr int PP_SMEM_BYTES = PP_SLAB /*Q*/ + PP_KSTAGES * PP_TILE + 2 * PP_TILE /*V*/ + PP_XCHG_BYTES + PP_MERGE_M +
                              PP_MERGE_L + 1024 /*align*/ + 256 /*barriers*/;
static_assert(PP_SMEM_BYTES <= 232448, "shared memory budget exceeded");
constexpr float PP_MASKED = -1.2676506002282294e30f;   // -2^100: stands in for finfo.min, exact in bf16 (row-max exchange)
constexpr float PP_LAZY = 8.0f;                  // raise the running max only when exceeded by 2^8

struct AttnPpParams {
    const float* key_mask;                       // [B, nk] or nullptr
    __nv_bfloat16* out; long long ldo;
    int num_heads, nq, nk, kv_batch_rows, q_batch_rows;
    float scale_log2;
    int num_items;                               // batch * num_heads / 2
};

// MN-major SW128 B operand: [K rows of 128 B][64 MN elements], 8-row groups 1024 B apart (SBO), 64-element MN blocks
// lbo_bytes apart (LBO) - the V tile used in place (attention_tc.cu)
UNIREC_DEVICE uint64_t pp_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// cta_group::1 MMA that leaves the TMEM lanes 64..127 (skip_upper) or 0..63 untouched (hardware-checked by
// tools/probe_mixed_cta_group.cu and the fused K/V attention kernel)
UNIREC_DEVICE void pp_umma_lanes(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
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
UNIREC_DEVICE void pp_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(PP_THREADS, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const AttnPpParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                   // [128 rows: Q_h, Q_h+1][64]
    uint8_t* sK = smem + PP_SLAB;                         // [4][2 slabs]  K(g), later P(g)
    uint8_t* sV = sK + PP_KSTAGES * PP_TILE;              // [2][2 slabs]
    uint8_t* sX = sV + 2 * PP_TILE;
    float* sM = reinterpret_cast<float*>(sX + PP_XCHG_BYTES);
    float* sL = reinterpret_cast<float*>(sX + PP_XCHG_BYTES + PP_MERGE_M);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + PP_XCHG_BYTES + PP_MERGE_M + PP_MERGE_L);
    uint64_t* k_full = bars;            // [4]
    uint64_t* k_empty = bars + 4;       // [4]  issuer (commit after PV: the slot held K, then P)
    uint64_t* v_full = bars + 8;        // [2]
    uint64_t* v_empty = bars + 10;      // [2]
    uint64_t* q_full = bars + 12;       // TMA -> issuer
    uint64_t* q_free = bars + 13;       // issuer (commit after the item's last S) -> producer
    uint64_t* s_full = bars + 14;       // [2 groups]  issuer (commit) -> softmax group
    uint64_t* s_free = bars + 16;       // [2]  softmax group -> issuer
    uint64_t* p_ready = bars + 18;      // [2]  softmax group -> issuer
    uint64_t* pv_done = bars + 20;      // [2]  issuer (commit) -> softmax group
    uint64_t* item_done = bars + 22;    // issuer (commit after the item's last PV) -> all softmax warps
    uint64_t* o_free = bars + 23;       // all softmax warps -> issuer (both O accumulators of the item have been read)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 24);

    const int T = (p.nk + PP_KT - 1) / PP_KT;             // >= 2 (host)
    const int my_items = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int G = my_items * T;
    const int pairs = p.num_heads >> 1;

    if (warp_idx == 16 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        for (int i = 0; i < PP_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 8);
            mbar_init(&p_ready[i], 8); mbar_init(&pv_done[i], 1);
        }
        mbar_init(q_full, 1);
        mbar_init(q_free, 1);
        mbar_init(item_done, 1);
        mbar_init(o_free, 16);
        fence_mbar_init();
    }
    if (warp_idx == 17) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // 640 threads start with 96 registers each; the control warp group shrinks to 64 and the softmax warp groups grow to 104
    // (16 warps x 8 = 4 warps x 32: the CTA's register pool balances exactly - an increase the pool cannot serve blocks forever).
    // The role code sits INSIDE the two branches: ptxas only gives the softmax code its 104 registers that way.
    if (warp_idx >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp_idx == 16) {
            // ===================== TMA producer: Q and the K ring =====================
            int it = 0, t = 0;
            for (int g = 0; g < G; ++g, ++t) {
                if (t == T) { t = 0; ++it; }
                const int w = blockIdx.x + it * gridDim.x;
                const int b = w / pairs, hp = w - b * pairs;
                if (t == 0) {
                    if (it > 0) mbar_wait(q_free, (it - 1) & 1);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(q_full, PP_SLAB);
                        tma_load_2d(&tmap_q, q_full, sQ, hp * 128, b * p.q_batch_rows);
                        tma_load_2d(&tmap_q, q_full, sQ + 64 * 128, hp * 128 + 64, b * p.q_batch_rows);
                    }
                    __syncwarp();
                }
                const int kslot = g & (PP_KSTAGES - 1);
                const int row = b * p.kv_batch_rows + t * PP_KT;
                mbar_wait(&k_empty[kslot], ((g >> 2) & 1) ^ 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&k_full[kslot], PP_TILE);
                    tma_load_2d(&tmap_k, &k_full[kslot], sK + kslot * PP_TILE, hp * 128, row);
                    tma_load_2d(&tmap_k, &k_full[kslot], sK + kslot * PP_TILE + PP_SLAB, hp * 128 + 64, row);
                }
                __syncwarp();
            }
        } else if (warp_idx == 18) {
            // ===================== TMA producer: the V ring (its own warp: a V slot only frees when PV(g-2) has completed,
            // and a producer that waits for it in line would hold back the K tiles the issuer needs two tiles ahead) ==========
            int it = 0, t = 0;
            for (int g = 0; g < G; ++g, ++t) {
                if (t == T) { t = 0; ++it; }
                const int w = blockIdx.x + it * gridDim.x;
                const int b = w / pairs, hp = w - b * pairs;
                const int vslot = g & 1;
                const int row = b * p.kv_batch_rows + t * PP_KT;
                mbar_wait(&v_empty[vslot], ((g >> 1) & 1) ^ 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&v_full[vslot], PP_TILE);
                    tma_load_2d(&tmap_v, &v_full[vslot], sV + vslot * PP_TILE, hp * 128, row);
                    tma_load_2d(&tmap_v, &v_full[vslot], sV + vslot * PP_TILE + PP_SLAB, hp * 128 + 64, row);
                }
                __syncwarp();
            }
        } else if (warp_idx == 17) {
            // ===================== UMMA issuer =====================
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128) | (1u << 16);      // B operand MN-major
            // The S stream (S(g) needs Q, K(g) and the group's free accumulator) and the PV stream (PV(g) needs V(g) and
            // P(g)) are each issued in order, but whichever is ready goes first: a PV must not queue behind the K tile of
            // an S two tiles ahead (measured: with the fixed order S(g+2), PV(g) the kernel ran 40 % slower than
            // attention_tc_kernel).
            auto s_ready = [&](int g) {
                const int it = g / T, t = g - it * T;
                const uint32_t n = static_cast<uint32_t>(g >> 1);
                if (t == 0 && !mbar_test_wait(q_full, it & 1)) return false;
                return mbar_test_wait(&k_full[g & (PP_KSTAGES - 1)], (g >> 2) & 1) && mbar_test_wait(&s_free[g & 1], (n & 1) ^ 1);
            };
            auto pv_ready = [&](int g) {
                const int it = g / T, t = g - it * T;
                const uint32_t n = static_cast<uint32_t>(g >> 1);
                if (t < 2 && it > 0 && !mbar_test_wait(o_free, (it - 1) & 1)) return false;
                return mbar_test_wait(&p_ready[g & 1], n & 1) && mbar_test_wait(&v_full[g & 1], n & 1);
            };
            auto issue_s = [&](int g) {
                const int it = g / T, t = g - it * T;
                const int grp = g & 1;
                const int kslot = g & (PP_KSTAGES - 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(sQ);
                    const uint32_t b_addr = smem_u32(sK + kslot * PP_TILE);
    #pragma unroll
                    for (int hd = 0; hd < 2; ++hd)
    #pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            pp_umma_lanes(tmem_base + grp * 128, umma_smem_desc_sw128(a_addr + ks * 32),
                                          umma_smem_desc_sw128(b_addr + hd * PP_SLAB + ks * 32), idesc_s, ks != 0 ? 1u : 0u,
                                          hd == 0);
                    umma_commit(&s_full[grp]);
                    if (t == T - 1) umma_commit(q_free);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int g) {
                const int it = g / T, t = g - it * T;
                const int grp = g & 1;
                const int kslot = g & (PP_KSTAGES - 1);
                const bool first = t < 2;                       // this group's first tile of the item: O starts from zero
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(sK + kslot * PP_TILE);     // P(g), written over K(g)
                    const uint32_t b_addr = smem_u32(sV + grp * PP_TILE);
                    const uint32_t tmem_o = tmem_base + 256 + grp * 128;
    #pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t a_off = (ks >> 2) * PP_SLAB + (ks & 3) * 32;     // 16 keys = 32 B along K
                        const uint32_t b_off = ks * 16 * 128;                            // 16 key rows of 128 B
                        umma_bf16_ss(tmem_o, umma_smem_desc_sw128(a_addr + a_off), pp_desc_mn_sw128(b_addr + b_off, PP_SLAB),
                                     idesc_pv, (!first || ks != 0) ? 1u : 0u);
                    }
                    umma_commit(&v_empty[grp]);
                    umma_commit(&k_empty[kslot]);
                    umma_commit(&pv_done[grp]);
                    if (t == T - 1) umma_commit(item_done);
                }
                __syncwarp();
            };
            int gs = 0, gp = 0;
            while (gp < G) {
                // lane 0's view of the barriers decides for the warp
                if (gs < G && __shfl_sync(0xffffffffu, s_ready(gs) ? 1 : 0, 0) != 0) {
                    issue_s(gs);
                    ++gs;
                }
                if (gp < gs && __shfl_sync(0xffffffffu, pv_ready(gp) ? 1 : 0, 0) != 0) {
                    issue_pv(gp);
                    ++gp;
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===================== softmax groups: thread = (accumulator row r, key half hf) of the group's tiles =============
        const int grp = warp_idx >> 3;
        const int quad = warp_idx & 3, hf = (warp_idx >> 2) & 1;
        const int r = quad * 32 + lane;            // 0..127
        const int hsel = r >> 6;                   // 0: head 2*hp, 1: head 2*hp + 1
        const int qrow = r & 63;
        const uint32_t lane_field = static_cast<uint32_t>(quad * 32) << 16;
        const int pair_bar = 1 + grp * 4 + quad;   // named barrier of the two warps that share these rows (64 threads)
        __nv_bfloat16* sXq = reinterpret_cast<__nv_bfloat16*>(sX + (grp * 4 + quad) * 256);
        const uint32_t tmem_s = tmem_base + grp * 128 + 64 * hf + lane_field;
        const uint32_t tmem_orow = tmem_base + 256 + grp * 128 + lane_field + hsel * 64 + hf * 32;   // 32 of my O columns

        // key_mask values of keys (64*hf + lane) and (+32) of tile (it_, t_), fetched one of the group's tiles ahead
        auto mask_fetch = [&](int it_, int t_, float& a, float& bq) {
            a = 1.0f; bq = 1.0f;
            if (p.key_mask == nullptr) return;
            const int b_ = (blockIdx.x + it_ * gridDim.x) / pairs;
            const int key = t_ * PP_KT + 64 * hf + lane;
            const float* mrow = p.key_mask + static_cast<long long>(b_) * p.nk;
            if (key < p.nk) a = __ldg(mrow + key);
            if (key + 32 < p.nk) bq = __ldg(mrow + key + 32);
        };
        float m_used = -INFINITY, l_part = 0.f;
        float mrawA = 1.0f, mrawB = 1.0f;
        int it = 0, t = grp, b = 0, hp = 0;        // T >= 2: tile `grp` of item 0 exists
        if (grp < G) mask_fetch(0, grp, mrawA, mrawB);
        uint32_t n = 0;
        for (int g = grp; g < G; g += 2, ++n, t += 2) {
            if (t >= T) { t -= T; ++it; }
            const bool first = t < 2;              // this group's first tile of the item
            const bool last = t + 2 >= T;          // ... and its last
            if (first) {
                const int w = blockIdx.x + it * gridDim.x;
                b = w / pairs;
                hp = w - b * pairs;
                m_used = -INFINITY;
                l_part = 0.f;
            }
            const int n_exist = min(max(p.nk - t * PP_KT - 64 * hf, 0), 64);      // keys of this half inside [0, nk)
            const uint32_t att_lo = __ballot_sync(0xffffffffu, lane < n_exist && mrawA != 0.f);
            const uint32_t att_hi = __ballot_sync(0xffffffffu, lane + 32 < n_exist && mrawB != 0.f);
            const bool plain = (att_lo & att_hi) == 0xffffffffu;
            if (g + 2 < G) {
                if (t + 2 >= T) mask_fetch(it + 1, t + 2 - T, mrawA, mrawB);
                else mask_fetch(it, t + 2, mrawA, mrawB);
            }

            mbar_wait(&s_full[grp], n & 1);
            tc_fence_after();
            uint32_t sv[2][32];
            tmem_ld_32x32(tmem_s, sv[0]);
            tmem_ld_32x32(tmem_s + 32, sv[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[grp]);

            // ---- x = scale * s (+ mask); row maximum.  Plain tiles keep raw s and fold the scale into the ex2 FFMA.
            float mx = -INFINITY;
            float xs = p.scale_log2;
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
                                                          : ((c * 32 + j < n_exist) ? PP_MASKED : -INFINITY);
                        sv[c][j] = __float_as_uint(x);
                        mx = fmaxf(mx, x);
                    }
                }
                xs = 1.0f;
            }
            // ---- row maximum over both halves (bf16-rounded on both sides, so the two threads of a row agree exactly)
            {
                const __nv_bfloat16 mine = __float2bfloat16_rn(mx);
                sXq[((n & 1) * 2 + hf) * 32 + lane] = mine;
                pp_bar_sync(pair_bar, 64);
                mx = fmaxf(__bfloat162float(mine), __bfloat162float(sXq[((n & 1) * 2 + (hf ^ 1)) * 32 + lane]));
            }
            // ---- lazy running max: rescale only when this tile exceeds the reference by more than 2^8
            const bool raise = mx > m_used + PP_LAZY;      // first tile: m_used = -inf -> true (mx is finite or PP_MASKED)
            float alpha = 1.0f;
            if (raise) {
                alpha = ex2_approx(m_used - mx);           // first tile: 0
                m_used = mx;
                l_part *= alpha;
            }
            bool pv_seen = (n == 0);
            if (!first && __any_sync(0xffffffffu, raise)) {
                mbar_wait(&pv_done[grp], (n - 1) & 1);     // this group's previous PV wrote the accumulator
                tc_fence_after();
                pv_seen = true;
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
            const uint32_t prow = smem_u32(sK + (g & (PP_KSTAGES - 1)) * PP_TILE + hf * PP_SLAB);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float psum = 0.f;
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(sv[c][j]), xs, neg_m));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(sv[c][j + 1]), xs, neg_m));
                    psum += p0 + p1;
                    pk[j >> 1] = pack_bf16(p0, p1);
                }
                l_part += psum;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    st_shared_v4(prow + swz128(r, c * 4 + j), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
            }
            const int par = it & 1;
            if (last) {
                // this group's share of the item: running max and partial row sum, read by both groups in the merge below
                // (published before the p_ready arrival: arrive -> issuer -> commit -> item_done orders it for the readers)
                if (hf == 0) sM[(par * 2 + grp) * 128 + r] = m_used;
                sL[((par * 2 + grp) * 2 + hf) * 128 + r] = l_part;
            }
            // every phase of pv_done[grp] is consumed in order, and phase n-1 before p_ready(n) is signalled
            if (!pv_seen) mbar_wait(&pv_done[grp], (n - 1) & 1);
            tc_fence_before();               // orders the O rescale (tcgen05.st) before the issuer's next MMA
            fence_proxy_async_smem();        // P writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[grp]);

            if (last) {
                // ---- merge the two groups' partial results of the item: thread = (row, 16 of its head's 64 columns)
                mbar_wait(item_done, par);
                tc_fence_after();
                pp_bar_sync(9, 512);         // both groups have published (m, l): a direct ordering next to the transitive one
                const float m0 = sM[(par * 2 + 0) * 128 + r], m1 = sM[(par * 2 + 1) * 128 + r];
                const float l0 = sL[((par * 2 + 0) * 2 + 0) * 128 + r] + sL[((par * 2 + 0) * 2 + 1) * 128 + r];
                const float l1 = sL[((par * 2 + 1) * 2 + 0) * 128 + r] + sL[((par * 2 + 1) * 2 + 1) * 128 + r];
                const float mm = fmaxf(m0, m1);
                const float w0 = ex2_approx(m0 - mm), w1 = ex2_approx(m1 - mm);
                const float inv = 1.0f / (l0 * w0 + l1 * w1);
                const float f0 = w0 * inv, f1 = w1 * inv;
                const int col0 = (grp * 2 + hf) * 16;
                uint32_t oa[16], ob[16];
                tmem_ld_32x16(tmem_base + 256 + lane_field + hsel * 64 + col0, oa);
                tmem_ld_32x16(tmem_base + 384 + lane_field + hsel * 64 + col0, ob);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_free);
                if (qrow < p.nq) {
                    __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.nq + qrow) * p.ldo + (2 * hp + hsel) * 64 + col0;
                    float o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = fmaf(__uint_as_float(oa[j]), f0, __uint_as_float(ob[j]) * f1);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        *reinterpret_cast<uint4*>(dst + j * 8) =
                            make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                                       pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 17) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

bool attention_pp_supported(long long num_heads, long long nq, long long nk, long long head_dim, long long ldk,
                            long long ldv, long long kv_batch_rows) {
    // two key tiles per item at least: each softmax group must own a tile of every item
    return head_dim == 64 && nq <= 64 && (num_heads % 2) == 0 && nk > PP_KT && ldk % 8 == 0 && ldv % 8 == 0 &&
           kv_batch_rows >= nk;
}

int attention_pp(const void* q, long long ldq, long long q_batch_rows, const void* k, long long ldk, const void* v,
                 long long ldv, long long kv_batch_rows, const float* key_mask, void* out, long long ldo, long long batch,
                 long long num_heads, long long nq, long long nk, float scale, cudaStream_t stream) {
    if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
        (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || ldq % 8 != 0 || ldo % 8 != 0 ||
        batch * kv_batch_rows > 2147483647LL || batch * q_batch_rows > 2147483647LL) {
        set_last_error("attention (tcgen05, two groups): pointers must be 16-byte aligned, row strides multiples of 8");
        return UNIREC_ERR_BAD_ARG;
    }
    AttnPpParams p;
    p.key_mask = key_mask;
    p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo;
    p.num_heads = static_cast<int>(num_heads); p.nq = static_cast<int>(nq); p.nk = static_cast<int>(nk);
    p.kv_batch_rows = static_cast<int>(kv_batch_rows);
    p.q_batch_rows = static_cast<int>(q_batch_rows);
    p.scale_log2 = scale * 1.4426950408889634f;
    p.num_items = static_cast<int>(batch * (num_heads / 2));
    CUtensorMap tq, tk, tv;
    // queries: rows beyond the tensor are zero-filled by TMA; with nq < 64 a box also covers rows of the next batch element -
    // they land in accumulator rows that are never stored
    const long long q_rows = q_batch_rows > 0 ? (batch - 1) * q_batch_rows + nq : nq;
    int rc = make_tmap_bf16_2d(&tq, q, q_rows, num_heads * 64, ldq, 64);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tk, k, batch * kv_batch_rows, num_heads * 64, ldk, PP_KT);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tv, v, batch * kv_batch_rows, num_heads * 64, ldv, PP_KT);
    if (rc != UNIREC_OK) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM_BYTES);
        if (e != cudaSuccess) {
            set_last_error("attention (tcgen05, two groups): cudaFuncSetAttribute(%d): %s", PP_SMEM_BYTES, cudaGetErrorString(e));
            return UNIREC_ERR_CUDA;
        }
        attr_set = true;
    }
    int grid = num_sms() > 0 ? num_sms() : 148;
    if (grid > p.num_items) grid = p.num_items;
    attention_pp_kernel<<<grid, PP_THREADS, PP_SMEM_BYTES, stream>>>(tq, tk, tv, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("attention (tcgen05, two groups) launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

}  // namespace unirec
