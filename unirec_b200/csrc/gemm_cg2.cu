// Projection GEMM, CTA-pair version (sm_100a):
//     out[M,N] (bf16) = epilogue( A[M,K] (bf16, row-major) x W[N,K]^T (bf16, row-major = nn.Linear.weight) )
//
// A thread-block cluster of two CTAs (one SM pair) owns a 256 x 256 output tile.  Each CTA stages ITS 128
// rows of A and ITS 128 rows of W per 64-wide K block (TMA, 128-byte swizzle, 5-stage ring); the leader CTA
// issues tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16), which reads A/W from both CTAs' shared memory and
// writes each CTA's 128 accumulator rows into that CTA's TMEM - so every operand byte is fetched from L2 once
// per pair and read from shared memory once per pair (2/3 of the L2->SM traffic and half of the W smem reads
// of the single-CTA 128 x 256 kernel in gemm_tcgen05.cu).
//
// The epilogue never touches global memory with per-thread loads/stores: the residual tile is brought into
// shared memory by TMA while the tile's MMAs run, the epilogue warps do tcgen05.ld -> bias / erf-GELU /
// +residual -> bf16 IN PLACE in that shared tile, and a dedicated warp writes it back with TMA bulk stores
// (full 128-byte lines, asynchronous).  The C tile is handled as four 64-column slabs so that loads, epilogue
// math and stores of neighbouring slabs/tiles overlap.
//
// Warp roles per CTA (384 threads):
//     warp 0      TMA producer (A and W k-blocks; signals the LEADER's full barrier)
//     warp 1      UMMA issuer (leader CTA only; tcgen05.commit multicast releases both CTAs' smem slots)
//     warp 2      TMEM allocator (cta_group::2, both CTAs)
//     warp 3      C mover: TMA store of finished slabs, TMA load of the next tile's residual slabs
//     warps 4-11  epilogue (TMEM lane quadrant = warp % 4, column half = (warp - 4) / 4)
//
// Replaces the same reference lines as gemm_tcgen05.cu (every nn.Linear of the path: models/qformer.py:185-198,
// :286, :359-360, :372; heads models/qformer_utils.py:50,53, training/user_qformer_training.py:38-43).
#include "common.cuh"
#include "cg2_ptx.cuh"
#include "umma_pipe.cuh"

namespace unirec {

enum : int { G2_EPI_BIAS = 0, G2_EPI_BIAS_GELU = 1, G2_EPI_BIAS_RESIDUAL = 2 };

constexpr int G2_TILE_M = 256;              // per cluster; 128 per CTA
constexpr int G2_TILE_N = 256;
constexpr int G2_BLOCK_K = 64;
constexpr int G2_STAGES = 5;
constexpr int G2_A_BYTES = 128 * G2_BLOCK_K * 2;          // this CTA's 128 rows of A
constexpr int G2_B_BYTES = 128 * G2_BLOCK_K * 2;          // this CTA's 128 rows of W
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;   // 32 KB
constexpr int G2_SLABS = 4;                               // 64-column slabs of this CTA's 128 x 256 C tile
constexpr int G2_SLAB_BYTES = 128 * 64 * 2;               // 16 KB
constexpr int G2_THREADS = 384;
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_BARRIER_BYTES = 512;
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + G2_SLABS * G2_SLAB_BYTES + 1024 + G2_BARRIER_BYTES;
constexpr int G2_TMEM_COLS = 512;                         // two 256-column accumulator stages
static_assert(G2_SMEM_BYTES <= 232448, "shared memory budget exceeded");

struct Gemm2Params {
    int M, N, K;
    const float* bias;      // [N] fp32 or nullptr
    int num_m_blocks, num_n_blocks;
    // GATHER kernels only: the A operand is never materialised - every 32-row group g of the virtual
    // [M, K] matrix is the 32 query tokens of item gather_ids[g] in the resident item-token table (tmap_a, box 32 rows), or,
    // for history slots beyond the user's length, 32 rows of a padding table (tmap_pad) at the slot's position
    const long long* gather_ids;    // [M / 32]
    const int* gather_len;          // [users] valid slots per user
    int slots_per_user;
    int table_rows;                 // rows of the token table viewed 2-D: an out-of-bounds coordinate (TMA zero fill)
    int res_period;                 // > 0: the residual tile of rows m.. is read at rows (m % res_period).. of its table
    int reverse;                    // 1: walk the output tiles from the last to the first (experiment, see common.cuh)
    // ---- LayerNorm folded into the GEMMs around it (unirec_linear_ln_bf16; models/qformer.py:288, :374: h = LN(pre)):
    // the LayerNorm output is never materialised.  Its PRODUCER (the dense + residual GEMM that writes `pre`) also adds the
    // row sums / sums of squares of the bf16 values it stores into stats_out; its CONSUMERS read `pre` and these statistics:
    //   * as the A operand (W already scaled by gamma): y = rstd (pre W'^T) - mu rstd c + b', c[n] = sum_k W'[n,k]
    //   * as the residual: h = (pre - mu) rstd gamma + beta, evaluated on the residual tile in the epilogue
    const float* ln_in_stats;       // [M, ln_parts, 2] (sum, sum of squares) partials of the rows of A, or nullptr
    const float* ln_in_c;           // [N] column sums of the gamma-scaled weight
    const float* ln_res_stats;      // [M, ln_parts, 2] of the rows of the residual tensor, or nullptr (residual is used as is)
    const float* ln_res_gamma;      // [N]
    const float* ln_res_beta;       // [N]
    float* stats_out;               // [M, 2 N / 256, 2]: one (sum, sum of squares) partial per 128-column half tile, or nullptr
    int ln_parts;                   // partials per row in ln_in_stats / ln_res_stats (summed in index order: deterministic)
    float ln_eps, ln_inv_h;         // LayerNorm epsilon, 1 / hidden size
};

template <int MODE, bool GATHER = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_bf16_cg2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                     const __grid_constant__ CUtensorMap tmap_pad, const Gemm2Params p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + G2_STAGES * G2_A_BYTES;
    uint8_t* smem_c = smem + G2_STAGES * G2_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + G2_SLABS * G2_SLAB_BYTES);
    uint64_t* full_bar = bars;                                   // [STAGES]  used in the leader
    uint64_t* empty_bar = bars + G2_STAGES;                      // [STAGES]  one per CTA
    uint64_t* tmem_full_bar = bars + 2 * G2_STAGES;              // [2]       one per CTA
    uint64_t* tmem_empty_bar = bars + 2 * G2_STAGES + 2;         // [2]       used in the leader
    uint64_t* slab_ready_bar = bars + 2 * G2_STAGES + 4;         // [SLABS]   buffer free / residual landed
    uint64_t* slab_written_bar = bars + 2 * G2_STAGES + 4 + G2_SLABS;   // [SLABS] epilogue done with the slab
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4 + 2 * G2_SLABS);

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_out);
        if constexpr (MODE == G2_EPI_BIAS_RESIDUAL) tma_prefetch_desc(&tmap_res);
        if constexpr (GATHER) tma_prefetch_desc(&tmap_pad);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < G2_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 2 * G2_EPI_WARPS);   // one arrival per epilogue warp of both CTAs
        }
        for (int i = 0; i < G2_SLABS; ++i) {
            mbar_init(&slab_ready_bar[i], 1);
            mbar_init(&slab_written_bar[i], G2_EPI_WARPS / 2);  // the four warps that share the slab's column half
        }
        fence_mbar_init();
    }
    if (warp_idx == 2) {
        tmem_alloc_cg2(tmem_ptr_smem, G2_TMEM_COLS);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_kb = p.K / G2_BLOCK_K;
    const int num_tiles = p.num_m_blocks * p.num_n_blocks;

    if (warp_idx == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int tile_e = p.reverse ? num_tiles - 1 - tile : tile;
            const int m_blk = tile_e / p.num_n_blocks;
            const int n_blk = tile_e % p.num_n_blocks;
            const int m_coord = m_blk * G2_TILE_M + static_cast<int>(cta_rank) * 128;
            const int n_coord = n_blk * G2_TILE_N + static_cast<int>(cta_rank) * 128;
            // GATHER: source row of each of this CTA's four 32-row groups (the same for every k-block of the tile)
            int grow[4] = {0, 0, 0, 0};
            bool gpad[4] = {false, false, false, false};
            if constexpr (GATHER) {
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const long long g = (m_coord >> 5) + i;
                        if (g * 32 >= p.M) {
                            grow[i] = p.table_rows;                          // beyond the matrix: zero rows
                        } else {
                            const long long u = g / p.slots_per_user;
                            const int j = static_cast<int>(g - u * p.slots_per_user);
                            if (j < __ldg(p.gather_len + u)) {
                                const long long id = __ldg(p.gather_ids + g);
                                grow[i] = (id >= 0 && id * 32 < p.table_rows) ? static_cast<int>(id * 32) : p.table_rows;
                            } else {
                                grow[i] = j * 32;                            // padding slot: rows of the padding table
                                gpad[i] = true;
                            }
                        }
                    }
                }
            }
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (lane == 0) {
                    const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                    if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
                    if constexpr (GATHER) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            tma_load_2d_cg2(gpad[i] ? &tmap_pad : &tmap_a, full_leader,
                                            smem_a + stage * G2_A_BYTES + i * (32 * G2_BLOCK_K * 2), kb * G2_BLOCK_K, grow[i],
                                            kCacheEvictNormal);
                    } else
                    tma_load_2d_cg2(&tmap_a, full_leader, smem_a + stage * G2_A_BYTES, kb * G2_BLOCK_K, m_coord,
                                    kCacheEvictNormal);
                    tma_load_2d_cg2(&tmap_b, full_leader, smem_b + stage * G2_B_BYTES, kb * G2_BLOCK_K, n_coord,
                                    kCacheEvictNormal);
                }
                __syncwarp();
                if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== UMMA issuer (leader CTA only) =====================
        if (is_leader) {
            constexpr uint32_t idesc = umma_idesc_bf16(G2_TILE_M, G2_TILE_N);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t iter = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++iter) {
                const uint32_t as = iter & 1u;
                const uint32_t aphase = (iter >> 1) & 1u;
                mbar_wait_cluster(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * G2_TILE_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_u32(smem_a + stage * G2_A_BYTES);
                        const uint32_t b_addr = smem_u32(smem_b + stage * G2_B_BYTES);
#pragma unroll
                        for (int k = 0; k < G2_BLOCK_K / 16; ++k) {
                            const uint64_t da = umma_smem_desc_sw128(a_addr + k * 32);
                            const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32);
                            umma_bf16_ss_cg2(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_cg2_mc(&empty_bar[stage], 0x3);                       // both CTAs' smem slots
                        if (kb == num_kb - 1) umma_commit_cg2_mc(&tmem_full_bar[as], 0x3);  // both CTAs' epilogues
                    }
                    __syncwarp();
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 3) {
        // ===================== C mover (both CTAs) =====================
        // Slab order 0,2,1,3: the two epilogue column halves finish slabs (0,2) first, then (1,3).
        int prev_m = 0, prev_n = 0;
        uint32_t iter = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++iter) {
            const int tile_e = p.reverse ? num_tiles - 1 - tile : tile;
            const int m_blk = tile_e / p.num_n_blocks;
            const int n_blk = tile_e % p.num_n_blocks;
            const int m_coord = m_blk * G2_TILE_M + static_cast<int>(cta_rank) * 128;
            const int n_coord = n_blk * G2_TILE_N;
#pragma unroll 1
            for (int si = 0; si < G2_SLABS; ++si) {
                const int slab = ((si & 1) << 1) | (si >> 1);
                if (iter > 0) {
                    mbar_wait(&slab_written_bar[slab], (iter - 1) & 1u);
                    if (lane == 0) {
                        tma_store_2d(&tmap_out, smem_c + slab * G2_SLAB_BYTES, prev_n + slab * 64, prev_m);
                        bulk_commit_group();
                        bulk_wait_group_read0();     // the slab buffer may be overwritten from here on
                    }
                    __syncwarp();
                }
                if (lane == 0) {
                    if constexpr (MODE == G2_EPI_BIAS_RESIDUAL) {
                        mbar_arrive_expect_tx(&slab_ready_bar[slab], G2_SLAB_BYTES);
                        tma_load_2d(&tmap_res, &slab_ready_bar[slab], smem_c + slab * G2_SLAB_BYTES, n_coord + slab * 64,
                                    p.res_period > 0 ? m_coord % p.res_period : m_coord);
                    } else {
                        mbar_arrive(&slab_ready_bar[slab]);
                    }
                }
                __syncwarp();
            }
            prev_m = m_coord;
            prev_n = n_coord;
        }
        if (iter > 0) {
#pragma unroll 1
            for (int si = 0; si < G2_SLABS; ++si) {
                const int slab = ((si & 1) << 1) | (si >> 1);
                mbar_wait(&slab_written_bar[slab], (iter - 1) & 1u);
                if (lane == 0) {
                    tma_store_2d(&tmap_out, smem_c + slab * G2_SLAB_BYTES, prev_n + slab * 64, prev_m);
                    bulk_commit_group();
                }
                __syncwarp();
            }
            if (lane == 0) bulk_wait_group0();       // global writes complete before the CTA may exit
            __syncwarp();
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue (both CTAs) =====================
        const int q = warp_idx & 3;              // TMEM lane quadrant this warp may access
        const int half = (warp_idx - 4) >> 2;    // column half: slabs {0,1} or {2,3}
        const int r = q * 32 + lane;             // row inside this CTA's 128-row tile
        const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
        uint32_t iter = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++iter) {
            const int n_blk = (p.reverse ? num_tiles - 1 - tile : tile) % p.num_n_blocks;
            const uint32_t as = iter & 1u;
            const uint32_t aphase = (iter >> 1) & 1u;
            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + as * G2_TILE_N + (static_cast<uint32_t>(q * 32) << 16);
            // LayerNorm folding: this thread's row scalars (rows beyond M: neutral values, their outputs are clipped)
            const long long grow = static_cast<long long>((p.reverse ? num_tiles - 1 - tile : tile) / p.num_n_blocks) * G2_TILE_M +
                                   cta_rank * 128 + r;
            float in_rstd = 1.f, in_shift = 0.f, res_rstd = 1.f, res_shift = 0.f, st1 = 0.f, st2 = 0.f;
            if (p.ln_in_stats != nullptr && grow < p.M) {
                const float2* st = reinterpret_cast<const float2*>(p.ln_in_stats) + grow * p.ln_parts;
                float s1 = 0.f, s2 = 0.f;
                for (int i = 0; i < p.ln_parts; ++i) { const float2 t = __ldg(st + i); s1 += t.x; s2 += t.y; }
                const float mu = s1 * p.ln_inv_h;
                in_rstd = rsqrtf(fmaxf(s2 * p.ln_inv_h - mu * mu, 0.f) + p.ln_eps);
                in_shift = -mu * in_rstd;
            }
            if (MODE == G2_EPI_BIAS_RESIDUAL && p.ln_res_stats != nullptr && grow < p.M) {
                const float2* st = reinterpret_cast<const float2*>(p.ln_res_stats) + grow * p.ln_parts;
                float s1 = 0.f, s2 = 0.f;
                for (int i = 0; i < p.ln_parts; ++i) { const float2 t = __ldg(st + i); s1 += t.x; s2 += t.y; }
                const float mu = s1 * p.ln_inv_h;
                res_rstd = rsqrtf(fmaxf(s2 * p.ln_inv_h - mu * mu, 0.f) + p.ln_eps);
                res_shift = -mu * res_rstd;
            }
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                const int slab = half * 2 + s;
                uint8_t* slab_smem = smem_c + slab * G2_SLAB_BYTES;
                mbar_wait(&slab_ready_bar[slab], iter & 1u);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col_in_tile = slab * 64 + c * 32;
                    const int n0 = n_blk * G2_TILE_N + col_in_tile;
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_acc + col_in_tile, v);
                    uint4 res[4];
                    if constexpr (MODE == G2_EPI_BIAS_RESIDUAL) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            res[j] = *reinterpret_cast<const uint4*>(slab_smem + swz128(r, c * 4 + j));
                    }
                    tmem_ld_wait();
                    if (s == 1 && c == 1) {
                        // every TMEM read of this accumulator stage is in registers: hand it back to the issuer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tmem_empty_leader0 + as * 8);
                    }
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (p.ln_in_stats != nullptr) {
                        // A = pre of a folded LayerNorm: y = rstd acc + (-mu rstd) c[n] + b'[n]
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 4));
                            const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.ln_in_c + n0 + j * 4));
                            f[4 * j + 0] = fmaf(f[4 * j + 0], in_rstd, fmaf(in_shift, c4.x, b.x));
                            f[4 * j + 1] = fmaf(f[4 * j + 1], in_rstd, fmaf(in_shift, c4.y, b.y));
                            f[4 * j + 2] = fmaf(f[4 * j + 2], in_rstd, fmaf(in_shift, c4.z, b.z));
                            f[4 * j + 3] = fmaf(f[4 * j + 3], in_rstd, fmaf(in_shift, c4.w, b.w));
                        }
                    } else if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 4));
                            f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                        }
                    }
                    if constexpr (MODE == G2_EPI_BIAS_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) gelu_erf_x2(f[j], f[j + 1]);
                    }
                    if constexpr (MODE == G2_EPI_BIAS_RESIDUAL) {
                        if (p.ln_res_stats != nullptr) {
                            // residual = LayerNorm of the tile just loaded: h = x (rstd gamma) + (beta - mu rstd gamma)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t w[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
                                const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_res_gamma + n0 + 8 * j));
                                const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.ln_res_gamma + n0 + 8 * j) + 1);
                                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln_res_beta + n0 + 8 * j));
                                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.ln_res_beta + n0 + 8 * j) + 1);
                                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    f[8 * j + 2 * t] += fmaf(bf16_lo(w[t]), res_rstd * g[2 * t], fmaf(res_shift, g[2 * t], bb[2 * t]));
                                    f[8 * j + 2 * t + 1] +=
                                        fmaf(bf16_hi(w[t]), res_rstd * g[2 * t + 1], fmaf(res_shift, g[2 * t + 1], bb[2 * t + 1]));
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t w[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    f[8 * j + 2 * t] += bf16_lo(w[t]);
                                    f[8 * j + 2 * t + 1] += bf16_hi(w[t]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 o = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                                   pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                        *reinterpret_cast<uint4*>(slab_smem + swz128(r, c * 4 + j)) = o;
                        if (p.stats_out != nullptr) {
                            // statistics of the values as stored (bf16): what the consumers of this tensor will read
                            const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float a = bf16_lo(w[t]), b = bf16_hi(w[t]);
                                st1 += a + b;
                                st2 = fmaf(a, a, fmaf(b, b, st2));
                            }
                        }
                    }
                }
                fence_proxy_async_smem();        // generic-proxy smem writes -> visible to the TMA store
                __syncwarp();
                if (lane == 0) mbar_arrive(&slab_written_bar[slab]);
            }
            if (p.stats_out != nullptr && grow < p.M)     // this thread's 128 columns of the row: its own slot, no atomics
                reinterpret_cast<float2*>(p.stats_out)[(grow * p.num_n_blocks + n_blk) * 2 + half] = make_float2(st1, st2);
        }
    }

    // ---- teardown: nobody may exit (or free TMEM) while the pair still uses this CTA's smem / barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, G2_TMEM_COLS);
    }
}

// (sum, sum of squares) partials per row that a LayerNorm-producer call writes for N output columns: one per column half of
// every 256-column tile (unirec_linear_ln_stats_parts)
long long gemm_cg2_stats_parts(long long N) { return 2 * (N / G2_TILE_N); }

template <int MODE, bool GATHER = false>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                        const Gemm2Params& p, int max_ctas, cudaStream_t stream, const CUtensorMap* tpad = nullptr) {
    auto kern = gemm_bf16_cg2_kernel<MODE, GATHER>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES);
        if (e != cudaSuccess) {
            set_last_error("cudaFuncSetAttribute(smem=%d): %s", G2_SMEM_BYTES, cudaGetErrorString(e));
            return UNIREC_ERR_CUDA;
        }
        attr_set = true;
    }
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    int clusters = num_sms() / 2;
    if (max_ctas > 0 && clusters > max_ctas / 2) clusters = max_ctas / 2 > 0 ? max_ctas / 2 : 1;
    if (clusters > tiles) clusters = tiles;
    kern<<<2 * clusters, G2_THREADS, G2_SMEM_BYTES, stream>>>(ta, tb, to, tr, tpad != nullptr ? *tpad : ta, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("gemm (cta pair) launch failed: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

// Shapes the CTA-pair kernel accepts (the caller falls back to the single-CTA kernel otherwise).
bool gemm_cg2_supported(long long M, long long N, long long K, int out_fp32, int res_row_mod, long long lda,
                        long long ldw, long long ldo, long long ldr, int epilogue) {
    (void)M;
    if (out_fp32 || res_row_mod > 0) return false;
    if (N % G2_TILE_N != 0 || K % G2_BLOCK_K != 0) return false;
    if (lda % 8 != 0 || ldw % 8 != 0 || ldo % 8 != 0) return false;
    if (epilogue == G2_EPI_BIAS_RESIDUAL && ldr % 8 != 0) return false;
    return true;
}

int gemm_bf16_cg2(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* residual,
                  long long ldr, void* out, long long ldo, long long M, long long N, long long K, int epilogue,
                  int max_ctas, cudaStream_t stream, const LnFold* ln) {
    Gemm2Params p;
    p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
    p.bias = bias;
    p.num_m_blocks = static_cast<int>((M + G2_TILE_M - 1) / G2_TILE_M);
    p.num_n_blocks = static_cast<int>(N / G2_TILE_N);
    p.gather_ids = nullptr; p.gather_len = nullptr; p.slots_per_user = 0; p.table_rows = 0; p.res_period = 0;
    p.reverse = 0;      // GEMMs walk forward; reversing them as well (alternating directions) measured slower, see common.cuh
    p.ln_in_stats = ln != nullptr ? ln->in_stats : nullptr;
    p.ln_in_c = ln != nullptr ? ln->in_c : nullptr;
    p.ln_res_stats = ln != nullptr ? ln->res_stats : nullptr;
    p.ln_res_gamma = ln != nullptr ? ln->res_gamma : nullptr;
    p.ln_res_beta = ln != nullptr ? ln->res_beta : nullptr;
    p.stats_out = ln != nullptr ? ln->stats_out : nullptr;
    p.ln_parts = ln != nullptr ? ln->parts : 0;
    p.ln_eps = ln != nullptr ? ln->eps : 0.f;
    p.ln_inv_h = ln != nullptr ? 1.0f / static_cast<float>(ln->hidden) : 0.f;
    CUtensorMap ta, tb, to, tr;
    int rc = make_tmap_bf16_2d(&ta, A, M, K, lda, 128);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tb, W, N, K, ldw, 128);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&to, out, M, N, ldo, 128);
    if (rc != UNIREC_OK) return rc;
    if (epilogue == G2_EPI_BIAS_RESIDUAL) {
        rc = make_tmap_bf16_2d(&tr, residual, M, N, ldr, 128);
        if (rc != UNIREC_OK) return rc;
    } else {
        tr = to;
    }
    switch (epilogue) {
        case G2_EPI_BIAS: return launch_gemm2<G2_EPI_BIAS>(ta, tb, to, tr, p, max_ctas, stream);
        case G2_EPI_BIAS_GELU: return launch_gemm2<G2_EPI_BIAS_GELU>(ta, tb, to, tr, p, max_ctas, stream);
        case G2_EPI_BIAS_RESIDUAL: return launch_gemm2<G2_EPI_BIAS_RESIDUAL>(ta, tb, to, tr, p, max_ctas, stream);
        default: set_last_error("gemm_bf16: unknown epilogue %d", epilogue); return UNIREC_ERR_BAD_ARG;
    }
}

// out[M, N] = gather(table)[M, K] x W[N, K]^T + bias + posbias[m % period, :]   (bf16 out)
// The cross-attention K/V projection of the user Q-Former straight from the resident item-token table: row group g of
// the virtual user-sequence matrix (M = users x slots_per_user x 32 rows) is the 32 tokens of item ids[g]; the sinusoidal
// position term of models/user_sequence_encoder.py:128-140 enters as the precomputed tile posbias = PE x W^T (linear in
// A), and slots beyond a user's length read pad_table = -PE so that their rows come out as the bias alone - what the
// reference's zero-padded sequence gives (training/user_qformer_training.py:153-161).
int gemm_bf16_cg2_gather(const void* table, long long ld_table, long long table_rows, const long long* ids,
                         const int* lengths, long long slots_per_user, const void* pad_table, long long ld_pad,
                         long long pad_rows, const void* W, long long ldw, const float* bias, const void* posbias,
                         long long ld_pos, long long pos_rows, long long period, void* out, long long ldo, long long M,
                         long long N, long long K, cudaStream_t stream) {
    if (table == nullptr || ids == nullptr || lengths == nullptr || pad_table == nullptr || W == nullptr ||
        posbias == nullptr || out == nullptr || M <= 0 || slots_per_user <= 0) {
        set_last_error("linear_gather: null pointer or empty shape");
        return UNIREC_ERR_BAD_ARG;
    }
    if (N % G2_TILE_N != 0 || K % G2_BLOCK_K != 0 || M % 32 != 0 || period != slots_per_user * 32 ||
        M % period != 0 || pos_rows < period + 128 || pad_rows < period || table_rows % 32 != 0 ||
        table_rows >= 2147483647LL - 64 || ld_table % 8 != 0 || ld_pad % 8 != 0 || ldw % 8 != 0 || ld_pos % 8 != 0 ||
        ldo % 8 != 0) {
        set_last_error("linear_gather: needs N %% 256 == 0, K %% 64 == 0, 32-token slots, period = slots x 32, posbias with "
                       "period + 128 rows, 16-byte aligned rows (M=%lld N=%lld K=%lld period=%lld)", M, N, K, period);
        return UNIREC_ERR_BAD_ARG;
    }
    Gemm2Params p;
    p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
    p.bias = bias;
    p.num_m_blocks = static_cast<int>((M + G2_TILE_M - 1) / G2_TILE_M);
    p.num_n_blocks = static_cast<int>(N / G2_TILE_N);
    p.gather_ids = ids; p.gather_len = lengths; p.slots_per_user = static_cast<int>(slots_per_user);
    p.table_rows = static_cast<int>(table_rows); p.res_period = static_cast<int>(period);
    p.reverse = 0;
    p.ln_in_stats = p.ln_in_c = p.ln_res_stats = p.ln_res_gamma = p.ln_res_beta = nullptr;
    p.stats_out = nullptr; p.ln_parts = 0; p.ln_eps = 0.f; p.ln_inv_h = 0.f;
    CUtensorMap ta, tb, to, tr, tp;
    int rc = make_tmap_bf16_2d(&ta, table, table_rows, K, ld_table, 32);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tp, pad_table, pad_rows, K, ld_pad, 32);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tb, W, N, K, ldw, 128);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&to, out, M, N, ldo, 128);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tr, posbias, pos_rows, N, ld_pos, 128);
    if (rc != UNIREC_OK) return rc;
    return launch_gemm2<G2_EPI_BIAS_RESIDUAL, true>(ta, tb, to, tr, p, 0, stream, &tp);
}

}  // namespace unirec
