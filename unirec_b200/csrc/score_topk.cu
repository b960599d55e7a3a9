// Fused cosine scoring + top-k over a large candidate pool (sm_100a).
//
//   score[b, n] = <users[b,:], cands[n,:]> * user_inv[b] * cand_inv[n]          (cosine similarity)
//   out[b, :]   = the k best candidates of row b, descending
//
// Replaces, at scale, the ranking idiom of training/train_item_individual_token_joint.py:405-415
// (F.normalize both sides, matmul, argsort descending) - the reference scores <= 100 candidates per
// user in a Python loop; here B users x N = 1M candidates are ranked without ever materialising the
// [B, N] score matrix (16 GB at B = 4096).
//
// Kernel 0 (score_tile_kernel<MODE_DENSE>, large pools only): the same tcgen05 pipeline scores a strided
// SAMPLE of the candidate tiles (1/32 of the pool, 8192..32768 candidates) and stores those scores densely;
// row_kth_kernel then finds, per user, the k-th best sampled score tau_s by bisection over values held in
// registers.  tau_s is a valid lower bound of the user's true k-th best score (k candidates >= tau_s exist),
// so the main pass starts its filter at tau_s instead of -inf and only ~k*N/N_sample scores per user ever
// leave the epilogue.  Small pools (N <= 32768) are scored densely and ranked by the same row kernel.
//
// Kernel 1 (score_tile_kernel<MODE_FILTER>): the tcgen05/TMEM/TMA pipeline of umma_pipe.cuh computes
// 128 x 256 score tiles.  The epilogue thread that owns (row, column-half) keeps a running threshold tau in a
// register and appends only candidates with score > tau to its private list in global scratch memory
// (interleaved [slot][lane] so that a warp scanning its 32 lists in lock-step reads coalesced lines).
// When a list is nearly full the warp compacts its lists: each lane bisects (on the order-preserving
// integer image of the float) for a threshold that keeps between k and KMAX entries, rewrites its
// list in place and raises tau (with the sampled start threshold this only happens for adversarially
// ordered pools, but it keeps the result exact for any input).  Almost nothing passes the filter, so the
// epilogue costs one multiply and one max per score.  The candidate range is split into R contiguous ranges; a work
// item = (128-user block, range); work items are dealt round-robin to persistent CTAs so that the CTAs
// that stream the same candidate range run concurrently (candidates come from HBM once, then L2).
//
// Kernel 2 (topk_select_kernel): per user, exact selection of the k largest among the <= 2*R*KMAX
// surviving entries (gathered into shared memory, bisection for the k-th value, then a 128-wide bitonic
// sort), scaled by user_inv.
// The same kernel merges per-GPU top-k lists after the NCCL all-gather (unirec_topk_merge).
#include "common.cuh"
#include "cg2_ptx.cuh"
#include "umma_pipe.cuh"
#include "../../include/unirec_b200.h"

#include <atomic>

namespace unirec {

extern std::atomic<long long> g_launch_count;

constexpr int SC_BLOCK_N = 256;
constexpr int SC_THREADS = 384;
constexpr int SC_EPI_THREADS = 256;
constexpr int SC_CAP = 512;    // scratch list capacity per (row, column half)
constexpr int SC_KMAX = 128;   // max k, and max entries kept by a compaction
constexpr int SC_DENSE_MAX_N = 32768;   // pools up to this size are scored densely (row_kth_kernel holds a row in registers)
constexpr int SC_ROW_THREADS = 512;
constexpr int SC_SEL_CAP = 8192;        // entries the select kernel can hold in shared memory
constexpr int SC_MAX_USERS_PER_PASS = 4096;
using ScPipe = UmmaPipe<SC_BLOCK_N, 4>;
// CTA-pair variant (more than 128 users per pass): a cluster of two CTAs owns a 256-user x 256-candidate tile
constexpr int SP_TILE_M = 256;
constexpr int SP_STAGES = 5;
constexpr int SP_A_BYTES = 128 * PIPE_BLOCK_K * 2;          // this CTA's 128 user rows
constexpr int SP_B_BYTES = 128 * PIPE_BLOCK_K * 2;          // this CTA's 128 candidate rows
constexpr int SP_STAGE_BYTES = SP_A_BYTES + SP_B_BYTES;
constexpr int SP_CI_BYTES = 2 * SC_BLOCK_N * 4;             // cand_inv of the current and the next tile
constexpr int SP_BARRIER_BYTES = 256;
constexpr int SP_SMEM_BYTES = SP_STAGES * SP_STAGE_BYTES + SP_CI_BYTES + 1024 + SP_BARRIER_BYTES;
static_assert(SP_SMEM_BYTES <= 232448, "shared memory budget exceeded");

enum : int { MODE_FILTER = 0, MODE_DENSE = 1 };

UNIREC_DEVICE uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
UNIREC_DEVICE float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
constexpr uint32_t ORD_NEG_INF = 0x007fffffu;   // f2ord(-inf): every finite score has a larger ordinal

struct ScoreParams {
    int B, N, D, k;
    const float* cand_inv;
    int num_m_blocks, n_tiles, R, tiles_per_range;
    // MODE_FILTER
    const float* row_tau;   // [B] start threshold per user (scores >= row_tau pass) or nullptr (-inf)
    uint2* scratch;         // [grid][8][SC_CAP][32]
    uint2* partial;         // [B_pad][2R][SC_KMAX]   (score bits, candidate index)
    int* partial_cnt;       // [B_pad][2R]
    // MODE_DENSE: tile j of the work list is candidate tile j * tile_stride; scores go to dense[row][j*256 + col]
    float* dense;
    long long ldd;
    int n_dense_tiles, tile_stride;
};

// Warp-lock-step compaction of the 32 per-lane lists of one epilogue warp.
// On return every participating lane has k <= cnt <= SC_KMAX (or cnt unchanged if it was <= SC_KMAX)
// and tau raised so that (score > tau) can only admit candidates that may still belong to the top k.
UNIREC_DEVICE void compact_lists(uint2* base, int lane, int& cnt, float& tau, int k) {
    const bool active = cnt > SC_KMAX;
    int maxcnt = active ? cnt : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));
    if (maxcnt == 0) return;
    // hi: smallest ordinal known to have count(>= hi) < k;  lo: ordinal with count(>= lo) >= k
    uint32_t mx = 0;
    for (int j = 0; j < maxcnt; ++j)
        if (active && j < cnt) mx = max(mx, f2ord(__uint_as_float(base[j * 32 + lane].x)));
    uint32_t lo = f2ord(tau);   // every stored entry is > tau
    uint32_t hi = mx + 1u;
    bool done = !active;
    for (int it = 0; it < 40; ++it) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (mid == lo) done = true;   // lo is the exact k-th ordinal (ties may exceed KMAX; handled below)
        if (__all_sync(0xffffffffu, done)) break;
        int c = 0;
        for (int j = 0; j < maxcnt; ++j)
            if (!done && j < cnt) c += (f2ord(__uint_as_float(base[j * 32 + lane].x)) >= mid) ? 1 : 0;
        if (!done) {
            if (c >= k) {
                lo = mid;
                if (c <= SC_KMAX) done = true;
            } else {
                hi = mid;
            }
        }
    }
    // in-place rewrite: keep ord > lo, plus ord == lo while room remains
    int c_gt = 0;
    for (int j = 0; j < maxcnt; ++j)
        if (active && j < cnt) c_gt += (f2ord(__uint_as_float(base[j * 32 + lane].x)) > lo) ? 1 : 0;
    int eq_quota = SC_KMAX - c_gt;
    int w = 0;
    for (int j = 0; j < maxcnt; ++j) {
        if (active && j < cnt) {
            const uint2 e = base[j * 32 + lane];
            const uint32_t o = f2ord(__uint_as_float(e.x));
            bool keep = o > lo;
            if (o == lo && eq_quota > 0) { keep = true; --eq_quota; }
            if (keep) { base[w * 32 + lane] = e; ++w; }
        }
    }
    if (active) {
        cnt = w;
        // ordinal lo back to float: entries equal to the new tau that arrive later are ties with kept ones
        tau = ord2f(lo);
    }
}

// Work list.  MODE_FILTER: item w = (m_blk = w % num_m_blocks, range r = w / num_m_blocks) covers candidate
// tiles [r*tiles_per_range, ...).  MODE_DENSE: item w = (m_blk, j = w / num_m_blocks) is the single sampled
// tile j*tile_stride.  Either way consecutive CTAs work on the same candidate tiles at the same time, so a
// candidate tile is fetched from HBM once and then served from L2.
template <int MODE>
UNIREC_DEVICE void work_item(const ScoreParams& p, int w, int& m_blk, int& r, int& t0, int& t1, int& tstep) {
    m_blk = w % p.num_m_blocks;
    r = w / p.num_m_blocks;
    if constexpr (MODE == MODE_FILTER) {
        t0 = r * p.tiles_per_range;
        t1 = min(t0 + p.tiles_per_range, p.n_tiles);
        tstep = 1;
    } else {
        t0 = r * p.tile_stride;
        t1 = t0 + 1;
        tstep = 1;
    }
}

template <int MODE>
__global__ void __launch_bounds__(SC_THREADS, 1)
score_tile_kernel(const __grid_constant__ CUtensorMap tmap_users, const __grid_constant__ CUtensorMap tmap_cands,
                  const ScoreParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_users);
        tma_prefetch_desc(&tmap_cands);
    }
    ScPipe pipe;
    pipe.setup(smem_raw, warp_idx, lane, SC_EPI_THREADS);
    const int num_kb = p.D / PIPE_BLOCK_K;
    const int num_items = p.num_m_blocks * (MODE == MODE_FILTER ? p.R : p.n_dense_tiles);

    if (warp_idx == 0) {
        RingState rs;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
            int m_blk, r, t0, t1, ts;
            work_item<MODE>(p, w, m_blk, r, t0, t1, ts);
            for (int t = t0; t < t1; t += ts)
                pipe_produce_tile(pipe, rs, &tmap_users, &tmap_cands, m_blk * PIPE_BLOCK_M, t * SC_BLOCK_N, num_kb, lane);
        }
    } else if (warp_idx == 1) {
        RingState rs;
        uint32_t iter = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
            int m_blk, r, t0, t1, ts;
            work_item<MODE>(p, w, m_blk, r, t0, t1, ts);
            for (int t = t0; t < t1; t += ts, ++iter) pipe_mma_tile<SC_BLOCK_N>(pipe, rs, iter, num_kb, lane);
        }
    } else if (warp_idx >= 4) {
        const int q = warp_idx & 3;
        const int half = (warp_idx - 4) >> 2;
        uint2* base = nullptr;
        if constexpr (MODE == MODE_FILTER)
            base = p.scratch + (static_cast<size_t>(blockIdx.x) * 8 + (warp_idx - 4)) * SC_CAP * 32;
        const int L = 2 * p.R;
        uint32_t iter = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
            int m_blk, r, t0, t1, ts;
            work_item<MODE>(p, w, m_blk, r, t0, t1, ts);
            const int row = m_blk * PIPE_BLOCK_M + q * 32 + lane;
            const bool row_ok = row < p.B;
            int cnt = 0;
            float tau = -INFINITY;
            if constexpr (MODE == MODE_FILTER) {
                if (p.row_tau != nullptr && row_ok) {
                    // scores >= row_tau must pass the strict (score > tau) filter: step one ordinal down
                    const float ts_ = __ldg(p.row_tau + row);
                    if (ts_ > -INFINITY) tau = ord2f(f2ord(ts_) - 1u);
                }
            }
            for (int t = t0; t < t1; t += ts, ++iter) {
                const uint32_t tmem_acc = pipe_epilogue_wait<SC_BLOCK_N>(pipe, iter);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int col = half * 128 + c * 32;
                    const int n0 = t * SC_BLOCK_N + col;
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_acc + col + (static_cast<uint32_t>(q * 32) << 16), v);
                    float ci[32];
                    if (n0 + 32 <= p.N) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 x = __ldg(reinterpret_cast<const float4*>(p.cand_inv + n0) + j);
                            ci[4 * j] = x.x; ci[4 * j + 1] = x.y; ci[4 * j + 2] = x.z; ci[4 * j + 3] = x.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) ci[j] = (n0 + j < p.N) ? __ldg(p.cand_inv + n0 + j) : 0.f;
                    }
                    tmem_ld_wait();
                    if (c == 3) pipe_epilogue_release(pipe, iter);
                    float s[32];
                    if constexpr (MODE == MODE_FILTER) {
                        float mx = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            s[j] = __uint_as_float(v[j]) * ci[j];
                            mx = fmaxf(mx, s[j]);
                        }
                        if (row_ok && mx > tau) {
                            const bool full_chunk = n0 + 32 <= p.N;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (s[j] > tau && (full_chunk || n0 + j < p.N)) {
                                    base[cnt * 32 + lane] = make_uint2(__float_as_uint(s[j]), static_cast<uint32_t>(n0 + j));
                                    ++cnt;
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            s[j] = (n0 + j < p.N) ? __uint_as_float(v[j]) * ci[j] : -INFINITY;
                        if (row_ok) {
                            float* o = p.dense + static_cast<long long>(row) * p.ldd + r * SC_BLOCK_N + col;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                *reinterpret_cast<float4*>(o + 4 * j) = make_float4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
                        }
                    }
                }
                if constexpr (MODE == MODE_FILTER) {
                    // room for the next tile's (at most 128) appends?
                    if (__any_sync(0xffffffffu, cnt > SC_CAP - 128)) compact_lists(base, lane, cnt, tau, p.k);
                }
            }
            if constexpr (MODE == MODE_FILTER) {
                if (__any_sync(0xffffffffu, cnt > SC_KMAX)) compact_lists(base, lane, cnt, tau, p.k);
                if (row_ok) {
                    const size_t list = static_cast<size_t>(row) * L + (r * 2 + half);
                    p.partial_cnt[list] = cnt;
                    uint2* dst = p.partial + list * SC_KMAX;
                    for (int j = 0; j < cnt; ++j) dst[j] = base[j * 32 + lane];
                }
            }
        }
    }
    pipe.teardown(warp_idx);
}

// ---------------------------------------------------------------------------------------------
// score_pair_kernel: the same two modes on the CTA-pair pipeline of gemm_cg2.cu.  A cluster of two CTAs owns a
// 256-user x 256-candidate tile: each CTA stages ITS 128 user rows and ITS 128 candidate rows per 64-wide k block, the
// leader issues tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16) and each CTA's 128 user rows x 256 scores land in its own
// TMEM - a candidate tile is read from shared memory once per PAIR of user blocks (half the smem operand reads per flop
// of the single-CTA kernel; on a power-capped part that is clock).  The epilogue is the single-CTA one (thread = one
// user row x one column half, running threshold in a register, private list in global scratch) with one change that ncu
// asked for (profiles/r01_f_ncu_stalls_score_tile.txt: 43 % long-scoreboard stalls on the cand_inv loads): the 256
// candidate norms of a tile are fetched ONE TILE AHEAD by the 256 epilogue threads (one coalesced 1 KB read) into a
// double-buffered shared array, so the per-chunk operand is a broadcast LDS instead of eight dependent global loads.
// ---------------------------------------------------------------------------------------------
UNIREC_DEVICE void sp_epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SC_EPI_THREADS) : "memory"); }

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SC_THREADS, 1)
score_pair_kernel(const __grid_constant__ CUtensorMap tmap_users, const __grid_constant__ CUtensorMap tmap_cands,
                  const ScoreParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + SP_STAGES * SP_A_BYTES;
    float* smem_ci = reinterpret_cast<float*>(smem + SP_STAGES * SP_STAGE_BYTES);        // [2][256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SP_STAGES * SP_STAGE_BYTES + SP_CI_BYTES);
    uint64_t* full_bar = bars;                              // [STAGES]  used in the leader
    uint64_t* empty_bar = bars + SP_STAGES;                 // [STAGES]  one per CTA
    uint64_t* tmem_full_bar = bars + 2 * SP_STAGES;         // [2]       one per CTA
    uint64_t* tmem_empty_bar = bars + 2 * SP_STAGES + 2;    // [2]       used in the leader
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * SP_STAGES + 4);

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_users);
        tma_prefetch_desc(&tmap_cands);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < SP_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 2 * (SC_EPI_THREADS / 32));   // one arrival per epilogue warp of both CTAs
        }
        fence_mbar_init();
    }
    if (warp_idx == 2) {
        tmem_alloc_cg2(tmem_ptr_smem, 2 * SC_BLOCK_N);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_kb = p.D / PIPE_BLOCK_K;
    const int num_items = p.num_m_blocks * (MODE == MODE_FILTER ? p.R : p.n_dense_tiles);

    if (warp_idx == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int w = cluster_id; w < num_items; w += num_clusters) {
            int m_blk, r, t0, t1, ts;
            work_item<MODE>(p, w, m_blk, r, t0, t1, ts);
            const int m_coord = m_blk * SP_TILE_M + static_cast<int>(cta_rank) * 128;
            for (int t = t0; t < t1; t += ts) {
                const int n_coord = t * SC_BLOCK_N + static_cast<int>(cta_rank) * 128;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (lane == 0) {
                        const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * SP_STAGE_BYTES);
                        tma_load_2d_cg2(&tmap_users, full_leader, smem_a + stage * SP_A_BYTES, kb * PIPE_BLOCK_K, m_coord,
                                        kCacheEvictLast);
                        tma_load_2d_cg2(&tmap_cands, full_leader, smem_b + stage * SP_B_BYTES, kb * PIPE_BLOCK_K, n_coord,
                                        kCacheEvictNormal);
                    }
                    __syncwarp();
                    if (++stage == SP_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== UMMA issuer (leader CTA only) =====================
        if (is_leader) {
            constexpr uint32_t idesc = umma_idesc_bf16(SP_TILE_M, SC_BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t iter = 0;
            for (int w = cluster_id; w < num_items; w += num_clusters) {
                int m_blk, r, t0, t1, ts;
                work_item<MODE>(p, w, m_blk, r, t0, t1, ts);
                for (int t = t0; t < t1; t += ts, ++iter) {
                    const uint32_t as = iter & 1u;
                    const uint32_t aphase = (iter >> 1) & 1u;
                    mbar_wait_cluster(&tmem_empty_bar[as], aphase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * SC_BLOCK_N;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint32_t a_addr = smem_u32(smem_a + stage * SP_A_BYTES);
                            const uint32_t b_addr = smem_u32(smem_b + stage * SP_B_BYTES);
#pragma unroll
                            for (int k = 0; k < PIPE_BLOCK_K / 16; ++k) {
                                const uint64_t da = umma_smem_desc_sw128(a_addr + k * 32);
                                const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32);
                                umma_bf16_ss_cg2(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                            umma_commit_cg2_mc(&empty_bar[stage], 0x3);
                            if (kb == num_kb - 1) umma_commit_cg2_mc(&tmem_full_bar[as], 0x3);
                        }
                        __syncwarp();
                        if (++stage == SP_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue (both CTAs) =====================
        const int q = warp_idx & 3;
        const int half = (warp_idx - 4) >> 2;
        const int et = threadIdx.x - 128;                   // 0..255: the column whose norm this thread prefetches
        const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
        uint2* base = nullptr;
        if constexpr (MODE == MODE_FILTER)
            base = p.scratch + (static_cast<size_t>(blockIdx.x) * 8 + (warp_idx - 4)) * SC_CAP * 32;
        const int L = 2 * p.R;
        uint32_t iter = 0;
        for (int w = cluster_id; w < num_items; w += num_clusters) {
            int m_blk, r, t0, t1, ts;
            work_item<MODE>(p, w, m_blk, r, t0, t1, ts);
            const int row = m_blk * SP_TILE_M + static_cast<int>(cta_rank) * 128 + q * 32 + lane;
            const bool row_ok = row < p.B;
            int cnt = 0;
            float tau = -INFINITY;
            if constexpr (MODE == MODE_FILTER) {
                if (p.row_tau != nullptr && row_ok) {
                    // scores >= row_tau must pass the strict (score > tau) filter: step one ordinal down
                    const float ts_ = __ldg(p.row_tau + row);
                    if (ts_ > -INFINITY) tau = ord2f(f2ord(ts_) - 1u);
                }
            }
            for (int t = t0; t < t1; t += ts, ++iter) {
                float* ci_cur = smem_ci + (iter & 1u) * SC_BLOCK_N;
                if (t == t0) {
                    const int n = t * SC_BLOCK_N + et;
                    ci_cur[et] = n < p.N ? __ldg(p.cand_inv + n) : 0.f;
                }
                const bool has_next = t + ts < t1;
                float ci_next = 0.f;
                if (has_next) {
                    const int n = (t + ts) * SC_BLOCK_N + et;
                    ci_next = n < p.N ? __ldg(p.cand_inv + n) : 0.f;
                }
                // this tile's norms are visible; every epilogue thread has finished the previous tile (the other buffer)
                sp_epi_bar_sync();
                const uint32_t as = iter & 1u;
                const uint32_t aphase = (iter >> 1) & 1u;
                mbar_wait(&tmem_full_bar[as], aphase);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + as * SC_BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int col = half * 128 + c * 32;
                    const int n0 = t * SC_BLOCK_N + col;
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_acc + col, v);
                    float ci[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 x = *reinterpret_cast<const float4*>(ci_cur + col + 4 * j);
                        ci[4 * j] = x.x; ci[4 * j + 1] = x.y; ci[4 * j + 2] = x.z; ci[4 * j + 3] = x.w;
                    }
                    tmem_ld_wait();
                    if (c == 3) {
                        // every TMEM read of this accumulator stage is in registers: hand it back to the issuer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tmem_empty_leader0 + as * 8);
                    }
                    float s[32];
                    if constexpr (MODE == MODE_FILTER) {
                        float mx = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            s[j] = __uint_as_float(v[j]) * ci[j];
                            mx = fmaxf(mx, s[j]);
                        }
                        if (row_ok && mx > tau) {
                            const bool full_chunk = n0 + 32 <= p.N;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (s[j] > tau && (full_chunk || n0 + j < p.N)) {
                                    base[cnt * 32 + lane] = make_uint2(__float_as_uint(s[j]), static_cast<uint32_t>(n0 + j));
                                    ++cnt;
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            s[j] = (n0 + j < p.N) ? __uint_as_float(v[j]) * ci[j] : -INFINITY;
                        if (row_ok) {
                            float* o = p.dense + static_cast<long long>(row) * p.ldd + r * SC_BLOCK_N + col;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                *reinterpret_cast<float4*>(o + 4 * j) = make_float4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
                        }
                    }
                }
                if (has_next) smem_ci[((iter & 1u) ^ 1u) * SC_BLOCK_N + et] = ci_next;
                if constexpr (MODE == MODE_FILTER) {
                    // room for the next tile's (at most 128) appends?
                    if (__any_sync(0xffffffffu, cnt > SC_CAP - 128)) compact_lists(base, lane, cnt, tau, p.k);
                }
            }
            if constexpr (MODE == MODE_FILTER) {
                if (__any_sync(0xffffffffu, cnt > SC_KMAX)) compact_lists(base, lane, cnt, tau, p.k);
                if (row_ok) {
                    const size_t list = static_cast<size_t>(row) * L + (r * 2 + half);
                    p.partial_cnt[list] = cnt;
                    uint2* dst = p.partial + list * SC_KMAX;
                    for (int j = 0; j < cnt; ++j) dst[j] = base[j * 32 + lane];
                }
            }
        }
    }

    // ---- teardown: nobody may exit (or free TMEM) while the pair still uses this CTA's smem / barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, 2 * SC_BLOCK_N);
    }
}

// ---------------------------------------------------------------------------------------------
// Block-wide helpers shared by the two row kernels
// ---------------------------------------------------------------------------------------------
template <int NWARPS>
UNIREC_DEVICE int block_sum_int(int v, int* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int i = 0; i < NWARPS; ++i) t += s_red[i];
    return t;
}

template <int NWARPS>
UNIREC_DEVICE uint32_t block_max_u32(uint32_t v, uint32_t* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < NWARPS; ++i) t = max(t, s_red[i]);
    return t;
}

// Bitonic sort of SC_KMAX (score desc, index asc) in shared memory; unused slots hold (-inf, max index).
UNIREC_DEVICE void sort_and_emit(float* s_val, long long* s_idx, int k, int k_eff, float scale, long long index_base,
                                 float* out_scores, long long* out_idx) {
    const int tid = threadIdx.x;
    for (int size = 2; size <= SC_KMAX; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (tid < SC_KMAX) {
                const int partner = tid ^ stride;
                if (partner > tid) {
                    const bool up = (tid & size) == 0;   // "up" block: best first
                    const float a = s_val[tid], b = s_val[partner];
                    const long long ia = s_idx[tid], ib = s_idx[partner];
                    const bool a_first = (a > b) || (a == b && ia < ib);
                    if (a_first != up) {
                        s_val[tid] = b; s_val[partner] = a;
                        s_idx[tid] = ib; s_idx[partner] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid < k) {
        const bool ok = tid < k_eff;
        out_scores[tid] = ok ? s_val[tid] * scale : -INFINITY;
        out_idx[tid] = ok ? s_idx[tid] + index_base : -1;
    }
}

// ---------------------------------------------------------------------------------------------
// row_kth_kernel: one CTA per user row of a dense score matrix (<= 512*VPT columns, held in registers as
// order-preserving integers).  Finds by bisection the largest t with count(score >= t) >= k and writes it
// to tau_out; with EMIT it also gathers, sorts and writes the row's top-k (dense ranking of small pools).
// ---------------------------------------------------------------------------------------------
struct RowKthParams {
    const float* dense; long long ldd; int ncols; int k;
    float* tau_out;                 // [rows] or nullptr
    const float* row_scale; long long index_base;
    float* out_scores; long long* out_idx;
};

template <int VPT, bool EMIT>
__global__ void __launch_bounds__(SC_ROW_THREADS)
row_kth_kernel(const RowKthParams p) {
    constexpr int NW = SC_ROW_THREADS / 32;
    __shared__ int s_red[NW];
    __shared__ uint32_t s_redu[NW];
    __shared__ int s_ngt, s_neq;
    __shared__ float s_val[SC_KMAX];
    __shared__ long long s_idx[SC_KMAX];
    const int row = blockIdx.x;
    const int tid = threadIdx.x;
    const float* src = p.dense + static_cast<long long>(row) * p.ldd;

    uint32_t ord[VPT];
    int nvalid = 0;
    uint32_t omax = 0;
#pragma unroll
    for (int i = 0; i < VPT / 4; ++i) {
        const int c = (i * SC_ROW_THREADS + tid) * 4;
        float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (c + 4 <= p.ncols) x = __ldg(reinterpret_cast<const float4*>(src + c));
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t o = f2ord(xs[j]);
            ord[4 * i + j] = o;
            nvalid += (o > ORD_NEG_INF) ? 1 : 0;
            omax = max(omax, o);
        }
    }
    nvalid = block_sum_int<NW>(nvalid, s_red);
    omax = block_max_u32<NW>(omax, s_redu);
    const int k_eff = min(p.k, nvalid);
    if (EMIT && tid < SC_KMAX) { s_val[tid] = -INFINITY; s_idx[tid] = 0x7fffffffffffffffLL; }
    if (tid == 0) { s_ngt = 0; s_neq = 0; }

    uint32_t lo = ORD_NEG_INF + 1u;   // count(>= lo) = nvalid >= k_eff
    if (k_eff > 0) {
        uint32_t hi = omax + 1u;      // count(>= hi) = 0
        while (hi - lo > 1u) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            int c = 0;
#pragma unroll
            for (int i = 0; i < VPT; ++i) c += (ord[i] >= mid) ? 1 : 0;
            c = block_sum_int<NW>(c, s_red);
            if (c >= k_eff) lo = mid; else hi = mid;
            if (c == k_eff) break;    // everything >= mid is exactly the top k_eff
        }
    }
    if (p.tau_out != nullptr && tid == 0) p.tau_out[row] = (k_eff == p.k) ? ord2f(lo) : -INFINITY;

    if constexpr (EMIT) {
        __syncthreads();
        if (k_eff > 0) {
#pragma unroll
            for (int i = 0; i < VPT; ++i) {
                if (ord[i] > lo) {
                    const int slot = atomicAdd(&s_ngt, 1);
                    if (slot < SC_KMAX) {
                        s_val[slot] = ord2f(ord[i]);
                        s_idx[slot] = ((i >> 2) * SC_ROW_THREADS + tid) * 4 + (i & 3);
                    }
                }
            }
            __syncthreads();
            const int ngt = min(s_ngt, k_eff);
#pragma unroll
            for (int i = 0; i < VPT; ++i) {
                if (ord[i] == lo) {
                    const int slot = ngt + atomicAdd(&s_neq, 1);
                    if (slot < k_eff) {
                        s_val[slot] = ord2f(ord[i]);
                        s_idx[slot] = ((i >> 2) * SC_ROW_THREADS + tid) * 4 + (i & 3);
                    }
                }
            }
        }
        __syncthreads();
        sort_and_emit(s_val, s_idx, p.k, k_eff, p.row_scale ? p.row_scale[row] : 1.0f, p.index_base,
                      p.out_scores + static_cast<size_t>(row) * p.k, p.out_idx + static_cast<size_t>(row) * p.k);
    }
}

// ---------------------------------------------------------------------------------------------
// Exact top-k selection over a row's candidate lists.
//   PACKED = true : lists are uint2 (score bits, int32 index) [rows][L][slots], counts [rows][L]
//   PACKED = false: scores fp32 / idx int64 [L][rows][slots] (all-gathered per-GPU lists), count = slots
// The valid entries are first gathered into shared memory (SC_SEL_CAP of them; (score, payload) pairs); if a
// row has more (adversarial pools only) the bisection runs over the lists in global memory instead.
// ---------------------------------------------------------------------------------------------
struct SelectParams {
    const uint2* packed; const int* counts;
    const float* scores; const long long* idx;
    int rows, L, slots, k;
    const float* row_scale;      // user_inv or nullptr
    long long index_base;
    float* out_scores; long long* out_idx;
};

template <bool PACKED>
__global__ void __launch_bounds__(256)
topk_select_kernel(const SelectParams p) {
    extern __shared__ __align__(16) uint8_t sel_smem[];
    uint2* s_ent = reinterpret_cast<uint2*>(sel_smem);           // [SC_SEL_CAP] (score bits, payload)
    __shared__ int s_red[8];
    __shared__ uint32_t s_redu[8];
    __shared__ int s_total, s_ngt, s_neq;
    __shared__ float s_val[SC_KMAX];
    __shared__ long long s_idx[SC_KMAX];
    const int row = blockIdx.x;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int total_slots = p.L * p.slots;

    auto score_at = [&](int e) -> float {
        if constexpr (PACKED) return __uint_as_float(p.packed[static_cast<size_t>(row) * total_slots + e].x);
        else return p.scores[(static_cast<size_t>(e / p.slots) * p.rows + row) * p.slots + e % p.slots];
    };
    auto idx_at = [&](int e) -> long long {
        if constexpr (PACKED) return static_cast<long long>(p.packed[static_cast<size_t>(row) * total_slots + e].y);
        else return p.idx[(static_cast<size_t>(e / p.slots) * p.rows + row) * p.slots + e % p.slots];
    };
    auto list_count = [&](int l) -> int {
        if constexpr (PACKED) return min(p.counts[static_cast<size_t>(row) * p.L + l], p.slots);
        else return p.slots;
    };

    if (tid == 0) { s_total = 0; s_ngt = 0; s_neq = 0; }
    if (tid < SC_KMAX) { s_val[tid] = -INFINITY; s_idx[tid] = 0x7fffffffffffffffLL; }
    __syncthreads();
    // ---- gather valid entries: one warp per list
    for (int l = warp; l < p.L; l += 8) {
        const int cnt = list_count(l);
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int j = j0 + lane;
            const bool ok = j < cnt;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            int basepos = 0;
            if (lane == 0) basepos = atomicAdd(&s_total, __popc(m));
            basepos = __shfl_sync(0xffffffffu, basepos, 0);
            if (ok) {
                const int pos = basepos + __popc(m & ((1u << lane) - 1u));
                if (pos < SC_SEL_CAP) {
                    const int e = l * p.slots + j;
                    s_ent[pos] = make_uint2(f2ord(score_at(e)), static_cast<uint32_t>(e));
                }
            }
        }
    }
    __syncthreads();
    const int nvalid = s_total;
    const bool in_smem = nvalid <= SC_SEL_CAP;
    const int k_eff = min(p.k, nvalid);

    // entry accessors for the selected storage: iterate i over [0, n_iter) and test ok(i)
    const int n_iter = in_smem ? nvalid : total_slots;
    auto ent_ok = [&](int i) -> bool {
        if (in_smem) return true;
        return (i % p.slots) < list_count(i / p.slots);
    };
    auto ent_ord = [&](int i) -> uint32_t { return in_smem ? s_ent[i].x : f2ord(score_at(i)); };
    auto ent_e = [&](int i) -> int { return in_smem ? static_cast<int>(s_ent[i].y) : i; };

    if (k_eff > 0) {
        uint32_t omax = 0, omin = 0xffffffffu;
        for (int i = tid; i < n_iter; i += 256) {
            if (ent_ok(i)) {
                const uint32_t o = ent_ord(i);
                omax = max(omax, o);
                omin = min(omin, o);
            }
        }
        omax = block_max_u32<8>(omax, s_redu);
        omin = ~block_max_u32<8>(~omin, s_redu);
        // largest ordinal t with count(ord >= t) >= k_eff
        uint32_t lo = omin;                       // count(>= omin) = nvalid >= k_eff
        uint32_t hi = omax + 1u;                  // count(>= hi) = 0   (omax < 0xffffffff for finite scores)
        while (hi - lo > 1u) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            int c = 0;
            for (int i = tid; i < n_iter; i += 256)
                if (ent_ok(i) && ent_ord(i) >= mid) ++c;
            c = block_sum_int<8>(c, s_red);
            if (c >= k_eff) lo = mid; else hi = mid;
            if (c == k_eff) break;                // everything >= mid is exactly the top k_eff
        }
        // gather: entries > lo always; entries == lo until k_eff is reached
        for (int i = tid; i < n_iter; i += 256) {
            if (!ent_ok(i)) continue;
            const uint32_t o = ent_ord(i);
            if (o > lo) {
                const int slot = atomicAdd(&s_ngt, 1);
                if (slot < SC_KMAX) { s_val[slot] = ord2f(o); s_idx[slot] = idx_at(ent_e(i)); }
            }
        }
        __syncthreads();
        const int ngt = min(s_ngt, k_eff);
        for (int i = tid; i < n_iter; i += 256) {
            if (!ent_ok(i)) continue;
            const uint32_t o = ent_ord(i);
            if (o == lo) {
                const int slot = ngt + atomicAdd(&s_neq, 1);
                if (slot < k_eff) { s_val[slot] = ord2f(o); s_idx[slot] = idx_at(ent_e(i)); }
            }
        }
    }
    __syncthreads();
    sort_and_emit(s_val, s_idx, p.k, k_eff, p.row_scale ? p.row_scale[row] : 1.0f, p.index_base,
                  p.out_scores + static_cast<size_t>(row) * p.k, p.out_idx + static_cast<size_t>(row) * p.k);
}

// ---------------------------------------------------------------------------------------------
// Host
// ---------------------------------------------------------------------------------------------
struct ScorePlan {
    int pair;                  // 1: CTA-pair kernel (256-user blocks, clusters of two CTAs); 0: single-CTA kernel (128)
    int num_m_blocks, n_tiles, R, tiles_per_range, grid;
    int dense_mode;            // 1: small pool, score everything densely and rank with row_kth_kernel
    int n_dense_tiles, tile_stride;
    size_t scratch_bytes, partial_bytes, cnt_bytes, dense_bytes, tau_bytes;
    size_t total() const { return scratch_bytes + partial_bytes + cnt_bytes + dense_bytes + tau_bytes; }
};

static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static ScorePlan make_plan(long long B, long long N) {
    ScorePlan pl;
    const int sms = num_sms() > 0 ? num_sms() : 148;
    // more than one 128-user block: pairs of CTAs share every candidate tile (score_pair_kernel); a single block (the
    // HBM-bound small-batch regime) keeps the single-CTA kernel, whose 148 CTAs each stream their own candidate range
    pl.pair = B > PIPE_BLOCK_M ? 1 : 0;
    const int block_m = pl.pair ? SP_TILE_M : PIPE_BLOCK_M;
    const int units = pl.pair ? sms / 2 : sms;          // clusters or CTAs that run concurrently
    const int ctas_per_unit = pl.pair ? 2 : 1;
    pl.num_m_blocks = static_cast<int>((B + block_m - 1) / block_m);
    pl.n_tiles = static_cast<int>((N + SC_BLOCK_N - 1) / SC_BLOCK_N);
    const size_t b_pad = static_cast<size_t>(pl.num_m_blocks) * block_m;
    pl.dense_mode = N <= SC_DENSE_MAX_N ? 1 : 0;
    if (pl.dense_mode) {
        pl.n_dense_tiles = pl.n_tiles;
        pl.tile_stride = 1;
        pl.R = 0; pl.tiles_per_range = 0;
        const long long items = static_cast<long long>(pl.n_dense_tiles) * pl.num_m_blocks;
        pl.grid = static_cast<int>(items < units ? items : units) * ctas_per_unit;
        pl.scratch_bytes = pl.partial_bytes = pl.cnt_bytes = pl.tau_bytes = 0;
        pl.dense_bytes = align256(b_pad * static_cast<size_t>(pl.n_dense_tiles) * SC_BLOCK_N * sizeof(float));
        return pl;
    }
    // sampled start threshold: every tile_stride-th candidate tile, 1/32 of the pool but at least 16 tiles (4096 candidates:
    // ~k * N / 4096 scores per user pass the filter) and at most 128 (what row_kth_kernel holds in registers)
    // Pools of a few 100 k rows (the per-rank shards of the 4- and 8-GPU runs) get a denser sample, 1/8 of the pool up to 64
    // tiles: the survivors of a user (~k * N / N_sample) are the same ~3000 whatever the pool size, so on a small pool they
    // are several per 256-candidate tile and the append path of the filter epilogue, not the MMAs, sets the pace
    // (profiles/r02_z_launches_score_topk.csv: 943 TFLOP/s at 4096 x 125 k against 1415 TFLOP/s at 4096 x 1 M)
    int nd = pl.n_tiles / 32;
    const int nd_small = pl.n_tiles / 8 < 64 ? pl.n_tiles / 8 : 64;
    nd = nd > nd_small ? nd : nd_small;
    nd = nd < 16 ? 16 : (nd > SC_DENSE_MAX_N / SC_BLOCK_N ? SC_DENSE_MAX_N / SC_BLOCK_N : nd);
    pl.n_dense_tiles = nd;
    pl.tile_stride = pl.n_tiles / nd;
    pl.dense_bytes = align256(b_pad * static_cast<size_t>(nd) * SC_BLOCK_N * sizeof(float));
    pl.tau_bytes = align256(b_pad * sizeof(float));
    // number of candidate ranges: enough work items to balance the SMs, at least ~8 tiles per range
    const int r_max = pl.n_tiles >= 16 ? pl.n_tiles / 8 : 1;
    int best_r = r_max;   // fallback: cannot fill the machine, use the most ranges allowed
    double best_waste = 1e30;
    for (int r = 1; r <= r_max; ++r) {
        const long long items = static_cast<long long>(r) * pl.num_m_blocks;
        if (items < units) continue;
        if (items > 16LL * units) break;
        const long long waves = (items + units - 1) / units;
        const double waste = static_cast<double>(waves * units) / static_cast<double>(items);
        if (waste < best_waste - 1e-9) { best_waste = waste; best_r = r; }   // ties: keep the smaller r
    }
    pl.tiles_per_range = (pl.n_tiles + best_r - 1) / best_r;
    pl.R = (pl.n_tiles + pl.tiles_per_range - 1) / pl.tiles_per_range;
    const long long items = static_cast<long long>(pl.R) * pl.num_m_blocks;
    pl.grid = static_cast<int>(items < units ? items : units) * ctas_per_unit;
    pl.scratch_bytes = align256(static_cast<size_t>(pl.grid) * 8 * SC_CAP * 32 * sizeof(uint2));
    pl.partial_bytes = align256(b_pad * 2 * pl.R * SC_KMAX * sizeof(uint2));
    pl.cnt_bytes = align256(b_pad * 2 * pl.R * sizeof(int));
    return pl;
}

template <class K>
static int set_smem_attr(K kern, int bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_last_error("%s: cudaFuncSetAttribute(%d): %s", what, bytes, cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

static int launch_row_kth(const RowKthParams& rp, int rows, bool emit, cudaStream_t stream) {
#define UNIREC_ROWK(VPT)                                                                             \
    do {                                                                                             \
        if (emit) row_kth_kernel<VPT, true><<<rows, SC_ROW_THREADS, 0, stream>>>(rp);                \
        else row_kth_kernel<VPT, false><<<rows, SC_ROW_THREADS, 0, stream>>>(rp);                    \
    } while (0)
    if (rp.ncols <= SC_ROW_THREADS * 4) UNIREC_ROWK(4);
    else if (rp.ncols <= SC_ROW_THREADS * 16) UNIREC_ROWK(16);
    else if (rp.ncols <= SC_ROW_THREADS * 32) UNIREC_ROWK(32);
    else UNIREC_ROWK(64);
#undef UNIREC_ROWK
    return UNIREC_OK;
}

// One pass over <= SC_MAX_USERS_PER_PASS users.
static int score_topk_pass(const void* users, int64_t ldu, const float* user_inv, const void* cands, int64_t ldc,
                           const float* cand_inv, int64_t B, int64_t N, int64_t D, int64_t k, int64_t index_base,
                           float* out_scores, int64_t* out_idx, uint8_t* ws, cudaStream_t stream) {
    const ScorePlan pl = make_plan(B, N);
    ScoreParams p{};
    p.B = (int)B; p.N = (int)N; p.D = (int)D; p.k = (int)k;
    p.cand_inv = cand_inv;
    p.num_m_blocks = pl.num_m_blocks; p.n_tiles = pl.n_tiles; p.R = pl.R; p.tiles_per_range = pl.tiles_per_range;
    p.partial_cnt = reinterpret_cast<int*>(ws);
    p.partial = reinterpret_cast<uint2*>(ws + pl.cnt_bytes);
    p.scratch = reinterpret_cast<uint2*>(ws + pl.cnt_bytes + pl.partial_bytes);
    p.dense = reinterpret_cast<float*>(ws + pl.cnt_bytes + pl.partial_bytes + pl.scratch_bytes);
    float* tau = reinterpret_cast<float*>(ws + pl.cnt_bytes + pl.partial_bytes + pl.scratch_bytes + pl.dense_bytes);
    p.ldd = static_cast<long long>(pl.n_dense_tiles) * SC_BLOCK_N;
    p.n_dense_tiles = pl.n_dense_tiles; p.tile_stride = pl.tile_stride;
    p.row_tau = nullptr;

    CUtensorMap tu, tc;
    int rc = make_tmap_bf16_2d(&tu, users, B, D, ldu, PIPE_BLOCK_M);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tc, cands, N, D, ldc, pl.pair ? 128 : SC_BLOCK_N);     // a CTA of a pair loads 128 candidate rows
    if (rc != UNIREC_OK) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        if ((rc = set_smem_attr(score_tile_kernel<MODE_FILTER>, ScPipe::SMEM_BYTES, "score_topk")) != UNIREC_OK) return rc;
        if ((rc = set_smem_attr(score_tile_kernel<MODE_DENSE>, ScPipe::SMEM_BYTES, "score_topk")) != UNIREC_OK) return rc;
        if ((rc = set_smem_attr(score_pair_kernel<MODE_FILTER>, SP_SMEM_BYTES, "score_topk")) != UNIREC_OK) return rc;
        if ((rc = set_smem_attr(score_pair_kernel<MODE_DENSE>, SP_SMEM_BYTES, "score_topk")) != UNIREC_OK) return rc;
        if ((rc = set_smem_attr(topk_select_kernel<true>, SC_SEL_CAP * 8, "score_topk")) != UNIREC_OK) return rc;
        if ((rc = set_smem_attr(topk_select_kernel<false>, SC_SEL_CAP * 8, "score_topk")) != UNIREC_OK) return rc;
        attr_set = true;
    }
    // ---- sampled (or, for small pools, complete) dense scores
    {
        const long long items = static_cast<long long>(pl.n_dense_tiles) * pl.num_m_blocks;
        const int sms = num_sms() > 0 ? num_sms() : 148;
        if (pl.pair) {
            const int clusters = static_cast<int>(items < sms / 2 ? items : sms / 2);
            score_pair_kernel<MODE_DENSE><<<2 * clusters, SC_THREADS, SP_SMEM_BYTES, stream>>>(tu, tc, p);
        } else {
            const int grid = static_cast<int>(items < sms ? items : sms);
            score_tile_kernel<MODE_DENSE><<<grid, SC_THREADS, ScPipe::SMEM_BYTES, stream>>>(tu, tc, p);
        }
    }
    RowKthParams rp{};
    rp.dense = p.dense; rp.ldd = p.ldd; rp.ncols = pl.n_dense_tiles * SC_BLOCK_N; rp.k = (int)k;
    rp.row_scale = user_inv; rp.index_base = index_base;
    rp.out_scores = out_scores; rp.out_idx = reinterpret_cast<long long*>(out_idx);
    if (pl.dense_mode) {
        rp.tau_out = nullptr;
        launch_row_kth(rp, (int)B, true, stream);
        g_launch_count.fetch_add(2, std::memory_order_relaxed);
    } else {
        rp.tau_out = tau;
        launch_row_kth(rp, (int)B, false, stream);
        p.row_tau = tau;
        if (pl.pair) score_pair_kernel<MODE_FILTER><<<pl.grid, SC_THREADS, SP_SMEM_BYTES, stream>>>(tu, tc, p);
        else score_tile_kernel<MODE_FILTER><<<pl.grid, SC_THREADS, ScPipe::SMEM_BYTES, stream>>>(tu, tc, p);
        SelectParams sp{};
        sp.packed = p.partial; sp.counts = p.partial_cnt; sp.scores = nullptr; sp.idx = nullptr;
        sp.rows = (int)B; sp.L = 2 * pl.R; sp.slots = SC_KMAX; sp.k = (int)k;
        sp.row_scale = user_inv; sp.index_base = index_base;
        sp.out_scores = out_scores; sp.out_idx = reinterpret_cast<long long*>(out_idx);
        topk_select_kernel<true><<<static_cast<unsigned>(B), 256, SC_SEL_CAP * 8, stream>>>(sp);
        g_launch_count.fetch_add(4, std::memory_order_relaxed);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("score_topk launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

}  // namespace unirec

using namespace unirec;

extern "C" {

int64_t unirec_score_topk_workspace_bytes(int64_t B, int64_t N, int64_t k) {
    if (B <= 0 || N <= 0 || k <= 0 || k > SC_KMAX) {
        set_last_error("score_topk: need B > 0, N > 0, 0 < k <= %d", SC_KMAX);
        return -1;
    }
    // the passes of a call share one workspace: full passes of SC_MAX_USERS_PER_PASS users and a shorter last one
    size_t need = make_plan(B < SC_MAX_USERS_PER_PASS ? B : SC_MAX_USERS_PER_PASS, N).total();
    if (B > SC_MAX_USERS_PER_PASS && B % SC_MAX_USERS_PER_PASS != 0) {
        const size_t last = make_plan(B % SC_MAX_USERS_PER_PASS, N).total();
        need = last > need ? last : need;
    }
    return static_cast<int64_t>(need + 1024);
}

int unirec_score_topk(const void* users, int64_t ldu, const float* user_inv, const void* cands, int64_t ldc,
                      const float* cand_inv, int64_t B, int64_t N, int64_t D, int64_t k, int64_t index_base,
                      float* out_scores, int64_t* out_idx, void* workspace, int64_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (users == nullptr || cands == nullptr || user_inv == nullptr || cand_inv == nullptr || out_scores == nullptr ||
        out_idx == nullptr || workspace == nullptr || B <= 0 || N <= 0 || k <= 0 || k > SC_KMAX || D % PIPE_BLOCK_K != 0 ||
        ldu % 8 != 0 || ldc % 8 != 0 || N > 2147483647LL - SC_BLOCK_N) {
        set_last_error("score_topk: bad arguments (B=%lld N=%lld D=%lld k=%lld; need D%%64==0, k<=%d)", (long long)B,
                       (long long)N, (long long)D, (long long)k, SC_KMAX);
        return UNIREC_ERR_BAD_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(cand_inv) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
        set_last_error("score_topk: cand_inv must be 16-byte and workspace 256-byte aligned");
        return UNIREC_ERR_BAD_ARG;
    }
    const int64_t need = unirec_score_topk_workspace_bytes(B, N, k) - 1024;
    if (workspace_bytes < need) {
        set_last_error("score_topk: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
        return UNIREC_ERR_BAD_ARG;
    }
    // users are processed in passes of <= 4096 (bounds the dense sample matrix and the list scratch)
    for (int64_t b0 = 0; b0 < B; b0 += SC_MAX_USERS_PER_PASS) {
        const int64_t nb = (B - b0 < SC_MAX_USERS_PER_PASS) ? (B - b0) : SC_MAX_USERS_PER_PASS;
        const int rc = score_topk_pass(reinterpret_cast<const __nv_bfloat16*>(users) + b0 * ldu, ldu, user_inv + b0, cands,
                                       ldc, cand_inv, nb, N, D, k, index_base, out_scores + b0 * k, out_idx + b0 * k,
                                       reinterpret_cast<uint8_t*>(workspace), stream);
        if (rc != UNIREC_OK) return rc;
    }
    return UNIREC_OK;
}

int unirec_topk_merge(const float* in_scores, const int64_t* in_idx, int64_t G, int64_t B, int64_t k,
                      float* out_scores, int64_t* out_idx, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (in_scores == nullptr || in_idx == nullptr || out_scores == nullptr || out_idx == nullptr || G <= 0 || B <= 0 ||
        k <= 0 || k > SC_KMAX) {
        set_last_error("topk_merge: bad arguments (G=%lld B=%lld k=%lld)", (long long)G, (long long)B, (long long)k);
        return UNIREC_ERR_BAD_ARG;
    }
    static bool attr_set = false;
    if (!attr_set) {
        const int rc = set_smem_attr(topk_select_kernel<false>, SC_SEL_CAP * 8, "topk_merge");
        if (rc != UNIREC_OK) return rc;
        attr_set = true;
    }
    SelectParams sp{};
    sp.packed = nullptr; sp.counts = nullptr; sp.scores = in_scores; sp.idx = reinterpret_cast<const long long*>(in_idx);
    sp.rows = (int)B; sp.L = (int)G; sp.slots = (int)k; sp.k = (int)k;
    sp.row_scale = nullptr; sp.index_base = 0;
    sp.out_scores = out_scores; sp.out_idx = reinterpret_cast<long long*>(out_idx);
    topk_select_kernel<false><<<static_cast<unsigned>(B), 256, SC_SEL_CAP * 8, stream>>>(sp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("topk_merge launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return UNIREC_OK;
}

}  // extern "C"
