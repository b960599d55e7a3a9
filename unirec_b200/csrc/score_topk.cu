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
// Kernel 1 (score_filter_kernel): the tcgen05/TMEM/TMA pipeline of umma_pipe.cuh computes 128 x 256
// score tiles.  The epilogue thread that owns (row, column-half) keeps a running threshold tau in a
// register and appends only candidates with score > tau to its private list in global scratch memory
// (interleaved [slot][lane] so that a warp scanning its 32 lists in lock-step reads coalesced lines).
// When a list is nearly full the warp compacts its lists: each lane bisects (on the order-preserving
// integer image of the float) for a threshold that keeps between k and KMAX entries, rewrites its
// list in place and raises tau.  After warm-up almost nothing passes the filter, so the epilogue costs
// one multiply and one max per score.  The candidate range is split into R contiguous ranges; a work
// item = (128-user block, range); work items are dealt round-robin to persistent CTAs so that the CTAs
// that stream the same candidate range run concurrently (candidates come from HBM once, then L2).
//
// Kernel 2 (topk_select_kernel): per user, exact selection of the k largest among the <= 2*R*KMAX
// surviving entries (bisection for the k-th value, then a 128-wide bitonic sort), scaled by user_inv.
// The same kernel merges per-GPU top-k lists after the NCCL all-gather (unirec_topk_merge).
#include "common.cuh"
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
using ScPipe = UmmaPipe<SC_BLOCK_N, 4>;

UNIREC_DEVICE uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct ScoreParams {
    int B, N, D, k;
    const float* cand_inv;
    int num_m_blocks, n_tiles, R, tiles_per_range;
    uint2* scratch;      // [grid][8][SC_CAP][32]
    uint2* partial;      // [B_pad][2R][SC_KMAX]   (score bits, candidate index)
    int* partial_cnt;    // [B_pad][2R]
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
        const uint32_t u = (lo & 0x80000000u) ? (lo & 0x7fffffffu) : ~lo;
        tau = __uint_as_float(u);
    }
}

__global__ void __launch_bounds__(SC_THREADS, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap tmap_users, const __grid_constant__ CUtensorMap tmap_cands,
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
    const int num_items = p.num_m_blocks * p.R;

    if (warp_idx == 0) {
        RingState rs;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
            const int m_blk = w % p.num_m_blocks, r = w / p.num_m_blocks;
            const int t0 = r * p.tiles_per_range, t1 = min(t0 + p.tiles_per_range, p.n_tiles);
            for (int t = t0; t < t1; ++t)
                pipe_produce_tile(pipe, rs, &tmap_users, &tmap_cands, m_blk * PIPE_BLOCK_M, t * SC_BLOCK_N, num_kb, lane);
        }
    } else if (warp_idx == 1) {
        RingState rs;
        uint32_t iter = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
            const int r = w / p.num_m_blocks;
            const int t0 = r * p.tiles_per_range, t1 = min(t0 + p.tiles_per_range, p.n_tiles);
            for (int t = t0; t < t1; ++t, ++iter) pipe_mma_tile<SC_BLOCK_N>(pipe, rs, iter, num_kb, lane);
        }
    } else if (warp_idx >= 4) {
        const int q = warp_idx & 3;
        const int half = (warp_idx - 4) >> 2;
        uint2* base = p.scratch + (static_cast<size_t>(blockIdx.x) * 8 + (warp_idx - 4)) * SC_CAP * 32;
        const int L = 2 * p.R;
        uint32_t iter = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
            const int m_blk = w % p.num_m_blocks, r = w / p.num_m_blocks;
            const int t0 = r * p.tiles_per_range, t1 = min(t0 + p.tiles_per_range, p.n_tiles);
            const int row = m_blk * PIPE_BLOCK_M + q * 32 + lane;
            const bool row_ok = row < p.B;
            int cnt = 0;
            float tau = -INFINITY;
            for (int t = t0; t < t1; ++t, ++iter) {
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
                }
                // room for the next tile's (at most 128) appends?
                if (__any_sync(0xffffffffu, cnt > SC_CAP - 128)) compact_lists(base, lane, cnt, tau, p.k);
            }
            if (__any_sync(0xffffffffu, cnt > SC_KMAX)) compact_lists(base, lane, cnt, tau, p.k);
            if (row_ok) {
                const size_t list = static_cast<size_t>(row) * L + (r * 2 + half);
                p.partial_cnt[list] = cnt;
                uint2* dst = p.partial + list * SC_KMAX;
                for (int j = 0; j < cnt; ++j) dst[j] = base[j * 32 + lane];
            }
        }
    }
    pipe.teardown(warp_idx);
}

// ---------------------------------------------------------------------------------------------
// Exact top-k selection over a row's candidate lists.
//   PACKED = true : lists are uint2 (score bits, int32 index) [rows][L][slots], counts [rows][L]
//   PACKED = false: scores fp32 / idx int64 [L][rows][slots] (all-gathered per-GPU lists), count = slots
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
    __shared__ int s_cnt[8];
    __shared__ int s_ngt, s_neq;
    __shared__ float s_val[SC_KMAX];
    __shared__ long long s_idx[SC_KMAX];
    const int row = blockIdx.x;
    const int tid = threadIdx.x;
    const int total_slots = p.L * p.slots;

    auto valid = [&](int e) -> bool {
        if constexpr (PACKED) return (e % p.slots) < p.counts[static_cast<size_t>(row) * p.L + e / p.slots];
        else return true;
    };
    auto score_at = [&](int e) -> float {
        if constexpr (PACKED) return __uint_as_float(p.packed[static_cast<size_t>(row) * total_slots + e].x);
        else return p.scores[(static_cast<size_t>(e / p.slots) * p.rows + row) * p.slots + e % p.slots];
    };
    auto idx_at = [&](int e) -> long long {
        if constexpr (PACKED) return static_cast<long long>(p.packed[static_cast<size_t>(row) * total_slots + e].y);
        else return p.idx[(static_cast<size_t>(e / p.slots) * p.rows + row) * p.slots + e % p.slots];
    };
    auto block_sum = [&](int v) -> int {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((tid & 31) == 0) s_cnt[tid >> 5] = v;
        __syncthreads();
        int t = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s_cnt[i];
        return t;
    };

    // number of valid entries and the ordinal range
    int nvalid = 0;
    uint32_t omax = 0, omin = 0xffffffffu;
    for (int e = tid; e < total_slots; e += blockDim.x) {
        if (valid(e)) {
            ++nvalid;
            const uint32_t o = f2ord(score_at(e));
            omax = max(omax, o);
            omin = min(omin, o);
        }
    }
    nvalid = block_sum(nvalid);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        omax = max(omax, __shfl_xor_sync(0xffffffffu, omax, o));
        omin = min(omin, __shfl_xor_sync(0xffffffffu, omin, o));
    }
    __shared__ uint32_t s_omax[8], s_omin[8];
    if ((tid & 31) == 0) { s_omax[tid >> 5] = omax; s_omin[tid >> 5] = omin; }
    __syncthreads();
    for (int i = 0; i < 8; ++i) { omax = max(omax, s_omax[i]); omin = min(omin, s_omin[i]); }

    const int k_eff = min(p.k, nvalid);
    if (tid < SC_KMAX) { s_val[tid] = -INFINITY; s_idx[tid] = 0x7fffffffffffffffLL; }
    if (tid == 0) { s_ngt = 0; s_neq = 0; }
    __syncthreads();

    if (k_eff > 0) {
        // largest ordinal t with count(ord >= t) >= k_eff
        uint32_t lo = omin;                       // count(>= omin) = nvalid >= k_eff
        uint32_t hi = omax + 1u;                  // count(>= hi) = 0   (omax < 0xffffffff for finite scores)
        while (hi - lo > 1u) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            int c = 0;
            for (int e = tid; e < total_slots; e += blockDim.x)
                if (valid(e) && f2ord(score_at(e)) >= mid) ++c;
            c = block_sum(c);
            if (c >= k_eff) lo = mid; else hi = mid;
            if (c == k_eff) break;                // everything >= mid is exactly the top k_eff
        }
        // gather: entries > lo always; entries == lo until k_eff is reached
        for (int e = tid; e < total_slots; e += blockDim.x) {
            if (!valid(e)) continue;
            const float sc = score_at(e);
            const uint32_t o = f2ord(sc);
            if (o > lo) {
                const int slot = atomicAdd(&s_ngt, 1);
                if (slot < SC_KMAX) { s_val[slot] = sc; s_idx[slot] = idx_at(e); }
            }
        }
        __syncthreads();
        const int ngt = min(s_ngt, k_eff);
        for (int e = tid; e < total_slots; e += blockDim.x) {
            if (!valid(e)) continue;
            const float sc = score_at(e);
            if (f2ord(sc) == lo) {
                const int slot = ngt + atomicAdd(&s_neq, 1);
                if (slot < k_eff) { s_val[slot] = sc; s_idx[slot] = idx_at(e); }
            }
        }
        __syncthreads();
    }

    // bitonic sort of SC_KMAX (score desc, index asc); unused slots hold (-inf, max index)
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
    if (tid < p.k) {
        const float scale = p.row_scale ? p.row_scale[row] : 1.0f;
        const bool ok = tid < k_eff;
        p.out_scores[static_cast<size_t>(row) * p.k + tid] = ok ? s_val[tid] * scale : -INFINITY;
        p.out_idx[static_cast<size_t>(row) * p.k + tid] = ok ? s_idx[tid] + p.index_base : -1;
    }
}

// ---------------------------------------------------------------------------------------------
// Host
// ---------------------------------------------------------------------------------------------
struct ScorePlan {
    int num_m_blocks, n_tiles, R, tiles_per_range, grid;
    size_t scratch_bytes, partial_bytes, cnt_bytes;
};

static ScorePlan make_plan(long long B, long long N) {
    ScorePlan pl;
    const int sms = num_sms() > 0 ? num_sms() : 148;
    pl.num_m_blocks = static_cast<int>((B + PIPE_BLOCK_M - 1) / PIPE_BLOCK_M);
    pl.n_tiles = static_cast<int>((N + SC_BLOCK_N - 1) / SC_BLOCK_N);
    // number of candidate ranges: enough work items to balance the SMs, at least ~8 tiles per range
    const int r_max = pl.n_tiles >= 16 ? pl.n_tiles / 8 : 1;
    int best_r = r_max;   // fallback: cannot fill the machine, use the most ranges allowed
    double best_waste = 1e30;
    for (int r = 1; r <= r_max; ++r) {
        const long long items = static_cast<long long>(r) * pl.num_m_blocks;
        if (items < sms) continue;
        if (items > 16LL * sms) break;
        const long long waves = (items + sms - 1) / sms;
        const double waste = static_cast<double>(waves * sms) / static_cast<double>(items);
        if (waste < best_waste - 1e-9) { best_waste = waste; best_r = r; }   // ties: keep the smaller r
    }
    pl.tiles_per_range = (pl.n_tiles + best_r - 1) / best_r;
    pl.R = (pl.n_tiles + pl.tiles_per_range - 1) / pl.tiles_per_range;
    const long long items = static_cast<long long>(pl.R) * pl.num_m_blocks;
    pl.grid = static_cast<int>(items < sms ? items : sms);
    const size_t b_pad = static_cast<size_t>(pl.num_m_blocks) * PIPE_BLOCK_M;
    pl.scratch_bytes = static_cast<size_t>(pl.grid) * 8 * SC_CAP * 32 * sizeof(uint2);
    pl.partial_bytes = b_pad * 2 * pl.R * SC_KMAX * sizeof(uint2);
    pl.cnt_bytes = ((b_pad * 2 * pl.R * sizeof(int)) + 255) & ~static_cast<size_t>(255);
    return pl;
}

}  // namespace unirec

using namespace unirec;

extern "C" {

int64_t unirec_score_topk_workspace_bytes(int64_t B, int64_t N, int64_t k) {
    if (B <= 0 || N <= 0 || k <= 0 || k > SC_KMAX) {
        set_last_error("score_topk: need B > 0, N > 0, 0 < k <= %d", SC_KMAX);
        return -1;
    }
    const ScorePlan pl = make_plan(B, N);
    return static_cast<int64_t>(pl.scratch_bytes + pl.partial_bytes + pl.cnt_bytes + 1024);
}

int unirec_score_topk(const void* users, int64_t ldu, const float* user_inv, const void* cands, int64_t ldc,
                      const float* cand_inv, int64_t B, int64_t N, int64_t D, int64_t k, int64_t index_base,
                      float* out_scores, int64_t* out_idx, void* workspace, int64_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (users == nullptr || cands == nullptr || user_inv == nullptr || cand_inv == nullptr || out_scores == nullptr ||
        out_idx == nullptr || workspace == nullptr || B <= 0 || N <= 0 || k <= 0 || k > SC_KMAX || D % PIPE_BLOCK_K != 0 ||
        ldu % 8 != 0 || ldc % 8 != 0 || N > 2147483647LL) {
        set_last_error("score_topk: bad arguments (B=%lld N=%lld D=%lld k=%lld; need D%%64==0, k<=%d)", (long long)B,
                       (long long)N, (long long)D, (long long)k, SC_KMAX);
        return UNIREC_ERR_BAD_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(cand_inv) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15)) {
        set_last_error("score_topk: cand_inv and workspace must be 16-byte aligned");
        return UNIREC_ERR_BAD_ARG;
    }
    const ScorePlan pl = make_plan(B, N);
    const size_t need = pl.scratch_bytes + pl.partial_bytes + pl.cnt_bytes;
    if (static_cast<size_t>(workspace_bytes) < need) {
        set_last_error("score_topk: workspace too small (%lld < %zu)", (long long)workspace_bytes, need);
        return UNIREC_ERR_BAD_ARG;
    }
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    ScoreParams p;
    p.B = (int)B; p.N = (int)N; p.D = (int)D; p.k = (int)k;
    p.cand_inv = cand_inv;
    p.num_m_blocks = pl.num_m_blocks; p.n_tiles = pl.n_tiles; p.R = pl.R; p.tiles_per_range = pl.tiles_per_range;
    p.partial_cnt = reinterpret_cast<int*>(ws);
    p.partial = reinterpret_cast<uint2*>(ws + pl.cnt_bytes);
    p.scratch = reinterpret_cast<uint2*>(ws + pl.cnt_bytes + pl.partial_bytes);

    CUtensorMap tu, tc;
    int rc = make_tmap_bf16_2d(&tu, users, B, D, ldu, PIPE_BLOCK_M);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tc, cands, N, D, ldc, SC_BLOCK_N);
    if (rc != UNIREC_OK) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(score_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             ScPipe::SMEM_BYTES);
        if (e != cudaSuccess) {
            set_last_error("score_topk: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return UNIREC_ERR_CUDA;
        }
        attr_set = true;
    }
    score_filter_kernel<<<pl.grid, SC_THREADS, ScPipe::SMEM_BYTES, stream>>>(tu, tc, p);
    SelectParams sp;
    sp.packed = p.partial; sp.counts = p.partial_cnt; sp.scores = nullptr; sp.idx = nullptr;
    sp.rows = (int)B; sp.L = 2 * pl.R; sp.slots = SC_KMAX; sp.k = (int)k;
    sp.row_scale = user_inv; sp.index_base = index_base;
    sp.out_scores = out_scores; sp.out_idx = reinterpret_cast<long long*>(out_idx);
    topk_select_kernel<true><<<static_cast<unsigned>(B), 256, 0, stream>>>(sp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("score_topk launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    g_launch_count.fetch_add(2, std::memory_order_relaxed);
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
    SelectParams sp;
    sp.packed = nullptr; sp.counts = nullptr; sp.scores = in_scores; sp.idx = reinterpret_cast<const long long*>(in_idx);
    sp.rows = (int)B; sp.L = (int)G; sp.slots = (int)k; sp.k = (int)k;
    sp.row_scale = nullptr; sp.index_base = 0;
    sp.out_scores = out_scores; sp.out_idx = reinterpret_cast<long long*>(out_idx);
    topk_select_kernel<false><<<static_cast<unsigned>(B), 256, 0, stream>>>(sp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("topk_merge launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return UNIREC_OK;
}

}  // extern "C"
