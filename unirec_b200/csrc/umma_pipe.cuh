// Shared TMA -> smem ring -> tcgen05.mma -> TMEM pipeline used by the projection GEMM
// (gemm_tcgen05.cu) and the fused scoring kernel (score_topk.cu).
//
// Tile: 128 (M) x BLOCK_N x 64 (K per stage), bf16 operands, both K-major, 128-byte swizzle.
// Barriers: full[s]/empty[s] for the STAGES-deep smem ring (TMA <-> MMA), tmem_full[a]/tmem_empty[a]
// for the two TMEM accumulator stages (MMA <-> epilogue).
#pragma once

#include "common.cuh"

namespace unirec {

constexpr int PIPE_BLOCK_M = 128;
constexpr int PIPE_BLOCK_K = 64;
constexpr int PIPE_UMMA_K = 16;
constexpr int PIPE_A_STAGE_BYTES = PIPE_BLOCK_M * PIPE_BLOCK_K * 2;

template <int BLOCK_N, int STAGES_>
struct UmmaPipe {
    static constexpr int STAGES = STAGES_;
    static constexpr int B_STAGE_BYTES = BLOCK_N * PIPE_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = PIPE_A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages; 256 or 512 (power of two)
    static constexpr int BARRIER_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + BARRIER_BYTES;
    static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N must be 128 or 256");
    static_assert((2 * STAGES + 4) * 8 + 4 <= BARRIER_BYTES, "barrier area too small");

    uint8_t* smem_a;
    uint8_t* smem_b;
    uint8_t* extra;       // first byte after the barrier area (kernel-specific use)
    uint64_t* full_bar;
    uint64_t* empty_bar;
    uint64_t* tmem_full_bar;
    uint64_t* tmem_empty_bar;
    uint32_t* tmem_ptr_smem;
    uint32_t tmem_base;

    // Carve shared memory, initialise barriers, allocate TMEM.  Must be called by all threads of the
    // CTA (contains __syncthreads).  Warp 1 initialises barriers, warp 2 allocates TMEM.
    UNIREC_DEVICE void setup(uint8_t* smem_raw, int warp_idx, int lane, uint32_t num_epilogue_threads) {
        uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
        smem_a = smem;
        smem_b = smem + STAGES * PIPE_A_STAGE_BYTES;
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
        full_bar = bars;
        empty_bar = bars + STAGES;
        tmem_full_bar = bars + 2 * STAGES;
        tmem_empty_bar = bars + 2 * STAGES + 2;
        tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
        extra = smem + STAGES * STAGE_BYTES + BARRIER_BYTES;
        if (warp_idx == 1 && lane == 0) {
            for (int i = 0; i < STAGES; ++i) {
                mbar_init(&full_bar[i], 1);
                mbar_init(&empty_bar[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&tmem_full_bar[i], 1);
                mbar_init(&tmem_empty_bar[i], num_epilogue_threads);
            }
            fence_mbar_init();
        }
        if (warp_idx == 2) {
            tmem_alloc(tmem_ptr_smem, TMEM_COLS);
            tmem_relinquish();
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem_base = *tmem_ptr_smem;
    }

    UNIREC_DEVICE void teardown(int warp_idx) {
        tc_fence_before();
        __syncthreads();
        if (warp_idx == 2) {
            tc_fence_after();
            tmem_dealloc(tmem_base, TMEM_COLS);
        }
    }
};

struct RingState {
    int stage = 0;
    uint32_t phase = 0;
};

// Producer warp: stream one output tile's K panels (A rows at m_coord, B rows at n_coord).
template <class Pipe>
UNIREC_DEVICE void pipe_produce_tile(Pipe& pipe, RingState& rs, const CUtensorMap* tmap_a, const CUtensorMap* tmap_b,
                                     int m_coord, int n_coord, int num_kb, int lane) {
    for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&pipe.empty_bar[rs.stage], rs.phase ^ 1);
        if (lane == 0) {
            mbar_arrive_expect_tx(&pipe.full_bar[rs.stage], Pipe::STAGE_BYTES);
            tma_load_2d(tmap_a, &pipe.full_bar[rs.stage], pipe.smem_a + rs.stage * PIPE_A_STAGE_BYTES,
                        kb * PIPE_BLOCK_K, m_coord);
            tma_load_2d(tmap_b, &pipe.full_bar[rs.stage], pipe.smem_b + rs.stage * Pipe::B_STAGE_BYTES,
                        kb * PIPE_BLOCK_K, n_coord);
        }
        __syncwarp();
        if (++rs.stage == Pipe::STAGES) { rs.stage = 0; rs.phase ^= 1; }
    }
}

// MMA warp: accumulate one output tile into TMEM accumulator stage `as` (iteration counter `iter`
// selects stage and barrier parity), then signal the epilogue.
template <int BLOCK_N, class Pipe>
UNIREC_DEVICE void pipe_mma_tile(Pipe& pipe, RingState& rs, uint32_t iter, int num_kb, int lane) {
    constexpr uint32_t idesc = umma_idesc_bf16(PIPE_BLOCK_M, BLOCK_N);
    const uint32_t as = iter & 1u;
    const uint32_t aphase = (iter >> 1) & 1u;
    mbar_wait(&pipe.tmem_empty_bar[as], aphase ^ 1);
    tc_fence_after();
    const uint32_t tmem_d = pipe.tmem_base + as * BLOCK_N;
    for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&pipe.full_bar[rs.stage], rs.phase);
        tc_fence_after();
        if (lane == 0) {
            const uint32_t a_addr = smem_u32(pipe.smem_a + rs.stage * PIPE_A_STAGE_BYTES);
            const uint32_t b_addr = smem_u32(pipe.smem_b + rs.stage * Pipe::B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < PIPE_BLOCK_K / PIPE_UMMA_K; ++k) {
                const uint64_t da = umma_smem_desc_sw128(a_addr + k * PIPE_UMMA_K * 2);
                const uint64_t db = umma_smem_desc_sw128(b_addr + k * PIPE_UMMA_K * 2);
                umma_bf16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&pipe.empty_bar[rs.stage]);                          // frees this smem slot
            if (kb == num_kb - 1) umma_commit(&pipe.tmem_full_bar[as]);      // accumulator complete
        }
        __syncwarp();
        if (++rs.stage == Pipe::STAGES) { rs.stage = 0; rs.phase ^= 1; }
    }
}

// Epilogue side: wait for accumulator stage of iteration `iter`; returns the TMEM column base.
template <int BLOCK_N, class Pipe>
UNIREC_DEVICE uint32_t pipe_epilogue_wait(Pipe& pipe, uint32_t iter) {
    const uint32_t as = iter & 1u;
    const uint32_t aphase = (iter >> 1) & 1u;
    mbar_wait(&pipe.tmem_full_bar[as], aphase);
    tc_fence_after();
    return pipe.tmem_base + as * BLOCK_N;
}

// Epilogue side: this thread has finished reading the accumulator stage of iteration `iter`.
template <class Pipe>
UNIREC_DEVICE void pipe_epilogue_release(Pipe& pipe, uint32_t iter) {
    tc_fence_before();
    mbar_arrive(&pipe.tmem_empty_bar[iter & 1u]);
}

// LayerNorm folded into the GEMMs around it (gemm_cg2.cu, unirec_linear_ln_bf16): device pointers, all optional
struct LnFold {
    const float* in_stats;    // [M, parts, 2] row (sum, sum of squares) partials of the tensor A was read from -> A is LayerNorm-ed on the fly
    const float* in_c;        // [N] column sums of the gamma-scaled weight (required with in_stats)
    const float* res_stats;   // [M, parts, 2] of the residual tensor -> the residual is LayerNorm-ed on the fly
    const float* res_gamma;   // [N]
    const float* res_beta;    // [N]
    float* stats_out;         // [M, 2 N / 256, 2]: (sum, sum of squares) of each 128-column piece of the bf16 output rows
    int parts;                // partials per row of in_stats / res_stats (= 2 hidden / 256 when a call of this kernel wrote them)
    float eps;
    long long hidden;         // width the statistics were taken over
};

// Host: 2-D bf16 row-major tensor map, box [box_rows, 64 cols], 128-byte swizzle (gemm_tcgen05.cu).
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows);
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, long long batch, long long rows, long long cols, long long ld,
                      long long batch_stride, int box_rows);
int num_sms();

}  // namespace unirec
