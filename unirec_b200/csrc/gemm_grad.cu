// General-layout tcgen05 GEMM for the backward pass of the projections (sm_100a):
//
//     out[M,N] (+)= sum_k A(m,k) * B(n,k)          bf16 operands, fp32 accumulation in TMEM
//
// where each operand is stored either K-major (the contraction index is the contiguous one: A[M,K] / B[N,K]
// row-major) or MN-major (the output index is the contiguous one: A stored as [K,M] / B stored as [K,N]
// row-major).  UMMA reads both kinds directly from TMA-written, 128-byte-swizzled shared memory (a_major /
// b_major bits of the instruction descriptor + the MN-major shared-memory descriptor), so the two gradient
// GEMMs of y = x W^T need NO transposed copies:
//     dgrad  dx[M,K_in]  = dy[M,N] . W[N,K_in]          A = dy  K-major,   B = W   MN-major (stored [N, K_in])
//     wgrad  dW[N,K_in]  = dy[M,N]^T . x[M,K_in]        A = dy  MN-major (stored [M, N]),  B = x  MN-major
// wgrad contracts over the (long) row dimension and has few output tiles, so the contraction can be split over
// `ksplit` CTAs per tile whose partial tiles are reduced with fp32 red.global.add (accumulate = 1).
// Same warp-specialised single-CTA pipeline as gemm_tcgen05.cu (TMA producer / UMMA issuer / 8 epilogue warps,
// two TMEM accumulator stages); epilogue: optional column bias, bf16 or fp32 stores, or fp32 atomic accumulate.
//
// Reference: autograd of every nn.Linear on the item Q-Former path (training/item_qformer_training.py:129,
// loss.backward() through models/qformer.py:185-198, :286, :359, :372).
#include "common.cuh"
#include "umma_pipe.cuh"

namespace unirec {

struct GemmGParams {
    int M, N, K;
    void* out; long long ldo;
    int out_fp32, accumulate;
    int num_m_blocks, num_n_blocks, ksplit, kb_per_split;
};

constexpr int GG_THREADS = 384;
constexpr int GG_EPI_THREADS = 256;
constexpr int GG_BLOCK_N = 256;
using GGPipe = UmmaPipe<GG_BLOCK_N, 4>;

UNIREC_DEVICE uint64_t gg_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

UNIREC_DEVICE void red_add_f32x4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GG_THREADS, 1)
gemm_bf16_general_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmGParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    GGPipe pipe;
    pipe.setup(smem_raw, warp_idx, lane, GG_EPI_THREADS);
    const int num_kb_total = (p.K + PIPE_BLOCK_K - 1) / PIPE_BLOCK_K;   // TMA zero-fills the K tail
    const int num_work = p.num_m_blocks * p.num_n_blocks * p.ksplit;

    auto decode = [&](int w, int& m_blk, int& n_blk, int& kb0, int& kb1) {
        const int tile = w / p.ksplit, ks = w - tile * p.ksplit;
        m_blk = tile / p.num_n_blocks;
        n_blk = tile - m_blk * p.num_n_blocks;
        kb0 = ks * p.kb_per_split;
        kb1 = min(kb0 + p.kb_per_split, num_kb_total);
    };

    if (warp_idx == 0) {
        // ===================== TMA producer =====================
        RingState rs;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
            int m_blk, n_blk, kb0, kb1;
            decode(w, m_blk, n_blk, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&pipe.empty_bar[rs.stage], rs.phase ^ 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&pipe.full_bar[rs.stage], GGPipe::STAGE_BYTES);
                    uint8_t* sa = pipe.smem_a + rs.stage * PIPE_A_STAGE_BYTES;
                    uint8_t* sb = pipe.smem_b + rs.stage * GGPipe::B_STAGE_BYTES;
                    if constexpr (A_MN) {
                        // A stored [K, M]: two [64 k-rows][64 m] slabs
                        tma_load_2d(&tmap_a, &pipe.full_bar[rs.stage], sa, m_blk * PIPE_BLOCK_M, kb * PIPE_BLOCK_K);
                        tma_load_2d(&tmap_a, &pipe.full_bar[rs.stage], sa + 8192, m_blk * PIPE_BLOCK_M + 64, kb * PIPE_BLOCK_K);
                    } else {
                        tma_load_2d(&tmap_a, &pipe.full_bar[rs.stage], sa, kb * PIPE_BLOCK_K, m_blk * PIPE_BLOCK_M);
                    }
                    if constexpr (B_MN) {
#pragma unroll
                        for (int s = 0; s < GG_BLOCK_N / 64; ++s)
                            tma_load_2d(&tmap_b, &pipe.full_bar[rs.stage], sb + s * 8192, n_blk * GG_BLOCK_N + s * 64,
                                        kb * PIPE_BLOCK_K);
                    } else {
                        tma_load_2d(&tmap_b, &pipe.full_bar[rs.stage], sb, kb * PIPE_BLOCK_K, n_blk * GG_BLOCK_N);
                    }
                }
                __syncwarp();
                if (++rs.stage == GGPipe::STAGES) { rs.stage = 0; rs.phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== UMMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_bf16(PIPE_BLOCK_M, GG_BLOCK_N) | (A_MN ? (1u << 15) : 0u) |
                                   (B_MN ? (1u << 16) : 0u);
        RingState rs;
        uint32_t iter = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++iter) {
            int m_blk, n_blk, kb0, kb1;
            decode(w, m_blk, n_blk, kb0, kb1);
            const uint32_t as = iter & 1u;
            const uint32_t aphase = (iter >> 1) & 1u;
            mbar_wait(&pipe.tmem_empty_bar[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = pipe.tmem_base + as * GG_BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&pipe.full_bar[rs.stage], rs.phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(pipe.smem_a + rs.stage * PIPE_A_STAGE_BYTES);
                    const uint32_t b_addr = smem_u32(pipe.smem_b + rs.stage * GGPipe::B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < PIPE_BLOCK_K / PIPE_UMMA_K; ++k) {
                        const uint64_t da = A_MN ? gg_desc_mn(a_addr + k * 2048, 8192) : umma_smem_desc_sw128(a_addr + k * 32);
                        const uint64_t db = B_MN ? gg_desc_mn(b_addr + k * 2048, 8192) : umma_smem_desc_sw128(b_addr + k * 32);
                        umma_bf16_ss(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&pipe.empty_bar[rs.stage]);
                    if (kb == kb1 - 1) umma_commit(&pipe.tmem_full_bar[as]);
                }
                __syncwarp();
                if (++rs.stage == GGPipe::STAGES) { rs.stage = 0; rs.phase ^= 1; }
            }
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue =====================
        const int q = warp_idx & 3;
        const int half = (warp_idx - 4) >> 2;
        uint32_t iter = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++iter) {
            int m_blk, n_blk, kb0, kb1;
            decode(w, m_blk, n_blk, kb0, kb1);
            const uint32_t tmem_acc = pipe_epilogue_wait<GG_BLOCK_N>(pipe, iter);
            const int row = m_blk * PIPE_BLOCK_M + q * 32 + lane;
            const bool row_ok = row < p.M;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col_in_tile = half * 128 + c * 32;
                const int n0 = n_blk * GG_BLOCK_N + col_in_tile;
                uint32_t v[32];
                tmem_ld_32x32(tmem_acc + col_in_tile + (static_cast<uint32_t>(q * 32) << 16), v);
                tmem_ld_wait();
                if (c == 3) pipe_epilogue_release(pipe, iter);
                if (row_ok && p.out_fp32) {
                    float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + n0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (n0 + j * 4 < p.N) {
                            const float a = __uint_as_float(v[4 * j]), b = __uint_as_float(v[4 * j + 1]);
                            const float cc = __uint_as_float(v[4 * j + 2]), d = __uint_as_float(v[4 * j + 3]);
                            if (p.accumulate) red_add_f32x4(o + j * 4, a, b, cc, d);
                            else *reinterpret_cast<float4*>(o + j * 4) = make_float4(a, b, cc, d);
                        }
                    }
                } else if (row_ok) {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(row) * p.ldo + n0;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n0 + j * 8 < p.N)
                            *reinterpret_cast<uint4*>(o + j * 8) = make_uint4(
                                pack_bf16(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                }
            }
        }
    }
    pipe.teardown(warp_idx);
}

template <bool A_MN, bool B_MN>
static int launch_general(const CUtensorMap& ta, const CUtensorMap& tb, const GemmGParams& p, cudaStream_t stream) {
    auto kern = gemm_bf16_general_kernel<A_MN, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GGPipe::SMEM_BYTES);
        if (e != cudaSuccess) {
            set_last_error("cudaFuncSetAttribute(smem=%d): %s", GGPipe::SMEM_BYTES, cudaGetErrorString(e));
            return UNIREC_ERR_CUDA;
        }
        attr_set = true;
    }
    const int work = p.num_m_blocks * p.num_n_blocks * p.ksplit;
    const int grid = work < num_sms() ? work : num_sms();
    kern<<<grid, GG_THREADS, GGPipe::SMEM_BYTES, stream>>>(ta, tb, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("general gemm launch failed: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

// out[M,N] (+)= A . B^T with per-operand storage order.
//   a_mn = 0: A is [M, K] row-major (lda >= K);  a_mn = 1: A is [K, M] row-major (lda >= M)
//   b_mn = 0: B is [N, K] row-major (ldb >= K);  b_mn = 1: B is [K, N] row-major (ldb >= N)
// accumulate = 1 (fp32 out only): out += result via fp32 atomics, and the contraction may be split (ksplit 0 = auto).
int gemm_bf16_general(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, void* out,
                      long long ldo, int out_fp32, int accumulate, long long M, long long N, long long K, int ksplit,
                      cudaStream_t stream) {
    if (A == nullptr || B == nullptr || out == nullptr || M <= 0 || N <= 0 || K <= 0) {
        set_last_error("gemm_general: null pointer or empty shape (M=%lld N=%lld K=%lld)", M, N, K);
        return UNIREC_ERR_BAD_ARG;
    }
    if (K % 8 != 0 || N % 8 != 0 || lda % 8 != 0 || ldb % 8 != 0 || ldo % (out_fp32 ? 4 : 8) != 0 ||
        (a_mn && M % 8 != 0) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
        (reinterpret_cast<uintptr_t>(out) & 15) || (accumulate && !out_fp32)) {
        set_last_error("gemm_general: need K%%8==0, N%%8==0, 16-byte aligned rows, fp32 out for accumulate "
                       "(M=%lld N=%lld K=%lld lda=%lld ldb=%lld ldo=%lld)", M, N, K, lda, ldb, ldo);
        return UNIREC_ERR_BAD_ARG;
    }
    GemmGParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.out = out; p.ldo = ldo; p.out_fp32 = out_fp32; p.accumulate = accumulate;
    p.num_m_blocks = (int)((M + PIPE_BLOCK_M - 1) / PIPE_BLOCK_M);
    p.num_n_blocks = (int)((N + GG_BLOCK_N - 1) / GG_BLOCK_N);
    const int num_kb = (int)((K + PIPE_BLOCK_K - 1) / PIPE_BLOCK_K);
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    if (!accumulate) ksplit = 1;
    else if (ksplit <= 0) {
        // fill the machine: at least ~2 work items per SM, at least 8 k-blocks per split
        ksplit = (2 * num_sms() + tiles - 1) / tiles;
        const int max_split = num_kb / 8 > 0 ? num_kb / 8 : 1;
        if (ksplit > max_split) ksplit = max_split;
        if (ksplit < 1) ksplit = 1;
    }
    if (ksplit > num_kb) ksplit = num_kb;
    p.kb_per_split = (num_kb + ksplit - 1) / ksplit;
    p.ksplit = (num_kb + p.kb_per_split - 1) / p.kb_per_split;

    CUtensorMap ta, tb;
    int rc = a_mn ? make_tmap_bf16_2d(&ta, A, K, M, lda, PIPE_BLOCK_K) : make_tmap_bf16_2d(&ta, A, M, K, lda, PIPE_BLOCK_M);
    if (rc != UNIREC_OK) return rc;
    rc = b_mn ? make_tmap_bf16_2d(&tb, B, K, N, ldb, PIPE_BLOCK_K) : make_tmap_bf16_2d(&tb, B, N, K, ldb, GG_BLOCK_N);
    if (rc != UNIREC_OK) return rc;
    if (a_mn && b_mn) return launch_general<true, true>(ta, tb, p, stream);
    if (a_mn) return launch_general<true, false>(ta, tb, p, stream);
    if (b_mn) return launch_general<false, true>(ta, tb, p, stream);
    return launch_general<false, false>(ta, tb, p, stream);
}

}  // namespace unirec
