// Dense projection GEMM for the Q-Former path on sm_100a:
//     out[M,N] = epilogue( A[M,K] (bf16, row-major) x W[N,K]^T (bf16, row-major = nn.Linear.weight) )
// tcgen05.mma (UMMA 128 x BLOCK_N x 16, bf16 -> fp32 accumulators in TMEM), operands staged in shared
// memory by TMA (128-byte swizzle), a persistent warp-specialised CTA per SM:
//     warp 0      TMA producer           (full/empty mbarrier ring, STAGES deep)
//     warp 1      UMMA issuer            (one elected lane; tcgen05.commit releases smem slots)
//     warp 2      TMEM allocator
//     warps 4-11  epilogue               (tcgen05.ld -> bias / erf-GELU / +residual -> bf16|fp32 stores)
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//
// Replaces (reference file:line): every nn.Linear on the path - Q/K/V projections
// models/qformer.py:185-198, attention output dense :286, FFN up :359 (+ GELU :360), FFN down :372,
// heads models/qformer_utils.py:50,53 and training/user_qformer_training.py:38-43.
#include "common.cuh"
#include "umma_pipe.cuh"

#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace unirec {

static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}
const char* get_last_error() { return g_last_error; }

enum : int { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RESIDUAL = 2 };

struct GemmParams {
    int M, N, K;
    const float* bias;               // [N] fp32 or nullptr
    const __nv_bfloat16* residual;   // [*, ldr] bf16 or nullptr (EPI_BIAS_RESIDUAL)
    long long ldr;
    int res_row_mod;                 // >0: residual row = row % res_row_mod (batch-invariant residual)
    void* out;
    long long ldo;
    int out_fp32;
    int num_m_blocks, num_n_blocks;
};

constexpr int BLOCK_M = PIPE_BLOCK_M;
constexpr int BLOCK_K = PIPE_BLOCK_K;
constexpr int NUM_THREADS = 384;
constexpr int NUM_EPI_THREADS = 256;

template <int BLOCK_N>
using Cfg = UmmaPipe<BLOCK_N, (BLOCK_N == 256) ? 4 : 6>;

template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    Cfg<BLOCK_N> pipe;
    pipe.setup(smem_raw, warp_idx, lane, NUM_EPI_THREADS);

    const int num_kb = p.K / BLOCK_K;
    const int num_tiles = p.num_m_blocks * p.num_n_blocks;

    if (warp_idx == 0) {
        // ===================== TMA producer =====================
        RingState rs;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.num_n_blocks;
            const int n_blk = tile % p.num_n_blocks;
            pipe_produce_tile(pipe, rs, &tmap_a, &tmap_b, m_blk * BLOCK_M, n_blk * BLOCK_N, num_kb, lane);
        }
    } else if (warp_idx == 1) {
        // ===================== UMMA issuer =====================
        RingState rs;
        uint32_t iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter)
            pipe_mma_tile<BLOCK_N>(pipe, rs, iter, num_kb, lane);
    } else if (warp_idx >= 4) {
        // ===================== epilogue =====================
        const int q = warp_idx & 3;             // TMEM lane quadrant this warp may access
        const int half = (warp_idx - 4) >> 2;   // which half of the tile's columns
        constexpr int COLS_PER_WARP = BLOCK_N / 2;
        uint32_t iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
            const int m_blk = tile / p.num_n_blocks;
            const int n_blk = tile % p.num_n_blocks;
            const uint32_t tmem_acc = pipe_epilogue_wait<BLOCK_N>(pipe, iter);

            const int row = m_blk * BLOCK_M + q * 32 + lane;
            const bool row_ok = row < p.M;
            const long long res_row = (p.res_row_mod > 0) ? (row % p.res_row_mod) : row;
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                const int col_in_tile = half * COLS_PER_WARP + c * 32;
                const int n0 = n_blk * BLOCK_N + col_in_tile;
                uint32_t v[32];
                tmem_ld_32x32(tmem_acc + col_in_tile + (static_cast<uint32_t>(q * 32) << 16), v);

                uint4 res[4];
                if constexpr (MODE == EPI_BIAS_RESIDUAL) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        res[j] = make_uint4(0, 0, 0, 0);
                        if (row_ok && n0 + j * 8 < p.N)
                            res[j] = __ldg(reinterpret_cast<const uint4*>(p.residual + res_row * p.ldr + n0 + j * 8));
                    }
                }
                tmem_ld_wait();
                if (c == COLS_PER_WARP / 32 - 1) {
                    // all TMEM reads of this accumulator stage are in registers: hand it back to the issuer
                    pipe_epilogue_release(pipe, iter);
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (p.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (n0 + j * 4 < p.N) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 4));
                            f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                        }
                    }
                }
                if constexpr (MODE == EPI_BIAS_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) gelu_erf_x2(f[j], f[j + 1]);
                }
                if constexpr (MODE == EPI_BIAS_RESIDUAL) {
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
                if (row_ok) {
                    if (p.out_fp32) {
                        float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + n0;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (n0 + j * 4 < p.N)
                                *reinterpret_cast<float4*>(o + j * 4) =
                                    make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
                        __nv_bfloat16* o =
                            reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(row) * p.ldo + n0;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (n0 + j * 8 < p.N)
                                *reinterpret_cast<uint4*>(o + j * 8) =
                                    make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                               pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                    }
                }
            }
        }
    }
    pipe.teardown(warp_idx);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    });
    return fn;
}

// 2-D bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                      int box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (fn == nullptr) {
        set_last_error("cuTensorMapEncodeTiled driver entry point not available");
        return UNIREC_ERR_TENSORMAP;
    }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%lld cols=%lld ld=%lld box_rows=%d base=%p)",
                       static_cast<int>(r), rows, cols, ld, box_rows, base);
        return UNIREC_ERR_TENSORMAP;
    }
    return UNIREC_OK;
}

// 3-D bf16 [batch, rows, cols]: element (b, r, c) at base + (b * batch_stride + r * ld + c); box = [1, box_rows, 64 cols],
// 128B swizzle.  Rows >= `rows` of a batch element are out of bounds: zero-filled on load, clipped on store.
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, long long batch, long long rows, long long cols, long long ld,
                      long long batch_stride, int box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (fn == nullptr) {
        set_last_error("cuTensorMapEncodeTiled driver entry point not available");
        return UNIREC_ERR_TENSORMAP;
    }
    if (batch <= 1) { batch = 1; batch_stride = rows * ld; }
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
    cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(batch_stride) * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(3d) failed: CUresult %d (batch=%lld rows=%lld cols=%lld ld=%lld stride=%lld "
                       "box_rows=%d base=%p)", static_cast<int>(r), batch, rows, cols, ld, batch_stride, box_rows, base);
        return UNIREC_ERR_TENSORMAP;
    }
    return UNIREC_OK;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <int BLOCK_N, int MODE>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int max_ctas,
                       cudaStream_t stream) {
    using C = Cfg<BLOCK_N>;
    auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) {
            set_last_error("cudaFuncSetAttribute(smem=%d): %s", C::SMEM_BYTES, cudaGetErrorString(e));
            return UNIREC_ERR_CUDA;
        }
        attr_set = true;
    }
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    int grid = tiles < num_sms() ? tiles : num_sms();
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    kern<<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("gemm launch failed: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

// CTA-pair kernel (gemm_cg2.cu)
bool gemm_cg2_supported(long long M, long long N, long long K, int out_fp32, int res_row_mod, long long lda,
                        long long ldw, long long ldo, long long ldr, int epilogue);
int gemm_bf16_cg2(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* residual,
                  long long ldr, void* out, long long ldo, long long M, long long N, long long K, int epilogue,
                  int max_ctas, cudaStream_t stream, const LnFold* ln = nullptr);

// out = epilogue(LN?(A) W^T + bias [+ LN?(residual)]) with the LayerNorms folded in (see LnFold); CTA-pair kernel only.
int gemm_bf16_ln(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* residual,
                 long long ldr, void* out, long long ldo, long long M, long long N, long long K, int epilogue,
                 const LnFold& ln, cudaStream_t stream) {
    if (A == nullptr || W == nullptr || out == nullptr || bias == nullptr || M <= 0 || N <= 0 || K <= 0) {
        set_last_error("linear_ln: null pointer or empty shape (M=%lld N=%lld K=%lld; bias is required)", M, N, K);
        return UNIREC_ERR_BAD_ARG;
    }
    if (!gemm_cg2_supported(M, N, K, 0, 0, lda, ldw, ldo, ldr, epilogue) || (reinterpret_cast<uintptr_t>(A) & 15) ||
        (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
        (reinterpret_cast<uintptr_t>(bias) & 15)) {
        set_last_error("linear_ln: needs bf16 output, N%%256==0, K%%64==0, 16-byte aligned rows (M=%lld N=%lld K=%lld)", M, N, K);
        return UNIREC_ERR_BAD_ARG;
    }
    if (epilogue == EPI_BIAS_RESIDUAL && (residual == nullptr || (reinterpret_cast<uintptr_t>(residual) & 15))) {
        set_last_error("linear_ln: residual epilogue needs a 16-byte aligned bf16 residual");
        return UNIREC_ERR_BAD_ARG;
    }
    if ((ln.in_stats != nullptr && (ln.in_c == nullptr || (reinterpret_cast<uintptr_t>(ln.in_c) & 15) ||
                                    (reinterpret_cast<uintptr_t>(ln.in_stats) & 7))) ||
        (ln.res_stats != nullptr && (epilogue != EPI_BIAS_RESIDUAL || ln.res_gamma == nullptr || ln.res_beta == nullptr ||
                                     (reinterpret_cast<uintptr_t>(ln.res_gamma) & 15) ||
                                     (reinterpret_cast<uintptr_t>(ln.res_beta) & 15) ||
                                     (reinterpret_cast<uintptr_t>(ln.res_stats) & 7))) ||
        (ln.stats_out != nullptr && (reinterpret_cast<uintptr_t>(ln.stats_out) & 7)) || ln.hidden <= 0 ||
        ((ln.in_stats != nullptr || ln.res_stats != nullptr) && (ln.parts <= 0 || ln.parts > 256))) {
        set_last_error("linear_ln: inconsistent LayerNorm-folding arguments (in_stats needs in_c; res_stats needs the residual "
                       "epilogue, gamma and beta; vectors 16-byte aligned; 1 <= ln_parts <= 256)");
        return UNIREC_ERR_BAD_ARG;
    }
    return gemm_bf16_cg2(A, lda, W, ldw, bias, residual, ldr, out, ldo, M, N, K, epilogue, 0, stream, &ln);
}

int gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* residual,
              long long ldr, int res_row_mod, void* out, long long ldo, int out_fp32, long long M, long long N,
              long long K, int epilogue, int block_n, int max_ctas, cudaStream_t stream) {
    if (A == nullptr || W == nullptr || out == nullptr || M <= 0 || N <= 0 || K <= 0) {
        set_last_error("gemm_bf16: null pointer or empty shape (M=%lld N=%lld K=%lld)", M, N, K);
        return UNIREC_ERR_BAD_ARG;
    }
    if (K % BLOCK_K != 0 || N % 8 != 0 || lda % 8 != 0 || ldw % 8 != 0 || ldo % (out_fp32 ? 4 : 8) != 0 ||
        (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) ||
        (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_last_error("gemm_bf16: need K%%64==0, N%%8==0, 16-byte aligned rows (K=%lld N=%lld lda=%lld ldw=%lld ldo=%lld)",
                       K, N, lda, ldw, ldo);
        return UNIREC_ERR_BAD_ARG;
    }
    if (epilogue == EPI_BIAS_RESIDUAL &&
        (residual == nullptr || ldr % 8 != 0 || (reinterpret_cast<uintptr_t>(residual) & 15))) {
        set_last_error("gemm_bf16: residual epilogue needs a 16-byte aligned bf16 residual");
        return UNIREC_ERR_BAD_ARG;
    }
    if (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15)) {
        set_last_error("gemm_bf16: bias must be 16-byte aligned");
        return UNIREC_ERR_BAD_ARG;
    }
    if (epilogue == EPI_BIAS_RESIDUAL && (reinterpret_cast<uintptr_t>(residual) & 15)) {
        set_last_error("gemm_bf16: residual must be 16-byte aligned");
        return UNIREC_ERR_BAD_ARG;
    }
    // block_n: 0 = auto, 128 / 256 = single-CTA kernel with that tile width, 2 = CTA-pair kernel (256 x 256 per pair)
    const bool cg2_ok = gemm_cg2_supported(M, N, K, out_fp32, res_row_mod, lda, ldw, ldo, ldr, epilogue);
    if (block_n == 2 && !cg2_ok) {
        set_last_error("gemm_bf16: the CTA-pair kernel needs bf16 output, N%%256==0, K%%64==0, no residual row broadcast");
        return UNIREC_ERR_BAD_ARG;
    }
    if (block_n == 0 && cg2_ok) {
        // the pair kernel whenever its 256 x 256 tiles give every SM pair at least one tile
        const long long tiles = ((M + 255) / 256) * (N / 256);
        if (tiles >= num_sms() / 2) block_n = 2;
    }
    if (block_n == 2)
        return gemm_bf16_cg2(A, lda, W, ldw, bias, residual, ldr, out, ldo, M, N, K, epilogue, max_ctas, stream);
    if (block_n == 0) {
        // 256-wide tiles when they fill the machine for >= 2 waves, else 128-wide for more parallelism
        const long long tiles256 = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + 255) / 256);
        block_n = (N % 256 == 0 && tiles256 >= 2LL * num_sms()) ? 256 : 128;
        if (N <= 128) block_n = 128;
    }
    if (block_n != 128 && block_n != 256) {
        set_last_error("gemm_bf16: block_n must be 0, 2, 128 or 256");
        return UNIREC_ERR_BAD_ARG;
    }
    GemmParams p;
    p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
    p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.ldr = ldr; p.res_row_mod = res_row_mod;
    p.out = out; p.ldo = ldo; p.out_fp32 = out_fp32;
    p.num_m_blocks = static_cast<int>((M + BLOCK_M - 1) / BLOCK_M);
    p.num_n_blocks = static_cast<int>((N + block_n - 1) / block_n);

    CUtensorMap ta, tb;
    int rc = make_tmap_bf16_2d(&ta, A, M, K, lda, BLOCK_M);
    if (rc != UNIREC_OK) return rc;
    rc = make_tmap_bf16_2d(&tb, W, N, K, ldw, block_n);
    if (rc != UNIREC_OK) return rc;

#define UNIREC_DISPATCH(BN)                                                                        \
    switch (epilogue) {                                                                            \
        case EPI_BIAS: return launch_gemm<BN, EPI_BIAS>(ta, tb, p, max_ctas, stream);              \
        case EPI_BIAS_GELU: return launch_gemm<BN, EPI_BIAS_GELU>(ta, tb, p, max_ctas, stream);    \
        case EPI_BIAS_RESIDUAL: return launch_gemm<BN, EPI_BIAS_RESIDUAL>(ta, tb, p, max_ctas, stream); \
        default: set_last_error("gemm_bf16: unknown epilogue %d", epilogue); return UNIREC_ERR_BAD_ARG; \
    }
    if (block_n == 256) { UNIREC_DISPATCH(256) } else { UNIREC_DISPATCH(128) }
#undef UNIREC_DISPATCH
}

}  // namespace unirec
