// Row-wise HBM-bound kernels of the Q-Former path (sm_100a): LayerNorm (optionally fused with a
// residual add and a row-broadcast input), dtype casts, token mean-pooling, the field projection
// head, row L2 norms and the user-sequence builder (gather + context + sinusoidal PE).
// All of them move each byte exactly once with 16-byte vector accesses; one warp per row.
#include "common.cuh"

namespace unirec {

// ---------------------------------------------------------------------------------------------
// LayerNorm: y = (x - mean) / sqrt(var + eps) * gamma + beta over the last dim H (H % 8 == 0).
// Replaces nn.LayerNorm at models/qformer.py:104 (embeddings), :288 (attention output), :374 (FFN
// output) and training/user_qformer_training.py:41.  x may be fp32 or bf16; `in_row_mod` > 0 reads
// row (r % in_row_mod) so a batch-invariant input (the learned query tokens) is broadcast for free.
// Statistics in fp32, two-pass over registers (mean first, then centred variance).
// ---------------------------------------------------------------------------------------------
template <bool IN_FP32, int MAX_VEC>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x_, long long ldx, int in_row_mod, const __nv_bfloat16* __restrict__ res,
                 long long ldres, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 void* __restrict__ out_, long long ldo, int out_fp32, int rows, int H) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const long long in_row = in_row_mod > 0 ? (warp % in_row_mod) : warp;
    const int nvec = H / 8;  // 8 elements per vector slot
    float v[MAX_VEC][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            if constexpr (IN_FP32) {
                const float* x = reinterpret_cast<const float*>(x_) + in_row * ldx + vi * 8;
                const float4 a = __ldg(reinterpret_cast<const float4*>(x));
                const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 1);
                v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
                v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
            } else {
                const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(x_) + in_row * ldx + vi * 8;
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(x));
                v[i][0] = bf16_lo(a.x); v[i][1] = bf16_hi(a.x); v[i][2] = bf16_lo(a.y); v[i][3] = bf16_hi(a.y);
                v[i][4] = bf16_lo(a.z); v[i][5] = bf16_hi(a.z); v[i][6] = bf16_lo(a.w); v[i][7] = bf16_hi(a.w);
            }
            if (res != nullptr) {
                const uint4 r = __ldg(reinterpret_cast<const uint4*>(res + static_cast<long long>(warp) * ldres + vi * 8));
                v[i][0] += bf16_lo(r.x); v[i][1] += bf16_hi(r.x); v[i][2] += bf16_lo(r.y); v[i][3] += bf16_hi(r.y);
                v[i][4] += bf16_lo(r.z); v[i][5] += bf16_hi(r.z); v[i][6] += bf16_lo(r.w); v[i][7] += bf16_hi(r.w);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[i][j];
        }
    }
    const float mean = warp_sum(sum) / static_cast<float>(H);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[i][j] - mean;
                sq += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(H) + eps);
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8) + 1);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
            if (out_fp32) {
                float* o = reinterpret_cast<float*>(out_) + static_cast<long long>(warp) * ldo + vi * 8;
                *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
                *(reinterpret_cast<float4*>(o) + 1) = make_float4(y[4], y[5], y[6], y[7]);
            } else {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_) + static_cast<long long>(warp) * ldo + vi * 8;
                *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]),
                                                          pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
            }
        }
    }
}

// Fast path for the layer outputs of the encoder (bf16 in, no residual, no row broadcast): persistent warps, the
// NEXT row's 16-byte loads are issued before the current row is reduced and normalised, and the row is kept
// packed (bf16 pairs) in registers, so that each SM keeps ~128 KB of loads in flight without a bubble between
// the load and the store phase of a row.
template <int MAX_VEC>
__global__ void __launch_bounds__(256)
layernorm_bf16_stream_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float eps, void* __restrict__ out_, long long ldo,
                             int out_fp32, int rows, int H, int reverse) {
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // reverse: visit row (rows - 1 - i) at step i (stream_reverse(), common.cuh); x / out are re-based so that the loop below
    // runs unchanged on mirrored row numbers
    if (reverse) {
        x += static_cast<long long>(rows - 1) * ldx;
        ldx = -ldx;
        out_ = out_fp32 ? static_cast<void*>(reinterpret_cast<float*>(out_) + static_cast<long long>(rows - 1) * ldo)
                        : static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(out_) + static_cast<long long>(rows - 1) * ldo);
        ldo = -ldo;
    }
    const int nvec = H / 8;
    uint4 cur[MAX_VEC], nxt[MAX_VEC];
    auto load = [&](int r, uint4 (&dst)[MAX_VEC]) {
#pragma unroll
        for (int i = 0; i < MAX_VEC; ++i) {
            const int vi = lane + i * 32;
            dst[i] = (vi < nvec) ? __ldg(reinterpret_cast<const uint4*>(x + static_cast<long long>(r) * ldx) + vi)
                                 : make_uint4(0, 0, 0, 0);
        }
    };
    if (row < rows) load(row, cur);
    const float inv_h = 1.0f / static_cast<float>(H);
    while (row < rows) {
        const int next = row + warps_total;
        if (next < rows) load(next, nxt);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < MAX_VEC; ++i) {
            const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) sum += bf16_lo(w[j]) + bf16_hi(w[j]);     // padding lanes hold zeros
        }
        const float mean = warp_sum(sum) * inv_h;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < MAX_VEC; ++i) {
            if (lane + i * 32 < nvec) {
                const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d0 = bf16_lo(w[j]) - mean, d1 = bf16_hi(w[j]) - mean;
                    sq = fmaf(d0, d0, sq);
                    sq = fmaf(d1, d1, sq);
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) * inv_h + eps);
#pragma unroll
        for (int i = 0; i < MAX_VEC; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8) + 1);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8) + 1);
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
                float y[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    y[2 * j] = (bf16_lo(w[j]) - mean) * rstd * g[2 * j] + b[2 * j];
                    y[2 * j + 1] = (bf16_hi(w[j]) - mean) * rstd * g[2 * j + 1] + b[2 * j + 1];
                }
                if (out_fp32) {
                    float* o = reinterpret_cast<float*>(out_) + static_cast<long long>(row) * ldo + vi * 8;
                    *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
                    *(reinterpret_cast<float4*>(o) + 1) = make_float4(y[4], y[5], y[6], y[7]);
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_) + static_cast<long long>(row) * ldo + vi * 8;
                    *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]),
                                                              pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MAX_VEC; ++i) cur[i] = nxt[i];
        row = next;
    }
}

int layernorm(const void* x, int x_fp32, long long ldx, int in_row_mod, const void* residual, long long ldres,
              const float* gamma, const float* beta, float eps, void* out, int out_fp32, long long ldo,
              long long rows, long long H, cudaStream_t stream) {
    if (x == nullptr || gamma == nullptr || beta == nullptr || out == nullptr || rows <= 0 || H <= 0 || H % 8 != 0 ||
        H > 4096 || ldx % 8 != 0 || ldo % 8 != 0) {
        set_last_error("layernorm: bad arguments (rows=%lld H=%lld ldx=%lld ldo=%lld)", rows, H, ldx, ldo);
        return UNIREC_ERR_BAD_ARG;
    }
    const int threads = 256;
    const long long blocks = (rows * 32 + threads - 1) / threads;
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(residual);
    if (!x_fp32 && res == nullptr && in_row_mod == 0 && H <= 1024 && rows >= 4096) {
        // streaming fast path: persistent grid, 4 CTAs of 8 warps per SM
        static int per_sm_1 = 0, per_sm_4 = 0, sms = 0;
        if (sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_1, layernorm_bf16_stream_kernel<1>, threads, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_4, layernorm_bf16_stream_kernel<4>, threads, 0);
            if (per_sm_1 < 1) per_sm_1 = 1;
            if (per_sm_4 < 1) per_sm_4 = 1;
            if (sms < 1) sms = 148;
        }
        long long grid = static_cast<long long>(sms) * (H <= 256 ? per_sm_1 : per_sm_4);
        if (grid > blocks) grid = blocks;
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
        const int rev = stream_reverse() ? 1 : 0;
        if (H <= 256)
            layernorm_bf16_stream_kernel<1><<<static_cast<unsigned>(grid), threads, 0, stream>>>(xb, ldx, gamma, beta, eps, out, ldo,
                                                                                              out_fp32, (int)rows, (int)H, rev);
        else
            layernorm_bf16_stream_kernel<4><<<static_cast<unsigned>(grid), threads, 0, stream>>>(xb, ldx, gamma, beta, eps, out, ldo,
                                                                                              out_fp32, (int)rows, (int)H, rev);
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) {
            set_last_error("layernorm launch: %s", cudaGetErrorString(e2));
            return UNIREC_ERR_CUDA;
        }
        return UNIREC_OK;
    }
#define UNIREC_LN(MV)                                                                                      \
    if (x_fp32)                                                                                            \
        layernorm_kernel<true, MV><<<blocks, threads, 0, stream>>>(x, ldx, in_row_mod, res, ldres, gamma, beta, eps, \
                                                                  out, ldo, out_fp32, (int)rows, (int)H);  \
    else                                                                                                   \
        layernorm_kernel<false, MV><<<blocks, threads, 0, stream>>>(x, ldx, in_row_mod, res, ldres, gamma, beta, eps, \
                                                                   out, ldo, out_fp32, (int)rows, (int)H);
    if (H <= 256) { UNIREC_LN(1) }
    else if (H <= 1024) { UNIREC_LN(4) }
    else { UNIREC_LN(16) }
#undef UNIREC_LN
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("layernorm launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

// ---------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (callers hand fp32 field embeddings, models/qformer_utils.py:37; the kernels
// compute in bf16).  n % 8 == 0.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long nvec) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
        reinterpret_cast<uint4*>(out)[i] =
            make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    }
}

int cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream) {
    if (in == nullptr || out == nullptr || n < 0 || n % 8 != 0) {
        set_last_error("cast_f32_to_bf16: n must be a multiple of 8 (n=%lld)", n);
        return UNIREC_ERR_BAD_ARG;
    }
    if (n == 0) return UNIREC_OK;
    const long long nvec = n / 8;
    long long blocks = (nvec + 255) / 256;
    const long long cap = static_cast<long long>(148) * 16;
    if (blocks > cap) blocks = cap;
    cast_f32_bf16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out), nvec);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("cast launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

// ---------------------------------------------------------------------------------------------
// Mean over the token axis: out[b, :] = mean_t x[b, t, :]   (models/qformer_utils.py:50,
// training/user_qformer_training.py:60, and the pooled scoring vector of SURVEY.md section 8d).
// x bf16 [B, T, H] contiguous rows (row stride ldx), out bf16 or fp32 [B, H].
// One thread per 8 columns; grid (H/8/128, B).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
mean_tokens_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int T, int H, void* __restrict__ out,
                   long long ldo, int out_fp32) {
    const int b = blockIdx.y;
    const int vi = blockIdx.x * blockDim.x + threadIdx.x;
    if (vi * 8 >= H) return;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const __nv_bfloat16* p = x + (static_cast<long long>(b) * T) * ldx + vi * 8;
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p + static_cast<long long>(t) * ldx));
        acc[0] += bf16_lo(a.x); acc[1] += bf16_hi(a.x); acc[2] += bf16_lo(a.y); acc[3] += bf16_hi(a.y);
        acc[4] += bf16_lo(a.z); acc[5] += bf16_hi(a.z); acc[6] += bf16_lo(a.w); acc[7] += bf16_hi(a.w);
    }
    const float inv = 1.0f / static_cast<float>(T);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    if (out_fp32) {
        float* o = reinterpret_cast<float*>(out) + static_cast<long long>(b) * ldo + vi * 8;
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *(reinterpret_cast<float4*>(o) + 1) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + static_cast<long long>(b) * ldo + vi * 8;
        *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]),
                                                  pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
    }
}

int mean_tokens(const void* x, long long ldx, long long B, long long T, long long H, void* out, long long ldo,
                int out_fp32, cudaStream_t stream) {
    if (x == nullptr || out == nullptr || B <= 0 || T <= 0 || H <= 0 || H % 8 != 0 || ldx % 8 != 0 || ldo % 8 != 0 ||
        B > 2147483647LL / 1) {
        set_last_error("mean_tokens: bad arguments (B=%lld T=%lld H=%lld)", B, T, H);
        return UNIREC_ERR_BAD_ARG;
    }
    // grid.y is limited to 65535: fold the batch in slices
    const int gx = static_cast<int>((H / 8 + 127) / 128);
    for (long long b0 = 0; b0 < B; b0 += 65535) {
        const long long nb = (B - b0 < 65535) ? (B - b0) : 65535;
        const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x) + b0 * T * ldx;
        void* op = out_fp32 ? static_cast<void*>(reinterpret_cast<float*>(out) + b0 * ldo)
                            : static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(out) + b0 * ldo);
        mean_tokens_kernel<<<dim3(gx, static_cast<unsigned>(nb)), 128, 0, stream>>>(xp, ldx, (int)T, (int)H, op, ldo,
                                                                                     out_fp32);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("mean_tokens launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

// ---------------------------------------------------------------------------------------------
// Field projection head: out[b, f, :] = sum_t Wp[f, t] * rec[b, t, :] + bp[f]
// (= field_projection(rec.transpose(1,2)).transpose(1,2), models/qformer_utils.py:54).
// rec bf16 [B, T, E]; Wp fp32 [F, T]; out bf16/fp32 [B, F, E].  T <= 64, F <= 32.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
field_projection_kernel(const __nv_bfloat16* __restrict__ rec, const float* __restrict__ Wp,
                        const float* __restrict__ bp, void* __restrict__ out, int out_fp32, int T, int F, int E) {
    extern __shared__ float s_w[];  // [F*T] + [F]
    for (int i = threadIdx.x; i < F * T; i += blockDim.x) s_w[i] = Wp[i];
    for (int i = threadIdx.x; i < F; i += blockDim.x) s_w[F * T + i] = bp[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int vi = blockIdx.x * blockDim.x + threadIdx.x;
    if (vi * 8 >= E) return;
    const __nv_bfloat16* p = rec + (static_cast<long long>(b) * T) * E + vi * 8;
    for (int f0 = 0; f0 < F; f0 += 8) {
        float acc[8][8];
#pragma unroll
        for (int f = 0; f < 8; ++f)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[f][j] = 0.f;
        for (int t = 0; t < T; ++t) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(p + static_cast<long long>(t) * E));
            const float x[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y),
                                bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                const float w = (f0 + f < F) ? s_w[(f0 + f) * T + t] : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[f][j] = fmaf(w, x[j], acc[f][j]);
            }
        }
#pragma unroll
        for (int f = 0; f < 8; ++f) {
            if (f0 + f >= F) break;
            const float bias = s_w[F * T + f0 + f];
            const long long off = (static_cast<long long>(b) * F + f0 + f) * E + vi * 8;
            if (out_fp32) {
                float* o = reinterpret_cast<float*>(out) + off;
                *reinterpret_cast<float4*>(o) =
                    make_float4(acc[f][0] + bias, acc[f][1] + bias, acc[f][2] + bias, acc[f][3] + bias);
                *(reinterpret_cast<float4*>(o) + 1) =
                    make_float4(acc[f][4] + bias, acc[f][5] + bias, acc[f][6] + bias, acc[f][7] + bias);
            } else {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + off;
                *reinterpret_cast<uint4*>(o) = make_uint4(
                    pack_bf16(acc[f][0] + bias, acc[f][1] + bias), pack_bf16(acc[f][2] + bias, acc[f][3] + bias),
                    pack_bf16(acc[f][4] + bias, acc[f][5] + bias), pack_bf16(acc[f][6] + bias, acc[f][7] + bias));
            }
        }
    }
}

int field_projection(const void* rec, const float* Wp, const float* bp, void* out, int out_fp32, long long B,
                     long long T, long long F, long long E, cudaStream_t stream) {
    if (rec == nullptr || Wp == nullptr || bp == nullptr || out == nullptr || B <= 0 || T <= 0 || T > 256 || F <= 0 ||
        F > 256 || E % 8 != 0) {
        set_last_error("field_projection: bad arguments (B=%lld T=%lld F=%lld E=%lld)", B, T, F, E);
        return UNIREC_ERR_BAD_ARG;
    }
    const int gx = static_cast<int>((E / 8 + 127) / 128);
    const size_t smem = static_cast<size_t>(F * T + F) * sizeof(float);
    for (long long b0 = 0; b0 < B; b0 += 65535) {
        const long long nb = (B - b0 < 65535) ? (B - b0) : 65535;
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(rec) + b0 * T * E;
        void* op = out_fp32 ? static_cast<void*>(reinterpret_cast<float*>(out) + b0 * F * E)
                            : static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(out) + b0 * F * E);
        field_projection_kernel<<<dim3(gx, static_cast<unsigned>(nb)), 128, smem, stream>>>(rp, Wp, bp, op, out_fp32,
                                                                                            (int)T, (int)F, (int)E);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("field_projection launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

// ---------------------------------------------------------------------------------------------
// User-sequence builder: seq[b, h*Q + q, :] = item_tokens[history[b,h], q, :] (+ ctx[b,h,:]) + PE[h*Q+q, :]
// for h < lengths[b], zero otherwise; mask[b, s] = s < lengths[b]*Q.
// Replaces models/user_sequence_encoder.py:128-140 (context add, flatten, positional encoding) and
// the right-padding of training/user_qformer_training.py:153-161.  PE is computed in-kernel from
// the closed form (user_sequence_encoder.py:20-24): pe[p, 2i] = sin(p * w_i), pe[p, 2i+1] = cos(p * w_i),
// w_i = exp(-(2i) ln(10000) / D).
// One warp per output row; 16-byte gathers from the item-token table.  With a precomputed table `pe` ([S, D] fp32,
// positional_encoding_kernel below - the same closed form evaluated once; 6.5 MB at S = 1600, L2 resident) the kernel
// is a pure HBM stream; without it the transcendental evaluation makes it issue-bound (ncu: 90 % issue slots).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
positional_encoding_kernel(float* __restrict__ pe, int S, int D) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // one (position, pair) per thread
    const int pairs = D / 2;
    if (i >= static_cast<long long>(S) * pairs) return;
    const int s = static_cast<int>(i / pairs), j = static_cast<int>(i % pairs);
    const float w = expf(static_cast<float>(2 * j) * (-9.210340371976184f / static_cast<float>(D)));
    float sn, cs;
    sincosf(static_cast<float>(s) * w, &sn, &cs);
    reinterpret_cast<float2*>(pe)[i] = make_float2(sn, cs);
}

int positional_encoding(float* pe, long long S, long long D, cudaStream_t stream) {
    if (pe == nullptr || S <= 0 || D <= 0 || D % 2 != 0) {
        set_last_error("positional_encoding: bad arguments (S=%lld D=%lld)", S, D);
        return UNIREC_ERR_BAD_ARG;
    }
    const long long n = S * (D / 2);
    positional_encoding_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(pe, (int)S, (int)D);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("positional_encoding launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

__global__ void __launch_bounds__(256)
build_user_sequence_kernel(const __nv_bfloat16* __restrict__ table, const long long* __restrict__ history,
                           const int* __restrict__ lengths, const __nv_bfloat16* __restrict__ ctx,
                           const float* __restrict__ pe, __nv_bfloat16* __restrict__ seq, float* __restrict__ mask,
                           int Hmax, int Q, int D, long long rows_total, long long num_items) {
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows_total) return;
    const int S = Hmax * Q;
    const long long b = row / S;
    const int s = static_cast<int>(row % S);
    const int h = s / Q;
    const int q = s % Q;
    const bool valid = h < lengths[b];
    if (lane == 0) mask[row] = valid ? 1.f : 0.f;
    __nv_bfloat16* o = seq + row * D;
    if (!valid) {
        for (int vi = lane; vi < D / 8; vi += 32) reinterpret_cast<uint4*>(o)[vi] = make_uint4(0, 0, 0, 0);
        return;
    }
    // ids outside [0, num_items) (sentinels, items missing from the token table) contribute a ZERO token row - the slot
    // keeps its context / position terms and stays attended - exactly what the gathered K/V projection does for them
    // (gemm_cg2.cu: TMA out-of-bounds fill); never an out-of-bounds read
    const long long item = history[b * Hmax + h];
    const bool known = item >= 0 && item < num_items;
    const __nv_bfloat16* src = table + ((known ? item : 0) * Q + q) * D;
    const float neg_ln1e4_over_d = -9.210340371976184f / static_cast<float>(D);
    for (int vi = lane; vi < D / 8; vi += 32) {
        const uint4 a = known ? __ldg(reinterpret_cast<const uint4*>(src) + vi) : make_uint4(0, 0, 0, 0);
        float x[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y),
                      bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
        if (ctx != nullptr) {
            const uint4 c = __ldg(reinterpret_cast<const uint4*>(ctx + (b * Hmax + h) * D) + vi);
            x[0] += bf16_lo(c.x); x[1] += bf16_hi(c.x); x[2] += bf16_lo(c.y); x[3] += bf16_hi(c.y);
            x[4] += bf16_lo(c.z); x[5] += bf16_hi(c.z); x[6] += bf16_lo(c.w); x[7] += bf16_hi(c.w);
        }
        if (pe != nullptr) {
            const float4 p0 = __ldg(reinterpret_cast<const float4*>(pe + static_cast<long long>(s) * D) + 2 * vi);
            const float4 p1 = __ldg(reinterpret_cast<const float4*>(pe + static_cast<long long>(s) * D) + 2 * vi + 1);
            x[0] += p0.x; x[1] += p0.y; x[2] += p0.z; x[3] += p0.w;
            x[4] += p1.x; x[5] += p1.y; x[6] += p1.z; x[7] += p1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float w = expf(static_cast<float>(vi * 8 + 2 * j) * neg_ln1e4_over_d);
                float sn, cs;
                sincosf(static_cast<float>(s) * w, &sn, &cs);
                x[2 * j] += sn;
                x[2 * j + 1] += cs;
            }
        }
        reinterpret_cast<uint4*>(o)[vi] =
            make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    }
}

int build_user_sequence(const void* table, long long num_items, const long long* history, const int* lengths,
                        const void* ctx, const float* pe, void* seq, float* mask, long long B, long long Hmax, long long Q,
                        long long D, cudaStream_t stream) {
    if (table == nullptr || history == nullptr || lengths == nullptr || seq == nullptr || mask == nullptr || B <= 0 ||
        Hmax <= 0 || Q <= 0 || D % 8 != 0 || num_items <= 0) {
        set_last_error("build_user_sequence: bad arguments (B=%lld Hmax=%lld Q=%lld D=%lld)", B, Hmax, Q, D);
        return UNIREC_ERR_BAD_ARG;
    }
    const long long rows = B * Hmax * Q;
    const long long blocks = (rows * 32 + 255) / 256;
    build_user_sequence_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(table), history, lengths, reinterpret_cast<const __nv_bfloat16*>(ctx), pe,
        reinterpret_cast<__nv_bfloat16*>(seq), mask, (int)Hmax, (int)Q, (int)D, rows, num_items);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("build_user_sequence launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

// ---------------------------------------------------------------------------------------------
// Row inverse L2 norms: inv[r] = 1 / max(||x_r||_2, eps)   (F.normalize(p=2, eps=1e-12),
// training/train_item_individual_token_joint.py:405-406,412).  x bf16 or fp32 [rows, D].
// ---------------------------------------------------------------------------------------------
template <bool IN_FP32>
__global__ void __launch_bounds__(256)
inv_l2_norm_kernel(const void* __restrict__ x_, long long ldx, float* __restrict__ inv, long long rows, int D,
                   float eps) {
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float ss = 0.f;
    if constexpr (IN_FP32) {
        const float* x = reinterpret_cast<const float*>(x_) + row * ldx;
        for (int vi = lane; vi < D / 4; vi += 32) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(x) + vi);
            ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
    } else {
        const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(x_) + row * ldx;
        for (int vi = lane; vi < D / 8; vi += 32) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(x) + vi);
            const float f[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y),
                                bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
        }
    }
    ss = warp_sum(ss);
    if (lane == 0) inv[row] = 1.0f / fmaxf(sqrtf(ss), eps);
}

int inv_l2_norm(const void* x, int x_fp32, long long ldx, float* inv, long long rows, long long D, float eps,
                cudaStream_t stream) {
    if (x == nullptr || inv == nullptr || rows <= 0 || D <= 0 || D % 8 != 0 || ldx % 8 != 0) {
        set_last_error("inv_l2_norm: bad arguments (rows=%lld D=%lld)", rows, D);
        return UNIREC_ERR_BAD_ARG;
    }
    const long long blocks = (rows * 32 + 255) / 256;
    if (x_fp32)
        inv_l2_norm_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, ldx, inv, rows, (int)D, eps);
    else
        inv_l2_norm_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, ldx, inv, rows, (int)D, eps);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("inv_l2_norm launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

}  // namespace unirec
