// Backward-pass kernels of the item Q-Former training step (sm_100a): everything that is not a GEMM.
//   gelu_fwd / gelu_bwd      erf-GELU and its derivative (models/qformer.py:360 in training, where the
//                            pre-activation has to be kept for the backward pass)
//   colsum                   bias gradients  db[n] = sum_rows dy[row, n]
//   layernorm_bwd            dx, dgamma, dbeta of LayerNorm(x) (models/qformer.py:104, :288, :374)
//   attention_bwd            dQ, dK, dV of the small-tile attention (models/qformer.py:205, 244-268), one key tile
//                            (nq <= 64, nk <= 64: item self-attention 32 x 32, item cross-attention 32 x 14)
// Reference semantics: autograd of training/item_qformer_training.py:129 (loss.backward()) with dropout disabled.
#include "common.cuh"
#include "dropout.cuh"

namespace unirec {

// q(x) = 0.5 (1 + erf(x / sqrt 2)) = Phi(x), same polynomial as gelu_erf_x2 (common.cuh)
UNIREC_DEVICE float gelu_phi(float x) {
    constexpr float kB[10] = {3.989351690e-01f, -6.643180549e-02f, 9.905591607e-03f, -1.150100143e-03f,
                              1.037442707e-04f, -7.123461273e-06f, 3.560621167e-07f, -1.206515066e-08f,
                              2.453015291e-10f, -2.242819645e-12f};
    const float xc = fminf(fmaxf(x, -4.2426405f), 4.2426405f);
    const float s = xc * xc;
    float p = kB[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) p = fmaf(p, s, kB[i]);
    return fminf(fmaxf(fmaf(xc, p, 0.5f), 0.f), 1.f);
}

template <bool BWD>
__global__ void __launch_bounds__(256)
gelu_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ out,
            long long nvec) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 zv = __ldg(reinterpret_cast<const uint4*>(z) + i);
        const uint32_t zw[4] = {zv.x, zv.y, zv.z, zv.w};
        uint32_t ow[4];
        uint32_t dw[4] = {0, 0, 0, 0};
        if constexpr (BWD) {
            const uint4 dv = __ldg(reinterpret_cast<const uint4*>(da) + i);
            dw[0] = dv.x; dw[1] = dv.y; dw[2] = dv.z; dw[3] = dv.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x0 = bf16_lo(zw[j]), x1 = bf16_hi(zw[j]);
            if constexpr (BWD) {
                // gelu'(x) = Phi(x) + x phi(x), phi(x) = exp(-x^2 / 2) / sqrt(2 pi)
                const float g0 = gelu_phi(x0) + x0 * 0.3989422804f * ex2_approx(-0.72134752f * x0 * x0);
                const float g1 = gelu_phi(x1) + x1 * 0.3989422804f * ex2_approx(-0.72134752f * x1 * x1);
                ow[j] = pack_bf16(bf16_lo(dw[j]) * g0, bf16_hi(dw[j]) * g1);
            } else {
                gelu_erf_x2(x0, x1);
                ow[j] = pack_bf16(x0, x1);
            }
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

int gelu_forward(const void* z, void* out, long long n, cudaStream_t stream) {
    if (z == nullptr || out == nullptr || n <= 0 || n % 8 != 0) {
        set_last_error("gelu_forward: bad arguments (n=%lld)", n);
        return UNIREC_ERR_BAD_ARG;
    }
    const long long nvec = n / 8;
    long long blocks = (nvec + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gelu_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(z), nullptr,
                                                                        reinterpret_cast<__nv_bfloat16*>(out), nvec);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("gelu_forward launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

int gelu_backward(const void* z, const void* da, void* dz, long long n, cudaStream_t stream) {
    if (z == nullptr || da == nullptr || dz == nullptr || n <= 0 || n % 8 != 0) {
        set_last_error("gelu_backward: bad arguments (n=%lld)", n);
        return UNIREC_ERR_BAD_ARG;
    }
    const long long nvec = n / 8;
    long long blocks = (nvec + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gelu_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(z), reinterpret_cast<const __nv_bfloat16*>(da),
        reinterpret_cast<__nv_bfloat16*>(dz), nvec);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("gelu_backward launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

// ---------------------------------------------------------------------------------------------
// colsum: out[n] += sum_r x[r, n]   (x bf16 [rows, N] with row stride ld; out fp32 [N], atomically accumulated)
// block = 32 column vectors (256 columns) x 8 row lanes; grid.y slices the rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, long long rows, int N, float* __restrict__ out) {
    __shared__ float red[8][256];
    const int cv = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + cv) * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (col < N) {
        for (long long r = static_cast<long long>(blockIdx.y) * 8 + rl; r < rows; r += static_cast<long long>(gridDim.y) * 8) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + r * ld + col));
            acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
            acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rl][cv * 8 + j] = acc[j];
    __syncthreads();
    const int c = threadIdx.x;
    if (blockIdx.x * 256 + c < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][c];
        atomicAdd(out + blockIdx.x * 256 + c, s);
    }
}

int colsum(const void* x, long long ld, long long rows, long long N, float* out, cudaStream_t stream) {
    if (x == nullptr || out == nullptr || rows <= 0 || N <= 0 || N % 8 != 0 || ld % 8 != 0) {
        set_last_error("colsum: bad arguments (rows=%lld N=%lld ld=%lld)", rows, N, ld);
        return UNIREC_ERR_BAD_ARG;
    }
    const unsigned gx = static_cast<unsigned>((N + 255) / 256);
    long long gy = (rows + 63) / 64;
    const long long cap = (148LL * 8 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    colsum_kernel<<<dim3(gx, static_cast<unsigned>(gy)), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, rows,
                                                                           (int)N, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("colsum launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward.  y = (x - mu) * rstd * gamma + beta over H (<= 1024, % 8 == 0).
//   xhat = (x - mu) rstd;  g = dy * gamma;  dx = rstd (g - mean(g) - xhat mean(g xhat))
//   dgamma += sum_rows dy xhat;  dbeta += sum_rows dy        (fp32 [H], atomically accumulated)
// x, dy, dx bf16 [rows, H]; one warp per row (persistent), statistics recomputed in fp32.
// dy2 (optional, bf16) is added to dy first: the residual branch hands its gradient to the same tensor.
// ---------------------------------------------------------------------------------------------
// A CTA is split into row GROUPS of `tpg` threads; thread t of a group owns the 8 columns [8 t, 8 t + 8) for EVERY row its
// group processes, so the dgamma / dbeta partial sums are 16 registers per thread.  A group takes LNB_ROWS rows at a
// time (the next rows' 3 x LNB_ROWS 16-byte loads are already in flight), two group-wide reductions (sum x / sum x^2, then sum g /
// sum g xhat) of 2 x LNB_ROWS values each through warp shuffles, a [warps][2 LNB_ROWS] exchange in shared memory and a
// NAMED barrier of the group (the groups of a CTA never wait for each other).  At the end the groups of a CTA add their
// column sums in shared memory and the CTA issues ONE global atomic per column: with one CTA per SM that is 148 atomics
// per address (a grid of 1184 small CTAs measured 0.144 ms, most of it 2.4 M atomics queueing on 64 cache lines).
// (The first version gave every warp a row and kept 64 accumulator registers per thread for the column sums: one CTA
// per SM, reductions on the critical path of every row - 33 % of HBM peak.)
constexpr int LNB_ROWS = 2;
constexpr int LNB_MAX_GROUPS = 8;
constexpr int LNB_MAX_WARPS = 16;

UNIREC_DEVICE void unpack_bf16x8(const uint4& q, float (&f)[8]) {
    f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x); f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
    f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z); f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}

// Sum of v[0 .. 2 LNB_ROWS) over the calling thread's group; every thread of the group gets the totals.  `slot` alternates
// between exchange areas so that one barrier per reduction is enough.
UNIREC_DEVICE void lnb_group_sum(float (&v)[2 * LNB_ROWS], float (*s_x)[LNB_MAX_WARPS][2 * LNB_ROWS], int slot, int warp0,
                                 int gwarps, int bar_id, int tpg) {
#pragma unroll
    for (int i = 0; i < 2 * LNB_ROWS; ++i) v[i] = warp_sum(v[i]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 2 * LNB_ROWS; ++i) s_x[slot][warp][i] = v[i];
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(tpg) : "memory");
#pragma unroll
    for (int i = 0; i < 2 * LNB_ROWS; ++i) {
        float t = 0.f;
        for (int w = 0; w < gwarps; ++w) t += s_x[slot][warp0 + w][i];
        v[i] = t;
    }
}

__global__ void __launch_bounds__(512, 1)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ dy, long long lddy,
                     const __nv_bfloat16* __restrict__ dy2, long long lddy2, const float* __restrict__ gamma, float eps,
                     __nv_bfloat16* __restrict__ dx, long long lddx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     int rows, int H, int tpg, const DropoutParams drop_in, __nv_bfloat16* __restrict__ dx_drop,
                     long long lddrop, float* __restrict__ dbias) {
    __shared__ float s_x[4][LNB_MAX_WARPS][2 * LNB_ROWS];      // [exchange slot][warp][value]
    __shared__ float s_acc[3][1024];                           // column sums of the CTA's groups
    const int groups = blockDim.x / tpg;
    const int group = threadIdx.x / tpg;
    const int tg = threadIdx.x - group * tpg;                  // thread inside the group
    const int gwarps = tpg >> 5, warp0 = group * gwarps;
    const int bar_id = 1 + group;
    const int col = tg * 8;
    const bool active = col < H;                               // tpg * 8 >= H; the last warp of a group may be partly idle
    const float inv_h = 1.0f / static_cast<float>(H);
    for (int i = threadIdx.x; i < 3 * 1024; i += blockDim.x) (&s_acc[0][0])[i] = 0.f;
    // optional second output: dx_drop = dx o mask * scale - the gradient of the dense layer whose dropped output entered this
    // LayerNorm's input sum (models/qformer.py:287-288 / :373-374 backwards) - and its column sums (that layer's bias gradient)
    const DropoutParams drop = dropout_resolve(drop_in);
    float gm[8], adg[8], adb[8], ads[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { gm[j] = 0.f; adg[j] = 0.f; adb[j] = 0.f; ads[j] = 0.f; }
    if (active) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col) + 1);
        gm[0] = g0.x; gm[1] = g0.y; gm[2] = g0.z; gm[3] = g0.w; gm[4] = g1.x; gm[5] = g1.y; gm[6] = g1.z; gm[7] = g1.w;
    }
    const int stride = gridDim.x * groups * LNB_ROWS;
    // raw rows of the NEXT iteration are requested before this iteration's arithmetic (software prefetch: 6 x 16 bytes per
    // thread are always in flight; without it the loads of an iteration only start when the previous one has stored)
    uint4 nx[LNB_ROWS], nd[LNB_ROWS], ne[LNB_ROWS];
    auto fetch = [&](int row0) {
#pragma unroll
        for (int r = 0; r < LNB_ROWS; ++r) {
            const long long row = row0 + r;
            const bool ok = active && row < rows;
            nx[r] = ok ? __ldg(reinterpret_cast<const uint4*>(x + row * ldx + col)) : make_uint4(0u, 0u, 0u, 0u);
            nd[r] = ok ? __ldg(reinterpret_cast<const uint4*>(dy + row * lddy + col)) : make_uint4(0u, 0u, 0u, 0u);
            ne[r] = (ok && dy2 != nullptr) ? __ldg(reinterpret_cast<const uint4*>(dy2 + row * lddy2 + col))
                                           : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    int it = 0;
    const int first = (blockIdx.x * groups + group) * LNB_ROWS;
    fetch(first);
    for (int row0 = first; row0 < rows; row0 += stride, ++it) {
        // ---- this iteration's rows as fp32: x and dy (+ dy2); rows beyond the end and idle columns are zeros
        float f[LNB_ROWS][8], d[LNB_ROWS][8];
#pragma unroll
        for (int r = 0; r < LNB_ROWS; ++r) {
            float e[8];
            unpack_bf16x8(nx[r], f[r]);
            unpack_bf16x8(nd[r], d[r]);
            unpack_bf16x8(ne[r], e);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[r][j] += e[j];
        }
        fetch(row0 + stride);
        // ---- row statistics: mean and E[(x - mean)^2] from sum x and sum x^2 in fp32 (|x| = O(1) pre-LayerNorm sums)
        float red[2 * LNB_ROWS];
#pragma unroll
        for (int r = 0; r < LNB_ROWS; ++r) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1 += f[r][j]; s2 = fmaf(f[r][j], f[r][j], s2); }
            red[2 * r] = s1; red[2 * r + 1] = s2;
        }
        lnb_group_sum(red, s_x, (2 * it) & 3, warp0, gwarps, bar_id, tpg);
        float rstd[LNB_ROWS];
        // ---- xhat in place of x;  g = dy * gamma in place of dy;  m1 = mean(g), m2 = mean(g * xhat);  column sums
#pragma unroll
        for (int r = 0; r < LNB_ROWS; ++r) {
            const float mean = red[2 * r] * inv_h;
            rstd[r] = rsqrtf(fmaxf(red[2 * r + 1] * inv_h - mean * mean, 0.f) + eps);
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (f[r][j] - mean) * rstd[r];
                const float dyv = d[r][j];
                adg[j] = fmaf(dyv, xh, adg[j]);
                adb[j] += dyv;
                const float g = dyv * gm[j];
                f[r][j] = xh;
                d[r][j] = g;
                m1 += g;
                m2 = fmaf(g, xh, m2);
            }
            red[2 * r] = m1; red[2 * r + 1] = m2;
        }
        lnb_group_sum(red, s_x, (2 * it + 1) & 3, warp0, gwarps, bar_id, tpg);
#pragma unroll
        for (int r = 0; r < LNB_ROWS; ++r) {
            const long long row = row0 + r;
            if (!(active && row < rows)) continue;
            const float m1 = red[2 * r] * inv_h, m2 = red[2 * r + 1] * inv_h;
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = rstd[r] * (d[r][j] - m1 - f[r][j] * m2);
            const uint4 packed =
                make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
            *reinterpret_cast<uint4*>(dx + row * lddx + col) = packed;
            if (dx_drop != nullptr) {
                // the mask applies to the bf16-rounded gradient, exactly like dropout_backward on the stored dx
                float q[8];
                unpack_bf16x8(packed, q);
                const uint32_t keep = drop.thr16 != 0
                    ? dropout_keep8(drop, static_cast<unsigned long long>(row), static_cast<uint32_t>(col >> 3)) : 0xffffu;
                const float sc = drop.thr16 != 0 ? drop.scale : 1.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = ((keep >> j) & 1u) ? q[j] * sc : 0.f;
                const uint4 pd =
                    make_uint4(pack_bf16(q[0], q[1]), pack_bf16(q[2], q[3]), pack_bf16(q[4], q[5]), pack_bf16(q[6], q[7]));
                *reinterpret_cast<uint4*>(dx_drop + row * lddrop + col) = pd;
                if (dbias != nullptr) {
                    unpack_bf16x8(pd, q);            // column sums of the values the consumer GEMMs will read
#pragma unroll
                    for (int j = 0; j < 8; ++j) ads[j] += q[j];
                }
            } else if (dbias != nullptr) {
                float q[8];
                unpack_bf16x8(packed, q);
#pragma unroll
                for (int j = 0; j < 8; ++j) ads[j] += q[j];
            }
        }
    }
    __syncthreads();                                 // s_acc is zeroed (every thread passed its zeroing loop)
    if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&s_acc[0][col + j], adg[j]);
            atomicAdd(&s_acc[1][col + j], adb[j]);
            if (dbias != nullptr) atomicAdd(&s_acc[2][col + j], ads[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        atomicAdd(dgamma + i, s_acc[0][i]);
        atomicAdd(dbeta + i, s_acc[1][i]);
        if (dbias != nullptr) atomicAdd(dbias + i, s_acc[2][i]);
    }
}

int layernorm_backward(const void* x, long long ldx, const void* dy, long long lddy, const void* dy2, long long lddy2,
                       const float* gamma, float eps, void* dx, long long lddx, float* dgamma, float* dbeta,
                       long long rows, long long H, unsigned drop_thr16, unsigned long long drop_seed, unsigned drop_site,
                       const unsigned long long* drop_seed_offset, void* dx_drop, long long lddrop, float* dbias,
                       cudaStream_t stream) {
    if (x == nullptr || dy == nullptr || gamma == nullptr || dx == nullptr || dgamma == nullptr || dbeta == nullptr ||
        rows <= 0 || H <= 0 || H > 1024 || H % 8 != 0 || ldx % 8 != 0 || lddy % 8 != 0 || lddx % 8 != 0 ||
        (dy2 != nullptr && lddy2 % 8 != 0) || rows > 2147483647LL || (dx_drop != nullptr && lddrop % 8 != 0) ||
        drop_thr16 >= 65536u || (drop_thr16 != 0 && dx_drop == nullptr)) {
        set_last_error("layernorm_backward: bad arguments (rows=%lld H=%lld)", rows, H);
        return UNIREC_ERR_BAD_ARG;
    }
    const int tpg = static_cast<int>(((H / 8 + 31) / 32) * 32);          // threads per row group: one per 8 columns
    int groups = 512 / tpg;
    if (groups > LNB_MAX_GROUPS) groups = LNB_MAX_GROUPS;
    long long blocks = (rows + groups * LNB_ROWS - 1) / (groups * LNB_ROWS);
    if (blocks > 148) blocks = 148;                                     // one CTA per SM (128 registers x 512 threads)
    DropoutParams dp;
    dp.thr16 = drop_thr16; dp.seed = drop_seed; dp.site = drop_site; dp.seed_offset = drop_seed_offset;
    dp.scale = 65536.0f / (65536.0f - static_cast<float>(drop_thr16));
    layernorm_bwd_kernel<<<static_cast<unsigned>(blocks), groups * tpg, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<const __nv_bfloat16*>(dy), lddy,
        reinterpret_cast<const __nv_bfloat16*>(dy2), lddy2, gamma, eps, reinterpret_cast<__nv_bfloat16*>(dx), lddx, dgamma,
        dbeta, (int)rows, (int)H, tpg, dp, reinterpret_cast<__nv_bfloat16*>(dx_drop), lddrop, dbias);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("layernorm_backward launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

// ---------------------------------------------------------------------------------------------
// Attention backward, one key tile.  Per (batch, head) CTA, each warp owns 16 query rows:
//   S = Q K^T scale + mask;  P = softmax(S);  dP = dO V^T;  dS = P o (dP - rowsum(dP o P)) scale
//   dQ = dS K;   dK = dS^T Q;   dV = P^T dO
// mma.sync m16n8k16 bf16 with fp32 accumulation; P and dS go through shared memory (bf16) for the two
// transposed products.  Layouts as in the forward kernel: rows of 128 B (64 dims) per token, XOR swizzle.
// ---------------------------------------------------------------------------------------------
struct AttnBwdParams {
    const __nv_bfloat16* q; long long ldq; long long q_batch_rows;
    const __nv_bfloat16* k; long long ldk;
    const __nv_bfloat16* v; long long ldv;
    long long kv_batch_rows;
    const float* key_mask;
    const __nv_bfloat16* dout; long long lddo;
    __nv_bfloat16* dq; long long lddq;
    __nv_bfloat16* dk; long long lddk;
    __nv_bfloat16* dv; long long lddv;
    int num_heads, nq, nk;
    float scale;
    DropoutParams drop;      // same (seed, site) as the forward pass; thr16 == 0: no dropout
};

template <int KT>
__global__ void __launch_bounds__(128)
attention_bwd_kernel(const AttnBwdParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.num_heads;
    const int h = blockIdx.x % p.num_heads;
    const int nq_pad = nwarps * 16;

    uint8_t* sQ = smem;                          // nq_pad x 128 B
    uint8_t* sdO = sQ + nq_pad * 128;            // nq_pad x 128 B
    uint8_t* sK = sdO + nq_pad * 128;            // KT x 128 B
    uint8_t* sV = sK + KT * 128;                 // KT x 128 B
    uint8_t* sP = sV + KT * 128;                 // nq_pad x 128 B  (KT <= 64 keys per row)
    uint8_t* sdS = sP + nq_pad * 128;            // nq_pad x 128 B
    float* sM = reinterpret_cast<float*>(sdS + nq_pad * 128);   // KT

    const __nv_bfloat16* qbase = p.q + (static_cast<long long>(b) * p.q_batch_rows) * p.ldq + h * 64;
    const __nv_bfloat16* dobase = p.dout + (static_cast<long long>(b) * p.nq) * p.lddo + h * 64;
    const __nv_bfloat16* kbase = p.k + (static_cast<long long>(b) * p.kv_batch_rows) * p.ldk + h * 64;
    const __nv_bfloat16* vbase = p.v + (static_cast<long long>(b) * p.kv_batch_rows) * p.ldv + h * 64;
    for (int i = threadIdx.x; i < nq_pad * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < p.nq;
        const long long rr = ok ? r : (p.nq - 1);
        cp_async_16(smem_u32(sQ) + swz128(r, c), qbase + rr * p.ldq + c * 8, ok);
        cp_async_16(smem_u32(sdO) + swz128(r, c), dobase + rr * p.lddo + c * 8, ok);
    }
    for (int i = threadIdx.x; i < KT * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < p.nk;
        const long long rr = ok ? r : (p.nk - 1);
        cp_async_16(smem_u32(sK) + swz128(r, c), kbase + rr * p.ldk + c * 8, ok);
        cp_async_16(smem_u32(sV) + swz128(r, c), vbase + rr * p.ldv + c * 8, ok);
    }
    for (int i = threadIdx.x; i < KT; i += blockDim.x) {
        float m = -INFINITY;
        if (i < p.nk) m = (p.key_mask != nullptr && p.key_mask[static_cast<long long>(b) * p.nk + i] == 0.f) ? -1.0e30f : 0.f;
        sM[i] = m;
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const int g4 = lane >> 2, t = lane & 3;
    const float scale_log2 = p.scale * 1.4426950408889634f;

    // ---- S and dP (16 x KT per warp)
    float s[KT / 8][4], dp[KT / 8][4];
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
    {
        uint32_t qf[4][4], dof[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int c = kk * 2 + (lane >> 4);
            ldmatrix_x4(smem_u32(sQ) + swz128(r, c), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
            ldmatrix_x4(smem_u32(sdO) + swz128(r, c), dof[kk][0], dof[kk][1], dof[kk][2], dof[kk][3]);
        }
#pragma unroll
        for (int jj = 0; jj < KT / 16; ++jj) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t b0, b1, b2, b3;
                const int r = jj * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = kk * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(smem_u32(sK) + swz128(r, c), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * jj], qf[kk], b0, b1);
                mma_bf16_16816(s[2 * jj + 1], qf[kk], b2, b3);
                ldmatrix_x4(smem_u32(sV) + swz128(r, c), b0, b1, b2, b3);
                mma_bf16_16816(dp[2 * jj], dof[kk], b0, b1);
                mma_bf16_16816(dp[2 * jj + 1], dof[kk], b2, b3);
            }
        }
    }
    // ---- softmax (single tile) and dS
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
        const float m0 = sM[8 * j + 2 * t], m1 = sM[8 * j + 2 * t + 1];
        s[j][0] = (m0 == 0.f) ? s[j][0] * scale_log2 : m0;
        s[j][1] = (m1 == 0.f) ? s[j][1] * scale_log2 : m1;
        s[j][2] = (m0 == 0.f) ? s[j][2] * scale_log2 : m0;
        s[j][3] = (m1 == 0.f) ? s[j][3] * scale_log2 : m1;
        mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
    float l[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
        s[j][0] = ex2_approx(s[j][0] - mx[0]); s[j][1] = ex2_approx(s[j][1] - mx[0]);
        s[j][2] = ex2_approx(s[j][2] - mx[1]); s[j][3] = ex2_approx(s[j][3] - mx[1]);
        l[0] += s[j][0] + s[j][1];
        l[1] += s[j][2] + s[j][3];
    }
    float rd[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
        l[r] = 1.0f / l[r];
    }
    // dropout (models/qformer.py:258): O = (P o m) V with m in {0, 1/keep}; dP = (dO V^T) o m, and dV uses P o m
    float mk[KT / 8][4];
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) { mk[j][0] = mk[j][1] = mk[j][2] = mk[j][3] = 1.f; }
    if (p.drop.thr16 != 0) {
        const DropoutParams drop = dropout_resolve(p.drop);
        const unsigned long long row0 = static_cast<unsigned long long>(blockIdx.x) * p.nq + warp * 16 + g4;
#pragma unroll
        for (int kb = 0; kb < (KT + 31) / 32; ++kb) {
            const uint32_t k0 = dropout_keep8(drop, row0, static_cast<uint32_t>(kb * 4 + t));
            const uint32_t k1 = dropout_keep8(drop, row0 + 8, static_cast<uint32_t>(kb * 4 + t));
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int j = kb * 4 + nt;
                if (j < KT / 8) {
                    mk[j][0] = ((k0 >> (2 * nt)) & 1u) ? drop.scale : 0.f;
                    mk[j][1] = ((k0 >> (2 * nt + 1)) & 1u) ? drop.scale : 0.f;
                    mk[j][2] = ((k1 >> (2 * nt)) & 1u) ? drop.scale : 0.f;
                    mk[j][3] = ((k1 >> (2 * nt + 1)) & 1u) ? drop.scale : 0.f;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
        s[j][0] *= l[0]; s[j][1] *= l[0]; s[j][2] *= l[1]; s[j][3] *= l[1];          // P
        dp[j][0] *= mk[j][0]; dp[j][1] *= mk[j][1]; dp[j][2] *= mk[j][2]; dp[j][3] *= mk[j][3];
        rd[0] += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
        rd[1] += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        rd[r] += __shfl_xor_sync(0xffffffffu, rd[r], 1);
        rd[r] += __shfl_xor_sync(0xffffffffu, rd[r], 2);
    }
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
        dp[j][0] = s[j][0] * (dp[j][0] - rd[0]) * p.scale; dp[j][1] = s[j][1] * (dp[j][1] - rd[0]) * p.scale;   // dS
        dp[j][2] = s[j][2] * (dp[j][2] - rd[1]) * p.scale; dp[j][3] = s[j][3] * (dp[j][3] - rd[1]) * p.scale;
        const int r0 = warp * 16 + g4;
        const uint32_t off0 = swz128(r0, j) + 4 * t, off1 = swz128(r0 + 8, j) + 4 * t;
        *reinterpret_cast<uint32_t*>(sP + off0) = pack_bf16(s[j][0] * mk[j][0], s[j][1] * mk[j][1]);   // P o m for dV
        *reinterpret_cast<uint32_t*>(sP + off1) = pack_bf16(s[j][2] * mk[j][2], s[j][3] * mk[j][3]);
        *reinterpret_cast<uint32_t*>(sdS + off0) = pack_bf16(dp[j][0], dp[j][1]);
        *reinterpret_cast<uint32_t*>(sdS + off1) = pack_bf16(dp[j][2], dp[j][3]);
    }
    // ---- dQ = dS K  (16 x 64 per warp; dS fragments straight from registers)
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < KT / 16; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
        a[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
        a[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
        a[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            uint32_t b0, b1, b2, b3;
            const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int c = jj * 2 + (lane >> 4);
            ldmatrix_x4_trans(smem_u32(sK) + swz128(r, c), b0, b1, b2, b3);
            mma_bf16_16816(dq[2 * jj], a, b0, b1);
            mma_bf16_16816(dq[2 * jj + 1], a, b2, b3);
        }
    }
    __syncthreads();     // P and dS of every warp are in shared memory

    // ---- dV = P^T dO,  dK = dS^T Q : 16 key rows per tile, key tiles dealt round-robin to the warps
    for (int kt = warp; kt < KT / 16; kt += nwarps) {
        float dv[8][4], dk[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f; dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f; }
        for (int qq = 0; qq < nq_pad / 16; ++qq) {
            // A fragments of P^T / dS^T: transposed 8x8 loads from the [query][key] tiles
            uint32_t ap[4], as_[4];
            const int mi = lane >> 3;
            const int r = qq * 16 + (mi >> 1) * 8 + (lane & 7);
            const int c = kt * 2 + (mi & 1);
            ldmatrix_x4_trans(smem_u32(sP) + swz128(r, c), ap[0], ap[1], ap[2], ap[3]);
            ldmatrix_x4_trans(smem_u32(sdS) + swz128(r, c), as_[0], as_[1], as_[2], as_[3]);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                uint32_t b0, b1, b2, b3;
                const int rb = qq * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int cb = jj * 2 + (lane >> 4);
                ldmatrix_x4_trans(smem_u32(sdO) + swz128(rb, cb), b0, b1, b2, b3);
                mma_bf16_16816(dv[2 * jj], ap, b0, b1);
                mma_bf16_16816(dv[2 * jj + 1], ap, b2, b3);
                ldmatrix_x4_trans(smem_u32(sQ) + swz128(rb, cb), b0, b1, b2, b3);
                mma_bf16_16816(dk[2 * jj], as_, b0, b1);
                mma_bf16_16816(dk[2 * jj + 1], as_, b2, b3);
            }
        }
        // direct 4-byte stores (rows of 128 B are completed by the 4 lanes of a quad x 8 column tiles)
        const int key0 = kt * 16 + g4;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int key = key0 + half * 8;
            if (key < p.nk) {
                __nv_bfloat16* dkrow = p.dk + (static_cast<long long>(b) * p.kv_batch_rows + key) * p.lddk + h * 64;
                __nv_bfloat16* dvrow = p.dv + (static_cast<long long>(b) * p.kv_batch_rows + key) * p.lddv + h * 64;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    *reinterpret_cast<uint32_t*>(dkrow + 8 * j + 2 * t) = pack_bf16(dk[j][2 * half], dk[j][2 * half + 1]);
                    *reinterpret_cast<uint32_t*>(dvrow + 8 * j + 2 * t) = pack_bf16(dv[j][2 * half], dv[j][2 * half + 1]);
                }
            }
        }
    }
    // ---- dQ rows
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int r = warp * 16 + g4 + half * 8;
        if (r < p.nq) {
            __nv_bfloat16* dqrow = p.dq + (static_cast<long long>(b) * p.nq + r) * p.lddq + h * 64;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint32_t*>(dqrow + 8 * j + 2 * t) = pack_bf16(dq[j][2 * half], dq[j][2 * half + 1]);
        }
    }
}

int attention_backward(const void* q, long long ldq, long long q_batch_rows, const void* k, long long ldk, const void* v,
                       long long ldv, long long kv_batch_rows, const float* key_mask, const void* dout, long long lddo,
                       void* dq, long long lddq, void* dk, long long lddk, void* dv, long long lddv, long long batch,
                       long long num_heads, long long nq, long long nk, long long head_dim, float scale,
                       unsigned drop_thr16, unsigned long long drop_seed, unsigned drop_site,
                       const unsigned long long* drop_seed_offset, cudaStream_t stream) {
    if (q == nullptr || k == nullptr || v == nullptr || dout == nullptr || dq == nullptr || dk == nullptr || dv == nullptr ||
        batch <= 0 || num_heads <= 0 || nq <= 0 || nk <= 0) {
        set_last_error("attention_backward: null pointer or empty shape");
        return UNIREC_ERR_BAD_ARG;
    }
    if (head_dim != 64 || nq > 64 || nk > 64 || q_batch_rows != nq || ldq % 8 != 0 || ldk % 8 != 0 || ldv % 8 != 0 ||
        lddo % 8 != 0 || lddq % 2 != 0 || lddk % 2 != 0 || lddv % 2 != 0 || batch * num_heads > 2147483647LL) {
        set_last_error("attention_backward: supports head_dim 64, <= 64 queries and keys, per-batch queries (nq=%lld nk=%lld)",
                       nq, nk);
        return UNIREC_ERR_BAD_ARG;
    }
    AttnBwdParams p;
    p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.ldq = ldq; p.q_batch_rows = q_batch_rows;
    p.k = reinterpret_cast<const __nv_bfloat16*>(k); p.ldk = ldk;
    p.v = reinterpret_cast<const __nv_bfloat16*>(v); p.ldv = ldv;
    p.kv_batch_rows = kv_batch_rows;
    p.key_mask = key_mask;
    p.dout = reinterpret_cast<const __nv_bfloat16*>(dout); p.lddo = lddo;
    p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.lddq = lddq;
    p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.lddk = lddk;
    p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.lddv = lddv;
    p.num_heads = (int)num_heads; p.nq = (int)nq; p.nk = (int)nk;
    p.scale = scale;
    if (drop_thr16 >= 65536u) { set_last_error("attention_backward: dropout probability must be < 1"); return UNIREC_ERR_BAD_ARG; }
    p.drop.thr16 = drop_thr16; p.drop.seed = drop_seed; p.drop.site = drop_site; p.drop.seed_offset = drop_seed_offset;
    p.drop.scale = 65536.0f / (65536.0f - static_cast<float>(drop_thr16));
    const int nwarps = (int)((nq + 15) / 16);
    const int threads = nwarps * 32;
    const unsigned grid = static_cast<unsigned>(batch * num_heads);
    const int kt = nk <= 16 ? 16 : (nk <= 32 ? 32 : 64);
    const size_t smem = static_cast<size_t>(nwarps) * 16 * 128 * 4 + 2 * static_cast<size_t>(kt) * 128 + kt * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attention_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(attention_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(attention_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    if (kt == 16) attention_bwd_kernel<16><<<grid, threads, smem, stream>>>(p);
    else if (kt == 32) attention_bwd_kernel<32><<<grid, threads, smem, stream>>>(p);
    else attention_bwd_kernel<64><<<grid, threads, smem, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("attention_backward launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

}  // namespace unirec
