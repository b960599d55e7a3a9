// Fused small-query attention for the Q-Former path (sm_100a), head_dim = 64.
//
//   ctx[b, q, h*64:(h+1)*64] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] / 8 + mask[b,k] ) @ V[b,k,h,:]
//
// Replaces models/qformer.py:161-167 (head split), :205 (QK^T), :244-250 (scale, additive mask,
// softmax), :264-268 (PV, head merge by permute+contiguous).  Used for
//   item self-attention   32 queries x 32 keys          (KT = 32, one tile)
//   item cross-attention  32 queries x 14 field keys    (KT = 16, one tile)
//   user self-attention   64 queries x 64 keys          (KT = 64, one tile)
//   user cross-attention  64 queries x <= 1600 keys     (KT = 64, key-tiled online softmax)
// One CTA per (batch, head); each warp owns 16 query rows.  Q, K and V tiles are staged in shared
// memory with cp.async (16-byte chunks, XOR swizzle), the two small matmuls run on mma.sync
// m16n8k16 (bf16 in, fp32 accumulate) from ldmatrix fragments, softmax statistics are fp32 with
// quad shuffles, and the context is written head-merged ([B*Q, H] row-major) through shared memory
// so that global stores are 128-byte row segments.  The kernel is HBM-bound by design: algorithmic
// bytes = (Q + K + V + O) tiles, each read/written exactly once.
//
// Mask semantics: key_mask[b,k] == 0 adds a large negative constant (the
// reference adds finfo.min, models/qformer.py:927-933 via HF invert_attention_mask) so masked keys get
// probability exactly 0 whenever any key is unmasked, and a row whose keys are ALL masked
// degenerates to uniform attention over all nk keys (not NaN, not zero) - exactly the fp32
// behaviour of the reference, where score + finfo.min == finfo.min for every key.
#include "common.cuh"

namespace unirec {

struct AttnParams {
    const __nv_bfloat16* q; long long ldq; long long q_batch_rows;   // 0 => same queries for every batch
    const __nv_bfloat16* k; long long ldk;
    const __nv_bfloat16* v; long long ldv;
    long long kv_batch_rows;
    const float* key_mask;   // [B, nk] (1 attend / 0 masked) or nullptr
    __nv_bfloat16* out; long long ldo;
    int num_heads, nq, nk;
    float scale_log2;        // softmax scale * log2(e)
};

constexpr float kMaskedLog2 = -1.0e30f;  // stands in for finfo.min (see header comment)

template <int KT>
__global__ void __launch_bounds__(128)
attention_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.num_heads;
    const int h = blockIdx.x % p.num_heads;
    const int nq_pad = nwarps * 16;

    uint8_t* sQ = smem;                               // nq_pad rows x 128 B
    uint8_t* sK = sQ + nq_pad * 128;                  // 2 x KT rows x 128 B
    uint8_t* sV = sK + 2 * KT * 128;                  // 2 x KT rows x 128 B
    float* sM = reinterpret_cast<float*>(sV + 2 * KT * 128);  // 2 x KT additive mask (log2 domain)

    const __nv_bfloat16* qbase = p.q + (static_cast<long long>(b) * p.q_batch_rows) * p.ldq + h * 64;
    const __nv_bfloat16* kbase = p.k + (static_cast<long long>(b) * p.kv_batch_rows) * p.ldk + h * 64;
    const __nv_bfloat16* vbase = p.v + (static_cast<long long>(b) * p.kv_batch_rows) * p.ldv + h * 64;
    const float* mbase = p.key_mask ? p.key_mask + static_cast<long long>(b) * p.nk : nullptr;

    auto load_kv_tile = [&](int tile, int buf) {
        const int base = tile * KT;
        for (int i = threadIdx.x; i < KT * 8; i += blockDim.x) {
            const int r = i >> 3, c = i & 7;
            const int key = base + r;
            const bool ok = key < p.nk;
            const long long kr = ok ? key : (p.nk - 1);
            cp_async_16(smem_u32(sK + buf * KT * 128) + swz128(r, c), kbase + kr * p.ldk + c * 8, ok);
            cp_async_16(smem_u32(sV + buf * KT * 128) + swz128(r, c), vbase + kr * p.ldv + c * 8, ok);
        }
        for (int i = threadIdx.x; i < KT; i += blockDim.x) {
            const int key = base + i;
            float m = -INFINITY;  // padding beyond nk: excluded
            if (key < p.nk) m = (mbase != nullptr && mbase[key] == 0.f) ? kMaskedLog2 : 0.f;
            sM[buf * KT + i] = m;
        }
    };

    // ---- prologue: Q tile + first K/V tile
    for (int i = threadIdx.x; i < nq_pad * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < p.nq;
        const long long qr = ok ? r : (p.nq - 1);
        cp_async_16(smem_u32(sQ) + swz128(r, c), qbase + qr * p.ldq + c * 8, ok);
    }
    load_kv_tile(0, 0);
    cp_async_commit();

    const int ntiles = (p.nk + KT - 1) / KT;
    const int g = lane >> 2, t = lane & 3;

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) {
            load_kv_tile(tile + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (tile == 0) {
            // Q fragments (A operand, 16 rows x 64 dims = 4 k-steps) stay in registers for all tiles
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = kk * 2 + (lane >> 4);
                ldmatrix_x4(smem_u32(sQ) + swz128(r, c), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
            }
        }

        // ---- S = Q K^T  (16 x KT per warp)
        float s[KT / 8][4];
#pragma unroll
        for (int j = 0; j < KT / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
        const uint32_t kaddr = smem_u32(sK + buf * KT * 128);
#pragma unroll
        for (int jj = 0; jj < KT / 16; ++jj) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t b0, b1, b2, b3;
                const int r = jj * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = kk * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(kaddr + swz128(r, c), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * jj], qf[kk], b0, b1);
                mma_bf16_16816(s[2 * jj + 1], qf[kk], b2, b3);
            }
        }

        // ---- scale + mask, online softmax (log2 domain)
        const float* mt = sM + buf * KT;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < KT / 8; ++j) {
            const float m0 = mt[8 * j + 2 * t], m1 = mt[8 * j + 2 * t + 1];
            // masked keys: the score is absorbed by the huge constant exactly as in the fp32 reference
            s[j][0] = (m0 == 0.f) ? s[j][0] * p.scale_log2 : m0;
            s[j][1] = (m1 == 0.f) ? s[j][1] * p.scale_log2 : m1;
            s[j][2] = (m0 == 0.f) ? s[j][2] * p.scale_log2 : m0;
            s[j][3] = (m1 == 0.f) ? s[j][3] * p.scale_log2 : m1;
            mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
        }
        float alpha[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);   // finite: key 0 of tile 0 always exists
            alpha[r] = exp2f(m_run[r] - m_new);           // first tile: exp2(-inf) = 0
            m_run[r] = m_new;
            l_run[r] *= alpha[r];
        }
#pragma unroll
        for (int j = 0; j < KT / 8; ++j) {
            s[j][0] = exp2f(s[j][0] - m_run[0]);
            s[j][1] = exp2f(s[j][1] - m_run[0]);
            s[j][2] = exp2f(s[j][2] - m_run[1]);
            s[j][3] = exp2f(s[j][3] - m_run[1]);
            l_run[0] += s[j][0] + s[j][1];
            l_run[1] += s[j][2] + s[j][3];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
            o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
        }

        // ---- O += P V
        const uint32_t vaddr = smem_u32(sV + buf * KT * 128);
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
            uint32_t a[4];
            a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
            a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
            a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                uint32_t b0, b1, b2, b3;
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = jj * 2 + (lane >> 4);
                ldmatrix_x4_trans(vaddr + swz128(r, c), b0, b1, b2, b3);
                mma_bf16_16816(o[2 * jj], a, b0, b1);
                mma_bf16_16816(o[2 * jj + 1], a, b2, b3);
            }
        }
        __syncthreads();  // everyone done with buffer `buf` before it is refilled two tiles later
    }

    // ---- normalise and write the context tile through shared memory (reuses this warp's Q rows)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
        l_run[r] = 1.0f / l_run[r];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int r0 = warp * 16 + g;
        *reinterpret_cast<uint32_t*>(sQ + swz128(r0, j) + 4 * t) = pack_bf16(o[j][0] * l_run[0], o[j][1] * l_run[0]);
        *reinterpret_cast<uint32_t*>(sQ + swz128(r0 + 8, j) + 4 * t) =
            pack_bf16(o[j][2] * l_run[1], o[j][3] * l_run[1]);
    }
    __syncwarp();
    __nv_bfloat16* obase = p.out + (static_cast<long long>(b) * p.nq) * p.ldo + h * 64;
    for (int i = lane; i < 16 * 8; i += 32) {
        const int r = warp * 16 + (i >> 3), c = i & 7;
        if (r < p.nq) {
            const uint4 val = *reinterpret_cast<const uint4*>(sQ + swz128(r, c));
            *reinterpret_cast<uint4*>(obase + static_cast<long long>(r) * p.ldo + c * 8) = val;
        }
    }
}

int attention(const void* q, long long ldq, long long q_batch_rows, const void* k, long long ldk, const void* v,
              long long ldv, long long kv_batch_rows, const float* key_mask, void* out, long long ldo, long long batch,
              long long num_heads, long long nq, long long nk, long long head_dim, float scale, cudaStream_t stream) {
    if (q == nullptr || k == nullptr || v == nullptr || out == nullptr || batch <= 0 || num_heads <= 0 || nq <= 0 ||
        nk <= 0) {
        set_last_error("attention: null pointer or empty shape");
        return UNIREC_ERR_BAD_ARG;
    }
    if (head_dim != 64 || nq > 64 || ldq % 8 != 0 || ldk % 8 != 0 || ldv % 8 != 0 || ldo % 8 != 0 ||
        batch * num_heads > 2147483647LL) {
        set_last_error("attention: supports head_dim 64, <= 64 queries, 16-byte aligned rows (head_dim=%lld nq=%lld)",
                       head_dim, nq);
        return UNIREC_ERR_BAD_ARG;
    }
    AttnParams p;
    p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.ldq = ldq; p.q_batch_rows = q_batch_rows;
    p.k = reinterpret_cast<const __nv_bfloat16*>(k); p.ldk = ldk;
    p.v = reinterpret_cast<const __nv_bfloat16*>(v); p.ldv = ldv;
    p.kv_batch_rows = kv_batch_rows;
    p.key_mask = key_mask;
    p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo;
    p.num_heads = static_cast<int>(num_heads); p.nq = static_cast<int>(nq); p.nk = static_cast<int>(nk);
    p.scale_log2 = scale * 1.4426950408889634f;
    const int nwarps = static_cast<int>((nq + 15) / 16);
    const int threads = nwarps * 32;
    const unsigned grid = static_cast<unsigned>(batch * num_heads);
    const int kt = nk <= 16 ? 16 : (nk <= 32 ? 32 : 64);
    const size_t smem = static_cast<size_t>(nwarps) * 16 * 128 + 4 * static_cast<size_t>(kt) * 128 + 2 * kt * sizeof(float);
    if (kt == 16) attention_kernel<16><<<grid, threads, smem, stream>>>(p);
    else if (kt == 32) attention_kernel<32><<<grid, threads, smem, stream>>>(p);
    else attention_kernel<64><<<grid, threads, smem, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("attention launch: %s", cudaGetErrorString(e));
        return UNIREC_ERR_CUDA;
    }
    return UNIREC_OK;
}

}  // namespace unirec
