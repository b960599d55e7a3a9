// Hidden-state dropout of the training step (models/qformer.py:107, :287, :373):
//   forward   out = dropout(x) [+ residual]          (the pre-LayerNorm sum of BertSelfOutput / BertOutput, or the
//                                                      dropped query embeddings when residual == NULL)
//   backward  dx  = dy o mask * scale
// bf16 rows, 8 elements (one Philox call, one 16-byte access) per thread step; HBM-bound streaming kernels.
#include "dropout.cuh"

namespace unirec {

template <bool BWD>
__global__ void __launch_bounds__(256)
dropout_rows_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int x_row_mod,
                    const __nv_bfloat16* __restrict__ res, long long ldres, __nv_bfloat16* __restrict__ out,
                    long long ldo, long long rows, int H, uint32_t thr16, unsigned long long seed, uint32_t site,
                    const unsigned long long* __restrict__ seed_offset) {
    const DropoutParams d = make_dropout(thr16, seed, site, seed_offset);
    const int groups = H >> 3;
    const long long total = rows * groups;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long row = i / groups;
        const int g = static_cast<int>(i - row * groups);
        const long long xr = x_row_mod > 0 ? row % x_row_mod : row;
        const uint4 xv = *reinterpret_cast<const uint4*>(x + xr * ldx + g * 8);
        const uint32_t keep = dropout_keep8(d, static_cast<unsigned long long>(row), static_cast<uint32_t>(g));
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
        uint32_t rs[4] = {0u, 0u, 0u, 0u};
        if (!BWD && res != nullptr) {
            const uint4 rv = *reinterpret_cast<const uint4*>(res + row * ldres + g * 8);
            rs[0] = rv.x; rs[1] = rv.y; rs[2] = rv.z; rs[3] = rv.w;
        }
        uint32_t os[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float lo = ((keep >> (2 * j)) & 1u) ? bf16_lo(xs[j]) * d.scale : 0.f;
            const float hi = ((keep >> (2 * j + 1)) & 1u) ? bf16_hi(xs[j]) * d.scale : 0.f;
            os[j] = pack_bf16(lo + bf16_lo(rs[j]), hi + bf16_hi(rs[j]));
        }
        *reinterpret_cast<uint4*>(out + row * ldo + g * 8) = make_uint4(os[0], os[1], os[2], os[3]);
    }
}

static int launch_rows(bool bwd, const void* x, long long ldx, long long x_row_mod, const void* res, long long ldres,
                       void* out, long long ldo, long long rows, long long H, unsigned thr16, unsigned long long seed,
                       unsigned site, const unsigned long long* seed_offset, cudaStream_t stream, const char* what) {
    if (x == nullptr || out == nullptr || rows <= 0 || H <= 0 || H % 8 != 0 || ldx % 8 != 0 || ldo % 8 != 0 ||
        (res != nullptr && ldres % 8 != 0) || thr16 >= 65536u) {
        set_last_error("%s: null pointer, H %% 8 != 0, unaligned rows or p >= 1 (H=%lld)", what, H);
        return UNIREC_ERR_BAD_ARG;
    }
    long long blocks = (rows * (H / 8) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (bwd)
        dropout_rows_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(x), ldx, 0, nullptr, 0, reinterpret_cast<__nv_bfloat16*>(out), ldo, rows,
            static_cast<int>(H), thr16, seed, site, seed_offset);
    else
        dropout_rows_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(x), ldx, static_cast<int>(x_row_mod),
            reinterpret_cast<const __nv_bfloat16*>(res), ldres, reinterpret_cast<__nv_bfloat16*>(out), ldo, rows,
            static_cast<int>(H), thr16, seed, site, seed_offset);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("%s launch: %s", what, cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

int dropout_add(const void* x, long long ldx, long long x_row_mod, const void* res, long long ldres, void* out,
                long long ldo, long long rows, long long H, unsigned thr16, unsigned long long seed, unsigned site,
                const unsigned long long* seed_offset, cudaStream_t stream) {
    return launch_rows(false, x, ldx, x_row_mod, res, ldres, out, ldo, rows, H, thr16, seed, site, seed_offset, stream,
                       "dropout_add");
}

int dropout_backward(const void* dy, long long lddy, void* dx, long long lddx, long long rows, long long H,
                     unsigned thr16, unsigned long long seed, unsigned site, const unsigned long long* seed_offset,
                     cudaStream_t stream) {
    return launch_rows(true, dy, lddy, 0, nullptr, 0, dx, lddx, rows, H, thr16, seed, site, seed_offset, stream,
                       "dropout_backward");
}

}  // namespace unirec
