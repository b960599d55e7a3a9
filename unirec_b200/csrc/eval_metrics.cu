// Reconstruction-quality metrics of evaluation/evaluate_item_qformer.py:66-95 in ONE pass over the two
// [B, F, E] tensors: per valid (item, field) row the squared error sum (masked MSE numerator, :74-75) and the cosine
// similarity of the reconstructed and the original embedding (:79-88), accumulated over the batch into three doubles
//     acc[0] += sum_valid ||rec - orig||^2      acc[1] += sum_valid cos(rec, orig)      acc[2] += #valid rows
// The reference does this with an unreduced mse_loss tensor, two boolean gathers, two normalisations and three
// .item() synchronisations per batch; here nothing is materialised and nothing synchronises (HBM-bound: both tensors
// are read exactly once, 16-byte loads, one warp per row).
#include "common.cuh"

namespace unirec {

template <bool REC_FP32>
__global__ void __launch_bounds__(256)
reconstruction_metrics_kernel(const void* __restrict__ rec_, const float* __restrict__ orig, const float* __restrict__ mask,
                              long long rows, int E, float eps, double* __restrict__ acc) {
    __shared__ float s_part[8][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float sq = 0.f, cs = 0.f, cnt = 0.f;
    for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < rows;
         row += static_cast<long long>(gridDim.x) * 8) {
        if (__ldg(mask + row) == 0.f) continue;
        const float* o = orig + row * E;
        float d2 = 0.f, dot = 0.f, nr = 0.f, no = 0.f;
        for (int vi = lane; vi < E / 8; vi += 32) {
            const float4 o0 = __ldg(reinterpret_cast<const float4*>(o) + 2 * vi);
            const float4 o1 = __ldg(reinterpret_cast<const float4*>(o) + 2 * vi + 1);
            const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
            float rv[8];
            if constexpr (REC_FP32) {
                const float* r = reinterpret_cast<const float*>(rec_) + row * E;
                const float4 r0 = __ldg(reinterpret_cast<const float4*>(r) + 2 * vi);
                const float4 r1 = __ldg(reinterpret_cast<const float4*>(r) + 2 * vi + 1);
                rv[0] = r0.x; rv[1] = r0.y; rv[2] = r0.z; rv[3] = r0.w; rv[4] = r1.x; rv[5] = r1.y; rv[6] = r1.z; rv[7] = r1.w;
            } else {
                const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(rec_) + row * E;
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(r) + vi);
                rv[0] = bf16_lo(a.x); rv[1] = bf16_hi(a.x); rv[2] = bf16_lo(a.y); rv[3] = bf16_hi(a.y);
                rv[4] = bf16_lo(a.z); rv[5] = bf16_hi(a.z); rv[6] = bf16_lo(a.w); rv[7] = bf16_hi(a.w);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = rv[j] - ov[j];
                d2 += d * d; dot += rv[j] * ov[j]; nr += rv[j] * rv[j]; no += ov[j] * ov[j];
            }
        }
        d2 = warp_sum(d2); dot = warp_sum(dot); nr = warp_sum(nr); no = warp_sum(no);
        sq += d2;
        cs += dot / (fmaxf(sqrtf(nr), eps) * fmaxf(sqrtf(no), eps));     // F.normalize on both sides, then the dot
        cnt += 1.f;
    }
    if (lane == 0) { s_part[warp][0] = sq; s_part[warp][1] = cs; s_part[warp][2] = cnt; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += static_cast<double>(s_part[w][threadIdx.x]);
        if (t != 0.0) atomicAdd(acc + threadIdx.x, t);
    }
}

int reconstruction_metrics(const void* rec, int rec_fp32, const float* orig, const float* mask, long long rows, long long E,
                           float eps, double* acc, cudaStream_t stream) {
    if (rec == nullptr || orig == nullptr || mask == nullptr || acc == nullptr || rows <= 0 || E <= 0 || E % 8 != 0) {
        set_last_error("reconstruction_metrics: null pointer, empty shape or E %% 8 != 0 (rows=%lld E=%lld)", rows, E);
        return UNIREC_ERR_BAD_ARG;
    }
    long long blocks = (rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (rec_fp32)
        reconstruction_metrics_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(rec, orig, mask, rows,
                                                                                              static_cast<int>(E), eps, acc);
    else
        reconstruction_metrics_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(rec, orig, mask, rows,
                                                                                               static_cast<int>(E), eps, acc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("reconstruction_metrics launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

}  // namespace unirec
