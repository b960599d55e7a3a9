// Feature encoders that sit in front of the user-sequence builder, as fused elementwise + MLP kernels:
//
//   * TimestampEncoder       models/mwne.py:504-566  9 features (secular + 4 sin/cos pairs) -> Linear(9, 2D) -> GELU -> Linear(2D, D)
//   * GeoCoordinateEncoder   models/mwne.py:569-610  (lat, lon) -> unit-sphere xyz        -> Linear(3, 2D) -> GELU -> Linear(2D, D)
//   * ImprovedMathematicalEncoder (+ the eval-mode scaling of MathematicallyAwareNormalizer)  models/mwne.py:91-183, :55-62
//
// The reference adds the two event embeddings to every query token of the event's item
// (models/user_sequence_encoder.py:125-131: context = time_emb + geo_emb).  Both MLPs end in a Linear(2D, D), so their
// sum is ONE GEMM over the concatenated hidden activations:
//     ctx = [gelu(W1t f_t + b1t) | gelu(W1g f_g + b1g)] x [W2t | W2g]^T + (b2t + b2g)
// `context_hidden_kernel` produces the bf16 [n, 4D] left operand (features are a handful of fp32 ops per event, the
// 9- and 3-wide first layers are evaluated in registers), the tcgen05 projection GEMM does the rest (unirec_linear_bf16).
// All feature arithmetic is fp32 in the reference's operation order (fmod-based remainder, division by the period, the
// 2 pi multiply, sinf / cosf with full range reduction) so the features agree with torch to a few ulp.
#include "common.cuh"

namespace unirec {

constexpr int ENC_THREADS = 256;

UNIREC_DEVICE float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// models/mwne.py:525-565.  `x` is the timestamp after the reference's `.float()`.
UNIREC_DEVICE void timestamp_features(float x, float (&f)[9]) {
    const float seconds_in_year = 31557600.0f;      // 365.25 * 24 * 60 * 60, exact in fp32
    const float seconds_in_day = 86400.0f;
    const float two_pi = 6.283185307179586f;
    f[0] = x / seconds_in_year;
    float r = fmodf(x, seconds_in_day);             // torch.remainder: fmod, then the sign fix-up
    if (r != 0.f && r < 0.f) r += seconds_in_day;
    const float day_phase = r / seconds_in_day;
    f[1] = sinf(two_pi * day_phase);
    f[2] = cosf(two_pi * day_phase);
    const float week_phase = ((x / seconds_in_day) + 4.0f) / 7.0f;
    f[3] = sinf(two_pi * week_phase);
    f[4] = cosf(two_pi * week_phase);
    float ry = fmodf(x, seconds_in_year);
    if (ry != 0.f && ry < 0.f) ry += seconds_in_year;
    const float year_phase = ry / seconds_in_year;
    f[5] = sinf(two_pi * year_phase);
    f[6] = cosf(two_pi * year_phase);
    const float month_phase = year_phase * 12.0f;
    f[7] = sinf(two_pi * month_phase);
    f[8] = cosf(two_pi * month_phase);
}

// models/mwne.py:596-608
UNIREC_DEVICE void geo_features(float lat_deg, float lon_deg, float (&f)[3]) {
    const float d2r = 0.017453292519943295f;        // torch.deg2rad: x * (pi / 180)
    const float lat = lat_deg * d2r, lon = lon_deg * d2r;
    f[0] = cosf(lat) * cosf(lon);
    f[1] = cosf(lat) * sinf(lon);
    f[2] = sinf(lat);
}

struct ContextParams {
    const void* timestamps;     // [n] fp32 or int64 (ts_int64), or nullptr (time half written as zeros)
    int ts_int64;
    const float* coords;        // [n, 2] fp32 (lat, lon in degrees), or nullptr
    const float* w1t;           // [2D, 9] fp32   (TimestampEncoder.projection.0.weight)
    const float* b1t;           // [2D]
    const float* w1g;           // [2D, 3] fp32   (GeoCoordinateEncoder.projection.0.weight)
    const float* b1g;           // [2D]
    long long n;
    int hidden;                 // 2D
};

// hidden[e, 0:2D] = gelu(W1t f_t(e) + b1t), hidden[e, 2D:4D] = gelu(W1g f_g(e) + b1g); bf16, row stride ld.
// Optionally also the raw features (fp32 [n, 12]) for tests.  One CTA per EV events; a thread owns output columns
// j, j + 256, ... for all EV events, so every first-layer weight row is read once per CTA.
template <int EV>
__global__ void __launch_bounds__(ENC_THREADS)
context_hidden_kernel(const ContextParams p, __nv_bfloat16* __restrict__ hidden, long long ld, float* __restrict__ feats) {
    __shared__ float s_ft[EV][9];
    __shared__ float s_fg[EV][3];
    const long long e0 = static_cast<long long>(blockIdx.x) * EV;
    if (threadIdx.x < EV) {
        const long long e = e0 + threadIdx.x;
        float ft[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, fg[3] = {0, 0, 0};
        if (e < p.n) {
            if (p.timestamps != nullptr) {
                const float x = p.ts_int64 ? __ll2float_rn(reinterpret_cast<const long long*>(p.timestamps)[e])
                                           : reinterpret_cast<const float*>(p.timestamps)[e];
                timestamp_features(x, ft);
            }
            if (p.coords != nullptr) geo_features(p.coords[2 * e], p.coords[2 * e + 1], fg);
            if (feats != nullptr) {
#pragma unroll
                for (int i = 0; i < 9; ++i) feats[e * 12 + i] = ft[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) feats[e * 12 + 9 + i] = fg[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) s_ft[threadIdx.x][i] = ft[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) s_fg[threadIdx.x][i] = fg[i];
    }
    __syncthreads();
    const int H2 = p.hidden;
    for (int j = threadIdx.x; j < 2 * H2; j += ENC_THREADS) {
        const bool geo = j >= H2;
        const int jj = geo ? j - H2 : j;
        float w[9];
        float b;
        bool live;
        if (!geo) {
            live = p.timestamps != nullptr;
#pragma unroll
            for (int i = 0; i < 9; ++i) w[i] = live ? __ldg(p.w1t + jj * 9 + i) : 0.f;
            b = live ? __ldg(p.b1t + jj) : 0.f;
        } else {
            live = p.coords != nullptr;
#pragma unroll
            for (int i = 0; i < 3; ++i) w[i] = live ? __ldg(p.w1g + jj * 3 + i) : 0.f;
            b = live ? __ldg(p.b1g + jj) : 0.f;
        }
#pragma unroll
        for (int ev = 0; ev < EV; ++ev) {
            const long long e = e0 + ev;
            if (e >= p.n) break;
            float acc = 0.f;
            if (!geo) {
#pragma unroll
                for (int i = 0; i < 9; ++i) acc = fmaf(s_ft[ev][i], w[i], acc);
            } else {
#pragma unroll
                for (int i = 0; i < 3; ++i) acc = fmaf(s_fg[ev][i], w[i], acc);
            }
            hidden[e * ld + j] = __float2bfloat16(live ? gelu_exact(acc + b) : 0.f);
        }
    }
}

int context_hidden(const void* timestamps, int ts_int64, const float* coords, const float* w1t, const float* b1t,
                   const float* w1g, const float* b1g, long long n, long long hidden, void* out, long long ldo,
                   float* feats, cudaStream_t stream) {
    if (out == nullptr || n <= 0 || hidden <= 0 || ldo < 2 * hidden || (timestamps == nullptr && coords == nullptr) ||
        (timestamps != nullptr && (w1t == nullptr || b1t == nullptr)) ||
        (coords != nullptr && (w1g == nullptr || b1g == nullptr))) {
        set_last_error("context_hidden: null pointer, empty shape or ldo < 2 * hidden (n=%lld hidden=%lld)", n, hidden);
        return UNIREC_ERR_BAD_ARG;
    }
    ContextParams p;
    p.timestamps = timestamps; p.ts_int64 = ts_int64; p.coords = coords; p.w1t = w1t; p.b1t = b1t; p.w1g = w1g; p.b1g = b1g;
    p.n = n; p.hidden = static_cast<int>(hidden);
    constexpr int EV = 8;
    const unsigned blocks = static_cast<unsigned>((n + EV - 1) / EV);
    context_hidden_kernel<EV><<<blocks, ENC_THREADS, 0, stream>>>(p, reinterpret_cast<__nv_bfloat16*>(out), ldo, feats);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("context_hidden launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

// ImprovedMathematicalEncoder.forward (models/mwne.py:134-183) followed, when `scale` is given, by the eval-mode scaling of
// MathematicallyAwareNormalizer (:55-62, scale[d] = clamp(target_std / (running_std[d] + 1e-8), 0.1, 10) precomputed by the
// host): out[i, :] = scale o [ interleaved (cos, sin)(x f_k) o fourier_weight | (x, sign x) o raw_scale | x * extra_w ].
struct MwneParams {
    const float* numbers;       // [n]
    const float* freqs;         // [F]
    const float* fourier_w;     // [2F]
    const float* raw_scale;     // [2] or nullptr (include_raw = False)
    const float* extra_w;       // [D - 2F - raw] = extra_proj.weight[:, 0], or nullptr
    const float* scale;         // [D] or nullptr
    long long n;
    int F, D;
};

template <bool OUT_FP32>
__global__ void __launch_bounds__(ENC_THREADS)
mwne_encode_kernel(const MwneParams p, void* __restrict__ out_) {
    const long long i = blockIdx.x;
    if (i >= p.n) return;
    const float x = __ldg(p.numbers + i);
    const int raw = p.raw_scale != nullptr ? 2 : 0;
    for (int d = threadIdx.x; d < p.D; d += ENC_THREADS) {
        float v;
        if (d < 2 * p.F) {
            const float ph = x * __ldg(p.freqs + (d >> 1));
            v = ((d & 1) ? sinf(ph) : cosf(ph)) * __ldg(p.fourier_w + d);
        } else if (d < 2 * p.F + raw) {
            const int r = d - 2 * p.F;
            const float sgn = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
            v = (r == 0 ? x : sgn) * __ldg(p.raw_scale + r);
        } else {
            v = x * __ldg(p.extra_w + (d - 2 * p.F - raw));
        }
        if (p.scale != nullptr) v *= __ldg(p.scale + d);
        if constexpr (OUT_FP32) reinterpret_cast<float*>(out_)[i * p.D + d] = v;
        else reinterpret_cast<__nv_bfloat16*>(out_)[i * p.D + d] = __float2bfloat16(v);
    }
}

int mwne_encode(const float* numbers, long long n, const float* freqs, long long F, const float* fourier_w,
                const float* raw_scale, const float* extra_w, const float* scale, long long D, void* out, int out_fp32,
                cudaStream_t stream) {
    const long long raw = raw_scale != nullptr ? 2 : 0;
    if (numbers == nullptr || freqs == nullptr || fourier_w == nullptr || out == nullptr || n <= 0 || F <= 0 ||
        D < 2 * F + raw || (D > 2 * F + raw && extra_w == nullptr) || n > 2147483647LL) {
        set_last_error("mwne_encode: null pointer or embedding_dim too small (n=%lld F=%lld D=%lld)", n, F, D);
        return UNIREC_ERR_BAD_ARG;
    }
    MwneParams p;
    p.numbers = numbers; p.freqs = freqs; p.fourier_w = fourier_w; p.raw_scale = raw_scale; p.extra_w = extra_w;
    p.scale = scale; p.n = n; p.F = static_cast<int>(F); p.D = static_cast<int>(D);
    if (out_fp32) mwne_encode_kernel<true><<<static_cast<unsigned>(n), ENC_THREADS, 0, stream>>>(p, out);
    else mwne_encode_kernel<false><<<static_cast<unsigned>(n), ENC_THREADS, 0, stream>>>(p, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error("mwne_encode launch: %s", cudaGetErrorString(e)); return UNIREC_ERR_CUDA; }
    return UNIREC_OK;
}

}  // namespace unirec
