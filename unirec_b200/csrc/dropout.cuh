// Train-mode dropout masks (models/qformer.py:107, :258, :287, :373 with the module in train()).
// The reference draws its masks from torch's global RNG; here a mask bit is a pure function of
// (seed, site, element) through Philox4x32-10, so the forward pass, the recomputing backward pass and the CPU
// oracle (oracle/dropout_masks.py, pinned against the Random123 known-answer vectors) agree bit for bit:
//   words = philox4x32_10(counter = (row lo32, group, site, row hi32), key = (seed lo32, seed hi32))
//   v[j]  = (words[j >> 1] >> (16 * (j & 1))) & 0xffff,  j = 0..7;   keep iff v[j] >= thr16 = round(p * 65536)
//   kept elements are scaled by 65536 / (65536 - thr16)
//   hidden-state sites ([rows, H]): group = col >> 3, j = col & 7
//   attention-probability sites: row = (b * heads + h) * nq + q, group = (k >> 5) * 4 + ((k & 7) >> 1),
//                                j = 2 * ((k & 31) >> 3) + (k & 1)   (= the 8 probabilities one lane of an
//                                mma.sync accumulator quad owns inside a 32-key block)
#pragma once

#include "common.cuh"

namespace unirec {

struct DropoutParams {
    uint32_t thr16;          // 0 = dropout off
    uint32_t site;
    unsigned long long seed;
    float scale;             // 65536 / (65536 - thr16)
    // optional device-resident addend of the seed (NULL = none): lets a CUDA graph that was captured once with a fixed
    // `seed` draw fresh masks on every replay - the graph itself increments *seed_offset (unirec_b200/training.py)
    const unsigned long long* seed_offset;
};

// Effective parameters of a kernel: the seed with the device-resident offset folded in (one uniform load per thread).
UNIREC_DEVICE DropoutParams dropout_resolve(DropoutParams d) {
    if (d.seed_offset != nullptr) d.seed += __ldg(d.seed_offset);
    d.seed_offset = nullptr;
    return d;
}

UNIREC_DEVICE DropoutParams make_dropout(uint32_t thr16, unsigned long long seed, uint32_t site,
                                         const unsigned long long* seed_offset = nullptr) {
    DropoutParams d;
    d.thr16 = thr16; d.site = site; d.seed = seed; d.seed_offset = seed_offset;
    d.scale = 65536.0f / (65536.0f - static_cast<float>(thr16));
    return dropout_resolve(d);
}

UNIREC_DEVICE uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1;
        c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// The 8 values of one (row, group) as a keep bitmask: bit j set iff element j is kept.
UNIREC_DEVICE uint32_t dropout_keep8(const DropoutParams& d, unsigned long long row, uint32_t group) {
    const uint4 w = philox4x32_10(static_cast<uint32_t>(row), group, d.site, static_cast<uint32_t>(row >> 32),
                                  static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bits |= ((ws[i] & 0xffffu) >= d.thr16 ? 1u : 0u) << (2 * i);
        bits |= ((ws[i] >> 16) >= d.thr16 ? 1u : 0u) << (2 * i + 1);
    }
    return bits;
}

}  // namespace unirec
