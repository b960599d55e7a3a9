// Hardware probe (not part of the library): can a kernel that allocates TMEM with tcgen05.alloc.cta_group::2 and issues
// tcgen05.mma.cta_group::2 ALSO issue tcgen05.mma.cta_group::1 (each CTA of the pair on its own shared memory, into its own
// TMEM columns)?  The fused K/V-projection + attention kernel wants exactly that: the 256 x 256 projection tile on the CTA
// pair, then per-CTA attention MMAs over the CTA's own 128 keys.  The probe runs
//     pair:  D2[256 x 256] = A[256 x 64] B[256 x 64]^T        (cta_group::2, TMEM columns 0..255)
//     each CTA r:  D1_r[128 x 128] = A_r[128 x 64] B_r[128 x 64]^T  (cta_group::1, TMEM columns 256..383)
//     pair again:  D2 += A B^T                                  (cta_group::2 after cta_group::1)
// and compares all three with a host reference (small integers: exact in bf16 / fp32).
// Build + run (GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I unirec_b200/csrc
//                              tools/probe_mixed_cta_group.cu -o /tmp/probe && /tmp/probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "cg2_ptx.cuh"

using namespace unirec;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* d2, float* d1, int order) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                       // this CTA's 128 rows of A, [128][64] bf16, 128-byte swizzle
    uint8_t* sB = smem + 16384;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768);
    uint64_t* done2 = bars;                   // cta_group::2 commit, multicast to both CTAs
    uint64_t* done1 = bars + 1;               // cta_group::1 commit, local
    uint64_t* ready = bars + 2;               // both CTAs' operands are in shared memory (leader's copy is used)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
    const uint32_t rank = cluster_ctarank();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // operands: row r of this CTA's half, 8 chunks of 8 bf16
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + swz128(r, c)) = *reinterpret_cast<const uint4*>(A + (rank * 128 + r) * 64 + c * 8);
        *reinterpret_cast<uint4*>(sB + swz128(r, c)) = *reinterpret_cast<const uint4*>(B + (rank * 128 + r) * 64 + c * 8);
    }
    fence_proxy_async_smem();
    if (tid == 0) {
        mbar_init(done2, 1);
        mbar_init(done1, 1);
        mbar_init(ready, 2);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc_cg2(tmem_ptr, 512);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    constexpr uint32_t idesc2 = umma_idesc_bf16(256, 256);
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 128);

    auto mma2 = [&](uint32_t acc) {           // leader only
        for (int k = 0; k < 4; ++k)
            umma_bf16_ss_cg2(tmem, umma_smem_desc_sw128(smem_u32(sA) + k * 32), umma_smem_desc_sw128(smem_u32(sB) + k * 32),
                             idesc2, (acc | k) != 0 ? 1u : 0u);
    };
    auto mma1 = [&]() {                       // every CTA, on its own operands
        for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem + 256, umma_smem_desc_sw128(smem_u32(sA) + k * 32), umma_smem_desc_sw128(smem_u32(sB) + k * 32),
                         idesc1, k != 0 ? 1u : 0u);
    };
    if (warp == 0 && lane == 0) {
        if (order == 0) {
            // cta_group::2, then cta_group::1 in both CTAs, then cta_group::2 accumulating on top
            if (rank == 0) { mma2(0); umma_commit_cg2_mc(done2, 0x3); }
            mbar_wait(done2, 0);
            tc_fence_after();
            mma1();
            umma_commit(done1);
            mbar_wait(done1, 0);
            tc_fence_after();
            // the leader may only continue once the peer's cta_group::1 MMAs are done (they read the peer's smem)
            mbar_arrive_cluster(mapa_u32(smem_u32(ready), 0));
            if (rank == 0) {
                mbar_wait_cluster(ready, 0);
                tc_fence_after();
                mma2(1);
                umma_commit_cg2_mc(done2, 0x3);
            }
            mbar_wait(done2, 1);
            tc_fence_after();
        } else {
            // both kinds in flight at once: the leader issues cta_group::2 and cta_group::1 back to back
            if (rank == 0) { mma2(0); mma2(1); umma_commit_cg2_mc(done2, 0x3); }
            mma1();
            umma_commit(done1);
            mbar_wait(done1, 0);
            mbar_wait(done2, 0);
            tc_fence_after();
        }
    }
    __syncthreads();
    tc_fence_after();
    // read back: thread = TMEM lane (row of this CTA's half)
    const uint32_t lane_field = static_cast<uint32_t>(warp * 32) << 16;
    const int row = rank * 128 + warp * 32 + lane;
    for (int c = 0; c < 384; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + lane_field + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) {
            if (c < 256) d2[row * 256 + c + j] = __uint_as_float(v[j]);
            else d1[row * 128 + (c - 256) + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem, 512);
    }
}

int main() {
    std::vector<__nv_bfloat16> hA(256 * 64), hB(256 * 64);
    std::vector<float> fA(256 * 64), fB(256 * 64);
    srand(7);
    for (int i = 0; i < 256 * 64; ++i) {
        fA[i] = static_cast<float>(rand() % 7 - 3);
        fB[i] = static_cast<float>(rand() % 5 - 2);
        hA[i] = __float2bfloat16(fA[i]);
        hB[i] = __float2bfloat16(fB[i]);
    }
    __nv_bfloat16 *dA, *dB;
    float *d2, *d1;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&d2, 256 * 256 * 4); cudaMalloc(&d1, 256 * 128 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    int fails = 0;
    for (int order = 0; order < 2; ++order) {
        cudaMemset(d2, 0xff, 256 * 256 * 4); cudaMemset(d1, 0xff, 256 * 128 * 4);
        probe_kernel<<<2, 128, 40000>>>(dA, dB, d2, d1, order);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("order %d: CUDA error %s\n", order, cudaGetErrorString(e)); return 2; }
        std::vector<float> h2(256 * 256), h1(256 * 128);
        cudaMemcpy(h2.data(), d2, h2.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h1.data(), d1, h1.size() * 4, cudaMemcpyDeviceToHost);
        int bad2 = 0, bad1 = 0;
        for (int m = 0; m < 256; ++m)
            for (int n = 0; n < 256; ++n) {
                float ref = 0.f;
                for (int k = 0; k < 64; ++k) ref += fA[m * 64 + k] * fB[n * 64 + k];
                if (h2[m * 256 + n] != 2.f * ref) ++bad2;
                if ((m >> 7) == (n >> 7) && h1[m * 128 + (n & 127)] != ref) ++bad1;
            }
        printf("order %d: cta_group::2 tile mismatches %d / 65536, cta_group::1 tile mismatches %d / 32768 -> %s\n", order,
               bad2, bad1, (bad2 | bad1) ? "FAIL" : "PASS");
        fails += (bad2 | bad1) ? 1 : 0;
    }
    printf(fails ? "MIXED CTA_GROUP: FAIL\n" : "MIXED CTA_GROUP: PASS\n");
    return fails ? 1 : 0;
}
