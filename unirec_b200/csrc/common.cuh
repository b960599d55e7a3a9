// Shared device helpers for the unirec_b200 sm_100a kernels: raw PTX wrappers for mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (UMMA issue, TMEM alloc/ld, commit), cp.async, ldmatrix and
// mma.sync, plus small numeric utilities.  Everything here is written against the PTX ISA for
// sm_100a; there is no other architecture path.
#pragma once
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace unirec {

// Traversal order.  The streaming kernels that consume what a GEMM has just written (LayerNorm, the small-tile attention)
// walk their rows / work items from the LAST to the FIRST: a persistent GEMM finishes with its highest row blocks, so those
// are the lines still in the 126 MB L2, and what the streaming kernel writes last (the lowest rows) is what the next GEMM
// reads first.  Any order is arithmetically the same; on a power-capped part the DRAM bytes saved are clock: +0.7 % users/s,
// +0.4 % items/s in an interleaved A/B on one box (profiles/r02_j_*).  Letting the directions ALTERNATE along the whole chain,
// GEMMs included, measured -0.4 % against all-forward (same profiles) and was dropped.  UNIREC_STREAM_REVERSE=0: forward.
inline bool stream_reverse() {
    const char* e = getenv("UNIREC_STREAM_REVERSE");
    return !(e != nullptr && e[0] == '0');
}

#define UNIREC_DEVICE __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// Error codes returned through the C ABI (include/unirec_b200.h)
// ---------------------------------------------------------------------------------------------
enum : int {
    UNIREC_OK = 0,
    UNIREC_ERR_BAD_ARG = 1,       // unsupported shape / alignment / null pointer
    UNIREC_ERR_CUDA = 2,          // a CUDA runtime call failed (see unirec_last_error())
    UNIREC_ERR_TENSORMAP = 3,     // cuTensorMapEncodeTiled failed or driver entry point missing
    UNIREC_ERR_NO_DEVICE = 4,     // device is not compute capability 10.x
};

void set_last_error(const char* fmt, ...);

// ---------------------------------------------------------------------------------------------
// Basic helpers
// ---------------------------------------------------------------------------------------------
UNIREC_DEVICE uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

UNIREC_DEVICE uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

UNIREC_DEVICE uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

UNIREC_DEVICE uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

UNIREC_DEVICE float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
UNIREC_DEVICE float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// Exact-erf GELU (ACT2FN["gelu"], models/qformer.py:353-354): gelu(x) = 0.5 x (1 + erf(x / sqrt 2)).
// erf is evaluated as an odd polynomial t * P(t^2) (degree 19, minimax fit on |t| <= 3, scaled so that the
// clamped tail gives 1 - 2.4e-7) - no MUFU, pure FMA, evaluated two elements at a time with the packed
// FFMA2/FMUL2 instructions of sm_100 (fma.rn.f32x2): ~8 issue slots per element instead of ~45 for erff().
// That matters because the FFN-up epilogue (32768 GELUs per 128 x 256 tile) competes with the tile's 8192
// tensor-core cycles on the FMA pipe.  |erf error| <= 3.6e-5, |gelu error| <= 6.6e-5 absolute over all x
// (checked against float64 on 2M points in [-60, 60]); the value is then stored as bf16 (2^-9 relative).
UNIREC_DEVICE unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
UNIREC_DEVICE void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
UNIREC_DEVICE unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
UNIREC_DEVICE unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
UNIREC_DEVICE unsigned long long mul_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// x0, x1 -> gelu(x0), gelu(x1).  gelu(x) = x * q(x), q = 0.5 + 0.5 erf(x / sqrt 2) = 0.5 + xc * B(xc^2) with
// xc = clamp(x, +-3 sqrt 2): the 1/sqrt 2 scaling and the 0.5 are folded into the coefficients, so the FMA
// pipe sees 12 packed instructions per pair (clamps run on the ALU pipe).
UNIREC_DEVICE void gelu_erf_x2(float& x0, float& x1) {
    constexpr float kB[10] = {3.989351690e-01f, -6.643180549e-02f, 9.905591607e-03f, -1.150100143e-03f,
                              1.037442707e-04f, -7.123461273e-06f, 3.560621167e-07f, -1.206515066e-08f,
                              2.453015291e-10f, -2.242819645e-12f};
    constexpr float kClamp = 4.2426405f;
    const float c0 = fminf(fmaxf(x0, -kClamp), kClamp);
    const float c1 = fminf(fmaxf(x1, -kClamp), kClamp);
    const unsigned long long xc = pack_f32x2(c0, c1);
    const unsigned long long s = mul_f32x2(xc, xc);
    unsigned long long p = fma_f32x2(pack_f32x2(kB[9], kB[9]), s, pack_f32x2(kB[8], kB[8]));
#pragma unroll
    for (int i = 7; i >= 0; --i) p = fma_f32x2(p, s, pack_f32x2(kB[i], kB[i]));
    float q0, q1;
    unpack_f32x2(fma_f32x2(xc, p, pack_f32x2(0.5f, 0.5f)), q0, q1);
    q0 = fmaxf(q0, 0.f);          // the fit leaves q(-clamp) = -3.5e-6: keep gelu(x) -> -0 for very negative x
    q1 = fmaxf(q1, 0.f);
    unpack_f32x2(mul_f32x2(pack_f32x2(x0, x1), pack_f32x2(q0, q1)), x0, x1);
}
UNIREC_DEVICE float gelu_erf(float x) {
    float y = x, z = 0.f;
    gelu_erf_x2(y, z);
    return y;
}

UNIREC_DEVICE float ex2_approx(float x) {   // MUFU.EX2: 2 ulp, ex2(-inf) = 0
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

UNIREC_DEVICE float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

UNIREC_DEVICE float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
UNIREC_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

UNIREC_DEVICE void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

UNIREC_DEVICE void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

UNIREC_DEVICE void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

UNIREC_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

UNIREC_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

UNIREC_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// Non-blocking test (try_wait may suspend the thread up to a system-dependent time limit before it reports failure:
// wrong for a thread that polls several barriers while it has other work to issue).
UNIREC_DEVICE bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
UNIREC_DEVICE bool mbar_test_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Cluster-scope variants (used by the LayerNorm-fused GEMM epilogue to exchange row statistics).
UNIREC_DEVICE uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}

UNIREC_DEVICE void mbar_arrive_cluster(uint32_t remote_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr)
                 : "memory");
}

UNIREC_DEVICE bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

UNIREC_DEVICE void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait_cluster(bar, parity)) {
    }
}

UNIREC_DEVICE void st_cluster_f32x2(uint32_t remote_addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote_addr), "f"(a), "f"(b) : "memory");
}

UNIREC_DEVICE void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
UNIREC_DEVICE void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// 2-D tile load global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
UNIREC_DEVICE void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

UNIREC_DEVICE void tma_load_2d_hint(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int32_t c0,
                                    int32_t c1, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "l"(hint)
        : "memory");
}

// 3-D tile load / store (coordinates innermost first: column, row, batch).
UNIREC_DEVICE void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
UNIREC_DEVICE void tma_store_3d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
UNIREC_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
UNIREC_DEVICE void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
UNIREC_DEVICE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr uint64_t kCacheEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kCacheEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kCacheEvictLast = 0x14F0000000000000ull;

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
UNIREC_DEVICE void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
UNIREC_DEVICE void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
UNIREC_DEVICE void tmem_dealloc(uint32_t tmem_addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}
UNIREC_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
UNIREC_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate, one CTA.
UNIREC_DEVICE void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
UNIREC_DEVICE void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32, both operands K-major.
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   [15] A major (0 = K)  [16] B major (0 = K)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// Shared-memory matrix descriptor for a K-major bf16 tile whose rows are 128 bytes (64 elements)
// written by TMA with CU_TENSOR_MAP_SWIZZLE_128B into a 1024-byte aligned buffer:
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (unused for swizzled K-major: 0)
//   [32,46) stride byte offset >> 4 = 8 rows * 128 B = 1024 B   [46,48) version = 1 (sm_100)
//   [61,64) layout = 2 (SWIZZLE_128B)
UNIREC_DEVICE uint64_t umma_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns (thread i gets lane
// base+i, columns c..c+31).  The lane field of `taddr` must be the warp's own quadrant
// ((warp_id % 4) * 32) << 16.
UNIREC_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 16-column variants (thread i: lane base+i, columns c..c+15)
UNIREC_DEVICE void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
UNIREC_DEVICE void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// explicit shared-state-space accesses (a generic pointer derived from the dynamic smem base compiles to LD.E / ST.E)
UNIREC_DEVICE void st_shared_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
UNIREC_DEVICE void st_shared_u16(uint32_t addr, unsigned short v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
UNIREC_DEVICE unsigned short ld_shared_u16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
UNIREC_DEVICE void st_shared_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
UNIREC_DEVICE float ld_shared_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

UNIREC_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

UNIREC_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
        "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]),
        "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
UNIREC_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// cp.async / ldmatrix / mma.sync (small-query attention kernels)
// ---------------------------------------------------------------------------------------------
UNIREC_DEVICE void cp_async_16(uint32_t smem_dst, const void* gsrc, bool pred) {
    const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
UNIREC_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
UNIREC_DEVICE void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

UNIREC_DEVICE void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
UNIREC_DEVICE void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}

// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
UNIREC_DEVICE void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 128-byte-row tile in shared memory with the 16-byte chunk index XOR-swizzled by (row & 7):
// conflict-free for both the cp.async fill and ldmatrix reads.
UNIREC_DEVICE uint32_t swz128(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }

}  // namespace unirec
