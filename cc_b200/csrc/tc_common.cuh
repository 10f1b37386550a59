// tc_common.cuh - device helpers shared by the TMA-fed tcgen05 convolution kernels (conv_tma.cu, conv_nhwc.cu):
// mbarrier / TMA / tcgen05 PTX wrappers, UMMA shared-memory descriptors, the tensor-map encoder entry point.
#pragma once
#include "ccb_common.cuh"
#ifndef CCB_CPU_SIM
#include <cstdlib>
#include <cuda.h>   // CUtensorMap + enums only; cuTensorMapEncodeTiled is fetched through the runtime (no libcuda link)

namespace ccb {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void tm_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool tm_mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_addr(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar))
        : "memory");
}
__device__ __forceinline__ void tm_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-uniform issue: every lane runs the surrounding code (so descriptors stay in uniform registers and no
// divergent region is entered), only the leader's predicate lets the instruction through.
// Descriptors are passed as their two 32-bit halves: the high half is constant per operand kind.
__device__ __forceinline__ void tm_umma_tf32_p(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate, uint32_t leader) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tm_commit_p(uint64_t* bar, uint32_t leader) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_addr(bar)), "r"(leader)
        : "memory");
}
// halves of the shared-memory matrix descriptors (version 1): low = start >> 4 | LBO >> 4 << 16, high = SBO >> 4 | 1 << 14 | layout << 29
constexpr uint32_t DESC_A_MN_LO = (4096u >> 4) << 16, DESC_A_MN_HI = (512u >> 4) | (1u << 14) | (1u << 29);   // MN-major SWIZZLE_128B_BASE32B
constexpr uint32_t DESC_K_LO = (16u >> 4) << 16, DESC_K_HI = (1024u >> 4) | (1u << 14) | (2u << 29);          // K-major SWIZZLE_128B
__device__ __forceinline__ void tm_prefetch_map(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tm_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor (version 1); layout 2 = SWIZZLE_128B (16 B chunks), 1 = SWIZZLE_128B_BASE32B (32 B
// chunks, 4-row atoms: the only layout the tensor core accepts for an MN-major tf32 operand)
__device__ __forceinline__ uint64_t tm_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// remainder of the tensor core's own tf32 reading of x (it drops the 13 low mantissa bits), ROUNDED to tf32: the hardware
// then reads the lo operand exactly (a truncated 13-bit remainder would lose its 2 low bits, always towards zero)
__device__ __forceinline__ float tf32_rest(float x) {
    const float r = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(r));
    return __uint_as_float(u);
}
__device__ __forceinline__ float4 tf32_rest4(const float4 v) {
    return make_float4(tf32_rest(v.x), tf32_rest(v.y), tf32_rest(v.z), tf32_rest(v.w));
}
__device__ __forceinline__ float tm_act(float v, int act, float slope) {
    switch (act) {
        case CCB_ACT_RELU: return fmaxf(v, 0.f);
        case CCB_ACT_LEAKY: return v > 0.f ? v : v * slope;
        case CCB_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}


}  // namespace ccb
#endif
