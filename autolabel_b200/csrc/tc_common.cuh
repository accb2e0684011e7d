// tcgen05 / TMEM / cp.async building blocks shared by the fused MLP kernels (mlp_tc.cu) and the tiled GEMM path
// for wide heads (gemm_tc.cu).  Conventions are verified on the device by tools/umma_probe.cu.
#pragma once
#include "common.cuh"

namespace tc {

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// cute/arch/mma_sm100_desc.hpp `SmemDescriptor`: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version 1 [46,48), layout type 0 (no swizzle) [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// `InstrDescriptor` for kind::f16: fp32 accumulate [4,6) = 1, A/B fp16 [7,13) = 0, a_major [15], b_major [16],
// N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(mbar) : "memory");
}
// Non-blocking phase test (a polling loop over several barriers must not park on one of them).
__device__ __forceinline__ bool mbar_test(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    return done != 0;
}
// Phase test that lets the hardware park the thread for up to ~`ns` nanoseconds (no issue slots burnt while waiting).
__device__ __forceinline__ bool mbar_try_wait_ns(uint32_t mbar, uint32_t parity, uint32_t ns) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mbar), "r"(parity), "r"(ns) : "memory");
    return done != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
// 16 columns of this warp's 32 TMEM lanes <- zero
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                 ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------ operand views
struct Operand {
    uint32_t addr, lbo, sbo, kstep;
};
// K-major view of a canonical tile with C columns: MN = tile row, K = tile column.
__device__ __forceinline__ Operand view_k(uint32_t base, int C) { return {base, 128u, (uint32_t)(C / 8) * 128u, 256u}; }
// MN-major view: MN = tile column, K = tile row.
__device__ __forceinline__ Operand view_mn(uint32_t base, int C) {
    return {base, (uint32_t)(C / 8) * 128u, 128u, (uint32_t)(C / 8) * 256u};
}

// D[M x N] (+)= A[M x K] B[N x K]^T, K/16 instructions, issued by ONE thread.  The issuing thread is a serial resource
// (one thread feeds the tensor pipe of the whole CTA): the two descriptors are built once and advanced by a 64-bit add
// per K step (the start-address field holds addr >> 4 in its low 14 bits; a tile never crosses the 256 KB window).
template <int M, int N, int K, bool A_MN, bool B_MN>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, const Operand a, const Operand b, bool accumulate) {
    static_assert(M == 64 || M == 128, "UMMA M");
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N");
    static_assert(K % 16 == 0, "UMMA K");
    constexpr uint32_t idesc = make_idesc(M, N, A_MN, B_MN);
    uint64_t da = make_desc(a.addr, a.lbo, a.sbo), db = make_desc(b.addr, b.lbo, b.sbo);
    const uint64_t sa = (uint64_t)(a.kstep >> 4), sb = (uint64_t)(b.kstep >> 4);
    if (accumulate) {
        #pragma unroll
        for (int k = 0; k < K / 16; ++k) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
            da += sa; db += sb;
        }
    } else {
        mma_f16(tmem_d, da, db, idesc, 0u);
        #pragma unroll
        for (int k = 1; k < K / 16; ++k) {
            da += sa; db += sb;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
        }
    }
}


__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;                                 // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace tc
