// umma_probe — validates, on a B200, every tcgen05 convention csrc/mlp_tc.cu relies on, one case per
// process (a faulting case must not poison the next):
//   * shared-memory matrix descriptors for the un-swizzled canonical layouts, K-major and MN-major
//     (field layout: cute/arch/mma_sm100_desc.hpp `SmemDescriptor`), operands written with ordinary
//     st.shared + fence.proxy.async, or brought in by a 1-D bulk copy (cp.async.bulk + mbarrier tx)
//   * the kind::f16 instruction descriptor (M, N, majors, fp32 accumulate)
//   * TMEM accumulator addressing for M = 128 (row i -> lane i) and M = 64 (row i -> lane (i/16)*32 + i%16)
//   * tcgen05.commit -> mbarrier, tcgen05.ld 32x32b
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
// Run:   umma_probe <M> <N> <K> <a_mn> <b_mn> <bulk>        prints "max_err ..." and exits 0 when exact.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

struct Args { int M, N, K, a_mn, b_mn, bulk; const __half* A; const __half* B; float* D; };

// canonical byte offset of logical element (r, k) of an operand with R rows (MN extent) and K columns
__host__ __device__ inline uint32_t canon_off(int r, int k, int R, int K, int mn_major) {
    if (!mn_major) return (uint32_t)((r / 8) * (K / 8) * 128 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2);
    return (uint32_t)((k / 8) * (R / 8) * 128 + (r / 8) * 128 + (k % 8) * 16 + (r % 8) * 2);
}

__global__ void __launch_bounds__(128, 1) probe(Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    unsigned char* sA = smem;
    unsigned char* sB = smem + 32768;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bytesA = (uint32_t)a.M * a.K * 2, bytesB = (uint32_t)a.N * a.K * 2;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (a.bulk) {
        // operands are pre-tiled in global memory (the exact shared-memory image)
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[1])), "r"(bytesA + bytesB) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sA)), "l"(a.A), "r"(bytesA), "r"(smem_u32(&mbar[1])) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sB)), "l"(a.B), "r"(bytesB), "r"(smem_u32(&mbar[1])) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&mbar[1])), "r"(0) : "memory");
    } else {
        for (int i = tid; i < a.M * a.K; i += blockDim.x) {
            const int r = i / a.K, k = i % a.K;
            *reinterpret_cast<__half*>(sA + canon_off(r, k, a.M, a.K, a.a_mn)) = a.A[i];
        }
        for (int i = tid; i < a.N * a.K; i += blockDim.x) {
            const int r = i / a.K, k = i % a.K;
            *reinterpret_cast<__half*>(sB + canon_off(r, k, a.N, a.K, a.b_mn)) = a.B[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (tid == 0) {
        const uint32_t idesc = make_idesc(a.M, a.N, a.a_mn, a.b_mn);
        // K-major: LBO = 128 (next 8 k), SBO = K/8*128 (next 8 rows).  MN-major: SBO = 128 (next 8 mn), LBO = R/8*128 (next 8 k)
        const uint32_t lboA = a.a_mn ? (uint32_t)(a.M / 8) * 128 : 128, sboA = a.a_mn ? 128 : (uint32_t)(a.K / 8) * 128;
        const uint32_t lboB = a.b_mn ? (uint32_t)(a.N / 8) * 128 : 128, sboB = a.b_mn ? 128 : (uint32_t)(a.K / 8) * 128;
        for (int k = 0; k < a.K / 16; ++k) {
            const uint64_t da = make_desc(smem_u32(sA) + k * 2 * lboA, lboA, sboA);
            const uint64_t db = make_desc(smem_u32(sB) + k * 2 * lboB, lboB, sboB);
            const uint32_t acc = k > 0;
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[0])) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&mbar[0])), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // read the accumulator: this warp owns TMEM lanes [32*warp, 32*warp+32)
    int row;
    bool valid;
    if (a.M == 128) { row = warp * 32 + lane; valid = true; }
    else { row = warp * 16 + lane; valid = lane < 16; }
    for (int c = 0; c < a.N; c += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid)
            for (int j = 0; j < 8; ++j) a.D[(size_t)row * a.N + c + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main(int argc, char** argv) {
    if (argc < 7) { printf("usage: umma_probe M N K a_mn b_mn bulk\n"); return 1; }
    Args a;
    a.M = atoi(argv[1]); a.N = atoi(argv[2]); a.K = atoi(argv[3]); a.a_mn = atoi(argv[4]); a.b_mn = atoi(argv[5]); a.bulk = atoi(argv[6]);
    std::vector<__half> hA((size_t)a.M * a.K), hB((size_t)a.N * a.K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    srand(1234);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 17) - 8) / 8.0f; hA[i] = __float2half(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 13) - 6) / 4.0f; hB[i] = __float2half(fB[i]); }
    std::vector<__half> gA = hA, gB = hB;
    if (a.bulk) {   // pre-tile on the host
        for (int r = 0; r < a.M; ++r) for (int k = 0; k < a.K; ++k) gA[canon_off(r, k, a.M, a.K, a.a_mn) / 2] = hA[(size_t)r * a.K + k];
        for (int r = 0; r < a.N; ++r) for (int k = 0; k < a.K; ++k) gB[canon_off(r, k, a.N, a.K, a.b_mn) / 2] = hB[(size_t)r * a.K + k];
    }
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, gA.size() * 2)); CK(cudaMalloc(&dB, gB.size() * 2)); CK(cudaMalloc(&dD, (size_t)a.M * a.N * 4));
    CK(cudaMemcpy(dA, gA.data(), gA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, gB.data(), gB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, (size_t)a.M * a.N * 4));
    a.A = dA; a.B = dB; a.D = dD;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    probe<<<1, 128, 65536>>>(a);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> hD((size_t)a.M * a.N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0; int bad = 0;
    for (int m = 0; m < a.M; ++m)
        for (int n = 0; n < a.N; ++n) {
            double ref = 0;
            for (int k = 0; k < a.K; ++k) ref += (double)fA[(size_t)m * a.K + k] * fB[(size_t)n * a.K + k];
            const double e = fabs(ref - (double)hD[(size_t)m * a.N + n]);
            if (!(e <= 1e-3)) { if (bad < 4) printf("  mismatch (%d,%d): got %f want %f\n", m, n, hD[(size_t)m * a.N + n], ref); ++bad; }
            if (e > max_err) max_err = e;
        }
    printf("M=%d N=%d K=%d a_mn=%d b_mn=%d bulk=%d  max_err %.3g  bad %d  %s\n", a.M, a.N, a.K, a.a_mn, a.b_mn, a.bulk, max_err, bad, bad ? "FAIL" : "OK");
    return bad ? 3 : 0;
}
