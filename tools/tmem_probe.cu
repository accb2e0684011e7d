// TMEM read throughput and tcgen05 round-trip latency on the device this runs on (B200):
//   (1) tcgen05.ld.32x32b.x16 / .x32 issued by 4, 8 or 16 warps of one CTA, 1 / 2 / 4 loads in flight per wait
//       -> bytes per clock per SM (the epilogue bound of the fused MLP kernels);
//   (2) the latency of  issue (one 128x128x16 MMA) -> commit -> mbarrier wait -> first tcgen05.ld  seen by a warp
//       -> the fixed cost of one phase of k_mlp_bwd_tc.
// Build: make -C tools tmem_probe ; run on the GPU box: tools/tmem_probe
#include "tc_common.cuh"
#include <cstdio>
#include <vector>

unsigned long long g_al_launches = 0;
void al_set_error(const char*, ...) {}
int al_num_sms() { return 148; }

using namespace tc;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// MODE 16: x16 loads, MODE 32: x32 loads.  INFL loads in flight per wait.  Every warp reads `iters * INFL` loads from its
// own lane quarter (warp % 4), column windows rotating over the 512 columns.
template <int MODE, int INFL>
__global__ void k_ld(int iters, unsigned long long* clocks, uint32_t* sink) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    int col = (warp >> 2) * MODE;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 16) {
            uint32_t v[INFL][16];
            #pragma unroll
            for (int q = 0; q < INFL; ++q) { tmem_ld16(tmem + lane_sel + ((col + q * 16) & 511 & ~15), v[q]); }
            tmem_ld_wait();
            #pragma unroll
            for (int q = 0; q < INFL; ++q)
                #pragma unroll
                for (int j = 0; j < 16; ++j) acc ^= v[q][j];
            col += INFL * 16;
        } else {
            uint32_t v[INFL][32];
            #pragma unroll
            for (int q = 0; q < INFL; ++q) { tmem_ld32(tmem + lane_sel + ((col + q * 32) & 511 & ~31), v[q]); }
            tmem_ld_wait();
            #pragma unroll
            for (int q = 0; q < INFL; ++q)
                #pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= v[q][j];
            col += INFL * 32;
        }
    }
    __syncthreads();
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// One phase of the fused MLP backward, stripped: thread 0 issues a 128 x N x K GEMM on zeroed smem operands and commits;
// all threads wait on the mbarrier, read NCOL accumulator columns (their share) and meet at a __syncthreads.
template <int N, int K, bool A_MN = false, bool B_MN = false, int REPEAT = 1>
__global__ void k_phase(int iters, unsigned long long* clocks, uint32_t* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int i = tid; i < (128 * K * 2 + N * K * 2) / 4; i += nthr) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t aA = smem_u32(smem), aB = aA + 128 * K * 2, b = smem_u32(&bar);
    const int warp = tid >> 5, parts = nthr / 128, part = tid >> 7;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t par = 0, acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (tid == 0) {
            tc_fence_after();
            #pragma unroll
            for (int rep = 0; rep < REPEAT; ++rep)
                issue_gemm<128, N, K, A_MN, B_MN>(tmem + rep * N, A_MN ? view_mn(aA, 128) : view_k(aA, K),
                                                  B_MN ? view_mn(aB, N) : view_k(aB, K), false);
            mma_commit(b);
        }
        mbar_wait(b, par); par ^= 1;
        tc_fence_after();
        for (int c = part * (N / parts); c < (part + 1) * (N / parts); c += 32) {
            uint32_t v0[16], v1[16];
            tmem_ld16(tmem + lane_sel + c, v0);
            tmem_ld16(tmem + lane_sel + c + 16, v1);
            tmem_ld_wait();
            #pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= v0[j] ^ v1[j];
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
    }
    const unsigned long long t1 = clock64();
    if (tid == 0) clocks[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int MODE, int INFL>
void run_ld(int warps, int ctas, unsigned long long* d_clk, uint32_t* d_sink) {
    const int iters = 2000;
    k_ld<MODE, INFL><<<ctas, warps * 32>>>(iters, d_clk, d_sink);
    cudaDeviceSynchronize();
    std::vector<unsigned long long> h(ctas);
    cudaMemcpy(h.data(), d_clk, ctas * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    const double bytes = (double)iters * INFL * warps * 32.0 * MODE * 4.0;
    printf("ld x%-2d  warps %2d  in flight %d  ctas %3d : %8llu clk  %7.1f B/clk/SM  (%s)\n", MODE, warps, INFL, ctas, mx, bytes / mx,
           cudaGetErrorString(cudaGetLastError()));
}

template <int N, int K, bool A_MN = false, bool B_MN = false, int REPEAT = 1>
void run_phase(int threads, unsigned long long* d_clk, uint32_t* d_sink) {
    const int iters = 1000;
    const int smem = 128 * K * 2 + N * K * 2;
    cudaFuncSetAttribute(k_phase<N, K, A_MN, B_MN, REPEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_phase<N, K, A_MN, B_MN, REPEAT><<<1, threads, smem>>>(iters, d_clk, d_sink);
    cudaDeviceSynchronize();
    unsigned long long h = 0;
    cudaMemcpy(&h, d_clk, sizeof h, cudaMemcpyDeviceToHost);
    printf("phase 128x%dx%d x%d A_%s B_%s threads %3d : %7.1f clk per phase (MMA floor %d clk)  (%s)\n", N, K, REPEAT, A_MN ? "mn" : "k ",
           B_MN ? "mn" : "k ", threads, (double)h / iters, REPEAT * 128 * N / 256 * (K / 16), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    unsigned long long* d_clk;
    uint32_t* d_sink;
    cudaMalloc(&d_clk, 256 * sizeof(unsigned long long));
    cudaMalloc(&d_sink, 16);
    for (int ctas : {148}) {
        for (int warps : {4, 8, 16}) {
            run_ld<16, 1>(warps, ctas, d_clk, d_sink);
            run_ld<16, 2>(warps, ctas, d_clk, d_sink);
            run_ld<16, 4>(warps, ctas, d_clk, d_sink);
            run_ld<32, 1>(warps, ctas, d_clk, d_sink);
            run_ld<32, 2>(warps, ctas, d_clk, d_sink);
        }
    }
    run_phase<128, 128>(512, d_clk, d_sink);
    run_phase<128, 128>(256, d_clk, d_sink);
    run_phase<128, 128>(128, d_clk, d_sink);
    run_phase<128, 16>(512, d_clk, d_sink);
    run_phase<64, 64>(512, d_clk, d_sink);
    run_phase<64, 64>(256, d_clk, d_sink);
    run_phase<32, 16>(512, d_clk, d_sink);
    // operand majors of the backward GEMMs: dgrad = A K-major, B MN-major; wgrad = both MN-major; REPEAT = GEMMs per commit
    run_phase<128, 128, false, true>(512, d_clk, d_sink);
    run_phase<128, 128, true, true>(512, d_clk, d_sink);
    run_phase<128, 128, true, false>(512, d_clk, d_sink);
    run_phase<128, 128, false, false, 2>(512, d_clk, d_sink);
    run_phase<128, 128, false, true, 2>(512, d_clk, d_sink);
    run_phase<128, 128, true, true, 2>(512, d_clk, d_sink);
    run_phase<128, 128, false, false, 4>(512, d_clk, d_sink);
    run_phase<128, 128, true, true, 4>(512, d_clk, d_sink);
    run_phase<48, 128, true, true, 2>(512, d_clk, d_sink);
    run_phase<16, 128, true, true, 2>(512, d_clk, d_sink);
    run_phase<64, 128, true, true, 2>(512, d_clk, d_sink);
    return 0;
}
