// Fully fused bias-free MLP on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM):
// forward, and a backward that recomputes the hidden activations, back-propagates through all
// layers and keeps the weight gradients in TMEM for the whole lifetime of a persistent CTA.
//
// Same contract as mlp.cu (the mma.sync back end, kept as the measured baseline); replaces
// tiny-cuda-nn's `tcnn.Network` as used by autolabel/models.py:84-136, design reference
// torch_ngp/ffmlp/src/ffmlp.cu:331-518 (wmma forward/backward + CUTLASS split-K weight gradients).
//
// Layout
//  * Every matrix operand lives in shared memory in the un-swizzled canonical UMMA layout: a tile of
//    R rows x C columns of fp16 is a grid of 8x8 "core matrices", each 128 contiguous bytes,
//        byte(r, c) = (r/8) * (C/8) * 128 + (c/8) * 128 + (r%8) * 16 + (c%8) * 2.
//    The SAME bytes serve as a K-major operand (MN = row, K = column: LBO 128, SBO (C/8)*128) and as an
//    MN-major operand (MN = column, K = row: SBO 128, LBO (C/8)*128), so one copy of each weight matrix
//    feeds forward (W, K-major) and dgrad (W^T, MN-major), and one copy of each activation / gradient
//    tile feeds the next layer (K-major, samples = M) and the weight gradient (MN-major, samples = K).
//    tools/umma_probe.cu checks each of these views against a CPU GEMM on the device.
//  * A tile is 128 samples = the 128 TMEM lanes; sample i of the tile is accumulator row i.  A thread
//    owns one sample row: it reads its TMEM lane with tcgen05.ld.32x32b, applies ReLU (or the ReLU mask of
//    the recomputed activation), packs to fp16 and writes its row of the next operand tile; 8 lanes
//    of a warp cover one 128-byte core matrix per store (bank-conflict free).
//  * Forward: G independent 128-thread groups per persistent CTA, each running
//    load -> MMA -> epilogue -> MMA ... on its own tiles and TMEM columns, so the tensor pipe works on
//    one group's layer while the others run their epilogues.
//  * Backward: one group of 256 threads (two column halves per TMEM lane quarter).  Per tile:
//    fwd1, fwd2 (recompute), dgrad chain, and the three weight-gradient GEMMs
//        dW1 [H x IN] += dH1^T A0,   dW2 [H x H] += dH2^T A1,   dWo^T [H x OUT] += A2^T dOut
//    with the 128 samples as the MMA K dimension, issued right behind the dgrad GEMM whose epilogue they
//    overlap.  dW accumulators stay in TMEM across ALL tiles of the CTA; one red.global.add per weight
//    per CTA at the end.  No activation or dH tensor is written to HBM.
#include "common.cuh"
#include "mlp_args.cuh"

namespace {

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

// D[M x N] (+)= A[M x K] B[N x K]^T, K/16 instructions, issued by ONE thread.
template <int M, int N, int K, bool A_MN, bool B_MN>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, const Operand a, const Operand b, bool accumulate) {
    static_assert(M == 64 || M == 128, "UMMA M");
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N");
    static_assert(K % 16 == 0, "UMMA K");
    constexpr uint32_t idesc = make_idesc(M, N, A_MN, B_MN);
    #pragma unroll
    for (int k = 0; k < K / 16; ++k) {
        const uint64_t da = make_desc(a.addr + k * a.kstep, a.lbo, a.sbo);
        const uint64_t db = make_desc(b.addr + k * b.kstep, b.lbo, b.sbo);
        mma_f16(tmem_d, da, db, idesc, (accumulate || k > 0) ? 1u : 0u);
    }
}

// ------------------------------------------------------------------------------------------ staging
// fp32 global row-major [R][C] -> fp16 canonical tile.
__device__ __forceinline__ void stage_weights(const float* __restrict__ g, unsigned char* s, int R, int C, int tid,
                                              int nthreads) {
    const int chunks = C / 8;
    for (int i = tid; i < R * chunks; i += nthreads) {
        const int r = i / chunks, ch = i - r * chunks;
        const float4 a = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * C + ch * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * C + ch * 8 + 4));
        uint4 v;
        v.x = pack_h2(a.x, a.y); v.y = pack_h2(a.z, a.w); v.z = pack_h2(b.x, b.y); v.w = pack_h2(b.z, b.w);
        *reinterpret_cast<uint4*>(s + (r >> 3) * chunks * 128 + ch * 128 + (r & 7) * 16) = v;
    }
}

// This thread's share of one input row (fp16, global): chunks [c0, c0 + NCH) of 8 halfs; zeros for rows >= n.
template <int NCH>
__device__ __forceinline__ void load_row(const __half* __restrict__ x, size_t ldx, long long row, long long n, int c0,
                                         uint4 (&v)[NCH]) {
    #pragma unroll
    for (int j = 0; j < NCH; ++j) v[j] = make_uint4(0, 0, 0, 0);
    if (row < n) {
        const uint4* p = reinterpret_cast<const uint4*>(x + (size_t)row * ldx) + c0;
        #pragma unroll
        for (int j = 0; j < NCH; ++j) v[j] = __ldg(p + j);
    }
}
template <int NCH>
__device__ __forceinline__ void store_row(unsigned char* tile, int C, int r, int c0, const uint4 (&v)[NCH]) {
    unsigned char* p = tile + (r >> 3) * (C / 8) * 128 + (r & 7) * 16 + c0 * 128;
    #pragma unroll
    for (int j = 0; j < NCH; ++j) *reinterpret_cast<uint4*>(p + j * 128) = v[j];
}

// TMEM accumulator columns [c_begin, c_end) of this thread's lane -> (ReLU | mask) -> fp16 -> canonical tile row r.
//   MODE 0: relu(acc);  MODE 1: acc where mask_tile(r, c) > 0 else 0 (mask_tile: fp16 canonical, same C).
template <int C, int MODE>
__device__ __forceinline__ void epi_to_tile(uint32_t taddr_lane, int c_begin, int c_end, unsigned char* dst,
                                            const unsigned char* mask_tile, int r) {
    const uint32_t row_off = (r >> 3) * (C / 8) * 128 + (r & 7) * 16;
    for (int c = c_begin; c < c_end; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr_lane + c, v);
        tmem_ld_wait();
        #pragma unroll
        for (int q = 0; q < 2; ++q) {
            float f[8];
            #pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[q * 8 + j]);
            const uint32_t off = row_off + ((c >> 3) + q) * 128;
            if (MODE == 0) {
                #pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            } else {
                const uint4 m = *reinterpret_cast<const uint4*>(mask_tile + off);
                const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&mw[j]));
                    f[2 * j] = a.x > 0.f ? f[2 * j] : 0.f;
                    f[2 * j + 1] = a.y > 0.f ? f[2 * j + 1] : 0.f;
                }
            }
            uint4 o;
            o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]); o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
            *reinterpret_cast<uint4*>(dst + off) = o;
        }
    }
}

template <int IN, int H, int OUT, int NH>
struct Shape {
    static constexpr int W1 = 0;                                   // [H][IN]   (flat fp32 parameter offsets)
    static constexpr int W2 = W1 + H * IN;                         // [H][H]
    static constexpr int WO = W2 + (NH == 2 ? H * H : 0);          // [OUT][H]
    static constexpr int NPARAMS = WO + OUT * H;
    static constexpr uint32_t bW1 = H * IN * 2, bW2 = NH == 2 ? H * H * 2 : 0, bWO = OUT * H * 2;
    static constexpr uint32_t bX = 128 * IN * 2, bH = 128 * H * 2, bO = 128 * OUT * 2;
};

// ========================================================================================== forward
template <int IN, int H, int OUT, int NH, int G>
struct FwdCfg {
    using S = Shape<IN, H, OUT, NH>;
    static constexpr uint32_t oW1 = 0, oW2 = oW1 + S::bW1, oWO = oW2 + S::bW2, oGrp = oWO + S::bWO;
    static constexpr uint32_t bGrp = S::bX + S::bH;
    static constexpr uint32_t oBar = oGrp + G * bGrp;
    static constexpr uint32_t BYTES = oBar + 8 * G + 16;
    static constexpr int TCOLS = H + OUT;                          // TMEM columns per group
    static_assert(G * TCOLS <= 512, "TMEM columns");
    static_assert(BYTES <= 227 * 1024, "shared memory");
};

template <int IN, int H, int OUT, int NH, int G>
__global__ void __launch_bounds__(G * 128, 1) k_mlp_fwd_tc(const MlpFwdArgs args) {
    using S = Shape<IN, H, OUT, NH>;
    using C = FwdCfg<IN, H, OUT, NH, G>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x;
    const int g = tid >> 7, tg = tid & 127, wq = (tid >> 5) & 3, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::oBar + 8 * G);

    stage_weights(args.params + S::W1, smem + C::oW1, H, IN, tid, G * 128);
    if (NH == 2) stage_weights(args.params + S::W2, smem + C::oW2, H, H, tid, G * 128);
    stage_weights(args.params + S::WO, smem + C::oWO, OUT, H, tid, G * 128);
    if (tid == 0) {
        for (int i = 0; i < G; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_h = tmem + g * C::TCOLS, t_o = t_h + H;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;

    unsigned char* sX = smem + C::oGrp + g * C::bGrp;
    unsigned char* sH = sX + S::bX;
    const uint32_t aW1 = smem_u32(smem + C::oW1), aW2 = smem_u32(smem + C::oW2), aWO = smem_u32(smem + C::oWO);
    const uint32_t aX = smem_u32(sX), aH = smem_u32(sH), bar = smem_u32(&bars[g]);
    const int r = tg;                                              // tile row == TMEM lane == sample
    const long long n = args.n_dev ? min((long long)args.cap, (long long)*args.n_dev) : (long long)args.cap;
    const long long n_tiles = (n + 127) / 128;
    uint32_t parity = 0;

    uint4 xr[IN / 8];
    long long tile = (long long)blockIdx.x * G + g;
    if (tile < n_tiles) load_row<IN / 8>(args.x, args.ldx, tile * 128 + r, n, 0, xr);
    for (; tile < n_tiles; tile += (long long)gridDim.x * G) {
        const long long row = tile * 128 + r;
        store_row<IN / 8>(sX, IN, r, 0, xr);
        {   // prefetch the next tile's row while this one goes through the layers
            const long long nt = tile + (long long)gridDim.x * G;
            if (nt < n_tiles) load_row<IN / 8>(args.x, args.ldx, nt * 128 + r, n, 0, xr);
        }
        fence_async_smem();
        tc_fence_before();
        named_bar(1 + g, 128);
        if (tg == 0) {
            tc_fence_after();
            issue_gemm<128, H, IN, false, false>(t_h, view_k(aX, IN), view_k(aW1, IN), false);
            mma_commit(bar);
        }
        mbar_wait(bar, parity); parity ^= 1;
        tc_fence_after();
        epi_to_tile<H, 0>(t_h + lane_sel, 0, H, sH, nullptr, r);
        fence_async_smem();
        tc_fence_before();
        named_bar(1 + g, 128);
        if (NH == 2) {
            if (tg == 0) {
                tc_fence_after();
                issue_gemm<128, H, H, false, false>(t_h, view_k(aH, H), view_k(aW2, H), false);
                mma_commit(bar);
            }
            mbar_wait(bar, parity); parity ^= 1;
            tc_fence_after();
            epi_to_tile<H, 0>(t_h + lane_sel, 0, H, sH, nullptr, r);
            fence_async_smem();
            tc_fence_before();
            named_bar(1 + g, 128);
        }
        if (tg == 0) {
            tc_fence_after();
            issue_gemm<128, OUT, H, false, false>(t_o, view_k(aH, H), view_k(aWO, H), false);
            mma_commit(bar);
        }
        mbar_wait(bar, parity); parity ^= 1;
        tc_fence_after();
        // output epilogue: this thread's row of y -> the three output windows
        #pragma unroll
        for (int c = 0; c < OUT; c += 16) {
            uint32_t v[16];
            tmem_ld16(t_o + lane_sel + c, v);
            tmem_ld_wait();
            if (row < n) {
                #pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float y = __uint_as_float(v[j]);
                    const int col = c + j;
                    if (args.o0.ptr) {
                        const int rel = col - args.o0.src0;
                        if (rel >= 0 && rel < args.o0.ncols)
                            args.o0.ptr[(size_t)row * args.o0.ld + args.o0.col0 + rel] = al_apply_act(y, args.o0.act);
                    }
                    if (args.o1.ptr) {
                        const int rel = col - args.o1.src0;
                        if (rel >= 0 && rel < args.o1.ncols)
                            args.o1.ptr[(size_t)row * args.o1.ld + args.o1.col0 + rel] = al_apply_act(y, args.o1.act);
                    }
                    if (args.h0.ptr) {
                        const int rel = col - args.h0.src0;
                        if (rel >= 0 && rel < args.h0.ncols)
                            args.h0.ptr[(size_t)row * args.h0.ld + args.h0.col0 + rel] =
                                __float2half_rn(args.h0.act == 1 ? fmaxf(y, 0.f) : y);
                    }
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ========================================================================================== backward
template <int IN, int H, int OUT, int NH>
struct BwdCfg {
    using S = Shape<IN, H, OUT, NH>;
    static constexpr uint32_t oW1 = 0, oW2 = oW1 + S::bW1, oWO = oW2 + S::bW2;
    static constexpr uint32_t oA0 = oWO + S::bWO;                  // [128][IN]  layer-1 input
    static constexpr uint32_t oA1 = oA0 + S::bX;                   // [128][H]   relu(h1)
    static constexpr uint32_t oA2 = oA1 + S::bH;                   // [128][H]   relu(h2), later d h1      (NH == 2)
    static constexpr uint32_t oDL = oA2 + (NH == 2 ? S::bH : 0);   // [128][H]   d h_last
    static constexpr uint32_t oDO = oDL + S::bH;                   // [128][OUT] d out (scaled, fp16)
    static constexpr uint32_t oBar = oDO + S::bO;
    static constexpr uint32_t BYTES = oBar + 16 + 16;
    // TMEM columns
    static constexpr int tACC = 0, tDX = tACC + H, tW1 = tDX + IN, tW2 = tW1 + IN, tWO = tW2 + (NH == 2 ? H : 0);
    static constexpr int TCOLS = tWO + OUT;
    static_assert(TCOLS <= 512, "TMEM columns");
    static_assert(BYTES <= 227 * 1024, "shared memory");
};

constexpr int kBwdThreads = 256;

// dW accumulator [MW x NW] in TMEM (row i -> lane i for MW = 128, lane (i/16)*32 + i%16 for MW = 64)
// -> red.global.add into dW, element (row, col) at dW[row * s_row + col * s_col].
template <int MW, int NW>
__device__ __forceinline__ void flush_dw(uint32_t taddr, float* __restrict__ dW, int s_row, int s_col, float inv_scale,
                                         int wq, int half, int lane) {
    const int row = MW == 128 ? wq * 32 + lane : wq * 16 + lane;
    const bool valid = MW == 128 || lane < 16;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    constexpr int NCH = NW / 16;                                   // 16-column chunks, split over the two halves
    for (int ch = half; ch < NCH; ch += 2) {
        uint32_t v[16];
        tmem_ld16(taddr + lane_sel + ch * 16, v);
        tmem_ld_wait();
        if (valid) {
            #pragma unroll
            for (int j = 0; j < 16; ++j)
                atomicAdd(dW + (size_t)row * s_row + (size_t)(ch * 16 + j) * s_col, __uint_as_float(v[j]) * inv_scale);
        }
    }
}

template <int IN, int H, int OUT, int NH>
__global__ void __launch_bounds__(kBwdThreads, 1) k_mlp_bwd_tc(const MlpBwdArgs args) {
    using S = Shape<IN, H, OUT, NH>;
    using C = BwdCfg<IN, H, OUT, NH>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x;
    const int wq = (tid >> 5) & 3, half = tid >> 7, lane = tid & 31;
    const int r = wq * 32 + lane;                                  // tile row == TMEM lane == sample
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::oBar + 16);

    stage_weights(args.params + S::W1, smem + C::oW1, H, IN, tid, kBwdThreads);
    if (NH == 2) stage_weights(args.params + S::W2, smem + C::oW2, H, H, tid, kBwdThreads);
    stage_weights(args.params + S::WO, smem + C::oWO, OUT, H, tid, kBwdThreads);
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    const uint32_t aW1 = smem_u32(smem + C::oW1), aW2 = smem_u32(smem + C::oW2), aWO = smem_u32(smem + C::oWO);
    unsigned char* sA0 = smem + C::oA0;
    unsigned char* sA1 = smem + C::oA1;
    unsigned char* sA2 = smem + C::oA2;
    unsigned char* sDL = smem + C::oDL;
    unsigned char* sDO = smem + C::oDO;
    unsigned char* sD1 = NH == 2 ? sA2 : sDL;                      // d h1 (first hidden layer's gradient)
    const uint32_t aA0 = smem_u32(sA0), aA1 = smem_u32(sA1), aA2 = smem_u32(sA2), aDL = smem_u32(sDL),
                   aDO = smem_u32(sDO), aD1 = smem_u32(sD1);
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
    uint32_t par0 = 0, par1 = 0;

    const long long n = args.n_dev ? min((long long)args.cap, (long long)*args.n_dev) : (long long)args.cap;
    const long long n_tiles = (n + 127) / 128;
    const float scale = al_grad_scale(args.amax_dev);
    const float inv_scale = 1.0f / scale;

    // this thread's share of a row: input chunks [xc0, xc0 + XCH), d-out columns [dc0, dc0 + DCH * 8)
    constexpr int XCH = (IN / 8 + 1) / 2;                          // chunks of the first half (second may have fewer)
    constexpr int DCH = OUT / 16;                                  // 8-column chunks per half
    const int xc0 = half * XCH;
    const int xn = half == 0 ? XCH : IN / 8 - XCH;
    const int dc0 = half * DCH * 8;

    uint4 xr[XCH];
    float dr[DCH * 8];
    auto prefetch = [&](long long t) {
        const long long row = t * 128 + r;
        #pragma unroll
        for (int j = 0; j < XCH; ++j) xr[j] = make_uint4(0, 0, 0, 0);
        #pragma unroll
        for (int j = 0; j < DCH * 8; ++j) dr[j] = 0.f;
        if (row < n) {
            const uint4* p = reinterpret_cast<const uint4*>(args.x + (size_t)row * args.ldx) + xc0;
            #pragma unroll
            for (int j = 0; j < XCH; ++j)
                if (j < xn) xr[j] = __ldg(p + j);
            const DoutSpec& sp = args.spec;
            if (sp.kind == 0) {
                const float* d = args.dout + (size_t)row * args.ld_dout + args.dcol0;
                #pragma unroll
                for (int j = 0; j < DCH * 8; ++j)
                    if (dc0 + j < args.dncols) dr[j] = __ldg(d + dc0 + j);
            } else {
                // G(row, c): materialised or rank-1 (see DoutSpec)
                const float* grow;
                float wrow = 1.0f;
                if (sp.w) { wrow = __ldg(sp.w + row); grow = sp.g_out + (size_t)__ldg(sp.sray + row) * sp.K; }
                else grow = sp.g_vals + (size_t)row * sp.ldg + 1;
                const float* vrow = sp.vals + (size_t)row * sp.ldv;
                if (sp.kind == 1) {
                    #pragma unroll
                    for (int j = 0; j < DCH * 8; ++j)
                        if (dc0 + j < sp.C) dr[j] = wrow * __ldg(grow + 3 + dc0 + j);
                } else if (sp.kind == 2) {
                    const float* gf = grow + 3 + sp.C;
                    const float* ft = vrow + 4 + sp.C;
                    const float* ds = sp.d_semo_in + (size_t)row * sp.ld_semo;
                    #pragma unroll
                    for (int j = 0; j < DCH * 8; ++j)
                        if (dc0 + j < sp.F) {
                            float v = wrow * __ldg(gf + dc0 + j);
                            if (__ldg(ft + dc0 + j) > 0.f) v += __ldg(ds + dc0 + j);
                            dr[j] = v;
                        }
                } else if (sp.kind == 3) {
                    #pragma unroll
                    for (int j = 0; j < DCH * 8; ++j)
                        if (dc0 + j < 3) {
                            const float rgb = __ldg(vrow + 1 + dc0 + j);
                            dr[j] = wrow * __ldg(grow + dc0 + j) * rgb * (1.0f - rgb);
                        }
                } else {
                    const float* ds = sp.d_semo_in + (size_t)row * sp.ld_semo + sp.F;
                    const float* d1 = sp.dgeo_semf + (size_t)row * 16;
                    const float* d2 = sp.dgeo_color + (size_t)row * 16;
                    #pragma unroll
                    for (int j = 0; j < DCH * 8; ++j) {
                        const int c = dc0 + j;
                        if (c == 0) {
                            const float gs = sp.w ? __ldg(sp.g_sigma + row) : __ldg(sp.g_vals + (size_t)row * sp.ldg);
                            dr[j] = gs * __expf(fminf(fmaxf(__ldg(sp.h16 + (size_t)row * 16), -15.f), 15.f));
                        } else if (c < 16) {
                            dr[j] = __ldg(ds + c - 1) + __ldg(d1 + c - 1) + __ldg(d2 + c - 1);
                        }
                    }
                }
            }
        }
    };

    bool any = false, pending = false;
    long long tile = blockIdx.x;
    if (tile < n_tiles) prefetch(tile);
    for (; tile < n_tiles; tile += gridDim.x) {
        const long long row = tile * 128 + r;
        if (pending) { mbar_wait(bar1, par1); par1 ^= 1; pending = false; }   // last tile's weight-gradient GEMMs done
        // ---- stage A0 and d-out
        {
            unsigned char* p = sA0 + (r >> 3) * (IN / 8) * 128 + (r & 7) * 16 + xc0 * 128;
            #pragma unroll
            for (int j = 0; j < XCH; ++j)
                if (j < xn) *reinterpret_cast<uint4*>(p + j * 128) = xr[j];
            unsigned char* q = sDO + (r >> 3) * (OUT / 8) * 128 + (r & 7) * 16 + (dc0 >> 3) * 128;
            #pragma unroll
            for (int j = 0; j < DCH; ++j) {
                float f[8];
                #pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(dr[j * 8 + e] * scale, -65504.f), 65504.f);
                uint4 o;
                o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]); o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
                *reinterpret_cast<uint4*>(q + j * 128) = o;
            }
        }
        {
            const long long nt = tile + gridDim.x;
            if (nt < n_tiles) prefetch(nt);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- forward recompute
        if (tid == 0) {
            tc_fence_after();
            issue_gemm<128, H, IN, false, false>(tmem + C::tACC, view_k(aA0, IN), view_k(aW1, IN), false);
            mma_commit(bar0);
        }
        mbar_wait(bar0, par0); par0 ^= 1;
        tc_fence_after();
        epi_to_tile<H, 0>(tmem + C::tACC + lane_sel, half * (H / 2), (half + 1) * (H / 2), sA1, nullptr, r);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (NH == 2) {
            if (tid == 0) {
                tc_fence_after();
                issue_gemm<128, H, H, false, false>(tmem + C::tACC, view_k(aA1, H), view_k(aW2, H), false);
                mma_commit(bar0);
            }
            mbar_wait(bar0, par0); par0 ^= 1;
            tc_fence_after();
            epi_to_tile<H, 0>(tmem + C::tACC + lane_sel, half * (H / 2), (half + 1) * (H / 2), sA2, nullptr, r);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
        }
        // ---- d h_last = (d out . Wo) * relu'(a_last);  dWo^T += a_last^T d out
        const uint32_t aAL = NH == 2 ? aA2 : aA1;
        unsigned char* sAL = NH == 2 ? sA2 : sA1;
        if (tid == 0) {
            tc_fence_after();
            issue_gemm<128, H, OUT, false, true>(tmem + C::tACC, view_k(aDO, OUT), view_mn(aWO, H), false);
            mma_commit(bar0);
            issue_gemm<H, OUT, 128, true, true>(tmem + C::tWO, view_mn(aAL, H), view_mn(aDO, OUT), any);
        }
        mbar_wait(bar0, par0); par0 ^= 1;
        tc_fence_after();
        epi_to_tile<H, 1>(tmem + C::tACC + lane_sel, half * (H / 2), (half + 1) * (H / 2), sDL, sAL, r);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (NH == 2) {
            // ---- d h1 = (d h2 . W2) * relu'(a1);  dW2 += d h2^T a1
            if (tid == 0) {
                tc_fence_after();
                issue_gemm<128, H, H, false, true>(tmem + C::tACC, view_k(aDL, H), view_mn(aW2, H), false);
                mma_commit(bar0);
                issue_gemm<H, H, 128, true, true>(tmem + C::tW2, view_mn(aDL, H), view_mn(aA1, H), any);
            }
            mbar_wait(bar0, par0); par0 ^= 1;     // also covers dWo^T: a2 is free to be overwritten by d h1
            tc_fence_after();
            epi_to_tile<H, 1>(tmem + C::tACC + lane_sel, half * (H / 2), (half + 1) * (H / 2), sD1, sA1, r);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
        }
        // ---- d x = d h1 . W1;  dW1 += d h1^T a0
        if (tid == 0) {
            tc_fence_after();
            if (args.dx) {
                issue_gemm<128, IN, H, false, true>(tmem + C::tDX, view_k(aD1, H), view_mn(aW1, IN), false);
                mma_commit(bar0);
            }
            issue_gemm<H, IN, 128, true, true>(tmem + C::tW1, view_mn(aD1, H), view_mn(aA0, IN), any);
            mma_commit(bar1);
        }
        pending = true;
        any = true;
        if (args.dx) {
            mbar_wait(bar0, par0); par0 ^= 1;
            tc_fence_after();
            constexpr int NCH = IN / 16;
            for (int ch = half; ch < NCH; ch += 2) {
                uint32_t v[16];
                tmem_ld16(tmem + C::tDX + lane_sel + ch * 16, v);
                tmem_ld_wait();
                if (row < n) {
                    #pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const int rel = ch * 16 + j - args.dx_c0;        // dx_c0 is even: a pair never straddles the window
                        if (rel >= 0 && rel < args.dx_n) {
                            const float v0 = __uint_as_float(v[j]) * inv_scale, v1 = __uint_as_float(v[j + 1]) * inv_scale;
                            if (args.dx_mode == 0) {
                                args.dx[(size_t)row * args.ld_dx + rel] = v0;
                                if (rel + 1 < args.dx_n) args.dx[(size_t)row * args.ld_dx + rel + 1] = v1;
                            } else if (rel + 1 < args.dx_n) {
                                *reinterpret_cast<float2*>(args.dx + ((size_t)(rel >> 1) * args.ld_dx + row) * 2) =
                                    make_float2(v0, v1);
                            } else {
                                args.dx[((size_t)(rel >> 1) * args.ld_dx + row) * 2] = v0;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }
    if (pending) { mbar_wait(bar1, par1); par1 ^= 1; }
    tc_fence_after();
    if (any && args.dparams) {
        // dW1 [H][IN]: lanes = out, columns = in;  dW2 [H][H] likewise;  dWo^T: lanes = in (H), columns = out
        flush_dw<H, IN>(tmem + C::tW1, args.dparams + S::W1, IN, 1, inv_scale, wq, half, lane);
        if (NH == 2) flush_dw<H, H>(tmem + C::tW2, args.dparams + S::W2, H, 1, inv_scale, wq, half, lane);
        flush_dw<H, OUT>(tmem + C::tWO, args.dparams + S::WO, 1, H, inv_scale, wq, half, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ------------------------------------------------------------------------------------------ launchers
template <int IN, int H, int OUT, int NH>
int launch_fwd_tc(const MlpFwdArgs& a, cudaStream_t st) {
    constexpr int G = 3;
    using C = FwdCfg<IN, H, OUT, NH, G>;
    static bool configured = false;
    if (!configured) {
        AL_CHECK(cudaFuncSetAttribute(k_mlp_fwd_tc<IN, H, OUT, NH, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES));
        configured = true;
    }
    const long long tiles = ((long long)a.cap + 127) / 128;
    const long long want = (tiles + G - 1) / G;
    const int grid = (int)(want < al_num_sms() ? want : al_num_sms());
    k_mlp_fwd_tc<IN, H, OUT, NH, G><<<grid, G * 128, C::BYTES, st>>>(a);
    AL_LAUNCH_CHECK();
    return 0;
}
template <int IN, int H, int OUT, int NH>
int launch_bwd_tc(const MlpBwdArgs& a, cudaStream_t st) {
    using C = BwdCfg<IN, H, OUT, NH>;
    static bool configured = false;
    if (!configured) {
        AL_CHECK(cudaFuncSetAttribute(k_mlp_bwd_tc<IN, H, OUT, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES));
        configured = true;
    }
    const long long tiles = ((long long)a.cap + 127) / 128;
    const int grid = (int)(tiles < al_num_sms() ? tiles : al_num_sms());
    k_mlp_bwd_tc<IN, H, OUT, NH><<<grid, kBwdThreads, C::BYTES, st>>>(a);
    AL_LAUNCH_CHECK();
    return 0;
}

}  // namespace

#define AL_TC_CONFIGS(X)    \
    X(48, 128, 16, 2)       \
    X(64, 128, 16, 2)       \
    X(32, 128, 16, 2)       \
    X(16, 64, 64, 2)        \
    X(80, 64, 16, 1)        \
    X(64, 64, 16, 2)        \
    X(48, 64, 16, 2)        \
    X(32, 64, 16, 2)        \
    X(144, 64, 16, 1)

int al_tc_mlp_forward(int in_pad, int hidden, int out_pad, int n_hidden, const MlpFwdArgs& a, cudaStream_t st) {
#define X(I, Hh, O, N) \
    if (in_pad == I && hidden == Hh && out_pad == O && n_hidden == N) return launch_fwd_tc<I, Hh, O, N>(a, st);
    AL_TC_CONFIGS(X)
#undef X
    return -1;
}
int al_tc_mlp_backward(int in_pad, int hidden, int out_pad, int n_hidden, const MlpBwdArgs& a, cudaStream_t st) {
    if (a.dx && (a.dx_c0 & 1)) return -1;   // the pair-wise d-x store needs an even window start
#define X(I, Hh, O, N) \
    if (in_pad == I && hidden == Hh && out_pad == O && n_hidden == N) return launch_bwd_tc<I, Hh, O, N>(a, st);
    AL_TC_CONFIGS(X)
#undef X
    return -1;
}
