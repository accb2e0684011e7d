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
#include "tc_common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {
using namespace tc;


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
// Packed fp16x2 arithmetic: convert first (cvt.rn.f16x2.f32, one instruction per pair), then apply ReLU / the ReLU mask
// on the packed halves.  Bit-identical to "select in fp32, then round": rounding is monotonic and keeps the sign, and a
// zero stays a zero (relu of -0.0 is +0.0 in both forms: max(x, +0) returns +0 for x = -0 on the half2 unit).
template <int C, int MODE>
__device__ __forceinline__ void epi_chunk16(const uint32_t (&v)[16], int c, uint32_t row_off, unsigned char* dst,
                                            const unsigned char* mask_tile) {
    #pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint32_t off = row_off + ((c >> 3) + q) * 128;
        uint32_t h[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = pack_h2(__uint_as_float(v[q * 8 + 2 * j]), __uint_as_float(v[q * 8 + 2 * j + 1]));
        if (MODE == 0) {
            const __half2 z = __float2half2_rn(0.f);
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&h[j]), z);
                h[j] = *reinterpret_cast<const uint32_t*>(&r);
            }
        } else {
            const uint4 m = *reinterpret_cast<const uint4*>(mask_tile + off);
            const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
            const __half2 z = __float2half2_rn(0.f);
            #pragma unroll
            for (int j = 0; j < 4; ++j)       // 0xFFFF per half where the recomputed activation is > 0
                h[j] &= __hgt2_mask(*reinterpret_cast<const __half2*>(&mw[j]), z);
        }
        *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}
// TMEM accumulator columns [c_begin, c_end) of this thread's lane -> (ReLU | mask) -> fp16 -> canonical tile row r.
//   MODE 0: relu(acc);  MODE 1: acc where mask_tile(r, c) > 0 else 0 (mask_tile: fp16 canonical, same C).
// Two 16-column TMEM loads are in flight per wait.
template <int C, int MODE>
__device__ __forceinline__ void epi_to_tile(uint32_t taddr_lane, int c_begin, int c_end, unsigned char* dst,
                                            const unsigned char* mask_tile, int r) {
    const uint32_t row_off = (r >> 3) * (C / 8) * 128 + (r & 7) * 16;
    int c = c_begin;
    for (; c + 32 <= c_end; c += 32) {
        uint32_t v0[16], v1[16];
        tmem_ld16(taddr_lane + c, v0);
        tmem_ld16(taddr_lane + c + 16, v1);
        tmem_ld_wait();
        epi_chunk16<C, MODE>(v0, c, row_off, dst, mask_tile);
        epi_chunk16<C, MODE>(v1, c + 16, row_off, dst, mask_tile);
    }
    for (; c < c_end; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr_lane + c, v);
        tmem_ld_wait();
        epi_chunk16<C, MODE>(v, c, row_off, dst, mask_tile);
    }
}

// ------------------------------------------------------------------------------------------ async staging
// 128 rows x IN halfs of x (row-major, global) -> canonical tile, 16-byte chunks; rows >= n are zero-filled.
template <int IN>
__device__ __forceinline__ void load_x_tile_async(const __half* __restrict__ x, size_t ldx, long long row0, long long n,
                                                  uint32_t tile, int tid, int nthreads) {
    // unit q of the canonical tile sits at byte q * 16 (q = ((r / 8) * CH + ch) * 8 + r % 8): consecutive threads write
    // consecutive 16-byte units (bank-conflict free) and read 8 rows x 64 contiguous bytes per warp (whole sectors)
    constexpr int CH = IN / 8;
    for (int q = tid; q < 128 * CH; q += nthreads) {
        const int t = q >> 3;
        const int rb = t / CH, ch = t - rb * CH;
        const int r = rb * 8 + (q & 7);
        const bool valid = row0 + r < n;
        const __half* src = valid ? x + (size_t)(row0 + r) * ldx + ch * 8 : x;
        cp_async16(tile + (uint32_t)q * 16u, src, valid);
    }
}

// Source windows of a tile, staged one tile ahead with cp.async.
//   unit 4 : [128 rows][ncols] 4-byte elements of a row-major matrix (ld, col0 in elements); staged row stride
//            (ncols | 1) words, consecutive threads on consecutive elements (coalesced)
//   unit 16: [128 rows][ncols] 16-byte chunks (ld = row stride in BYTES, col0 = byte offset, both multiples of 16);
//            staged row stride (ncols | 1) * 16 bytes
// The odd strides make "thread r reads row r" bank-conflict free.  Rows >= n are zero-filled.
struct Win {
    const void* ptr;
    int ld, col0, ncols, unit;
    int r0, j0, dr, dj;                                            // this thread's first element and its step (tile invariant)
};
__device__ __forceinline__ void win_init(Win& w, int tid, int nthreads) {
    if (!w.ptr || w.ncols <= 0) { w.ptr = nullptr; return; }
    w.r0 = tid / w.ncols; w.j0 = tid - w.r0 * w.ncols;
    w.dr = nthreads / w.ncols; w.dj = nthreads - w.dr * w.ncols;
}
__device__ __forceinline__ void stage_window_async(const Win& w, long long row0, long long n, uint32_t stage) {
    if (!w.ptr) return;
    const int stride = w.ncols | 1;
    int r = w.r0, j = w.j0;
    const int dr = w.dr, dj = w.dj;
    const char* base = reinterpret_cast<const char*>(w.ptr);
    if (w.unit == 16) {
        while (r < 128) {
            const bool valid = row0 + r < n;
            const char* src = valid ? base + (size_t)(row0 + r) * w.ld + w.col0 + j * 16 : base;
            cp_async16(stage + (uint32_t)(r * stride + j) * 16u, src, valid);
            r += dr; j += dj;
            if (j >= w.ncols) { j -= w.ncols; ++r; }
        }
    } else {
        while (r < 128) {
            const bool valid = row0 + r < n;
            const char* src = valid ? base + ((size_t)(row0 + r) * w.ld + w.col0 + j) * 4 : base;
            cp_async4(stage + (uint32_t)(r * stride + j) * 4u, src, valid);
            r += dr; j += dj;
            if (j >= w.ncols) { j -= w.ncols; ++r; }
        }
    }
}

// This warp's staged rows (fp32, row stride `stride` words) -> a column window of a row-major global matrix,
// consecutive lanes on consecutive elements (coalesced); one activation per window; acc: += instead of =.
// Fast path: windows of an even, power-of-two pair count (2, 4, ..., 64 columns) with 8-byte aligned rows go out as one
// 8-byte (fp32) / 4-byte (fp16) vector per lane with shift-only indexing; everything else takes the generic loop.
template <typename T>
__device__ __forceinline__ void write_window(T* __restrict__ ptr, int ld, int col0, int src0, int ncols, int act,
                                             const float* __restrict__ rows, int stride, long long row0, long long n,
                                             int lane, int nrows = 32, int acc = 0) {
    if (!ptr || ncols <= 0) return;
    const long long left = n - row0;
    const int nvalid = left < (long long)nrows ? (int)(left < 0 ? 0 : left) : nrows;
    const int pairs = ncols >> 1;
    if (((ncols | ld | col0) & 1) == 0 && (pairs & (pairs - 1)) == 0 && pairs <= 32 &&
        (reinterpret_cast<uintptr_t>(ptr) & 7u) == 0) {
        const int lg = 31 - __clz(pairs);
        const int rstep = 32 >> lg;
        const int j = (lane & (pairs - 1)) * 2;
        // pointer-stepping loops, the mode decided once outside them (this loop is the whole cost of a wide window)
        int r = lane >> lg;
        const float* src = rows + src0 + j + r * stride;
        T* d = ptr + ((size_t)row0 + r) * ld + col0 + j;
        const int sstep = rstep * stride;
        const size_t dstep = (size_t)rstep * ld;
        if constexpr (sizeof(T) == 4) {
            if (acc) {
                #pragma unroll 4
                for (; r < nvalid; r += rstep, src += sstep, d += dstep)
                    atomicAdd(reinterpret_cast<float2*>(d), make_float2(src[0], src[1]));    // red.global.add.v2.f32
            } else if (act == 0) {
                #pragma unroll 4
                for (; r < nvalid; r += rstep, src += sstep, d += dstep)
                    *reinterpret_cast<float2*>(d) = make_float2(src[0], src[1]);
            } else {
                #pragma unroll 4
                for (; r < nvalid; r += rstep, src += sstep, d += dstep)
                    *reinterpret_cast<float2*>(d) = make_float2(al_apply_act(src[0], act), al_apply_act(src[1], act));
            }
        } else {
            if (act == 1) {
                #pragma unroll 4
                for (; r < nvalid; r += rstep, src += sstep, d += dstep)
                    *reinterpret_cast<__half2*>(d) = __floats2half2_rn(fmaxf(src[0], 0.f), fmaxf(src[1], 0.f));
            } else {
                #pragma unroll 4
                for (; r < nvalid; r += rstep, src += sstep, d += dstep)
                    *reinterpret_cast<__half2*>(d) = __floats2half2_rn(src[0], src[1]);
            }
        }
        return;
    }
    if ((ncols & 31) == 0 && !acc) {
        // wide windows at an odd column (the feature block of vals behind 4 + C columns): lane = column, one row per
        // step, 128-byte coalesced rows
        const float* src = rows + src0 + lane;
        T* d = ptr + (size_t)row0 * ld + col0 + lane;
        #pragma unroll 4
        for (int r = 0; r < nvalid; ++r, src += stride, d += ld) {
            for (int k = 0; k < ncols; k += 32) {
                const float y = src[k];
                if constexpr (sizeof(T) == 4) d[k] = al_apply_act(y, act);
                else d[k] = __float2half_rn(act == 1 ? fmaxf(y, 0.f) : y);
            }
        }
        return;
    }
    if (ncols == 1) {
        if (lane < nvalid) {
            const float y = rows[lane * stride + src0];
            T* d = ptr + (size_t)(row0 + lane) * ld + col0;
            if constexpr (sizeof(T) == 4) {
                if (acc) atomicAdd(d, y);
                else *d = al_apply_act(y, act);
            } else {
                *d = __float2half_rn(act == 1 ? fmaxf(y, 0.f) : y);
            }
        }
        return;
    }
    int r = lane / ncols, j = lane - r * ncols;
    const int dr = 32 / ncols, dj = 32 - dr * ncols;
    while (r < nvalid) {
        const float y = rows[r * stride + src0 + j];
        T* dst = ptr + (size_t)(row0 + r) * ld + col0 + j;
        if constexpr (sizeof(T) == 4) {
            if (acc) atomicAdd(dst, y);                        // red.global.add: no read-back latency
            else *dst = al_apply_act(y, act);
        } else {
            *dst = __float2half_rn(act == 1 ? fmaxf(y, 0.f) : y);
        }
        r += dr; j += dj;
        if (j >= ncols) { j -= ncols; ++r; }
    }
}

template <int IN, int H, int OUT, int NH>
struct Shape {
    static constexpr int W1 = 0;                                   // [H][IN]   (flat fp32 parameter offsets)
    static constexpr int W2 = W1 + H * IN;                         // [H][H]
    static constexpr int WO = W2 + (NH == 2 ? H * H : 0);          // [OUT][H]
    static constexpr uint32_t bW1 = H * IN * 2, bW2 = NH == 2 ? H * H * 2 : 0, bWO = OUT * H * 2;
    static constexpr uint32_t bX = 128 * IN * 2, bH = 128 * H * 2, bO = 128 * OUT * 2;
};

// ========================================================================================== forward
template <int IN, int H, int OUT, int NH, int G>
struct FwdCfg {
    using S = Shape<IN, H, OUT, NH>;
    static constexpr int YS = OUT + 1;                             // staging row stride (words, odd)
    static constexpr uint32_t bY = (128 * YS * 4 + 127) / 128 * 128;
    // The output staging aliases the hidden tile when it fits: the hidden tile is dead once the output layer's MMA has
    // completed, and the group barrier at the top of the tile loop orders the staging reads before the next writes.
    static constexpr bool kAliasY = bY <= S::bH;
    static constexpr uint32_t oW1 = 0, oW2 = oW1 + S::bW1, oWO = oW2 + S::bW2, oGrp = oWO + S::bWO;
    static constexpr uint32_t bGrp = S::bX + S::bH + (kAliasY ? 0 : bY);
    static constexpr uint32_t oBar = oGrp + G * bGrp;
    static constexpr uint32_t BYTES = oBar + 8 * G + 16;
    // TMEM columns per group: the output accumulator reuses the hidden accumulator's columns (free after the last
    // hidden epilogue has read them)
    static constexpr int TCOLS = H > OUT ? H : OUT;
    static constexpr bool kFits = G * TCOLS <= 512 && BYTES <= 227 * 1024;
};
// Independent 128-thread groups per CTA: as many as shared memory and TMEM allow, at most 4 (512 threads).
template <int IN, int H, int OUT, int NH>
constexpr int fwd_groups() {
    return FwdCfg<IN, H, OUT, NH, 4>::kFits ? 4 : (FwdCfg<IN, H, OUT, NH, 3>::kFits ? 3 : 2);
}

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
    const uint32_t t_h = tmem + g * C::TCOLS, t_o = t_h;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;

    unsigned char* sX = smem + C::oGrp + g * C::bGrp;
    unsigned char* sH = sX + S::bX;
    float* sY = reinterpret_cast<float*>(C::kAliasY ? sH : sH + S::bH);
    const uint32_t aW1 = smem_u32(smem + C::oW1), aW2 = smem_u32(smem + C::oW2), aWO = smem_u32(smem + C::oWO);
    const uint32_t aX = smem_u32(sX), aH = smem_u32(sH), bar = smem_u32(&bars[g]);
    const int r = tg;                                              // tile row == TMEM lane == sample
    const long long n = args.n_dev ? min((long long)args.cap, (long long)*args.n_dev) : (long long)args.cap;
    const long long n_tiles = (n + 127) / 128;
    const long long tile_step = (long long)gridDim.x * G;
    uint32_t parity = 0;

    long long tile = (long long)blockIdx.x * G + g;
    if (tile < n_tiles) load_x_tile_async<IN>(args.x, args.ldx, tile * 128, n, aX, tg, 128);
    for (; tile < n_tiles; tile += tile_step) {
        cp_async_wait_all();
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
        // the x tile is free again: bring in the next one while this tile goes through the layers
        if (tile + tile_step < n_tiles) load_x_tile_async<IN>(args.x, args.ldx, (tile + tile_step) * 128, n, aX, tg, 128);
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
        // output epilogue.  The fp16 copy for the next MLP (h0) leaves straight from this thread's registers when its
        // window is 16-column / 16-byte aligned: the thread owns the whole row, so it writes 32-byte sectors of that
        // row (two 16-byte stores per 16 columns).  Everything else: y row -> staging (this warp's 32 rows), then
        // coalesced window writes.
        const bool h0_direct = args.h0.ptr && !((args.h0.src0 | args.h0.ncols) & 15) && !((args.h0.col0 | args.h0.ld) & 7) &&
                               !(reinterpret_cast<uintptr_t>(args.h0.ptr) & 15);
        const bool staged = args.o0.ptr || args.o1.ptr || args.sum.out || (args.h0.ptr && !h0_direct);
        {
            const long long grow = tile * 128 + r;
            __half* hrow = args.h0.ptr + (size_t)grow * args.h0.ld + args.h0.col0 - args.h0.src0;
            #pragma unroll
            for (int c = 0; c < OUT; c += 16) {
                uint32_t v[16];
                tmem_ld16(t_o + lane_sel + c, v);
                tmem_ld_wait();
                if (staged) {
                    #pragma unroll
                    for (int j = 0; j < 16; ++j) sY[r * C::YS + c + j] = __uint_as_float(v[j]);
                }
                if (OUT == 16 && args.hin.color_in && grow < n) {
                    // density MLP: the three heads' input rows, from this row's registers (HeadIn, mlp_args.cuh)
                    uint32_t geo[8];
                    #pragma unroll
                    for (int j = 0; j < 7; ++j) geo[j] = pack_h2(__uint_as_float(v[2 * j + 1]), __uint_as_float(v[2 * j + 2]));
                    geo[7] = pack_h2(__uint_as_float(v[15]), 1.0f);
                    const uint4 g0 = make_uint4(geo[0], geo[1], geo[2], geo[3]), g1 = make_uint4(geo[4], geo[5], geo[6], geo[7]);
                    const size_t rd = args.hin.sray ? (size_t)__ldg(args.hin.sray + grow) : (size_t)grow;
                    // (d + 1) / 2 then 2 x - 1, as the reference + tcnn do
                    const float dx = ((__ldg(args.hin.dirs + rd * 3) + 1.f) * 0.5f) * 2.f - 1.f;
                    const float dy = ((__ldg(args.hin.dirs + rd * 3 + 1) + 1.f) * 0.5f) * 2.f - 1.f;
                    const float dz = ((__ldg(args.hin.dirs + rd * 3 + 2) + 1.f) * 0.5f) * 2.f - 1.f;
                    float sh[16];
                    al_sh4(dx, dy, dz, sh);
                    uint4* ci = reinterpret_cast<uint4*>(args.hin.color_in + (size_t)grow * 32);
                    ci[0] = make_uint4(pack_h2(sh[0], sh[1]), pack_h2(sh[2], sh[3]), pack_h2(sh[4], sh[5]), pack_h2(sh[6], sh[7]));
                    ci[1] = make_uint4(pack_h2(sh[8], sh[9]), pack_h2(sh[10], sh[11]), pack_h2(sh[12], sh[13]), pack_h2(sh[14], sh[15]));
                    ci[2] = g0; ci[3] = g1;
                    uint4* sf = reinterpret_cast<uint4*>(args.hin.semf_in + (size_t)grow * 16);
                    sf[0] = g0; sf[1] = g1;
                    uint4* so = reinterpret_cast<uint4*>(args.hin.semo_in + (size_t)grow * args.hin.ld_semo + args.hin.F);
                    so[0] = g0; so[1] = g1;
                }
                if (h0_direct && c >= args.h0.src0 && c < args.h0.src0 + args.h0.ncols && grow < n) {
                    uint32_t h[8];
                    #pragma unroll
                    for (int j = 0; j < 8; ++j) h[j] = pack_h2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                    if (args.h0.act == 1) {
                        const __half2 z = __float2half2_rn(0.f);
                        #pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&h[j]), z);
                            h[j] = *reinterpret_cast<const uint32_t*>(&m);
                        }
                    }
                    uint4* dst = reinterpret_cast<uint4*>(hrow + c);
                    dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
                    dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (staged) {
            const float* rows = sY + wq * 32 * C::YS;
            const long long row0 = tile * 128 + wq * 32;
            write_window<float>(args.o0.ptr, args.o0.ld, args.o0.col0, args.o0.src0, args.o0.ncols, args.o0.act, rows, C::YS, row0, n, lane);
            write_window<float>(args.o1.ptr, args.o1.ld, args.o1.col0, args.o1.src0, args.o1.ncols, args.o1.act, rows, C::YS, row0, n, lane);
            if (!h0_direct)
                write_window<__half>(args.h0.ptr, args.h0.ld, args.h0.col0, args.h0.src0, args.h0.ncols, args.h0.act, rows, C::YS, row0, n, lane);
            if (args.sum.out) {
                // compositing fused into the epilogue (OutSum): weighted column sums of this warp's 32 rows, one
                // reduction per (ray, channel)
                const long long rowi = row0 + lane;
                float my_w = 0.f;
                int my_ray = -1;
                if (rowi < n) { my_w = __ldg(args.sum.w + rowi); my_ray = __ldg(args.sum.ray + rowi); }
                const int nc = args.sum.ncols, act = args.sum.act;
                const float* ys = rows + args.sum.src0;
                const unsigned nz = __ballot_sync(0xffffffffu, my_w != 0.f);
                if (nz) {
                    const int ray0 = __shfl_sync(0xffffffffu, my_ray, __ffs(nz) - 1);
                    if (__all_sync(0xffffffffu, my_w == 0.f || my_ray == ray0)) {
                        // the usual case (waves of 32 k samples per ray): all weighted rows belong to one ray.  No
                        // data-dependent control flow; a zero weight masks its row (unused slots may hold anything).
                        float* o = args.sum.out + (size_t)ray0 * args.sum.ld + args.sum.col0;
                        if (nc <= 4) {
                            // few channels (rgb): lane = row, butterfly over the rows
                            for (int c = 0; c < nc; ++c) {
                                float v = my_w != 0.f ? my_w * al_apply_act(ys[lane * C::YS + c], act) : 0.f;
                                #pragma unroll
                                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                                if (lane == 0) atomicAdd(o + c, v);
                            }
                        } else {
                            // lane = channel (lane, lane + 32): 32 independent row terms, four partial sums each
                            const int c0 = min((int)lane, nc - 1), c1 = min((int)lane + 32, nc - 1);
                            float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
                            auto rows_sum = [&](auto act_c, auto wide_c) {      // activation and width fixed per instantiation
                                constexpr int ACT = decltype(act_c)::value;
                                constexpr bool WIDE = decltype(wide_c)::value;
                                #pragma unroll
                                for (int rr = 0; rr < 32; ++rr) {
                                    const float wr = __shfl_sync(0xffffffffu, my_w, rr);
                                    const bool on = wr != 0.f;
                                    a0[rr & 3] = fmaf(wr, on ? al_apply_act(ys[rr * C::YS + c0], ACT) : 0.f, a0[rr & 3]);
                                    if (WIDE) a1[rr & 3] = fmaf(wr, on ? al_apply_act(ys[rr * C::YS + c1], ACT) : 0.f, a1[rr & 3]);
                                }
                            };
                            using std::integral_constant;
                            if (act == 0) {
                                if (nc > 32) rows_sum(integral_constant<int, 0>{}, integral_constant<bool, true>{});
                                else rows_sum(integral_constant<int, 0>{}, integral_constant<bool, false>{});
                            } else if (act == 1) {
                                rows_sum(integral_constant<int, 1>{}, integral_constant<bool, true>{});
                            } else {
                                rows_sum(integral_constant<int, 2>{}, integral_constant<bool, true>{});
                            }
                            if ((int)lane < nc) atomicAdd(o + lane, (a0[0] + a0[1]) + (a0[2] + a0[3]));
                            if ((int)lane + 32 < nc) atomicAdd(o + lane + 32, (a1[0] + a1[1]) + (a1[2] + a1[3]));
                        }
                    } else {
                        // several rays among the 32 rows (wave lengths that are not multiples of 32): ray segments in row order
                        float acc0 = 0.f, acc1 = 0.f;
                        int cur = -1;
                        auto flush = [&]() {
                            if (cur < 0) return;
                            float* o = args.sum.out + (size_t)cur * args.sum.ld + args.sum.col0;
                            if ((int)lane < nc) atomicAdd(o + lane, acc0);
                            if ((int)lane + 32 < nc) atomicAdd(o + lane + 32, acc1);
                        };
                        for (int rr = 0; rr < 32; ++rr) {
                            const float wr = __shfl_sync(0xffffffffu, my_w, rr);
                            const int ry = __shfl_sync(0xffffffffu, my_ray, rr);
                            if (wr == 0.f) continue;                  // warp-uniform
                            if (ry != cur) { flush(); cur = ry; acc0 = acc1 = 0.f; }
                            if ((int)lane < nc) acc0 = fmaf(wr, al_apply_act(ys[rr * C::YS + lane], act), acc0);
                            if ((int)lane + 32 < nc) acc1 = fmaf(wr, al_apply_act(ys[rr * C::YS + lane + 32], act), acc1);
                        }
                        flush();
                    }
                }
            }
        }
        __syncwarp();
    }
    cp_async_wait_all();
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ========================================================================================== backward
template <int IN, int H, int OUT, int NH>
struct BwdCfg {
    using S = Shape<IN, H, OUT, NH>;
    static constexpr uint32_t oW1 = 0, oW2 = oW1 + S::bW1, oWO = oW2 + S::bW2;
    static constexpr uint32_t oA0 = oWO + S::bWO;                  // 2 x [128][IN]  layer-1 input (double buffered)
    static constexpr uint32_t oA1 = oA0 + 2 * S::bX;               // [128][H]   relu(h1)
    static constexpr uint32_t oA2 = oA1 + S::bH;                   // [128][H]   relu(h2), later d h1      (NH == 2)
    static constexpr uint32_t oDL = oA2 + (NH == 2 ? S::bH : 0);   // [128][H]   d h_last
    static constexpr uint32_t oDO = oDL + S::bH;                   // [128][OUT] d out (scaled, fp16)
    // staged source windows of the output gradient, filled by cp.async one tile ahead: 3 big slots (up to OUT fp32
    // columns, or OUT/4 16-byte chunks) and 2 one-column slots
    static constexpr int BIG_ROW = ((OUT | 1) * 4 > ((OUT / 4) | 1) * 16) ? (OUT | 1) * 4 : ((OUT / 4) | 1) * 16;
    static constexpr int BIG_ROW16 = BIG_ROW > 80 ? BIG_ROW : 80;  // kind 4 stages [128][16] fp32 as 4 chunks (stride 5)
    static constexpr uint32_t bBig = 128 * BIG_ROW16, bSmall = 128 * 4;
    static constexpr uint32_t oSlot = oDO + S::bO;                  // big0 big1 big2 small0 small1
    static constexpr uint32_t oDX = oSlot + 3 * bBig + 2 * bSmall;  // d x staging (row-major d x only, when it fits)
    static constexpr int DX_W = IN | 1;
    static constexpr uint32_t bDX = (128 * DX_W * 4 + 127) / 128 * 128;
    static constexpr bool kDxStage = oDX + bDX + 64 <= 227 * 1024;
    static constexpr uint32_t oBar = oDX + (kDxStage ? bDX : 0);
    static constexpr uint32_t BYTES = oBar + 16 + 16;
    // TMEM columns
    static constexpr int tACC = 0, tDX = tACC + H, tW1 = tDX + IN, tW2 = tW1 + IN, tWO = tW2 + (NH == 2 ? H : 0);
    static constexpr int TCOLS = tWO + OUT;
    static_assert(TCOLS <= 512, "TMEM columns");
    static_assert(BYTES <= 227 * 1024, "shared memory");
};

// dW accumulator [MW x NW] in TMEM (row i -> lane i for MW = 128, lane (i/16)*32 + i%16 for MW = 64)
// -> red.global.add into dW, element (row, col) at dW[row * s_row + col * s_col].
template <int MW, int NW, int NP>
__device__ __forceinline__ void flush_dw(uint32_t taddr, float* __restrict__ dW, int s_row, int s_col, float inv_scale,
                                         int wq, int part, int lane) {
    const int row = MW == 128 ? wq * 32 + lane : wq * 16 + lane;
    const bool valid = MW == 128 || lane < 16;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    constexpr int NCH = NW / 16;                                   // 16-column chunks, split over the column parts
    // row-major weight matrices (s_col == 1): a thread's 16 columns are contiguous -> four 16-byte vector reductions
    const bool vec = s_col == 1 && (s_row & 3) == 0 && (reinterpret_cast<uintptr_t>(dW) & 15u) == 0;
    for (int ch = part; ch < NCH; ch += NP) {
        uint32_t v[16];
        tmem_ld16(taddr + lane_sel + ch * 16, v);
        tmem_ld_wait();
        if (valid) {
            if (vec) {
                float4* d = reinterpret_cast<float4*>(dW + (size_t)row * s_row + ch * 16);
                #pragma unroll
                for (int j = 0; j < 4; ++j)
                    atomicAdd(d + j, make_float4(__uint_as_float(v[4 * j]) * inv_scale, __uint_as_float(v[4 * j + 1]) * inv_scale,
                                                 __uint_as_float(v[4 * j + 2]) * inv_scale, __uint_as_float(v[4 * j + 3]) * inv_scale));
            } else {
                #pragma unroll
                for (int j = 0; j < 16; ++j)
                    atomicAdd(dW + (size_t)row * s_row + (size_t)(ch * 16 + j) * s_col, __uint_as_float(v[j]) * inv_scale);
            }
        }
    }
}

// The source windows of the output gradient (see DoutSpec), in slot order big0 big1 big2 small0 small1.
__device__ __forceinline__ void dout_windows(const MlpBwdArgs& a, Win (&w)[5]) {
    const DoutSpec& sp = a.spec;
    #pragma unroll
    for (int i = 0; i < 5; ++i) w[i] = {nullptr, 0, 0, 0, 4, 0, 0, 0, 0};
    const bool r1 = sp.w != nullptr;
    if (r1 && sp.kind != 0 && sp.kind != 4) { w[3] = {sp.w, 1, 0, 1, 4, 0, 0, 0, 0}; w[4] = {sp.sray, 1, 0, 1, 4, 0, 0, 0, 0}; }
    switch (sp.kind) {
    case 0: w[0] = {a.dout, a.ld_dout, a.dcol0, a.dncols, 4, 0, 0, 0, 0}; break;
    case 1:
        if (!r1) w[0] = {sp.g_vals, sp.ldg, 1 + 3, sp.C, 4, 0, 0, 0, 0};
        break;
    case 2:
        w[0] = {sp.relu_feat, sp.ld_relu * 2, 0, sp.F / 8, 16, 0, 0, 0, 0};
        w[1] = {sp.d_feat, sp.ld_dfeat * 4, 0, sp.F / 4, 16, 0, 0, 0, 0};
        if (!r1) w[2] = {sp.g_vals, sp.ldg, 1 + 3 + sp.C, sp.F, 4, 0, 0, 0, 0};
        break;
    case 3:
        w[0] = {sp.vals, sp.ldv, 1, 3, 4, 0, 0, 0, 0};
        if (!r1) w[1] = {sp.g_vals, sp.ldg, 1, 3, 4, 0, 0, 0, 0};
        break;
    default:
        w[0] = {sp.dgeo, 64, 0, 4, 16, 0, 0, 0, 0};
        w[3] = {sp.h16, 16, 0, 1, 4, 0, 0, 0, 0};
        if (r1) w[4] = {sp.g_sigma, 1, 0, 1, 4, 0, 0, 0, 0};
        else w[4] = {sp.g_vals, sp.ldg, 0, 1, 4, 0, 0, 0, 0};
        break;
    }
}

// NP = column parts per TMEM lane quarter: the CTA has 4 * NP warps (128 * NP threads).
template <int IN, int H, int OUT, int NH, int NP>
__global__ void __launch_bounds__(128 * NP, 1) k_mlp_bwd_tc(const MlpBwdArgs args) {
    using S = Shape<IN, H, OUT, NH>;
    using C = BwdCfg<IN, H, OUT, NH>;
    constexpr int NT = 128 * NP;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x;
    const int wq = (tid >> 5) & 3, part = tid >> 7, lane = tid & 31, warp = tid >> 5;
    const int r = wq * 32 + lane;                                  // tile row == TMEM lane == sample
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::oBar + 16);

    stage_weights(args.params + S::W1, smem + C::oW1, H, IN, tid, NT);
    if (NH == 2) stage_weights(args.params + S::W2, smem + C::oW2, H, H, tid, NT);
    stage_weights(args.params + S::WO, smem + C::oWO, OUT, H, tid, NT);
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
    unsigned char* sA1 = smem + C::oA1;
    unsigned char* sA2 = smem + C::oA2;
    unsigned char* sDL = smem + C::oDL;
    unsigned char* sDO = smem + C::oDO;
    unsigned char* sD1 = NH == 2 ? sA2 : sDL;                      // d h1 (first hidden layer's gradient)
    const uint32_t aA0base = smem_u32(smem + C::oA0);
    const uint32_t aA1 = smem_u32(sA1), aA2 = smem_u32(sA2), aDL = smem_u32(sDL), aDO = smem_u32(sDO), aD1 = smem_u32(sD1);
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
    uint32_t par0 = 0, par1 = 0;

    const long long n = args.n_dev ? min((long long)args.cap, (long long)*args.n_dev) : (long long)args.cap;
    const long long n_tiles = (n + 127) / 128;
    const float scale = al_grad_scale(args.amax_dev);
    const float inv_scale = 1.0f / scale;
    const DoutSpec& sp = args.spec;

    // slots: big0 big1 big2 small0 small1
    auto slot_off = [](int i) -> uint32_t { return C::oSlot + (i < 3 ? i * C::bBig : 3 * C::bBig + (i - 3) * C::bSmall); };
    // one tile's inputs, asynchronously: x rows -> A0[buf], output-gradient source windows -> slots
    Win wins[5];
    dout_windows(args, wins);
    #pragma unroll
    for (int i = 0; i < 5; ++i) win_init(wins[i], tid, NT);
    auto prefetch = [&](long long t, int buf) {
        load_x_tile_async<IN>(args.x, args.ldx, t * 128, n, aA0base + buf * S::bX, tid, NT);
        #pragma unroll
        for (int i = 0; i < 5; ++i) stage_window_async(wins[i], t * 128, n, smem_u32(smem + slot_off(i)));
    };
    auto slotf = [&](int i) { return reinterpret_cast<const float*>(smem + slot_off(i)); };

    // d out is assembled in 8-column chunks: chunk q of row r by the thread (r, part) with q % NP == part
    constexpr int NCHUNK = OUT / 8;
    constexpr int HP = H / NP;                                     // hidden columns per part in the epilogues

    bool any = false, pending = false;
    int buf = 0;
    long long tile = blockIdx.x;
    if (tile < n_tiles) prefetch(tile, 0);
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const long long row = tile * 128 + r;
        if (pending) { mbar_wait(bar1, par1); par1 ^= 1; pending = false; }   // last tile's weight-gradient GEMMs done
        cp_async_wait_all();
        __syncthreads();                                           // every thread's copies have landed
        // ---- assemble d out (scaled, fp16) from the staged windows
        for (int q = part; q < NCHUNK; q += NP) {
            const int c0 = q * 8;
            float dr[8];
            #pragma unroll
            for (int j = 0; j < 8; ++j) dr[j] = 0.f;
            const bool r1 = sp.w != nullptr;
            float wrow = 1.0f;
            const float* grow = nullptr;                            // rank-1: this sample's ray row of g_out
            if (r1 && sp.kind != 0 && sp.kind != 4) {
                wrow = slotf(3)[r];
                grow = sp.g_out + (size_t)reinterpret_cast<const int*>(slotf(4))[r] * sp.K;
            }
            if (sp.kind == 0) {
                const float* s0 = slotf(0) + r * (args.dncols | 1);
                #pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (c0 + j < args.dncols) dr[j] = s0[c0 + j];
            } else if (sp.kind == 1) {
                if (r1) {
                    #pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c0 + j < sp.C) dr[j] = wrow * __ldg(grow + 3 + c0 + j);
                } else {
                    const float* s0 = slotf(0) + r * (sp.C | 1);
                    #pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c0 + j < sp.C) dr[j] = s0[c0 + j];
                }
            } else if (sp.kind == 2) {
                if (c0 < sp.F) {                                    // F is a multiple of 16: whole chunks
                    const uint4 mk = *reinterpret_cast<const uint4*>(smem + slot_off(0) + (r * ((sp.F / 8) | 1) + q) * 16);
                    const float4 d0 = *reinterpret_cast<const float4*>(smem + slot_off(1) + (r * ((sp.F / 4) | 1) + 2 * q) * 16);
                    const float4 d1 = *reinterpret_cast<const float4*>(smem + slot_off(1) + (r * ((sp.F / 4) | 1) + 2 * q + 1) * 16);
                    const float df[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                    const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
                    float gq[8];
                    if (r1) {
                        #pragma unroll
                        for (int j = 0; j < 8; ++j) gq[j] = wrow * __ldg(grow + 3 + sp.C + c0 + j);
                    } else {
                        const float* gf = slotf(2) + r * (sp.F | 1) + c0;
                        #pragma unroll
                        for (int j = 0; j < 8; ++j) gq[j] = gf[j];
                    }
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 m = __half22float2(*reinterpret_cast<const __half2*>(&mw[j]));
                        dr[2 * j] = gq[2 * j] + (m.x > 0.f ? df[2 * j] : 0.f);
                        dr[2 * j + 1] = gq[2 * j + 1] + (m.y > 0.f ? df[2 * j + 1] : 0.f);
                    }
                }
            } else if (sp.kind == 3) {
                if (c0 == 0) {
                    const float* rgbp = slotf(0) + r * 3;
                    #pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float c = rgbp[j];
                        const float gj = r1 ? wrow * __ldg(grow + j) : slotf(1)[r * 3 + j];
                        dr[j] = gj * c * (1.0f - c);
                    }
                }
            } else if (c0 < 16) {
                const float* dg = reinterpret_cast<const float*>(smem + slot_off(0) + r * 80);
                #pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + j;
                    if (c == 0) dr[j] = slotf(4)[r] * __expf(fminf(fmaxf(slotf(3)[r], -15.f), 15.f));
                    else dr[j] = dg[c - 1];
                }
            }
            float f[8];
            #pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(dr[e] * scale, -65504.f), 65504.f);
            uint4 o;
            o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]); o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
            *reinterpret_cast<uint4*>(sDO + (r >> 3) * (OUT / 8) * 128 + (r & 7) * 16 + q * 128) = o;
        }
        const uint32_t aA0 = aA0base + buf * S::bX;
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // the slots are consumed and A0[buf ^ 1] is no longer read: fetch the next tile behind this one's GEMMs
        if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x, buf ^ 1);
        // ---- forward recompute
        if (tid == 0) {
            tc_fence_after();
            issue_gemm<128, H, IN, false, false>(tmem + C::tACC, view_k(aA0, IN), view_k(aW1, IN), false);
            mma_commit(bar0);
        }
        mbar_wait(bar0, par0); par0 ^= 1;
        tc_fence_after();
        epi_to_tile<H, 0>(tmem + C::tACC + lane_sel, part * HP, (part + 1) * HP, sA1, nullptr, r);
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
            epi_to_tile<H, 0>(tmem + C::tACC + lane_sel, part * HP, (part + 1) * HP, sA2, nullptr, r);
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
        epi_to_tile<H, 1>(tmem + C::tACC + lane_sel, part * HP, (part + 1) * HP, sDL, sAL, r);
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
            epi_to_tile<H, 1>(tmem + C::tACC + lane_sel, part * HP, (part + 1) * HP, sD1, sA1, r);
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
            if (C::kDxStage && args.dx_mode == 0) {
                // row-major d x: stage the tile, then coalesced window writes (4 NP warps x 32 / NP rows)
                float* sDX = reinterpret_cast<float*>(smem + C::oDX);
                for (int ch = part; ch < NCH; ch += NP) {
                    uint32_t v[16];
                    tmem_ld16(tmem + C::tDX + lane_sel + ch * 16, v);
                    tmem_ld_wait();
                    #pragma unroll
                    for (int j = 0; j < 16; ++j) sDX[r * C::DX_W + ch * 16 + j] = __uint_as_float(v[j]) * inv_scale;
                }
                tc_fence_before();
                __syncthreads();
                constexpr int RW = 32 / NP;                        // rows per warp
                write_window<float>(args.dx, args.ld_dx, 0, args.dx_c0, args.dx_n, 0, sDX + warp * RW * C::DX_W, C::DX_W,
                                    tile * 128 + warp * RW, n, lane, RW, args.dx_acc);
                write_window<float>(args.dx2, args.ld_dx2, 0, args.dx2_c0, args.dx2_n, 0, sDX + warp * RW * C::DX_W, C::DX_W,
                                    tile * 128 + warp * RW, n, lane, RW, args.dx2_acc);
            } else {
                for (int ch = part; ch < NCH; ch += NP) {
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
                                    float* d = args.dx + (size_t)row * args.ld_dx + rel;
                                    d[0] = args.dx_acc ? d[0] + v0 : v0;
                                    if (rel + 1 < args.dx_n) d[1] = args.dx_acc ? d[1] + v1 : v1;
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
    }
    cp_async_wait_all();
    if (pending) { mbar_wait(bar1, par1); par1 ^= 1; }
    tc_fence_after();
    if (any && args.dparams) {
        // dW1 [H][IN]: lanes = out, columns = in;  dW2 [H][H] likewise;  dWo^T: lanes = in (H), columns = out
        flush_dw<H, IN, NP>(tmem + C::tW1, args.dparams + S::W1, IN, 1, inv_scale, wq, part, lane);
        if (NH == 2) flush_dw<H, H, NP>(tmem + C::tW2, args.dparams + S::W2, H, 1, inv_scale, wq, part, lane);
        flush_dw<H, OUT, NP>(tmem + C::tWO, args.dparams + S::WO, 1, H, inv_scale, wq, part, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ========================================================================================== backward, two tiles in flight
// Same arithmetic as k_mlp_bwd_tc (identical GEMMs, identical epilogues), rescheduled so that the tensor pipe and the
// TMEM reads overlap ACROSS tiles:
//  * two groups of 256 threads (8 warps: 4 TMEM lane quarters x 2 column halves), each walking its own sequence of
//    128-sample tiles with its own activation buffers and its own accumulator columns; while one group runs an
//    epilogue (tcgen05.ld -> ReLU / mask -> smem) the other group's GEMMs run;
//  * ONE issuing warp (warp 16) serves both groups: a group posts "operands ready" on its mbarrier (one arrive per
//    warp after fence.proxy.async), the issuer polls the two barriers, issues the phase's GEMMs and commits to the
//    group's "done" mbarrier.  A single issuer keeps every accumulation into the SHARED weight-gradient columns in
//    program order;
//  * to fit two tiles next to the weights (208 KB for the 48->128->128->16 density MLP) the gradient tiles are written
//    IN PLACE over the activation they are masked by (d h2 over relu(h2), d h1 over relu(h1): the element a thread
//    overwrites is the one it just read as its mask); the commit of a phase therefore covers the phase's weight-gradient
//    GEMM too (it still reads the tile that the epilogue is about to overwrite);
//  * the output gradient of the NEXT tile is assembled straight from global memory while the last GEMMs of the current
//    tile run; d x goes from TMEM registers to global memory (row = thread: whole 32 B sectors per thread).
// Staging of the output-gradient sources of ONE tile (per group), filled by cp.async a whole tile ahead:
//   s0 [128] : w[row] (kinds 1-3)  |  h16[row, 0] (kind 4)             s1 [128] : sray[row] (kinds 1-3)  |  d sigma (kind 4)
//   sA, sB   : kind 2 only (OUT >= 32): relu(features) mask rows (16-byte chunks, odd chunk stride) and d_feat rows
// so that the assembly has no dependent global loads (sray -> g_out row) and no latency per chunk.  Kind 0 and the
// materialised-gradient forms (w == NULL, kinds 1-3) take the one-tile kernel.
template <int OUT>
struct DoutStage {
    static constexpr bool kBulk = OUT >= 32;
    static constexpr int A_CH = OUT / 8, B_CH = OUT / 4;           // 16-byte chunks per row: fp16 mask, fp32 d_feat
    // row r reads ITS chunks: rows of a multiple of 8 chunks are XOR-swizzled inside each 128-byte group (chunk ^ (r & 7):
    // 8 consecutive rows cover all 32 banks, no padding bytes), other widths get an odd chunk stride
    static constexpr int A_STRIDE = A_CH % 8 == 0 ? A_CH : (A_CH | 1), B_STRIDE = B_CH % 8 == 0 ? B_CH : (B_CH | 1);
    template <int CH>
    __device__ static __forceinline__ int chunk(int r, int j) {
        if constexpr (CH % 8 == 0) return r * CH + ((j & ~7) | ((j & 7) ^ (r & 7)));
        else return r * (CH | 1) + j;
    }
    static constexpr uint32_t oS0 = 0, oS1 = 512, oA = 1024;
    static constexpr uint32_t oB = oA + (kBulk ? 128 * A_STRIDE * 16 : 0);
    static constexpr uint32_t BYTES = oB + (kBulk ? 128 * B_STRIDE * 16 : 0);
};

// Timing experiments only (results are wrong when set; tools/job_bwd_dbg.sh): al_set_bwd_debug(bits)  1: skip the epilogues'
// TMEM reads / conversions / smem writes, 2: issue no GEMMs (commit at once), 4: skip output-gradient assembly and the
// d x write-out.  0 in normal operation (a uniform branch per phase).
static int g_bwd_dbg = 0;
#define AL_BWD_DBG (args.dbg)
#ifndef AL_BWD_DIRECT
#define AL_BWD_DIRECT 1
#endif
#ifndef AL_BWD_FOUR_GROUPS
#define AL_BWD_FOUR_GROUPS 0      // measured (profiles/): four 128-thread groups are slower than two 256-thread groups at H = 64
#endif
template <int IN, int H, int OUT, int NH>
struct Bwd2Cfg {
    using S = Shape<IN, H, OUT, NH>;
    // Tiles in flight per CTA.  128-wide hidden layers: two groups of 256 threads (two column halves per TMEM lane quarter)
    // is what shared memory holds next to the weights; 64-wide heads: four groups of 128 threads.
    static constexpr int NP = (H >= 128 || !AL_BWD_FOUR_GROUPS) ? 2 : 1;   // column parts per lane quarter
    static constexpr int NG = (H >= 128 || !AL_BWD_FOUR_GROUPS) ? 2 : 4;   // groups = tiles in flight
    static constexpr int GT = 128 * NP;                           // threads per group
    static constexpr int GW = 4 * NP;                             // warps per group
    static constexpr uint32_t oW1 = 0, oW2 = oW1 + S::bW1, oWO = oW2 + S::bW2, oGrp = oWO + S::bWO;
    static constexpr uint32_t gA0 = 0, gA1 = gA0 + S::bX, gA2 = gA1 + S::bH, gDO = gA1 + (NH == 2 ? 2 : 1) * S::bH;
    static constexpr uint32_t gST = gDO + S::bO;                   // staged sources of the next tile's output gradient
    static constexpr uint32_t gX2 = gST + DoutStage<OUT>::BYTES;   // second x tile (double buffer), where shared memory allows
    static constexpr uint32_t bGrp1 = gX2;
    static constexpr bool kX2 = (S::bW1 + S::bW2 + S::bWO) + NG * (bGrp1 + S::bX) + 16 * NG + 16 <= 227 * 1024;
    static constexpr uint32_t bGrp = bGrp1 + (kX2 ? S::bX : 0);
    static constexpr uint32_t oBar = oGrp + NG * bGrp;            // ready[NG], done[NG], tmem slot
    static constexpr uint32_t BYTES = oBar + 16 * NG + 16;
    static constexpr int TA = H > IN ? H : IN;                    // per-group accumulator columns (hidden / d x)
    static constexpr int tW1 = NG * TA, tW2 = tW1 + IN, tWO = tW2 + (NH == 2 ? H : 0), TCOLS = tWO + OUT;
    static constexpr bool kFits = TCOLS <= 512 && BYTES <= 227 * 1024;
    static constexpr int NPH = NH == 2 ? 5 : 3;                   // GEMM phases per tile
    // AL_BWD_DIRECT: every group's first thread issues the group's GEMMs itself (after a group barrier) instead of posting
    // to an issuing warp -- one mbarrier hop less per phase.  GEMMs of BOTH groups then accumulate into the shared
    // weight-gradient columns from two issuing threads: each tcgen05.mma is a whole read-modify-write of its accumulator
    // in the CTA's single tensor pipe, so the sums are complete whatever the interleaving
    // (tests/test_mlp_gpu.py::test_backward_weight_gradients_are_exact_sums); the accumulators are zero-filled up front
    // because "first GEMM overwrites" has no owner any more.
    static constexpr bool kDirect = AL_BWD_DIRECT != 0;
    static constexpr int NT = NG * GT + (kDirect ? 0 : 32);
    static constexpr int MMA_WARP = kDirect ? -1 : NG * GW;
};

template <int OUT>
__device__ __forceinline__ void prefetch_dout(const MlpBwdArgs& a, long long row0, long long n, unsigned char* st, int tg, int gt) {
    using D = DoutStage<OUT>;
    const DoutSpec& sp = a.spec;
    if (sp.kind == 0) return;
    const uint32_t base = smem_u32(st);
    for (int i = tg; i < 256; i += gt) {
        const int r = i & 127;
        const long long row = row0 + r;
        const bool valid = row < n;
        const void* src;
        if (i < 128) src = sp.kind == 4 ? (const void*)(sp.h16 + (size_t)row * 16) : (const void*)(sp.w + row);
        else src = sp.kind == 4 ? (sp.w ? (const void*)(sp.g_sigma + row) : (const void*)(sp.g_vals + (size_t)row * sp.ldg))
                                : (const void*)(sp.sray + row);
        cp_async4(base + (i < 128 ? D::oS0 : D::oS1) + r * 4, valid ? src : (const void*)a.params, valid);
    }
    if constexpr (D::kBulk) {
        if (sp.kind == 2) {
            constexpr int ach = D::A_CH, bch = D::B_CH;            // F == OUT (checked by the launcher): shifts, no divisions
            for (int i = tg; i < 128 * ach; i += gt) {
                const int r = i / ach, j = i - r * ach;
                const bool valid = row0 + r < n;
                const void* src = valid ? (const void*)(sp.relu_feat + (size_t)(row0 + r) * sp.ld_relu + j * 8) : (const void*)a.params;
                cp_async16(base + D::oA + (uint32_t)D::template chunk<D::A_CH>(r, j) * 16u, src, valid);
            }
            for (int i = tg; i < 128 * bch; i += gt) {
                const int r = i / bch, j = i - r * bch;
                const bool valid = row0 + r < n;
                const void* src = valid ? (const void*)(sp.d_feat + (size_t)(row0 + r) * sp.ld_dfeat + j * 4) : (const void*)a.params;
                cp_async16(base + D::oB + (uint32_t)D::template chunk<D::B_CH>(r, j) * 16u, src, valid);
            }
        }
    }
}

// the four heads of the C2 field must take the two-tile schedule
static_assert(Bwd2Cfg<48, 128, 16, 2>::kFits && Bwd2Cfg<32, 128, 16, 2>::kFits && Bwd2Cfg<16, 64, 64, 2>::kFits &&
              Bwd2Cfg<80, 64, 16, 1>::kFits, "two-tile backward: C2 head shapes must fit");

// One 8-column chunk (columns c0 .. c0 + 7) of row r's output gradient from the staged sources (+ independent global loads).
template <int OUT, int KIND>
__device__ __forceinline__ void dout_chunk(const MlpBwdArgs& a, const unsigned char* st, int r, long long row, long long n,
                                           int c0, float (&dr)[8]) {
    using D = DoutStage<OUT>;
    const DoutSpec& sp = a.spec;
    #pragma unroll
    for (int j = 0; j < 8; ++j) dr[j] = 0.f;
    if (row >= n) return;
    if (KIND == 0) {                                           // a plain d-out matrix (the tcnn.Network operator): direct loads
        const float* s0 = a.dout + (size_t)row * a.ld_dout + a.dcol0;
        #pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c0 + j < a.dncols) dr[j] = __ldg(s0 + c0 + j);
        return;
    }
    const float s0 = reinterpret_cast<const float*>(st + D::oS0)[r];
    if (KIND == 4) {
        if (c0 >= 16) return;
        const float4* dg = reinterpret_cast<const float4*>(sp.dgeo + (size_t)row * 16 + c0);
        const float4 a0 = __ldg(dg), a1 = __ldg(dg + 1);
        const float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (c0 == 0) {
            // column 0: d sigma through trunc_exp; columns 1..7: d geo 0..6 (dgeo row = [d geo (15) | unused])
            const float gsig = reinterpret_cast<const float*>(st + D::oS1)[r];
            dr[0] = gsig * __expf(fminf(fmaxf(s0, -15.f), 15.f));
        } else {
            dr[0] = __ldg(sp.dgeo + (size_t)row * 16 + 7);        // columns 8..15: d geo 7..14
        }
        #pragma unroll
        for (int j = 1; j < 8; ++j) dr[j] = v[j - 1];
        return;
    }
    const float wrow = s0;
    const float* grow = sp.g_out + (size_t)reinterpret_cast<const int*>(st + D::oS1)[r] * sp.K;
    if (KIND == 1) {
        #pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c0 + j < sp.C) dr[j] = wrow * __ldg(grow + 3 + c0 + j);
    } else if (KIND == 2) {
        if constexpr (D::kBulk) {
            if (c0 < sp.F) {                                      // F is a multiple of 16: whole chunks
                const int q = c0 >> 3;
                const uint4 mk = *reinterpret_cast<const uint4*>(st + D::oA + D::template chunk<D::A_CH>(r, q) * 16);
                const float4 d0 = *reinterpret_cast<const float4*>(st + D::oB + D::template chunk<D::B_CH>(r, 2 * q) * 16);
                const float4 d1 = *reinterpret_cast<const float4*>(st + D::oB + D::template chunk<D::B_CH>(r, 2 * q + 1) * 16);
                const float df[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
                const float* gs = grow + 3 + sp.C + c0;
                float gq[8];
                #pragma unroll
                for (int j = 0; j < 8; ++j) gq[j] = __ldg(gs + j);
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 m = __half22float2(*reinterpret_cast<const __half2*>(&mw[j]));
                    dr[2 * j] = wrow * gq[2 * j] + (m.x > 0.f ? df[2 * j] : 0.f);
                    dr[2 * j + 1] = wrow * gq[2 * j + 1] + (m.y > 0.f ? df[2 * j + 1] : 0.f);
                }
            }
        }
    } else if (c0 == 0) {                                         // kind 3
        const float* rgbp = sp.vals + (size_t)row * sp.ldv + 1;
        #pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float c = __ldg(rgbp + j);
            dr[j] = wrow * __ldg(grow + j) * c * (1.0f - c);
        }
    }
}

// Row-major windows of d x: dst[row * ld + j] (+)= val[c0 + j], j < n, from this thread's 16 accumulator columns
// [cbase, cbase + 16).  Hot path: windows whose bounds are multiples of 16 columns with 16-byte aligned rows (every use of the
// field: d_feat / dgeo of the semantic heads, dgeo of the colour head) -- a chunk lies wholly inside one window or outside
// all of them, four 16-byte stores (or fire-and-forget 16-byte reductions for +=: a load-add-store would put a global round
// trip on the tile's critical path; the row belongs to this thread alone within a launch).  Anything else takes the
// out-of-line generic routine (keeps the unrolled tile loop small: its instruction footprint showed up as fetch stalls).
struct DxWin {
    float* ptr;
    int ld, c0, n, acc;
};
__device__ __noinline__ void dx_store_generic(const DxWin w0, const DxWin w1, long long row, bool valid, int cbase, uint32_t taddr,
                                              float inv_scale) {
    // Re-reads its 16 accumulator columns itself (the caller's register copy would have to live in local memory otherwise).
    // tcgen05.ld is warp-collective: the whole warp gets here, `valid` masks the rows past the end.
    uint32_t v[16];
    tmem_ld16(taddr, v);
    tmem_ld_wait();
    if (!valid) return;
    #pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        const DxWin& w = k ? w1 : w0;
        if (!w.ptr || w.n <= 0) continue;
        float* drow = w.ptr + (size_t)row * w.ld;
        for (int e = 0; e < 16; ++e) {
            const int re = cbase + e - w.c0;
            if (re >= 0 && re < w.n) {
                const float x = __uint_as_float(v[e]) * inv_scale;
                drow[re] = w.acc ? drow[re] + x : x;
            }
        }
    }
}
__device__ __forceinline__ bool dx_win_fast(const DxWin& w) {
    return !w.ptr || (((w.ld & 3) | (w.c0 & 15) | (w.n & 15)) == 0 && (reinterpret_cast<uintptr_t>(w.ptr) & 15u) == 0);
}
__device__ __forceinline__ void dx_store_fast(const DxWin& w, long long row, int cbase, const float (&val)[16]) {
    if (!w.ptr || cbase < w.c0 || cbase >= w.c0 + w.n) return;
    float4* d = reinterpret_cast<float4*>(w.ptr + (size_t)row * w.ld + (cbase - w.c0));
    #pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 o = make_float4(val[4 * q], val[4 * q + 1], val[4 * q + 2], val[4 * q + 3]);
        if (w.acc) atomicAdd(d + q, o);
        else d[q] = o;
    }
}

// The GEMMs of phase p of one tile (one issuing thread).  acc_*: accumulate into the weight-gradient columns (false only for
// the very first GEMM into an accumulator that was not zero-filled).
template <int IN, int H, int OUT, int NH>
__device__ __forceinline__ void issue_bwd_phase(int p, uint32_t tmem, uint32_t tACC, uint32_t base, uint32_t aA0, uint32_t aW1,
                                                uint32_t aW2, uint32_t aWO, bool want_dx, bool acc1, bool acc2, bool acco) {
    using C = Bwd2Cfg<IN, H, OUT, NH>;
    const uint32_t aA1 = base + C::gA1, aA2 = base + C::gA2, aDO = base + C::gDO;
    const uint32_t aAL = NH == 2 ? aA2 : aA1;                      // last hidden activation, later d h_last in place
    if (p == 0) {
        issue_gemm<128, H, IN, false, false>(tACC, view_k(aA0, IN), view_k(aW1, IN), false);
    } else if (NH == 2 && p == 1) {
        issue_gemm<128, H, H, false, false>(tACC, view_k(aA1, H), view_k(aW2, H), false);
    } else if (p == NH) {
        // d h_last = d out . Wo ;  dWo^T += a_last^T d out
        issue_gemm<128, H, OUT, false, true>(tACC, view_k(aDO, OUT), view_mn(aWO, H), false);
        issue_gemm<H, OUT, 128, true, true>(tmem + C::tWO, view_mn(aAL, H), view_mn(aDO, OUT), acco);
    } else if (NH == 2 && p == 3) {
        // d h1 = d h2 . W2 ;  dW2 += d h2^T a1      (d h2 sits where relu(h2) was)
        issue_gemm<128, H, H, false, true>(tACC, view_k(aA2, H), view_mn(aW2, H), false);
        issue_gemm<H, H, 128, true, true>(tmem + C::tW2, view_mn(aA2, H), view_mn(aA1, H), acc2);
    } else {
        // d x = d h1 . W1 ;  dW1 += d h1^T a0       (d h1 sits where relu(h1) was)
        if (want_dx) issue_gemm<128, IN, H, false, true>(tACC, view_k(aA1, H), view_mn(aW1, IN), false);
        issue_gemm<H, IN, 128, true, true>(tmem + C::tW1, view_mn(aA1, H), view_mn(aA0, IN), acc1);
    }
}

template <int IN, int H, int OUT, int NH>
__global__ void __launch_bounds__(Bwd2Cfg<IN, H, OUT, NH>::NT, 1) k_mlp_bwd_tc2(const MlpBwdArgs args) {
    using S = Shape<IN, H, OUT, NH>;
    using C = Bwd2Cfg<IN, H, OUT, NH>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NG = C::NG;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBar);  // ready[0..NG), done[NG..2NG)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::oBar + 16 * NG);

    stage_weights(args.params + S::W1, smem + C::oW1, H, IN, tid, C::NT);
    if (NH == 2) stage_weights(args.params + S::W2, smem + C::oW2, H, H, tid, C::NT);
    stage_weights(args.params + S::WO, smem + C::oWO, OUT, H, tid, C::NT);
    if (tid == 0) {
        for (int g = 0; g < NG; ++g) {
            mbar_init(smem_u32(&bars[g]), C::GW);                  // ready[g]: one arrive per warp of the group
            mbar_init(smem_u32(&bars[NG + g]), 1);                 // done[g]: tcgen05.commit
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr int ALLOC_WARP = C::kDirect ? 0 : C::MMA_WARP;
    if (warp == ALLOC_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t aW1 = smem_u32(smem + C::oW1), aW2 = smem_u32(smem + C::oW2), aWO = smem_u32(smem + C::oWO);
    const long long n = args.n_dev ? min((long long)args.cap, (long long)*args.n_dev) : (long long)args.cap;
    const long long n_tiles = (n + 127) / 128;
    const long long tile_step = (long long)gridDim.x * NG;
    const bool any = (long long)blockIdx.x * NG < n_tiles;
    if (C::kDirect) {
        // zero the weight-gradient accumulators: 16 warps, 4 column parts per lane quarter
        const uint32_t lsel = (uint32_t)((warp & 3) * 32) << 16;
        for (int c = C::tW1 + (warp >> 2) * 16; c < C::TCOLS; c += 64) tmem_st16_zero(tmem + lsel + c);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    if (!C::kDirect && warp == C::MMA_WARP) {
        // ------------------------------------------------------------------ the issuing warp
        long long cnt[NG];
        int ph[NG];
        uint32_t rpar[NG];
        long long live = 0;
        #pragma unroll
        for (int g = 0; g < NG; ++g) {
            const long long first = (long long)blockIdx.x * NG + g;
            cnt[g] = first < n_tiles ? (n_tiles - 1 - first) / tile_step + 1 : 0;
            ph[g] = 0;
            rpar[g] = 0;
            live += cnt[g];
        }
        bool f1 = false, f2 = false, fo = false;                  // weight-gradient accumulators already written
        while (live > 0) {
            #pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (cnt[g] <= 0) continue;
                // parked by the hardware while waiting (a spinning poll takes issue slots away from the epilogue warps
                // that share this scheduler): short naps, so that no group waits long behind another
                if (!mbar_try_wait_ns(smem_u32(&bars[g]), rpar[g], 40u)) continue;
                rpar[g] ^= 1;
                tc_fence_after();
                const uint32_t base = smem_u32(smem + C::oGrp + g * C::bGrp);
                const uint32_t tACC = tmem + g * C::TA;
                const int p = ph[g];
                if (lane == 0) {
                    if (!(AL_BWD_DBG & 2))
                        issue_bwd_phase<IN, H, OUT, NH>(p, tmem, tACC, base, base + C::gA0, aW1, aW2, aWO, args.dx != nullptr, f1, f2, fo);
                    mma_commit(smem_u32(&bars[NG + g]));
                }
                __syncwarp();
                if (p == NH) fo = true;
                else if (NH == 2 && p == 3) f2 = true;
                else if (p == C::NPH - 1) f1 = true;
                if (++ph[g] == C::NPH) { ph[g] = 0; --cnt[g]; --live; }
            }
        }
    } else {
        // ------------------------------------------------------------------ the tile groups
        const int g = tid / C::GT, tg = tid - g * C::GT;
        const int wq = (tg >> 5) & 3, part = tg >> 7;
        const int r = wq * 32 + lane;                              // tile row == TMEM lane == sample
        unsigned char* gb = smem + C::oGrp + g * C::bGrp;
        unsigned char* sA1 = gb + C::gA1;
        unsigned char* sA2 = gb + C::gA2;
        unsigned char* sAL = NH == 2 ? sA2 : sA1;
        unsigned char* sDO = gb + C::gDO;
        // x tiles: double buffered where shared memory allows (direct issue only: the issuing warp reads the fixed slot) --
        // the next tile's rows are then fetched a whole tile ahead instead of behind this tile's last GEMMs
        constexpr bool kX2 = C::kX2 && C::kDirect;
        const uint32_t aA0 = smem_u32(gb + C::gA0), aA0b = kX2 ? smem_u32(gb + C::gX2) : aA0;
        uint32_t aA0cur = aA0;
        const uint32_t ready = smem_u32(&bars[g]), done = smem_u32(&bars[NG + g]);
        const uint32_t tACC = tmem + g * C::TA + ((uint32_t)(wq * 32) << 16);
        uint32_t dpar = 0;
        const float scale = al_grad_scale(args.amax_dev);
        const float inv_scale = 1.0f / scale;
        constexpr int NCHUNK = OUT / 8;
        constexpr int NP = C::NP;
        constexpr int HP = H / NP;

        const DxWin dxw0 = {args.dx, args.ld_dx, args.dx_c0, args.dx_n, args.dx_acc};
        const DxWin dxw1 = {args.dx2, args.ld_dx2, args.dx2_c0, args.dx2_n, args.dx2_acc};
        const bool dx_fast = dx_win_fast(dxw0) && dx_win_fast(dxw1);
        int phase = 0;
        auto post = [&]() {                                        // this thread's smem writes / TMEM reads are done
            fence_async_smem();
            tc_fence_before();
            if (C::kDirect) {
                named_bar(1 + g, C::GT);                           // ... and the whole group's
                if (tg == 0) {
                    tc_fence_after();
                    if (!(AL_BWD_DBG & 2))
                        issue_bwd_phase<IN, H, OUT, NH>(phase, tmem, tmem + g * C::TA, smem_u32(gb), aA0cur, aW1, aW2, aWO,
                                                        args.dx != nullptr, true, true, true);
                    mma_commit(done);
                }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(ready);
            }
            phase = phase + 1 == C::NPH ? 0 : phase + 1;
        };
        auto wait_done = [&]() {
            mbar_wait(done, dpar);
            dpar ^= 1;
            tc_fence_after();
        };
        unsigned char* sST = gb + C::gST;
        auto assemble = [&](long long t) {                         // scaled fp16 d out of tile t -> sDO
            cp_async_wait_all();                                   // this thread's share of the staged sources ...
            named_bar(1 + g, C::GT);                               // ... and everybody else's
            const long long row = t * 128 + r;
            // the kind is uniform for the launch: one specialised, contiguous copy of the chunk loop per kind
            auto rows = [&](auto kind_c) {
                constexpr int KIND = decltype(kind_c)::value;
                #pragma unroll
                for (int q = 0; q < NCHUNK; ++q) {
                    if (NP == 2 && (NCHUNK >= 4 ? (q >> 1) & 1 : q & 1) != part) continue;
                    float dr[8];
                    dout_chunk<OUT, KIND>(args, sST, r, row, n, q * 8, dr);
                    float f[8];
                    #pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(dr[e] * scale, -65504.f), 65504.f);
                    uint4 o;
                    o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]); o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
                    *reinterpret_cast<uint4*>(sDO + (r >> 3) * (OUT / 8) * 128 + (r & 7) * 16 + q * 128) = o;
                }
            };
            switch (args.spec.kind) {
            case 0: rows(std::integral_constant<int, 0>{}); break;
            case 1: rows(std::integral_constant<int, 1>{}); break;
            case 2: rows(std::integral_constant<int, 2>{}); break;
            case 3: rows(std::integral_constant<int, 3>{}); break;
            default: rows(std::integral_constant<int, 4>{}); break;
            }
            // the staging buffer is refilled after the next post(): with direct issue that post() is a group barrier
            if (!C::kDirect) named_bar(1 + g, C::GT);
        };

        long long tile = (long long)blockIdx.x * NG + g;
        if (tile < n_tiles) {
            load_x_tile_async<IN>(args.x, args.ldx, tile * 128, n, aA0, tg, C::GT);
            prefetch_dout<OUT>(args, tile * 128, n, sST, tg, C::GT);
            assemble(tile);
        }
        for (; tile < n_tiles; tile += tile_step) {
            const long long row = tile * 128 + r;
            cp_async_wait_all();
            post();                                                // A0 + d out ready            -> fwd1
            const long long next = tile + tile_step;
            if (next < n_tiles) {
                prefetch_dout<OUT>(args, next * 128, n, sST, tg, C::GT);   // consumed at the end of this tile
                if (kX2) load_x_tile_async<IN>(args.x, args.ldx, next * 128, n, aA0cur == aA0 ? aA0b : aA0, tg, C::GT);
            }
            wait_done();
            if (!(AL_BWD_DBG & 1)) epi_to_tile<H, 0>(tACC, part * HP, (part + 1) * HP, sA1, nullptr, r);
            post();                                                // relu(h1) ready              -> fwd2 | d h_last
            if (NH == 2) {
                wait_done();
                if (!(AL_BWD_DBG & 1)) epi_to_tile<H, 0>(tACC, part * HP, (part + 1) * HP, sA2, nullptr, r);
                post();                                            // relu(h2) ready              -> d h2, dWo
            }
            wait_done();                                           // covers dWo: a_last may be overwritten
            if (!(AL_BWD_DBG & 1)) epi_to_tile<H, 1>(tACC, part * HP, (part + 1) * HP, sAL, sAL, r);
            post();                                                // d h_last ready              -> d h1, dW2 | d x, dW1
            if (NH == 2) {
                wait_done();                                       // covers dW2: relu(h1) may be overwritten
                if (!(AL_BWD_DBG & 1)) epi_to_tile<H, 1>(tACC, part * HP, (part + 1) * HP, sA1, sA1, r);
                post();                                            // d h1 ready                  -> d x, dW1
            }
            // behind the last GEMMs of this tile: the next tile's output gradient (d out was last read by dWo)
            if (next < n_tiles && !(AL_BWD_DBG & 4)) assemble(next);
            wait_done();                                           // d x ready; dW1 done: A0 and A1 are free
            if (kX2) aA0cur = aA0cur == aA0 ? aA0b : aA0;
            else if (next < n_tiles) load_x_tile_async<IN>(args.x, args.ldx, next * 128, n, aA0, tg, C::GT);
            if (args.dx && !(AL_BWD_DBG & 4)) {
                constexpr int NCH = IN / 16;
                #pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    if (NP == 2 && (ch & 1) != part) continue;
                    if (args.dx_mode == 0 && !dx_fast) {           // warp-uniform: windows that are not 16-column aligned
                        dx_store_generic(dxw0, dxw1, row, row < n, ch * 16, tACC + ch * 16, inv_scale);
                        continue;
                    }
                    uint32_t v[16];
                    tmem_ld16(tACC + ch * 16, v);
                    tmem_ld_wait();
                    if (row < n) {
                        float val[16];
                        #pragma unroll
                        for (int j = 0; j < 16; ++j) val[j] = __uint_as_float(v[j]) * inv_scale;
                        if (args.dx_mode == 0) {
                            dx_store_fast(dxw0, row, ch * 16, val);
                            dx_store_fast(dxw1, row, ch * 16, val);
                        } else {
                            #pragma unroll
                            for (int j = 0; j < 16; j += 2) {
                                const int rel = ch * 16 + j - args.dx_c0;   // dx_c0 is even: a pair never straddles the window
                                if (rel >= 0 && rel + 1 < args.dx_n)
                                    *reinterpret_cast<float2*>(args.dx + ((size_t)(rel >> 1) * args.ld_dx + row) * 2) =
                                        make_float2(val[j], val[j + 1]);
                                else if (rel >= 0 && rel < args.dx_n)
                                    args.dx[((size_t)(rel >> 1) * args.ld_dx + row) * 2] = val[j];
                            }
                        }
                    }
                }
            }
        }
        cp_async_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (any && args.dparams && warp < 16) {                      // NG * GW = 16 epilogue warps in either configuration
        const float inv_scale = 1.0f / al_grad_scale(args.amax_dev);
        const int wq = warp & 3, part = warp >> 2;
        flush_dw<H, IN, 4>(tmem + C::tW1, args.dparams + S::W1, IN, 1, inv_scale, wq, part, lane);
        if (NH == 2) flush_dw<H, H, 4>(tmem + C::tW2, args.dparams + S::W2, H, 1, inv_scale, wq, part, lane);
        flush_dw<H, OUT, 4>(tmem + C::tWO, args.dparams + S::WO, 1, H, inv_scale, wq, part, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ALLOC_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ------------------------------------------------------------------------------------------ launchers
template <int IN, int H, int OUT, int NH>
int launch_fwd_tc(const MlpFwdArgs& a, cudaStream_t st) {
    constexpr int G = fwd_groups<IN, H, OUT, NH>();
    static_assert(FwdCfg<IN, H, OUT, NH, G>::kFits, "forward MLP: shared memory / TMEM budget");
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
template <int IN, int H, int OUT, int NH, int NP>
int launch_bwd_tc_np(const MlpBwdArgs& a, cudaStream_t st) {
    using C = BwdCfg<IN, H, OUT, NH>;
    static bool configured = false;
    if (!configured) {
        AL_CHECK(cudaFuncSetAttribute(k_mlp_bwd_tc<IN, H, OUT, NH, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES));
        configured = true;
    }
    if (a.dx2 && !(C::kDxStage && a.dx_mode == 0)) {
        al_set_error("al_mlp_backward: a second d-x window needs the staged row-major path (shape in=%d hidden=%d)", IN, H);
        return (int)cudaErrorInvalidValue;
    }
    const long long tiles = ((long long)a.cap + 127) / 128;
    const int grid = (int)(tiles < al_num_sms() ? tiles : al_num_sms());
    k_mlp_bwd_tc<IN, H, OUT, NH, NP><<<grid, 128 * NP, C::BYTES, st>>>(a);
    AL_LAUNCH_CHECK();
    return 0;
}
template <int IN, int H, int OUT, int NH>
int launch_bwd_tc2(const MlpBwdArgs& a, cudaStream_t st) {
    using C = Bwd2Cfg<IN, H, OUT, NH>;
    static_assert(C::kFits, "two-tile backward: shared memory / TMEM budget");
    static bool configured = false;
    if (!configured) {
        AL_CHECK(cudaFuncSetAttribute(k_mlp_bwd_tc2<IN, H, OUT, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES));
        configured = true;
    }
    const long long tiles = ((long long)a.cap + 127) / 128;
    const long long want = (tiles + C::NG - 1) / C::NG;
    const int grid = (int)(want < al_num_sms() ? want : al_num_sms());
    MlpBwdArgs b = a;
    b.dbg = g_bwd_dbg;
    k_mlp_bwd_tc2<IN, H, OUT, NH><<<grid, C::NT, C::BYTES, st>>>(b);
    AL_LAUNCH_CHECK();
    return 0;
}
// Backward schedule: AL_BWD_SCHED=1 -> one tile in flight (k_mlp_bwd_tc), 2 (default) -> two tiles in flight.
static int bwd_sched() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("AL_BWD_SCHED");
        v = (e && e[0] == '1') ? 1 : 2;
    }
    return v;
}
// Column parts per TMEM lane quarter in the backward (threads = 128 * parts): AL_BWD_PARTS=2|4.
static int bwd_parts() {
    static int np = 0;
    if (np == 0) {
        const char* e = getenv("AL_BWD_PARTS");
        np = (e && e[0] == '2') ? 2 : 4;
    }
    return np;
}
template <int IN, int H, int OUT, int NH>
int launch_bwd_tc(const MlpBwdArgs& a, cudaStream_t st) {
    if constexpr (Bwd2Cfg<IN, H, OUT, NH>::kFits) {
        // the two-tile schedule takes a plain d-out matrix and the rank-1 / density forms of the heads' output gradient;
        // the materialised-gradient forms of the heads take the one-tile kernel
        const int k = a.spec.kind;
        const bool staged = k == 0 || k == 4 || ((k == 1 || k == 3) && a.spec.w) || (k == 2 && a.spec.w && OUT >= 32 && a.spec.F == OUT);
        if (bwd_sched() == 2 && staged) return launch_bwd_tc2<IN, H, OUT, NH>(a, st);
    }
    if (bwd_parts() == 4) return launch_bwd_tc_np<IN, H, OUT, NH, 4>(a, st);
    return launch_bwd_tc_np<IN, H, OUT, NH, 2>(a, st);
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

AL_API int al_set_bwd_debug(int bits) {
    const int prev = g_bwd_dbg;
    if (bits >= 0) g_bwd_dbg = bits;
    return prev;
}

int al_tc_mlp_forward(int in_pad, int hidden, int out_pad, int n_hidden, const MlpFwdArgs& a, cudaStream_t st) {
#define X(I, Hh, O, N) \
    if (in_pad == I && hidden == Hh && out_pad == O && n_hidden == N) return launch_fwd_tc<I, Hh, O, N>(a, st);
    AL_TC_CONFIGS(X)
#undef X
    return -1;
}
bool al_tc_mlp_has(int in_pad, int hidden, int out_pad, int n_hidden) {
#define X(I, Hh, O, N) \
    if (in_pad == I && hidden == Hh && out_pad == O && n_hidden == N) return true;
    AL_TC_CONFIGS(X)
#undef X
    return false;
}
int al_tc_mlp_backward(int in_pad, int hidden, int out_pad, int n_hidden, const MlpBwdArgs& a, cudaStream_t st) {
    if (a.dx && (a.dx_c0 & 1)) return -1;   // the pair-wise d-x store needs an even window start
    // staged source windows: slot capacity (BwdCfg::SLOT_W / NSLOT)
    if (a.spec.kind == 4 && out_pad != 16) return -1;
    if (a.spec.kind == 2 && (a.spec.F > out_pad || a.spec.F % 16 != 0)) return -1;
    if (a.spec.kind == 1 && a.spec.C > 16) return -1;
    if (a.spec.kind == 0 && a.dncols > out_pad) return -1;
#define X(I, Hh, O, N) \
    if (in_pad == I && hidden == Hh && out_pad == O && n_hidden == N) return launch_bwd_tc<I, Hh, O, N>(a, st);
    AL_TC_CONFIGS(X)
#undef X
    return -1;
}
