// The large layers of the wide MLP heads (512-d LSeg feature head: autolabel/models.py:115-136 with --feature-dim 512) on
// a TMA-fed, warp-specialised tcgen05 pipeline.  Same GEMMs and epilogues as k_gemm_tc (gemm_tc.cu), which stays the
// path for every shape this file does not take (K = 16 first layers, narrow outputs, unaligned operands):
//   F  forward   Y[M, N]   = act(X[M, K] W[N, K]^T)               both operands K-major
//   D  dgrad     dX[M, N]  = (dY[M, K] Wt[N, K]^T) * [mask > 0]    Wt = the transposed fp16 weight copy, K-major again
//   W  wgrad     G[P, Q]  += sum_s U[s, P] V[s, Q]                 samples = K, both operands MN-major
// Design reference for the reference's side of this: torch_ngp/ffmlp/src/ffmlp.cu:742-895 (CUTLASS GEMMs per layer).
//
// One persistent CTA per SM, 20 warps:
//   warp 0      producer: one lane issues cp.async.bulk.tensor (TMA) loads of 64-wide K chunks into a 3-stage (W: 4) ring of
//               128-byte-swizzled tiles (A 128 x 64, B 256 x 64 halfs; W mode: 64 x 64 boxes, samples along rows)
//   warp 1      MMA: one lane issues 4 tcgen05.mma (K = 16) per chunk into one of TWO 256-column TMEM accumulators and
//               commits the stage back to the producer; the last chunk also commits "accumulator full"
//   warps 4-19  epilogue: TMEM -> registers -> (ReLU | mask | scale) -> fp32 staging tile in shared memory -> coalesced
//               fp16 / fp32 window stores, 128 columns at a time; W mode: red.global.add straight from registers.
// The epilogue of item i overlaps the MMAs of item i + 1 (the other accumulator), and the TMA engine keeps three
// chunks in flight without spending LSU issue slots or L1 tag cycles on them.
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "gemm_args.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int kEpiWarps = 16;                     // 4 column parts per TMEM lane quarter
constexpr int kEpiWarp0 = 4, kEpiThreads = kEpiWarps * 32, kParts = kEpiWarps / 4;
constexpr int kThreads = kEpiWarp0 * 32 + kEpiThreads;
constexpr int kBK = 64;
constexpr uint32_t kATile = 128 * kBK * 2;        // 16 KB
constexpr uint32_t kBTile = 256 * kBK * 2;        // 32 KB
constexpr uint32_t kStageBytes = kATile + kBTile;
constexpr int kRow32 = 132;                       // words per staged fp32 row (128 columns + 4)
constexpr int kRow16 = 528;                       // bytes per staged fp16 row (256 columns + 16)
constexpr uint32_t kStagingBytes = 128 * kRow32 * 4;   // == 128 * kRow16
// Ring depth and layout per mode.  F / D: three 48 KB stages next to the epilogue's staging tile.  W has no staging tile
// (its epilogue reduces straight from registers) and takes a fourth stage: the main loops are bound by the bytes they keep
// in flight, and the fourth stage was worth 14 % there (a fourth stage for F / D by shrinking the staging tile to half /
// quarter passes was measured: -6 % on the plain layers, +60 % on the masked and windowed ones).
template <int MODE>
struct Ring {
    static constexpr int kStages = MODE == 2 ? 4 : 3;
    static constexpr uint32_t kStagingOff = kStages * kStageBytes;
    static constexpr uint32_t kBarOff = kStagingOff + (MODE == 2 ? 0u : kStagingBytes);
    static constexpr uint32_t kSmemBytes = kBarOff + 128 + 1024;   // + slack for the 1024-byte alignment of the ring
};

struct alignas(64) TmaArgs {
    CUtensorMap tmA, tmB;
    GemmArgs g;
};

// ---- PTX helpers local to the TMA path
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
// The same load delivered to the same shared-memory offset (and signalled on the same mbarrier offset) of every CTA in mask
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t mbar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(mbar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint32_t mbar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(mbar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait_wd(uint32_t mbar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_ns(mbar, parity, 1000u)) {
        if (++spins > 4000000u) __trap();            // > ~4 s
    }
}
// 128-byte-swizzled tiles (1024-byte aligned): K-major [rows][64 halfs], 8-row groups 1024 bytes apart;
// MN-major blocks of 64 MN elements x 64 K rows (8 KB), 8-row K groups 1024 bytes apart.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// Copy-out of the staged fp32 tile: this warp's rows (every kEpiWarps-th, `nr` of them) of the pass-local columns [lo, hi) to a
// window whose row `ew` / column 0 of the pass is `d`.  Everything row-invariant (vector width, activation) is decided
// once per window and pass: the row loop is a load, a store and two pointer steps.
template <int V>
__device__ __forceinline__ void copy_rows_f32(const float* __restrict__ s, float* __restrict__ d, size_t dstep, int lo, int hi,
                                              int nr, int act, float oscale, int lane) {
    const bool plain = act == 0 && oscale == 1.0f;
    #pragma unroll 2
    for (int i = 0; i < nr; ++i, s += kEpiWarps * kRow32, d += dstep) {
        for (int c = lo + lane * V; c < hi; c += 32 * V) {
            if (V == 4) {
                float4 v = *reinterpret_cast<const float4*>(s + c);
                if (!plain) v = make_float4(al_apply_act(v.x * oscale, act), al_apply_act(v.y * oscale, act),
                                            al_apply_act(v.z * oscale, act), al_apply_act(v.w * oscale, act));
                *reinterpret_cast<float4*>(d + c) = v;
            } else if (V == 2) {
                float2 v = *reinterpret_cast<const float2*>(s + c);
                if (!plain) v = make_float2(al_apply_act(v.x * oscale, act), al_apply_act(v.y * oscale, act));
                *reinterpret_cast<float2*>(d + c) = v;
            } else {
                const float v = s[c];
                d[c] = plain ? v : al_apply_act(v * oscale, act);
            }
        }
    }
}
__device__ __forceinline__ void copy_window_f32(const float* s, float* d, size_t ld, int lo, int hi, int nr, int act, float oscale,
                                                int lane) {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(d + lo);
    const int w = hi - lo;
    if (!(a0 & 15u) && !(ld & 3) && !(lo & 3) && !(w & 3)) copy_rows_f32<4>(s, d, (size_t)kEpiWarps * ld, lo, hi, nr, act, oscale, lane);
    else if (!(a0 & 7u) && !(ld & 1) && !(lo & 1) && !(w & 1)) copy_rows_f32<2>(s, d, (size_t)kEpiWarps * ld, lo, hi, nr, act, oscale, lane);
    else copy_rows_f32<1>(s, d, (size_t)kEpiWarps * ld, lo, hi, nr, act, oscale, lane);
}
template <int V, bool RELU>
__device__ __forceinline__ void copy_rows_f16(const float* __restrict__ s, __half* __restrict__ d, size_t dstep, int lo, int hi,
                                              int nr, int lane) {
    #pragma unroll 2
    for (int i = 0; i < nr; ++i, s += kEpiWarps * kRow32, d += dstep) {
        for (int c = lo + lane * V; c < hi; c += 32 * V) {
            if (V == 4) {
                float4 v = *reinterpret_cast<const float4*>(s + c);
                if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<uint2*>(d + c) = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
            } else {
                const float v = s[c];
                d[c] = __float2half_rn(RELU ? fmaxf(v, 0.f) : v);
            }
        }
    }
}
__device__ __forceinline__ void copy_window_f16(const float* s, __half* d, size_t ld, int lo, int hi, int nr, int act, int lane) {
    const bool vec = !(reinterpret_cast<uintptr_t>(d + lo) & 7u) && !(ld & 3) && !(lo & 3) && !((hi - lo) & 3);
    if (vec) { if (act == 1) copy_rows_f16<4, true>(s, d, (size_t)kEpiWarps * ld, lo, hi, nr, lane); else copy_rows_f16<4, false>(s, d, (size_t)kEpiWarps * ld, lo, hi, nr, lane); }
    else { if (act == 1) copy_rows_f16<1, true>(s, d, (size_t)kEpiWarps * ld, lo, hi, nr, lane); else copy_rows_f16<1, false>(s, d, (size_t)kEpiWarps * ld, lo, hi, nr, lane); }
}

// MODE 0: F / D (K-major operands); MODE 2: W (MN-major operands).  MASK: dgrad ReLU mask.  WIN: fp32 / fp16 output windows
// (else only the fp16 matrix Yh).
// CL = 2 (F / D only): the kernel runs in clusters of two CTAs that take the two 128-row tiles of a 256-row block against
// the SAME weight tile; each CTA loads half of that weight tile and multicasts it into both CTAs' shared memory, so the
// L2 -> SM operand traffic per chunk falls from 48 KB to 32 KB per CTA.  A stage is released to both producers by both
// MMA warps (the "empty" barriers count two multicast commits).
template <int MODE, bool MASK, bool WIN, int CL>
__global__ void __launch_bounds__(kThreads, 1) k_gemm_tma(const __grid_constant__ TmaArgs ta) {
    extern __shared__ unsigned char smem_raw[];
    constexpr int kStages = Ring<MODE>::kStages;
    constexpr uint32_t kStagingOff = Ring<MODE>::kStagingOff, kBarOff = Ring<MODE>::kBarOff;
    const GemmArgs& a = ta.g;
    const uint32_t s_raw = smem_u32(smem_raw);
    const uint32_t s0 = (s_raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (s0 - s_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
    // bars: [0, S) full, [S, 2S) empty, [2S, 2S + 2) accumulator full, [2S + 2, 2S + 4) accumulator empty
    const uint32_t b_full = s0 + kBarOff, b_empty = b_full + 8 * kStages, b_tfull = b_empty + 8 * kStages, b_tempty = b_tfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kBarOff + 8 * (2 * kStages + 4));
    uint32_t crank = 0;
    if (CL == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    if (tid == 0) {
        for (int i = 0; i < 2 * kStages + 2; ++i) mbar_init(smem_u32(&bars[i]), 1);
        if (CL == 2)
            for (int i = 0; i < kStages; ++i) mbar_init(b_empty + 8 * i, 2);   // both CTAs' MMA warps release a stage
        for (int i = 0; i < 2; ++i) mbar_init(b_tempty + 8 * i, kEpiThreads / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (CL == 2) cluster_sync_all();                              // the peer's barriers exist before anything arrives on them
    const uint32_t tmem = *tmem_slot;

    const long long n = a.n_dev ? min((long long)a.M, (long long)*a.n_dev) : (long long)a.M;
    // ---- work decomposition (the same for every role)
    const int bm = 128;
    const int n_tiles_n = (a.N + 255) / 256;
    // F / D: items = output tiles, strided over the CTAs (neighbouring CTAs share the A tile through L2).
    // W: an item is one split (rows_per_item samples) of one output tile.  CTA c owns tile c % T and the c / T-th
    // range of that tile's splits, accumulated in TMEM and flushed once: the T CTAs of a group walk the SAME sample
    // range at the same time, so every row of U and V comes from HBM once and from L2 for the other tiles.  (With
    // more tiles than CTAs the items fall back to contiguous (tile, split) ranges.)
    long long item_first, item_last, item_step;
    int n_split = 0, w_tile = -1;
    if (MODE == 2) {
        n_split = (int)((n + a.rows_per_item - 1) / a.rows_per_item);
        const int T = ((a.P + bm - 1) / bm) * n_tiles_n;
        const int groups = (int)gridDim.x / T;
        if (groups >= 1) {
            const int spg = (n_split + groups - 1) / groups;
            const int grp = (int)blockIdx.x / T;
            w_tile = (int)blockIdx.x % T;
            item_first = grp < groups ? (long long)grp * spg : 0;
            item_last = grp < groups ? min((long long)n_split, item_first + spg) : 0;
        } else {
            const long long items = (long long)T * n_split;
            const long long per_cta = (items + gridDim.x - 1) / gridDim.x;
            item_first = (long long)blockIdx.x * per_cta;
            item_last = min(items, item_first + per_cta);
        }
        item_step = 1;
    } else if (CL == 2) {
        // an item = a pair of 128-row tiles x one weight tile; CTA rank r of the cluster takes tile 2 * pair + r (a tile
        // past the live rows is computed and dropped: the two CTAs must stay in step)
        const long long m_pairs = ((n + 127) / 128 + 1) / 2;
        item_first = blockIdx.x / 2;
        item_last = m_pairs * n_tiles_n;
        item_step = gridDim.x / 2;
    } else {
        item_first = blockIdx.x;
        item_last = ((n + 127) / 128) * n_tiles_n;
        item_step = gridDim.x;
    }

    struct Item { long long m0, k_begin, k_end; int n0, bn, p0; bool acc_first, flush; };
    auto decode = [&](long long item) {
        Item it;
        if (MODE == 2) {
            const long long tile = w_tile >= 0 ? w_tile : item / n_split;
            const long long sp = w_tile >= 0 ? item : item - tile * n_split;
            const int tn = (int)(tile % n_tiles_n), tm = (int)(tile / n_tiles_n);
            it.n0 = tn * 256; it.bn = min(256, a.N - it.n0); it.p0 = tm * bm; it.m0 = 0;
            it.k_begin = sp * a.rows_per_item; it.k_end = min(n, it.k_begin + a.rows_per_item);
            it.acc_first = item == item_first || sp == 0;
            it.flush = item + 1 >= item_last || sp + 1 == n_split;
        } else {
            const int tn = (int)(item % n_tiles_n);
            it.m0 = (CL == 2 ? 2 * (item / n_tiles_n) + crank : item / n_tiles_n) * 128; it.n0 = tn * 256; it.bn = min(256, a.N - it.n0); it.p0 = 0;
            it.k_begin = 0; it.k_end = a.K; it.acc_first = true; it.flush = true;
        }
        return it;
    };

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            uint32_t chunk = 0;
            const uint32_t b_box_bytes = (uint32_t)min(256, a.N) * (kBK * 2);   // F / D: the B box holds min(256, N) weight rows
            for (long long item = item_first; item < item_last; item += item_step) {
                const Item it = decode(item);
                const int nk = (int)((it.k_end - it.k_begin + kBK - 1) / kBK);
                for (int kc = 0; kc < nk; ++kc, ++chunk) {
                    const int s = chunk % kStages;
                    const uint32_t ph = (chunk / kStages) & 1u;
                    mbar_wait_wd(b_empty + 8 * s, ph ^ 1u);                    // a fresh barrier passes the parity-1 wait
                    const uint32_t dA = s0 + s * kStageBytes, dB = dA + kATile, full = b_full + 8 * s;
                    const int k0 = (int)(it.k_begin + (long long)kc * kBK);
                    if (MODE == 2) {
                        const int nb = (it.bn + 63) / 64;             // whole boxes; columns past N / P are zero-filled
                        mbar_expect_tx(full, (uint32_t)(2 + nb) * 8192u);
                        for (int i = 0; i < 2; ++i) tma_load_2d(dA + i * 8192u, &ta.tmA, it.p0 + 64 * i, k0, full);
                        for (int i = 0; i < nb; ++i) tma_load_2d(dB + i * 8192u, &ta.tmB, it.n0 + 64 * i, k0, full);
                    } else {
                        mbar_expect_tx(full, kATile + b_box_bytes);             // whole boxes always (out-of-range parts are zero-filled)
                        tma_load_2d(dA, &ta.tmA, k0, (int)it.m0, full);
                        if (CL == 2) {                                          // my half of the weight tile, to both CTAs
                            const uint32_t hb = (uint32_t)min(256, a.N) / 2;
                            tma_load_2d_mc(dB + crank * hb * (kBK * 2), &ta.tmB, k0, it.n0 + (int)(crank * hb), full, (uint16_t)3);
                        } else {
                            tma_load_2d(dB, &ta.tmB, k0, it.n0, full);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            uint32_t chunk = 0, tile_iter = 0;
            for (long long item = item_first; item < item_last; item += item_step) {
                const Item it = decode(item);
                const int nk = (int)((it.k_end - it.k_begin + kBK - 1) / kBK);
                const uint32_t acc = tile_iter & 1u;
                const uint32_t d_tmem = tmem + acc * 256u;
                if (it.acc_first) {
                    mbar_wait_wd(b_tempty + 8 * acc, ((tile_iter >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
                    tc_fence_after();
                }
                const uint32_t idesc = make_idesc(bm, it.bn, MODE == 2, MODE == 2);
                for (int kc = 0; kc < nk; ++kc, ++chunk) {
                    const int s = chunk % kStages;
                    const uint32_t ph = (chunk / kStages) & 1u;
                    mbar_wait_wd(b_full + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t dA = s0 + s * kStageBytes, dB = dA + kATile;
                    uint64_t da, db, step;
                    if (MODE == 2) { da = desc_sw128(dA, 8192u, 1024u); db = desc_sw128(dB, 8192u, 1024u); step = 2048u >> 4; }
                    else { da = desc_sw128(dA, 16u, 1024u); db = desc_sw128(dB, 16u, 1024u); step = 32u >> 4; }
                    #pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        mma_f16(d_tmem, da, db, idesc, (!it.acc_first || kc > 0 || k > 0) ? 1u : 0u);
                        da += step; db += step;
                    }
                    if (CL == 2) mma_commit_mc(b_empty + 8 * s, (uint16_t)3);
                    else mma_commit(b_empty + 8 * s);
                }
                if (it.flush) {
                    if (nk > 0) mma_commit(b_tfull + 8 * acc);
                    else mbar_arrive(b_tfull + 8 * acc);                      // nothing was issued: release the epilogue by hand
                    ++tile_iter;
                }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================================================================== epilogue
        const int ew = warp - kEpiWarp0;                       // 0 .. kEpiWarps - 1
        const int wq = warp & 3, part = ew >> 2;               // TMEM lane quarter (hardware: warp id mod 4), column part
        const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
        float* stage = reinterpret_cast<float*>(smem + kStagingOff);
        const float scale = (a.amax_dev != nullptr) ? al_grad_scale(a.amax_dev) : 1.0f;
        const float inv_scale = 1.0f / scale;
        uint32_t tile_iter = 0;
        for (long long item = item_first; item < item_last; item += item_step) {
            const Item it = decode(item);
            if (!it.flush) continue;
            const uint32_t acc = tile_iter & 1u;
            const uint32_t t_acc = tmem + acc * 256u + lane_sel;
            if (MODE != 2 && MASK && !WIN) {
                // The ReLU-mask block of this warp (32 rows x 256 / kParts columns) goes into the warp's own part of the
                // staging tile BEFORE the accumulator is waited for: read coalesced (8 lanes per 128-byte row segment)
                // instead of 32 bytes per lane from 32 different rows, and each thread later overwrites exactly the
                // 16-byte units it has consumed.  (The previous item's copy-out ended with the epilogue barrier.)
                constexpr int kUnits = (256 / kParts) / 8;         // 16-byte units per row of the warp's block
                unsigned char* st16 = smem + kStagingOff;
                for (int q = lane; q < 32 * kUnits; q += 32) {
                    const int rr = q / kUnits, ch = q - rr * kUnits;
                    const int col = part * (256 / kParts) + ch * 8;
                    const long long grow = it.m0 + wq * 32 + rr;
                    uint4 m = make_uint4(0, 0, 0, 0);
                    if (grow < n && col < it.bn) m = __ldg(reinterpret_cast<const uint4*>(a.mask + (size_t)grow * a.ldmask + it.n0 + col));
                    *reinterpret_cast<uint4*>(st16 + (wq * 32 + rr) * kRow16 + col * 2) = m;
                }
                __syncwarp();
            }
            mbar_wait_wd(b_tfull + 8 * acc, (tile_iter >> 1) & 1u);
            tc_fence_after();
            ++tile_iter;
            if (MODE == 2) {
                const int prow = it.p0 + wq * 32 + lane;
                for (int c = part * 16; c < it.bn; c += 16 * kParts) {
                    uint32_t v[16];
                    tmem_ld16(t_acc + c, v);
                    tmem_ld_wait();
                    if (prow < a.P) {                            // P not a multiple of 128: the tile's last rows are padding
                        #pragma unroll
                        for (int j = 0; j < 16; ++j)
                            atomicAdd(a.G + (size_t)prow * a.sp + (size_t)(it.n0 + c + j) * a.sq, __uint_as_float(v[j]) * inv_scale);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_tempty + 8 * acc);
                continue;
            }
            const int row_l = wq * 32 + lane;
            const long long row = it.m0 + row_l;
            // one 16-column accumulator chunk of this thread's row -> ReLU / mask, in registers
            auto load_chunk = [&](int c, float (&f)[16]) {
                uint32_t v[16];
                tmem_ld16(t_acc + c, v);
                uint4 m0v = make_uint4(0, 0, 0, 0), m1v = m0v;
                if (MASK && !WIN) {                            // staged by this warp above
                    const uint4* mp = reinterpret_cast<const uint4*>(smem + kStagingOff + row_l * kRow16 + c * 2);
                    m0v = mp[0]; m1v = mp[1];
                } else if (MASK && row < n) {
                    const uint4* mp = reinterpret_cast<const uint4*>(a.mask + (size_t)row * a.ldmask + it.n0 + c);
                    m0v = __ldg(mp); m1v = __ldg(mp + 1);
                }
                tmem_ld_wait();
                #pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                if (a.relu) {
                    #pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                if (MASK) {
                    const uint32_t mw[8] = {m0v.x, m0v.y, m0v.z, m0v.w, m1v.x, m1v.y, m1v.z, m1v.w};
                    #pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 mm = __half22float2(*reinterpret_cast<const __half2*>(&mw[j]));
                        f[2 * j] = mm.x > 0.f ? f[2 * j] : 0.f;
                        f[2 * j + 1] = mm.y > 0.f ? f[2 * j + 1] : 0.f;
                    }
                }
            };
            auto release_acc = [&]() {                         // last TMEM read of this item: the MMA warp may reuse the accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_tempty + 8 * acc);
            };
            if (!WIN) {
                // ---- fp16 matrix only: the whole [128 x bn] tile is staged as halfs (row stride 528 bytes: 16-byte stores of
                // lanes = rows stay conflict free), then leaves as 16 bytes per lane, one 512-byte row per store instruction
                unsigned char* st16 = smem + kStagingOff;
                #pragma unroll 2
                for (int cc = 0; cc < 256 / kParts; cc += 16) {
                    const int c = part * (256 / kParts) + cc;
                    if (c >= it.bn) break;                     // warp-uniform
                    float f[16];
                    load_chunk(c, f);
                    uint4 o0v, o1v;
                    o0v.x = pack_h2(f[0], f[1]); o0v.y = pack_h2(f[2], f[3]); o0v.z = pack_h2(f[4], f[5]); o0v.w = pack_h2(f[6], f[7]);
                    o1v.x = pack_h2(f[8], f[9]); o1v.y = pack_h2(f[10], f[11]); o1v.z = pack_h2(f[12], f[13]); o1v.w = pack_h2(f[14], f[15]);
                    uint4* sp = reinterpret_cast<uint4*>(st16 + row_l * kRow16 + c * 2);
                    sp[0] = o0v; sp[1] = o1v;
                }
                release_acc();
                named_bar(1, kEpiThreads);
                if (lane < (it.bn >> 3)) {                      // 16-byte units per row
                    const int valid = (int)min(128LL, n - it.m0);
                    const int nr = valid > ew ? (valid - ew + kEpiWarps - 1) / kEpiWarps : 0;
                    const unsigned char* sp = st16 + ew * kRow16 + lane * 16;
                    __half* dp = a.Yh + (size_t)(it.m0 + ew) * a.ldyh + it.n0 + lane * 8;
                    const size_t dstep = (size_t)kEpiWarps * a.ldyh;
                    #pragma unroll 4
                    for (int i = 0; i < nr; ++i, sp += kEpiWarps * kRow16, dp += dstep)
                        *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<const uint4*>(sp);
                }
                named_bar(1, kEpiThreads);                     // staging free for the next item
                continue;
            }
            // ---- windows: fp32 staging, 128 columns per pass (row stride 132 words: float4 stores of lanes = rows and float4
            // reads along a row are both conflict free), vector stores where a window's alignment allows
            const float oscale = a.unscale ? inv_scale : 1.0f;
            #pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int cbase = half * 128;                  // tile-local first column of this pass
                if (cbase >= it.bn) break;                     // narrow tile: pass 0 already released the accumulator
                #pragma unroll
                for (int cc = 0; cc < 128 / kParts; cc += 16) {
                    const int c = cbase + part * (128 / kParts) + cc;
                    if (c >= it.bn) break;                     // warp-uniform
                    float f[16];
                    load_chunk(c, f);
                    float4* sp = reinterpret_cast<float4*>(stage + row_l * kRow32 + (c - cbase));
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) sp[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                }
                if (half == 1 || cbase + 128 >= it.bn) release_acc();
                named_bar(1, kEpiThreads);
                // ---- staging -> global: warp ew takes rows ew, ew + kEpiWarps, ...
                const int ncol = min(128, it.bn - cbase);
                const int g0 = it.n0 + cbase;                  // first global output column of this pass
                auto range = [&](int src0, int ncols, int& lo, int& hi) {   // pass-local columns [lo, hi) of a window
                    lo = max(src0, g0) - g0;
                    hi = min(src0 + ncols, g0 + ncol) - g0;
                };
                int lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0, loh = 0, hih = 0;
                if (a.o0.ptr) range(a.o0.src0, a.o0.ncols, lo0, hi0);
                if (a.o1.ptr) range(a.o1.src0, a.o1.ncols, lo1, hi1);
                if (a.h0.ptr) range(a.h0.src0, a.h0.ncols, loh, hih);
                {
                    const int valid = (int)min(128LL, n - it.m0);
                    const int nr = valid > ew ? (valid - ew + kEpiWarps - 1) / kEpiWarps : 0;
                    const float* s = stage + ew * kRow32;
                    const size_t grow = (size_t)(it.m0 + ew);
                    if (a.Yh) copy_window_f16(s, a.Yh + grow * a.ldyh + g0, a.ldyh, 0, ncol, nr, 0, lane);
                    if (hi0 > lo0) copy_window_f32(s, a.o0.ptr + grow * a.o0.ld + a.o0.col0 + (g0 - a.o0.src0), a.o0.ld, lo0, hi0, nr, a.o0.act, oscale, lane);
                    if (hi1 > lo1) copy_window_f32(s, a.o1.ptr + grow * a.o1.ld + a.o1.col0 + (g0 - a.o1.src0), a.o1.ld, lo1, hi1, nr, a.o1.act, oscale, lane);
                    if (hih > loh) copy_window_f16(s, a.h0.ptr + grow * a.h0.ld + a.h0.col0 + (g0 - a.h0.src0), a.h0.ld, loh, hih, nr, a.h0.act, lane);
                }
                named_bar(1, kEpiThreads);                     // staging free for the next pass / item
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();                              // no CTA leaves while its peer can still write into it
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// Row-major fp16 matrix [rows, cols] with leading dimension ld (halfs); box = box_cols x box_rows, 128-byte swizzle.
bool make_map(CUtensorMap* m, const __half* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Rows [n, round_up(n, 64)) of a [cap, ncols] fp16 matrix <- 0: the wgrad reads whole 64-sample chunks, and what lies
// behind the live rows is dead storage (possibly never written).
__global__ void k_zero_tail(__half* __restrict__ m, int ld, int ncols, int cap, const int* __restrict__ n_dev) {
    const long long n = n_dev ? min((long long)cap, (long long)*n_dev) : (long long)cap;
    const long long end = min((long long)cap, (n + 63) / 64 * 64);
    const long long total = (end - n) * ncols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = n + i / ncols;
        m[(size_t)r * ld + (int)(i % ncols)] = __float2half_rn(0.f);
    }
}

template <int MODE, bool MASK, bool WIN, int CL>
int launch_t(const TmaArgs& ta, int grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        AL_CHECK(cudaFuncSetAttribute(k_gemm_tma<MODE, MASK, WIN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Ring<MODE>::kSmemBytes));
        configured = true;
    }
    if (CL == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Ring<MODE>::kSmemBytes; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        AL_CHECK(cudaLaunchKernelEx(&cfg, k_gemm_tma<MODE, MASK, WIN, CL>, ta));
    } else {
        k_gemm_tma<MODE, MASK, WIN, CL><<<grid, kThreads, Ring<MODE>::kSmemBytes, st>>>(ta);
    }
    AL_LAUNCH_CHECK();
    return 0;
}

// Off by default: measured neutral on the C5 GEMMs (profiles/r3_gemm_tma_progress.md) -- with three 48 KB stages the
// kernel is bound by the bytes it can keep in flight per SM, not by the L2 read bandwidth the multicast saves.
// AL_GEMM_CLUSTER=1 selects it.
bool cluster_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AL_GEMM_CLUSTER");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

bool tma_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("AL_GEMM_TMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

}  // namespace

int al_gemm_tma_launch(const GemmArgs& a, cudaStream_t st) {
    if (!tma_enabled() || a.M <= 0) return -1;
    TmaArgs ta;
    ta.g = a;
    long long items;
    bool cl2 = false;
    const int tn = (a.N + 255) / 256;
    if (a.mode == 2) {
        // tiles of 128 x 256 outputs; partial tiles load zero-filled boxes and skip the padding in the epilogue
        if (a.P % 16 != 0 || a.N % 16 != 0 || a.rows_per_item % 64 != 0) return -1;
        if (a.lda % 8 != 0 || a.ldb % 8 != 0 || !aligned16(a.A) || !aligned16(a.B)) return -1;
        if (!make_map(&ta.tmA, a.A, a.M, a.P, a.lda, 64, 64) || !make_map(&ta.tmB, a.B, a.M, a.N, a.ldb, 64, 64)) return -1;
        k_zero_tail<<<8, 256, 0, st>>>(const_cast<__half*>(a.A), a.lda, a.P, a.M, a.n_dev);
        AL_LAUNCH_CHECK();
        k_zero_tail<<<8, 256, 0, st>>>(const_cast<__half*>(a.B), a.ldb, a.N, a.M, a.n_dev);
        AL_LAUNCH_CHECK();
        items = (long long)((a.P + 127) / 128) * tn * (((long long)a.M + a.rows_per_item - 1) / a.rows_per_item);
    } else {
        const __half* B = a.mode == 1 ? a.Bt : a.B;
        const int ldb = a.mode == 1 ? a.ldbt : a.ldb;
        if (!B || a.K < 16 || a.N < 16 || a.N % 16 != 0) return -1;   // short K / narrow N: the boxes' out-of-range parts are zero-filled
        if (a.lda % 8 != 0 || ldb % 8 != 0 || !aligned16(a.A) || !aligned16(B)) return -1;
        if (a.Yh && (a.ldyh % 8 != 0 || !aligned16(a.Yh))) return -1;
        if (!a.Yh && !(a.o0.ptr || a.o1.ptr || a.h0.ptr)) return -1;
        if (a.mask && (a.ldmask % 8 != 0 || !aligned16(a.mask))) return -1;
        const int b_rows = a.N < 256 ? a.N : 256;
        // clusters of two CTAs (multicast weight tiles) when there are at least two 128-row tiles and the weight tile
        // splits into two halves of whole 8-row swizzle groups
        cl2 = cluster_enabled() && a.M > 128 && b_rows % 16 == 0 && a.K >= 64;
        if (!make_map(&ta.tmA, a.A, a.M, a.K, a.lda, 64, 128) || !make_map(&ta.tmB, B, a.N, a.K, ldb, 64, cl2 ? b_rows / 2 : b_rows)) return -1;
        items = (((long long)a.M + 127) / 128) * tn;
    }
    if (items <= 0) return 0;
    int grid = (int)(items < al_num_sms() ? items : al_num_sms());
    if (a.mode == 2) return launch_t<2, false, false, 1>(ta, grid, st);
    const bool win = a.o0.ptr || a.o1.ptr || a.h0.ptr;
    if (cl2) {
        grid = (grid + 1) & ~1;                                 // whole clusters (a tile past the live rows is dropped)
        if (grid > al_num_sms()) grid = al_num_sms() & ~1;
        if (a.mask) return win ? launch_t<0, true, true, 2>(ta, grid, st) : launch_t<0, true, false, 2>(ta, grid, st);
        return win ? launch_t<0, false, true, 2>(ta, grid, st) : launch_t<0, false, false, 2>(ta, grid, st);
    }
    if (a.mask) return win ? launch_t<0, true, true, 1>(ta, grid, st) : launch_t<0, true, false, 1>(ta, grid, st);
    return win ? launch_t<0, false, true, 1>(ta, grid, st) : launch_t<0, false, false, 1>(ta, grid, st);
}
