// Wide MLP heads (hidden / output widths beyond what fits the weight-resident fused kernels of mlp_tc.cu):
// a tiled tcgen05 GEMM with fused epilogues, run layer by layer with fp16 activations in HBM.
//
// Replaces tiny-cuda-nn's CutlassMLP as autolabel uses it for the 512-d LSeg feature head
// (autolabel/models.py:115-136 with --feature-dim 512; scripts/language/*) and for ScanNet label sets
// (semantic_out 64 -> up to 606 classes, scripts/convert_scannet.py:117-121); design reference
// torch_ngp/ffmlp/src/ffmlp.cu:742-895 (CUTLASS GEMMs for layers wider than the fused kernel, split-K weight
// gradients).
//
// One kernel, three modes, all operands fp16 row-major matrices in global memory, accumulators in TMEM:
//   F  forward   Y[M, N]   = act(X[M, K] W[N, K]^T)              A, B K-major tiles
//   D  dgrad     dX[M, N]  = (dY[M, K] W[K, N]) * [mask > 0]      A K-major, B = W tile viewed MN-major
//   W  wgrad     G[P, Q]  += sum_s U[s, P] V[s, Q]                samples = K, both tiles viewed MN-major,
//                                                                 split over sample ranges, red.global.add
// A work item is a 128-row (64 for wgrad of a 64-wide side) x BN-column output tile; the K loop streams 64-wide
// chunks of both operands through a 3-stage cp.async ring into the un-swizzled canonical UMMA layout (the tile
// loader and views of tc_common.cuh), one thread issues the MMAs, tcgen05.commit frees a stage.
#include "common.cuh"
#include "mlp_args.cuh"
#include "tc_common.cuh"
#include "gemm_args.cuh"
#include "../../include/autolabel_b200.h"

namespace {
using namespace tc;

constexpr int kThreads = 512;                    // 16 warps: 4 column parts per TMEM lane quarter in the epilogues
constexpr int kColStep = 16 * (kThreads / 128);   // epilogue: column stride of one part
constexpr int kStages = 3;
constexpr int kBK = 64;
constexpr uint32_t kATile = 128 * kBK * 2;        // 16 KB
constexpr uint32_t kBTile = 256 * kBK * 2;        // 32 KB
constexpr uint32_t kStageBytes = kATile + kBTile;
constexpr uint32_t kRowPad = 16;                  // bytes added to every staged row: lanes = rows stay bank-conflict free
constexpr uint32_t kMaskBytes = 128 * (256 * 2 + kRowPad);   // dgrad: the ReLU-mask tile [128 x bn] fp16, prefetched
constexpr uint32_t kMaskOff = kStages * kStageBytes + 64;
constexpr uint32_t kSmemBytes = kMaskOff + kMaskBytes;

// [nrows x ncols] block of a row-major fp16 matrix at (row0, col0) -> canonical tile with `ncols` columns;
// rows >= row_limit are zero-filled.  ncols is a multiple of 8.
__device__ __forceinline__ void load_tile_async(const __half* __restrict__ src, size_t ld, long long row0, int nrows,
                                                long long row_limit, int col0, int ncols, uint32_t dst, int tid) {
    // 16-byte unit q of the tile lives at byte q * 16 of the canonical layout: q = ((r / 8) * chunks + ch) * 8 + r % 8.
    // Consecutive threads take consecutive q: shared-memory writes are linear (no bank conflicts; the row-major order
    // "consecutive threads on consecutive chunks of a row" puts 8 threads on the same banks, stride 128 B), global reads
    // stay sector-efficient (a warp covers 8 rows x 64 contiguous bytes).  nrows is a multiple of 8.
    const int chunks = ncols >> 3;
    const int total = nrows * chunks;
    for (int q = tid; q < total; q += kThreads) {
        const int t = q >> 3;
        const int rb = t / chunks, ch = t - rb * chunks;
        const int r = rb * 8 + (q & 7);
        const bool valid = row0 + r < row_limit;
        const __half* p = valid ? src + (size_t)(row0 + r) * ld + col0 + ch * 8 : src;
        cp_async16(dst + (uint32_t)q * 16u, p, valid);
    }
}

// MODE 0 F, 1 D, 2 W; MASK: dgrad with a ReLU mask; WIN: fp32 / fp16 output windows (else the fp16 matrix Yh, or both).
// The epilogue variants are compile-time: the per-chunk instruction stream of the drain loop is what paces an item.
template <int MODE, bool MASK, bool WIN>
__global__ void __launch_bounds__(kThreads, 1) k_gemm_tc(const GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x;
    const int wq = (tid >> 5) & 3, part = tid >> 7, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);      // [kStages] stage free, [kStages] = tile done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * kStageBytes + 40);
    if (tid == 0) {
        for (int i = 0; i <= kStages; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    const uint32_t s0 = smem_u32(smem);
    uint32_t stage_par = 0;                                   // bit s: parity to wait for on bars[s]
    uint32_t done_par = 0;
    uint32_t stage_used = 0;                                  // bit s: the stage has an uncommitted-or-unwaited MMA batch

    const long long n = a.n_dev ? min((long long)a.M, (long long)*a.n_dev) : (long long)a.M;
    const float scale = (MODE != 0) ? al_grad_scale(a.amax_dev) : 1.0f;
    const float inv_scale = 1.0f / scale;

    // ---- work decomposition
    const int bm = (MODE == 2) ? ((a.P % 128 == 0) ? 128 : 64) : 128;
    const int n_tiles_n = (a.N + 255) / 256;
    long long items;
    int n_tiles_m = 0, n_split = 0;
    if (MODE == 2) {
        n_tiles_m = a.P / bm;
        n_split = (int)((n + a.rows_per_item - 1) / a.rows_per_item);
        items = (long long)n_tiles_m * n_tiles_n * n_split;
    } else {
        items = ((n + 127) / 128) * n_tiles_n;
    }

    // F / D: items strided over the CTAs.  W: every CTA takes a CONTIGUOUS range of items ordered (tile, split), so the
    // splits of one output tile that land on the same CTA accumulate in TMEM and are flushed (red.global.add) once.
    const long long per_cta = (items + gridDim.x - 1) / gridDim.x;
    const long long item_first = MODE == 2 ? (long long)blockIdx.x * per_cta : (long long)blockIdx.x;
    const long long item_last = MODE == 2 ? min(items, item_first + per_cta) : items;
    const long long item_step = MODE == 2 ? 1 : (long long)gridDim.x;
    for (long long item = item_first; item < item_last; item += item_step) {
        // ---- this item's tile and K range
        long long m0;                 // F/D: first sample row.  W: first sample of the split
        int n0, bn, p0 = 0;
        long long k_begin, k_end;     // F/D: reduction columns.  W: sample rows
        bool acc_first = true, flush = true;   // W: start a new accumulation / write the tile out after this item
        if (MODE == 2) {
            const long long tile = item / n_split;
            const long long sp = item - tile * n_split;
            const int tn = (int)(tile % n_tiles_n);
            const int tm = (int)(tile / n_tiles_n);
            n0 = tn * 256; bn = min(256, a.N - n0); p0 = tm * bm;
            k_begin = sp * a.rows_per_item; k_end = min(n, k_begin + a.rows_per_item);
            m0 = 0;
            acc_first = item == item_first || sp == 0;
            flush = item + 1 >= item_last || sp + 1 == n_split;
        } else {
            const int tn = (int)(item % n_tiles_n);
            m0 = (item / n_tiles_n) * 128;
            n0 = tn * 256; bn = min(256, a.N - n0);
            k_begin = 0; k_end = a.K;
        }
        const int nk = (int)((k_end - k_begin + kBK - 1) / kBK);
        const uint32_t idesc = make_idesc(bm, bn, MODE == 2, MODE != 0);
        const uint32_t row_bytes = (uint32_t)bn * 2u + kRowPad;           // staged fp16 row (mask tile, output tile)
        if (MODE != 2 && MASK) {
            // the ReLU-mask tile of this item, coalesced (a warp reads whole rows), in its own (oldest) cp.async group:
            // it has landed by the time the first operand chunk has
            const int chunks = bn >> 3;
            for (int q = tid; q < 128 * chunks; q += kThreads) {
                const int r = q / chunks, ch = q - r * chunks;
                const bool valid = m0 + r < n;
                const __half* src = valid ? a.mask + (size_t)(m0 + r) * a.ldmask + n0 + ch * 8 : a.mask;
                cp_async16(s0 + kMaskOff + (uint32_t)r * row_bytes + (uint32_t)ch * 16u, src, valid);
            }
            cp_async_commit();
        }

        // ---- per-thread copy plan of a FULL 64-wide chunk, computed once per item: the 16-byte units this thread moves
        // (2048 / kThreads at most per tile), each a base pointer at k = 0 plus a stride per unit of k.  Per chunk only
        // "base + k0 * stride" remains: the address arithmetic of 3072 cp.async per chunk (divisions, 64-bit
        // multiplies) would otherwise cost more issue slots than the MMAs of the chunk take.
        //   K along columns (F: A, B;  D: A):   row fixed, column = k0 + ..  -> stride 1,  validity fixed
        //   K along rows    (D: B;  W: A, B):   row = k0 + r                -> stride ld, validity r < limit - k0
        constexpr int kUnits = 2048 / kThreads;   // 16-byte units of the largest tile (256 x 64 halfs) per thread
        struct Plan { const __half* base[kUnits]; int r[kUnits]; int count; long long kmul; long long limit; bool krows; };
        auto make_plan = [&](const __half* mat, long long ld, bool krows, long long rowfix, long long colfix, int nrows,
                             int ncols, long long row_limit) {
            Plan pl;
            pl.krows = krows; pl.kmul = krows ? ld : 1; pl.limit = row_limit; pl.count = 0;
            const int chunks = ncols >> 3, total = nrows * chunks;
            #pragma unroll
            for (int i = 0; i < kUnits; ++i) {
                const int q = tid + i * kThreads;
                pl.base[i] = mat; pl.r[i] = 0;
                if (q < total) {
                    const int t = q >> 3;
                    const int rb = t / chunks, ch = t - rb * chunks;
                    const int r = rb * 8 + (q & 7);
                    pl.base[i] = mat + (krows ? (long long)r : rowfix + r) * ld + colfix + ch * 8;
                    pl.r[i] = krows ? r : (rowfix + r < row_limit ? 0 : 1 << 30);     // fixed validity folded into r
                    pl.count = i + 1;
                }
            }
            return pl;
        };
        auto run_plan = [&](const Plan& pl, long long k0, uint32_t dst) {
            const long long lim = pl.krows ? pl.limit - k0 : (long long)(1 << 29);
            #pragma unroll
            for (int i = 0; i < kUnits; ++i) {
                if (i < pl.count) {
                    const bool valid = pl.r[i] < lim;
                    cp_async16(dst + (uint32_t)(tid + i * kThreads) * 16u, valid ? pl.base[i] + k0 * pl.kmul : pl.base[i], valid);
                }
            }
        };
        Plan planA, planB;
        if (MODE == 0) {
            planA = make_plan(a.A, a.lda, false, m0, 0, 128, kBK, n);
            planB = make_plan(a.B, a.ldb, false, n0, 0, bn, kBK, a.N);
        } else if (MODE == 1) {
            planA = make_plan(a.A, a.lda, false, m0, 0, 128, kBK, n);
            planB = make_plan(a.B, a.ldb, true, 0, n0, kBK, bn, a.K);
        } else {
            planA = make_plan(a.A, a.lda, true, 0, p0, kBK, bm, n);
            planB = make_plan(a.B, a.ldb, true, 0, n0, kBK, bn, n);
        }

        auto load_chunk = [&](int kc) {
            const int s = kc % kStages;
            const uint32_t dA = s0 + s * kStageBytes, dB = dA + kATile;
            const long long k0 = k_begin + (long long)kc * kBK;
            const int kw = (int)min((long long)kBK, k_end - k0);                 // F/D: multiple of 16
            if (kw == kBK || MODE == 2) {                                       // W always stages 64 rows (zero-filled past n)
                run_plan(planA, k0, dA);
                run_plan(planB, k0, dB);
            } else if (MODE == 0) {
                load_tile_async(a.A, a.lda, m0, 128, n, (int)k0, kw, dA, tid);
                load_tile_async(a.B, a.ldb, n0, bn, a.N, (int)k0, kw, dB, tid);
            } else {
                load_tile_async(a.A, a.lda, m0, 128, n, (int)k0, kw, dA, tid);
                load_tile_async(a.B, a.ldb, k0, kw, a.K, n0, bn, dB, tid);
            }
            cp_async_commit();
        };
        auto wait_stage_free = [&](int s) {
            if ((stage_used >> s) & 1u) {
                mbar_wait(smem_u32(&bars[s]), (stage_par >> s) & 1u);
                stage_par ^= 1u << s;
                stage_used &= ~(1u << s);
            }
        };

        // ---- prologue: fill up to kStages - 1 stages
        for (int kc = 0; kc < kStages - 1; ++kc) {
            if (kc < nk) { wait_stage_free(kc % kStages); load_chunk(kc); }
            else cp_async_commit();
        }
        for (int kc = 0; kc < nk; ++kc) {
            cp_async_wait_group<kStages - 2>();                // chunk kc has landed (this thread's copies); kc + 1 may be in flight
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            const int s = kc % kStages;
            if (tid == 0) {
                tc_fence_after();
                const uint32_t dA = s0 + s * kStageBytes, dB = dA + kATile;
                const long long k0 = k_begin + (long long)kc * kBK;
                const int kw = (MODE == 2) ? kBK : (int)min((long long)kBK, k_end - k0);
                Operand oa, ob;
                if (MODE == 0) { oa = view_k(dA, kw); ob = view_k(dB, kw); }
                else if (MODE == 1) { oa = view_k(dA, kw); ob = view_mn(dB, bn); }
                else { oa = view_mn(dA, bm); ob = view_mn(dB, bn); }
                for (int k = 0; k < kw / 16; ++k) {
                    const uint64_t da = make_desc(oa.addr + k * oa.kstep, oa.lbo, oa.sbo);
                    const uint64_t db = make_desc(ob.addr + k * ob.kstep, ob.lbo, ob.sbo);
                    mma_f16(tmem, da, db, idesc, (!acc_first || kc > 0 || k > 0) ? 1u : 0u);
                }
                mma_commit(smem_u32(&bars[s]));
                if (kc == nk - 1 && flush) mma_commit(smem_u32(&bars[kStages]));
            }
            stage_used |= 1u << s;
            // With MMA(kc) queued behind MMA(kc - 1), refill the stage chunk kc - 1 used: the tensor pipe always has the next
            // batch waiting while the threads sit on the mbarrier of the previous one.
            const int kn = kc + kStages - 1;
            if (kn < nk) { wait_stage_free(kn % kStages); load_chunk(kn); }
            else cp_async_commit();
        }
        if (nk == 0 || !flush) continue;     // W: the next item of this CTA continues the same accumulation
        mbar_wait(smem_u32(&bars[kStages]), done_par);
        done_par ^= 1;
        tc_fence_after();

        // ---- epilogue
        if (MODE == 2) {
            const int prow = bm == 128 ? wq * 32 + lane : wq * 16 + lane;
            const bool valid = bm == 128 || lane < 16;
            for (int c = part * 16; c < bn; c += kColStep) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_sel + c, v);
                tmem_ld_wait();
                if (valid) {
                    #pragma unroll
                    for (int j = 0; j < 16; ++j)
                        atomicAdd(a.G + (size_t)(p0 + prow) * a.sp + (size_t)(n0 + c + j) * a.sq, __uint_as_float(v[j]) * inv_scale);
                }
            }
        } else {
            const long long row = m0 + wq * 32 + lane;
            const float oscale = a.unscale ? inv_scale : 1.0f;
            constexpr bool windows = WIN;
            const bool stage16 = a.Yh && !windows;
            float* stage = reinterpret_cast<float*>(smem);
            const int sstride = bn + 1;
            for (int c = part * 16; c < bn; c += kColStep) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_sel + c, v);
                tmem_ld_wait();
                if (row < n) {
                    float f[16];
                    #pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                    if (a.relu) {
                        #pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    if (MASK) {
                        const uint4* mp = reinterpret_cast<const uint4*>(smem + kMaskOff + (size_t)(wq * 32 + lane) * row_bytes + c * 2);
                        const uint4 m0v = mp[0], m1v = mp[1];
                        const uint32_t mw[8] = {m0v.x, m0v.y, m0v.z, m0v.w, m1v.x, m1v.y, m1v.z, m1v.w};
                        #pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 mm = __half22float2(*reinterpret_cast<const __half2*>(&mw[j]));
                            f[2 * j] = mm.x > 0.f ? f[2 * j] : 0.f;
                            f[2 * j + 1] = mm.y > 0.f ? f[2 * j + 1] : 0.f;
                        }
                    }
                    if (a.Yh) {
                        uint4 o0v, o1v;
                        o0v.x = pack_h2(f[0], f[1]); o0v.y = pack_h2(f[2], f[3]); o0v.z = pack_h2(f[4], f[5]); o0v.w = pack_h2(f[6], f[7]);
                        o1v.x = pack_h2(f[8], f[9]); o1v.y = pack_h2(f[10], f[11]); o1v.z = pack_h2(f[12], f[13]); o1v.w = pack_h2(f[14], f[15]);
                        // staged in the (idle) operand ring when no fp32 window shares it, written out coalesced below
                        uint4* yp = stage16 ? reinterpret_cast<uint4*>(smem + (size_t)(wq * 32 + lane) * row_bytes + c * 2)
                                            : reinterpret_cast<uint4*>(a.Yh + (size_t)row * a.ldyh + n0 + c);
                        yp[0] = o0v; yp[1] = o1v;
                    }
                    if (windows) {          // stage the fp32 row; the windows are written coalesced below
                        float* srow = stage + (size_t)(wq * 32 + lane) * sstride + c;
                        #pragma unroll
                        for (int j = 0; j < 16; ++j) srow[j] = f[j] * oscale;
                    }
                }
            }
            if (stage16) {
                // the fp16 tile -> global, a warp per row: 16 bytes per lane, whole 512-byte rows per store instruction
                __syncthreads();
                const int warp = tid >> 5;
                const int chunks = bn >> 3;
                for (int r = warp; r < 128; r += kThreads / 32) {
                    const long long grow = m0 + r;
                    if (grow >= n) break;
                    for (int ch = lane; ch < chunks; ch += 32)
                        *reinterpret_cast<uint4*>(a.Yh + (size_t)grow * a.ldyh + n0 + ch * 8) =
                            *reinterpret_cast<const uint4*>(smem + (size_t)r * row_bytes + ch * 16);
                }
            }
            if (windows) {
                // All MMAs of the item are complete (bars[kStages]) and the next item's loads start after the barrier at
                // the bottom, so the operand ring is free: it holds the [128 x bn] fp32 tile (row stride bn + 1 words).
                // Warp w writes rows w, w + 8, ...: consecutive lanes on consecutive columns of a window.
                __syncthreads();
                const int warp = tid >> 5;
                // per window: the tile columns [lo, hi) it covers, then one tight loop (lanes on consecutive columns)
                auto range = [&](int src0, int ncols, int& lo, int& hi) {
                    lo = max(src0, n0) - n0;
                    hi = min(src0 + ncols, n0 + bn) - n0;
                };
                int lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0, loh = 0, hih = 0;
                if (a.o0.ptr) range(a.o0.src0, a.o0.ncols, lo0, hi0);
                if (a.o1.ptr) range(a.o1.src0, a.o1.ncols, lo1, hi1);
                if (a.h0.ptr) range(a.h0.src0, a.h0.ncols, loh, hih);
                for (int r = warp; r < 128; r += kThreads / 32) {
                    const long long grow = m0 + r;
                    if (grow >= n) break;
                    const float* srow = stage + (size_t)r * sstride;
                    if (hi0 > lo0) {
                        float* d = a.o0.ptr + (size_t)grow * a.o0.ld + a.o0.col0 + (n0 - a.o0.src0);
                        for (int c = lo0 + lane; c < hi0; c += 32) d[c] = al_apply_act(srow[c], a.o0.act);
                    }
                    if (hi1 > lo1) {
                        float* d = a.o1.ptr + (size_t)grow * a.o1.ld + a.o1.col0 + (n0 - a.o1.src0);
                        for (int c = lo1 + lane; c < hi1; c += 32) d[c] = al_apply_act(srow[c], a.o1.act);
                    }
                    if (hih > loh) {
                        __half* d = a.h0.ptr + (size_t)grow * a.h0.ld + a.h0.col0 + (n0 - a.h0.src0);
                        if (a.h0.act == 1)
                            for (int c = loh + lane; c < hih; c += 32) d[c] = __float2half_rn(fmaxf(srow[c], 0.f));
                        else
                            for (int c = loh + lane; c < hih; c += 32) d[c] = __float2half_rn(srow[c]);
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();                                       // TMEM drained before the next item's first MMA
    }
    // drain outstanding stage commits so no mbarrier arrival is pending at exit
    for (int s = 0; s < kStages; ++s)
        if ((stage_used >> s) & 1u) mbar_wait(smem_u32(&bars[s]), (stage_par >> s) & 1u);
    cp_async_wait_all();
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

template <int MODE, bool MASK, bool WIN>
int launch_gemm_t(const GemmArgs& a, int grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        AL_CHECK(cudaFuncSetAttribute(k_gemm_tc<MODE, MASK, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        configured = true;
    }
    k_gemm_tc<MODE, MASK, WIN><<<grid, kThreads, kSmemBytes, st>>>(a);
    AL_LAUNCH_CHECK();
    return 0;
}

int launch_gemm(const GemmArgs& a, cudaStream_t st) {
    long long items;
    const int tn = (a.N + 255) / 256;
    if (a.mode == 2) {
        const int bm = (a.P % 128 == 0) ? 128 : 64;
        items = (long long)(a.P / bm) * tn * (((long long)a.M + a.rows_per_item - 1) / a.rows_per_item);
    } else {
        items = (((long long)a.M + 127) / 128) * tn;
    }
    if (items <= 0) return 0;
    {   // the large aligned layers run on the TMA pipeline (gemm_tma.cu); -1 = not its shape
        const int r = al_gemm_tma_launch(a, st);
        if (r != -1) return r;
    }
    const int grid = (int)(items < al_num_sms() ? items : al_num_sms());
    const bool win = a.o0.ptr || a.o1.ptr || a.h0.ptr;
    if (a.mode == 2) return launch_gemm_t<2, false, false>(a, grid, st);
    if (a.mode == 0) return win ? launch_gemm_t<0, false, true>(a, grid, st) : launch_gemm_t<0, false, false>(a, grid, st);
    if (a.mask) return win ? launch_gemm_t<1, true, true>(a, grid, st) : launch_gemm_t<1, true, false>(a, grid, st);
    return win ? launch_gemm_t<1, false, true>(a, grid, st) : launch_gemm_t<1, false, false>(a, grid, st);
}

// fp32 -> fp16 copy of the flat parameter vector.
__global__ void k_cast_params(const float* __restrict__ src, __half* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2half_rn(src[i]);
}

// dst[c * rows + r] = src[r * cols + c]  (fp16 weight matrix [rows, cols] -> its transpose)
__global__ void k_transpose_half(const __half* __restrict__ src, __half* __restrict__ dst, int rows, int cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) {
        const int c = i / rows, r = i - c * rows;
        dst[i] = src[(size_t)r * cols + c];
    }
}

// Output gradient window (fp32) -> scaled fp16 [live rows, out_pad], zero outside the window.
__global__ void k_cast_dout(const float* __restrict__ dout, int ld_dout, int dcol0, int dncols, int out_pad, int cap,
                            const int* __restrict__ n_dev, const float* __restrict__ amax_dev, __half* __restrict__ dst) {
    const long long n = n_dev ? min((long long)cap, (long long)*n_dev) : (long long)cap;
    const float scale = al_grad_scale(amax_dev);
    // a thread converts 8 consecutive columns of one row (one 16-byte store); out_pad is a multiple of 16.  Rows past
    // the live count are never read (the loaders zero-fill them).
    const int groups = out_pad >> 3;
    const long long total = n * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i - r * groups) * 8;
        const float* src = dout + (size_t)r * ld_dout + dcol0 + c;
        uint32_t h[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float v0 = (c + 2 * j < dncols) ? src[2 * j] * scale : 0.f;
            const float v1 = (c + 2 * j + 1 < dncols) ? src[2 * j + 1] * scale : 0.f;
            h[j] = tc::pack_h2(fminf(fmaxf(v0, -65504.f), 65504.f), fminf(fmaxf(v1, -65504.f), 65504.f));
        }
        *reinterpret_cast<uint4*>(dst + (size_t)r * out_pad + c) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}

struct WideWs {
    __half* Wh;      // fp16 parameters
    __half* A1;      // [cap, H] relu(h1)
    __half* A2;      // [cap, H] relu(h2)           (n_hidden == 2)
    __half* dY;      // [cap, out_pad] scaled output gradient
    __half* dAl;     // [cap, H] d h_last
    __half* dA1;     // [cap, H] d h1               (n_hidden == 2)
    __half* WhT;     // every weight matrix transposed (dgrad on the TMA path: both operands K-major)
    size_t bytes;
};
WideWs wide_carve(int in_pad, int H, int out_pad, int nh, int cap, int training, void* base) {
    WideWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? (void*)((char*)base + off) : nullptr;
        off += (bytes + 255) / 256 * 256;
        return p;
    };
    const size_t np = (size_t)H * in_pad + (nh == 2 ? (size_t)H * H : 0) + (size_t)out_pad * H;
    w.Wh = (__half*)take(np * 2);
    w.A1 = (__half*)take((size_t)cap * H * 2);
    w.A2 = nh == 2 ? (__half*)take((size_t)cap * H * 2) : nullptr;
    if (training) {
        w.dY = (__half*)take((size_t)cap * out_pad * 2);
        w.dAl = (__half*)take((size_t)cap * H * 2);
        w.dA1 = nh == 2 ? (__half*)take((size_t)cap * H * 2) : nullptr;
        w.WhT = (__half*)take(np * 2);
    } else {
        w.dY = w.dAl = w.dA1 = w.WhT = nullptr;
    }
    w.bytes = off;
    return w;
}

bool wide_shape_ok(int in_pad, int H, int out_pad, int nh) {
    return in_pad >= 16 && in_pad % 16 == 0 && in_pad <= 2048 && H % 64 == 0 && H >= 64 && H <= 1024 && out_pad % 16 == 0 &&
           out_pad >= 16 && out_pad <= 1024 && (nh == 1 || nh == 2);
}

}  // namespace

// ---------------------------------------------------------------------------------------------- C ABI
AL_API int al_mlp_wide_num_params(int in_pad, int hidden, int out_pad, int n_hidden) {
    if (!wide_shape_ok(in_pad, hidden, out_pad, n_hidden)) return -1;
    return hidden * in_pad + (n_hidden == 2 ? hidden * hidden : 0) + out_pad * hidden;
}
AL_API size_t al_mlp_wide_workspace(int in_pad, int hidden, int out_pad, int n_hidden, int cap, int training) {
    if (!wide_shape_ok(in_pad, hidden, out_pad, n_hidden)) return 0;
    return wide_carve(in_pad, hidden, out_pad, n_hidden, cap, training, nullptr).bytes;
}

// tcnn.Network forward, layer by layer (same contract and parameter layout as al_mlp_forward); the workspace keeps
// the fp16 parameters and hidden activations for al_mlp_wide_backward.
AL_API int al_mlp_wide_forward(int in_pad, int hidden, int out_pad, int n_hidden, const float* params,
                               const void* x_half, int ldx, int cap, const int* n_dev,
                               float* o0, int o0_ld, int o0_col0, int o0_src0, int o0_ncols, int o0_act,
                               float* o1, int o1_ld, int o1_col0, int o1_src0, int o1_ncols, int o1_act,
                               void* h0_half, int h0_ld, int h0_col0, int h0_src0, int h0_ncols, int h0_act,
                               void* workspace, void* stream) {
    if (cap <= 0) return 0;
    AL_REQUIRE(params && x_half && workspace, "null pointer");
    AL_REQUIRE(wide_shape_ok(in_pad, hidden, out_pad, n_hidden), "unsupported wide MLP shape (hidden multiple of 64, widths multiples of 16)");
    AL_REQUIRE(ldx >= in_pad && ldx % 8 == 0, "ldx must be >= in_pad and a multiple of 8");
    cudaStream_t st = (cudaStream_t)stream;
    const WideWs w = wide_carve(in_pad, hidden, out_pad, n_hidden, cap, 0, workspace);
    const int np = al_mlp_wide_num_params(in_pad, hidden, out_pad, n_hidden);
    k_cast_params<<<al_div_up(np, 256), 256, 0, st>>>(params, w.Wh, np);
    AL_LAUNCH_CHECK();
    const __half* W1 = w.Wh;
    const __half* W2 = W1 + (size_t)hidden * in_pad;
    const __half* WO = W2 + (n_hidden == 2 ? (size_t)hidden * hidden : 0);
    GemmArgs g = {};
    g.mode = 0; g.M = cap; g.n_dev = n_dev; g.relu = 1;
    g.A = (const __half*)x_half; g.lda = ldx; g.B = W1; g.ldb = in_pad; g.N = hidden; g.K = in_pad; g.Yh = w.A1; g.ldyh = hidden;
    { const int r = launch_gemm(g, st); if (r) return r; }
    const __half* last = w.A1;
    if (n_hidden == 2) {
        g.A = w.A1; g.lda = hidden; g.B = W2; g.ldb = hidden; g.N = hidden; g.K = hidden; g.Yh = w.A2; g.ldyh = hidden;
        { const int r = launch_gemm(g, st); if (r) return r; }
        last = w.A2;
    }
    g.relu = 0; g.A = last; g.lda = hidden; g.B = WO; g.ldb = hidden; g.N = out_pad; g.K = hidden; g.Yh = nullptr;
    g.o0 = {o0, o0_ld, o0_col0, o0_src0, o0_ncols, o0_act};
    g.o1 = {o1, o1_ld, o1_col0, o1_src0, o1_ncols, o1_act};
    g.h0 = {(__half*)h0_half, h0_ld, h0_col0, h0_src0, h0_ncols, h0_act};
    return launch_gemm(g, st);
}

// The scaled fp16 output-gradient buffer [cap, out_pad] of a wide MLP's training workspace (filled by the caller
// before al_wide_backward_dy; csrc/field.cu assembles the semantic heads' gradients straight into it).
void* al_wide_dy(int in_pad, int hidden, int out_pad, int n_hidden, int cap, void* workspace) {
    return wide_carve(in_pad, hidden, out_pad, n_hidden, cap, 1, workspace).dY;
}

// Backward of the wide path with dY (scaled by al_grad_scale(amax_dev), fp16) already in the workspace.
int al_wide_backward_dy(int in_pad, int hidden, int out_pad, int n_hidden, const void* x_half, int ldx, int cap,
                        const int* n_dev, const float* amax_dev, float* dparams, float* dx, int ld_dx, int dx_c0,
                        int dx_n, void* workspace, cudaStream_t st) {
    const WideWs w = wide_carve(in_pad, hidden, out_pad, n_hidden, cap, 1, workspace);
    const __half* W1 = w.Wh;
    const __half* W2 = W1 + (size_t)hidden * in_pad;
    const __half* WO = W2 + (n_hidden == 2 ? (size_t)hidden * hidden : 0);
    float* g1 = dparams;
    float* g2 = dparams ? g1 + (size_t)hidden * in_pad : nullptr;
    float* go = dparams ? g2 + (n_hidden == 2 ? (size_t)hidden * hidden : 0) : nullptr;
    const __half* a_last = n_hidden == 2 ? w.A2 : w.A1;
    // transposed copies of the output and hidden->hidden matrices: WOt [hidden, out_pad], W2t [hidden (in), hidden (out)]
    __half* WOt = w.WhT;
    __half* W2t = WOt + (size_t)out_pad * hidden;
    k_transpose_half<<<al_div_up(out_pad * hidden, 256), 256, 0, st>>>(WO, WOt, out_pad, hidden);
    AL_LAUNCH_CHECK();
    if (n_hidden == 2) {
        k_transpose_half<<<al_div_up(hidden * hidden, 256), 256, 0, st>>>(W2, W2t, hidden, hidden);
        AL_LAUNCH_CHECK();
    }
    __half* W1t = W2t + (n_hidden == 2 ? (size_t)hidden * hidden : 0);          // [in_pad, hidden]
    if (dx) {
        k_transpose_half<<<al_div_up(hidden * in_pad, 256), 256, 0, st>>>(W1, W1t, hidden, in_pad);
        AL_LAUNCH_CHECK();
    }
    GemmArgs g = {};
    g.M = cap; g.n_dev = n_dev; g.amax_dev = amax_dev;
    // d h_last = (dY Wo) * relu'(a_last)
    g.mode = 1; g.A = w.dY; g.lda = out_pad; g.B = WO; g.ldb = hidden; g.N = hidden; g.K = out_pad;
    g.Bt = WOt; g.ldbt = out_pad;
    g.mask = a_last; g.ldmask = hidden; g.Yh = w.dAl; g.ldyh = hidden;
    { const int r = launch_gemm(g, st); if (r) return r; }
    auto wgrad = [&](const __half* dYm, int n_out, const __half* Xm, int ldxm, int n_in, float* G) -> int {
        // G[out, in] += dY^T X: the side that is a multiple of 64 becomes the MMA M dimension
        if (!G) return 0;
        GemmArgs h = {};
        h.mode = 2; h.M = cap; h.n_dev = n_dev; h.amax_dev = amax_dev; h.rows_per_item = 2048; h.G = G;
        if (n_out % 64 == 0) { h.A = dYm; h.lda = n_out; h.P = n_out; h.B = Xm; h.ldb = ldxm; h.N = n_in; h.sp = n_in; h.sq = 1; }
        else { h.A = Xm; h.lda = ldxm; h.P = n_in; h.B = dYm; h.ldb = n_out; h.N = n_out; h.sp = 1; h.sq = n_in; }
        return launch_gemm(h, st);
    };
    { const int r = wgrad(w.dY, out_pad, a_last, hidden, hidden, go); if (r) return r; }
    const __half* d1 = w.dAl;
    if (n_hidden == 2) {
        GemmArgs h = g;
        h.A = w.dAl; h.lda = hidden; h.B = W2; h.ldb = hidden; h.N = hidden; h.K = hidden; h.mask = w.A1; h.ldmask = hidden;
        h.Bt = W2t; h.ldbt = hidden;
        h.Yh = w.dA1; h.ldyh = hidden;
        { const int r = launch_gemm(h, st); if (r) return r; }
        { const int r = wgrad(w.dAl, hidden, w.A1, hidden, hidden, g2); if (r) return r; }
        d1 = w.dA1;
    }
    if (dx) {
        GemmArgs h = {};
        h.mode = 1; h.M = cap; h.n_dev = n_dev; h.amax_dev = amax_dev; h.unscale = 1;
        h.A = d1; h.lda = hidden; h.B = W1; h.ldb = in_pad; h.N = in_pad; h.K = hidden; h.Bt = W1t; h.ldbt = hidden;
        h.o0 = {dx, ld_dx, 0, dx_c0, dx_n, 0};
        { const int r = launch_gemm(h, st); if (r) return r; }
    }
    return wgrad(d1, hidden, (const __half*)x_half, ldx, in_pad, g1);
}

// tcnn.Network backward for the wide path: dparams += dL/dparams, dx (optional, row-major window) = dL/dx.
// Must follow al_mlp_wide_forward on the same workspace (allocated with training = 1).
AL_API int al_mlp_wide_backward(int in_pad, int hidden, int out_pad, int n_hidden, const float* params,
                                const void* x_half, int ldx, int cap, const int* n_dev, const float* dout,
                                int ld_dout, int dcol0, int dncols, const float* amax_dev, float* dparams,
                                float* dx, int ld_dx, int dx_c0, int dx_n, void* workspace, void* stream) {
    if (cap <= 0) return 0;
    AL_REQUIRE(params && x_half && dout && workspace && amax_dev, "null pointer");
    AL_REQUIRE(wide_shape_ok(in_pad, hidden, out_pad, n_hidden), "unsupported wide MLP shape");
    AL_REQUIRE(dncols <= out_pad, "dncols exceeds the padded output width");
    cudaStream_t st = (cudaStream_t)stream;
    const WideWs w = wide_carve(in_pad, hidden, out_pad, n_hidden, cap, 1, workspace);
    const unsigned long long cast_blocks = al_div_up((unsigned long long)cap * (out_pad / 8), 256);
    const unsigned long long cast_full = (unsigned long long)al_num_sms() * 8;
    k_cast_dout<<<(unsigned)(cast_blocks < cast_full ? cast_blocks : cast_full), 256, 0, st>>>(dout, ld_dout, dcol0, dncols, out_pad, cap,
                                                                                  n_dev, amax_dev, w.dY);
    AL_LAUNCH_CHECK();
    return al_wide_backward_dy(in_pad, hidden, out_pad, n_hidden, x_half, ldx, cap, n_dev, amax_dev, dparams, dx, ld_dx,
                               dx_c0, dx_n, workspace, st);
}
