// Fully fused bias-free MLP (ReLU hidden, linear output) on tensor cores: forward, and a
// backward that recomputes the hidden activations, back-propagates through all layers and
// accumulates the weight gradients inside the same kernel.
//
// Replaces tiny-cuda-nn's `tcnn.Network` (FullyFusedMLP / CutlassMLP) as used by
//   autolabel/models.py:84-136  (sigma_net, color_net, semantic_features, semantic_out)
// whose in-tree design reference is torch_ngp/ffmlp/src/ffmlp.cu:331-518 (wmma forward /
// backward, CUTLASS split-K weight gradients on side streams, cutlass_matmul.h).
//
// Contract (oracle/field_oracle.py::mlp): x [n, IN] (columns beyond the logical input width are
// 1.0 = bias column), W1 [H, IN], (W2 [H, H] if NH == 2), Wo [OUT, H], all row-major [out, in] in
// one flat fp32 parameter vector;  y = Wo relu(W2 relu(W1 x)).  Operands are rounded to fp16,
// products accumulate in fp32 (mma.sync.m16n8k16).
//
// Layout / schedule (differs from ffmlp on purpose):
//  * persistent CTAs (one per SM, 8 warps); the whole weight set lives in shared memory as fp16
//    for the lifetime of the CTA (converted from the fp32 master copy once per launch);
//  * a warp owns 16 sample rows and carries them through ALL layers in registers: the m16n8
//    accumulator fragment of layer l is re-packed in place as the A fragment of layer l+1, so
//    hidden activations never touch shared or global memory in the forward pass;
//  * backward: per 128-row tile, phase A (per warp) recomputes the activations, runs the dgrad
//    chain and parks fp16 activations / output-gradients in shared memory; phase B (whole CTA)
//    computes dW_l = dH_l^T A_{l-1} with the 128 samples as the MMA K dimension.  Each warp owns a
//    fixed slice of every dW and keeps it in registers across ALL tiles of the persistent CTA;
//    one fp32 red.global.add per weight per CTA at the very end.  No activation or dH tensor is
//    written to HBM (the reference stores both and runs separate split-K GEMMs).
//  * gradients enter scaled by a power of two derived on the device from their running max
//    (`amax_dev`), so fp16 dH does not underflow; dW / dx leave unscaled in fp32.
#include "common.cuh"
#include "mlp_args.cuh"
#include <mma.h>
#include <stdlib.h>

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTileRows = kWarps * 16;
constexpr int kPad = 8;  // halfs of row padding -> conflict-free ldmatrix

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// ---------------------------------------------------------------- building blocks
// fp32 global [rows][cols] -> fp16 shared [rows][cols + kPad]
__device__ __forceinline__ void stage_weights(const float* __restrict__ g, __half* s, int rows, int cols) {
    const int stride = cols + kPad;
    const int vec_per_row = cols / 4;
    for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
        const int r = i / vec_per_row, c = (i - r * vec_per_row) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * cols + c));
        __half2* d = reinterpret_cast<__half2*>(s + r * stride + c);
        d[0] = __floats2half2_rn(v.x, v.y);
        d[1] = __floats2half2_rn(v.z, v.w);
    }
}

// A fragments of a [16 x 16*KS] fp16 tile in shared memory (row stride `stride` halfs).
template <int KS>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[KS][4], const __half* tile, int stride, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    const __half* p = tile + (r + (mi & 1) * 8) * stride + (mi >> 1) * 8;
    #pragma unroll
    for (int kk = 0; kk < KS; ++kk) ldsm_x4(a[kk], p + kk * 16);
}

// acc[NT][4] += A[16 x 16*KS] * W^T, W = [8*NT][16*KS] fp16 in shared ([out][in], stride halfs).
template <int KS, int NT>
__device__ __forceinline__ void layer_fwd(float (&acc)[NT][4], const uint32_t (&a)[KS][4], const __half* W,
                                          int stride, int lane) {
    static_assert(NT % 2 == 0, "NT must be even");
    const int mi = lane >> 3, r = lane & 7;
    const __half* p = W + ((mi >> 1) * 8 + r) * stride + (mi & 1) * 8;
    #pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
        #pragma unroll
        for (int j = 0; j < NT; j += 2) {
            uint32_t b[4];
            ldsm_x4(b, p + (j * 8) * stride + kk * 16);
            mma16816(acc[j], a[kk], b[0], b[1]);
            mma16816(acc[j + 1], a[kk], b[2], b[3]);
        }
    }
}

// acc[NT][4] += D[16 x 16*KS] * W, W = [16*KS][8*NT] fp16 in shared ([out][in]): dgrad.
template <int KS, int NT>
__device__ __forceinline__ void layer_dgrad(float (&acc)[NT][4], const uint32_t (&d)[KS][4], const __half* W,
                                            int stride, int lane) {
    static_assert(NT % 2 == 0, "NT must be even");
    const int mi = lane >> 3, r = lane & 7;
    const __half* p = W + ((mi & 1) * 8 + r) * stride + (mi >> 1) * 8;
    #pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
        #pragma unroll
        for (int j = 0; j < NT; j += 2) {
            uint32_t b[4];
            ldsm_x4_t(b, p + (kk * 16) * stride + j * 8);
            mma16816(acc[j], d[kk], b[0], b[1]);
            mma16816(acc[j + 1], d[kk], b[2], b[3]);
        }
    }
}

template <int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[NT][4]) {
    #pragma unroll
    for (int j = 0; j < NT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
}

// ReLU + repack accumulator fragments [16 x 8*NT] as A fragments [16 x 16*(NT/2)].
template <int NT>
__device__ __forceinline__ void relu_to_a(const float (&acc)[NT][4], uint32_t (&a)[NT / 2][4]) {
    #pragma unroll
    for (int k = 0; k < NT / 2; ++k) {
        a[k][0] = pack_h2(fmaxf(acc[2 * k][0], 0.f), fmaxf(acc[2 * k][1], 0.f));
        a[k][1] = pack_h2(fmaxf(acc[2 * k][2], 0.f), fmaxf(acc[2 * k][3], 0.f));
        a[k][2] = pack_h2(fmaxf(acc[2 * k + 1][0], 0.f), fmaxf(acc[2 * k + 1][1], 0.f));
        a[k][3] = pack_h2(fmaxf(acc[2 * k + 1][2], 0.f), fmaxf(acc[2 * k + 1][3], 0.f));
    }
}

// Store A-fragment-layout registers [16 x 16*KS] to a shared tile (row-major, stride halfs).
template <int KS>
__device__ __forceinline__ void store_a_frags(const uint32_t (&a)[KS][4], __half* tile, int stride, int lane) {
    const int g = lane >> 2, tq = lane & 3;
    #pragma unroll
    for (int k = 0; k < KS; ++k) {
        *reinterpret_cast<uint32_t*>(tile + g * stride + k * 16 + tq * 2) = a[k][0];
        *reinterpret_cast<uint32_t*>(tile + (g + 8) * stride + k * 16 + tq * 2) = a[k][1];
        *reinterpret_cast<uint32_t*>(tile + g * stride + k * 16 + 8 + tq * 2) = a[k][2];
        *reinterpret_cast<uint32_t*>(tile + (g + 8) * stride + k * 16 + 8 + tq * 2) = a[k][3];
    }
}

// dgrad accumulators * relu'(activation) -> A fragments.  The activation tile (fp16, post-ReLU) is
// read back from shared memory at exactly the positions this lane wrote in the forward recompute.
template <int NT>
__device__ __forceinline__ void mask_to_a(const float (&acc)[NT][4], const __half* act_tile, int stride,
                                          int lane, uint32_t (&a)[NT / 2][4]) {
    const int g = lane >> 2, tq = lane & 3;
    #pragma unroll
    for (int k = 0; k < NT / 2; ++k) {
        const __half2 m0 = *reinterpret_cast<const __half2*>(act_tile + g * stride + k * 16 + tq * 2);
        const __half2 m1 = *reinterpret_cast<const __half2*>(act_tile + (g + 8) * stride + k * 16 + tq * 2);
        const __half2 m2 = *reinterpret_cast<const __half2*>(act_tile + g * stride + k * 16 + 8 + tq * 2);
        const __half2 m3 = *reinterpret_cast<const __half2*>(act_tile + (g + 8) * stride + k * 16 + 8 + tq * 2);
        const float2 f0 = __half22float2(m0), f1 = __half22float2(m1), f2 = __half22float2(m2),
                     f3 = __half22float2(m3);
        a[k][0] = pack_h2(f0.x > 0.f ? acc[2 * k][0] : 0.f, f0.y > 0.f ? acc[2 * k][1] : 0.f);
        a[k][1] = pack_h2(f1.x > 0.f ? acc[2 * k][2] : 0.f, f1.y > 0.f ? acc[2 * k][3] : 0.f);
        a[k][2] = pack_h2(f2.x > 0.f ? acc[2 * k + 1][0] : 0.f, f2.y > 0.f ? acc[2 * k + 1][1] : 0.f);
        a[k][3] = pack_h2(f3.x > 0.f ? acc[2 * k + 1][2] : 0.f, f3.y > 0.f ? acc[2 * k + 1][3] : 0.f);
    }
}

// Cooperative copy of this warp's 16 input rows (fp16, global) into its shared tile; rows >= n are
// zero-filled.  cols is a multiple of 8 (16-byte chunks).
__device__ __forceinline__ void stage_rows(const __half* __restrict__ x, size_t ldx, int cols, long long row0,
                                           long long n, __half* tile, int stride, int lane) {
    const int chunks = cols / 8;
    for (int i = lane; i < 16 * chunks; i += 32) {
        const int r = i / chunks, c = (i - r * chunks) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row0 + r < n) v = __ldg(reinterpret_cast<const uint4*>(x + (size_t)(row0 + r) * ldx + c));
        *reinterpret_cast<uint4*>(tile + r * stride + c) = v;
    }
}

// ---------------------------------------------------------------- epilogue descriptors (mlp_args.cuh)
__device__ __forceinline__ float apply_act(float v, int act) { return al_apply_act(v, act); }

template <int NT>
__device__ __forceinline__ void write_out_f32(const OutF32& o, const float (&acc)[NT][4], long long row0,
                                              long long n, int lane) {
    if (!o.ptr) return;
    const int g = lane >> 2, tq = lane & 3;
    #pragma unroll
    for (int j = 0; j < NT; ++j) {
        #pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int col = j * 8 + tq * 2 + (e & 1);
            const long long row = row0 + g + ((e >> 1) ? 8 : 0);
            const int rel = col - o.src0;
            if (rel >= 0 && rel < o.ncols && row < n)
                o.ptr[(size_t)row * o.ld + o.col0 + rel] = apply_act(acc[j][e], o.act);
        }
    }
}
template <int NT>
__device__ __forceinline__ void write_out_f16(const OutF16& o, const float (&acc)[NT][4], long long row0,
                                              long long n, int lane) {
    if (!o.ptr) return;
    const int g = lane >> 2, tq = lane & 3;
    #pragma unroll
    for (int j = 0; j < NT; ++j) {
        #pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int col = j * 8 + tq * 2 + (e & 1);
            const long long row = row0 + g + ((e >> 1) ? 8 : 0);
            const int rel = col - o.src0;
            if (rel >= 0 && rel < o.ncols && row < n) {
                float v = acc[j][e];
                if (o.act == 1) v = fmaxf(v, 0.f);
                o.ptr[(size_t)row * o.ld + o.col0 + rel] = __float2half_rn(v);
            }
        }
    }
}

template <int IN, int H, int OUT, int NH>
struct Dims {
    static constexpr int SI = IN + kPad, SH = H + kPad, SO = OUT + kPad;
    static constexpr int W1 = 0;                                   // [H][IN]
    static constexpr int W2 = W1 + H * IN;                         // [H][H]
    static constexpr int WO = W2 + (NH == 2 ? H * H : 0);          // [OUT][H]
    static constexpr int NPARAMS = WO + OUT * H;
    // shared (halfs)
    static constexpr int sW1 = 0;
    static constexpr int sW2 = sW1 + H * SI;
    static constexpr int sWO = sW2 + (NH == 2 ? H * SH : 0);
    static constexpr int sWend = sWO + OUT * SH;
};

// ================================================================ forward kernel
template <int IN, int H, int OUT, int NH>
__global__ void __launch_bounds__(kThreads, 1) k_mlp_fwd(const MlpFwdArgs args) {
    using D = Dims<IN, H, OUT, NH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sm = reinterpret_cast<__half*>(smem_raw);
    __half* sW1 = sm + D::sW1;
    __half* sW2 = sm + D::sW2;
    __half* sWO = sm + D::sWO;
    __half* sX = sm + D::sWend;  // [kWarps][16][SI]

    stage_weights(args.params + D::W1, sW1, H, IN);
    if (NH == 2) stage_weights(args.params + D::W2, sW2, H, H);
    stage_weights(args.params + D::WO, sWO, OUT, H);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n = args.n_dev ? min((long long)args.cap, (long long)*args.n_dev) : (long long)args.cap;
    __half* xt = sX + warp * 16 * D::SI;
    for (long long tile = blockIdx.x; tile * kTileRows < n; tile += gridDim.x) {
        const long long row0 = tile * kTileRows + warp * 16;
        if (row0 >= n) continue;
        stage_rows(args.x, args.ldx, IN, row0, n, xt, D::SI, lane);
        __syncwarp();
        uint32_t a0[IN / 16][4];
        load_a_frags<IN / 16>(a0, xt, D::SI, lane);
        __syncwarp();
        float acc[H / 8][4];
        zero_acc<H / 8>(acc);
        layer_fwd<IN / 16, H / 8>(acc, a0, sW1, D::SI, lane);
        uint32_t h[H / 16][4];
        relu_to_a<H / 8>(acc, h);
        if constexpr (NH == 2) {
            zero_acc<H / 8>(acc);
            layer_fwd<H / 16, H / 8>(acc, h, sW2, D::SH, lane);
            relu_to_a<H / 8>(acc, h);
        }
        float out[OUT / 8][4];
        zero_acc<OUT / 8>(out);
        layer_fwd<H / 16, OUT / 8>(out, h, sWO, D::SH, lane);
        write_out_f32<OUT / 8>(args.o0, out, row0, n, lane);
        write_out_f32<OUT / 8>(args.o1, out, row0, n, lane);
        write_out_f16<OUT / 8>(args.h0, out, row0, n, lane);
    }
}

// ================================================================ backward kernel
// dW slice owned by one warp: tiles of m16 (out) x n8 (in).
template <int OUTD, int IND>
struct WSlice {
    static constexpr int MT = OUTD / 16, NT = IND / 8, T = MT * NT;
    static_assert(T % kWarps == 0, "dW tiles must divide evenly over the warps");
    static constexpr int TPW = T / kWarps;                        // tiles per warp
    static constexpr int NPER = TPW < NT ? TPW : NT;              // n-tiles per m-tile handled
    static constexpr int MPER = TPW / NPER;                       // m-tiles handled
    static_assert(NPER * MPER == TPW && (NT % NPER) == 0, "unsupported dW split");
};

// acc[MPER][NPER][4] += dH^T A over the tile's 128 samples.
//   dh: [128][SD] fp16 (sample-major), act: [128][SA] fp16 (sample-major)
template <int OUTD, int IND>
__device__ __forceinline__ void wgrad_tile(float (&acc)[WSlice<OUTD, IND>::MPER][WSlice<OUTD, IND>::NPER][4],
                                           const __half* dh, int SD, const __half* act, int SA, int warp,
                                           int lane) {
    using S = WSlice<OUTD, IND>;
    const int t0 = warp * S::TPW;
    const int m0 = t0 / S::NT, n0 = t0 % S::NT;
    const int mi = lane >> 3, r = lane & 7;
    #pragma unroll 2
    for (int kk = 0; kk < kTileRows / 16; ++kk) {
        #pragma unroll
        for (int mm = 0; mm < S::MPER; ++mm) {
            // A(m = out, k = sample) = dh[sample][out]  (transposed load)
            uint32_t a[4];
            {
                // matrices: (k lo, m lo) (k lo, m hi) (k hi, m lo) (k hi, m hi) -> a0 a1 a2 a3
                const __half* p = dh + (kk * 16 + (mi >> 1) * 8 + r) * SD + (m0 + mm) * 16 + (mi & 1) * 8;
                ldsm_x4_t(a, p);
            }
            #pragma unroll
            for (int j = 0; j + 1 < S::NPER; j += 2) {
                // B(k = sample, n = in) = act[sample][in]; matrices (k lo, n j) (k hi, n j) (k lo, n j+1) (k hi, n j+1)
                uint32_t b[4];
                const __half* p = act + (kk * 16 + (mi & 1) * 8 + r) * SA + (n0 + j + (mi >> 1)) * 8;
                ldsm_x4_t(b, p);
                mma16816(acc[mm][j], a, b[0], b[1]);
                mma16816(acc[mm][j + 1], a, b[2], b[3]);
            }
            if (S::NPER & 1) {
                constexpr int j = S::NPER - 1;
                uint32_t b[2];
                const __half* p = act + (kk * 16 + ((lane >> 3) & 1) * 8 + r) * SA + (n0 + j) * 8;
                ldsm_x2_t(b, p);
                mma16816(acc[mm][j], a, b[0], b[1]);
            }
        }
    }
}

template <int OUTD, int IND>
__device__ __forceinline__ void wgrad_flush(const float (&acc)[WSlice<OUTD, IND>::MPER][WSlice<OUTD, IND>::NPER][4],
                                            float* __restrict__ dW, float inv_scale, int warp, int lane) {
    using S = WSlice<OUTD, IND>;
    const int t0 = warp * S::TPW;
    const int m0 = t0 / S::NT, n0 = t0 % S::NT;
    const int g = lane >> 2, tq = lane & 3;
    #pragma unroll
    for (int mm = 0; mm < S::MPER; ++mm)
        #pragma unroll
        for (int j = 0; j < S::NPER; ++j) {
            const int o = (m0 + mm) * 16 + g, i = (n0 + j) * 8 + tq * 2;
            atomicAdd(dW + (size_t)o * IND + i, acc[mm][j][0] * inv_scale);
            atomicAdd(dW + (size_t)o * IND + i + 1, acc[mm][j][1] * inv_scale);
            atomicAdd(dW + (size_t)(o + 8) * IND + i, acc[mm][j][2] * inv_scale);
            atomicAdd(dW + (size_t)(o + 8) * IND + i + 1, acc[mm][j][3] * inv_scale);
        }
}

template <int IN, int H, int OUT, int NH>
struct BwdSmem {
    using D = Dims<IN, H, OUT, NH>;
    static constexpr int A0 = D::sWend;                       // [128][SI]  layer-1 input
    static constexpr int A1 = A0 + kTileRows * D::SI;         // [128][SH]  relu(h1)
    static constexpr int A2 = A1 + kTileRows * D::SH;         // [128][SH]  relu(h2)   (NH == 2)
    static constexpr int DO = A2 + (NH == 2 ? kTileRows * D::SH : 0);  // [128][SO]  d out
    static constexpr int D2 = DO + kTileRows * D::SO;         // [128][SH]  d h_last
    static constexpr int D1 = D2 + kTileRows * D::SH;         // [128][SH]  d h1        (NH == 2)
    static constexpr int END = D1 + (NH == 2 ? kTileRows * D::SH : 0);
    static constexpr size_t BYTES = (size_t)END * 2;
};

template <int IN, int H, int OUT, int NH>
__global__ void __launch_bounds__(kThreads, 1) k_mlp_bwd(const MlpBwdArgs args) {
    using D = Dims<IN, H, OUT, NH>;
    using SM = BwdSmem<IN, H, OUT, NH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sm = reinterpret_cast<__half*>(smem_raw);
    __half* sW1 = sm + D::sW1;
    __half* sW2 = sm + D::sW2;
    __half* sWO = sm + D::sWO;
    __half* sA0 = sm + SM::A0;
    __half* sA1 = sm + SM::A1;
    __half* sA2 = sm + SM::A2;
    __half* sDO = sm + SM::DO;
    __half* sD2 = sm + SM::D2;
    __half* sD1 = sm + SM::D1;

    stage_weights(args.params + D::W1, sW1, H, IN);
    if (NH == 2) stage_weights(args.params + D::W2, sW2, H, H);
    stage_weights(args.params + D::WO, sWO, OUT, H);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const long long n = args.n_dev ? min((long long)args.cap, (long long)*args.n_dev) : (long long)args.cap;

    const float scale = al_grad_scale(args.amax_dev);
    const float inv_scale = 1.0f / scale;

    // persistent weight-gradient accumulators
    float gW1[WSlice<H, IN>::MPER][WSlice<H, IN>::NPER][4];
    float gW2[NH == 2 ? WSlice<H, H>::MPER : 1][NH == 2 ? WSlice<H, H>::NPER : 1][4];
    float gWO[WSlice<OUT, H>::MPER][WSlice<OUT, H>::NPER][4];
    #pragma unroll
    for (int a = 0; a < WSlice<H, IN>::MPER; ++a)
        #pragma unroll
        for (int b = 0; b < WSlice<H, IN>::NPER; ++b) gW1[a][b][0] = gW1[a][b][1] = gW1[a][b][2] = gW1[a][b][3] = 0.f;
    #pragma unroll
    for (int a = 0; a < (NH == 2 ? WSlice<H, H>::MPER : 1); ++a)
        #pragma unroll
        for (int b = 0; b < (NH == 2 ? WSlice<H, H>::NPER : 1); ++b) gW2[a][b][0] = gW2[a][b][1] = gW2[a][b][2] = gW2[a][b][3] = 0.f;
    #pragma unroll
    for (int a = 0; a < WSlice<OUT, H>::MPER; ++a)
        #pragma unroll
        for (int b = 0; b < WSlice<OUT, H>::NPER; ++b) gWO[a][b][0] = gWO[a][b][1] = gWO[a][b][2] = gWO[a][b][3] = 0.f;

    __syncthreads();

    for (long long tile = blockIdx.x; tile * kTileRows < n; tile += gridDim.x) {
        const long long row0 = tile * kTileRows + warp * 16;
        __half* a0t = sA0 + warp * 16 * D::SI;
        __half* a1t = sA1 + warp * 16 * D::SH;
        __half* a2t = sA2 + warp * 16 * D::SH;
        __half* dot = sDO + warp * 16 * D::SO;
        __half* d2t = sD2 + warp * 16 * D::SH;
        __half* d1t = sD1 + warp * 16 * D::SH;

        // ---------------- phase A: this warp's 16 rows through fwd recompute + dgrad chain
        stage_rows(args.x, args.ldx, IN, row0, n, a0t, D::SI, lane);
        // output gradient tile -> fp16 (scaled), zero outside [dcol0, dcol0+dncols) and for rows >= n.
        // All loads of the tile are issued before the first use (one L2 round trip, not 16*OUT/32).
        {
            constexpr int PER = 16 * OUT / 32;
            float v[PER];
            #pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = lane + 32 * j;
                const int r = i / OUT, c = i - r * OUT;
                v[j] = 0.f;
                if (row0 + r < n && c < args.dncols)
                    v[j] = __ldg(args.dout + (size_t)(row0 + r) * args.ld_dout + args.dcol0 + c);
            }
            #pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = lane + 32 * j;
                const int r = i / OUT, c = i - r * OUT;
                dot[r * D::SO + c] = __float2half_rn(fminf(fmaxf(v[j] * scale, -65504.f), 65504.f));
            }
        }
        __syncwarp();
        {
            uint32_t h[H / 16][4];
            {
                uint32_t a0[IN / 16][4];
                load_a_frags<IN / 16>(a0, a0t, D::SI, lane);
                float acc[H / 8][4];
                zero_acc<H / 8>(acc);
                layer_fwd<IN / 16, H / 8>(acc, a0, sW1, D::SI, lane);
                relu_to_a<H / 8>(acc, h);
                store_a_frags<H / 16>(h, a1t, D::SH, lane);
                if constexpr (NH == 2) {
                    zero_acc<H / 8>(acc);
                    layer_fwd<H / 16, H / 8>(acc, h, sW2, D::SH, lane);
                    relu_to_a<H / 8>(acc, h);
                    store_a_frags<H / 16>(h, a2t, D::SH, lane);
                }
            }
            __syncwarp();
            // d h_last = (dout Wo) * relu'
            uint32_t dlast[H / 16][4];
            {
                uint32_t dfr[OUT / 16][4];
                load_a_frags<OUT / 16>(dfr, dot, D::SO, lane);
                float acc[H / 8][4];
                zero_acc<H / 8>(acc);
                layer_dgrad<OUT / 16, H / 8>(acc, dfr, sWO, D::SH, lane);
                mask_to_a<H / 8>(acc, NH == 2 ? a2t : a1t, D::SH, lane, dlast);
                store_a_frags<H / 16>(dlast, d2t, D::SH, lane);
            }
            if constexpr (NH == 2) {
                float acc[H / 8][4];
                zero_acc<H / 8>(acc);
                layer_dgrad<H / 16, H / 8>(acc, dlast, sW2, D::SH, lane);
                mask_to_a<H / 8>(acc, a1t, D::SH, lane, dlast);
                store_a_frags<H / 16>(dlast, d1t, D::SH, lane);
            }
            if (args.dx) {
                float acc[IN / 8][4];
                zero_acc<IN / 8>(acc);
                layer_dgrad<H / 16, IN / 8>(acc, dlast, sW1, D::SI, lane);
                #pragma unroll
                for (int j = 0; j < IN / 8; ++j)
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = j * 8 + tq * 2 + (e & 1);
                        const long long row = row0 + g + ((e >> 1) ? 8 : 0);
                        const int rel = col - args.dx_c0;
                        if (rel >= 0 && rel < args.dx_n && row < n) {
                            const float v = acc[j][e] * inv_scale;
                            if (args.dx_mode == 0) args.dx[(size_t)row * args.ld_dx + rel] = v;
                            else args.dx[((size_t)(rel >> 1) * args.ld_dx + row) * 2 + (rel & 1)] = v;
                        }
                    }
            }
        }
        __syncthreads();
        // ---------------- phase B: weight gradients with the tile's 128 samples as K
        wgrad_tile<OUT, H>(gWO, sDO, D::SO, NH == 2 ? sA2 : sA1, D::SH, warp, lane);
        if constexpr (NH == 2) {
            wgrad_tile<H, H>(gW2, sD2, D::SH, sA1, D::SH, warp, lane);
            wgrad_tile<H, IN>(gW1, sD1, D::SH, sA0, D::SI, warp, lane);
        } else {
            wgrad_tile<H, IN>(gW1, sD2, D::SH, sA0, D::SI, warp, lane);
        }
        __syncthreads();
    }

    if (args.dparams) {
        wgrad_flush<H, IN>(gW1, args.dparams + D::W1, inv_scale, warp, lane);
        if constexpr (NH == 2) wgrad_flush<H, H>(gW2, args.dparams + D::W2, inv_scale, warp, lane);
        wgrad_flush<OUT, H>(gWO, args.dparams + D::WO, inv_scale, warp, lane);
    }
}

template <int IN, int H, int OUT, int NH>
int launch_fwd(const MlpFwdArgs& a, cudaStream_t st) {
    using D = Dims<IN, H, OUT, NH>;
    const size_t smem = ((size_t)D::sWend + (size_t)kWarps * 16 * D::SI) * 2;
    static bool configured = false;
    if (!configured) {
        AL_CHECK(cudaFuncSetAttribute(k_mlp_fwd<IN, H, OUT, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int tiles = (a.cap + kTileRows - 1) / kTileRows;
    const int grid = tiles < al_num_sms() ? tiles : al_num_sms();
    k_mlp_fwd<IN, H, OUT, NH><<<grid, kThreads, smem, st>>>(a);
    AL_LAUNCH_CHECK();
    return 0;
}
template <int IN, int H, int OUT, int NH>
int launch_bwd(const MlpBwdArgs& a, cudaStream_t st) {
    using SM = BwdSmem<IN, H, OUT, NH>;
    static_assert(SM::BYTES <= 227 * 1024, "backward tile does not fit in shared memory");
    static bool configured = false;
    if (!configured) {
        AL_CHECK(cudaFuncSetAttribute(k_mlp_bwd<IN, H, OUT, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES));
        configured = true;
    }
    const int tiles = (a.cap + kTileRows - 1) / kTileRows;
    const int grid = tiles < al_num_sms() ? tiles : al_num_sms();
    k_mlp_bwd<IN, H, OUT, NH><<<grid, kThreads, SM::BYTES, st>>>(a);
    AL_LAUNCH_CHECK();
    return 0;
}

// amax of |v| over a strided fp32 matrix region, accumulated with atomicMax on the bit pattern.
__global__ void k_amax(const float* __restrict__ v, int ld, int col0, int ncols, int cap,
                       const int* __restrict__ n_dev, float* __restrict__ amax) {
    const long long n = n_dev ? min((long long)cap, (long long)*n_dev) : (long long)cap;
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * ncols;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ncols;
        const int c = (int)(i - r * ncols);
        const float a = fabsf(v[(size_t)r * ld + col0 + c]);
        if (a < 3.0e38f) m = fmaxf(m, a);
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));
}

}  // namespace

// Back end of al_mlp_forward / al_mlp_backward: 1 = tcgen05 / TMEM (mlp_tc.cu, default), 0 = mma.sync (this file).
static int g_mlp_backend = -1;
static int mlp_backend() {
    if (g_mlp_backend < 0) {
        const char* e = getenv("AL_MLP_BACKEND");
        g_mlp_backend = (e && (e[0] == 'm' || e[0] == '0')) ? 0 : 1;
    }
    return g_mlp_backend;
}
AL_API int al_set_mlp_backend(int backend) {
    const int prev = mlp_backend();
    if (backend == 0 || backend == 1) g_mlp_backend = backend;
    return prev;
}

#define AL_MLP_CONFIGS(X)   \
    X(48, 128, 16, 2)       \
    X(64, 128, 16, 2)       \
    X(32, 128, 16, 2)       \
    X(16, 64, 64, 2)        \
    X(80, 64, 16, 1)        \
    X(64, 64, 16, 2)        \
    X(48, 64, 16, 2)        \
    X(32, 64, 16, 2)        \
    X(144, 64, 16, 1)

// Number of fp32 parameters of a supported configuration, or -1.
AL_API int al_mlp_num_params(int in_pad, int hidden, int out_pad, int n_hidden) {
#define X(I, Hh, O, N) \
    if (in_pad == I && hidden == Hh && out_pad == O && n_hidden == N) return Dims<I, Hh, O, N>::NPARAMS;
    AL_MLP_CONFIGS(X)
#undef X
    return -1;
}

// tcnn.Network forward.  x: fp16 [cap, ldx] (ldx >= in_pad, rows 16-byte aligned).
// Two fp32 outputs and one fp16 output, each a column window of y with an activation
// (fp32: 0 none, 1 sigmoid, 2 exp; fp16: 0 none, 1 relu); null pointer = unused.
// n_dev (optional device int) bounds the live rows without a host sync.
AL_API int al_mlp_forward(int in_pad, int hidden, int out_pad, int n_hidden, const float* params,
                          const void* x_half, int ldx, int cap, const int* n_dev,
                          float* o0, int o0_ld, int o0_col0, int o0_src0, int o0_ncols, int o0_act,
                          float* o1, int o1_ld, int o1_col0, int o1_src0, int o1_ncols, int o1_act,
                          void* h0_half, int h0_ld, int h0_col0, int h0_src0, int h0_ncols, int h0_act,
                          void* stream) {
    if (cap <= 0) return 0;
    AL_REQUIRE(params && x_half, "null pointer");
    AL_REQUIRE(ldx >= in_pad && ldx % 8 == 0, "ldx must be >= in_pad and a multiple of 8");
    MlpFwdArgs a;
    a.params = params; a.x = (const __half*)x_half; a.ldx = ldx; a.cap = cap; a.n_dev = n_dev;
    a.o0 = {o0, o0_ld, o0_col0, o0_src0, o0_ncols, o0_act};
    a.o1 = {o1, o1_ld, o1_col0, o1_src0, o1_ncols, o1_act};
    a.h0 = {(__half*)h0_half, h0_ld, h0_col0, h0_src0, h0_ncols, h0_act};
    a.sum = {nullptr, 0, 0, 0, 0, 0, nullptr, nullptr};
    a.hin = {nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr};
    if (mlp_backend() == 1) {
        const int r = al_tc_mlp_forward(in_pad, hidden, out_pad, n_hidden, a, (cudaStream_t)stream);
        if (r != -1) return r;
    }
#define X(I, Hh, O, N) \
    if (in_pad == I && hidden == Hh && out_pad == O && n_hidden == N) return launch_fwd<I, Hh, O, N>(a, (cudaStream_t)stream);
    AL_MLP_CONFIGS(X)
#undef X
    al_set_error("al_mlp_forward: unsupported MLP shape in=%d hidden=%d out=%d n_hidden=%d", in_pad, hidden, out_pad, n_hidden);
    return (int)cudaErrorInvalidValue;
}

int al_mlp_forward_args(int in_pad, int hidden, int out_pad, int n_hidden, const MlpFwdArgs& a, cudaStream_t st) {
    if (a.cap <= 0) return 0;
    if (mlp_backend() == 1) {
        const int r = al_tc_mlp_forward(in_pad, hidden, out_pad, n_hidden, a, st);
        if (r != -1) return r;
    }
    al_set_error("al_mlp_forward_args: shape in=%d hidden=%d out=%d n_hidden=%d needs the tcgen05 back end", in_pad, hidden, out_pad, n_hidden);
    return (int)cudaErrorInvalidValue;
}

// tcnn.Network backward: dparams += dL/dparams (fp32 atomics), dx = dL/dx (optional).
AL_API int al_mlp_backward(int in_pad, int hidden, int out_pad, int n_hidden, const float* params,
                           const void* x_half, int ldx, int cap, const int* n_dev, const float* dout,
                           int ld_dout, int dcol0, int dncols, const float* amax_dev, float* dparams,
                           float* dx, int dx_mode, int ld_dx, int dx_c0, int dx_n, void* stream) {
    if (cap <= 0) return 0;
    AL_REQUIRE(params && x_half && dout, "null pointer");
    AL_REQUIRE(ldx >= in_pad && ldx % 8 == 0, "ldx must be >= in_pad and a multiple of 8");
    AL_REQUIRE(dncols <= out_pad, "dncols exceeds the padded output width");
    MlpBwdArgs a = {};
    a.params = params; a.x = (const __half*)x_half; a.ldx = ldx; a.cap = cap; a.n_dev = n_dev;
    a.dout = dout; a.ld_dout = ld_dout; a.dcol0 = dcol0; a.dncols = dncols; a.amax_dev = amax_dev;
    a.dparams = dparams; a.dx = dx; a.dx_mode = dx_mode; a.ld_dx = ld_dx; a.dx_c0 = dx_c0; a.dx_n = dx_n;
    if (mlp_backend() == 1) {
        const int r = al_tc_mlp_backward(in_pad, hidden, out_pad, n_hidden, a, (cudaStream_t)stream);
        if (r != -1) return r;
    }
#define X(I, Hh, O, N) \
    if (in_pad == I && hidden == Hh && out_pad == O && n_hidden == N) return launch_bwd<I, Hh, O, N>(a, (cudaStream_t)stream);
    AL_MLP_CONFIGS(X)
#undef X
    al_set_error("al_mlp_backward: unsupported MLP shape in=%d hidden=%d out=%d n_hidden=%d", in_pad, hidden, out_pad, n_hidden);
    return (int)cudaErrorInvalidValue;
}

// amax[0] = max(amax[0], max |v|) over rows < n, columns [col0, col0+ncols).
AL_API int al_amax(const float* v, int ld, int col0, int ncols, int cap, const int* n_dev, float* amax,
                   void* stream) {
    if (cap <= 0) return 0;
    AL_REQUIRE(v && amax, "null pointer");
    k_amax<<<al_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(v, ld, col0, ncols, cap, n_dev, amax);
    AL_LAUNCH_CHECK();
    return 0;
}
