// Front-to-back compositing of K value channels (RGB + semantic logits + feature vector),
// depth and accumulated opacity along marched ray segments; forward, backward (training) and
// the in-place inference variant.
//
// Replaces and generalises the reference's 3-channel kernels:
//   torch_ngp/raymarching/src/raymarching.cu:547-636  composite_rays_train_forward
//   torch_ngp/raymarching/src/raymarching.cu:649-740  composite_rays_train_backward
//   torch_ngp/raymarching/src/raymarching.cu:868-961  composite_rays
// and the PyTorch compositing of semantic logits / features in
//   torch_ngp/nerf/renderer.py:243-311 (weights, depth, depth_variance, coordinates_map,
//   image, semantic, semantic_features).
//
// Differences from the reference kernels:
//  * K channels instead of 3; one warp owns a ray, lanes stride over the channels, so every
//    sample row is read with one coalesced request (the reference is one thread per ray).
//  * backward propagates the depth gradient (the reference drops it, raymarching.py:437-438).
//  * depth can be accumulated over the sample positions `tpos` (renderer.run() semantics,
//    renderer.py:273-275) instead of the running sum of deltas[.,1].
//  * optional second moment (for depth_variance) and position (coordinates_map) outputs.
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct RaySeg {
    uint32_t id, offset, count;
    bool valid;
};
__device__ __forceinline__ RaySeg load_seg(const int* __restrict__ rays, uint32_t n, uint32_t M) {
    RaySeg s;
    s.id = (uint32_t)rays[n * 3];
    s.offset = (uint32_t)rays[n * 3 + 1];
    s.count = (uint32_t)rays[n * 3 + 2];
    // empty ray, or ray whose segment overflowed the sample budget (raymarching.cu:568)
    s.valid = !(s.count == 0 || (unsigned long long)s.offset + s.count >= (unsigned long long)M);
    return s;
}

// Inclusive warp scans (product / sum) over the 32 samples of a chunk.
__device__ __forceinline__ float warp_scan_mul(float v, uint32_t lane) {
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v *= u;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_add(float v, uint32_t lane) {
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += u;
    }
    return v;
}

// Per-chunk compositing weights of 32 consecutive samples of one ray (lane = sample):
//   alpha = 1 - exp(-sigma scale dt),  T_i = prod_{j<i} (1 - alpha_j),  w = alpha T,  T_next = T (1 - alpha).
// The serial recurrence of the reference (raymarching.cu:587-615) becomes a warp-level multiplicative scan,
// which takes the transmittance chain off the critical path of the channel loads.
struct ChunkW {
    float w, T_next, dt, t;   // T_next: transmittance after this lane's sample, T_i (1 - alpha_i)
};
__device__ __forceinline__ ChunkW chunk_weights(const float* __restrict__ sigmas, uint32_t ld_sigma,
                                                const float* __restrict__ deltas, const float* __restrict__ tpos,
                                                size_t idx, bool valid, float sigma_scale, float& T_carry,
                                                float& t_carry, uint32_t lane) {
    float2 del = make_float2(0.f, 0.f);
    float sg = 0.f, t = 0.f;
    if (valid) {
        del = *reinterpret_cast<const float2*>(deltas + idx * 2);
        sg = sigmas[idx * ld_sigma];
        if (tpos) t = tpos[idx];
    }
    const float alpha = 1.0f - __expf(-(sg * sigma_scale) * del.x);
    const float p = warp_scan_mul(1.0f - alpha, lane);
    float Tex = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) Tex = 1.0f;
    Tex *= T_carry;
    ChunkW r;
    r.w = alpha * Tex;
    r.T_next = T_carry * p;
    r.dt = del.x;
    if (!tpos) {   // reference depth: running sum of deltas[.,1] (raymarching.cu:600-601)
        t = warp_scan_add(del.y, lane) + t_carry;
        t_carry = __shfl_sync(0xffffffffu, t, 31);
    }
    r.t = t;
    T_carry *= __shfl_sync(0xffffffffu, p, 31);
    return r;
}

// NC = channels per lane (K <= 32*NC).  One warp per ray; samples are taken 32 at a time: lane = sample for the
// weights, lane = channel for the K-channel accumulation (row reads are coalesced, 8 rows in flight).
template <int NC>
__global__ void __launch_bounds__(256) k_composite_train_fwd(
    const float* __restrict__ sigmas, uint32_t ld_sigma, const float* __restrict__ vals, uint32_t ldv,
    uint32_t K, const float* __restrict__ deltas, const float* __restrict__ tpos,
    const float* __restrict__ xyzs, const int* __restrict__ rays, uint32_t M, uint32_t N,
    float sigma_scale, float* __restrict__ weights_sum, float* __restrict__ depth,
    float* __restrict__ depth_sq, float* __restrict__ out, float* __restrict__ coords, uint32_t slices) {
    // `slices` > 1 (wide value rows, few rays): several warps per ray, each with its own 32 NC channels.  Every slice
    // recomputes the (cheap) weights and accumulates its channels in the same sample order, so the results do not
    // depend on `slices`; slice 0 writes the per-ray scalars.
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n = wid / slices, c0 = (wid - n * slices) * 32u * NC;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t K_all = K;
    K = K > c0 ? min(K - c0, 32u * NC) : 0u;
    vals += c0;
    const bool scalars = c0 == 0;
    if (!scalars) coords = nullptr;
    const RaySeg seg = load_seg(rays, n, M);
    float acc[NC];
    #pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = 0.f;
    float ws = 0.f, d = 0.f, d2 = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
    if (seg.valid) {
        float T_carry = 1.f, t_carry = 0.f;
        const float* vp = vals + (size_t)seg.offset * ldv;
        for (uint32_t base = 0; base < seg.count; base += 32) {
            const bool valid = base + lane < seg.count;
            const size_t idx = (size_t)seg.offset + base + lane;
            const ChunkW cw = chunk_weights(sigmas, ld_sigma, deltas, tpos, idx, valid, sigma_scale, T_carry, t_carry, lane);
            ws += cw.w;
            d = fmaf(cw.w, cw.t, d);
            d2 = fmaf(cw.w * cw.t, cw.t, d2);
            if (coords && valid) {
                cx = fmaf(cw.w, xyzs[idx * 3], cx);
                cy = fmaf(cw.w, xyzs[idx * 3 + 1], cy);
                cz = fmaf(cw.w, xyzs[idx * 3 + 2], cz);
            }
            const uint32_t nn = min(32u, seg.count - base);
            const float* rowp = vp + (size_t)base * ldv;
            constexpr int U = NC <= 5 ? 8 : (NC <= 20 ? 2 : 1);   // rows in flight (register budget)
            uint32_t i = 0;
            for (; i + U <= nn; i += U) {
                float v[U][NC];
                #pragma unroll
                for (int u = 0; u < U; ++u)
                    #pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        const uint32_t c = lane + 32 * j;
                        v[u][j] = (c < K) ? rowp[(size_t)(i + u) * ldv + c] : 0.f;
                    }
                #pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float wi = __shfl_sync(0xffffffffu, cw.w, i + u);
                    #pragma unroll
                    for (int j = 0; j < NC; ++j) acc[j] = fmaf(wi, v[u][j], acc[j]);
                }
            }
            for (; i < nn; ++i) {
                const float wi = __shfl_sync(0xffffffffu, cw.w, i);
                #pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const uint32_t c = lane + 32 * j;
                    if (c < K) acc[j] = fmaf(wi, rowp[(size_t)i * ldv + c], acc[j]);
                }
            }
        }
    }
    ws = warp_sum(ws); d = warp_sum(d); d2 = warp_sum(d2);
    if (coords) { cx = warp_sum(cx); cy = warp_sum(cy); cz = warp_sum(cz); }
    if (lane == 0 && scalars) {
        weights_sum[seg.id] = ws;
        depth[seg.id] = d;
        if (depth_sq) depth_sq[seg.id] = d2;
        if (coords) {
            coords[(size_t)seg.id * 3] = cx;
            coords[(size_t)seg.id * 3 + 1] = cy;
            coords[(size_t)seg.id * 3 + 2] = cz;
        }
    }
    #pragma unroll
    for (int j = 0; j < NC; ++j) {
        const uint32_t c = lane + 32 * j;
        if (c < K) out[(size_t)seg.id * K_all + c0 + c] = acc[j];
    }
}

// Backward.  With w_i = alpha_i T_i, T_{i+1} = T_i (1 - alpha_i), d alpha_i / d sigma_i =
// scale dt_i (1 - alpha_i):   dL/dsigma_i = scale dt_i [ T_{i+1} s_i - sum_{j>i} w_j s_j ],
// s_j = <g, v_j> + g_ws + g_depth t_j, and the tail sum is (S_final - S_prefix) exactly like the
// reference's (r_final - r) trick (raymarching.cu:711-716), with S_final recomputed from the
// saved per-ray outputs.  dL/dv_i = w_i g.
template <int NC>
__global__ void __launch_bounds__(256) k_composite_train_bwd(
    const float* __restrict__ g_ws, const float* __restrict__ g_depth, const float* __restrict__ g_out,
    const float* __restrict__ sigmas, uint32_t ld_sigma, const float* __restrict__ vals, uint32_t ldv,
    uint32_t K, const float* __restrict__ deltas, const float* __restrict__ tpos,
    const int* __restrict__ rays, const float* __restrict__ weights_sum, const float* __restrict__ depth,
    const float* __restrict__ out, uint32_t M, uint32_t N, float sigma_scale,
    float* __restrict__ g_sigmas, uint32_t ld_gsigma, float* __restrict__ g_vals, uint32_t ld_gv,
    float* __restrict__ amax_out) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const RaySeg seg = load_seg(rays, n, M);
    if (!seg.valid) return;
    float am = 0.f;   // running max |gradient| written by this warp (feeds the MLP backward's fp16 scale)
    float g[NC];
    float sfin = 0.f;
    #pragma unroll
    for (int j = 0; j < NC; ++j) {
        const uint32_t c = lane + 32 * j;
        g[j] = (c < K) ? g_out[(size_t)seg.id * K + c] : 0.f;
        if (c < K) sfin = fmaf(g[j], out[(size_t)seg.id * K + c], sfin);
    }
    const float gw = g_ws ? g_ws[seg.id] : 0.f;
    const float gd = g_depth ? g_depth[seg.id] : 0.f;
    sfin = warp_sum(sfin) + gw * weights_sum[seg.id] + gd * depth[seg.id];

    const float* sg = sigmas + (size_t)seg.offset * ld_sigma;
    const float* vp = vals + (size_t)seg.offset * ldv;
    const float* dl = deltas + (size_t)seg.offset * 2;
    float* gs = g_sigmas + (size_t)seg.offset * ld_gsigma;
    float* gv = g_vals + (size_t)seg.offset * ld_gv;
    float T = 1.f, srun = 0.f, trun = 0.f;
    #pragma unroll 2
    for (uint32_t s = 0; s < seg.count; ++s) {
        const float2 del = *reinterpret_cast<const float2*>(dl + 2 * s);
        const float alpha = 1.0f - __expf(-(sg[(size_t)s * ld_sigma] * sigma_scale) * del.x);
        const float w = alpha * T;
        T *= 1.0f - alpha;
        float p = 0.f;
        #pragma unroll
        for (int j = 0; j < NC; ++j) {
            const uint32_t c = lane + 32 * j;
            if (c < K) {
                p = fmaf(g[j], vp[(size_t)s * ldv + c], p);
                const float gvv = w * g[j];
                gv[(size_t)s * ld_gv + c] = gvv;
                am = fmaxf(am, fabsf(gvv));
            }
        }
        float t;
        if (tpos) t = tpos[seg.offset + s];
        else { trun += del.y; t = trun; }
        const float si = warp_sum(p) + gw + gd * t;
        srun = fmaf(w, si, srun);
        const float gsv = sigma_scale * del.x * (T * si - (sfin - srun));
        if (lane == 0) gs[(size_t)s * ld_gsigma] = gsv;
        am = fmaxf(am, fabsf(gsv));
    }
    if (amax_out) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
        if (lane == 0 && am > 0.f && am < 3.0e38f && am > *amax_out)   // racy pre-filter, atomicMax decides
            atomicMax(reinterpret_cast<int*>(amax_out), __float_as_int(am));
    }
}

// Narrow form of the materialised backward (K <= 4: the reference's own 3-channel operator, raymarching.cu:649-740):
// lane = sample.  The per-ray vectors g and <g, out> live in registers of every lane, each lane reads its own
// sample row (K consecutive floats) and the three serial recurrences of the reference loop (transmittance product,
// running sum of w s, running depth) become warp scans over 32-sample chunks, as in the forward.
template <int KS>
__global__ void __launch_bounds__(256) k_composite_train_bwd_narrow(
    const float* __restrict__ g_ws, const float* __restrict__ g_depth, const float* __restrict__ g_out,
    const float* __restrict__ sigmas, uint32_t ld_sigma, const float* __restrict__ vals, uint32_t ldv,
    uint32_t K, const float* __restrict__ deltas, const float* __restrict__ tpos,
    const int* __restrict__ rays, const float* __restrict__ weights_sum, const float* __restrict__ depth,
    const float* __restrict__ out, uint32_t M, uint32_t N, float sigma_scale,
    float* __restrict__ g_sigmas, uint32_t ld_gsigma, float* __restrict__ g_vals, uint32_t ld_gv,
    float* __restrict__ amax_out) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const RaySeg seg = load_seg(rays, n, M);
    if (!seg.valid) return;
    float g[KS];
    const float gw = g_ws ? g_ws[seg.id] : 0.f;
    const float gd = g_depth ? g_depth[seg.id] : 0.f;
    float sfin = gw * weights_sum[seg.id] + gd * depth[seg.id];
    #pragma unroll
    for (int c = 0; c < KS; ++c) {
        g[c] = ((uint32_t)c < K) ? g_out[(size_t)seg.id * K + c] : 0.f;
        if ((uint32_t)c < K) sfin = fmaf(g[c], out[(size_t)seg.id * K + c], sfin);
    }
    float am = 0.f, T_carry = 1.f, t_carry = 0.f, s_carry = 0.f;
    for (uint32_t base = 0; base < seg.count; base += 32) {
        const bool valid = base + lane < seg.count;
        const size_t idx = (size_t)seg.offset + base + lane;
        const ChunkW cw = chunk_weights(sigmas, ld_sigma, deltas, tpos, idx, valid, sigma_scale, T_carry, t_carry, lane);
        float p = 0.f;
        if (valid) {
            const float* row = vals + idx * ldv;
            #pragma unroll
            for (int c = 0; c < KS; ++c)
                if ((uint32_t)c < K) p = fmaf(g[c], row[c], p);
        }
        const float si = p + gw + gd * cw.t;
        const float srun = warp_scan_add(valid ? cw.w * si : 0.f, lane) + s_carry;   // inclusive: sum_{j<=i} w_j s_j
        s_carry = __shfl_sync(0xffffffffu, srun, 31);
        if (valid) {
            // dL/dsigma_i = scale dt_i [ T_{i+1} s_i - sum_{j>i} w_j s_j ]  (raymarching.cu:711-716)
            const float gsv = sigma_scale * cw.dt * (cw.T_next * si - (sfin - srun));
            g_sigmas[idx * ld_gsigma] = gsv;
            am = fmaxf(am, fabsf(gsv));
            float* grow = g_vals + idx * ld_gv;
            #pragma unroll
            for (int c = 0; c < KS; ++c)
                if ((uint32_t)c < K) { const float gvv = cw.w * g[c]; grow[c] = gvv; am = fmaxf(am, fabsf(gvv)); }
        }
    }
    if (amax_out) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
        if (lane == 0 && am > 0.f && am < 3.0e38f && am > *amax_out)
            atomicMax(reinterpret_cast<int*>(amax_out), __float_as_int(am));
    }
}

// Rank-1 form of the backward: dL/dvals[i, c] = w_i * g_out[ray(i), c] is a product of one number per
// sample and one vector per ray, so only w_i and dL/dsigma_i are written (8 bytes per sample instead of
// 4 (1 + K)); the consumers (the fused MLP backward loaders, csrc/mlp_tc.cu) rebuild dL/dvals on the fly.
// Lane = channel while the 32 rows of a chunk are read (all row loads in flight), then a butterfly
// transpose-reduction (31 shuffles for 32 dot products) makes lane = sample for the scans.
template <int NC>
__global__ void __launch_bounds__(256) k_composite_train_bwd_w(
    const float* __restrict__ g_ws, const float* __restrict__ g_depth, const float* __restrict__ g_out,
    const float* __restrict__ sigmas, uint32_t ld_sigma, const float* __restrict__ vals, uint32_t ldv,
    uint32_t K, const float* __restrict__ deltas, const float* __restrict__ tpos,
    const int* __restrict__ rays, const float* __restrict__ weights_sum, const float* __restrict__ depth,
    const float* __restrict__ out, uint32_t M, uint32_t N, float sigma_scale,
    float* __restrict__ w_out, float* __restrict__ g_sigmas, float* __restrict__ amax_out, int phase) {
    // phase 0: everything in one pass, a warp per ray.
    // phases 1 + 2 (wide value rows, few rays): phase 1 is a warp per (ray, 32-row chunk) that only forms the row dot
    // products <g, v_i> -- the part that reads vals -- and parks them in g_sigmas; phase 2 is the per-ray scan reading
    // them back.  The dot products are formed by the same instructions in both forms, so the results are identical.
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t n = wid, only_base = 0xffffffffu;
    if (phase == 1) {
        // warp -> (ray, chunk): 32 warps per ray, warp c takes chunks c, c + 32, ... (rays rarely exceed 1024 samples)
        n = wid >> 5;
        only_base = (wid & 31u) * 32u;
    }
    if (n >= N) return;
    const RaySeg seg = load_seg(rays, n, M);
    if (!seg.valid) return;
    if (phase == 1 && only_base >= seg.count) return;
    float g[NC];
    float sfin = 0.f, gmax = 0.f;
    #pragma unroll
    for (int j = 0; j < NC; ++j) {
        const uint32_t c = lane + 32 * j;
        g[j] = (c < K) ? g_out[(size_t)seg.id * K + c] : 0.f;
        if (c < K) sfin = fmaf(g[j], out[(size_t)seg.id * K + c], sfin);
        gmax = fmaxf(gmax, fabsf(g[j]));
    }
    const float gw = g_ws ? g_ws[seg.id] : 0.f;
    const float gd = g_depth ? g_depth[seg.id] : 0.f;
    sfin = warp_sum(sfin) + gw * weights_sum[seg.id] + gd * depth[seg.id];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));

    const float* vp = vals + (size_t)seg.offset * ldv;
    float T_carry = 1.f, t_carry = 0.f, s_carry = 0.f, am = 0.f;
    for (uint32_t base = (phase == 1 ? only_base : 0u); base < seg.count; base += (phase == 1 ? 1024u : 32u)) {
        const uint32_t nn = min(32u, seg.count - base);
        float p[32];
        if (phase == 2) {
            p[0] = (lane < nn) ? g_sigmas[(size_t)seg.offset + base + lane] : 0.f;
        } else {
        // p[i] = this lane's share of <g, v_i>.  Loads are unconditional (row / channel indices clamped, g = 0 for
        // channels >= K, rows >= nn are ignored later) and issued U rows at a time ahead of the FMAs.
        const float* rowp = vp + (size_t)base * ldv;
        constexpr int U = NC <= 5 ? 8 : (NC <= 20 ? 2 : 1);
        uint32_t cc[NC];
        #pragma unroll
        for (int j = 0; j < NC; ++j) cc[j] = min(lane + 32 * j, K - 1);
        #pragma unroll
        for (int i0 = 0; i0 < 32; i0 += U) {
            float v[U][NC];
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t ii = min((uint32_t)(i0 + u), nn - 1);
                #pragma unroll
                for (int j = 0; j < NC; ++j) v[u][j] = __ldg(rowp + (size_t)ii * ldv + cc[j]);
            }
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                float a = 0.f;
                #pragma unroll
                for (int j = 0; j < NC; ++j) a = fmaf(g[j], v[u][j], a);
                p[i0 + u] = a;
            }
        }
        // butterfly: afterwards p[0] of lane L = sum over lanes of p[L]
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const bool hi = (lane & o) != 0;
            #pragma unroll
            for (int k = 0; k < o; ++k) {
                const float send = hi ? p[k] : p[k + o];
                const float keep = hi ? p[k + o] : p[k];
                p[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        }
        if (phase == 1) {
            if (lane < nn) g_sigmas[(size_t)seg.offset + base + lane] = p[0];
            continue;
        }
        const bool valid = lane < nn;
        const size_t idx = (size_t)seg.offset + base + lane;
        const ChunkW cw = chunk_weights(sigmas, ld_sigma, deltas, tpos, idx, valid, sigma_scale, T_carry, t_carry, lane);
        const float si = p[0] + gw + gd * cw.t;
        const float srun = warp_scan_add(cw.w * si, lane) + s_carry;
        s_carry = __shfl_sync(0xffffffffu, srun, 31);
        const float gsv = sigma_scale * cw.dt * (cw.T_next * si - (sfin - srun));
        if (valid) {
            w_out[idx] = cw.w;
            g_sigmas[idx] = gsv;
            am = fmaxf(am, fmaxf(fabsf(gsv), cw.w * gmax));
        }
    }
    if (amax_out) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
        if (lane == 0 && am > 0.f && am < 3.0e38f && am > *amax_out)   // racy pre-filter, atomicMax decides
            atomicMax(reinterpret_cast<int*>(amax_out), __float_as_int(am));
    }
}

// In-place inference compositing (raymarching.cu:868-952), K channels, warp per alive ray.
// T_i = 1 - weight_sum; stops at deltas[.,0] == 0 (exhausted ray) or after a sample that started
// with T < 1e-4; rays that stopped early get rays_t = -1.
// WONLY: the weights-only form (al_composite_rays_weights): no value channels are read or accumulated; instead the
// compositing weight of every slot of the wave goes to w_out (0 for the slots behind the point where the ray stopped), for
// the head kernels that fold the K-channel sums into their output epilogue (OutSum, mlp_args.cuh).
template <int NC, bool WONLY = false>
__global__ void __launch_bounds__(256) k_composite_rays(
    uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive, float* __restrict__ rays_t,
    const float* __restrict__ sigmas, uint32_t ld_sigma, const float* __restrict__ vals, uint32_t ldv,
    uint32_t K, const float* __restrict__ deltas, const float* __restrict__ tpos,
    const float* __restrict__ xyzs, float sigma_scale, float* __restrict__ weights_sum,
    float* __restrict__ depth, float* __restrict__ depth_sq, float* __restrict__ out,
    float* __restrict__ coords, float* __restrict__ w_out = nullptr) {
    // Warp per ray, 32 steps per round.  The per-step arithmetic and its ORDER are the reference's (raymarching.cu:
    // 895-947: T = 1 - weight_sum, w = alpha T, sums by fused multiply-add in step order), so results stay bit-identical;
    // what changes is the schedule: the 32 alphas of a round are evaluated by the 32 lanes at once (one load latency
    // and one exp per round instead of per step), the short dependent chain ws += alpha (1 - ws) runs on broadcast
    // registers, and the K-channel rows of the round are then accumulated with several row loads in flight.
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    float t = rays_t[n];
    const size_t base = (size_t)n * n_step;
    float acc[NC];
    #pragma unroll
    for (int j = 0; j < NC; ++j) {
        const uint32_t c = lane + 32 * j;
        acc[j] = (!WONLY && c < K) ? out[(size_t)index * K + c] : 0.f;
    }
    float ws = weights_sum[index], d = depth[index];
    float d2 = depth_sq ? depth_sq[index] : 0.f;
    float cacc = (coords && lane < 3) ? coords[(size_t)index * 3 + lane] : 0.f;
    uint32_t step = 0;
    bool stopped = false;
    while (step < n_step && !stopped) {
        const uint32_t s = step + lane;
        float2 del = make_float2(0.f, 0.f);
        float alpha = 0.f, td = 0.f;
        if (s < n_step) {
            del = *reinterpret_cast<const float2*>(deltas + (base + s) * 2);
            if (del.x != 0.f) {
                alpha = 1.0f - __expf(-(sigmas[(base + s) * ld_sigma] * sigma_scale) * del.x);
                if (tpos) td = tpos[base + s];
            }
        }
        const uint32_t round = min(32u, n_step - step);
        // steps before the first exhausted slot (deltas[., 0] == 0)
        const uint32_t live = __ffs(__ballot_sync(0xffffffffu, !(s < n_step && del.x != 0.f))) - 1u;   // 32 if none (ffs(0) = 0)
        const uint32_t n_live = min(round, live);
        float myw = 0.f;
        uint32_t n_acc = 0;                                          // steps of this round that are accumulated
        for (uint32_t k = 0; k < n_live; ++k) {
            const float a = __shfl_sync(0xffffffffu, alpha, k);
            const float dy = __shfl_sync(0xffffffffu, del.y, k);
            const float T = 1.0f - ws;
            const float w = a * T;
            ws += w;
            t += dy;
            const float tdk = tpos ? __shfl_sync(0xffffffffu, td, k) : t;
            d = fmaf(w, tdk, d);
            d2 = fmaf(w * tdk, tdk, d2);
            if (lane == k) myw = w;
            n_acc = k + 1;
            if (T < 1e-4f) { stopped = true; break; }                // the sample that started with T < 1e-4 is the last one
        }
        if (!stopped && n_live < round) stopped = true;              // exhausted ray
        if (WONLY) {
            // the weights of this round's slots (0 behind the stopping point), and the weighted position sum
            if (lane < round) w_out[base + step + lane] = lane < n_acc ? myw : 0.f;
            if (coords) {
                for (uint32_t q = 0; q < n_acc; ++q) {
                    const float w = __shfl_sync(0xffffffffu, myw, q);
                    if (lane < 3) cacc = fmaf(w, xyzs[(base + step + q) * 3 + lane], cacc);
                }
            }
        } else {
        // K-channel accumulation of the round's n_acc rows, in step order per channel
        const float* vrow = vals + (base + step) * ldv;
        uint32_t k = 0;
        for (; k + 4 <= n_acc; k += 4) {
            float v[4][NC];
            #pragma unroll
            for (int u = 0; u < 4; ++u)
                #pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const uint32_t c = lane + 32 * j;
                    v[u][j] = (c < K) ? vrow[(size_t)(k + u) * ldv + c] : 0.f;
                }
            float xc[4] = {0.f, 0.f, 0.f, 0.f};
            if (coords && lane < 3) {
                #pragma unroll
                for (int u = 0; u < 4; ++u) xc[u] = xyzs[(base + step + k + u) * 3 + lane];
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float w = __shfl_sync(0xffffffffu, myw, k + u);
                #pragma unroll
                for (int j = 0; j < NC; ++j) acc[j] = fmaf(w, v[u][j], acc[j]);
                cacc = fmaf(w, xc[u], cacc);                         // lanes >= 3 (or no coords): adds w * 0 to an unused value
            }
        }
        for (; k < n_acc; ++k) {
            const float w = __shfl_sync(0xffffffffu, myw, k);
            #pragma unroll
            for (int j = 0; j < NC; ++j) {
                const uint32_t c = lane + 32 * j;
                if (c < K) acc[j] = fmaf(w, vrow[(size_t)k * ldv + c], acc[j]);
            }
            if (coords && lane < 3) cacc = fmaf(w, xyzs[(base + step + k) * 3 + lane], cacc);
        }
        }
        step += round;
    }
    if (WONLY) {
        // slots of the rounds this ray never reached (it stopped earlier): weight 0
        for (uint32_t sidx = step + lane; sidx < n_step; sidx += 32) w_out[base + sidx] = 0.f;
    }
    if (lane == 0) {
        rays_t[n] = stopped ? -1.0f : t;
        weights_sum[index] = ws;
        depth[index] = d;
        if (depth_sq) depth_sq[index] = d2;
    }
    if (coords && lane < 3) coords[(size_t)index * 3 + lane] = cacc;
    if (!WONLY) {
        #pragma unroll
        for (int j = 0; j < NC; ++j) {
            const uint32_t c = lane + 32 * j;
            if (c < K) out[(size_t)index * K + c] = acc[j];
        }
    }
}


// ------------------------------------------------------------------------------------------ alive-prefix compaction
// Training-time early termination.  The reference's marched inference kernel stops a ray after the sample that brings
// its transmittance below 1e-4 (raymarching.cu:929-935); its training kernels composite every marched sample
// (the break is commented out, raymarching.cu:593-594,697), although everything behind that point carries a total
// weight < 1e-4.  The samples of a ray whose transmittance BEFORE the sample is still >= t_thresh form a prefix of
// its segment, so "dropping the dead tail" is a per-ray prefix copy into a dense sample set, and every later kernel
// (colour / semantic heads, K-channel compositing, the whole backward, the hash-grid scatter) runs on the alive
// samples only.  Same alpha / transmittance arithmetic as chunk_weights above.

// alive[n] = number of leading samples of ray n with T_before >= t_thresh (0 for empty / dropped rays).
__global__ void __launch_bounds__(256) k_alive_count(const float* __restrict__ sigmas, uint32_t ld_sigma,
                                                     const float* __restrict__ deltas, const int* __restrict__ rays,
                                                     uint32_t M, uint32_t N, float sigma_scale, float t_thresh,
                                                     int* __restrict__ alive) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const RaySeg seg = load_seg(rays, n, M);
    uint32_t cnt = 0;
    if (seg.valid) {
        float T_carry = 1.f;
        for (uint32_t base = 0; base < seg.count; base += 32) {
            const bool valid = base + lane < seg.count;
            const size_t idx = (size_t)seg.offset + base + lane;
            float dt = 0.f, sg = 0.f;
            if (valid) {
                dt = deltas[idx * 2];
                sg = sigmas[idx * ld_sigma];
            }
            const float alpha = 1.0f - __expf(-(sg * sigma_scale) * dt);
            const float p = warp_scan_mul(1.0f - alpha, lane);
            float Tex = __shfl_up_sync(0xffffffffu, p, 1);
            if (lane == 0) Tex = 1.0f;
            Tex *= T_carry;
            const unsigned live = __ballot_sync(0xffffffffu, valid && Tex >= t_thresh);
            // transmittance never increases (alpha in [0,1]): the alive lanes are a prefix; stop at the first dead one
            const unsigned dead = ~live;
            const uint32_t lead = dead ? (uint32_t)(__ffs(dead) - 1) : 32u;
            cnt += lead;
            if (lead < 32u) break;
            T_carry *= __shfl_sync(0xffffffffu, p, 31);
        }
    }
    if (lane == 0) alive[n] = (int)cnt;
}

// rays_c[n] = (ray id, exclusive scan of alive, alive[n]); meta_c = {total, total}.  One CTA (N is a ray batch).
__global__ void __launch_bounds__(1024) k_alive_scan(const int* __restrict__ alive, const int* __restrict__ rays,
                                                     uint32_t N, int* __restrict__ rays_c, int* __restrict__ meta_c) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t start = 0; start < N; start += blockDim.x) {
        const uint32_t n = start + tid;
        const uint32_t c = n < N ? (uint32_t)alive[n] : 0u;
        uint32_t v = c;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= (uint32_t)o) v += u;
        }
        if (lane == 31) warp_sums[wid] = v;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_sums[lane];
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= (uint32_t)o) w += u;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = v + (wid ? warp_sums[wid - 1] : 0u) + carry_s;
        if (n < N) {
            rays_c[n * 3] = rays[n * 3];
            rays_c[n * 3 + 1] = (int)(incl - c);
            rays_c[n * 3 + 2] = (int)c;
        }
        __syncthreads();
        if (tid == blockDim.x - 1) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) { meta_c[0] = (int)carry_s; meta_c[1] = (int)carry_s; }
}

struct AliveCopy {
    const float* xyzs; const float* deltas; const float* tpos; const int* sray; const float* sigma;
    const uint4* x_enc; const uint4* h16; uint32_t enc_vec;          // 16-byte vectors per x_enc row (in_pad / 8)
    float* xyzs_c; float* deltas_c; float* tpos_c; int* sray_c; float* sigma_c; uint32_t ld_sigma_c;
    uint4* x_enc_c; uint4* h16_c;
};

// n elements src -> dst by one warp, four loads in flight per lane before the first store (the two ranges never
// overlap, but the compiler cannot know that from the pointers inside a struct, so the batching is explicit).
template <typename T>
__device__ __forceinline__ void copy_range(T* __restrict__ dst, const T* __restrict__ src, uint32_t n, uint32_t lane) {
    uint32_t i = lane;
    for (; i + 96 < n; i += 128) {
        const T v0 = src[i], v1 = src[i + 32], v2 = src[i + 64], v3 = src[i + 96];
        dst[i] = v0; dst[i + 32] = v1; dst[i + 64] = v2; dst[i + 96] = v3;
    }
    for (; i < n; i += 32) dst[i] = src[i];
}

// One warp per ray: the alive prefix [offset, offset + a) of every per-sample array -> [offset_c, offset_c + a).
// Each array is a contiguous range on both sides: lanes stride over its words / 16-byte vectors (coalesced).
__global__ void __launch_bounds__(256) k_alive_copy(const int* __restrict__ rays, const int* __restrict__ rays_c,
                                                    uint32_t N, AliveCopy a) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t cnt = (uint32_t)rays_c[n * 3 + 2];
    if (cnt == 0) return;
    const size_t src = (size_t)(uint32_t)rays[n * 3 + 1], dst = (size_t)(uint32_t)rays_c[n * 3 + 1];
    copy_range(a.xyzs_c + dst * 3, a.xyzs + src * 3, cnt * 3, lane);
    copy_range(a.deltas_c + dst * 2, a.deltas + src * 2, cnt * 2, lane);
    if (a.tpos) copy_range(a.tpos_c + dst, a.tpos + src, cnt, lane);
    copy_range(a.sray_c + dst, a.sray + src, cnt, lane);
    if (a.ld_sigma_c == 1) {
        copy_range(a.sigma_c + dst, a.sigma + src, cnt, lane);
    } else {
        for (uint32_t i = lane; i < cnt; i += 32) a.sigma_c[(dst + i) * a.ld_sigma_c] = a.sigma[src + i];
    }
    copy_range(a.h16_c + dst * 4, a.h16 + src * 4, cnt * 4, lane);
    copy_range(a.x_enc_c + dst * a.enc_vec, a.x_enc + src * a.enc_vec, cnt * a.enc_vec, lane);
}

}  // namespace

#define AL_DISPATCH_NC(K, CALL)                                   \
    do {                                                          \
        if ((K) <= 32) { constexpr int NC = 1; CALL; }            \
        else if ((K) <= 96) { constexpr int NC = 3; CALL; }       \
        else if ((K) <= 160) { constexpr int NC = 5; CALL; }      \
        else if ((K) <= 640) { constexpr int NC = 20; CALL; }     \
        else { constexpr int NC = 40; CALL; }                     \
    } while (0)

// composite_rays_train_forward (raymarching.h:12), K channels.
//   sigmas: element i at sigmas[i*ld_sigma]; vals: row i at vals + i*ldv, K channels used
//   deltas [M,2]; tpos [M] optional (null -> reference depth: running sum of deltas[.,1])
//   xyzs/coords optional (coordinates_map); depth_sq optional (sum w t^2)
//   out [N,K], weights_sum [N], depth [N]: indexed by ray id rays[n,0]
AL_API int al_composite_train_fwd(const float* sigmas, uint32_t ld_sigma, const float* vals, uint32_t ldv,
                                  uint32_t K, const float* deltas, const float* tpos, const float* xyzs,
                                  const int* rays, uint32_t M, uint32_t N, float sigma_scale,
                                  float* weights_sum, float* depth, float* depth_sq, float* out,
                                  float* coords, void* stream) {
    if (N == 0) return 0;
    AL_REQUIRE(sigmas && vals && deltas && rays && weights_sum && depth && out, "null pointer");
    AL_REQUIRE(K >= 1 && K <= 1280 && ldv >= K && ld_sigma >= 1, "bad channel layout");
    AL_REQUIRE(!coords || xyzs, "coords output needs xyzs");
    if (K > 160 && (unsigned long long)N * 32 < (unsigned long long)al_num_sms() * 2048) {
        // wide value rows and too few rays to fill the machine with one warp each (C5: 1024 rays x 517 channels):
        // warps per ray = slices of 160 channels, eight rows in flight each
        const uint32_t slices = (K + 159) / 160;
        const unsigned grid = al_div_up((unsigned long long)N * slices * 32, 256);
        k_composite_train_fwd<5><<<grid, 256, 0, (cudaStream_t)stream>>>(sigmas, ld_sigma, vals, ldv, K, deltas, tpos, xyzs, rays,
                                                                         M, N, sigma_scale, weights_sum, depth, depth_sq, out,
                                                                         coords, slices);
        AL_LAUNCH_CHECK();
        return 0;
    }
    const unsigned grid = al_div_up((unsigned long long)N * 32, 256);
    AL_DISPATCH_NC(K, (k_composite_train_fwd<NC><<<grid, 256, 0, (cudaStream_t)stream>>>(
                          sigmas, ld_sigma, vals, ldv, K, deltas, tpos, xyzs, rays, M, N, sigma_scale,
                          weights_sum, depth, depth_sq, out, coords, 1u)));
    AL_LAUNCH_CHECK();
    return 0;
}

// composite_rays_train_backward (raymarching.h:13), K channels + depth gradient.
// g_ws / g_depth may be null (treated as zero).  Gradients of samples outside valid segments
// are NOT written (the caller zero-fills, as the reference wrapper does).  amax_out (optional
// device float, zeroed by the caller) receives max |gradient written| via atomicMax.
AL_API int al_composite_train_bwd(const float* g_ws, const float* g_depth, const float* g_out,
                                  const float* sigmas, uint32_t ld_sigma, const float* vals, uint32_t ldv,
                                  uint32_t K, const float* deltas, const float* tpos, const int* rays,
                                  const float* weights_sum, const float* depth, const float* out, uint32_t M,
                                  uint32_t N, float sigma_scale, float* g_sigmas, uint32_t ld_gsigma,
                                  float* g_vals, uint32_t ld_gv, float* amax_out, void* stream) {
    if (N == 0) return 0;
    AL_REQUIRE(g_out && sigmas && vals && deltas && rays && weights_sum && depth && out && g_sigmas && g_vals,
               "null pointer");
    AL_REQUIRE(K >= 1 && K <= 1280 && ldv >= K && ld_gv >= K, "bad channel layout");
    const unsigned grid = al_div_up((unsigned long long)N * 32, 256);
    if (K <= 4) {
        k_composite_train_bwd_narrow<4><<<grid, 256, 0, (cudaStream_t)stream>>>(
            g_ws, g_depth, g_out, sigmas, ld_sigma, vals, ldv, K, deltas, tpos, rays, weights_sum, depth, out, M, N,
            sigma_scale, g_sigmas, ld_gsigma, g_vals, ld_gv, amax_out);
        AL_LAUNCH_CHECK();
        return 0;
    }
    AL_DISPATCH_NC(K, (k_composite_train_bwd<NC><<<grid, 256, 0, (cudaStream_t)stream>>>(
                          g_ws, g_depth, g_out, sigmas, ld_sigma, vals, ldv, K, deltas, tpos, rays, weights_sum,
                          depth, out, M, N, sigma_scale, g_sigmas, ld_gsigma, g_vals, ld_gv, amax_out)));
    AL_LAUNCH_CHECK();
    return 0;
}

// Rank-1 backward used by the fused training path: writes the compositing weight w [M] and dL/dsigma [M] only
// (dL/dvals[i, c] = w[i] * g_out[ray(i), c] is rebuilt by the consumers).  Same inputs as al_composite_train_bwd.
AL_API int al_composite_train_bwd_weights(const float* g_ws, const float* g_depth, const float* g_out,
                                          const float* sigmas, uint32_t ld_sigma, const float* vals, uint32_t ldv,
                                          uint32_t K, const float* deltas, const float* tpos, const int* rays,
                                          const float* weights_sum, const float* depth, const float* out, uint32_t M,
                                          uint32_t N, float sigma_scale, float* w_out, float* g_sigmas,
                                          float* amax_out, void* stream) {
    if (N == 0) return 0;
    AL_REQUIRE(g_out && sigmas && vals && deltas && rays && weights_sum && depth && out && w_out && g_sigmas,
               "null pointer");
    AL_REQUIRE(K >= 1 && K <= 1280 && ldv >= K, "bad channel layout");
    const unsigned grid = al_div_up((unsigned long long)N * 32, 256);
    if (K > 160 && (unsigned long long)N * 32 < (unsigned long long)al_num_sms() * 2048) {
        // wide value rows, few rays (C5): the row dot products in a chunk-parallel pass, then the per-ray scan
        const unsigned grid1 = al_div_up((unsigned long long)N * 32 * 32, 256);
        AL_DISPATCH_NC(K, (k_composite_train_bwd_w<NC><<<grid1, 256, 0, (cudaStream_t)stream>>>(
                              g_ws, g_depth, g_out, sigmas, ld_sigma, vals, ldv, K, deltas, tpos, rays, weights_sum,
                              depth, out, M, N, sigma_scale, w_out, g_sigmas, amax_out, 1)));
        AL_LAUNCH_CHECK();
        AL_DISPATCH_NC(K, (k_composite_train_bwd_w<NC><<<grid, 256, 0, (cudaStream_t)stream>>>(
                              g_ws, g_depth, g_out, sigmas, ld_sigma, vals, ldv, K, deltas, tpos, rays, weights_sum,
                              depth, out, M, N, sigma_scale, w_out, g_sigmas, amax_out, 2)));
        AL_LAUNCH_CHECK();
        return 0;
    }
    AL_DISPATCH_NC(K, (k_composite_train_bwd_w<NC><<<grid, 256, 0, (cudaStream_t)stream>>>(
                          g_ws, g_depth, g_out, sigmas, ld_sigma, vals, ldv, K, deltas, tpos, rays, weights_sum,
                          depth, out, M, N, sigma_scale, w_out, g_sigmas, amax_out, 0)));
    AL_LAUNCH_CHECK();
    return 0;
}

// composite_rays (raymarching.h:16), K channels, in place.
AL_API int al_composite_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, float* rays_t,
                             const float* sigmas, uint32_t ld_sigma, const float* vals, uint32_t ldv, uint32_t K,
                             const float* deltas, const float* tpos, const float* xyzs, float sigma_scale,
                             float* weights_sum, float* depth, float* depth_sq, float* out, float* coords,
                             void* stream) {
    if (n_alive == 0) return 0;
    AL_REQUIRE(rays_alive && rays_t && sigmas && vals && deltas && weights_sum && depth && out, "null pointer");
    AL_REQUIRE(K >= 1 && K <= 1280 && ldv >= K, "bad channel layout");
    AL_REQUIRE(!coords || xyzs, "coords output needs xyzs");
    const unsigned grid = al_div_up((unsigned long long)n_alive * 32, 256);
    AL_DISPATCH_NC(K, (k_composite_rays<NC><<<grid, 256, 0, (cudaStream_t)stream>>>(
                          n_alive, n_step, rays_alive, rays_t, sigmas, ld_sigma, vals, ldv, K, deltas, tpos, xyzs,
                          sigma_scale, weights_sum, depth, depth_sq, out, coords)));
    AL_LAUNCH_CHECK();
    return 0;
}

// The weights-only half of composite_rays: everything that does not need the value channels (weights_sum, depth,
// depth_sq, coords, rays_t, the stopping rule) plus the per-slot compositing weights w_out [n_alive * n_step] for the
// head kernels that fold the K-channel sums into their epilogue (al_field_heads_forward_sum).
AL_API int al_composite_rays_weights(uint32_t n_alive, uint32_t n_step, const int* rays_alive, float* rays_t,
                                     const float* sigmas, uint32_t ld_sigma, const float* deltas, const float* tpos,
                                     const float* xyzs, float sigma_scale, float* weights_sum, float* depth,
                                     float* depth_sq, float* coords, float* w_out, void* stream) {
    if (n_alive == 0) return 0;
    AL_REQUIRE(rays_alive && rays_t && sigmas && deltas && weights_sum && depth && w_out, "null pointer");
    AL_REQUIRE(!coords || xyzs, "coords output needs xyzs");
    const unsigned grid = al_div_up((unsigned long long)n_alive * 32, 256);
    k_composite_rays<1, true><<<grid, 256, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, sigmas, ld_sigma,
                                                                   nullptr, 0, 0, deltas, tpos, xyzs, sigma_scale, weights_sum,
                                                                   depth, depth_sq, nullptr, coords, w_out);
    AL_LAUNCH_CHECK();
    return 0;
}

// Alive-prefix compaction of a marched sample set (training-time early termination, see k_alive_count).
//   in : sigma [M] (density per sample), deltas [M,2], rays [N,3], xyzs [M,3], tpos [M] (optional), sray [M],
//        x_enc [M, in_pad] fp16 (in_pad a multiple of 8), h16 [M,16] fp32
//   out: rays_c [N,3] = (ray id, compact offset, alive count), meta_c [2] = {alive samples, alive samples},
//        the alive rows of every array, densely packed in ray order; sigma goes to sigma_c[i * ld_sigma_c]
//   alive_ws: int [N] scratch.  t_thresh <= 0 keeps every sample (a plain copy).
AL_API int al_compact_alive(const float* sigma, const float* deltas, const int* rays, uint32_t M, uint32_t N,
                            float sigma_scale, float t_thresh, const float* xyzs, const float* tpos, const int* sray,
                            const void* x_enc, uint32_t in_pad, const float* h16, int* rays_c, int* meta_c,
                            float* xyzs_c, float* deltas_c, float* tpos_c, int* sray_c, void* x_enc_c, float* h16_c,
                            float* sigma_c, uint32_t ld_sigma_c, int* alive_ws, void* stream) {
    if (N == 0) return 0;
    AL_REQUIRE(sigma && deltas && rays && xyzs && sray && x_enc && h16, "null input");
    AL_REQUIRE(rays_c && meta_c && xyzs_c && deltas_c && sray_c && x_enc_c && h16_c && sigma_c && alive_ws, "null output");
    AL_REQUIRE(!tpos || tpos_c, "tpos_c required with tpos");
    AL_REQUIRE(in_pad % 8 == 0 && ld_sigma_c >= 1, "in_pad must be a multiple of 8");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = al_div_up((unsigned long long)N * 32, 256);
    k_alive_count<<<blocks, 256, 0, st>>>(sigma, 1, deltas, rays, M, N, sigma_scale, t_thresh, alive_ws);
    AL_LAUNCH_CHECK();
    k_alive_scan<<<1, 1024, 0, st>>>(alive_ws, rays, N, rays_c, meta_c);
    AL_LAUNCH_CHECK();
    AliveCopy a;
    a.xyzs = xyzs; a.deltas = deltas; a.tpos = tpos; a.sray = sray; a.sigma = sigma;
    a.x_enc = (const uint4*)x_enc; a.h16 = (const uint4*)h16; a.enc_vec = in_pad / 8;
    a.xyzs_c = xyzs_c; a.deltas_c = deltas_c; a.tpos_c = tpos_c; a.sray_c = sray_c; a.sigma_c = sigma_c;
    a.ld_sigma_c = ld_sigma_c; a.x_enc_c = (uint4*)x_enc_c; a.h16_c = (uint4*)h16_c;
    k_alive_copy<<<blocks, 256, 0, st>>>(rays, rays_c, N, a);
    AL_LAUNCH_CHECK();
    return 0;
}
