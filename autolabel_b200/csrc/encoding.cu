// Multiresolution hash-grid encoding (forward / backward), frequency and spherical-harmonics
// encodings for sm_100a.
//
// Replaces:
//   torch_ngp/gridencoder/src/gridencoder.cu:35-72    fast_hash / get_grid_index
//   torch_ngp/gridencoder/src/gridencoder.cu:75-223   kernel_grid           (forward, dy_dx)
//   torch_ngp/gridencoder/src/gridencoder.cu:226-312  kernel_grid_backward  (scatter-add)
//   torch_ngp/gridencoder/src/gridencoder.cu:315-341  kernel_input_backward
//   torch_ngp/shencoder/src/shencoder.cu:50-73        SH basis, degree <= 4 (tcnn SphericalHarmonics
//                                                      as used by autolabel/models.py:97-101,205-207)
//   tcnn Frequency encoding as used by autolabel/models.py:15-59 (external dependency, unpinned:
//   sin/cos(2^k pi x), dim-major, (sin, cos) interleaved per frequency; DESIGN.md "oracle")
//
// Index arithmetic (scale, resolution, dense stride vs. hash, % hashmap_size) is bit-exact with
// the reference; interpolation uses the same operation order so values agree to rounding.
//
// B200 layout: the whole table (<= 57 MB fp32 for the hg+freq configuration) is L2-resident
// (126 MB L2), so the fused encoder runs one thread per sample over all levels with the eight
// corner gathers of a level in flight together (8-byte float2 loads through the read-only
// path) and writes one contiguous fp16 feature row per sample for the tensor-core MLP; the
// reference-layout entry points ([L,B,C] outputs, level-major grid) are kept for drop-in use.
// The backward scatter uses vectorised red.global.add.v2.f32 (one 8-byte reduction per corner)
// in level-major launch order so concurrent CTAs hit the same level's slice of the table.
#include "common.cuh"
#include "mlp_args.cuh"

namespace {

__device__ __forceinline__ uint32_t hash3(uint32_t x, uint32_t y, uint32_t z) {
    return (x * 1u) ^ (y * 2654435761u) ^ (z * 805459861u);
}

// index % m without the integer division in the two common cases: m a power of two (every hashed
// level: 2^19) and index < m (every dense level).  Bit-identical to `%`.
__device__ __forceinline__ uint32_t fast_mod(uint32_t index, uint32_t m) {
    if ((m & (m - 1u)) == 0u) return index & (m - 1u);
    return index < m ? index : index % m;
}

// gridencoder.cu:54-72 for D = 3, returning the entry index (without the channel factor).
__device__ __forceinline__ uint32_t grid_index3(uint32_t gridtype, uint32_t hashmap_size,
                                                uint32_t resolution, uint32_t x, uint32_t y, uint32_t z) {
    uint32_t stride = 1, index = 0;
    if (stride <= hashmap_size) { index += x * stride; stride *= (resolution + 1); }
    if (stride <= hashmap_size) { index += y * stride; stride *= (resolution + 1); }
    if (stride <= hashmap_size) { index += z * stride; stride *= (resolution + 1); }
    if (gridtype == 0 && stride > hashmap_size) index = hash3(x, y, z);
    return fast_mod(index, hashmap_size);
}
__device__ __forceinline__ uint32_t grid_index2(uint32_t gridtype, uint32_t hashmap_size,
                                                uint32_t resolution, uint32_t x, uint32_t y) {
    uint32_t stride = 1, index = 0;
    if (stride <= hashmap_size) { index += x * stride; stride *= (resolution + 1); }
    if (stride <= hashmap_size) { index += y * stride; stride *= (resolution + 1); }
    if (gridtype == 0 && stride > hashmap_size) index = (x * 1u) ^ (y * 2654435761u);
    return fast_mod(index, hashmap_size);
}

struct LevelGeom {
    float scale;
    uint32_t resolution;
};
__device__ __forceinline__ LevelGeom level_geom(uint32_t level, float S, uint32_t H) {
    LevelGeom g;
    g.scale = __fadd_rn(__fmul_rn(exp2f(__fmul_rn((float)level, S)), (float)H), -1.0f);
    g.resolution = (uint32_t)ceilf(g.scale) + 1u;
    return g;
}

template <typename T> struct VecC;
template <int C> struct ChanVec { float v[C]; };

template <int C>
__device__ __forceinline__ void load_entry(const float* __restrict__ p, float (&v)[C]) {
    if constexpr (C == 1) {
        v[0] = __ldg(p);
    } else if constexpr (C == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
        #pragma unroll
        for (int c = 0; c < C; c += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p + c));
            v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
        }
    }
}

// Trilinear interpolation of one level at one point (D = 3).  x in [0,1]^3 (caller handles OOB).
template <int C>
__device__ __forceinline__ void interp_level3(const float* __restrict__ table, uint32_t hashmap_size,
                                              uint32_t gridtype, const LevelGeom& g, float x0, float x1,
                                              float x2, float (&res)[C]) {
    float p[3];
    uint32_t pg[3];
    const float in[3] = {x0, x1, x2};
    #pragma unroll
    for (int d = 0; d < 3; ++d) {
        p[d] = __fmaf_rn(in[d], g.scale, 0.5f);
        const float fl = floorf(p[d]);
        pg[d] = (uint32_t)fl;
        p[d] = __fadd_rn(p[d], -(float)pg[d]);
    }
    uint32_t idx[8];
    float w[8];
    #pragma unroll
    for (int c = 0; c < 8; ++c) {
        float ww = 1.0f;
        uint32_t q[3];
        #pragma unroll
        for (int d = 0; d < 3; ++d) {
            if ((c & (1 << d)) == 0) { ww = __fmul_rn(ww, __fadd_rn(1.0f, -p[d])); q[d] = pg[d]; }
            else { ww = __fmul_rn(ww, p[d]); q[d] = pg[d] + 1; }
        }
        w[c] = ww;
        idx[c] = grid_index3(gridtype, hashmap_size, g.resolution, q[0], q[1], q[2]);
    }
    float e[8][C];
    if constexpr (C == 2) {
        // The two x-neighbours of a corner pair are very often the two halves of one aligned 16-byte pair of entries
        // (hashed levels: x even -> h(x+1) = h(x) ^ 1; dense levels: consecutive indices), which is one 32-byte
        // sector either way: fetch the aligned pair around corner 0 with ONE 16-byte load and skip the second
        // gather when corner 1 is its other half (25 % fewer L1/L2 requests on average, same values).
        #pragma unroll
        for (int c = 0; c < 8; c += 2) {
            const uint32_t i0 = idx[c], i1 = idx[c + 1];
            const bool merged = (i0 ^ i1) == 1u;
            const float4 a = __ldg(reinterpret_cast<const float4*>(table + (size_t)(i0 & ~1u) * 2));
            float2 b = make_float2(0.f, 0.f);
            if (!merged) b = __ldg(reinterpret_cast<const float2*>(table + (size_t)i1 * 2));
            const bool hi = (i0 & 1u) != 0u;
            e[c][0] = hi ? a.z : a.x; e[c][1] = hi ? a.w : a.y;
            e[c + 1][0] = merged ? (hi ? a.x : a.z) : b.x;
            e[c + 1][1] = merged ? (hi ? a.y : a.w) : b.y;
        }
    } else {
        #pragma unroll
        for (int c = 0; c < 8; ++c) load_entry<C>(table + (size_t)idx[c] * C, e[c]);
    }
    #pragma unroll
    for (int ch = 0; ch < C; ++ch) res[ch] = 0.f;
    #pragma unroll
    for (int c = 0; c < 8; ++c) {
        #pragma unroll
        for (int ch = 0; ch < C; ++ch) res[ch] = __fmaf_rn(w[c], e[c][ch], res[ch]);
    }
}

// ---------------------------------------------------------------- reference-layout forward
// outputs [L,B,C]; dy_dx [B,L,D,C] when calc_grad_inputs.  grid = (ceil(B/256), L).
template <int D, int C>
__global__ void __launch_bounds__(256) k_grid_fwd(const float* __restrict__ inputs,
                                                  const float* __restrict__ table,
                                                  const int* __restrict__ offsets, float* __restrict__ outputs,
                                                  uint32_t B, uint32_t L, float S, uint32_t H,
                                                  bool calc_grad_inputs, float* __restrict__ dy_dx,
                                                  uint32_t gridtype, int* __restrict__ dbg_indices) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const float* tab = table + (size_t)(uint32_t)offsets[level] * C;
    float x[3] = {0.f, 0.f, 0.f};
    bool oob = false;
    #pragma unroll
    for (int d = 0; d < D; ++d) {
        x[d] = inputs[(size_t)b * D + d];
        if (x[d] < 0.f || x[d] > 1.f) oob = true;
    }
    float* out = outputs + ((size_t)level * B + b) * C;
    if (oob) {
        #pragma unroll
        for (int ch = 0; ch < C; ++ch) out[ch] = 0.f;
        if (calc_grad_inputs) {
            float* g = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;
            #pragma unroll
            for (int i = 0; i < D * C; ++i) g[i] = 0.f;
        }
        if (dbg_indices) {
            for (int c = 0; c < (1 << D); ++c) dbg_indices[((size_t)b * L + level) * (1 << D) + c] = -1;
        }
        return;
    }
    const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    const LevelGeom g = level_geom(level, S, H);
    float p[D];
    uint32_t pg[D];
    #pragma unroll
    for (int d = 0; d < D; ++d) {
        p[d] = __fmaf_rn(x[d], g.scale, 0.5f);
        pg[d] = (uint32_t)floorf(p[d]);
        p[d] = __fadd_rn(p[d], -(float)pg[d]);
    }
    float res[C];
    #pragma unroll
    for (int ch = 0; ch < C; ++ch) res[ch] = 0.f;
    #pragma unroll
    for (int c = 0; c < (1 << D); ++c) {
        float w = 1.0f;
        uint32_t q[3] = {0, 0, 0};
        #pragma unroll
        for (int d = 0; d < D; ++d) {
            if ((c & (1 << d)) == 0) { w = __fmul_rn(w, __fadd_rn(1.0f, -p[d])); q[d] = pg[d]; }
            else { w = __fmul_rn(w, p[d]); q[d] = pg[d] + 1; }
        }
        const uint32_t idx = (D == 3) ? grid_index3(gridtype, hashmap_size, g.resolution, q[0], q[1], q[2])
                                      : grid_index2(gridtype, hashmap_size, g.resolution, q[0], q[1]);
        if (dbg_indices) dbg_indices[((size_t)b * L + level) * (1 << D) + c] = (int)idx;
        float e[C];
        load_entry<C>(tab + (size_t)idx * C, e);
        #pragma unroll
        for (int ch = 0; ch < C; ++ch) res[ch] = __fmaf_rn(w, e[ch], res[ch]);
    }
    #pragma unroll
    for (int ch = 0; ch < C; ++ch) out[ch] = res[ch];

    if (calc_grad_inputs) {
        float* gout = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;
        #pragma unroll
        for (int gd = 0; gd < D; ++gd) {
            float rg[C];
            #pragma unroll
            for (int ch = 0; ch < C; ++ch) rg[ch] = 0.f;
            #pragma unroll
            for (int c = 0; c < (1 << (D - 1)); ++c) {
                float w = g.scale;
                uint32_t q[3] = {0, 0, 0};
                #pragma unroll
                for (int nd = 0; nd < D - 1; ++nd) {
                    const int d = (nd >= gd) ? (nd + 1) : nd;
                    if ((c & (1 << nd)) == 0) { w = __fmul_rn(w, __fadd_rn(1.0f, -p[d])); q[d] = pg[d]; }
                    else { w = __fmul_rn(w, p[d]); q[d] = pg[d] + 1; }
                }
                q[gd] = pg[gd];
                const uint32_t il = (D == 3) ? grid_index3(gridtype, hashmap_size, g.resolution, q[0], q[1], q[2])
                                             : grid_index2(gridtype, hashmap_size, g.resolution, q[0], q[1]);
                q[gd] = pg[gd] + 1;
                const uint32_t ir = (D == 3) ? grid_index3(gridtype, hashmap_size, g.resolution, q[0], q[1], q[2])
                                             : grid_index2(gridtype, hashmap_size, g.resolution, q[0], q[1]);
                float el[C], er[C];
                load_entry<C>(tab + (size_t)il * C, el);
                load_entry<C>(tab + (size_t)ir * C, er);
                #pragma unroll
                for (int ch = 0; ch < C; ++ch) rg[ch] = __fmaf_rn(w, __fadd_rn(er[ch], -el[ch]), rg[ch]);
            }
            #pragma unroll
            for (int ch = 0; ch < C; ++ch) gout[gd * C + ch] = rg[ch];
        }
    }
}

// ---------------------------------------------------------------- reference-layout backward
// grad [L,B,C] -> grad_table (+=).  One thread per (sample, level); the C channels of a corner go
// out as 8-byte (C=2) / 16-byte (C=4,8) vector reductions.
template <int C>
__device__ __forceinline__ void red_add(float* __restrict__ p, const float (&v)[C], float w) {
    if constexpr (C == 1) {
        atomicAdd(p, w * v[0]);
    } else if constexpr (C == 2) {
        atomicAdd(reinterpret_cast<float2*>(p), make_float2(w * v[0], w * v[1]));
    } else {
        #pragma unroll
        for (int c = 0; c < C; c += 4)
            atomicAdd(reinterpret_cast<float4*>(p + c),
                      make_float4(w * v[c], w * v[c + 1], w * v[c + 2], w * v[c + 3]));
    }
}

template <int D, int C>
__global__ void __launch_bounds__(256) k_grid_bwd(const float* __restrict__ grad, uint32_t ld_level,
                                                  const float* __restrict__ inputs,
                                                  const int* __restrict__ offsets,
                                                  float* __restrict__ grad_table, uint32_t B,
                                                  const int* __restrict__ n_dev, uint32_t L, float S,
                                                  uint32_t H, uint32_t gridtype, float in_lo, float in_scale,
                                                  bool clip_inputs, uint32_t level0) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = n_dev ? min(B, (uint32_t)*n_dev) : B;
    if (b >= n) return;
    const uint32_t level = blockIdx.y;   // relative to the (possibly shifted) grad / offsets pointers
    float x[3] = {0.f, 0.f, 0.f};
    #pragma unroll
    for (int d = 0; d < D; ++d) {
        float v = inputs[(size_t)b * D + d];
        if (in_scale != 0.f) v = __fmul_rn(__fadd_rn(v, in_lo), in_scale);  // (x + bound) / (2 bound)
        if (clip_inputs) v = fminf(fmaxf(v, 0.f), 1.f);
        x[d] = v;
        if (v < 0.f || v > 1.f) return;  // gradient of an OOB sample is zero
    }
    float gv[C];
    #pragma unroll
    for (int ch = 0; ch < C; ++ch) gv[ch] = grad[((size_t)level * ld_level + b) * C + ch];
    const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    float* tab = grad_table + (size_t)(uint32_t)offsets[level] * C;
    const LevelGeom g = level_geom(level + level0, S, H);
    float p[D];
    uint32_t pg[D];
    #pragma unroll
    for (int d = 0; d < D; ++d) {
        p[d] = __fmaf_rn(x[d], g.scale, 0.5f);
        pg[d] = (uint32_t)floorf(p[d]);
        p[d] = __fadd_rn(p[d], -(float)pg[d]);
    }
    if constexpr (D == 3 && C == 2) {
        // x-neighbour corners that are the two halves of one aligned 16-byte pair of entries (see interp_level3) go out
        // as ONE red.global.add.v4.f32 instead of two v2 reductions: the L2 atomic unit is the limit of this kernel
        // (lts 92 % busy), and this removes a quarter of its operations on average.
        #pragma unroll
        for (int c = 0; c < 8; c += 2) {
            uint32_t q[3] = {pg[0], pg[1], pg[2]};
            // same operation order as the generic loop: w = ((1 * wx) * wy) * wz
            float w0 = __fmul_rn(1.0f, __fadd_rn(1.0f, -p[0])), w1 = __fmul_rn(1.0f, p[0]);
            #pragma unroll
            for (int d = 1; d < 3; ++d) {
                const bool up = (c & (1 << d)) != 0;
                const float f = up ? p[d] : __fadd_rn(1.0f, -p[d]);
                if (up) q[d] = pg[d] + 1;
                w0 = __fmul_rn(w0, f); w1 = __fmul_rn(w1, f);
            }
            const uint32_t i0 = grid_index3(gridtype, hashmap_size, g.resolution, q[0], q[1], q[2]);
            const uint32_t i1 = grid_index3(gridtype, hashmap_size, g.resolution, q[0] + 1, q[1], q[2]);
            if ((i0 ^ i1) == 1u) {
                const bool hi = (i0 & 1u) != 0u;
                const float ax = w0 * gv[0], ay = w0 * gv[1], bx = w1 * gv[0], by = w1 * gv[1];
                atomicAdd(reinterpret_cast<float4*>(tab + (size_t)(i0 & ~1u) * 2),
                          hi ? make_float4(bx, by, ax, ay) : make_float4(ax, ay, bx, by));
            } else {
                red_add<C>(tab + (size_t)i0 * C, gv, w0);
                red_add<C>(tab + (size_t)i1 * C, gv, w1);
            }
        }
    } else {
        #pragma unroll
        for (int c = 0; c < (1 << D); ++c) {
            float w = 1.0f;
            uint32_t q[3] = {0, 0, 0};
            #pragma unroll
            for (int d = 0; d < D; ++d) {
                if ((c & (1 << d)) == 0) { w = __fmul_rn(w, __fadd_rn(1.0f, -p[d])); q[d] = pg[d]; }
                else { w = __fmul_rn(w, p[d]); q[d] = pg[d] + 1; }
            }
            const uint32_t idx = (D == 3) ? grid_index3(gridtype, hashmap_size, g.resolution, q[0], q[1], q[2])
                                          : grid_index2(gridtype, hashmap_size, g.resolution, q[0], q[1]);
            red_add<C>(tab + (size_t)idx * C, gv, w);
        }
    }
}

// Coarse-level variant of the scatter (C = 2, D = 3).  Consecutive samples of a ray are dt apart, far
// less than a coarse cell, so neighbouring lanes mostly hit the SAME eight entries: lanes are grouped
// into runs of equal entry index (compare with the previous lane), a segmented shuffle scan sums each
// run and only its last lane issues the reduction.  Runs that are not adjacent simply issue separate
// reductions, so the result is the same sum in a different order.  The whole warp stays convergent:
// out-of-range lanes carry zero weight instead of returning.
__global__ void __launch_bounds__(256) k_grid_bwd_runs(const float* __restrict__ grad, uint32_t ld_level,
                                                       const float* __restrict__ inputs,
                                                       const int* __restrict__ offsets,
                                                       float* __restrict__ grad_table, uint32_t B,
                                                       const int* __restrict__ n_dev, uint32_t L, float S,
                                                       uint32_t H, uint32_t gridtype, float in_lo, float in_scale,
                                                       bool clip_inputs) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = n_dev ? min(B, (uint32_t)*n_dev) : B;
    const uint32_t lane = threadIdx.x & 31;
    if ((b & ~31u) >= n) return;                  // whole warp past the end
    const uint32_t level = blockIdx.y;
    bool live = b < n;
    float x[3] = {0.f, 0.f, 0.f};
    #pragma unroll
    for (int d = 0; d < 3; ++d) {
        float v = live ? inputs[(size_t)b * 3 + d] : 0.f;
        if (in_scale != 0.f) v = __fmul_rn(__fadd_rn(v, in_lo), in_scale);
        if (clip_inputs) v = fminf(fmaxf(v, 0.f), 1.f);
        x[d] = v;
        if (v < 0.f || v > 1.f) live = false;
    }
    float2 gv = make_float2(0.f, 0.f);
    if (live) gv = *reinterpret_cast<const float2*>(grad + ((size_t)level * ld_level + b) * 2);
    if (!live) { x[0] = x[1] = x[2] = 0.f; }
    const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    float* tab = grad_table + (size_t)(uint32_t)offsets[level] * 2;
    const LevelGeom g = level_geom(level, S, H);
    float p[3];
    uint32_t pg[3];
    #pragma unroll
    for (int d = 0; d < 3; ++d) {
        p[d] = __fmaf_rn(x[d], g.scale, 0.5f);
        pg[d] = (uint32_t)floorf(p[d]);
        p[d] = __fadd_rn(p[d], -(float)pg[d]);
    }
    #pragma unroll
    for (int c = 0; c < 8; ++c) {
        float w = 1.0f;
        uint32_t q[3];
        #pragma unroll
        for (int d = 0; d < 3; ++d) {
            if ((c & (1 << d)) == 0) { w = __fmul_rn(w, __fadd_rn(1.0f, -p[d])); q[d] = pg[d]; }
            else { w = __fmul_rn(w, p[d]); q[d] = pg[d] + 1; }
        }
        const uint32_t idx = grid_index3(gridtype, hashmap_size, g.resolution, q[0], q[1], q[2]);
        float vx = w * gv.x, vy = w * gv.y;
        const uint32_t prev = __shfl_up_sync(0xffffffffu, idx, 1);
        const bool head = (lane == 0) || (prev != idx);
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const uint32_t below = heads & (0xffffffffu >> (31 - lane));   // heads at or below my lane
        const int run_start = 31 - __clz(below);
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float ux = __shfl_up_sync(0xffffffffu, vx, o);
            const float uy = __shfl_up_sync(0xffffffffu, vy, o);
            if ((int)lane - o >= run_start) { vx += ux; vy += uy; }
        }
        const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
        if (tail && (vx != 0.f || vy != 0.f))
            atomicAdd(reinterpret_cast<float2*>(tab + (size_t)idx * 2), make_float2(vx, vy));
    }
}

// gridencoder.cu:315-341
template <int D, int C>
__global__ void k_grid_input_bwd(const float* __restrict__ grad, const float* __restrict__ dy_dx,
                                 float* __restrict__ grad_inputs, uint32_t B, uint32_t L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* dd = dy_dx + (size_t)b * L * D * C;
    float r = 0.f;
    for (uint32_t l = 0; l < L; ++l)
        #pragma unroll
        for (int ch = 0; ch < C; ++ch)
            r = __fmaf_rn(grad[((size_t)l * B + b) * C + ch], dd[l * D * C + d * C + ch], r);
    grad_inputs[t] = r;
}

// ---------------------------------------------------------------- fused position encoder
// One thread per sample builds the whole network input row in fp16:
//   mode 0 'freq'   : Frequency(10) of (x+b)/2b                      -> 60 (+4 ones)  = 64
//   mode 1 'hg'     : grid((x+b)/2b), OOB -> zeros                   -> 2L (+ones to a multiple of 16)
//   mode 2 'hg+freq': Frequency(2) of raw x (12) ++ grid(clip((x+b)/2b,0,1)) (2L) (+ones)
// (autolabel/models.py:15-59,138-148).  Padding columns are 1.0 (bias column of the bias-free MLP).
__device__ __forceinline__ void freq_encode3(const float* x, int n_freq, __half* out) {
    #pragma unroll
    for (int d = 0; d < 3; ++d) {
        for (int k = 0; k < n_freq; ++k) {
            float s, c;
            sincospif(scalbnf(x[d], k), &s, &c);
            out[(d * n_freq + k) * 2] = __float2half_rn(s);
            out[(d * n_freq + k) * 2 + 1] = __float2half_rn(c);
        }
    }
}

template <int LMAX>
__global__ void __launch_bounds__(128) k_encode_position(const float* __restrict__ xyz, uint32_t cap,
                                                         const int* __restrict__ n_dev, float bound,
                                                         int mode, const float* __restrict__ table,
                                                         const int* __restrict__ offsets, uint32_t L, float S,
                                                         uint32_t H, uint32_t gridtype,
                                                         __half* __restrict__ out, uint32_t ldo) {
    // Rows are assembled in shared memory (odd word stride: a thread writing its own row is bank-conflict free) and
    // leave the block as ONE contiguous, fully coalesced run of 16-byte streaming stores: a thread-per-row store
    // pattern costs 32 L1 wavefronts per 4-byte store instruction (24 of them per row), a third of this kernel's L1
    // load next to its gathers, and the rows are read exactly once, by the next kernel: evict-first keeps them from
    // displacing the hash table in L2.
    __shared__ uint32_t srow[128 * 33];
    const uint32_t b0 = blockIdx.x * blockDim.x;
    const uint32_t b = b0 + threadIdx.x;
    const uint32_t n = n_dev ? min(cap, (uint32_t)*n_dev) : cap;
    if (b0 >= n) return;
    const uint32_t W = ldo >> 1, stride = W | 1;                 // words per row; ldo is a multiple of 16 halfs, <= 64
    if (b < n) {
    const float p[3] = {xyz[(size_t)b * 3], xyz[(size_t)b * 3 + 1], xyz[(size_t)b * 3 + 2]};
    __half* o = reinterpret_cast<__half*>(srow + threadIdx.x * stride);
    uint32_t col = 0;
    const float inv2b = __fdiv_rn(1.0f, __fmul_rn(2.0f, bound));
    float xn[3];
    #pragma unroll
    for (int d = 0; d < 3; ++d) xn[d] = __fmul_rn(__fadd_rn(p[d], bound), inv2b);
    if (mode == 0) {
        __half tmp[60];
        freq_encode3(xn, 10, tmp);
        for (int i = 0; i < 60; ++i) o[i] = tmp[i];
        col = 60;
    } else {
        if (mode == 2) {
            __half tmp[12];
            freq_encode3(p, 2, tmp);
            #pragma unroll
            for (int i = 0; i < 12; ++i) o[i] = tmp[i];
            col = 12;
            #pragma unroll
            for (int d = 0; d < 3; ++d) xn[d] = fminf(fmaxf(xn[d], 0.f), 1.f);
        }
        const bool oob = xn[0] < 0.f || xn[0] > 1.f || xn[1] < 0.f || xn[1] > 1.f || xn[2] < 0.f || xn[2] > 1.f;
        for (uint32_t l = 0; l < L; ++l) {
            float r[2] = {0.f, 0.f};
            if (!oob) {
                const uint32_t hs = (uint32_t)(offsets[l + 1] - offsets[l]);
                const LevelGeom g = level_geom(l, S, H);
                interp_level3<2>(table + (size_t)(uint32_t)offsets[l] * 2, hs, gridtype, g, xn[0], xn[1], xn[2], r);
            }
            *reinterpret_cast<__half2*>(o + col + 2 * l) = __floats2half2_rn(r[0], r[1]);
        }
        col += 2 * L;
    }
    for (; col < ldo; ++col) o[col] = __float2half_rn(1.0f);
    }
    __syncthreads();
    const uint32_t rows = min(128u, n - b0);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)b0 * ldo);
    for (uint32_t i = threadIdx.x; i < rows * (W >> 2); i += 128) {
        const uint32_t r = (i << 2) / W, w = (i << 2) - r * W;   // W is a multiple of 4: a chunk never straddles rows
        const uint32_t* sp = srow + r * stride + w;
        __stcs(dst + i, make_uint4(sp[0], sp[1], sp[2], sp[3]));
    }
}

// ---------------------------------------------------------------- standalone small encodings
// tcnn Frequency: out [B, D*2*n_freq] fp32.
__global__ void k_freq_encode(const float* __restrict__ x, uint32_t B, uint32_t D, uint32_t n_freq,
                              float* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t per = D * n_freq;
    if (t >= B * per) return;
    const uint32_t b = t / per, r = t - b * per, d = r / n_freq, k = r - d * n_freq;
    float s, c;
    sincospif(scalbnf(x[(size_t)b * D + d], (int)k), &s, &c);
    out[(size_t)b * per * 2 + r * 2] = s;
    out[(size_t)b * per * 2 + r * 2 + 1] = c;
}

__global__ void k_sh_encode(const float* __restrict__ d01, uint32_t B, float* __restrict__ out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float o[16];
    al_sh4(d01[(size_t)b * 3] * 2.f - 1.f, d01[(size_t)b * 3 + 1] * 2.f - 1.f, d01[(size_t)b * 3 + 2] * 2.f - 1.f, o);
    #pragma unroll
    for (int i = 0; i < 16; ++i) out[(size_t)b * 16 + i] = o[i];
}

// Head inputs of the fused field path, built from the density MLP output h16 = [h0, geo(15)]:
//   color_in [n,32] = [SH4(dir) (16), geo (15), 1]     (models.py:205-209; dir of the sample's ray)
//   semf_in  [n,16] = [geo (15), 1]                    (models.py:253)
//   semo_in  [n, F+16] columns F.. = [geo (15), 1]     (models.py:254-255; columns 0..F-1 are
//                                                       written by the feature MLP's epilogue)
// dirs_mode 0: dirs [n,3] per sample; 1: rays_d [N,3] indexed through sray [n].
__global__ void __launch_bounds__(256) k_head_inputs(const float* __restrict__ h16, uint32_t cap,
                                                     const int* __restrict__ n_dev,
                                                     const float* __restrict__ dirs,
                                                     const int* __restrict__ sray, __half* __restrict__ color_in,
                                                     __half* __restrict__ semf_in, __half* __restrict__ semo_in,
                                                     uint32_t ld_semo, uint32_t F) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = n_dev ? min(cap, (uint32_t)*n_dev) : cap;
    if (b >= n) return;
    float h[16];
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = reinterpret_cast<const float4*>(h16 + (size_t)b * 16)[i];
        h[4 * i] = t.x; h[4 * i + 1] = t.y; h[4 * i + 2] = t.z; h[4 * i + 3] = t.w;
    }
    __align__(16) __half geo[16];
    #pragma unroll
    for (int i = 0; i < 15; ++i) geo[i] = __float2half_rn(h[i + 1]);
    geo[15] = __float2half_rn(1.0f);
    if (color_in) {
        const size_t r = sray ? (size_t)sray[b] : (size_t)b;
        // (d + 1) / 2 then 2 x - 1, as the reference + tcnn do
        const float dx = ((dirs[r * 3] + 1.f) * 0.5f) * 2.f - 1.f;
        const float dy = ((dirs[r * 3 + 1] + 1.f) * 0.5f) * 2.f - 1.f;
        const float dz = ((dirs[r * 3 + 2] + 1.f) * 0.5f) * 2.f - 1.f;
        float s[16];
        al_sh4(dx, dy, dz, s);
        __align__(16) __half sh[16];
        #pragma unroll
        for (int i = 0; i < 16; ++i) sh[i] = __float2half_rn(s[i]);
        uint4* dst = reinterpret_cast<uint4*>(color_in + (size_t)b * 32);
        dst[0] = reinterpret_cast<const uint4*>(sh)[0];
        dst[1] = reinterpret_cast<const uint4*>(sh)[1];
        dst[2] = reinterpret_cast<const uint4*>(geo)[0];
        dst[3] = reinterpret_cast<const uint4*>(geo)[1];
    }
    if (semf_in) {
        uint4* dst = reinterpret_cast<uint4*>(semf_in + (size_t)b * 16);
        dst[0] = reinterpret_cast<const uint4*>(geo)[0];
        dst[1] = reinterpret_cast<const uint4*>(geo)[1];
    }
    if (semo_in) {
        uint4* dst = reinterpret_cast<uint4*>(semo_in + (size_t)b * ld_semo + F);
        dst[0] = reinterpret_cast<const uint4*>(geo)[0];
        dst[1] = reinterpret_cast<const uint4*>(geo)[1];
    }
}

}  // namespace

// ================================================================ C ABI
#define AL_GRID_DISPATCH(D, C, CALL)                                                         \
    do {                                                                                     \
        if (D == 3 && C == 1) { constexpr int DD = 3, CC = 1; CALL; }                        \
        else if (D == 3 && C == 2) { constexpr int DD = 3, CC = 2; CALL; }                   \
        else if (D == 3 && C == 4) { constexpr int DD = 3, CC = 4; CALL; }                   \
        else if (D == 3 && C == 8) { constexpr int DD = 3, CC = 8; CALL; }                   \
        else if (D == 2 && C == 1) { constexpr int DD = 2, CC = 1; CALL; }                   \
        else if (D == 2 && C == 2) { constexpr int DD = 2, CC = 2; CALL; }                   \
        else if (D == 2 && C == 4) { constexpr int DD = 2, CC = 4; CALL; }                   \
        else if (D == 2 && C == 8) { constexpr int DD = 2, CC = 8; CALL; }                   \
        else { al_set_error("GridEncoding: D must be 2 or 3 and C one of 1, 2, 4, 8");       \
               return (int)cudaErrorInvalidValue; }                                          \
    } while (0)

// grid_encode_forward (gridencoder.h:11).  fp32 tables.  dbg_indices (optional, int [B,L,2^D])
// receives the table entry index of every corner (-1 for out-of-range samples): parity probe.
AL_API int al_grid_encode_forward(const float* inputs, const float* embeddings, const int* offsets,
                                  float* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                  uint32_t H, int calc_grad_inputs, float* dy_dx, uint32_t gridtype,
                                  int* dbg_indices, void* stream) {
    if (B == 0) return 0;
    AL_REQUIRE(inputs && embeddings && offsets && outputs, "null pointer");
    AL_REQUIRE(!calc_grad_inputs || dy_dx, "dy_dx required when calc_grad_inputs");
    AL_REQUIRE(L >= 1 && L <= 65535, "bad level count");
    const dim3 grid(al_div_up(B, 256), L, 1);
    AL_GRID_DISPATCH(D, C, (k_grid_fwd<DD, CC><<<grid, 256, 0, (cudaStream_t)stream>>>(
                               inputs, embeddings, offsets, outputs, B, L, S, H, calc_grad_inputs != 0, dy_dx,
                               gridtype, dbg_indices)));
    AL_LAUNCH_CHECK();
    return 0;
}

// grid_encode_backward (gridencoder.h:12).  grad_embeddings is accumulated into (+=): the caller
// zero-fills it (grid.py:72) or passes its persistent gradient buffer.
AL_API int al_grid_encode_backward(const float* grad, const float* inputs, const int* offsets,
                                   float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                                   float S, uint32_t H, int calc_grad_inputs, const float* dy_dx,
                                   float* grad_inputs, uint32_t gridtype, void* stream) {
    if (B == 0) return 0;
    AL_REQUIRE(grad && inputs && offsets && grad_embeddings, "null pointer");
    const dim3 grid(al_div_up(B, 256), L, 1);
    AL_GRID_DISPATCH(D, C, (k_grid_bwd<DD, CC><<<grid, 256, 0, (cudaStream_t)stream>>>(
                               grad, B, inputs, offsets, grad_embeddings, B, nullptr, L, S, H, gridtype, 0.f, 0.f,
                               false, 0u)));
    AL_LAUNCH_CHECK();
    if (calc_grad_inputs) {
        AL_REQUIRE(dy_dx && grad_inputs, "dy_dx / grad_inputs required");
        AL_GRID_DISPATCH(D, C, (k_grid_input_bwd<DD, CC><<<al_div_up((unsigned long long)B * D, 256), 256, 0,
                                                           (cudaStream_t)stream>>>(grad, dy_dx, grad_inputs, B, L)));
        AL_LAUNCH_CHECK();
    }
    return 0;
}

// Fused-path scatter: grad is level-major [L, ld_level, 2] (written by the density-MLP backward),
// positions are raw xyz [cap,3] in [-bound,bound]; normalisation (and the hg+freq clip) is redone
// here exactly as in the encoder.  n_dev (optional, device int) bounds the live samples.
AL_API int al_grid_scatter_xyz(const float* grad, uint32_t ld_level, const float* xyz, uint32_t cap,
                               const int* n_dev, float bound, int clip, const int* offsets,
                               float* grad_embeddings, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                               void* stream) {
    if (cap == 0) return 0;
    AL_REQUIRE(grad && xyz && offsets && grad_embeddings, "null pointer");
    const float inv2b = 1.0f / (2.0f * bound);
    // Levels whose cells are wider than a few marching steps (2 bound / (H 2^(l S)) >> dt_min) go through the
    // run-aggregating kernel; finer levels scatter directly.  The split only changes the summation order.
    uint32_t n_coarse = 0;
    for (uint32_t l = 0; l < L; ++l) {
        const float cell = 2.0f * bound / ((float)H * exp2f((float)l * S));
        if (cell > 2.0f * 0.00338f) n_coarse = l + 1;
    }
    if (n_coarse > 0) {
        const dim3 grid(al_div_up(cap, 256), n_coarse, 1);
        k_grid_bwd_runs<<<grid, 256, 0, (cudaStream_t)stream>>>(grad, ld_level, xyz, offsets, grad_embeddings, cap, n_dev,
                                                               L, S, H, gridtype, bound, inv2b, clip != 0);
        AL_LAUNCH_CHECK();
    }
    if (n_coarse < L) {
        const dim3 grid(al_div_up(cap, 256), L - n_coarse, 1);
        k_grid_bwd<3, 2><<<grid, 256, 0, (cudaStream_t)stream>>>(grad + (size_t)n_coarse * ld_level * 2, ld_level, xyz,
                                                                  offsets + n_coarse, grad_embeddings, cap, n_dev,
                                                                  L - n_coarse, S, H, gridtype, bound, inv2b, clip != 0,
                                                                  n_coarse);
        AL_LAUNCH_CHECK();
    }
    return 0;
}

// Position encoder of the fused path (fp16 rows for the tensor-core MLP).  mode: 0 freq, 1 hg,
// 2 hg+freq.  ldo must be a multiple of 16 and >= the encoded width.
AL_API int al_encode_position(const float* xyz, uint32_t cap, const int* n_dev, float bound, int mode,
                              const float* table, const int* offsets, uint32_t L, float S, uint32_t H,
                              uint32_t gridtype, void* out_half, uint32_t ldo, void* stream) {
    if (cap == 0) return 0;
    AL_REQUIRE(xyz && out_half, "null pointer");
    AL_REQUIRE(mode == 0 || (table && offsets), "grid modes need a table");
    const uint32_t width = mode == 0 ? 60 : (mode == 1 ? 2 * L : 12 + 2 * L);
    AL_REQUIRE(ldo >= width && ldo % 8 == 0 && ldo <= 64, "ldo too small / unaligned / above 64 halfs");
    AL_REQUIRE(((uintptr_t)out_half & 15) == 0, "out must be 16-byte aligned");
    k_encode_position<16><<<al_div_up(cap, 128), 128, 0, (cudaStream_t)stream>>>(
        xyz, cap, n_dev, bound, mode, table, offsets, L, S, H, gridtype, (__half*)out_half, ldo);
    AL_LAUNCH_CHECK();
    return 0;
}

AL_API int al_freq_encode(const float* x, uint32_t B, uint32_t D, uint32_t n_freq, float* out, void* stream) {
    if (B == 0) return 0;
    AL_REQUIRE(x && out, "null pointer");
    k_freq_encode<<<al_div_up((unsigned long long)B * D * n_freq, 256), 256, 0, (cudaStream_t)stream>>>(x, B, D, n_freq, out);
    AL_LAUNCH_CHECK();
    return 0;
}

AL_API int al_sh_encode(const float* d01, uint32_t B, float* out, void* stream) {
    if (B == 0) return 0;
    AL_REQUIRE(d01 && out, "null pointer");
    k_sh_encode<<<al_div_up(B, 256), 256, 0, (cudaStream_t)stream>>>(d01, B, out);
    AL_LAUNCH_CHECK();
    return 0;
}

AL_API int al_head_inputs(const float* h16, uint32_t cap, const int* n_dev, const float* dirs, const int* sray,
                          void* color_in, void* semf_in, void* semo_in, uint32_t ld_semo, uint32_t F,
                          void* stream) {
    if (cap == 0) return 0;
    AL_REQUIRE(h16 && (!color_in || dirs), "null pointer");
    AL_REQUIRE(!semo_in || (F % 8 == 0 && ld_semo % 8 == 0), "semo_in layout must be 16-byte aligned");
    k_head_inputs<<<al_div_up(cap, 256), 256, 0, (cudaStream_t)stream>>>(
        h16, cap, n_dev, dirs, sray, (__half*)color_in, (__half*)semf_in, (__half*)semo_in, ld_semo, F);
    AL_LAUNCH_CHECK();
    return 0;
}
