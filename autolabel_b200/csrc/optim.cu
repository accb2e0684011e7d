// Occupancy-grid maintenance and the optimiser step.
//
//   al_density_grid_update : the EMA-max + mean part of NeRFRenderer.update_extra_state
//                            (torch_ngp/nerf/renderer.py:662-667)
//   al_adam_step           : torch.optim.Adam as configured by scripts/train.py:50-63, fused with
//                            the gradient unscale of torch.cuda.amp.GradScaler and with zeroing the
//                            gradient buffer (the largest unavoidable HBM term of a step: 32 B/param)
#include "common.cuh"
#include "../../include/autolabel_b200.h"

namespace {

__global__ void __launch_bounds__(256) k_density_update(float* __restrict__ grid, const float* __restrict__ tmp,
                                                        uint32_t n, float decay, float inv_n,
                                                        float* __restrict__ mean_out) {
    float s = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g = grid[i];
        const float t = tmp[i];
        if (g >= 0.f && t >= 0.f) {
            g = fmaxf(g * decay, t);
            grid[i] = g;
        }
        s += fmaxf(g, 0.f);
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ float ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += ws[w];
        atomicAdd(mean_out, t * inv_n);
    }
}

// p, g, m, v as float4 streams; tail handled by the scalar path of the last thread block.
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, size_t n, float lr, float b1, float b2,
                                              float eps, float wd, float bc1, float bc2_sqrt, float gscale,
                                              int zero_grad) {
    const size_t n4 = n / 4;
    const float step_size = lr / bc1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 P = reinterpret_cast<float4*>(p)[i];
        float4 G = reinterpret_cast<float4*>(g)[i];
        float4 M = reinterpret_cast<float4*>(m)[i];
        float4 V = reinterpret_cast<float4*>(v)[i];
        float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gr = gg[k] * gscale;
            if (wd != 0.f) gr = fmaf(wd, pp[k], gr);
            mm[k] = fmaf(b1, mm[k], (1.f - b1) * gr);           // lerp form of torch: m + (g - m)(1 - b1)
            vv[k] = fmaf(b2, vv[k], (1.f - b2) * gr * gr);
            const float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
            pp[k] = pp[k] - step_size * (mm[k] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
        if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = n4 * 4 + threadIdx.x;
        float gr = g[i] * gscale;
        if (wd != 0.f) gr = fmaf(wd, p[i], gr);
        const float mk = fmaf(b1, m[i], (1.f - b1) * gr);
        const float vk = fmaf(b2, v[i], (1.f - b2) * gr * gr);
        m[i] = mk; v[i] = vk;
        p[i] = p[i] - step_size * (mk / (sqrtf(vk) / bc2_sqrt + eps));
        if (zero_grad) g[i] = 0.f;
    }
}


// The same update for up to AL_ADAM_MAX_TENSORS tensors in ONE launch, with the step count and the learning rate read
// from device memory: the launch has no host-side state, so it can sit inside the CUDA graph of a training step
// (k_adam_tick, enqueued right before, advances the counter).  Bias corrections are evaluated in double, as torch does
// on the host (torch/optim/adam.py: 1 - beta ** step).
struct AdamMulti {
    al_adam_tensor_t t[AL_ADAM_MAX_TENSORS];
    int count;
};
__global__ void k_adam_tick(int* step) { *step += 1; }

__global__ void __launch_bounds__(256) k_adam_multi(const AdamMulti a, const float* __restrict__ lr_dev,
                                                    const int* __restrict__ step_dev, double b1d, double b2d, float eps,
                                                    float gscale, int zero_grad) {
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) {
        const double step = (double)*step_dev;
        s_bc[0] = (float)(1.0 - pow(b1d, step));
        s_bc[1] = (float)sqrt(1.0 - pow(b2d, step));
    }
    __syncthreads();
    const float b1 = (float)b1d, b2 = (float)b2d;
    const float step_size = *lr_dev / s_bc[0], bc2_sqrt = s_bc[1];
    for (int ti = 0; ti < a.count; ++ti) {
        float* __restrict__ p = a.t[ti].param; float* __restrict__ g = a.t[ti].grad;
        float* __restrict__ m = a.t[ti].exp_avg; float* __restrict__ v = a.t[ti].exp_avg_sq;
        const size_t n = a.t[ti].n, n4 = n / 4;
        const float wd = a.t[ti].weight_decay;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            float4 P = reinterpret_cast<float4*>(p)[i];
            float4 G = reinterpret_cast<float4*>(g)[i];
            float4 M = reinterpret_cast<float4*>(m)[i];
            float4 V = reinterpret_cast<float4*>(v)[i];
            float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                float gr = gg[k] * gscale;
                if (wd != 0.f) gr = fmaf(wd, pp[k], gr);
                mm[k] = fmaf(b1, mm[k], (1.f - b1) * gr);
                vv[k] = fmaf(b2, vv[k], (1.f - b2) * gr * gr);
                const float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
                pp[k] = pp[k] - step_size * (mm[k] / denom);
            }
            reinterpret_cast<float4*>(p)[i] = P;
            reinterpret_cast<float4*>(m)[i] = M;
            reinterpret_cast<float4*>(v)[i] = V;
            if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
            const size_t i = n4 * 4 + threadIdx.x;
            float gr = g[i] * gscale;
            if (wd != 0.f) gr = fmaf(wd, p[i], gr);
            const float mk = fmaf(b1, m[i], (1.f - b1) * gr);
            const float vk = fmaf(b2, v[i], (1.f - b2) * gr * gr);
            m[i] = mk; v[i] = vk;
            p[i] = p[i] - step_size * (mk / (sqrtf(vk) / bc2_sqrt + eps));
            if (zero_grad) g[i] = 0.f;
        }
    }
}


__device__ __forceinline__ uint32_t part1by2(uint32_t x) {
    x &= 0x000003ff;
    x = (x ^ (x << 16)) & 0xff0000ff;
    x = (x ^ (x << 8)) & 0x0300f00f;
    x = (x ^ (x << 4)) & 0x030c30c3;
    x = (x ^ (x << 2)) & 0x09249249;
    return x;
}

// NeRFRenderer.mark_untrained_grid (torch_ngp/nerf/renderer.py:479-561): one thread per (cascade, cell) instead of
// the reference's five nested Python loops; the cell centre is tested against every camera frustum
//   cam = (p - t) R,  seen if  z > 0  and  |x| < cx/fx z + 2 half  and  |y| < cy/fy z + 2 half
// and cells no camera sees get density -1 (in Morton order, like the grid itself).
__global__ void __launch_bounds__(256) k_mark_untrained(float* __restrict__ grid, const float* __restrict__ poses,
                                                        uint32_t n_poses, float fx, float fy, float cx, float cy,
                                                        float bound, uint32_t C, uint32_t H) {
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cas = blockIdx.y;
    if (cell >= H * H * H || cas >= C) return;
    const uint32_t x = cell / (H * H), y = (cell / H) % H, z = cell % H;
    const float cb = fminf((float)(1u << cas), bound);
    const float half = cb / (float)H;
    const float s = cb - half;
    const float hm1 = (float)(H - 1);
    const float px = (2.0f * (float)x / hm1 - 1.0f) * s;
    const float py = (2.0f * (float)y / hm1 - 1.0f) * s;
    const float pz = (2.0f * (float)z / hm1 - 1.0f) * s;
    const float kx = cx / fx, ky = cy / fy, margin = half * 2.0f;
    bool seen = false;
    for (uint32_t b = 0; b < n_poses && !seen; ++b) {
        const float* P = poses + (size_t)b * 16;                   // row-major 4x4 camera-to-world
        const float dx = px - P[3], dy = py - P[7], dz = pz - P[11];
        const float camx = dx * P[0] + dy * P[4] + dz * P[8];
        const float camy = dx * P[1] + dy * P[5] + dz * P[9];
        const float camz = dx * P[2] + dy * P[6] + dz * P[10];
        seen = camz > 0.f && fabsf(camx) < kx * camz + margin && fabsf(camy) < ky * camz + margin;
    }
    if (!seen) {
        const uint32_t m = part1by2(x) | (part1by2(y) << 1) | (part1by2(z) << 2);
        grid[(size_t)cas * H * H * H + m] = -1.0f;
    }
}

}  // namespace

AL_API int al_density_grid_update(float* grid, const float* tmp_grid, uint32_t n_cells, float decay,
                                  float* mean_out, void* stream) {
    if (n_cells == 0) return 0;
    AL_REQUIRE(grid && tmp_grid && mean_out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    AL_CHECK(cudaMemsetAsync(mean_out, 0, sizeof(float), st));
    const unsigned blocks = min(al_div_up(n_cells, 256), (unsigned)al_num_sms() * 8);
    k_density_update<<<blocks, 256, 0, st>>>(grid, tmp_grid, n_cells, decay, 1.0f / (float)n_cells, mean_out);
    AL_LAUNCH_CHECK();
    return 0;
}

AL_API int al_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                        int zero_grad, void* stream) {
    if (n == 0) return 0;
    AL_REQUIRE(param && grad && exp_avg && exp_avg_sq, "null pointer");
    AL_REQUIRE(step >= 1, "step must be >= 1");
    AL_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
               "buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const unsigned blocks = min(al_div_up(n / 4 + 1, 256), (unsigned)al_num_sms() * 16);
    k_adam<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                     weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale, zero_grad);
    AL_LAUNCH_CHECK();
    return 0;
}

AL_API int al_adam_multi(const al_adam_tensor_t* tensors, int count, const float* lr_dev, int* step_dev, double beta1,
                         double beta2, float eps, float grad_scale, int zero_grad, void* stream) {
    if (count == 0) return 0;
    AL_REQUIRE(tensors && lr_dev && step_dev, "null pointer");
    AL_REQUIRE(count > 0 && count <= AL_ADAM_MAX_TENSORS, "1..AL_ADAM_MAX_TENSORS tensors per launch");
    AdamMulti a;
    a.count = count;
    size_t total = 0;
    for (int i = 0; i < count; ++i) {
        a.t[i] = tensors[i];
        AL_REQUIRE(a.t[i].param && a.t[i].grad && a.t[i].exp_avg && a.t[i].exp_avg_sq, "null tensor pointer");
        AL_REQUIRE((((uintptr_t)a.t[i].param | (uintptr_t)a.t[i].grad | (uintptr_t)a.t[i].exp_avg |
                     (uintptr_t)a.t[i].exp_avg_sq) & 15) == 0, "buffers must be 16-byte aligned");
        total += a.t[i].n;
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_adam_tick<<<1, 1, 0, st>>>(step_dev);
    AL_LAUNCH_CHECK();
    const unsigned blocks = min(al_div_up(total / 4 + 1, 256), (unsigned)al_num_sms() * 16);
    k_adam_multi<<<blocks, 256, 0, st>>>(a, lr_dev, step_dev, beta1, beta2, eps, grad_scale, zero_grad);
    AL_LAUNCH_CHECK();
    return 0;
}

// mark_untrained_grid (renderer.py:479-561).  poses: device fp32 [n_poses, 4, 4] camera-to-world (row-major);
// grid: [C, H^3] in Morton order; cells outside every frustum are set to -1, the others are left untouched.
AL_API int al_mark_untrained_grid(float* grid, const float* poses, uint32_t n_poses, float fx, float fy, float cx,
                                  float cy, float bound, uint32_t C, uint32_t H, void* stream) {
    AL_REQUIRE(grid && poses && n_poses > 0, "null pointer");
    AL_REQUIRE(C >= 1 && C <= 8 && H >= 8 && H <= 1024, "bad grid parameters");
    const dim3 g(al_div_up((unsigned long long)H * H * H, 256), C);
    k_mark_untrained<<<g, 256, 0, (cudaStream_t)stream>>>(grid, poses, n_poses, fx, fy, cx, cy, bound, C, H);
    AL_LAUNCH_CHECK();
    return 0;
}
