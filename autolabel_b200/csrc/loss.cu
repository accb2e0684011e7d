// Training losses of autolabel's SimpleTrainer.train_step (autolabel/trainer.py:54-94) and their gradients with
// respect to the compositing outputs, fused into two launches with no host synchronisation:
//
//   image = out[:, 0:3] + (1 - weights_sum) * 1          white background (torch_ngp/nerf/renderer.py:294-297)
//   depth = depth_raw / direction_norm                   metric depth      (renderer.py:273-275)
//   loss  = rgb_w  * mean((image - gt_rgb)^2)                                       (trainer.py:72-73, MSELoss)
//         + depth_w * sum(|depth - gt_depth| [gt_depth > eps]) / max(count, 1)      (trainer.py:76-80)
//         + feat_w * mean(|features[:, :Fg] - gt_features|)                         (trainer.py:82-85, l1_loss)
//         + sem_w  * sum(CE(logits, label) [label >= 0]) / max(count, 1)            (trainer.py:87-91)
//
// The reference evaluates the masked means with boolean-mask indexing (two device syncs per step) and lets
// autograd walk ~40 small kernels; here one warp owns a ray, lanes own channels.
#include "common.cuh"

namespace {

__device__ __forceinline__ float wsum(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(256) k_loss_counts(const float* __restrict__ gt_depth, const long long* __restrict__ gt_sem,
                                                     uint32_t N, float depth_eps, int* __restrict__ counts) {
    int cd = 0, cs = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        cd += gt_depth && gt_depth[i] > depth_eps;
        cs += gt_sem && gt_sem[i] >= 0;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cd += __shfl_xor_sync(0xffffffffu, cd, o);
        cs += __shfl_xor_sync(0xffffffffu, cs, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cd) atomicAdd(counts, cd);
        if (cs) atomicAdd(counts + 1, cs);
    }
}

template <int NC>
__global__ void __launch_bounds__(256) k_loss(const float* __restrict__ ws, const float* __restrict__ depth_raw,
                                              const float* __restrict__ out, uint32_t N, uint32_t C, uint32_t F,
                                              const float* __restrict__ norms, const float* __restrict__ gt_rgb,
                                              const float* __restrict__ gt_depth, const long long* __restrict__ gt_sem,
                                              const float* __restrict__ gt_feat, uint32_t Fg, float rgb_w, float depth_w,
                                              float sem_w, float feat_w, float depth_eps, float grad_scale,
                                              const int* __restrict__ counts, float* __restrict__ loss,
                                              float* __restrict__ g_ws, float* __restrict__ g_depth,
                                              float* __restrict__ g_out) {
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t K = 3 + C + F;
    const float inv_rgb = 1.0f / (3.0f * (float)N);
    const float inv_feat = Fg ? 1.0f / ((float)N * (float)Fg) : 0.f;
    const float inv_cd = 1.0f / (float)max(counts[0], 1);
    const float inv_cs = 1.0f / (float)max(counts[1], 1);
    float l_rgb = 0.f, l_depth = 0.f, l_feat = 0.f, l_sem = 0.f;
    for (uint32_t n = blockIdx.x * warps_per_block + (threadIdx.x >> 5); n < N; n += gridDim.x * warps_per_block) {
        const float w = ws[n];
        const float* o = out + (size_t)n * K;
        float* g = g_out + (size_t)n * K;
        float gws = 0.f;
        // semantic: log-softmax over the C logits in channels 3 .. 3 + C - 1 (lanes stride over them: any C, e.g. the
        // 606-class ScanNet label set; two warp reductions)
        const long long label = gt_sem ? gt_sem[n] : -1;
        float mx = -INFINITY;
        if (label >= 0) {
            for (uint32_t c = 3 + lane; c < 3 + C; c += 32) mx = fmaxf(mx, o[c]);
        }
        mx = wmax(mx);
        float se = 0.f;
        if (label >= 0) {
            for (uint32_t c = 3 + lane; c < 3 + C; c += 32) se += __expf(o[c] - mx);
        }
        se = wsum(se);
        #pragma unroll
        for (int j = 0; j < NC; ++j) {
            const uint32_t c = lane + 32 * j;
            if (c >= K) continue;
            float gv = 0.f;
            if (c < 3) {
                const float diff = o[c] + (1.0f - w) - gt_rgb[(size_t)n * 3 + c];
                l_rgb = fmaf(diff, diff, l_rgb);
                gv = rgb_w * 2.0f * diff * inv_rgb;
                gws -= gv;
            } else if (c < 3 + C) {
                if (label >= 0) {
                    const float logit = o[c];
                    const float p = __expf(logit - mx) / se;
                    const bool hit = (long long)(c - 3) == label;
                    gv = sem_w * inv_cs * (p - (hit ? 1.0f : 0.f));
                    if (hit) l_sem += (mx + __logf(se)) - logit;
                }
            } else {
                const uint32_t f = c - 3 - C;
                if (f < Fg) {
                    const float diff = o[c] - gt_feat[(size_t)n * Fg + f];
                    l_feat += fabsf(diff);
                    gv = feat_w * inv_feat * (diff > 0.f ? 1.0f : (diff < 0.f ? -1.0f : 0.f));
                }
            }
            g[c] = gv * grad_scale;
        }
        gws = wsum(gws);
        if (lane == 0) {
            g_ws[n] = gws * grad_scale;
            float gd = 0.f;
            const float gtd = gt_depth ? gt_depth[n] : 0.f;
            if (gtd > depth_eps) {
                const float inv_norm = 1.0f / norms[n];
                const float diff = depth_raw[n] * inv_norm - gtd;
                l_depth += fabsf(diff);
                gd = depth_w * inv_cd * (diff > 0.f ? 1.0f : (diff < 0.f ? -1.0f : 0.f)) * inv_norm;
            }
            g_depth[n] = gd * grad_scale;
        }
    }
    l_rgb = wsum(l_rgb) * rgb_w * inv_rgb;
    l_depth = wsum(l_depth) * depth_w * inv_cd;
    l_feat = wsum(l_feat) * feat_w * inv_feat;
    l_sem = wsum(l_sem) * sem_w * inv_cs;
    if (lane == 0) {
        atomicAdd(loss + 0, l_rgb + l_depth + l_feat + l_sem);
        atomicAdd(loss + 1, l_rgb);
        atomicAdd(loss + 2, l_depth);
        atomicAdd(loss + 3, l_feat);
        atomicAdd(loss + 4, l_sem);
    }
}

}  // namespace

// loss5 [5] = (total, rgb, depth, feature, semantic) and counts2 [2] are zeroed here; gt_depth / gt_sem / gt_feat may
// be NULL (term skipped).  g_depth is the gradient w.r.t. depth_raw (before the division by the direction norm).
AL_API int al_loss_fwd_bwd(const float* ws, const float* depth_raw, const float* out, uint32_t N, uint32_t C, uint32_t F,
                           const float* norms, const float* gt_rgb, const float* gt_depth, const long long* gt_sem,
                           const float* gt_feat, uint32_t Fg, float rgb_w, float depth_w, float sem_w, float feat_w,
                           float depth_eps, float grad_scale, float* loss5, int* counts2, float* g_ws, float* g_depth,
                           float* g_out, void* stream) {
    if (N == 0) return 0;
    AL_REQUIRE(ws && depth_raw && out && norms && gt_rgb && loss5 && counts2 && g_ws && g_depth && g_out, "null pointer");
    AL_REQUIRE(C >= 1, "at least one semantic class");
    AL_REQUIRE(Fg <= F, "ground-truth feature width exceeds the feature head");
    cudaStream_t st = (cudaStream_t)stream;
    AL_CHECK(cudaMemsetAsync(loss5, 0, 5 * sizeof(float), st));
    AL_CHECK(cudaMemsetAsync(counts2, 0, 2 * sizeof(int), st));
    k_loss_counts<<<al_div_up(N, 1024) < 64 ? al_div_up(N, 1024) : 64, 256, 0, st>>>(gt_depth, gt_sem, N, depth_eps, counts2);
    AL_LAUNCH_CHECK();
    const uint32_t K = 3 + C + F;
    const unsigned grid = al_div_up(N, 8) < (unsigned)al_num_sms() * 4 ? al_div_up(N, 8) : (unsigned)al_num_sms() * 4;
#define AL_LOSS_LAUNCH(NCV)                                                                                              \
    k_loss<NCV><<<grid, 256, 0, st>>>(ws, depth_raw, out, N, C, F, norms, gt_rgb, gt_depth, gt_sem, gt_feat, Fg, rgb_w,  \
                                      depth_w, sem_w, feat_w, depth_eps, grad_scale, counts2, loss5, g_ws, g_depth, g_out)
    if (K <= 32) AL_LOSS_LAUNCH(1);
    else if (K <= 96) AL_LOSS_LAUNCH(3);
    else if (K <= 160) AL_LOSS_LAUNCH(5);
    else if (K <= 640) AL_LOSS_LAUNCH(20);
    else { AL_REQUIRE(K <= 1280, "too many channels"); AL_LOSS_LAUNCH(40); }
#undef AL_LOSS_LAUNCH
    AL_LAUNCH_CHECK();
    return 0;
}
