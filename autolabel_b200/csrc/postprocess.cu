// Render epilogues of the export / render scripts, fused into one pass over the composited frame
// (SURVEY 8(f) rank 3): what the reference does per frame on the host or in per-row Python loops after
// `model.render(...)`:
//
//   label      = outputs['semantic'].argmax(-1)                                   scripts/export.py:78-90
//   text_label = argmax_t  <f / ||f||, text_features[t]>                          scripts/render.py:69-82,
//                                                                                  autolabel/evaluation.py:295-318
//   pca8       = uint8(255 clip(((f - mean) . components^T - min) / range, 0, 1)) scripts/render.py:61-66
//                                                                                  (sklearn PCA.transform, no whitening)
//   rgb8       = uint8(255 image)                                                 scripts/render.py:104
//
// so that 14 bytes per pixel go back to the host instead of the 4 (6 + C + F) bytes of the full maps.
// One warp per pixel: lanes stride over the feature channels (coalesced row read), the row is kept in
// registers / shared memory, text rows stream from L2.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxF = 1024;

__device__ __forceinline__ float wsum(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint8_t to_u8(float x) {   // numpy's float -> uint8 cast truncates; inputs lie in [0, 255]
    return (uint8_t)(int)fminf(fmaxf(x, 0.f), 255.f);
}

__global__ void __launch_bounds__(kWarps * 32) k_render_epilogue(
    const float* __restrict__ image, const float* __restrict__ logits, uint32_t ld_logits,
    const float* __restrict__ feat, uint32_t ld_feat, uint32_t N, uint32_t C, uint32_t F,
    const float* __restrict__ text, uint32_t T, const float* __restrict__ pca_mean,
    const float* __restrict__ pca_comp, const float* __restrict__ pca_min, const float* __restrict__ pca_range,
    uint8_t* __restrict__ rgb8, int* __restrict__ label, int* __restrict__ text_label, uint8_t* __restrict__ pca8) {
    __shared__ float rows[kWarps][kMaxF];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* row = rows[warp];
    for (uint32_t i = blockIdx.x * kWarps + warp; i < N; i += gridDim.x * kWarps) {
        if (rgb8 && lane < 3) rgb8[(size_t)i * 3 + lane] = to_u8(image[(size_t)i * 3 + lane] * 255.0f);
        if (label) {
            // first maximal value, as torch.argmax
            float best = -INFINITY;
            int arg = 0x7fffffff;
            for (uint32_t c = lane; c < C; c += 32) {
                const float v = logits[(size_t)i * ld_logits + c];
                if (v > best || (v == best && (int)c < arg) || arg == 0x7fffffff) { best = v; arg = (int)c; }
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (oa != 0x7fffffff && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
            }
            if (lane == 0) label[i] = arg;
        }
        if (!(text_label || pca8)) continue;
        float ss = 0.f;
        for (uint32_t j = lane; j < F; j += 32) {
            const float v = feat[(size_t)i * ld_feat + j];
            row[j] = v;
            ss += v * v;
        }
        ss = wsum(ss);
        __syncwarp();
        if (text_label) {
            const float inv = 1.0f / sqrtf(ss);                    // features / torch.norm(features)
            float best = -INFINITY;
            int arg = 0;
            for (uint32_t t = 0; t < T; t += 4) {                   // four text rows per pass: independent reductions
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                #pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (t + q < T) {
                        const float* tr = text + (size_t)(t + q) * F;
                        for (uint32_t j = lane; j < F; j += 32) acc[q] = fmaf(row[j] * inv, __ldg(tr + j), acc[q]);
                    }
                }
                #pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float s = wsum(acc[q]);
                    if (t + q < T && (s > best)) { best = s; arg = (int)(t + q); }
                }
            }
            if (lane == 0) text_label[i] = arg;
        }
        if (pca8) {
            float acc[3] = {0.f, 0.f, 0.f};
            for (uint32_t j = lane; j < F; j += 32) {
                const float c = row[j] - __ldg(pca_mean + j);
                #pragma unroll
                for (int q = 0; q < 3; ++q) acc[q] = fmaf(c, __ldg(pca_comp + (size_t)q * F + j), acc[q]);
            }
            #pragma unroll
            for (int q = 0; q < 3; ++q) acc[q] = wsum(acc[q]);
            if (lane < 3) {
                const float a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2]);
                const float u = fminf(fmaxf((a - pca_min[lane]) / pca_range[lane], 0.f), 1.f);
                pca8[(size_t)i * 3 + lane] = to_u8(u * 255.0f);
            }
        }
        __syncwarp();
    }
}

}  // namespace

AL_API int al_render_epilogue(const float* image, const float* logits, uint32_t ld_logits, const float* feat,
                              uint32_t ld_feat, uint32_t N, uint32_t C, uint32_t F, const float* text, uint32_t T,
                              const float* pca_mean, const float* pca_comp, const float* pca_min,
                              const float* pca_range, uint8_t* rgb8, int* label, int* text_label, uint8_t* pca8,
                              void* stream) {
    if (N == 0) return 0;
    AL_REQUIRE(!rgb8 || image, "rgb8 needs image");
    AL_REQUIRE(!label || (logits && C >= 1 && ld_logits >= C), "label needs logits [N, >= C]");
    AL_REQUIRE(!(text_label || pca8) || (feat && F >= 1 && F <= (uint32_t)kMaxF && ld_feat >= F),
               "feature epilogues need features [N, >= F], F <= 1024");
    AL_REQUIRE(!text_label || (text && T >= 1), "text_label needs text features [T, F]");
    AL_REQUIRE(!pca8 || (pca_mean && pca_comp && pca_min && pca_range), "pca8 needs mean [F], components [3, F], min [3], range [3]");
    const unsigned want = al_div_up(N, kWarps);
    const unsigned cap = (unsigned)al_num_sms() * 8;
    k_render_epilogue<<<want < cap ? want : cap, kWarps * 32, 0, (cudaStream_t)stream>>>(
        image, logits, ld_logits, feat, ld_feat, N, C, F, text, T, pca_mean, pca_comp, pca_min, pca_range, rgb8, label,
        text_label, pca8);
    AL_LAUNCH_CHECK();
    return 0;
}
