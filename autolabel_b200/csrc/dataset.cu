// Device-resident training-batch sampler and full-frame ray generation.
//
// Replaces the CPU side of autolabel/dataset.py: `_compute_direction` (:17-37, numba),
// `BaseDataset._next_train` (:182-242: per 512-ray chunk one image, fancy-indexed gathers of pixels /
// depth / semantic / features, broadcast origin, jittered directions) and the ray part of `_get_test`
// (:244-266).  The scene arrays stay in HBM in the layouts the reference keeps them in host memory
// (images fp32 [n,HW,3], depths uint16 millimetres [n,HW], semantics uint8 [n,HW] with 0 = unlabeled,
// features fp16 [n, fh*fw, F]); the random draws (image per chunk, pixel index and jitter per ray)
// are inputs, so a batch is reproducible against the reference on identical draws.
//
// One warp per ray: lanes 0..2 produce the per-ray scalars (one coordinate each), all lanes stream the
// feature row (fp16 -> fp32: 4 bytes in / 8 bytes out per lane and step, coalesced).
#include "common.cuh"
#include "../../include/autolabel_b200.h"

namespace {

struct SampleArgs {
    const float* images; const uint16_t* depths; const uint8_t* semantics; const __half* features;
    const float* rotations; const float* origins;
    uint32_t HW, w, fw, fhw, F;
    double sx, sy, fx, fy, cx, cy;
    const int* image_index; int image0; const int* ray_indices; const float* jitter;
    uint32_t n, chunk;
    float* rays_o; float* rays_d; float* norms; float* pixels; float* depth; long long* semantic; float* feat_out;
};

__global__ void __launch_bounds__(256) k_dataset_sample(const SampleArgs a) {
    const uint32_t ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ray >= a.n) return;
    const int img = a.image_index ? __ldg(a.image_index + ray / a.chunk) : a.image0;
    const long long pix = a.ray_indices ? (long long)__ldg(a.ray_indices + ray) : (long long)ray;
    const uint32_t x = (uint32_t)(pix % a.w), y = (uint32_t)(pix / a.w);
    if (lane < 3) {
        // dataset.py:17-37.  (xs - cx) / fx is evaluated in double (float32 array op float64 scalar under numba)
        // and rounded on the store; the squared norm accumulates in float32 in x, y, z order.
        float jx = 0.5f, jy = 0.5f;
        if (a.jitter) { jx = __ldg(a.jitter + 2 * (size_t)ray); jy = __ldg(a.jitter + 2 * (size_t)ray + 1); }
        const float xs = __fadd_rn((float)x, jx), ys = __fadd_rn((float)y, jy);
        const float dx = (float)(((double)xs - a.cx) / a.fx), dy = (float)(((double)ys - a.cy) / a.fy);
        const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 1.0f));
        const float ux = __fdiv_rn(dx, nrm), uy = __fdiv_rn(dy, nrm), uz = __fdiv_rn(1.0f, nrm);
        if (a.rays_d) {
            const float* R = a.rotations + (size_t)img * 9 + lane * 3;       // row `lane` of R_WC
            a.rays_d[(size_t)ray * 3 + lane] =
                __fadd_rn(__fadd_rn(__fmul_rn(__ldg(R), ux), __fmul_rn(__ldg(R + 1), uy)), __fmul_rn(__ldg(R + 2), uz));
        }
        if (a.rays_o) a.rays_o[(size_t)ray * 3 + lane] = __ldg(a.origins + (size_t)img * 3 + lane);
        if (a.pixels) a.pixels[(size_t)ray * 3 + lane] = __ldg(a.images + ((size_t)img * a.HW + pix) * 3 + lane);
        if (lane == 0) {
            if (a.norms) a.norms[ray] = nrm;
            // depth / 1000.0 in double, rounded to float32 on the store (dataset.py:220)
            if (a.depth) a.depth[ray] = (float)((double)__ldg(a.depths + (size_t)img * a.HW + pix) / 1000.0);
            if (a.semantic) a.semantic[ray] = (long long)__ldg(a.semantics + (size_t)img * a.HW + pix) - 1;
        }
    }
    if (a.feat_out) {
        // dataset.py:231-240: xy_features = (xy * scale_factor).astype(int): truncation of a double product
        const long long fxi = (long long)((double)x * a.sx), fyi = (long long)((double)y * a.sy);
        const __half2* src = reinterpret_cast<const __half2*>(a.features + ((size_t)img * a.fhw + (size_t)fyi * a.fw + fxi) * a.F);
        float2* dst = reinterpret_cast<float2*>(a.feat_out + (size_t)ray * a.F);
        for (uint32_t j = lane; j < a.F / 2; j += 32) dst[j] = __half22float2(__ldg(src + j));
    }
}

}  // namespace

AL_API int al_dataset_sample(const float* images, const uint16_t* depths, const uint8_t* semantics, const void* features,
                             const float* rotations, const float* origins, uint32_t w, uint32_t h, uint32_t fw,
                             uint32_t fh, uint32_t F, double fx, double fy, double cx, double cy,
                             const int* image_index, int image0, const int* ray_indices, const float* jitter,
                             uint32_t n_rays, uint32_t chunk, float* rays_o, float* rays_d, float* norms, float* pixels,
                             float* depth, long long* semantic, float* feat_out, void* stream) {
    if (n_rays == 0) return 0;
    AL_REQUIRE(w > 0 && h > 0 && chunk > 0, "empty image or chunk");
    AL_REQUIRE(rotations || !rays_d, "rays_d needs rotations");
    AL_REQUIRE(origins || !rays_o, "rays_o needs origins");
    AL_REQUIRE(images || !pixels, "pixels needs images");
    AL_REQUIRE(depths || !depth, "depth needs depths");
    AL_REQUIRE(semantics || !semantic, "semantic needs semantics");
    AL_REQUIRE(!feat_out || (features && fw > 0 && fh > 0 && F > 0 && F % 2 == 0), "features need a map and an even width");
    AL_REQUIRE(fx != 0.0 && fy != 0.0, "zero focal length");
    SampleArgs a;
    a.images = images; a.depths = depths; a.semantics = semantics; a.features = (const __half*)features;
    a.rotations = rotations; a.origins = origins;
    a.HW = w * h; a.w = w; a.fw = fw; a.fhw = fw * fh; a.F = F;
    a.sx = (double)fw / (double)w; a.sy = (double)fh / (double)h;   // scale_factor of dataset.py:448-449
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy;
    a.image_index = image_index; a.image0 = image0; a.ray_indices = ray_indices; a.jitter = jitter;
    a.n = n_rays; a.chunk = chunk;
    a.rays_o = rays_o; a.rays_d = rays_d; a.norms = norms; a.pixels = pixels; a.depth = depth; a.semantic = semantic;
    a.feat_out = feat_out;
    k_dataset_sample<<<al_div_up((unsigned long long)n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(a);
    AL_LAUNCH_CHECK();
    return 0;
}
