// ALNetwork as one device-side pipeline: position encoding -> density MLP -> colour / feature /
// semantic heads, forward and backward, sequenced on one stream with no host synchronisation
// (the live sample count is read from device memory by every kernel).
//
// Mirrors autolabel/models.py:150-256 in the variant NeRFRenderer.run() uses
// (torch_ngp/nerf/renderer.py:235-311): density() -> raw geo_feat (no ReLU), color() -> sigmoid,
// semantic() -> (logits, pre-ReLU features); sigma = trunc_exp(h0) (torch_ngp/activation.py).
#include "common.cuh"
#include "mlp_args.cuh"
#include "../../include/autolabel_b200.h"

namespace {

struct Ws {
    __half* x_enc;
    float* h16;
    __half* color_in;
    __half* semf_in;
    __half* semo_in;
    // training only
    float* d_semo_in;
    float* dout_semf;
    float* dgeo_semf;
    float* dgeo_color;
    float* dout_color;
    float* dout_sigma;
    float* d_enc;
    float* amax;
    // heads wider than the weight-resident fused kernels (F = 512 LSeg features, ScanNet label sets): tiled GEMM path
    bool semf_wide, semo_wide;
    int c_pad;                  // semantic_out's padded output width
    void* ws_semf;              // al_mlp_wide_workspace(16, F, F, 2)
    void* ws_semo;              // al_mlp_wide_workspace(F + 16, 64, c_pad, 1)
    size_t bytes;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Ws carve(const al_field_t* f, uint32_t cap, int training, void* base) {
    Ws w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? (void*)((char*)base + off) : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    const size_t F = (size_t)f->feat_dim, c = cap;
    w.c_pad = (f->n_classes + 15) / 16 * 16;
    w.semf_wide = al_mlp_num_params(16, (int)F, (int)F, 2) < 0;
    // the fused semantic_out backward stages d x [128, F + 16] fp32 in shared memory for its two output windows,
    // which fits for F = 64 only; other widths take the tiled GEMM path
    w.semo_wide = al_mlp_num_params((int)F + 16, 64, w.c_pad, 1) < 0 || F > 64;
    w.x_enc = (__half*)take(c * f->in_pad * 2);
    w.h16 = (float*)take(c * 16 * 4);
    w.color_in = (__half*)take(c * 32 * 2);
    w.semf_in = (__half*)take(c * 16 * 2);
    w.semo_in = (__half*)take(c * (F + 16) * 2);
    // wide workspaces: forward buffers first inside each, so the offsets the forward sees do not depend on `training`
    // (semantic_out's is always training-sized: its extras are small; the feature head's comes last)
    w.ws_semo = w.semo_wide ? take(al_mlp_wide_workspace((int)F + 16, 64, w.c_pad, 1, (int)cap, 1)) : nullptr;
    w.ws_semf = w.semf_wide ? take(al_mlp_wide_workspace(16, (int)F, (int)F, 2, (int)cap, training)) : nullptr;
    if (training) {
        w.d_semo_in = (float*)take(c * (F + 16) * 4);
        w.dout_semf = (float*)take(c * F * 4);
        w.dgeo_semf = (float*)take(c * 16 * 4);
        w.dgeo_color = (float*)take(c * 16 * 4);
        w.dout_color = (float*)take(c * 4 * 4);
        w.dout_sigma = (float*)take(c * 16 * 4);
        w.d_enc = (float*)take((size_t)f->L * c * 2 * 4);
        w.amax = (float*)take(4 * 4);
    } else {
        w.d_semo_in = w.dout_semf = w.dgeo_semf = w.dgeo_color = w.dout_color = w.dout_sigma = w.d_enc = w.amax = nullptr;
    }
    w.bytes = off;
    return w;
}

// After the semantic_out backward: output gradients of the feature and colour MLPs.
//   dout_semf[j] = g_feat[j] + relu'(feat[j]) * d_semo_in[j]        (models.py:253-255)
//   dout_color[c] = g_rgb[c] * rgb (1 - rgb)                        (sigmoid, models.py:213)
__global__ void __launch_bounds__(256) k_prep_heads_dout(const float* __restrict__ vals,
                                                         const float* __restrict__ g_vals, uint32_t ldv,
                                                         const float* __restrict__ d_semo_in, uint32_t ld_semo,
                                                         uint32_t F, uint32_t C, uint32_t cap,
                                                         const int* __restrict__ n_dev,
                                                         float* __restrict__ dout_semf,
                                                         float* __restrict__ dout_color) {
    const long long n = n_dev ? min((long long)cap, (long long)*n_dev) : (long long)cap;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t per = F + 4;
    if (i < n * per) {
        const long long r = i / per;
        const uint32_t c = (uint32_t)(i - r * per);
        if (c < F) {
            const float feat = vals[(size_t)r * ldv + 4 + C + c];
            float v = g_vals[(size_t)r * ldv + 4 + C + c];
            if (feat > 0.f) v += d_semo_in[(size_t)r * ld_semo + c];
            dout_semf[(size_t)r * F + c] = v;
        } else {
            const uint32_t k = c - F;
            float v = 0.f;
            if (k < 3) {
                const float rgb = vals[(size_t)r * ldv + 1 + k];
                v = g_vals[(size_t)r * ldv + 1 + k] * rgb * (1.0f - rgb);
            }
            dout_color[(size_t)r * 4 + k] = v;
        }
    }
}

// Output gradient of the density MLP:
//   d h0      = g_sigma * exp(clamp(h0, -15, 15))                    (trunc_exp backward)
//   d geo[i]  = d_semo_in[F+i] + dgeo_semf[i] + dgeo_color[i]        (geo_feat feeds three heads)
__global__ void __launch_bounds__(256) k_prep_sigma_dout(const float* __restrict__ h16,
                                                         const float* __restrict__ g_vals, uint32_t ldv,
                                                         const float* __restrict__ d_semo_in, uint32_t ld_semo,
                                                         uint32_t F, const float* __restrict__ dgeo_semf,
                                                         const float* __restrict__ dgeo_color, uint32_t cap,
                                                         const int* __restrict__ n_dev,
                                                         float* __restrict__ dout_sigma, int density_only) {
    const long long n = n_dev ? min((long long)cap, (long long)*n_dev) : (long long)cap;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * 16) {
        const long long r = i >> 4;
        const uint32_t c = (uint32_t)(i & 15);
        float v;
        if (c == 0) {
            const float h0 = h16[(size_t)r * 16];
            v = g_vals[(size_t)r * ldv] * __expf(fminf(fmaxf(h0, -15.f), 15.f));
        } else if (density_only) {
            v = 0.f;
        } else {
            v = d_semo_in[(size_t)r * ld_semo + F + (c - 1)] + dgeo_semf[(size_t)r * 16 + (c - 1)] +
                dgeo_color[(size_t)r * 16 + (c - 1)];
        }
        dout_sigma[i] = v;
    }
}

int check_field(const al_field_t* f) {
    AL_REQUIRE(f, "null field");
    AL_REQUIRE(f->encoding >= 0 && f->encoding <= 2, "encoding must be 0 (freq), 1 (hg) or 2 (hg+freq)");
    const int width = f->encoding == 0 ? 60 : (f->encoding == 1 ? 2 * (int)f->L : 12 + 2 * (int)f->L);
    AL_REQUIRE(f->in_pad == (width + 15) / 16 * 16, "in_pad must be the encoder width rounded up to 16");
    AL_REQUIRE(f->feat_dim % 16 == 0 && f->feat_dim >= 16, "feat_dim must be a multiple of 16");
    AL_REQUIRE(f->n_classes >= 1 && f->n_classes <= 1008, "n_classes must be in [1,1008]");
    AL_REQUIRE(al_mlp_num_params(16, f->feat_dim, f->feat_dim, 2) >= 0 || al_mlp_wide_num_params(16, f->feat_dim, f->feat_dim, 2) >= 0,
               "feat_dim: neither a fused shape (64) nor a wide shape (multiple of 64, <= 1024)");
    AL_REQUIRE(f->w_sigma, "null sigma parameters");
    AL_REQUIRE(f->encoding == 0 || (f->table && f->offsets && f->L >= 1 && f->L <= 16), "grid encodings need table/offsets, L <= 16");
    return 0;
}

}  // namespace

AL_API size_t al_field_workspace(const al_field_t* f, uint32_t cap, int training) {
    if (!f) return 0;
    return carve(f, cap, training, nullptr).bytes;
}

#define AL_TRY(expr)                 \
    do {                             \
        int _r = (expr);             \
        if (_r != 0) return _r;      \
    } while (0)

// colour / feature / semantic heads on the rows of w.h16 (raw density-MLP outputs): vals[:, 1:4 + C + F]
static int field_heads(const al_field_t* f, const Ws& w, const float* dirs, const int* sray, uint32_t cap,
                       const int* n_dev, float* vals, uint32_t ldv, void* stream, bool inputs_ready = false) {
    const int F = f->feat_dim, C = f->n_classes;
    const float* h16 = w.h16;
    if (!inputs_ready)
        AL_TRY(al_head_inputs(h16, cap, n_dev, dirs, sray, w.color_in, w.semf_in, w.semo_in, (uint32_t)(F + 16),
                              (uint32_t)F, stream));
    // colour MLP -> sigmoid -> vals[:,1:4]
    AL_TRY(al_mlp_forward(32, f->hidden_color, 16, 2, f->w_color, w.color_in, 32, (int)cap, n_dev,
                          vals, (int)ldv, 1, 0, 3, 1,
                          nullptr, 0, 0, 0, 0, 0,
                          nullptr, 0, 0, 0, 0, 0, stream));
    // feature MLP -> vals[:, 4+C : 4+C+F] (pre-ReLU) and relu(.) -> semo_in[:, 0:F]
    if (w.semf_wide)
        AL_TRY(al_mlp_wide_forward(16, F, F, 2, f->w_semf, w.semf_in, 16, (int)cap, n_dev,
                                   vals, (int)ldv, 4 + C, 0, F, 0,
                                   nullptr, 0, 0, 0, 0, 0,
                                   w.semo_in, F + 16, 0, 0, F, 1, w.ws_semf, stream));
    else
        AL_TRY(al_mlp_forward(16, F, F, 2, f->w_semf, w.semf_in, 16, (int)cap, n_dev,
                              vals, (int)ldv, 4 + C, 0, F, 0,
                              nullptr, 0, 0, 0, 0, 0,
                              w.semo_in, F + 16, 0, 0, F, 1, stream));
    // semantic MLP -> logits vals[:, 4:4+C]
    if (w.semo_wide)
        AL_TRY(al_mlp_wide_forward(F + 16, 64, w.c_pad, 1, f->w_semo, w.semo_in, F + 16, (int)cap, n_dev,
                                   vals, (int)ldv, 4, 0, C, 0,
                                   nullptr, 0, 0, 0, 0, 0,
                                   nullptr, 0, 0, 0, 0, 0, w.ws_semo, stream));
    else
        AL_TRY(al_mlp_forward(F + 16, 64, 16, 1, f->w_semo, w.semo_in, F + 16, (int)cap, n_dev,
                              vals, (int)ldv, 4, 0, C, 0,
                              nullptr, 0, 0, 0, 0, 0,
                              nullptr, 0, 0, 0, 0, 0, stream));
    return 0;
}

// Whether the density MLP can build the heads' input rows in its own epilogue (HeadIn: tcgen05 back end).
static bool fused_head_inputs(const al_field_t* f) {
    return al_set_mlp_backend(-1) == 1 && al_tc_mlp_has(f->in_pad, f->hidden, 16, 2);
}
// Density MLP on the encoded rows of w.x_enc: h16 (raw 16 outputs, optional), sigma = exp(h0) at sigma[row * ld_sigma],
// and -- with dirs -- the heads' input rows straight from its epilogue.
static int density_mlp(const al_field_t* f, const Ws& w, uint32_t cap, const int* n_dev, float* h16, float* sigma,
                       uint32_t ld_sigma, const float* dirs, const int* sray, cudaStream_t st) {
    MlpFwdArgs a;
    a.params = f->w_sigma; a.x = w.x_enc; a.ldx = f->in_pad; a.cap = (int)cap; a.n_dev = n_dev;
    a.o0 = {h16, 16, 0, 0, h16 ? 16 : 0, 0};
    a.o1 = {sigma, (int)ld_sigma, 0, 0, 1, 2};
    a.h0 = {nullptr, 0, 0, 0, 0, 0};
    a.sum = {nullptr, 0, 0, 0, 0, 0, nullptr, nullptr};
    a.hin = {nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr};
    if (dirs) a.hin = {w.color_in, w.semf_in, w.semo_in, f->feat_dim + 16, f->feat_dim, dirs, sray};
    return al_mlp_forward_args(f->in_pad, f->hidden, 16, 2, a, st);
}

AL_API int al_field_forward(const al_field_t* f, const float* xyz, const float* dirs, const int* sray,
                            uint32_t cap, const int* n_dev, float* vals, uint32_t ldv, float* h16_out,
                            int density_only, void* workspace, void* stream) {
    if (cap == 0) return 0;
    AL_TRY(check_field(f));
    AL_REQUIRE(xyz && vals && workspace, "null pointer");
    const int F = f->feat_dim, C = f->n_classes;
    AL_REQUIRE(density_only || ldv >= (uint32_t)(4 + C + F), "ldv too small");
    AL_REQUIRE(density_only || (dirs && f->w_color && f->w_semf && f->w_semo), "heads need dirs and parameters");
    const Ws w = carve(f, cap, 0, workspace);
    float* h16 = w.h16;

    AL_TRY(al_encode_position(xyz, cap, n_dev, f->bound, f->encoding, f->table, f->offsets, f->L, f->S, f->H,
                              f->gridtype, w.x_enc, (uint32_t)f->in_pad, stream));
    // density MLP: h16 raw (16 columns) + vals[:,0] = exp(h0) (+ the heads' input rows when its epilogue can build them)
    const bool fuse = !density_only && fused_head_inputs(f);
    if (fuse)
        AL_TRY(density_mlp(f, w, cap, n_dev, h16, vals, ldv, dirs, sray, (cudaStream_t)stream));
    else
        AL_TRY(al_mlp_forward(f->in_pad, f->hidden, 16, 2, f->w_sigma, w.x_enc, f->in_pad, (int)cap, n_dev,
                              h16, 16, 0, 0, 16, 0,
                              vals, (int)ldv, 0, 0, 1, 2,
                              nullptr, 0, 0, 0, 0, 0, stream));
    if (h16_out)
        AL_CHECK(cudaMemcpyAsync(h16_out, h16, (size_t)cap * 16 * sizeof(float), cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
    if (density_only) return 0;
    return field_heads(f, w, dirs, sray, cap, n_dev, vals, ldv, stream, fuse);
}

// The two halves of al_field_forward as separate calls, for the training step with early termination
// (al_compact_alive sits between them): position encoding + density MLP into caller buffers ...
AL_API int al_field_density_pre(const al_field_t* f, const float* xyz, uint32_t cap, const int* n_dev, void* x_enc,
                                float* h16, float* sigma, void* stream) {
    if (cap == 0) return 0;
    AL_TRY(check_field(f));
    AL_REQUIRE(xyz && x_enc && h16 && sigma, "null pointer");
    AL_TRY(al_encode_position(xyz, cap, n_dev, f->bound, f->encoding, f->table, f->offsets, f->L, f->S, f->H,
                              f->gridtype, x_enc, (uint32_t)f->in_pad, stream));
    return al_mlp_forward(f->in_pad, f->hidden, 16, 2, f->w_sigma, x_enc, f->in_pad, (int)cap, n_dev,
                          h16, 16, 0, 0, 16, 0,
                          sigma, 1, 0, 0, 1, 2,
                          nullptr, 0, 0, 0, 0, 0, stream);
}

// ... where the workspace keeps the encoded positions and the raw density-MLP outputs ...
AL_API int al_field_workspace_slots(const al_field_t* f, uint32_t cap, int training, void* workspace, void** x_enc,
                                    void** h16) {
    AL_TRY(check_field(f));
    AL_REQUIRE(workspace && x_enc && h16, "null pointer");
    const Ws w = carve(f, cap, training, workspace);
    *x_enc = w.x_enc;
    *h16 = w.h16;
    return 0;
}

// ... and the colour / feature / semantic heads on the rows already present in those two slots (vals[:, 0] = sigma
// is the caller's; columns 1.. are written here).
AL_API int al_field_heads_forward(const al_field_t* f, const float* dirs, const int* sray, uint32_t cap,
                                  const int* n_dev, float* vals, uint32_t ldv, void* workspace, void* stream) {
    if (cap == 0) return 0;
    AL_TRY(check_field(f));
    const int F = f->feat_dim, C = f->n_classes;
    AL_REQUIRE(dirs && vals && workspace && f->w_color && f->w_semf && f->w_semo, "null pointer");
    AL_REQUIRE(ldv >= (uint32_t)(4 + C + F), "ldv too small");
    return field_heads(f, carve(f, cap, 0, workspace), dirs, sray, cap, n_dev, vals, ldv, stream);
}

// al_field_density_pre for the inference waves: position encoding + density MLP -> sigma [cap], with the heads' input
// rows built in the field workspace by the density MLP's epilogue (no fp32 copy of its outputs, no k_head_inputs pass).
// Follow with al_composite_rays_weights and al_field_heads_forward_sum(inputs_ready = 1) on the same workspace and cap.
AL_API int al_field_density_inputs(const al_field_t* f, const float* xyz, const float* dirs, const int* sray,
                                   uint32_t cap, const int* n_dev, float* sigma, void* workspace, void* stream) {
    if (cap == 0) return 0;
    AL_TRY(check_field(f));
    AL_REQUIRE(xyz && dirs && sigma && workspace, "null pointer");
    AL_REQUIRE(fused_head_inputs(f), "needs the tcgen05 MLP back end");
    const Ws w = carve(f, cap, 0, workspace);
    AL_TRY(al_encode_position(xyz, cap, n_dev, f->bound, f->encoding, f->table, f->offsets, f->L, f->S, f->H,
                              f->gridtype, w.x_enc, (uint32_t)f->in_pad, stream));
    return density_mlp(f, w, cap, n_dev, nullptr, sigma, 1, dirs, sray, (cudaStream_t)stream);
}

// The same three heads with compositing folded into their output epilogues (inference waves): instead of the value
// matrix, each head adds  w[row] * value  to  out[sray[row], channel]  (channels in compositing order rgb | logits |
// features, renderer.py:302-311).  `w` comes from al_composite_rays_weights on the sigma of al_field_density_pre.
// Fused head shapes on the tcgen05 back end only (feat_dim 64, n_classes <= 16): returns an error otherwise, the
// caller keeps the al_field_heads_forward + al_composite_rays pair for those.
AL_API int al_field_heads_forward_sum(const al_field_t* f, const float* dirs, const int* sray, uint32_t cap,
                                      const int* n_dev, const float* w_samples, float* out, uint32_t ld_out,
                                      int inputs_ready, void* workspace, void* stream) {
    if (cap == 0) return 0;
    AL_TRY(check_field(f));
    const int F = f->feat_dim, C = f->n_classes;
    AL_REQUIRE(dirs && sray && w_samples && out && workspace && f->w_color && f->w_semf && f->w_semo, "null pointer");
    AL_REQUIRE(ld_out >= (uint32_t)(3 + C + F), "ld_out too small");
    const Ws w = carve(f, cap, 0, workspace);
    AL_REQUIRE(!w.semf_wide && !w.semo_wide, "fused compositing needs the weight-resident head shapes");
    if (!inputs_ready)
        AL_TRY(al_head_inputs(w.h16, cap, n_dev, dirs, sray, w.color_in, w.semf_in, w.semo_in, (uint32_t)(F + 16),
                              (uint32_t)F, stream));
    const cudaStream_t st = (cudaStream_t)stream;
    MlpFwdArgs a;
    a.hin = {nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr};
    a.cap = (int)cap;
    a.n_dev = n_dev;
    a.o0 = a.o1 = {nullptr, 0, 0, 0, 0, 0};
    a.h0 = {nullptr, 0, 0, 0, 0, 0};
    // colour: sigmoid(y[0:3]) -> out[:, 0:3]
    a.params = f->w_color; a.x = w.color_in; a.ldx = 32;
    a.sum = {out, (int)ld_out, 0, 0, 3, 1, w_samples, sray};
    AL_TRY(al_mlp_forward_args(32, f->hidden_color, 16, 2, a, st));
    // features: y -> out[:, 3 + C : 3 + C + F], relu(y) -> semo_in[:, 0:F]
    a.params = f->w_semf; a.x = w.semf_in; a.ldx = 16;
    a.h0 = {w.semo_in, F + 16, 0, 0, F, 1};
    a.sum = {out, (int)ld_out, 3 + C, 0, F, 0, w_samples, sray};
    AL_TRY(al_mlp_forward_args(16, F, F, 2, a, st));
    // logits -> out[:, 3 : 3 + C]
    a.params = f->w_semo; a.x = w.semo_in; a.ldx = F + 16;
    a.h0 = {nullptr, 0, 0, 0, 0, 0};
    a.sum = {out, (int)ld_out, 3, 0, C, 0, w_samples, sray};
    return al_mlp_forward_args(F + 16, 64, 16, 1, a, st);
}

// Where dL/d(vals) comes from: a materialised matrix (al_composite_train_bwd) or the rank-1 form
// (al_composite_train_bwd_weights).
struct GradSrc {
    const float* g_vals;                  // [cap, ldv] or null
    const float* w; const float* g_sigma; const float* g_out; const int* sray; int K;   // rank-1 (w != null)
};

// Grid of a grid-stride kernel over at most `threads` work items: enough blocks of 256 to fill the machine, no more.
static inline unsigned wide_grid(unsigned long long threads) {
    const unsigned long long want = (threads + 255) / 256;
    const unsigned long long full = (unsigned long long)al_num_sms() * 8;
    return (unsigned)(want < full ? (want ? want : 1) : full);
}

// Wide heads: the scaled fp16 output gradient [cap, out_pad] of a semantic head, written straight into the wide
// MLP's workspace (DoutSpec kinds 1 and 2 of mlp_args.cuh, same formulas):
//   kind 1 semantic_out       dY[r, j] = G(r, 3 + j),                                             j < C
//   kind 2 semantic_features  dY[r, j] = G(r, 3 + C + j) + [relu_feat[r, j] > 0] d_feat[r, j],   j < F
//   G(r, c) = g_vals[r * ldg + 1 + c]  or  w[r] * g_out[sray[r] * K + c]
// One thread per 8 output columns (out_pad, F, ld_relu, ld_dfeat are multiples of 8 / 4: 16-byte accesses), grid-stride
// over the LIVE rows only: rows past n are never read (the GEMM loaders zero-fill them) and a capacity-sized grid of
// empty blocks costs more than the work itself.
__global__ void __launch_bounds__(256) k_wide_dout(int kind, GradSrc gs, int ldg, int C, int F,
                                                   const __half* __restrict__ relu_feat, int ld_relu,
                                                   const float* __restrict__ d_feat, int ld_dfeat, int out_pad,
                                                   uint32_t cap, const int* __restrict__ n_dev,
                                                   const float* __restrict__ amax_dev, __half* __restrict__ dY) {
    const long long n = n_dev ? min((long long)cap, (long long)*n_dev) : (long long)cap;
    const float scale = al_grad_scale(amax_dev);
    const int groups = out_pad >> 3;
    const int ncols = kind == 1 ? C : F;
    const int cbase = kind == 1 ? 3 : 3 + C;
    const long long total = n * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int j0 = (int)(i - r * groups) * 8;
        float v[8];
        #pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (j0 < ncols) {
            if (gs.w) {
                const float wr = gs.w[r];
                const float* g = gs.g_out + (size_t)gs.sray[r] * gs.K + cbase + j0;
                #pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (j0 + e < ncols) v[e] = wr * __ldg(g + e);
            } else {
                const float* g = gs.g_vals + (size_t)r * ldg + 1 + cbase + j0;
                #pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (j0 + e < ncols) v[e] = g[e];
            }
            if (kind == 2) {          // F is a multiple of 16: whole groups
                const uint4 m = *reinterpret_cast<const uint4*>(relu_feat + (size_t)r * ld_relu + j0);
                const float4 d0 = *reinterpret_cast<const float4*>(d_feat + (size_t)r * ld_dfeat + j0);
                const float4 d1 = *reinterpret_cast<const float4*>(d_feat + (size_t)r * ld_dfeat + j0 + 4);
                const float df[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 mm = __half22float2(*reinterpret_cast<const __half2*>(&mw[e]));
                    if (mm.x > 0.f) v[2 * e] += df[2 * e];
                    if (mm.y > 0.f) v[2 * e + 1] += df[2 * e + 1];
                }
            }
        }
        __half2 h[4];
        #pragma unroll
        for (int e = 0; e < 4; ++e)
            h[e] = __floats2half2_rn(fminf(fmaxf(v[2 * e] * scale, -65504.f), 65504.f),
                                     fminf(fmaxf(v[2 * e + 1] * scale, -65504.f), 65504.f));
        *reinterpret_cast<uint4*>(dY + (size_t)r * out_pad + j0) = *reinterpret_cast<const uint4*>(h);
    }
}

// dgeo[r, j] = d_semo_in[r, F + j] (+ extra[r, j]), j < 16: geo_feat's gradient from semantic_out's wide backward
// (and the wide feature head's), the starting value the fused colour / feature kernels accumulate onto.
__global__ void __launch_bounds__(256) k_dgeo_init(const float* __restrict__ d_semo_in, int ld_semo, int F,
                                                   const float* __restrict__ extra, uint32_t cap,
                                                   const int* __restrict__ n_dev, float* __restrict__ dgeo) {
    const long long n = n_dev ? min((long long)cap, (long long)*n_dev) : (long long)cap;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * 16; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i >> 4;
        const int j = (int)(i & 15);
        float v = d_semo_in ? d_semo_in[(size_t)r * ld_semo + F + j] : dgeo[i];
        if (extra) v += extra[i];
        dgeo[i] = v;
    }
}

// tcgen05 back end: the four head backward kernels assemble their output gradients themselves (DoutSpec),
// no glue kernels and no dout buffers in HBM.
static int field_backward_tc(const al_field_t* f, const float* xyz, uint32_t cap, const int* n_dev,
                             const float* vals, uint32_t ldv, const GradSrc& gs, const float* amax,
                             float* g_table, float* g_sigma, float* g_color, float* g_semf, float* g_semo,
                             const Ws& w, cudaStream_t st) {
    const int F = f->feat_dim, C = f->n_classes;
    float* d_feat = w.d_semo_in;          // [cap, F]  semantic_out's input gradient w.r.t. relu(features)
    float* dgeo = w.dgeo_semf;            // [cap, 16] sum of the three heads' input gradients w.r.t. geo_feat
    DoutSpec sp = {};
    sp.g_vals = gs.g_vals; sp.ldg = (int)ldv;
    sp.w = gs.w; sp.g_sigma = gs.g_sigma; sp.g_out = gs.g_out; sp.sray = gs.sray; sp.K = gs.K;
    sp.vals = vals; sp.ldv = (int)ldv; sp.C = C; sp.F = F;
    sp.relu_feat = w.semo_in; sp.ld_relu = F + 16;
    sp.d_feat = d_feat; sp.ld_dfeat = F;
    sp.dgeo = dgeo; sp.h16 = w.h16;
    auto base = [&](int kind, int in_pad, const float* params, const __half* x, int dncols, float* dparams) {
        MlpBwdArgs a = {};
        a.params = params; a.x = x; a.ldx = in_pad; a.cap = (int)cap; a.n_dev = n_dev;
        a.dncols = dncols; a.amax_dev = amax; a.dparams = dparams;
        a.spec = sp; a.spec.kind = kind;
        return a;
    };
    auto run = [&](const MlpBwdArgs& a, int in_pad, int hidden, int out_pad, int nh) -> int {
        const int r = al_tc_mlp_backward(in_pad, hidden, out_pad, nh, a, st);
        if (r == -1) {
            al_set_error("al_field_backward: MLP shape in=%d hidden=%d out=%d is not instantiated in mlp_tc.cu", in_pad, hidden, out_pad);
            return (int)cudaErrorInvalidValue;
        }
        return r;
    };
    if (w.semo_wide) {
        // semantic_out through the tiled GEMM path: d x [cap, F + 16] fp32 = (d relu(features) | d geo | d 1)
        sp.d_feat = w.d_semo_in; sp.ld_dfeat = F + 16;
        __half* dY = (__half*)al_wide_dy(F + 16, 64, w.c_pad, 1, (int)cap, w.ws_semo);
        k_wide_dout<<<wide_grid((unsigned long long)cap * w.c_pad / 8), 256, 0, st>>>(
            1, gs, (int)ldv, C, F, nullptr, 0, nullptr, 0, w.c_pad, cap, n_dev, amax, dY);
        AL_LAUNCH_CHECK();
        AL_TRY(al_wide_backward_dy(F + 16, 64, w.c_pad, 1, w.semo_in, F + 16, (int)cap, n_dev, amax, g_semo,
                                   w.d_semo_in, F + 16, 0, F + 16, w.ws_semo, st));
        if (!w.semf_wide) {
            k_dgeo_init<<<wide_grid((unsigned long long)cap * 16), 256, 0, st>>>(w.d_semo_in, F + 16, F, nullptr, cap,
                                                                                    n_dev, dgeo);
            AL_LAUNCH_CHECK();
        }
    } else {   // semantic_out: input = [relu(features) (F) | geo (15) | 1]  ->  d_feat (store), dgeo (store)
        MlpBwdArgs a = base(1, F + 16, f->w_semo, w.semo_in, C, g_semo);
        a.dx = d_feat; a.dx_mode = 0; a.ld_dx = F; a.dx_c0 = 0; a.dx_n = F;
        a.dx2 = dgeo; a.ld_dx2 = 16; a.dx2_c0 = F; a.dx2_n = 16;
        AL_TRY(run(a, F + 16, 64, 16, 1));
    }
    if (w.semf_wide) {
        // semantic_features through the tiled GEMM path: d x [cap, 16] -> dgeo_color (scratch), then
        // dgeo = (semantic_out's d geo) + (this head's d geo)
        __half* dY = (__half*)al_wide_dy(16, F, F, 2, (int)cap, w.ws_semf);
        k_wide_dout<<<wide_grid((unsigned long long)cap * F / 8), 256, 0, st>>>(
            2, gs, (int)ldv, C, F, w.semo_in, F + 16, sp.d_feat, sp.ld_dfeat, F, cap, n_dev, amax, dY);
        AL_LAUNCH_CHECK();
        AL_TRY(al_wide_backward_dy(16, F, F, 2, w.semf_in, 16, (int)cap, n_dev, amax, g_semf, w.dgeo_color, 16, 0, 16,
                                   w.ws_semf, st));
        k_dgeo_init<<<wide_grid((unsigned long long)cap * 16), 256, 0, st>>>(
            w.semo_wide ? w.d_semo_in : nullptr, F + 16, F, w.dgeo_color, cap, n_dev, dgeo);
        AL_LAUNCH_CHECK();
    } else {   // semantic_features: input = [geo (15) | 1]  ->  dgeo +=
        MlpBwdArgs a = base(2, 16, f->w_semf, w.semf_in, F, g_semf);
        a.spec.d_feat = sp.d_feat; a.spec.ld_dfeat = sp.ld_dfeat;
        a.dx = dgeo; a.dx_mode = 0; a.ld_dx = 16; a.dx_c0 = 0; a.dx_n = 16; a.dx_acc = 1;
        AL_TRY(run(a, 16, F, F, 2));
    }
    {   // color_net: input = [SH (16) | geo (15) | 1]  ->  dgeo +=
        MlpBwdArgs a = base(3, 32, f->w_color, w.color_in, 3, g_color);
        a.dx = dgeo; a.dx_mode = 0; a.ld_dx = 16; a.dx_c0 = 16; a.dx_n = 16; a.dx_acc = 1;
        AL_TRY(run(a, 32, f->hidden_color, 16, 2));
    }
    const bool has_grid = f->encoding != 0 && g_table;
    {   // sigma_net: d out = (trunc_exp' g_sigma | dgeo); the grid part of d x goes out level-major for the scatter
        MlpBwdArgs a = base(4, f->in_pad, f->w_sigma, w.x_enc, 16, g_sigma);
        if (has_grid) {
            a.dx = w.d_enc; a.dx_mode = 1; a.ld_dx = (int)cap; a.dx_c0 = f->encoding == 2 ? 12 : 0; a.dx_n = 2 * (int)f->L;
        }
        AL_TRY(run(a, f->in_pad, f->hidden, 16, 2));
    }
    if (has_grid)
        AL_TRY(al_grid_scatter_xyz(w.d_enc, cap, xyz, cap, n_dev, f->bound, f->encoding == 2 ? 1 : 0, f->offsets,
                                   g_table, f->L, f->S, f->H, f->gridtype, st));
    return 0;
}

AL_API int al_field_backward_rays(const al_field_t* f, const float* xyz, uint32_t cap, const int* n_dev,
                                  const float* vals, uint32_t ldv, const float* w_samples, const float* g_sigma_samples,
                                  const float* g_out, const int* sray, const float* g_amax, float* g_table,
                                  float* g_sigma, float* g_color, float* g_semf, float* g_semo, void* workspace,
                                  void* stream) {
    if (cap == 0) return 0;
    AL_TRY(check_field(f));
    AL_REQUIRE(xyz && vals && w_samples && g_sigma_samples && g_out && sray && g_amax && workspace, "null pointer");
    AL_REQUIRE(al_set_mlp_backend(-1) == 1, "the rank-1 backward needs the tcgen05 MLP back end (al_set_mlp_backend(1))");
    const Ws w = carve(f, cap, 1, workspace);
    GradSrc gs = {nullptr, w_samples, g_sigma_samples, g_out, sray, 3 + f->n_classes + f->feat_dim};
    return field_backward_tc(f, xyz, cap, n_dev, vals, ldv, gs, g_amax, g_table, g_sigma, g_color, g_semf, g_semo, w,
                             (cudaStream_t)stream);
}

AL_API int al_field_backward(const al_field_t* f, const float* xyz, uint32_t cap, const int* n_dev,
                             const float* vals, const float* g_vals, const float* g_amax, uint32_t ldv, float* g_table,
                             float* g_sigma, float* g_color, float* g_semf, float* g_semo, void* workspace,
                             void* stream) {
    if (cap == 0) return 0;
    AL_TRY(check_field(f));
    AL_REQUIRE(xyz && vals && g_vals && workspace, "null pointer");
    const int F = f->feat_dim, C = f->n_classes;
    const Ws w = carve(f, cap, 1, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    // One power-of-two gradient scale for all four MLP backward kernels, derived from max |g_vals|
    // (hidden gradients stay within ~1e3 x of it: 2^6 target leaves 2^10 of fp16 head-room).
    const float* amax = g_amax;
    if (!amax) {
        AL_CHECK(cudaMemsetAsync(w.amax, 0, 4 * sizeof(float), st));
        AL_TRY(al_amax(g_vals, (int)ldv, 0, 4 + C + F, (int)cap, n_dev, w.amax, stream));
        amax = w.amax;
    }
    if (al_set_mlp_backend(-1) == 1) {
        GradSrc gs = {g_vals, nullptr, nullptr, nullptr, nullptr, 0};
        return field_backward_tc(f, xyz, cap, n_dev, vals, ldv, gs, amax, g_table, g_sigma, g_color, g_semf, g_semo, w, st);
    }

    // semantic_out backward: dout = g_logits
    AL_TRY(al_mlp_backward(F + 16, 64, 16, 1, f->w_semo, w.semo_in, F + 16, (int)cap, n_dev, g_vals, (int)ldv, 4, C,
                           amax, g_semo, w.d_semo_in, 0, F + 16, 0, F + 16, stream));
    {
        const unsigned long long work = (unsigned long long)cap * (F + 4);
        k_prep_heads_dout<<<al_div_up(work, 256), 256, 0, st>>>(vals, g_vals, ldv, w.d_semo_in, (uint32_t)(F + 16),
                                                                (uint32_t)F, (uint32_t)C, cap, n_dev, w.dout_semf,
                                                                w.dout_color);
        AL_LAUNCH_CHECK();
    }
    // feature MLP backward -> d geo (columns 0..14 of its input, column 15 is the bias column)
    AL_TRY(al_mlp_backward(16, F, F, 2, f->w_semf, w.semf_in, 16, (int)cap, n_dev, w.dout_semf, F, 0, F, amax,
                           g_semf, w.dgeo_semf, 0, 16, 0, 16, stream));
    // colour MLP backward -> d geo (input columns 16..30)
    AL_TRY(al_mlp_backward(32, f->hidden_color, 16, 2, f->w_color, w.color_in, 32, (int)cap, n_dev, w.dout_color, 4,
                           0, 3, amax, g_color, w.dgeo_color, 0, 16, 16, 16, stream));
    {
        const unsigned long long work = (unsigned long long)cap * 16;
        k_prep_sigma_dout<<<al_div_up(work, 256), 256, 0, st>>>(w.h16, g_vals, ldv, w.d_semo_in, (uint32_t)(F + 16),
                                                                (uint32_t)F, w.dgeo_semf, w.dgeo_color, cap, n_dev,
                                                                w.dout_sigma, 0);
        AL_LAUNCH_CHECK();
    }
    // density MLP backward; grid part of d x goes out level-major for the scatter
    const bool has_grid = f->encoding != 0 && g_table;
    const int grid_c0 = f->encoding == 2 ? 12 : 0;
    AL_TRY(al_mlp_backward(f->in_pad, f->hidden, 16, 2, f->w_sigma, w.x_enc, f->in_pad, (int)cap, n_dev,
                           w.dout_sigma, 16, 0, 16, amax, g_sigma, has_grid ? w.d_enc : nullptr, 1, (int)cap,
                           grid_c0, 2 * (int)f->L, stream));
    if (has_grid)
        AL_TRY(al_grid_scatter_xyz(w.d_enc, cap, xyz, cap, n_dev, f->bound, f->encoding == 2 ? 1 : 0, f->offsets,
                                   g_table, f->L, f->S, f->H, f->gridtype, stream));
    return 0;
}
