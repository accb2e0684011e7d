// Argument blocks shared by the two MLP back ends (mlp.cu: mma.sync; mlp_tc.cu: tcgen05 / TMEM).
#pragma once
#include "common.cuh"

struct OutF32 {          // dst[row*ld + col0 + j] = act(y[src0 + j]), j < ncols
    float* ptr;
    int ld, col0, src0, ncols, act;  // act: 0 none, 1 sigmoid, 2 exp
};
struct OutF16 {          // fp16 copy (optionally ReLU'd) for the next MLP's input
    __half* ptr;
    int ld, col0, src0, ncols, act;  // act: 0 none, 1 relu
};
// Compositing fused into the output epilogue (tcgen05 back end): instead of writing y[row, src0 : src0 + ncols], the
// epilogue adds  w[row] * act(y[row, src0 + j])  to  out[ray[row] * ld + col0 + j]  -- the front-to-back sum of
// renderer.py:302-311 / raymarching.cu:917-927 with the weights already known (al_composite_rays_weights), so the
// [samples, channels] value matrix never exists.  One reduction per (ray segment of a warp's 32 rows, channel);
// rows with w == 0 are skipped (terminated / exhausted slots may hold garbage).  ncols <= 64.
struct OutSum {
    float* out;
    int ld, col0, src0, ncols, act;
    const float* w;        // [cap] compositing weight per sample
    const int* ray;        // [cap] ray index per sample
};
// Density MLP only (out_pad 16, tcgen05 back end): its output epilogue also builds the input rows of the three heads from
// the row's 16 outputs y = [h0, geo (15)] and the sample's ray direction -- what k_head_inputs (encoding.cu) does in a
// pass of its own over an fp32 copy of y:
//   color_in [n,32] = [SH4(dir) (16), geo (15), 1], semf_in [n,16] = [geo (15), 1], semo_in[:, F : F + 16] = [geo (15), 1]
// (autolabel/models.py:205-209, 253-255).  dirs: [n,3] per sample (sray null) or rays_d [N,3] through sray [n].
struct HeadIn {
    __half* color_in;      // null: unused
    __half* semf_in;
    __half* semo_in;
    int ld_semo, F;
    const float* dirs;
    const int* sray;
};
struct MlpFwdArgs {
    const float* params;
    const __half* x;
    int ldx;
    int cap;
    const int* n_dev;
    OutF32 o0, o1;
    OutF16 h0;
    OutSum sum;            // sum.out == nullptr: unused
    HeadIn hin;            // hin.color_in == nullptr: unused
};
// Where the output gradient of an MLP comes from.  kind 0: a plain fp32 matrix (MlpBwdArgs::dout).  kinds 1-4: the
// four heads of ALNetwork (autolabel/models.py:150-256), whose output gradients are assembled on the fly from
// the compositing backward and the gradients the downstream heads already produced (tcgen05 back end only):
//   G(row, c)  = dL/d vals[row, 1 + c],  c over (rgb 3 | logits C | features F):
//                g_vals[row * ldg + 1 + c]                      (materialised), or
//                w[row] * g_out[sray[row] * K + c]              (rank-1 form of al_composite_train_bwd_weights)
//   gsig(row)  = g_vals[row * ldg]  or  g_sigma[row]
//   1 semantic_out      d y[j] = G(row, 3 + j),                                   j < C
//   2 semantic_features d y[j] = G(row, 3 + C + j) + [relu_feat_j > 0] d_feat[row, j],   j < F   (models.py:253-255)
//                       relu_feat = the fp16 relu(features) the forward wrote as semantic_out's input,
//                       d_feat    = semantic_out's input gradient, columns [0, F)
//   3 color_net         d y[k] = G(row, k) rgb_k (1 - rgb_k),                       k < 3   (sigmoid, models.py:213)
//   4 sigma_net         d y[0] = gsig(row) exp(clamp(h0, -15, 15))                  (trunc_exp, activation.py)
//                       d y[1 + k] = dgeo[row, k], k < 15: the sum of the three heads' input gradients w.r.t.
//                       geo_feat, accumulated by their backward kernels (MlpBwdArgs::dx_acc)
struct DoutSpec {
    int kind;
    const float* g_vals; int ldg;
    const float* w; const float* g_sigma; const float* g_out; const int* sray; int K;
    const float* vals; int ldv; int C, F;
    const __half* relu_feat; int ld_relu;
    const float* d_feat; int ld_dfeat;
    const float* dgeo;                    // [cap, 16]
    const float* h16;                     // [cap, 16] raw density-MLP output
};

struct MlpBwdArgs {
    const float* params;
    const __half* x;       // [cap, ldx] forward input rows
    int ldx;
    int cap;
    const int* n_dev;
    const float* dout;     // fp32, element (row, j) at dout[row*ld_dout + dcol0 + j], j < dncols; rest 0
    int ld_dout, dcol0, dncols;
    const float* amax_dev; // optional: running max |dout| -> power-of-two scale
    float* dparams;        // fp32 [NPARAMS], accumulated with atomics
    float* dx;             // optional
    int dx_mode;           // 0: dx[row*ld_dx + j] = d/dx[dx_c0 + j], j < dx_n
                           // 1: level-major pairs: dx[((j/2)*ld_dx + row)*2 + (j&1)], j < dx_n
    int ld_dx, dx_c0, dx_n;
    int dx_acc;            // row-major d x only (tcgen05 back end): 1 = accumulate (+=) instead of store
    float* dx2;            // optional second row-major window of d x (tcgen05 back end)
    int ld_dx2, dx2_c0, dx2_n, dx2_acc;
    DoutSpec spec;         // spec.kind == 0: use dout above
    int dbg;               // timing experiments only (mlp_tc.cu: al_set_bwd_debug); 0 in normal operation
};

// Real spherical harmonics, degree 4 (16 values), of a direction given in tcnn's [0,1] convention
// (autolabel/models.py:205-207 maps d -> (d+1)/2; tcnn maps back 2x-1).
__device__ __forceinline__ void al_sh4(float x, float y, float z, float* o) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

__device__ __forceinline__ float al_apply_act(float v, int act) {
    if (act == 1) return 1.0f / (1.0f + __expf(-v));
    if (act == 2) return __expf(v);
    return v;
}

// Power-of-two gradient scale: largest 2^k with amax * 2^k <= 64 (fp16 hidden gradients stay in range).
__device__ __forceinline__ float al_grad_scale(const float* amax_dev) {
    float scale = 1.0f;
    if (amax_dev) {
        const float am = *amax_dev;
        if (am > 0.f && am < 3.0e38f) {
            int e;
            frexpf(am, &e);                 // am in [2^(e-1), 2^e)
            e = 6 - e;
            e = max(-40, min(40, e));
            scale = scalbnf(1.0f, e);
        }
    }
    return scale;
}

// tcgen05 back end (mlp_tc.cu): returns -1 when the shape is not instantiated there.
int al_tc_mlp_forward(int in_pad, int hidden, int out_pad, int n_hidden, const MlpFwdArgs& a, cudaStream_t st);
// al_mlp_forward with the full argument block (field.cu: the compositing epilogue has no C-ABI form of its own)
int al_mlp_forward_args(int in_pad, int hidden, int out_pad, int n_hidden, const MlpFwdArgs& a, cudaStream_t st);
int al_tc_mlp_backward(int in_pad, int hidden, int out_pad, int n_hidden, const MlpBwdArgs& a, cudaStream_t st);
// whether the tcgen05 back end instantiates this shape
bool al_tc_mlp_has(int in_pad, int hidden, int out_pad, int n_hidden);

// Wide heads (gemm_tc.cu), for csrc/field.cu: the scaled fp16 output-gradient buffer [cap, out_pad] inside a wide MLP's
// training workspace, and the backward that consumes it (dY filled by the caller, scale = al_grad_scale(amax_dev)).
void* al_wide_dy(int in_pad, int hidden, int out_pad, int n_hidden, int cap, void* workspace);
int al_wide_backward_dy(int in_pad, int hidden, int out_pad, int n_hidden, const void* x_half, int ldx, int cap,
                        const int* n_dev, const float* amax_dev, float* dparams, float* dx, int ld_dx, int dx_c0,
                        int dx_n, void* workspace, cudaStream_t st);
