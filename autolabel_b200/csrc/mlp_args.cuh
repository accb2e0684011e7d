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
struct MlpFwdArgs {
    const float* params;
    const __half* x;
    int ldx;
    int cap;
    const int* n_dev;
    OutF32 o0, o1;
    OutF16 h0;
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
};

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
int al_tc_mlp_backward(int in_pad, int hidden, int out_pad, int n_hidden, const MlpBwdArgs& a, cudaStream_t st);
