// Argument block of the tiled tcgen05 GEMMs behind the wide MLP heads (gemm_tc.cu: cp.async ring, any shape;
// gemm_tma.cu: TMA + warp-specialised pipeline for the large layers).
#pragma once
#include "mlp_args.cuh"

struct GemmArgs {
    int mode;                 // 0 F, 1 D, 2 W
    const __half* A; int lda;
    const __half* B; int ldb;
    const __half* Bt; int ldbt;   // D only, optional: the same weights transposed ([N, K] row-major, K contiguous) -- lets the
                              // TMA pipeline run the dgrad with both operands K-major
    int M;                    // rows (samples) capacity
    const int* n_dev;         // live rows
    int N, K;                 // F/D: output columns, reduction length.  W: N = Q extent, K unused
    int P;                    // W: extent of the M side (multiple of 64)
    // F / D epilogue
    int relu;
    const __half* mask; int ldmask;
    __half* Yh; int ldyh;     // fp16 output (optional)
    const float* amax_dev;    // D / W: gradient scale (fp32 outputs are unscaled)
    int unscale;              // multiply fp32 window outputs by 1 / scale
    OutF32 o0, o1;
    OutF16 h0;
    // W epilogue: G(p, q) at G[p * sp + q * sq] += acc / scale
    float* G; int sp, sq;
    int rows_per_item;        // W: samples per work item (multiple of 64)
};

// gemm_tma.cu: runs the GEMM on the TMA pipeline when its shape and alignment qualify; returns -1 when it does not
// (the caller falls back to k_gemm_tc), 0 on success, a CUDA error code otherwise.
int al_gemm_tma_launch(const GemmArgs& a, cudaStream_t st);
