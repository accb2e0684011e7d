/* autolabel_b200 — C ABI of the B200-native feature-field hot path.
 *
 * One shared library (autolabel_b200/libautolabel_b200.so), plain pointers and sizes, no torch
 * types.  Every function returns 0 on success or a cudaError_t value; al_last_error() returns a
 * human-readable description of the last failure on the calling thread.  All pointers are DEVICE
 * pointers unless stated otherwise; buffers are caller-allocated, nothing is retained after
 * return, every launch goes to the `stream` argument (a cudaStream_t passed as void*).
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * tree, ethz-asl/autolabel @ 7d06358).  INTEGRATION.md shows the binding a reference maintainer
 * would add.
 */
#ifndef AUTOLABEL_B200_H
#define AUTOLABEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* al_last_error(void);
int al_abi_version(void);
int al_sm_count(void);
/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
unsigned long long al_launch_count(void);

/* ------------------------------------------------------------------ _raymarching (bindings.cpp:5-19) */

/* near_far_from_aabb — torch_ngp/raymarching/src/raymarching.h:7, raymarching.cu:98-199.
 * near_idx / far_idx (hit-face ids, 255 = miss) may be NULL. */
int al_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                          float min_near, float* nears, float* fars, uint8_t* near_idx,
                          uint8_t* far_idx, void* stream);

/* morton3D / morton3D_invert — raymarching.h:9-10, raymarching.cu:257-303. */
int al_morton3d(const int* coords, uint32_t N, int* indices, void* stream);
int al_morton3d_invert(const int* indices, uint32_t N, int* coords, void* stream);

/* packbits — raymarching.h:11, raymarching.cu:310-343.  N = output bytes.  thresh_dev (optional
 * device float): effective threshold = min(thresh, *thresh_dev), replacing the host-side
 * min(mean_density, density_thresh) of torch_ngp/nerf/renderer.py:671. */
int al_packbits(const float* grid, uint32_t N, float thresh, const float* thresh_dev,
                uint8_t* bitfield, void* stream);

/* march_rays_train — raymarching.h:13, raymarching.cu:354-537.
 * Same outputs as the reference (xyzs [M,3], dirs [M,3], deltas [M,2], ts [M], rays [N,3] =
 * (ray id, offset, count), counter[0] += samples, counter[1] += rays); segment offsets are the
 * exclusive scan of the counts in ray order.  Optional pointers may be NULL.
 *   nears/fars NULL  -> slab test fused in from (aabb, min_near); nears_out/fars_out receive it
 *   tpos [M]         -> chain parameter t at which each position was evaluated
 *   sray [M]         -> ray id of each sample
 *   meta  int[2]     -> {samples actually written, samples counted}
 *   workspace        -> al_march_rays_train_workspace(N, max_steps) bytes */
size_t al_march_rays_train_workspace(uint32_t N, uint32_t max_steps);
/* Two-phase form: _count runs the DDA + scan (rays, counter, meta), _write expands the recorded
 * chain into sample records; a caller may size its buffers from meta[1] in between. */
int al_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid,
                              float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C,
                              uint32_t H, uint32_t M, const float* nears, const float* fars,
                              const float* aabb, float min_near, float* nears_out, float* fars_out,
                              int* rays, int* counter, int* meta, uint32_t perturb, void* workspace,
                              void* stream);
int al_march_rays_train_write(const float* rays_o, const float* rays_d, float bound, float dt_gamma,
                              uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                              const int* rays, float* xyzs, float* dirs, float* deltas, float* ts,
                              float* tpos, int* sray, const void* workspace, void* stream);
int al_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                        float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                        uint32_t M, const float* nears, const float* fars, const float* aabb,
                        float min_near, float* nears_out, float* fars_out, float* xyzs, float* dirs,
                        float* deltas, float* ts, float* tpos, int* sray, int* rays, int* counter,
                        int* meta, uint32_t perturb, void* workspace, void* stream);
/* Same, with the sample budget of raymarching.py:324-327 (`mean_count` rounded up to 128) read from DEVICE memory:
 * M is the capacity of the sample buffers, the overflow rule of raymarching.cu:458-459 uses min(M, *budget_dev).
 * The budget can then follow the running mean of the last steps' totals without changing the launch geometry
 * (one captured CUDA graph, no host read-back).  budget_dev NULL == al_march_rays_train; with a budget, dropped
 * rays are reported with count 0 in `rays`. */
int al_march_rays_train_budget(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                               float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                               uint32_t M, const int* budget_dev, const float* nears, const float* fars,
                               const float* aabb, float min_near, float* nears_out, float* fars_out,
                               float* xyzs, float* dirs, float* deltas, float* ts, float* tpos, int* sray,
                               int* rays, int* counter, int* meta, uint32_t perturb, void* workspace,
                               void* stream);

/* composite_rays_train_forward / _backward — raymarching.h:14-15, raymarching.cu:547-740,
 * generalised to K value channels (image + semantic logits + feature vector; the reference
 * composites those in PyTorch, torch_ngp/nerf/renderer.py:243-311) and with the depth gradient
 * the reference drops (raymarching.py:437-438).
 *   sigmas: element i at sigmas[i*ld_sigma]; vals: row i at vals + i*ldv (K channels)
 *   tpos NULL -> depth accumulates the running sum of deltas[.,1] (reference kernel);
 *   tpos given -> depth = sum w * tpos (renderer.run() semantics, renderer.py:273-275)
 *   depth_sq (sum w t^2), xyzs/coords (sum w xyz) optional. */
int al_composite_train_fwd(const float* sigmas, uint32_t ld_sigma, const float* vals, uint32_t ldv,
                           uint32_t K, const float* deltas, const float* tpos, const float* xyzs,
                           const int* rays, uint32_t M, uint32_t N, float sigma_scale,
                           float* weights_sum, float* depth, float* depth_sq, float* out,
                           float* coords, void* stream);
int al_composite_train_bwd(const float* g_ws, const float* g_depth, const float* g_out,
                           const float* sigmas, uint32_t ld_sigma, const float* vals, uint32_t ldv,
                           uint32_t K, const float* deltas, const float* tpos, const int* rays,
                           const float* weights_sum, const float* depth, const float* out, uint32_t M,
                           uint32_t N, float sigma_scale, float* g_sigmas, uint32_t ld_gsigma,
                           float* g_vals, uint32_t ld_gv, float* amax_out, void* stream);

/* Rank-1 form of the backward for the fused training path: dL/dvals[i, c] = w[i] * g_out[ray(i), c], so only the
 * compositing weight w [M] and dL/dsigma [M] are written (8 bytes per sample instead of 4 (1 + K)).  Same inputs
 * as al_composite_train_bwd; amax_out = max(|dL/dsigma|, w * max_c |g_out|). */
int al_composite_train_bwd_weights(const float* g_ws, const float* g_depth, const float* g_out,
                                   const float* sigmas, uint32_t ld_sigma, const float* vals, uint32_t ldv,
                                   uint32_t K, const float* deltas, const float* tpos, const int* rays,
                                   const float* weights_sum, const float* depth, const float* out, uint32_t M,
                                   uint32_t N, float sigma_scale, float* w_out, float* g_sigmas,
                                   float* amax_out, void* stream);

/* march_rays / composite_rays / compact_rays — raymarching.h:17-19, raymarching.cu:747-990. */
int al_march_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, const float* rays_t,
                  const float* rays_o, const float* rays_d, float bound, float dt_gamma,
                  uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears,
                  const float* fars, float* xyzs, float* dirs, float* deltas, float* tpos, int* sray,
                  uint32_t perturb, void* stream);
int al_composite_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, float* rays_t,
                      const float* sigmas, uint32_t ld_sigma, const float* vals, uint32_t ldv,
                      uint32_t K, const float* deltas, const float* tpos, const float* xyzs,
                      float sigma_scale, float* weights_sum, float* depth, float* depth_sq, float* out,
                      float* coords, void* stream);
int al_compact_rays(uint32_t n_alive, int* rays_alive, const int* rays_alive_old, float* rays_t,
                    const float* rays_t_old, int* alive_counter, void* stream);
/* composite_rays without the value channels: the ray-level sums (weights_sum, depth, depth_sq, coords), rays_t and the
 * stopping rule of raymarching.cu:895-947, plus the compositing weight of every slot, w_out [n_alive * n_step]
 * (0 for the slots behind the point where a ray stopped).  The channel sums are then taken inside the head kernels
 * (al_field_heads_forward_sum), so the [samples, channels] value matrix of renderer.py:440-460 is never written. */
int al_composite_rays_weights(uint32_t n_alive, uint32_t n_step, const int* rays_alive, float* rays_t,
                              const float* sigmas, uint32_t ld_sigma, const float* deltas, const float* tpos,
                              const float* xyzs, float sigma_scale, float* weights_sum, float* depth,
                              float* depth_sq, float* coords, float* w_out, void* stream);

/* ------------------------------------------------------------------ _gridencoder (bindings.cpp:5-6) */

/* grid_encode_forward — torch_ngp/gridencoder/src/gridencoder.h:11, gridencoder.cu:75-223,344-371.
 * fp32 tables; outputs [L,B,C]; dy_dx [B,L,D,C].  dbg_indices (optional int [B,L,2^D]) receives
 * the table entry index of every corner (-1 for out-of-range inputs). */
int al_grid_encode_forward(const float* inputs, const float* embeddings, const int* offsets,
                           float* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                           uint32_t H, int calc_grad_inputs, float* dy_dx, uint32_t gridtype,
                           int* dbg_indices, void* stream);
/* grid_encode_backward — gridencoder.h:12, gridencoder.cu:226-341,373-413 (accumulates, +=). */
int al_grid_encode_backward(const float* grad, const float* inputs, const int* offsets,
                            float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                            float S, uint32_t H, int calc_grad_inputs, const float* dy_dx,
                            float* grad_inputs, uint32_t gridtype, void* stream);

/* ------------------------------------------------------------------ tinycudann (external; models.py:10) */

/* tcnn.Encoding {"otype": "Frequency"} — autolabel/models.py:19-22,34-38. out [B, D*2*n_freq]. */
int al_freq_encode(const float* x, uint32_t B, uint32_t D, uint32_t n_freq, float* out, void* stream);
/* tcnn.Encoding {"otype": "SphericalHarmonics", "degree": 4} — models.py:97-101; table
 * torch_ngp/shencoder/src/shencoder.cu:50-73.  Input in tcnn's [0,1] convention. out [B,16]. */
int al_sh_encode(const float* d01, uint32_t B, float* out, void* stream);

/* tcnn.Network (FullyFusedMLP / CutlassMLP, bias-free, ReLU hidden, linear out) — models.py:84-136.
 * params: flat fp32 [W1 (hidden x in_pad), W2 (hidden x hidden) if n_hidden == 2, Wo (out_pad x
 * hidden)], each row-major [out, in].  x: fp16 [cap, ldx]. */
int al_mlp_num_params(int in_pad, int hidden, int out_pad, int n_hidden);
int al_mlp_forward(int in_pad, int hidden, int out_pad, int n_hidden, const float* params,
                   const void* x_half, int ldx, int cap, const int* n_dev,
                   float* o0, int o0_ld, int o0_col0, int o0_src0, int o0_ncols, int o0_act,
                   float* o1, int o1_ld, int o1_col0, int o1_src0, int o1_ncols, int o1_act,
                   void* h0_half, int h0_ld, int h0_col0, int h0_src0, int h0_ncols, int h0_act,
                   void* stream);
int al_mlp_backward(int in_pad, int hidden, int out_pad, int n_hidden, const float* params,
                    const void* x_half, int ldx, int cap, const int* n_dev, const float* dout,
                    int ld_dout, int dcol0, int dncols, const float* amax_dev, float* dparams,
                    float* dx, int dx_mode, int ld_dx, int dx_c0, int dx_n, void* stream);
/* Wide heads (tcnn CutlassMLP: the 512-d LSeg feature head, ScanNet label sets; autolabel/models.py:115-136):
 * same contract and parameter layout as al_mlp_forward / al_mlp_backward for shapes whose weights do not fit the
 * fused kernels (hidden a multiple of 64 up to 1024, widths multiples of 16), run layer by layer as tiled tcgen05
 * GEMMs with fp16 activations in `workspace` (al_mlp_wide_workspace bytes; training != 0 adds the gradient buffers).
 * al_mlp_wide_backward must follow al_mlp_wide_forward on the same workspace; dx is a row-major window. */
int al_mlp_wide_num_params(int in_pad, int hidden, int out_pad, int n_hidden);
size_t al_mlp_wide_workspace(int in_pad, int hidden, int out_pad, int n_hidden, int cap, int training);
int al_mlp_wide_forward(int in_pad, int hidden, int out_pad, int n_hidden, const float* params,
                        const void* x_half, int ldx, int cap, const int* n_dev,
                        float* o0, int o0_ld, int o0_col0, int o0_src0, int o0_ncols, int o0_act,
                        float* o1, int o1_ld, int o1_col0, int o1_src0, int o1_ncols, int o1_act,
                        void* h0_half, int h0_ld, int h0_col0, int h0_src0, int h0_ncols, int h0_act,
                        void* workspace, void* stream);
int al_mlp_wide_backward(int in_pad, int hidden, int out_pad, int n_hidden, const float* params,
                         const void* x_half, int ldx, int cap, const int* n_dev, const float* dout,
                         int ld_dout, int dcol0, int dncols, const float* amax_dev, float* dparams,
                         float* dx, int ld_dx, int dx_c0, int dx_n, void* workspace, void* stream);

/* Back end of al_mlp_forward / al_mlp_backward (and of the fused field built on them):
 *   1 = tcgen05.mma with TMEM accumulators (csrc/mlp_tc.cu; default), 0 = mma.sync (csrc/mlp.cu, the
 *   recompiled-legacy-tensor-path baseline).  Returns the previous value; any other argument only queries.
 *   The environment variable AL_MLP_BACKEND=mma|tc sets the initial value. */
int al_set_mlp_backend(int backend);
/* Timing experiments on the two-tile MLP backward (tools/job_bwd_dbg.sh): bit 0 skips the epilogues, bit 1 issues no GEMMs,
 * bit 2 skips output-gradient assembly and the d-x write-out -- RESULTS ARE WRONG while a bit is set.  bits < 0: query.
 * Returns the previous value; 0 (the default) is normal operation. */
int al_set_bwd_debug(int bits);
int al_amax(const float* v, int ld, int col0, int ncols, int cap, const int* n_dev, float* amax,
            void* stream);

/* ------------------------------------------------------------------ fused field (ALNetwork) */

/* Position encoder of autolabel/models.py:15-59,138-148 as fp16 MLP input rows.
 * mode 0 'freq', 1 'hg', 2 'hg+freq'. */
int al_encode_position(const float* xyz, uint32_t cap, const int* n_dev, float bound, int mode,
                       const float* table, const int* offsets, uint32_t L, float S, uint32_t H,
                       uint32_t gridtype, void* out_half, uint32_t ldo, void* stream);
int al_head_inputs(const float* h16, uint32_t cap, const int* n_dev, const float* dirs, const int* sray,
                   void* color_in, void* semf_in, void* semo_in, uint32_t ld_semo, uint32_t F,
                   void* stream);
int al_grid_scatter_xyz(const float* grad, uint32_t ld_level, const float* xyz, uint32_t cap,
                        const int* n_dev, float bound, int clip, const int* offsets,
                        float* grad_embeddings, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                        void* stream);

/* ALNetwork description (autolabel/models.py:62-136; sizes from autolabel/model_utils.py:61-74). */
typedef struct al_field {
    int encoding;      /* 0 freq, 1 hg, 2 hg+freq                          (models.py:138-148) */
    int in_pad;        /* encoder width rounded up to 16: 64 / 32 / 48                         */
    int hidden;        /* sigma_net width, n_hidden = 2                     (models.py:84-92)  */
    int hidden_color;  /* color_net width, n_hidden = 2                     (models.py:104-113)*/
    int feat_dim;      /* hidden_dim_semantic F                             (models.py:115-126)*/
    int n_classes;     /* semantic_classes C (<= 16 in this build)          (models.py:127-136)*/
    float bound;
    uint32_t L, H, gridtype;
    float S;           /* log2(per_level_scale)                                                */
    const int* offsets;      /* device int [L+1]                                               */
    const float* table;      /* device fp32 [offsets[L], 2]                                    */
    const float* w_sigma;    /* flat fp32 parameter vectors (layout: al_mlp_forward)           */
    const float* w_color;
    const float* w_semf;
    const float* w_semo;
} al_field_t;

/* Bytes of per-sample scratch the field needs for `cap` samples (forward / training). */
size_t al_field_workspace(const al_field_t* f, uint32_t cap, int training);

/* ALNetwork.density + color + semantic (models.py:175-256) on `cap` (live: *n_dev) samples.
 *   xyz [cap,3]; dirs: per-sample [cap,3] (sray NULL) or per-ray [N,3] indexed by sray [cap]
 *   vals [cap, ldv] row = [sigma, r, g, b, logits (C), features (F)],  ldv >= 4 + C + F
 *   h16 [cap,16] (optional out) = raw density-MLP output [h0, geo_feat(15)]
 *   density_only != 0: only vals[.,0] (and h16) are produced.
 * The workspace keeps the fp16 MLP inputs for al_field_backward. */
int al_field_forward(const al_field_t* f, const float* xyz, const float* dirs, const int* sray,
                     uint32_t cap, const int* n_dev, float* vals, uint32_t ldv, float* h16_out,
                     int density_only, void* workspace, void* stream);

/* Backward of al_field_forward: g_vals [cap, ldv] -> parameter gradients (accumulated, +=).
 * g_* may be NULL to skip a parameter group.  Must follow al_field_forward on the same workspace.
 * g_amax (optional device float) = max |g_vals| (al_composite_train_bwd's amax_out); when NULL it is
 * computed here.  It fixes the power-of-two scale that keeps the fp16 hidden gradients in range. */
int al_field_backward(const al_field_t* f, const float* xyz, uint32_t cap, const int* n_dev,
                      const float* vals, const float* g_vals, const float* g_amax, uint32_t ldv, float* g_table,
                      float* g_sigma, float* g_color, float* g_semf, float* g_semo, void* workspace,
                      void* stream);

/* al_field_backward with the output gradient in the rank-1 form of al_composite_train_bwd_weights
 * (w_samples [cap], g_sigma_samples [cap], g_out [N, 3 + C + F], sray [cap]); tcgen05 back end only. */
int al_field_backward_rays(const al_field_t* f, const float* xyz, uint32_t cap, const int* n_dev,
                           const float* vals, uint32_t ldv, const float* w_samples, const float* g_sigma_samples,
                           const float* g_out, const int* sray, const float* g_amax, float* g_table,
                           float* g_sigma, float* g_color, float* g_semf, float* g_semo, void* workspace,
                           void* stream);

/* ------------------------------------------------------------------ occupancy grid + optimiser */

/* EMA-max update + mean of NeRFRenderer.update_extra_state (torch_ngp/nerf/renderer.py:662-667):
 *   valid = grid >= 0 && tmp >= 0;  grid[valid] = max(grid*decay, tmp);  *mean = mean(max(grid,0)).
 * mean_out: device float (zeroed by this call). */
int al_density_grid_update(float* grid, const float* tmp_grid, uint32_t n_cells, float decay,
                           float* mean_out, void* stream);

/* NeRFRenderer.mark_untrained_grid (torch_ngp/nerf/renderer.py:479-561): cells of the cascaded occupancy grid
 * ([C, H^3], Morton order) that lie outside every camera frustum are set to -1.  poses: device fp32
 * [n_poses, 4, 4] camera-to-world. */
int al_mark_untrained_grid(float* grid, const float* poses, uint32_t n_poses, float fx, float fy, float cx,
                           float cy, float bound, uint32_t C, uint32_t H, void* stream);

/* Losses of SimpleTrainer.train_step (autolabel/trainer.py:54-94) and their gradients w.r.t. the compositing outputs,
 * without host synchronisation:  loss = rgb_w MSE(image, gt_rgb) + depth_w mean|depth - gt_depth| over gt_depth > eps
 * + feat_w L1(features[:, :Fg], gt_feat) + sem_w CE(logits[label >= 0]);  image = out[:, :3] + (1 - ws) (white
 * background), depth = depth_raw / norms.  out / g_out: [N, 3 + C + F] rows (rgb, logits, features).
 * loss5 [5] = (total, rgb, depth, feature, semantic), counts2 [2] scratch; gt_depth / gt_sem (int64, -1 = unlabeled) /
 * gt_feat may be NULL.  g_depth is w.r.t. depth_raw.  grad_scale multiplies every gradient. */
int al_loss_fwd_bwd(const float* ws, const float* depth_raw, const float* out, uint32_t N, uint32_t C, uint32_t F,
                    const float* norms, const float* gt_rgb, const float* gt_depth, const long long* gt_sem,
                    const float* gt_feat, uint32_t Fg, float rgb_w, float depth_w, float sem_w, float feat_w,
                    float depth_eps, float grad_scale, float* loss5, int* counts2, float* g_ws, float* g_depth,
                    float* g_out, void* stream);

/* torch.optim.Adam step (scripts/train.py:50-63: lr 5e-3, betas (0.9,0.99), eps 1e-15, L2 weight
 * decay on the MLPs) fused with gradient unscale and zeroing.  step >= 1. */
int al_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                 float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                 int zero_grad, void* stream);

/* The same update for up to AL_ADAM_MAX_TENSORS tensors in ONE launch, with the step count and the learning rate in
 * device memory (step_dev is incremented on the stream first, then read): no host-side state, so the optimiser sits
 * inside the CUDA graph of a training step.  Betas are doubles: the bias corrections 1 - beta^step are evaluated in
 * double like torch does on the host (torch/optim/adam.py). */
#define AL_ADAM_MAX_TENSORS 8
typedef struct {
    float* param;
    float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    size_t n;
    float weight_decay;
} al_adam_tensor_t;
int al_adam_multi(const al_adam_tensor_t* tensors, int count, const float* lr_dev, int* step_dev, double beta1,
                  double beta2, float eps, float grad_scale, int zero_grad, void* stream);

/* Data-parallel training: gradient exchange + Adam as ONE kernel over NVLink / NVSwitch peer memory.
 * Replaces all_reduce(param.grad) followed by the optimiser on every rank (the reference's multi-GPU form is
 * DistributedDataParallel + torch.optim.Adam, torch_ngp/nerf/utils.py:378-380, scripts/train.py:50-63).
 * The flat gradient and parameter buffers of every rank live in symmetric memory; rank `rank` owns elements
 * [shard_begin, shard_end): it sums their gradients over all replicas (multimem.ld_reduce on the multicast address
 * when mc_grad != NULL, else loads through the peer pointers in rank order), applies Adam with ITS moments of the shard
 * (exp_avg / exp_avg_sq hold shard_end - shard_begin elements) and writes the new parameters into every replica.
 * grad_ptrs / param_ptrs: host arrays of `world` device pointers in rank order.  Elements >= wd_begin take
 * `weight_decay`.  Bounds are multiples of 4 elements.  The caller synchronises the ranks before (all backward passes
 * done) and after (all parameter writes landed, all gradients read) the launch, then clears its own gradient buffer. */
int al_peer_adam_step(const void* const* grad_ptrs, const void* const* param_ptrs, float* mc_grad, float* mc_param,
                      float* exp_avg, float* exp_avg_sq, size_t shard_begin, size_t shard_end, size_t wd_begin,
                      int world, int rank, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int step, float grad_scale, void* stream);

/* ------------------------------------------------------------------ device-resident dataset */

/* Training-batch sampler and full-frame ray generation: autolabel/dataset.py `_compute_direction` (:17-37),
 * `BaseDataset._next_train` (:182-242) and the ray part of `_get_test` (:244-266), with the scene arrays resident
 * in HBM in the reference's own layouts: images fp32 [n, h*w, 3], depths uint16 millimetres [n, h*w], semantics
 * uint8 [n, h*w] (0 = unlabeled), features fp16 [n, fh*fw, F], rotations fp32 [n, 3, 3] (R_WC), origins fp32 [n, 3].
 * The random draws are inputs: image_index [ceil(n_rays / chunk)] (NULL: every ray from image `image0`),
 * ray_indices [n_rays] flat pixel indices (NULL: 0..n_rays-1, i.e. a full frame), jitter [n_rays, 2] in [0,1)
 * (NULL: pixel centres, +0.5).  Outputs (each may be NULL): rays_o / rays_d [n_rays, 3], norms [n_rays]
 * (direction_norms), pixels [n_rays, 3], depth [n_rays] metres, semantic int64 [n_rays] (label - 1, -1 = unlabeled),
 * feat_out fp32 [n_rays, F] (nearest feature-map cell, `(xy * scale_factor).astype(int)`). */
int al_dataset_sample(const float* images, const uint16_t* depths, const uint8_t* semantics, const void* features,
                      const float* rotations, const float* origins, uint32_t w, uint32_t h, uint32_t fw, uint32_t fh,
                      uint32_t F, double fx, double fy, double cx, double cy, const int* image_index, int image0,
                      const int* ray_indices, const float* jitter, uint32_t n_rays, uint32_t chunk, float* rays_o,
                      float* rays_d, float* norms, float* pixels, float* depth, long long* semantic, float* feat_out,
                      void* stream);

/* ------------------------------------------------------------------ training-time early termination */

/* The two halves of al_field_forward as separate calls, with al_compact_alive between them:
 *   al_field_density_pre      position encoding + density MLP (models.py:175-190) on every marched sample, into
 *                             caller buffers: x_enc [cap, in_pad] fp16, h16 [cap,16] fp32, sigma [cap] = exp(h0)
 *   al_field_workspace_slots  where a field workspace keeps x_enc / h16 (the compaction writes the alive rows there)
 *   al_field_heads_forward    colour / feature / semantic heads on the rows of those slots -> vals[:, 1:] */
int al_field_density_pre(const al_field_t* f, const float* xyz, uint32_t cap, const int* n_dev, void* x_enc,
                         float* h16, float* sigma, void* stream);
int al_field_workspace_slots(const al_field_t* f, uint32_t cap, int training, void* workspace, void** x_enc,
                             void** h16);
int al_field_heads_forward(const al_field_t* f, const float* dirs, const int* sray, uint32_t cap, const int* n_dev,
                           float* vals, uint32_t ldv, void* workspace, void* stream);
/* The heads with compositing folded into their output epilogues (inference waves, renderer.py:440-460): adds
 * w_samples[row] * (rgb | logits | features)[row] to out[sray[row] * ld_out + channel].  Weight-resident head shapes
 * on the tcgen05 back end only (feat_dim 64, n_classes <= 16); an error otherwise. */
int al_field_heads_forward_sum(const al_field_t* f, const float* dirs, const int* sray, uint32_t cap,
                               const int* n_dev, const float* w_samples, float* out, uint32_t ld_out,
                               int inputs_ready, void* workspace, void* stream);
/* al_field_density_pre for those waves: encoder + density MLP -> sigma [cap]; the heads' input rows (models.py:205-209,
 * 253-255) are built in the workspace by the density MLP's epilogue, so al_field_heads_forward_sum runs with
 * inputs_ready = 1 on the same workspace and cap. */
int al_field_density_inputs(const al_field_t* f, const float* xyz, const float* dirs, const int* sray, uint32_t cap,
                            const int* n_dev, float* sigma, void* workspace, void* stream);

/* Alive-prefix compaction.  The reference's marched inference kernel stops a ray after the sample that brings its
 * transmittance below 1e-4 (raymarching.cu:929-935); its training kernels composite every marched sample
 * (raymarching.cu:593-594,697: the break is commented out) although everything behind that point carries a total
 * weight < 1e-4.  The samples of ray n whose transmittance BEFORE the sample is >= t_thresh are a prefix of its
 * segment; this call packs those prefixes densely, in ray order:
 *   in : sigma [M], deltas [M,2], rays [N,3] (march_rays_train), xyzs [M,3], tpos [M] (optional), sray [M],
 *        x_enc [M, in_pad] fp16 (in_pad % 8 == 0), h16 [M,16]
 *   out: rays_c [N,3] = (ray id, compact offset, alive count), meta_c [2] = {alive samples, alive samples}, the
 *        alive rows of each array; sigma goes to sigma_c[i * ld_sigma_c] (column 0 of the vals matrix)
 * alive_ws: int [N] scratch.  t_thresh <= 0 keeps every sample. */
int al_compact_alive(const float* sigma, const float* deltas, const int* rays, uint32_t M, uint32_t N,
                     float sigma_scale, float t_thresh, const float* xyzs, const float* tpos, const int* sray,
                     const void* x_enc, uint32_t in_pad, const float* h16, int* rays_c, int* meta_c, float* xyzs_c,
                     float* deltas_c, float* tpos_c, int* sray_c, void* x_enc_c, float* h16_c, float* sigma_c,
                     uint32_t ld_sigma_c, int* alive_ws, void* stream);

/* ------------------------------------------------------------------ render epilogues */

/* What scripts/export.py:78-90, scripts/render.py:61-82,104 and autolabel/evaluation.py:295-318 compute per frame
 * after model.render(), in one pass over the composited maps:
 *   label      int32 [N]   = argmax_c logits[i, c]                         (first maximal value, as torch.argmax)
 *   text_label int32 [N]   = argmax_t < feat[i] / ||feat[i]||, text[t] >     text: [T, F] fp32 (encoded class prompts)
 *   pca8       uint8 [N,3] = uint8(255 clip(((feat[i] - pca_mean) . pca_comp^T - pca_min) / pca_range, 0, 1))
 *                            pca_mean [F], pca_comp [3, F] (sklearn PCA.mean_ / components_), pca_min / pca_range [3]
 *   rgb8       uint8 [N,3] = uint8(255 image[i])                           image: [N, 3]
 * Each output (with its inputs) may be NULL.  logits rows have stride ld_logits, feature rows ld_feat (so the
 * compositing buffer [N, 3 + C + F] can be passed in place).  F <= 1024. */
int al_render_epilogue(const float* image, const float* logits, uint32_t ld_logits, const float* feat,
                       uint32_t ld_feat, uint32_t N, uint32_t C, uint32_t F, const float* text, uint32_t T,
                       const float* pca_mean, const float* pca_comp, const float* pca_min,
                       const float* pca_range, uint8_t* rgb8, int* label, int* text_label, uint8_t* pca8,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUTOLABEL_B200_H */
