// Gradient exchange + optimiser as ONE kernel over NVLink / NVSwitch peer memory (ray-sharded data-parallel training,
// SURVEY 8(e)): reduce-scatter of the parameter gradients, Adam on this rank's shard, all-gather of the updated
// parameters, without a staging copy and without NCCL.
//
// Replaces, for world > 1, the sequence  all_reduce(param.grad) -> Adam on every rank  (the reference trains on
// one GPU: scripts/train.py:50-63 builds torch.optim.Adam over encoder + network parameters; its multi-GPU form in
// torch_ngp is DistributedDataParallel, torch_ngp/nerf/utils.py:378-380).
//
// Memory model: the flat gradient buffer G and the flat parameter buffer P of every rank live in symmetric memory
// (torch.distributed._symmetric_memory), so each rank holds device pointers to all replicas -- and, when the fabric
// supports it, one MULTICAST address per buffer (NVLS):
//   multicast:  g = multimem.ld_reduce.add(G_mc + i)   the switch sums the W replicas and returns one value
//               multimem.st(P_mc + i, p')              the switch writes the new parameter into every replica
//   peer:       g = sum_k G_k[i] in rank order (plain loads through the peer mappings), stores to every replica.
// Gradients are NOT zeroed here: every rank clears its own replica with a local memset after the closing barrier
// (57 MB of HBM writes, ~10 us) instead of (W-1)/W * 57 MB of zeros over NVLink.
// Rank r owns elements [shard_begin, shard_end): it alone reads their gradients and writes their parameters, so every
// replica receives bit-identical parameters whatever the reduction order.  The Adam moments exist only for the owned
// shard (1/W of the optimiser state and of its 16 B/parameter of HBM traffic per rank).
// The caller brackets the launch with two symmetric-memory barriers: all backward passes done before, all parameter
// writes landed after.
#include "common.cuh"
#include "../../include/autolabel_b200.h"

namespace {

constexpr int kMaxWorld = 16;

struct PeerPtrs {
    float* g[kMaxWorld];
    float* p[kMaxWorld];
};

__device__ __forceinline__ float4 mc_ld_reduce(const float* addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* addr, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Same arithmetic as k_adam (optim.cu): torch.optim.Adam with L2 weight decay, bias corrections folded into
// step_size and bc2_sqrt.
__device__ __forceinline__ void adam4(float4& P, const float4& G, float4& M, float4& V, float lr_bc1, float b1, float b2,
                                      float eps, float wd, float bc2_sqrt, float gscale) {
    float* pp = &P.x; const float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        float gr = gg[k] * gscale;
        if (wd != 0.f) gr = fmaf(wd, pp[k], gr);
        mm[k] = fmaf(b1, mm[k], (1.f - b1) * gr);
        vv[k] = fmaf(b2, vv[k], (1.f - b2) * gr * gr);
        const float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
        pp[k] = pp[k] - lr_bc1 * (mm[k] / denom);
    }
}

// U float4 per thread and iteration: all gradient loads (W peer loads or one in-switch reduction each) are issued before
// the first use, so a thread keeps U * W (resp. U) NVLink round trips in flight.
template <bool MC, int U>
__global__ void __launch_bounds__(256) k_peer_adam(const PeerPtrs ptrs, float* __restrict__ mc_grad,
                                                   float* __restrict__ mc_param, const float* __restrict__ local_param,
                                                   float* __restrict__ m, float* __restrict__ v, size_t begin4,
                                                   size_t end4, size_t wd_begin4, int world, float lr_bc1, float b1,
                                                   float b2, float eps, float wd, float bc2_sqrt, float gscale) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end4; i0 += stride * U) {
        float4 G[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            G[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < end4) {
                if (MC) {
                    G[u] = mc_ld_reduce(mc_grad + i * 4);
                } else {
                    G[u] = __ldcg(reinterpret_cast<const float4*>(ptrs.g[0]) + i);
                    for (int k = 1; k < world; ++k) {
                        const float4 t = __ldcg(reinterpret_cast<const float4*>(ptrs.g[k]) + i);
                        G[u].x += t.x; G[u].y += t.y; G[u].z += t.z; G[u].w += t.w;
                    }
                }
            }
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            if (i >= end4) break;
            float4 P = __ldcg(reinterpret_cast<const float4*>(local_param) + i);   // every replica holds the same value
            float4 M = reinterpret_cast<float4*>(m)[i - begin4];
            float4 V = reinterpret_cast<float4*>(v)[i - begin4];
            adam4(P, G[u], M, V, lr_bc1, b1, b2, eps, i >= wd_begin4 ? wd : 0.f, bc2_sqrt, gscale);
            reinterpret_cast<float4*>(m)[i - begin4] = M;
            reinterpret_cast<float4*>(v)[i - begin4] = V;
            if (MC) {
                mc_st(mc_param + i * 4, P);
            } else {
                for (int k = 0; k < world; ++k) __stcg(reinterpret_cast<float4*>(ptrs.p[k]) + i, P);
            }
        }
    }
}

}  // namespace

// One optimiser step of rank `rank`'s shard [shard_begin, shard_end) of a flat parameter vector replicated on `world`
// GPUs.  grad_ptrs / param_ptrs: HOST arrays of `world` device pointers (the peer mappings of every replica's flat
// gradient / parameter buffer, this rank's own included, in rank order); mc_grad / mc_param: multicast addresses of
// the same buffers or NULL (then the peer pointers are used).  exp_avg / exp_avg_sq: this rank's moments of its shard.
// Elements >= wd_begin take `weight_decay` (the MLP parameters, scripts/train.py:57-62), the others none.
// shard_begin, shard_end and wd_begin are multiples of 4; all buffers 16-byte aligned.
AL_API int al_peer_adam_step(const void* const* grad_ptrs, const void* const* param_ptrs, float* mc_grad, float* mc_param,
                             float* exp_avg, float* exp_avg_sq, size_t shard_begin, size_t shard_end, size_t wd_begin,
                             int world, int rank, float lr, float beta1, float beta2, float eps, float weight_decay,
                             int step, float grad_scale, void* stream) {
    if (shard_end <= shard_begin) return 0;
    AL_REQUIRE(grad_ptrs && param_ptrs && exp_avg && exp_avg_sq, "null pointer");
    AL_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "bad world size / rank");
    AL_REQUIRE(step >= 1, "step must be >= 1");
    AL_REQUIRE(((shard_begin | shard_end | wd_begin) & 3) == 0, "shard bounds must be multiples of 4 elements");
    AL_REQUIRE((mc_grad == nullptr) == (mc_param == nullptr), "both multicast addresses or none");
    PeerPtrs ptrs = {};
    for (int k = 0; k < world; ++k) {
        ptrs.g[k] = (float*)grad_ptrs[k];
        ptrs.p[k] = (float*)param_ptrs[k];
        AL_REQUIRE(ptrs.g[k] && ptrs.p[k], "null peer pointer");
        AL_REQUIRE((((uintptr_t)ptrs.g[k] | (uintptr_t)ptrs.p[k]) & 15) == 0, "peer buffers must be 16-byte aligned");
    }
    AL_REQUIRE((((uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)mc_grad | (uintptr_t)mc_param) & 15) == 0,
               "buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const size_t n4 = (shard_end - shard_begin) / 4;
    const unsigned blocks = (unsigned)min((unsigned long long)al_div_up(n4, 256), (unsigned long long)al_num_sms() * 16ull);
    const float lr_bc1 = lr / (float)bc1;              // as k_adam computes it (fp32 division)
    if (mc_grad)
        k_peer_adam<true, 4><<<blocks, 256, 0, (cudaStream_t)stream>>>(ptrs, mc_grad, mc_param, ptrs.p[rank], exp_avg, exp_avg_sq,
                                                                    shard_begin / 4, shard_end / 4, wd_begin / 4, world, lr_bc1,
                                                                    beta1, beta2, eps, weight_decay, (float)sqrt(bc2), grad_scale);
    else
        k_peer_adam<false, 2><<<blocks, 256, 0, (cudaStream_t)stream>>>(ptrs, nullptr, nullptr, ptrs.p[rank], exp_avg, exp_avg_sq,
                                                                     shard_begin / 4, shard_end / 4, wd_begin / 4, world, lr_bc1,
                                                                     beta1, beta2, eps, weight_decay, (float)sqrt(bc2), grad_scale);
    AL_LAUNCH_CHECK();
    return 0;
}
