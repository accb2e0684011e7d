/* oracle/ngp_oracle.c — CPU restatement of the reference's ray-marching / compositing / hash-grid
 * algorithms.  TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the checker; never imported by the product (autolabel_b200/).
 *
 * Pinning: this file is checked (tests/test_oracle_pinned.py) against tests/golden/ref_*.npz,
 * which hold outputs of the reference's OWN kernels (oracle/_ref, compiled unmodified from
 * /root/reference by oracle/build_ref.py) run on a B200 by tests/golden/make_golden.py.
 *
 * Every function names the reference lines it follows.  Scalar, single-threaded C; float
 * arithmetic is written with explicit fmaf() in exactly the fused form the reference binary
 * executes (read off its SASS: FFMA for o + t*d, x*rbound + 1, the voxel-exit expression, ...),
 * and this file must be compiled with -ffp-contract=off so the compiler adds no others.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ---------------------------------------------------------------- pcg32 (raymarching/src/pcg32.h:57-116,149-170) */
typedef struct { uint64_t state, inc; } pcg32_t;
#define PCG32_MULT 0x5851f42d4c957f2dULL

static uint32_t pcg_next_uint(pcg32_t* r) {
    uint64_t old = r->state;
    r->state = old * PCG32_MULT + r->inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
static void pcg_seed(pcg32_t* r, uint64_t initstate, uint64_t initseq) {
    r->state = 0u;
    r->inc = (initseq << 1u) | 1u;
    pcg_next_uint(r);
    r->state += initstate;
    pcg_next_uint(r);
}
static void pcg_advance(pcg32_t* r, int64_t delta_) {
    uint64_t cur_mult = PCG32_MULT, cur_plus = r->inc, acc_mult = 1u, acc_plus = 0u;
    uint64_t delta = (uint64_t)delta_;
    while (delta > 0) {
        if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta /= 2;
    }
    r->state = acc_mult * r->state + acc_plus;
}
static float pcg_next_float(pcg32_t* r) {
    union { uint32_t u; float f; } x;
    x.u = (pcg_next_uint(r) >> 9) | 0x3f800000u;
    return x.f - 1.0f;
}

/* ---------------------------------------------------------------- helpers (raymarching.cu:32-88) */
static float signf_(float x) { return copysignf(1.0f, x); }
static float clampf_(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

static uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static uint32_t morton3D_(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static uint32_t morton3D_invert_(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}
static int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}
static int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}

/* ---------------------------------------------------------------- raymarching.cu:98-188 */
void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                            float min_near, float* nears, float* fars, uint8_t* near_indices,
                            uint8_t* far_indices) {
    for (uint32_t n = 0; n < N; ++n) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx;
        uint8_t near_idx = 0, far_idx = 3, t8;
        float tf;
        if (near > far) { tf = near; near = far; far = tf; t8 = near_idx; near_idx = far_idx; far_idx = t8; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        uint8_t near_y_idx = 1, far_y_idx = 4;
        if (near_y > far_y) { tf = near_y; near_y = far_y; far_y = tf; t8 = near_y_idx; near_y_idx = far_y_idx; far_y_idx = t8; }
        if (near > far_y || near_y > far) {
            nears[n] = fars[n] = FLT_MAX;
            if (near_indices) near_indices[n] = 255;
            if (far_indices) far_indices[n] = 255;
            continue;
        }
        if (near_y > near) { near = near_y; near_idx = near_y_idx; }
        if (far_y < far) { far = far_y; far_idx = far_y_idx; }
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        uint8_t near_z_idx = 2, far_z_idx = 5;
        if (near_z > far_z) { tf = near_z; near_z = far_z; far_z = tf; t8 = near_z_idx; near_z_idx = far_z_idx; far_z_idx = t8; }
        if (near > far_z || near_z > far) {
            nears[n] = fars[n] = FLT_MAX;
            if (near_indices) near_indices[n] = 255;
            if (far_indices) far_indices[n] = 255;
            continue;
        }
        if (near_z > near) { near = near_z; near_idx = near_z_idx; }
        if (far_z < far) { far = far_z; far_idx = far_z_idx; }
        if (near < min_near) near = min_near;
        nears[n] = near; fars[n] = far;
        if (near_indices) near_indices[n] = near_idx;
        if (far_indices) far_indices[n] = far_idx;
    }
}

/* raymarching.cu:257-303 */
void orc_morton3D(const int* coords, uint32_t N, int* indices) {
    for (uint32_t n = 0; n < N; ++n)
        indices[n] = (int)morton3D_((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}
void orc_morton3D_invert(const int* indices, uint32_t N, int* coords) {
    for (uint32_t n = 0; n < N; ++n) {
        const int ind = indices[n];
        coords[n * 3] = (int)morton3D_invert_((uint32_t)ind >> 0);
        coords[n * 3 + 1] = (int)morton3D_invert_((uint32_t)ind >> 1);
        coords[n * 3 + 2] = (int)morton3D_invert_((uint32_t)ind >> 2);
    }
}

/* raymarching.cu:310-343 */
void orc_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield) {
    for (uint32_t n = 0; n < N; ++n) {
        uint8_t bits = 0;
        for (uint8_t i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > density_thresh) ? ((uint8_t)1 << i) : 0;
        bitfield[n] = bits;
    }
}

/* One DDA evaluation at parameter t (raymarching.cu:403-442): returns 1 and dt if the cell is
 * occupied, else advances *t past the voxel. */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH;
    uint32_t C, H;
    const uint8_t* grid;
} dda_t;

static int dda_eval(const dda_t* a, float* t, float* px, float* py, float* pz, float* pdt) {
    const float x = clampf_(fmaf(*t, a->dx, a->ox), -a->bound, a->bound);
    const float y = clampf_(fmaf(*t, a->dy, a->oy), -a->bound, a->bound);
    const float z = clampf_(fmaf(*t, a->dz, a->oz), -a->bound, a->bound);
    const float dt = clampf_(*t * a->dt_gamma, a->dt_min, a->dt_max);
    const int l1 = mip_from_pos(x, y, z, (float)a->C), l2 = mip_from_dt(dt, (float)a->H, (float)a->C);
    const int level = l1 > l2 ? l1 : l2;
    const float mip_bound = fminf((float)(1 << level), a->bound);
    const float mip_rbound = 1 / mip_bound;
    const int nx = (int)clampf_((float)(0.5 * (double)fmaf(x, mip_rbound, 1.0f) * (double)a->H), 0.0f, (float)(a->H - 1));
    const int ny = (int)clampf_((float)(0.5 * (double)fmaf(y, mip_rbound, 1.0f) * (double)a->H), 0.0f, (float)(a->H - 1));
    const int nz = (int)clampf_((float)(0.5 * (double)fmaf(z, mip_rbound, 1.0f) * (double)a->H), 0.0f, (float)(a->H - 1));
    const uint32_t index = (uint32_t)level * a->H * a->H * a->H + morton3D_((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    const int occ = a->grid[index / 8] & (1 << (index % 8));
    *px = x; *py = y; *pz = z; *pdt = dt;
    if (occ) return 1;
    const float tx = (fmaf(fmaf((fmaf(0.5f, signf_(a->dx), (float)nx + 0.5f)) * a->rH, 2.0f, -1.0f), mip_bound, -x)) * a->rdx;
    const float ty = (fmaf(fmaf((fmaf(0.5f, signf_(a->dy), (float)ny + 0.5f)) * a->rH, 2.0f, -1.0f), mip_bound, -y)) * a->rdy;
    const float tz = (fmaf(fmaf((fmaf(0.5f, signf_(a->dz), (float)nz + 0.5f)) * a->rH, 2.0f, -1.0f), mip_bound, -z)) * a->rdz;
    const float tt = *t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do {
        *t += clampf_(*t * a->dt_gamma, a->dt_min, a->dt_max);
    } while (*t < tt);
    return 0;
}

static void dda_setup(dda_t* a, const float* rays_o, const float* rays_d, uint32_t n, const uint8_t* grid,
                      float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    a->ox = rays_o[n * 3]; a->oy = rays_o[n * 3 + 1]; a->oz = rays_o[n * 3 + 2];
    a->dx = rays_d[n * 3]; a->dy = rays_d[n * 3 + 1]; a->dz = rays_d[n * 3 + 2];
    a->rdx = 1 / a->dx; a->rdy = 1 / a->dy; a->rdz = 1 / a->dz;
    a->bound = bound; a->dt_gamma = dt_gamma;
    const float two_sqrt3 = 2 * 1.7320508075688772f;
    a->dt_min = two_sqrt3 / (float)max_steps;                  /* raymarching.cu:386 */
    a->dt_max = two_sqrt3 * (float)(1 << (C - 1)) / (float)H;  /* raymarching.cu:387 */
    a->rH = 1 / (float)H;
    a->C = C; a->H = H; a->grid = grid;
}

/* raymarching.cu:354-537.  Sequential in ray order: segment offsets are the exclusive scan of the
 * counts (one valid outcome of the reference's atomicAdd order).  Buffers are NOT cleared here;
 * the caller zero-fills like raymarching.py:329-334. */
void orc_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                          float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                          const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                          float* ts, int* rays, int* counter, uint32_t perturb) {
    pcg32_t rng0;
    pcg_seed(&rng0, 42, 1);
    for (uint32_t n = 0; n < N; ++n) {
        dda_t a;
        dda_setup(&a, rays_o, rays_d, n, grid, bound, dt_gamma, max_steps, C, H);
        const float far = fars[n];
        float t0 = nears[n];
        if (perturb) {
            pcg32_t rng = rng0;
            pcg_advance(&rng, (int64_t)n);
            t0 = fmaf(a.dt_min, pcg_next_float(&rng), t0);
        }
        float t = t0, x, y, z, dt;
        uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps) {
            if (dda_eval(&a, &t, &x, &y, &z, &dt)) { num_steps++; t += dt; }
        }
        const uint32_t point_index = (uint32_t)counter[0];
        counter[0] += (int)num_steps;
        const uint32_t ray_index = (uint32_t)counter[1];
        counter[1] += 1;
        rays[ray_index * 3] = (int)n;
        rays[ray_index * 3 + 1] = (int)point_index;
        rays[ray_index * 3 + 2] = (int)num_steps;
        if (num_steps == 0) continue;
        if (point_index + num_steps >= M) continue;
        float* px = xyzs + (size_t)point_index * 3;
        float* pd = dirs ? dirs + (size_t)point_index * 3 : NULL;
        float* pl = deltas + (size_t)point_index * 2;
        float* pt = ts ? ts + point_index : NULL;
        t = t0;
        uint32_t step = 0;
        float last_t = t;
        while (t < far && step < num_steps) {
            if (dda_eval(&a, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                if (pd) { pd[0] = a.dx; pd[1] = a.dy; pd[2] = a.dz; pd += 3; }
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                if (pt) { pt[0] = t; pt++; }
                last_t = t;
                px += 3; pl += 2;
                step++;
            }
        }
    }
}

/* raymarching.cu:547-625 (K value channels instead of 3; K = 3 is the reference kernel). */
void orc_composite_rays_train_forward(const float* sigmas, const float* vals, uint32_t K, const float* deltas,
                                      const int* rays, uint32_t M, uint32_t N, float* weights_sum,
                                      float* depth, float* image) {
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        float* im = image + (size_t)index * K;
        for (uint32_t c = 0; c < K; ++c) im[c] = 0;
        weights_sum[index] = 0; depth[index] = 0;
        if (num_steps == 0 || offset + num_steps >= M) continue;
        float T = 1.0f, ws = 0, t = 0, d = 0;
        for (uint32_t s = 0; s < num_steps; ++s) {
            const size_t i = (size_t)offset + s;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[i * 2]);
            const float weight = alpha * T;
            for (uint32_t c = 0; c < K; ++c) im[c] += weight * vals[i * K + c];
            t += deltas[i * 2 + 1];
            d += weight * t;
            ws += weight;
            T *= 1.0f - alpha;
        }
        weights_sum[index] = ws; depth[index] = d;
    }
}

/* raymarching.cu:649-729 (K channels; no depth gradient, exactly like the reference). */
void orc_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                       const float* vals, uint32_t K, const float* deltas, const int* rays,
                                       const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                       float* grad_sigmas, float* grad_vals) {
    float* run = (float*)malloc(sizeof(float) * (K ? K : 1));
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps >= M) continue;
        const float* gi = grad_image + (size_t)index * K;
        const float* fin = image + (size_t)index * K;
        const float ws_final = weights_sum[index], gws = grad_weights_sum[index];
        float T = 1.0f, ws = 0;
        for (uint32_t c = 0; c < K; ++c) run[c] = 0;
        for (uint32_t s = 0; s < num_steps; ++s) {
            const size_t i = (size_t)offset + s;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[i * 2]);
            const float weight = alpha * T;
            for (uint32_t c = 0; c < K; ++c) run[c] += weight * vals[i * K + c];
            ws += weight;
            T *= 1.0f - alpha;
            float acc = 0;
            for (uint32_t c = 0; c < K; ++c) {
                grad_vals[i * K + c] = gi[c] * weight;
                acc += gi[c] * (T * vals[i * K + c] - (fin[c] - run[c]));
            }
            grad_sigmas[i] = deltas[i * 2] * (acc + gws * (T - (ws_final - ws)));
        }
    }
    free(run);
}

/* raymarching.cu:747-854 */
void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, const float* rays_t,
                    const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                    uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars,
                    float* xyzs, float* dirs, float* deltas, uint32_t perturb) {
    pcg32_t rng0;
    pcg_seed(&rng0, (uint64_t)perturb, 1);
    (void)nears;
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int index = rays_alive[n];
        float t = rays_t[n];
        dda_t a;
        dda_setup(&a, rays_o, rays_d, (uint32_t)index, grid, bound, dt_gamma, max_steps, C, H);
        const float far = fars[index];
        if (perturb) {
            pcg32_t rng = rng0;
            pcg_advance(&rng, (int64_t)n);
            t = fmaf(a.dt_min, pcg_next_float(&rng), t);
        }
        float last_t = t, x, y, z, dt;
        uint32_t step = 0;
        size_t i = (size_t)n * n_step;
        while (t < far && step < n_step) {
            if (dda_eval(&a, &t, &x, &y, &z, &dt)) {
                xyzs[i * 3] = x; xyzs[i * 3 + 1] = y; xyzs[i * 3 + 2] = z;
                if (dirs) { dirs[i * 3] = a.dx; dirs[i * 3 + 1] = a.dy; dirs[i * 3 + 2] = a.dz; }
                t += dt;
                deltas[i * 2] = dt; deltas[i * 2 + 1] = t - last_t;
                last_t = t;
                ++i; ++step;
            }
        }
    }
}

/* raymarching.cu:868-952 (K channels) */
void orc_composite_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, float* rays_t,
                        const float* sigmas, const float* vals, uint32_t K, const float* deltas,
                        float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int index = rays_alive[n];
        float t = rays_t[n];
        float weight_sum = weights_sum[index], d = depth[index];
        float* im = image + (size_t)index * K;
        uint32_t step = 0;
        while (step < n_step) {
            const size_t i = (size_t)n * n_step + step;
            if (deltas[i * 2] == 0) break;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[i * 2]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t += deltas[i * 2 + 1];
            d += weight * t;
            for (uint32_t c = 0; c < K; ++c) im[c] += weight * vals[i * K + c];
            if (T < 1e-4f) break;
            step++;
        }
        rays_t[n] = (step < n_step) ? -1.0f : t;
        weights_sum[index] = weight_sum;
        depth[index] = d;
    }
}

/* raymarching.cu:964-982 (sequential: order preserved) */
void orc_compact_rays(uint32_t n_alive, int* rays_alive, const int* rays_alive_old, float* rays_t,
                      const float* rays_t_old, int* alive_counter) {
    for (uint32_t n = 0; n < n_alive; ++n) {
        if (rays_t_old[n] >= 0) {
            const int index = alive_counter[0]++;
            rays_alive[index] = rays_alive_old[n];
            rays_t[index] = rays_t_old[n];
        }
    }
}

/* ---------------------------------------------------------------- gridencoder.cu:35-72 */
static uint32_t fast_hash(const uint32_t* pos_grid, uint32_t D) {
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t result = 0;
    for (uint32_t i = 0; i < D; ++i) result ^= pos_grid[i] * primes[i];
    return result;
}
static uint32_t get_grid_index(uint32_t D, uint32_t C, uint32_t gridtype, uint32_t ch, uint32_t hashmap_size,
                               uint32_t resolution, const uint32_t* pos_grid) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) index = fast_hash(pos_grid, D);
    return (index % hashmap_size) * C + ch;
}

/* gridencoder.cu:75-175 (forward values + the entry index of every corner).
 * outputs [L,B,C]; indices (optional) [B,L,2^D] entry indices (without the channel factor), -1 if
 * the sample is out of [0,1].  `level_scales` (optional, [L]) overrides exp2f(level*S)*H-1 so a
 * caller can inject the GPU's exp2f values when S is not an integer. */
void orc_grid_encode_forward(const float* inputs, const float* grid, const int* offsets, float* outputs,
                             uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             uint32_t gridtype, int* indices, const float* level_scales) {
    for (uint32_t level = 0; level < L; ++level) {
        const float* g = grid + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const float scale = level_scales ? level_scales[level] : exp2f(level * S) * H - 1.0f;
        const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
        for (uint32_t b = 0; b < B; ++b) {
            const float* in = inputs + (size_t)b * D;
            float* out = outputs + ((size_t)level * B + b) * C;
            int oob = 0;
            for (uint32_t d = 0; d < D; ++d) if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) {
                for (uint32_t ch = 0; ch < C; ++ch) out[ch] = 0;
                if (indices) for (uint32_t i = 0; i < (1u << D); ++i) indices[((size_t)b * L + level) * (1u << D) + i] = -1;
                continue;
            }
            float pos[3];
            uint32_t pos_grid[3];
            for (uint32_t d = 0; d < D; ++d) {
                pos[d] = fmaf(in[d], scale, 0.5f);
                pos_grid[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pos_grid[d];
            }
            float results[8] = {0};
            for (uint32_t idx = 0; idx < (1u << D); ++idx) {
                float w = 1;
                uint32_t pl[3];
                for (uint32_t d = 0; d < D; ++d) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pl[d] = pos_grid[d] + 1; }
                }
                const uint32_t index = get_grid_index(D, C, gridtype, 0, hashmap_size, resolution, pl);
                if (indices) indices[((size_t)b * L + level) * (1u << D) + idx] = (int)(index / C);
                for (uint32_t ch = 0; ch < C; ++ch) results[ch] = fmaf(w, g[index + ch], results[ch]);
            }
            for (uint32_t ch = 0; ch < C; ++ch) out[ch] = results[ch];
        }
    }
}

/* gridencoder.cu:226-312: grad [L,B,C] scattered into grad_grid (+=), double accumulation so the
 * result is order independent (the reference's float atomics are order dependent). */
void orc_grid_encode_backward(const float* grad, const float* inputs, const int* offsets, double* grad_grid,
                              uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                              uint32_t gridtype, const float* level_scales) {
    for (uint32_t level = 0; level < L; ++level) {
        double* gg = grad_grid + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const float scale = level_scales ? level_scales[level] : exp2f(level * S) * H - 1.0f;
        const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
        for (uint32_t b = 0; b < B; ++b) {
            const float* in = inputs + (size_t)b * D;
            int oob = 0;
            for (uint32_t d = 0; d < D; ++d) if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) continue;
            float pos[3];
            uint32_t pos_grid[3];
            for (uint32_t d = 0; d < D; ++d) {
                pos[d] = fmaf(in[d], scale, 0.5f);
                pos_grid[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pos_grid[d];
            }
            for (uint32_t idx = 0; idx < (1u << D); ++idx) {
                float w = 1;
                uint32_t pl[3];
                for (uint32_t d = 0; d < D; ++d) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pl[d] = pos_grid[d] + 1; }
                }
                const uint32_t index = get_grid_index(D, C, gridtype, 0, hashmap_size, resolution, pl);
                for (uint32_t ch = 0; ch < C; ++ch)
                    gg[index + ch] += (double)(w * grad[((size_t)level * B + b) * C + ch]);
            }
        }
    }
}
