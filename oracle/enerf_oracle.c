/*
 * enerf_oracle.c — CPU restatement of the reference's hot-path arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA kernels are compared with
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference).  Nothing under
 * enerf_b200/ may call it.  It restates, in plain sequential C, what the reference's CUDA
 * kernels compute; every function cites the reference lines it follows (paths relative to the
 * reference repository knelk/enerf @ 3fb17cd).
 *
 * PARITY PIN: the reference ships no tests, fixtures or golden vectors for this path
 * (SURVEY.md §4, §8c).  The oracle is pinned instead (a) against closed forms / published
 * known answers (tests/test_oracle.py) and (b) against the reference's own CUDA build
 * (oracle/_ref, built by oracle/build_ref.py) executed on the B200: tests/test_ref_parity.py
 * and the fixtures under tests/golden/ generated from it.
 *
 * fp16 is emulated with _Float16 (round-to-nearest-even on every conversion), which is what
 * c10::Half / __half arithmetic does on the device.  Build: see oracle/Makefile.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

typedef _Float16 half_t;

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

/* ------------------------------------------------------------------------------------ */
/* pcg32 — raymarching/src/pcg32.h:57-72 (seed, next_uint), :107-116 (next_float)         */
/* ------------------------------------------------------------------------------------ */
typedef struct { uint64_t state, inc; } pcg32_t;

static uint32_t pcg32_next(pcg32_t* r) {
    uint64_t old = r->state;
    r->state = old * 0x5851f42d4c957f2dULL + r->inc;
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((~rot + 1u) & 31));
}
static void pcg32_seed(pcg32_t* r, uint64_t initstate, uint64_t initseq) {
    r->state = 0u;
    r->inc = (initseq << 1u) | 1u;
    pcg32_next(r);
    r->state += initstate;
    pcg32_next(r);
}
static float pcg32_float(pcg32_t* r) {
    union { uint32_t u; float f; } x;
    x.u = (pcg32_next(r) >> 9) | 0x3f800000u;
    return x.f - 1.0f;
}
/* exported for the known-answer test: first `n` uint32 outputs of seed(initstate, initseq) */
void oracle_pcg32_stream(uint64_t initstate, uint64_t initseq, uint32_t n, uint32_t* out) {
    pcg32_t r;
    pcg32_seed(&r, initstate, initseq);
    for (uint32_t i = 0; i < n; i++) out[i] = pcg32_next(&r);
}
float oracle_pcg32_first_float(uint64_t initstate, uint64_t initseq) {
    pcg32_t r;
    pcg32_seed(&r, initstate, initseq);
    return pcg32_float(&r);
}

/* ------------------------------------------------------------------------------------ */
/* Morton codes — raymarching/src/raymarching.cu:58-83                                    */
/* ------------------------------------------------------------------------------------ */
static uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static uint32_t morton3D_1(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static uint32_t morton3D_invert_1(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}
/* raymarching.cu:216-228 */
void oracle_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {
    for (uint32_t n = 0; n < N; n++)
        indices[n] = (int32_t)morton3D_1((uint32_t)coords[3 * n], (uint32_t)coords[3 * n + 1], (uint32_t)coords[3 * n + 2]);
}
/* raymarching.cu:239-256 */
void oracle_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) {
    for (uint32_t n = 0; n < N; n++) {
        int32_t ind = indices[n];
        coords[3 * n] = (int32_t)morton3D_invert_1((uint32_t)(ind >> 0));
        coords[3 * n + 1] = (int32_t)morton3D_invert_1((uint32_t)(ind >> 1));
        coords[3 * n + 2] = (int32_t)morton3D_invert_1((uint32_t)(ind >> 2));
    }
}
/* raymarching.cu:269-291 */
void oracle_packbits(const float* grid, uint32_t N, float thresh, uint8_t* bitfield) {
    for (uint32_t n = 0; n < N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= grid[(size_t)n * 8 + i] > thresh ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* ------------------------------------------------------------------------------------ */
/* near_far_from_aabb — raymarching/src/raymarching.cu:93-147                             */
/* ------------------------------------------------------------------------------------ */
void oracle_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                               float* nears, float* fars) {
    for (uint32_t n = 0; n < N; n++) {
        const float* o = rays_o + 3 * (size_t)n;
        const float* d = rays_d + 3 * (size_t)n;
        const float rdx = 1 / d[0], rdy = 1 / d[1], rdz = 1 / d[2];
        float near = (aabb[0] - o[0]) * rdx, far = (aabb[3] - o[0]) * rdx, t;
        if (near > far) { t = near; near = far; far = t; }
        float near_y = (aabb[1] - o[1]) * rdy, far_y = (aabb[4] - o[1]) * rdy;
        if (near_y > far_y) { t = near_y; near_y = far_y; far_y = t; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - o[2]) * rdz, far_z = (aabb[5] - o[2]) * rdz;
        if (near_z > far_z) { t = near_z; near_z = far_z; far_z = t; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* ------------------------------------------------------------------------------------ */
/* marcher — raymarching/src/raymarching.cu:44-56 (mip levels), :344-479 (train),          */
/*           :735-803 (inference).  One sequential march per ray, exactly the loop of the  */
/*           reference thread.  fmaf() marks where nvcc's default -fmad contracts.         */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max;
    uint32_t C, H;
    const uint8_t* grid;
} march_t;

static int mip_from_pos(float x, float y, float z, float max_cascade) {
    float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}
static int mip_from_dt(float dt, float H, float max_cascade) {
    float mx = (float)(dt * H * 0.5);
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}
static float signf1(float x) { return copysignf(1.0f, x); }

/* marches from t until `limit` samples were produced or t >= far; returns the sample count.
 * out_* may be NULL (counting pass).  *t_io receives the final t. */
static uint32_t march_one(const march_t* m, float t0, float far, uint32_t limit, float* xyzs, float* dirs, float* deltas) {
    const float H = (float)m->H;
    float t = t0, last_t = t0;
    uint32_t step = 0;
    while (t < far && step < limit) {
        const float x = clampf(fmaf(t, m->dx, m->ox), -m->bound, m->bound);
        const float y = clampf(fmaf(t, m->dy, m->oy), -m->bound, m->bound);
        const float z = clampf(fmaf(t, m->dz, m->oz), -m->bound, m->bound);
        const float dt = clampf(t * m->dt_gamma, m->dt_min, m->dt_max);
        int level = mip_from_pos(x, y, z, (float)m->C);
        int l2 = mip_from_dt(dt, H, (float)m->C);
        if (l2 > level) level = l2;
        const float mip_bound = fminf((float)(1 << level), m->bound);
        const float mip_rbound = 1 / mip_bound;
        const int nx = (int)clampf((float)(0.5 * (double)fmaf(x, mip_rbound, 1.0f) * (double)m->H), 0.0f, (float)(m->H - 1));
        const int ny = (int)clampf((float)(0.5 * (double)fmaf(y, mip_rbound, 1.0f) * (double)m->H), 0.0f, (float)(m->H - 1));
        const int nz = (int)clampf((float)(0.5 * (double)fmaf(z, mip_rbound, 1.0f) * (double)m->H), 0.0f, (float)(m->H - 1));
        const uint32_t index = (uint32_t)level * m->H * m->H * m->H + morton3D_1((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
        const int occ = m->grid[index / 8] & (1 << (index % 8));
        if (occ) {
            if (xyzs) {
                xyzs[3 * step] = x; xyzs[3 * step + 1] = y; xyzs[3 * step + 2] = z;
                dirs[3 * step] = m->dx; dirs[3 * step + 1] = m->dy; dirs[3 * step + 2] = m->dz;
            }
            t += dt;
            if (deltas) { deltas[2 * step] = dt; deltas[2 * step + 1] = t - last_t; }
            last_t = t;
            step++;
        } else {
            const float Hm1 = (float)(m->H - 1);
            const float tx = fmaf(fmaf(((float)nx + 0.5f + 0.5f * signf1(m->dx)) / Hm1, 2.0f, -1.0f), mip_bound, -x) * m->rdx;
            const float ty = fmaf(fmaf(((float)ny + 0.5f + 0.5f * signf1(m->dy)) / Hm1, 2.0f, -1.0f), mip_bound, -y) * m->rdy;
            const float tz = fmaf(fmaf(((float)nz + 0.5f + 0.5f * signf1(m->dz)) / Hm1, 2.0f, -1.0f), mip_bound, -z) * m->rdz;
            const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
            do { t += clampf(t * m->dt_gamma, m->dt_min, m->dt_max); } while (t < tt);
        }
    }
    return step;
}

static void march_setup(march_t* m, const float* o, const float* d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                        uint32_t H, const uint8_t* grid) {
    const float SQRT3 = 1.7320508075688772f;
    m->ox = o[0]; m->oy = o[1]; m->oz = o[2];
    m->dx = d[0]; m->dy = d[1]; m->dz = d[2];
    m->rdx = 1 / m->dx; m->rdy = 1 / m->dy; m->rdz = 1 / m->dz;
    m->bound = bound; m->dt_gamma = dt_gamma;
    m->dt_min = 2 * SQRT3 / max_steps;
    m->dt_max = 2 * SQRT3 * (1 << (C - 1)) / H;
    m->C = C; m->H = H; m->grid = grid;
}

/* raymarching.cu:313-480.  Ranges are reserved in ray order (the reference's atomics make the
 * order arbitrary; any order is a valid result, tests compare per ray through rays[:,0]). */
void oracle_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                             uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                             const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                             uint32_t perturb) {
    for (uint32_t n = 0; n < N; n++) {
        march_t m;
        march_setup(&m, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H, grid);
        float t0 = nears[n];
        if (perturb) {
            pcg32_t rng;
            pcg32_seed(&rng, (uint64_t)n, 1u);
            t0 = fmaf(m.dt_min, pcg32_float(&rng), t0);   /* nvcc contracts `t0 += dt_min * r` (verified against the reference build) */
        }
        const uint32_t num_steps = march_one(&m, t0, fars[n], max_steps, NULL, NULL, NULL);
        const uint32_t point_index = (uint32_t)counter[0];
        counter[0] += (int32_t)num_steps;
        const uint32_t ray_index = (uint32_t)counter[1];
        counter[1] += 1;
        rays[3 * ray_index] = (int32_t)n;
        rays[3 * ray_index + 1] = (int32_t)point_index;
        rays[3 * ray_index + 2] = (int32_t)num_steps;
        if (num_steps == 0) continue;
        if (point_index + num_steps >= M) continue;
        march_one(&m, t0, fars[n], num_steps, xyzs + 3 * (size_t)point_index, dirs + 3 * (size_t)point_index,
                  deltas + 2 * (size_t)point_index);
    }
}

/* raymarching.cu:700-804 */
void oracle_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                       const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                       const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                       uint32_t perturb) {
    (void)nears;
    for (uint32_t n = 0; n < n_alive; n++) {
        const int index = rays_alive[n];
        march_t m;
        march_setup(&m, rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, bound, dt_gamma, max_steps, C, H, grid);
        float t = rays_t[n];
        if (perturb) {
            pcg32_t rng;
            pcg32_seed(&rng, (uint64_t)n, (uint64_t)perturb);
            t = fmaf(m.dt_min, pcg32_float(&rng), t);
        }
        march_one(&m, t, fars[index], n_step, xyzs + 3 * (size_t)n * n_step, dirs + 3 * (size_t)n * n_step, deltas + 2 * (size_t)n * n_step);
    }
}

/* ------------------------------------------------------------------------------------ */
/* compositing                                                                            */
/* ------------------------------------------------------------------------------------ */
/* raymarching.cu:500-578 (n_ch = 3 there) */
void oracle_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                         uint32_t M, uint32_t N, uint32_t n_ch, float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], num_steps = (uint32_t)rays[3 * n + 2];
        if (num_steps == 0 || offset + num_steps >= M) {
            weights_sum[index] = 0; depth[index] = 0;
            for (uint32_t c = 0; c < n_ch; c++) image[index * n_ch + c] = 0;
            continue;
        }
        float T = 1.0f, ws = 0, t = 0, d = 0, col[8] = {0};
        for (uint32_t s = 0; s < num_steps; s++) {
            const size_t i = (size_t)offset + s;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[2 * i]);
            const float w = alpha * T;
            for (uint32_t c = 0; c < n_ch; c++) col[c] += w * rgbs[i * n_ch + c];
            t += deltas[2 * i + 1];
            d += w * t;
            ws += w;
            T *= 1.0f - alpha;
        }
        weights_sum[index] = ws; depth[index] = d;
        for (uint32_t c = 0; c < n_ch; c++) image[index * n_ch + c] = col[c];
    }
}

/* raymarching.cu:602-682 */
void oracle_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                          const float* rgbs, const float* deltas, const int32_t* rays, const float* weights_sum,
                                          const float* image, uint32_t M, uint32_t N, uint32_t n_ch, float* grad_sigmas,
                                          float* grad_rgbs) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], num_steps = (uint32_t)rays[3 * n + 2];
        if (num_steps == 0 || offset + num_steps >= M) continue;
        float T = 1.0f, ws = 0, col[8] = {0};
        const float ws_final = weights_sum[index];
        for (uint32_t s = 0; s < num_steps; s++) {
            const size_t i = (size_t)offset + s;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[2 * i]);
            const float w = alpha * T;
            for (uint32_t c = 0; c < n_ch; c++) col[c] += w * rgbs[i * n_ch + c];
            ws += w;
            T *= 1.0f - alpha;
            float acc = 0;
            for (uint32_t c = 0; c < n_ch; c++) {
                grad_rgbs[i * n_ch + c] = grad_image[index * n_ch + c] * w;
                acc += grad_image[index * n_ch + c] * (T * rgbs[i * n_ch + c] - (image[index * n_ch + c] - col[c]));
            }
            acc += grad_weights_sum[index] * (T - (ws_final - ws));
            grad_sigmas[i] = deltas[2 * i] * acc;
        }
    }
}

/* raymarching.cu:816-900 */
void oracle_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t, const float* sigmas,
                           const float* rgbs, const float* deltas, uint32_t n_ch, float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int index = rays_alive[n];
        float t = rays_t[n];
        const float* sg = sigmas + (size_t)n * n_step;
        const float* cl = rgbs + (size_t)n * n_step * n_ch;
        const float* dl = deltas + (size_t)n * n_step * 2;
        float weight_sum = weights_sum[index], d = depth[index], col[8];
        for (uint32_t c = 0; c < n_ch; c++) col[c] = image[index * n_ch + c];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = 1.0f - expf(-sg[0] * dl[0]);
            const float T = 1 - weight_sum;
            const float w = alpha * T;
            weight_sum += w;
            t += dl[1];
            d += w * t;
            for (uint32_t c = 0; c < n_ch; c++) col[c] += w * cl[c];
            if (T < 1e-5) break;
            sg++; cl += n_ch; dl += 2; step++;
        }
        rays_t[n] = step < n_step ? -1.0f : t;
        weights_sum[index] = weight_sum; depth[index] = d;
        for (uint32_t c = 0; c < n_ch; c++) image[index * n_ch + c] = col[c];
    }
}

/* raymarching.cu:912-930 (slot order = input order; the reference's atomic order is arbitrary) */
void oracle_compact_rays(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t,
                         const float* rays_t_old, int32_t* alive_counter) {
    for (uint32_t n = 0; n < n_alive; n++) {
        if (rays_t_old[n] >= 0) {
            const int index = alive_counter[0]++;
            rays_alive[index] = rays_alive_old[n];
            rays_t[index] = rays_t_old[n];
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* hash grid — gridencoder/src/gridencoder.cu:34-71 (index), :96-136 (range check, pos),  */
/*             :143-168 (blend), :178-220 (dy_dx), :250-310 (backward), :314-340 (input)  */
/* ------------------------------------------------------------------------------------ */
static uint32_t grid_index(uint32_t gridtype, uint32_t D, uint32_t Cc, uint32_t ch, uint32_t hashmap_size, uint32_t resolution,
                           const uint32_t* pos_grid) {
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) {
        index = 0;
        for (uint32_t d = 0; d < D; d++) index ^= pos_grid[d] * primes[d];
    }
    return (index % hashmap_size) * Cc + ch;
}

#define GRID_IMPL(NAME, T, ACC_STMT, SUB_EXPR, ZERO)                                                                              \
    void NAME(const float* inputs, const T* grid_all, const int32_t* offsets, T* outputs, uint32_t B, uint32_t D, uint32_t Cc,    \
              uint32_t L, float S, uint32_t H, int calc_grad_inputs, T* dy_dx_all, uint32_t gridtype, const float* level_scales) { \
        _Pragma("omp parallel for schedule(static)") for (int64_t bb = 0; bb < (int64_t)B; bb++) {                                \
            const uint32_t b = (uint32_t)bb;                                                                                      \
            const float* x = inputs + (size_t)b * D;                                                                              \
            int oob = 0;                                                                                                          \
            for (uint32_t d = 0; d < D; d++)                                                                                      \
                if (x[d] < 0 || x[d] > 1) oob = 1;                                                                                \
            for (uint32_t level = 0; level < L; level++) {                                                                        \
                const T* grid = grid_all + (size_t)(uint32_t)offsets[level] * Cc;                                                 \
                T* out = outputs + ((size_t)level * B + b) * Cc;                                                                  \
                T* dy_dx = dy_dx_all + ((size_t)b * L + level) * D * Cc;                                                          \
                if (oob) {                                                                                                        \
                    for (uint32_t ch = 0; ch < Cc; ch++) out[ch] = ZERO;                                                          \
                    if (calc_grad_inputs)                                                                                         \
                        for (uint32_t i = 0; i < D * Cc; i++) dy_dx[i] = ZERO;                                                    \
                    continue;                                                                                                     \
                }                                                                                                                 \
                const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);                                    \
                const float scale = level_scales ? level_scales[level] : exp2f(level * S) * H - 1.0f;                             \
                const uint32_t resolution = (uint32_t)ceilf(scale) + 1;                                                           \
                float pos[3];                                                                                                     \
                uint32_t pos_grid[3];                                                                                             \
                for (uint32_t d = 0; d < D; d++) {                                                                                \
                    pos[d] = fmaf(x[d], scale, 0.5f);                                                                             \
                    pos_grid[d] = (uint32_t)floorf(pos[d]);                                                                       \
                    pos[d] -= (float)pos_grid[d];                                                                                 \
                }                                                                                                                 \
                T results[8];                                                                                                     \
                for (uint32_t ch = 0; ch < Cc; ch++) results[ch] = ZERO;                                                          \
                for (uint32_t idx = 0; idx < (1u << D); idx++) {                                                                  \
                    float w = 1;                                                                                                  \
                    uint32_t pl[3];                                                                                               \
                    for (uint32_t d = 0; d < D; d++) {                                                                            \
                        if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pos_grid[d]; }                                     \
                        else { w *= pos[d]; pl[d] = pos_grid[d] + 1; }                                                            \
                    }                                                                                                             \
                    const uint32_t index = grid_index(gridtype, D, Cc, 0, hashmap_size, resolution, pl);                          \
                    for (uint32_t ch = 0; ch < Cc; ch++) { const T gval = grid[index + ch]; T* r = &results[ch]; ACC_STMT; }      \
                }                                                                                                                 \
                for (uint32_t ch = 0; ch < Cc; ch++) out[ch] = results[ch];                                                       \
                if (calc_grad_inputs) {                                                                                           \
                    for (uint32_t gd = 0; gd < D; gd++) {                                                                         \
                        T rg[8];                                                                                                  \
                        for (uint32_t ch = 0; ch < Cc; ch++) rg[ch] = ZERO;                                                       \
                        for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {                                                    \
                            float w = scale;                                                                                      \
                            uint32_t pl[3];                                                                                       \
                            for (uint32_t nd = 0; nd < D - 1; nd++) {                                                             \
                                const uint32_t d = (nd >= gd) ? (nd + 1) : nd;                                                    \
                                if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pl[d] = pos_grid[d]; }                            \
                                else { w *= pos[d]; pl[d] = pos_grid[d] + 1; }                                                    \
                            }                                                                                                     \
                            pl[gd] = pos_grid[gd];                                                                                \
                            const uint32_t il = grid_index(gridtype, D, Cc, 0, hashmap_size, resolution, pl);                     \
                            pl[gd] = pos_grid[gd] + 1;                                                                            \
                            const uint32_t ir = grid_index(gridtype, D, Cc, 0, hashmap_size, resolution, pl);                     \
                            for (uint32_t ch = 0; ch < Cc; ch++) {                                                                \
                                const T a_ = grid[ir + ch], b_ = grid[il + ch];                                                   \
                                const T gval = SUB_EXPR;                                                                          \
                                T* r = &rg[ch];                                                                                   \
                                ACC_STMT;                                                                                         \
                            }                                                                                                     \
                        }                                                                                                         \
                        for (uint32_t ch = 0; ch < Cc; ch++) dy_dx[gd * Cc + ch] = rg[ch];                                        \
                    }                                                                                                             \
                }                                                                                                                 \
            }                                                                                                                     \
        }                                                                                                                         \
    }

/* float table: `results += w * grid` contracts to one fma */
GRID_IMPL(oracle_grid_encode_forward_f32, float, *r = fmaf(w, gval, *r), (a_ - b_), 0.0f)
/* half table (c10::Half): product rounded to half, sum computed in fp32 and rounded to half */
GRID_IMPL(oracle_grid_encode_forward_f16, half_t, *r = (half_t)((float)*r + (float)(half_t)(w * (float)gval)),
          ((half_t)((float)a_ - (float)b_)), (half_t)0.0f)

/* backward: grad [L,B,C] (reference layout) -> grad_grid.  Accumulation in double is the
 * order-independent "true" sum the atomics approximate; `mode` 0: fp32 products (fp32 table),
 * 1: each product rounded to half first (gridencoder.cu:300, half table). */
void oracle_grid_encode_backward(const float* grad, const float* inputs, const int32_t* offsets, double* grad_grid, uint32_t B,
                                 uint32_t D, uint32_t Cc, uint32_t L, float S, uint32_t H, uint32_t gridtype, int mode,
                                 const float* level_scales) {
    /* levels own disjoint slices of grad_grid -> race-free across threads */
    _Pragma("omp parallel for schedule(dynamic, 1)") for (int64_t lv = 0; lv < (int64_t)L; lv++) {
        const uint32_t level = (uint32_t)lv;
        double* gg = grad_grid + (size_t)(uint32_t)offsets[level] * Cc;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const float scale = level_scales ? level_scales[level] : exp2f(level * S) * H - 1.0f;
        const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
        for (uint32_t b = 0; b < B; b++) {
            const float* x = inputs + (size_t)b * D;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++)
                if (x[d] < 0 || x[d] > 1) oob = 1;
            if (oob) continue;
            float pos[3];
            uint32_t pos_grid[3];
            for (uint32_t d = 0; d < D; d++) {
                pos[d] = fmaf(x[d], scale, 0.5f);
                pos_grid[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pos_grid[d];
            }
            const float* g = grad + ((size_t)level * B + b) * Cc;
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                uint32_t pl[3];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pl[d] = pos_grid[d] + 1; }
                }
                const uint32_t index = grid_index(gridtype, D, Cc, 0, hashmap_size, resolution, pl);
                for (uint32_t ch = 0; ch < Cc; ch++) {
                    const float p = w * g[ch];
                    gg[index + ch] += mode ? (double)(float)(half_t)p : (double)p;
                }
            }
        }
    }
}

/* gridencoder.cu:314-340, fp32 */
void oracle_grid_input_backward_f32(const float* grad, const float* dy_dx, float* grad_inputs, uint32_t B, uint32_t D, uint32_t Cc,
                                    uint32_t L) {
    for (uint32_t t = 0; t < B * D; t++) {
        const uint32_t b = t / D, d = t - b * D;
        const float* dd = dy_dx + (size_t)b * L * D * Cc;
        float result = 0;
        for (uint32_t l = 0; l < L; l++)
            for (uint32_t ch = 0; ch < Cc; ch++) result = fmaf(grad[((size_t)l * B + b) * Cc + ch], dd[(l * D + d) * Cc + ch], result);
        grad_inputs[t] = result;
    }
}

/* helpers so Python can move fp16 buffers through the oracle */
void oracle_f32_to_f16(const float* src, uint16_t* dst, size_t n) {
    for (size_t i = 0; i < n; i++) { half_t h = (half_t)src[i]; memcpy(&dst[i], &h, 2); }
}
void oracle_f16_to_f32(const uint16_t* src, float* dst, size_t n) {
    for (size_t i = 0; i < n; i++) { half_t h; memcpy(&h, &src[i], 2); dst[i] = (float)h; }
}

/* ------------------------------------------------------------------------------------ */
/* spherical harmonics, degree <= 4 — shencoder/src/shencoder.cu:51-69 (closed forms)     */
/* ------------------------------------------------------------------------------------ */
void oracle_sh_encode_forward_f32(const float* inputs, float* outputs, uint32_t B, uint32_t C) {
    const uint32_t C2 = C * C;
    for (uint32_t b = 0; b < B; b++) {
        const float x = inputs[3 * b], y = inputs[3 * b + 1], z = inputs[3 * b + 2];
        const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
        float* o = outputs + (size_t)b * C2;
        o[0] = 0.28209479177387814f;
        if (C <= 1) continue;
        o[1] = -0.48860251190291987f * y;
        o[2] = 0.48860251190291987f * z;
        o[3] = -0.48860251190291987f * x;
        if (C <= 2) continue;
        o[4] = 1.0925484305920792f * xy;
        o[5] = -1.0925484305920792f * yz;
        o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
        o[7] = -1.0925484305920792f * xz;
        o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
        if (C <= 3) continue;
        o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
        o[10] = 2.8906114426405538f * xy * z;
        o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
        o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
        o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
        o[14] = 1.4453057213202769f * z * (x2 - y2);
        o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
    }
}
