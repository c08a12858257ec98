// Occupancy-grid ray marcher + alpha-composite integrator for sm_100a.
//
// Replaces the 11 entry points of the reference's `_raymarching` extension
// (raymarching/src/raymarching.h:7-18).  Results follow raymarching/src/raymarching.cu; the
// execution model does not:
//   * march kernels are WARP-PER-RAY.  Both branches of the reference loop advance t by the
//     same function of t only (`t += clamp(t*dt_gamma, dt_min, dt_max)`, raymarching.cu:387,397),
//     so the candidate sequence t_k is independent of the grid.  A warp generates 32
//     consecutive candidates (repeated fp32 adds, so every t_k is bit-identical to the
//     reference's), tests their occupancy bits in parallel, and resolves "emit" / "skip to the
//     next voxel" with ballots.  One atomic per ray reserves the output range.
//   * composite kernels are WARP-PER-RAY with __shfl scans (multiplicative scan of 1-alpha,
//     additive scans for the prefix sums the backward needs) and coalesced segment loads.
//   * grids are sized from the work, blocks are 256 threads (8 rays per block).
#include "common.cuh"
#include <float.h>
#include <math.h>

namespace enerf {

static constexpr float kSqrt3 = 1.7320508075688772f;
static constexpr float kRPi = 0.3183098861837907f;
static constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// per-ray kernels (thread per ray; 44 B of traffic per ray, nothing to optimise)
// ------------------------------------------------------------------------------------------

// raymarching.cu:110-146
__global__ void k_near_far_from_aabb(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                     const float* __restrict__ aabb, uint32_t N, float min_near,
                                     float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float lo = -FLT_MAX, hi = FLT_MAX;  // running slab intersection
    bool first = true, miss = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float o = rays_o[n * 3 + a];
        const float rd = 1.0f / rays_d[n * 3 + a];
        float t0 = (aabb[a] - o) * rd;
        float t1 = (aabb[a + 3] - o) * rd;
        if (t0 > t1) { float s = t0; t0 = t1; t1 = s; }
        if (first) { lo = t0; hi = t1; first = false; continue; }
        if (miss) continue;
        if (lo > t1 || t0 > hi) { miss = true; continue; }
        if (t0 > lo) lo = t0;
        if (t1 < hi) hi = t1;
    }
    if (miss) {
        nears[n] = FLT_MAX;
        fars[n] = FLT_MAX;
        return;
    }
    if (lo < min_near) lo = min_near;
    nears[n] = lo;
    fars[n] = hi;
}

// raymarching.cu:181-199
__global__ void k_polar_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                 float radius, uint32_t N, float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float Bh = ox * dx + oy * dy + oz * dz;
    const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-Bh + sqrtf(Bh * Bh - A * Cc)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    coords[n * 2] = 2 * theta * kRPi - 1;
    coords[n * 2 + 1] = phi * kRPi;
}

__global__ void k_morton3D(const int32_t* __restrict__ coords, uint32_t N, int32_t* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int32_t)morton3((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}

__global__ void k_morton3D_invert(const int32_t* __restrict__ indices, uint32_t N, int32_t* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int32_t ind = indices[n];  // arithmetic shifts, as the reference's `ind >> k` on int
    coords[n * 3] = (int32_t)compact3((uint32_t)(ind >> 0));
    coords[n * 3 + 1] = (int32_t)compact3((uint32_t)(ind >> 1));
    coords[n * 3 + 2] = (int32_t)compact3((uint32_t)(ind >> 2));
}

// packbits: one warp packs 32 consecutive cells per load instruction (coalesced 128 B), the
// byte is assembled with a ballot; 8 loads per warp iteration -> 32 output bytes.
// Bit order: bit i of byte n <- grid[8n + i] > thresh (raymarching.cu:283-290).
__global__ void k_packbits(const float* __restrict__ grid, uint32_t n_bytes, float thresh,
                           uint8_t* __restrict__ bitfield) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = lane_id();
    const uint32_t word0 = warp * 8;  // this warp produces 8 x 32 bits = 32 bytes
    const uint32_t n_words = n_bytes >> 2;
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t w = word0 + j;
        float v = -FLT_MAX;
        if (w < n_words) v = grid[(size_t)w * 32 + lane];
        const uint32_t bits = __ballot_sync(kFull, v > thresh);
        if (lane == (uint32_t)j) mine = bits;
    }
    if (lane < 8 && word0 + lane < n_words) reinterpret_cast<uint32_t*>(bitfield)[word0 + lane] = mine;
    // tail bytes when n_bytes is not a multiple of 4 (never for H=128): done by warp 0
    if (warp == 0 && lane < (n_bytes & 3u)) {
        const uint32_t b = (n_words << 2) + lane;
        uint8_t bits = 0;
        for (int i = 0; i < 8; ++i) bits |= grid[(size_t)b * 8 + i] > thresh ? (uint8_t)(1u << i) : (uint8_t)0;
        bitfield[b] = bits;
    }
}

// ------------------------------------------------------------------------------------------
// the marcher
// ------------------------------------------------------------------------------------------

struct MarchCfg {
    float bound, dt_gamma, dt_min, dt_max, Hf, inv_Hm1_dummy;
    uint32_t C, H, H3;
};

struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
};

__device__ __forceinline__ int frexp_exponent(float v) {
    // exponent e of frexpf(v) = m * 2^e, m in [0.5,1); only its clamp to [0, C-1] is used, so
    // zero / denormals (true exponent < 0) may return anything <= 0.
    return (int)((__float_as_uint(v) >> 23) & 0xffu) - 126;
}

__device__ __forceinline__ float step_of(float t, const MarchCfg& c) {
    return clampf(t * c.dt_gamma, c.dt_min, c.dt_max);
}

// Occupancy of the candidate at parameter t (raymarching.cu:362-380); on an empty cell also
// the parameter `tt` at which the ray leaves the voxel (raymarching.cu:391-394).
__device__ __forceinline__ bool probe(const Ray& r, const MarchCfg& c, const uint8_t* __restrict__ grid,
                                      float t, float dt, float& x, float& y, float& z, float& tt) {
    x = clampf(__fmaf_rn(t, r.dx, r.ox), -c.bound, c.bound);
    y = clampf(__fmaf_rn(t, r.dy, r.oy), -c.bound, c.bound);
    z = clampf(__fmaf_rn(t, r.dz, r.oz), -c.bound, c.bound);
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    const int top = (int)c.C - 1;
    const int lp = min(top, max(0, frexp_exponent(mx)));
    const int ld = min(top, max(0, frexp_exponent(dt * c.Hf * 0.5f)));
    const int level = max(lp, ld);
    const float mip_bound = fminf((float)(1 << level), c.bound);
    const float mip_rbound = 1.0f / mip_bound;
    // (float)(0.5 * (double)(x*rb+1) * (double)H): 0.5*f is exact and f*H fits a double, so the
    // single fp32 rounding below is the same value.
    const float top_cell = (float)(c.H - 1);
    const int nx = (int)clampf(__fmul_rn(0.5f * __fmaf_rn(x, mip_rbound, 1.0f), c.Hf), 0.0f, top_cell);
    const int ny = (int)clampf(__fmul_rn(0.5f * __fmaf_rn(y, mip_rbound, 1.0f), c.Hf), 0.0f, top_cell);
    const int nz = (int)clampf(__fmul_rn(0.5f * __fmaf_rn(z, mip_rbound, 1.0f), c.Hf), 0.0f, top_cell);
    const uint32_t index = (uint32_t)level * c.H3 + morton3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    const bool occ = (__ldg(grid + (index >> 3)) >> (index & 7u)) & 1u;
    if (!occ) {
        const float fx = ((float)nx + 0.5f) + 0.5f * copysignf(1.0f, r.dx);
        const float fy = ((float)ny + 0.5f) + 0.5f * copysignf(1.0f, r.dy);
        const float fz = ((float)nz + 0.5f) + 0.5f * copysignf(1.0f, r.dz);
        const float tx = __fmaf_rn(__fmaf_rn(fx / top_cell, 2.0f, -1.0f), mip_bound, -x) * r.rdx;
        const float ty = __fmaf_rn(__fmaf_rn(fy / top_cell, 2.0f, -1.0f), mip_bound, -y) * r.rdy;
        const float tz = __fmaf_rn(__fmaf_rn(fz / top_cell, 2.0f, -1.0f), mip_bound, -z) * r.rdz;
        tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    }
    return occ;
}

// Candidate `lane` of a batch that starts at tb when the step is the constant c: the value of `lane` sequential fp32 additions
// t <- t + c (the reference's loop, raymarching.cu:387,397), bit for bit.
// While t stays inside one binade every t is a multiple of u = ulp(t), so RN(t + c) = t + c' with the same c' = RN_u(c) at every
// step (a tie c mod u == u/2 settles on one increment after its first step) and the sequence is the arithmetic progression
// tb + k*c', exactly representable.  That case is detected (two equal successive increments, same exponent at both ends of the
// batch) and evaluated with one fma; otherwise the 31 dependent additions are done literally.
__device__ __forceinline__ float const_step_candidate(float tb, float c, unsigned lane) {
    const float t1 = tb + c;
    const float d1 = t1 - tb;                          // exact: t1 and tb are within a factor of two
    const float d2 = (t1 + c) - t1;
    const float t_end = __fmaf_rn(32.0f, d1, tb);      // one step past the batch
    const bool same_binade = (__float_as_uint(tb) >> 23) == (__float_as_uint(t_end) >> 23);
    if (d1 == d2 && same_binade && tb >= c && c > 0.0f) return __fmaf_rn((float)lane, d1, tb);   // tb >= c > 0: the subtractions above are exact
    float t = tb;
#pragma unroll
    for (int j = 0; j < 31; ++j) {
        const float tn = t + c;
        t = ((unsigned)j < lane) ? tn : t;
    }
    return t;
}

// Candidate 32 of a batch (= candidate 0 of the next one) without generating the 32 in between: what 32 sequential steps leave in t
__device__ __forceinline__ float advance_batch(float tb, const MarchCfg& c) {
    if (c.dt_gamma == 0.0f) {
        const float s = clampf(0.0f, c.dt_min, c.dt_max);
        const float t1 = tb + s;
        const float d1 = t1 - tb;
        const float d2 = (t1 + s) - t1;
        const float t_end = __fmaf_rn(32.0f, d1, tb);
        const bool same_binade = (__float_as_uint(tb) >> 23) == (__float_as_uint(t_end) >> 23);
        if (d1 == d2 && same_binade && tb >= s && s > 0.0f) return t_end;      // see const_step_candidate
        float t = tb;
#pragma unroll
        for (int j = 0; j < 32; ++j) t += s;
        return t;
    }
    float t = tb;
    for (int j = 0; j < 32; ++j) t += step_of(t, c);
    return t;
}

// The parameter interval in which a ray can meet an occupied cell at all: its intersection with the box around the occupied cells of
// every cascade level (`occ_bounds`: int32 [C][6] = min x, y, z, max x, y, z in cells; enerf_occupancy_bounds).  The sequence of
// candidate parameters t0, t0 + dt, ... does not depend on the occupancy (an empty voxel is left by repeated `t += dt`,
// raymarching.cu:390-398), so candidates outside the interval can be stepped over without looking at the grid: the samples, their
// order and every bit of them stay what the exhaustive march produces.  Without bounds: (-inf, +inf).
struct TRange {
    float lo, hi;
};
// The box around the occupied cells.  enerf_occupancy_bounds leaves it per cascade level, in units of the level's half extent
// (`k_occupancy_bounds`: float rows behind the integer rows); a ray's warp scales the rows by mip_bound and unites them — two 16-byte
// loads, six multiplies and six min / max per level where the integer rows took ~80 instructions per level, a third of what a ray
// costs per inference round (ncu, `profiles/r2_60`).
struct OccBox {
    float lo[3], hi[3];
    int state;                    // 0: no bounds given (march exhaustively), 1: box valid, 2: nothing occupied anywhere
};
__host__ __device__ __forceinline__ uint32_t occ_box_offset(uint32_t C) { return (6u * C + 3u) & ~3u; }   // in 4-byte words, 16-byte aligned
__device__ __forceinline__ OccBox occupied_box(const MarchCfg& c, const int32_t* __restrict__ occ_bounds) {
    OccBox box;
    box.state = 0;
    if (!occ_bounds) return box;
    const float4* rows = reinterpret_cast<const float4*>(occ_bounds + occ_box_offset(c.C));
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll 1
    for (uint32_t l = 0; l < c.C; ++l) {
        const float mb = fminf((float)(1u << l), c.bound);
        const float4 a = __ldg(rows + 2 * l), b = __ldg(rows + 2 * l + 1);
        if (a.w != 0.0f) {        // the level has occupied cells
            lo[0] = fminf(lo[0], a.x * mb); lo[1] = fminf(lo[1], a.y * mb); lo[2] = fminf(lo[2], a.z * mb);
            hi[0] = fmaxf(hi[0], b.x * mb); hi[1] = fmaxf(hi[1], b.y * mb); hi[2] = fmaxf(hi[2], b.z * mb);
        }
    }
    // one cell of the finest level plus rounding of the cell index / the slab test: far more than either can be off by
    const float margin = 2.0f * fminf(1.0f, c.bound) / c.Hf + 1e-3f * c.bound;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        box.lo[a] = lo[a] - margin;
        box.hi[a] = hi[a] + margin;
    }
    box.state = (hi[0] < lo[0]) ? 2 : 1;
    return box;
}
__device__ __forceinline__ TRange occupied_range(const Ray& r, const OccBox& box) {
    TRange o;
    o.lo = -FLT_MAX;
    o.hi = FLT_MAX;
    if (box.state == 0) return o;
    if (box.state == 2) {         // nothing occupied anywhere
        o.lo = FLT_MAX;
        o.hi = -FLT_MAX;
        return o;
    }
    const float org[3] = {r.ox, r.oy, r.oz}, rd[3] = {r.rdx, r.rdy, r.rdz};
    float t_in = -FLT_MAX, t_out = FLT_MAX;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float ta = (box.lo[a] - org[a]) * rd[a], tb = (box.hi[a] - org[a]) * rd[a];
        // a NaN (origin on a slab plane of an axis-parallel ray) leaves the axis unconstrained: fminf / fmaxf return the other operand
        t_in = fmaxf(t_in, fminf(ta, tb));
        t_out = fminf(t_out, fmaxf(ta, tb));
    }
    const float slack = 1e-4f * (1.0f + fabsf(t_in) + fabsf(t_out));
    o.lo = t_in - slack;
    o.hi = t_out + slack;
    if (!(o.lo <= o.hi)) {        // the ray misses the box (or a NaN crept in: then be exhaustive)
        if (o.lo > o.hi) {
            o.lo = FLT_MAX;
            o.hi = -FLT_MAX;
        } else {
            o.lo = -FLT_MAX;
            o.hi = FLT_MAX;
        }
    }
    return o;
}

// Marches one ray with one warp.  Emits at most `limit` samples, starting at t0; returns the
// number emitted (warp-uniform).  With WRITE, sample i of this ray goes to row i of
// xyzs/dirs/deltas (already offset to the ray's range).
// Log of the batches that emitted samples during the counting pass (per warp, in shared memory): with it the
// writing pass only regenerates t for those batches and never touches the grid again.
static constexpr int kLogBatches = 96;
struct BatchLog {
    float tb[kLogBatches];        // candidate 0 of the batch
    uint32_t emit[kLogBatches];   // lanes that emitted
};

template <bool WRITE>
__device__ uint32_t march_warp(const Ray& r, const MarchCfg& c, const uint8_t* __restrict__ grid,
                               float t0, float far, uint32_t limit, float* __restrict__ xyzs,
                               float* __restrict__ dirs, float* __restrict__ deltas, TRange occ, BatchLog* log = nullptr,
                               uint32_t* n_logged = nullptr) {
    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t count = 0;
    uint32_t logged = 0;          // kLogBatches + 1 = overflow (log unusable)
    float tb = t0;                // candidate 0 of the current batch
    float skip_to = -FLT_MAX;     // pending "skip until t >= skip_to" from an empty voxel
    float last_t = t0;            // t after the previously emitted sample (raymarching.cu:425,462)

    while (tb < far && count < limit) {
        if (tb > occ.hi) break;                   // no occupied cell from here on: nothing more to emit
        if (tb < occ.lo) {                        // the whole batch may lie before the first occupied cell: step over it unprobed
            const float te = advance_batch(tb, c);
            if (te <= occ.lo) {
                tb = te;
                continue;
            }
        }
        // 32 consecutive candidates; lane i applies the step function i times so that its t is
        // bit-identical to the sequential loop's.
        float t = tb;
        if (c.dt_gamma == 0.0f) {
            t = const_step_candidate(tb, clampf(0.0f, c.dt_min, c.dt_max), lane);
        } else {
            for (int j = 0; j < 31; ++j) {
                const float tn = t + step_of(t, c);
                t = ((unsigned)j < lane) ? tn : t;
            }
        }
        const float dt = step_of(t, c);
        const float tnext = t + dt;
        const bool in_range = t < far;

        float x = 0.f, y = 0.f, z = 0.f, tt = 0.f;
        bool occ = false;
        if (in_range) occ = probe(r, c, grid, t, dt, x, y, z, tt);

        const unsigned m_range = __ballot_sync(kFull, in_range);  // a prefix of the warp
        const unsigned m_occ = __ballot_sync(kFull, occ);
        // first candidate not swallowed by a skip that started in an earlier batch; if the whole
        // batch is swallowed the pending target survives to the next one.
        const unsigned m_unskipped = __ballot_sync(kFull, t >= skip_to);
        int cur = 32;
        if (m_unskipped) {
            cur = __ffs(m_unskipped) - 1;
            skip_to = -FLT_MAX;
        }

        unsigned emit = 0;
        uint32_t room = limit - count;
        while (cur < 32 && ((m_range >> cur) & 1u) && room > 0) {
            if ((m_occ >> cur) & 1u) {
                // maximal run of occupied in-range candidates starting at cur
                const unsigned stop = (~(m_occ & m_range)) & (kFull << cur);
                const int e = stop ? (__ffs(stop) - 1) : 32;
                // the run is the contiguous candidates [cur, e): when there is room for fewer, keep its first `room` (the later ones
                // are dropped and `cur = e` ends the batch's scan either way, as when they were trimmed one bit at a time)
                const int e_keep = (int)min((uint32_t)e, (uint32_t)cur + room);
                const unsigned run = (e_keep >= 32 ? kFull : ((1u << e_keep) - 1u)) & (kFull << cur);
                emit |= run;
                room -= (uint32_t)(e_keep - cur);
                cur = e;
            } else {
                const float target = __shfl_sync(kFull, tt, cur);
                const unsigned later = (cur >= 31) ? 0u : (kFull << (cur + 1));
                const unsigned m = __ballot_sync(kFull, t >= target) & later;
                if (m) {
                    cur = __ffs(m) - 1;
                } else {
                    skip_to = target;
                    cur = 32;
                }
            }
        }

        if (WRITE && emit) {
            const bool me = (emit >> lane) & 1u;
            const unsigned below = emit & lt_mask;
            const int prev = below ? (31 - __clz(below)) : 0;
            const float prev_tnext = __shfl_sync(kFull, tnext, prev);
            if (me) {
                const size_t row = (size_t)count + (size_t)__popc(below);
                xyzs[row * 3 + 0] = x;
                xyzs[row * 3 + 1] = y;
                xyzs[row * 3 + 2] = z;
                dirs[row * 3 + 0] = r.dx;
                dirs[row * 3 + 1] = r.dy;
                dirs[row * 3 + 2] = r.dz;
                deltas[row * 2 + 0] = dt;
                deltas[row * 2 + 1] = tnext - (below ? prev_tnext : last_t);
            }
            last_t = __shfl_sync(kFull, tnext, 31 - __clz(emit));
        }
        count += (uint32_t)__popc(emit);
        if (!WRITE && log != nullptr && emit) {
            if (logged < (uint32_t)kLogBatches) {
                if (lane == 0) {
                    log->tb[logged] = tb;
                    log->emit[logged] = emit;
                }
                ++logged;
            } else {
                logged = kLogBatches + 1;
            }
        }

        if (~m_range) break;                      // some candidate reached `far`: ray finished
        tb = __shfl_sync(kFull, tnext, 31);       // candidate 32 = next batch's candidate 0
    }
    if (n_logged) *n_logged = logged;
    return count;
}

// Writing pass from the batch log: same t values (same repeated adds), same outputs, no grid probes.
__device__ void replay_warp(const Ray& r, const MarchCfg& c, float t0, const BatchLog* log, uint32_t n_logged,
                            float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas) {
    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t count = 0;
    float last_t = t0;
    for (uint32_t b = 0; b < n_logged; ++b) {
        const float tb = log->tb[b];
        const unsigned emit = log->emit[b];
        float t = tb;
        if (c.dt_gamma == 0.0f) {
            t = const_step_candidate(tb, clampf(0.0f, c.dt_min, c.dt_max), lane);
        } else {
            for (int j = 0; j < 31; ++j) {
                const float tn = t + step_of(t, c);
                t = ((unsigned)j < lane) ? tn : t;
            }
        }
        const float dt = step_of(t, c);
        const float tnext = t + dt;
        const bool me = (emit >> lane) & 1u;
        const unsigned below = emit & lt_mask;
        const int prev = below ? (31 - __clz(below)) : 0;
        const float prev_tnext = __shfl_sync(kFull, tnext, prev);
        if (me) {
            const size_t row = (size_t)count + (size_t)__popc(below);
            xyzs[row * 3 + 0] = clampf(__fmaf_rn(t, r.dx, r.ox), -c.bound, c.bound);
            xyzs[row * 3 + 1] = clampf(__fmaf_rn(t, r.dy, r.oy), -c.bound, c.bound);
            xyzs[row * 3 + 2] = clampf(__fmaf_rn(t, r.dz, r.oz), -c.bound, c.bound);
            dirs[row * 3 + 0] = r.dx;
            dirs[row * 3 + 1] = r.dy;
            dirs[row * 3 + 2] = r.dz;
            deltas[row * 2 + 0] = dt;
            deltas[row * 2 + 1] = tnext - (below ? prev_tnext : last_t);
        }
        last_t = __shfl_sync(kFull, tnext, 31 - __clz(emit));
        count += (uint32_t)__popc(emit);
    }
}

__device__ __forceinline__ Ray load_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, uint32_t n) {
    Ray r;
    r.ox = rays_o[n * 3]; r.oy = rays_o[n * 3 + 1]; r.oz = rays_o[n * 3 + 2];
    r.dx = rays_d[n * 3]; r.dy = rays_d[n * 3 + 1]; r.dz = rays_d[n * 3 + 2];
    r.rdx = 1.0f / r.dx; r.rdy = 1.0f / r.dy; r.rdz = 1.0f / r.dz;
    return r;
}

// raymarching.cu:313-480.  One warp per ray; pass 1 counts, one atomic pair reserves, pass 2 writes.
__global__ void __launch_bounds__(256)
k_march_rays_train(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                   const uint8_t* __restrict__ grid, MarchCfg c, uint32_t max_steps, uint32_t N, uint32_t M,
                   const float* __restrict__ nears, const float* __restrict__ fars,
                   float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                   int32_t* __restrict__ rays, int32_t* __restrict__ counter, uint32_t perturb, const int32_t* __restrict__ occ_bounds) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const unsigned lane = lane_id();
    const Ray r = load_ray(rays_o, rays_d, n);
    const TRange occ = occupied_range(r, occupied_box(c, occ_bounds));
    const float near = nears[n], far = fars[n];
    float t0 = near;
    if (perturb) {
        Pcg32 rng((uint64_t)n, 1u);
        t0 = __fmaf_rn(c.dt_min, rng.next_float(), t0);   // the reference's `t0 += dt_min * r` is contracted by nvcc
    }
    __shared__ BatchLog logs[8];
    BatchLog* log = &logs[threadIdx.x >> 5];
    uint32_t n_logged = 0;
    const uint32_t num_steps = march_warp<false>(r, c, grid, t0, far, max_steps, nullptr, nullptr, nullptr, occ, log, &n_logged);
    __syncwarp();

    uint32_t point_index = 0;
    if (lane == 0) {
        point_index = (uint32_t)atomicAdd(counter, (int)num_steps);
        const uint32_t ray_index = (uint32_t)atomicAdd(counter + 1, 1);
        rays[ray_index * 3] = (int32_t)n;
        rays[ray_index * 3 + 1] = (int32_t)point_index;
        rays[ray_index * 3 + 2] = (int32_t)num_steps;
    }
    point_index = __shfl_sync(kFull, point_index, 0);
    if (num_steps == 0) return;
    if (point_index + num_steps >= M) return;  // `>=` as raymarching.cu:416
    float* px = xyzs + (size_t)point_index * 3;
    float* pd = dirs + (size_t)point_index * 3;
    float* pl = deltas + (size_t)point_index * 2;
    if (n_logged <= (uint32_t)kLogBatches) replay_warp(r, c, t0, log, n_logged, px, pd, pl);
    else march_warp<true>(r, c, grid, t0, far, num_steps, px, pd, pl, occ);     // more emitting batches than the log holds: re-march
}

// The inference loop of NeRFRenderer.run_cuda (renderer.py:364-391) reads the number of alive rays back to the host after every
// compaction.  Here the kernels also accept that count as a DEVICE pointer: `n_alive` is then only an upper bound (rays never come
// back to life, so any earlier count is one) and slots at or beyond *n_alive_dev are skipped — the host can keep enqueueing rounds
// and synchronise only every few rounds.
__device__ __forceinline__ uint32_t alive_count(uint32_t n_alive, const int32_t* __restrict__ n_alive_dev) {
    if (!n_alive_dev) return n_alive;
    const int32_t v = *n_alive_dev;
    return v < 0 ? 0u : min(n_alive, (uint32_t)v);
}

// rows [from, to) of one slot's xyzs / dirs / deltas <- 0, by the slot's warp (consecutive floats, lane-strided)
__device__ __forceinline__ void zero_rows(float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas, uint32_t from,
                                          uint32_t to, unsigned lane) {
    for (uint32_t j = from * 3u + lane; j < to * 3u; j += 32u) {
        xyzs[j] = 0.0f;
        dirs[j] = 0.0f;
    }
    for (uint32_t j = from * 2u + lane; j < to * 2u; j += 32u) deltas[j] = 0.0f;
}

// raymarching.cu:700-804.  One warp per alive ray, at most n_step samples from rays_t.
__global__ void __launch_bounds__(256, 5)
k_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
             const float* __restrict__ rays_t, const float* __restrict__ rays_o,
             const float* __restrict__ rays_d, MarchCfg c, const uint8_t* __restrict__ grid,
             const float* __restrict__ nears, const float* __restrict__ fars, float* __restrict__ xyzs,
             float* __restrict__ dirs, float* __restrict__ deltas, uint32_t perturb, const int32_t* __restrict__ n_alive_dev,
             const int32_t* __restrict__ occ_bounds) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= n_alive) return;
    // slot n exists in both arrays whatever the device-side count says: its loads go out together with the count's, and so do the
    // rows of the box, which depend on neither
    const uint32_t index = (uint32_t)rays_alive[n];
    float t = rays_t[n];
    const OccBox box = occupied_box(c, occ_bounds);
    const size_t base = (size_t)n * n_step;
    const unsigned lane = lane_id();
    // Every row of the slot is written here, samples first and zeros behind them ("slots without a sample stay zero",
    // raymarching.py:205-207; a zero delta is how composite_rays recognises the end of a ray's round), dead slots included: the caller
    // does not have to clear n_alive * n_step * 32 bytes per round first.
    if (n >= alive_count(n_alive, n_alive_dev)) {
        zero_rows(xyzs + base * 3, dirs + base * 3, deltas + base * 2, 0u, n_step, lane);
        return;
    }
    const Ray r = load_ray(rays_o, rays_d, index);
    const float far = fars[index];
    if (perturb) {
        Pcg32 rng((uint64_t)n, (uint64_t)perturb);  // seeded by the alive SLOT, raymarching.cu:743
        t = __fmaf_rn(c.dt_min, rng.next_float(), t);
    }
    const uint32_t emitted = march_warp<true>(r, c, grid, t, far, n_step, xyzs + base * 3, dirs + base * 3, deltas + base * 2, occupied_range(r, box));
    zero_rows(xyzs + base * 3, dirs + base * 3, deltas + base * 2, emitted, n_step, lane);
}

// Box around the occupied cells of each cascade level: one CTA per level scans the level's bits 128 at a time (128 consecutive Morton
// codes = an aligned 8 x 4 x 4 block of cells, taken whole when any of its bits is set).  out: int32 [C][6] = min x, y, z, max x, y, z;
// an empty level gets min = H, max = -1.  Behind them, from word occ_box_offset(C): float [C][8], the box the marchers read.
__global__ void __launch_bounds__(1024)
k_occupancy_bounds(const uint8_t* __restrict__ grid, uint32_t H, uint32_t words_per_level, int32_t* __restrict__ out) {
    const uint32_t level = blockIdx.x;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(grid) + (size_t)level * words_per_level;
    int mn[3] = {(int)H, (int)H, (int)H}, mx[3] = {-1, -1, -1};
    // `span` consecutive Morton codes starting at a multiple of span = an aligned block of ex x ey x ez cells, taken whole
    auto take = [&](uint32_t code, int ex, int ey, int ez) {
        const int x = (int)compact3(code), y = (int)compact3(code >> 1), z = (int)compact3(code >> 2);
        mn[0] = min(mn[0], x); mn[1] = min(mn[1], y); mn[2] = min(mn[2], z);
        mx[0] = max(mx[0], x + ex - 1); mx[1] = max(mx[1], y + ey - 1); mx[2] = max(mx[2], z + ez - 1);
    };
    // one CTA reads a whole level (256 KB at H = 128): eight independent 16-byte loads in flight per thread (a plain loop is a chain of
    // memory latencies: 46 us for three levels), and one decode per non-zero 16 bytes = 128 codes = an 8 x 4 x 4 block
    const uint32_t n4 = (words_per_level % 4u == 0u) ? words_per_level / 4u : 0u;      // levels stay 16-byte aligned only then
    const uint4* w4 = reinterpret_cast<const uint4*>(w);
    for (uint32_t base = 0; base < n4; base += 8u * blockDim.x) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t i = base + (uint32_t)k * blockDim.x + threadIdx.x;
            v[k] = i < n4 ? __ldg(w4 + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if ((v[k].x | v[k].y | v[k].z | v[k].w) != 0u) take((base + (uint32_t)k * blockDim.x + threadIdx.x) * 128u, 8, 4, 4);
    }
    for (uint32_t i = n4 * 4u + threadIdx.x; i < words_per_level; i += blockDim.x)
        if (__ldg(w + i) != 0u) take(i * 32u, 4, 4, 2);
    __shared__ int s_mn[3][32], s_mx[3][32];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) {
            mn[a] = min(mn[a], __shfl_xor_sync(kFull, mn[a], k));
            mx[a] = max(mx[a], __shfl_xor_sync(kFull, mx[a], k));
        }
        if (lane == 0) {
            s_mn[a][warp] = mn[a];
            s_mx[a][warp] = mx[a];
        }
    }
    __syncthreads();
    if (warp == 0) {
        const unsigned n_warps = blockDim.x >> 5;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int v0 = lane < n_warps ? s_mn[a][lane] : (int)H, v1 = lane < n_warps ? s_mx[a][lane] : -1;
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) {
                v0 = min(v0, __shfl_xor_sync(kFull, v0, k));
                v1 = max(v1, __shfl_xor_sync(kFull, v1, k));
            }
            if (lane == 0) {
                const int top = min(v1, (int)H - 1);
                out[level * 6 + a] = v0;
                out[level * 6 + 3 + a] = top;
                // the same box in units of the level's half extent ([-1, 1]; cell i spans [2 i / H - 1, 2 (i + 1) / H - 1]), w = 1 when
                // the level has occupied cells: what the marchers read (occupied_box)
                float* row = reinterpret_cast<float*>(out + occ_box_offset(gridDim.x)) + level * 8;
                const float cell = 2.0f / (float)H;
                row[a] = __fmaf_rn((float)v0, cell, -1.0f);
                row[4 + a] = __fmaf_rn((float)(top + 1), cell, -1.0f);
                if (a == 0) {
                    row[3] = v1 >= v0 ? 1.0f : 0.0f;
                    row[7] = 0.0f;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// compositing
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ float scan_mul_incl(float v, unsigned lane) {
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const float o = __shfl_up_sync(kFull, v, k);
        if (lane >= (unsigned)k) v *= o;
    }
    return v;
}
__device__ __forceinline__ float scan_add_incl(float v, unsigned lane) {
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const float o = __shfl_up_sync(kFull, v, k);
        if (lane >= (unsigned)k) v += o;
    }
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(kFull, v, k);
    return v;
}

static constexpr int kCompositeGroup = 4;       // 32-sample chunks fetched together (one group ahead of the one being composited)
// one lane's sample of a 32-sample chunk (zeros beyond the ray's last sample)
template <int NCH>
struct CompositeChunk {
    float sigma, d0, d1, c[NCH];
};
template <int NCH>
__device__ __forceinline__ CompositeChunk<NCH> composite_fetch(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                               const float* __restrict__ deltas, uint32_t i, uint32_t num_steps) {
    CompositeChunk<NCH> k;
    k.sigma = k.d0 = k.d1 = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) k.c[ch] = 0.f;
    if (i < num_steps) {
        const float2 dl = *reinterpret_cast<const float2*>(deltas + (size_t)i * 2);
        k.sigma = sigmas[i];
        k.d0 = dl.x;
        k.d1 = dl.y;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) k.c[ch] = rgbs[(size_t)i * NCH + ch];
    }
    return k;
}

// raymarching.cu:500-578.  Warp per ray.
template <int NCH>
__global__ void __launch_bounds__(256)
k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                      const float* __restrict__ deltas, const int32_t* __restrict__ rays, uint32_t M,
                      uint32_t N, float* __restrict__ weights_sum, float* __restrict__ depth,
                      float* __restrict__ image) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const unsigned lane = lane_id();
    const uint32_t index = (uint32_t)rays[n * 3];
    const uint32_t offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];

    if (num_steps == 0 || offset + num_steps >= M) {
        if (lane == 0) {
            weights_sum[index] = 0;
            depth[index] = 0;
        }
        if (lane < NCH) image[index * NCH + lane] = 0;
        return;
    }
    sigmas += offset;
    rgbs += (size_t)offset * NCH;
    deltas += (size_t)offset * 2;

    float T_carry = 1.0f, t_carry = 0.0f;
    float acc_c[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) acc_c[ch] = 0.f;
    float acc_d = 0.f, acc_ws = 0.f;

    // 4096 rays are 28 warps per SM, each a chain of ~25 dependent iterations: the loads of the next kCompositeGroup x 32 samples are
    // in flight while the scans of the current group run, or every iteration waits for a round trip to memory
    auto process = [&](const CompositeChunk<NCH>& cur, uint32_t base) {
        const uint32_t i = base + lane;
        const bool valid = i < num_steps;
        float alpha = 0.f, d1 = 0.f;
        if (valid) {
            alpha = 1.0f - __expf(-cur.sigma * cur.d0);
            d1 = cur.d1;
        }
        const float incl = scan_mul_incl(1.0f - alpha, lane);
        float excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float w = alpha * (T_carry * excl);
        const float t_incl = scan_add_incl(d1, lane);
        const float t = t_carry + t_incl;
        if (valid) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) acc_c[ch] += w * cur.c[ch];
            acc_d += w * t;
            acc_ws += w;
        }
        T_carry *= __shfl_sync(kFull, incl, 31);
        t_carry += __shfl_sync(kFull, t_incl, 31);
    };
    CompositeChunk<NCH> cur[kCompositeGroup], nxt[kCompositeGroup];
#pragma unroll
    for (int k = 0; k < kCompositeGroup; ++k) cur[k] = composite_fetch<NCH>(sigmas, rgbs, deltas, 32u * k + lane, num_steps);
    for (uint32_t gbase = 0; gbase < num_steps; gbase += 32u * kCompositeGroup) {
        const uint32_t nbase = gbase + 32u * kCompositeGroup;
        if (nbase < num_steps) {
#pragma unroll
            for (int k = 0; k < kCompositeGroup; ++k) nxt[k] = composite_fetch<NCH>(sigmas, rgbs, deltas, nbase + 32u * k + lane, num_steps);
        }
#pragma unroll
        for (int k = 0; k < kCompositeGroup; ++k)
            if (gbase + 32u * k < num_steps) process(cur[k], gbase + 32u * k);
#pragma unroll
        for (int k = 0; k < kCompositeGroup; ++k) cur[k] = nxt[k];
    }
    acc_d = warp_sum(acc_d);
    acc_ws = warp_sum(acc_ws);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) acc_c[ch] = warp_sum(acc_c[ch]);
    if (lane == 0) {
        weights_sum[index] = acc_ws;
        depth[index] = acc_d;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) image[index * NCH + ch] = acc_c[ch];
    }
}

// raymarching.cu:602-682.  Warp per ray.
template <int NCH>
__global__ void __launch_bounds__(256)
k_composite_train_bwd(const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image,
                      const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                      const float* __restrict__ deltas, const int32_t* __restrict__ rays,
                      const float* __restrict__ weights_sum, const float* __restrict__ image, uint32_t M,
                      uint32_t N, float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const unsigned lane = lane_id();
    const uint32_t index = (uint32_t)rays[n * 3];
    const uint32_t offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps >= M) return;

    sigmas += offset;
    rgbs += (size_t)offset * NCH;
    deltas += (size_t)offset * 2;
    grad_sigmas += offset;
    grad_rgbs += (size_t)offset * NCH;

    float g_c[NCH], final_c[NCH], pre_c[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        g_c[ch] = grad_image[index * NCH + ch];
        final_c[ch] = image[index * NCH + ch];
        pre_c[ch] = 0.f;
    }
    const float g_ws = grad_weights_sum[index];
    const float ws_final = weights_sum[index];
    float pre_ws = 0.f, T_carry = 1.0f;

    auto process = [&](const CompositeChunk<NCH>& cur, uint32_t base) {
        const uint32_t i = base + lane;
        const bool valid = i < num_steps;
        float alpha = 0.f, d0 = 0.f;
        float c[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) c[ch] = 0.f;
        if (valid) {
            d0 = cur.d0;
            alpha = 1.0f - __expf(-cur.sigma * d0);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) c[ch] = cur.c[ch];
        }
        const float incl = scan_mul_incl(1.0f - alpha, lane);
        float excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float w = alpha * (T_carry * excl);
        const float T_after = T_carry * incl;  // T after this sample (raymarching.cu:656)

        const float ws_incl = pre_ws + scan_add_incl(w, lane);
        float acc = g_ws * (T_after - (ws_final - ws_incl));
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const float c_incl = pre_c[ch] + scan_add_incl(w * c[ch], lane);
            acc += g_c[ch] * (T_after * c[ch] - (final_c[ch] - c_incl));
            pre_c[ch] = __shfl_sync(kFull, c_incl, 31);
            if (valid) grad_rgbs[(size_t)i * NCH + ch] = g_c[ch] * w;
        }
        if (valid) grad_sigmas[i] = d0 * acc;
        pre_ws = __shfl_sync(kFull, ws_incl, 31);
        T_carry *= __shfl_sync(kFull, incl, 31);
    };
    // loads one group of chunks ahead, as in the forward kernel
    CompositeChunk<NCH> cur[kCompositeGroup], nxt[kCompositeGroup];
#pragma unroll
    for (int k = 0; k < kCompositeGroup; ++k) cur[k] = composite_fetch<NCH>(sigmas, rgbs, deltas, 32u * k + lane, num_steps);
    for (uint32_t gbase = 0; gbase < num_steps; gbase += 32u * kCompositeGroup) {
        const uint32_t nbase = gbase + 32u * kCompositeGroup;
        if (nbase < num_steps) {
#pragma unroll
            for (int k = 0; k < kCompositeGroup; ++k) nxt[k] = composite_fetch<NCH>(sigmas, rgbs, deltas, nbase + 32u * k + lane, num_steps);
        }
#pragma unroll
        for (int k = 0; k < kCompositeGroup; ++k)
            if (gbase + 32u * k < num_steps) process(cur[k], gbase + 32u * k);
#pragma unroll
        for (int k = 0; k < kCompositeGroup; ++k) cur[k] = nxt[k];
    }
}

// raymarching.cu:816-900.  n_step <= 8 samples per ray per call: thread per ray.
template <int NCH>
__global__ void k_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
                                 float* __restrict__ rays_t, const float* __restrict__ sigmas,
                                 const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                 float* __restrict__ weights_sum, float* __restrict__ depth,
                                 float* __restrict__ image, const int32_t* __restrict__ n_alive_dev) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= alive_count(n_alive, n_alive_dev)) return;
    const uint32_t index = (uint32_t)rays_alive[n];
    float t = rays_t[n];
    sigmas += (size_t)n * n_step;
    rgbs += (size_t)n * n_step * NCH;
    deltas += (size_t)n * n_step * 2;

    float ws = weights_sum[index];
    float d = depth[index];
    float c[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) c[ch] = image[index * NCH + ch];

    uint32_t step = 0;
    while (step < n_step) {
        const float d0 = deltas[step * 2];
        if (d0 == 0) break;
        const float alpha = 1.0f - __expf(-sigmas[step] * d0);
        const float T = 1 - ws;
        const float w = alpha * T;
        ws += w;
        t += deltas[step * 2 + 1];
        d += w * t;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) c[ch] += w * rgbs[step * NCH + ch];
        if (T < 1e-5f) break;   // NOTE reference compares against the double literal 1e-5
        step++;
    }
    rays_t[n] = (step < n_step) ? -1.0f : t;
    weights_sum[index] = ws;
    depth[index] = d;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) image[index * NCH + ch] = c[ch];
}

// The same in-place accumulation with one WARP per ray, for rounds of many steps (the mirror's inference loop marches up to 2^23
// samples per round, i.e. n_step in the tens or hundreds): the ray's n_step consecutive samples are read 32 at a time, fully coalesced;
// T_i = (1 - ws) * prod_{j<i} (1 - alpha_j) comes from a multiplicative shuffle scan (mathematically the reference's running
// `T = 1 - weight_sum`, raymarching.cu:872), t_i from an additive scan of deltas[:,1]; the two stopping rules become ballots:
//   * a sample with delta == 0 is not consumed and ends the ray's round (padding written by march_rays);
//   * a sample that sees T < 1e-5 IS consumed and then ends the ray (raymarching.cu:884) -> rays_t = -1.
template <int NCH>
__global__ void __launch_bounds__(256)
k_composite_rays_warp(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive, float* __restrict__ rays_t,
                      const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                      float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image,
                      const int32_t* __restrict__ n_alive_dev) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= n_alive) return;
    const unsigned lane = lane_id();
    const uint32_t index = (uint32_t)rays_alive[n];      // loads of slot n next to the load of the device-side count, not behind it
    float t = rays_t[n];
    if (n >= alive_count(n_alive, n_alive_dev)) return;
    sigmas += (size_t)n * n_step;
    rgbs += (size_t)n * n_step * NCH;
    deltas += (size_t)n * n_step * 2;
    float T = 1.0f - weights_sum[index];           // transmittance in front of the first sample of this round
    float acc_ws = 0.f, acc_d = 0.f, acc_c[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) acc_c[ch] = 0.f;
    bool ended = false;
    for (uint32_t base = 0; base < n_step && !ended; base += 32) {
        const uint32_t i = base + lane;
        const bool in = i < n_step;
        float d0 = 0.f, d1 = 0.f, sg = 0.f;
        if (in) {
            const float2 dd = *reinterpret_cast<const float2*>(deltas + (size_t)i * 2);
            d0 = dd.x;
            d1 = dd.y;
            sg = sigmas[i];
        }
        const bool real = in && d0 != 0.f;
        const float alpha = real ? 1.0f - __expf(-sg * d0) : 0.f;
        const float incl = scan_mul_incl(1.0f - alpha, lane);
        float excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float Ti = T * excl;                                   // transmittance seen by sample i
        const float ti = t + scan_add_incl(real ? d1 : 0.f, lane);   // t after consuming sample i
        // first sample that is padding (not consumed) / first consumed sample that saw T < 1e-5
        const unsigned m_pad = __ballot_sync(kFull, !real);
        const unsigned m_small = __ballot_sync(kFull, real && Ti < 1e-5f);
        const uint32_t first_pad = m_pad ? (uint32_t)(__ffs(m_pad) - 1) : 32u;
        const uint32_t first_small = m_small ? (uint32_t)(__ffs(m_small) - 1) : 32u;
        const uint32_t consumed = min(first_pad, first_small + 1u);  // lanes [0, consumed) contribute
        if (lane < consumed) {
            const float w = alpha * Ti;
            acc_ws += w;
            acc_d += w * ti;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) acc_c[ch] += w * rgbs[(size_t)i * NCH + ch];
        }
        const bool stop_small = first_small < 32u && first_small + 1u == consumed;      // the transmittance rule fired (no padding before it)
        const bool stop_pad = first_pad < 32u && consumed == first_pad;                 // padding, or the end of the round's n_step samples
        if (stop_small || stop_pad) {
            ended = true;
            if (consumed > 0) t = __shfl_sync(kFull, ti, consumed - 1);
            // `step < n_step` in the reference <=> the ray stopped before using up every sample of the round: it is finished
            if (stop_small || base + consumed < n_step) t = -1.0f;
        } else {
            t = __shfl_sync(kFull, ti, 31);
            T *= __shfl_sync(kFull, incl, 31);
        }
    }
    acc_ws = warp_sum(acc_ws);
    acc_d = warp_sum(acc_d);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) acc_c[ch] = warp_sum(acc_c[ch]);
    if (lane == 0) {
        rays_t[n] = t;
        weights_sum[index] += acc_ws;
        depth[index] += acc_d;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) image[index * NCH + ch] += acc_c[ch];
    }
}

// raymarching.cu:912-930 with one atomic per warp (ballot-aggregated).
__global__ void k_compact_rays(uint32_t n_alive, int32_t* __restrict__ rays_alive,
                               const int32_t* __restrict__ rays_alive_old, float* __restrict__ rays_t,
                               const float* __restrict__ rays_t_old, int32_t* __restrict__ alive_counter,
                               const int32_t* __restrict__ n_alive_dev) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = lane_id();
    const uint32_t n_old = alive_count(n_alive, n_alive_dev);
    float t = -1.f;
    int32_t id = 0;
    if (n < n_old) {
        t = rays_t_old[n];
        id = rays_alive_old[n];
    }
    const bool keep = (n < n_old) && (t >= 0);
    const unsigned m = __ballot_sync(kFull, keep);
    if (!m) return;
    int base = 0;
    if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(alive_counter, __popc(m));
    base = __shfl_sync(kFull, base, __ffs(m) - 1);
    if (keep) {
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        rays_alive[slot] = id;
        rays_t[slot] = t;
    }
}

static MarchCfg make_cfg(float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    MarchCfg c;
    c.bound = bound;
    c.dt_gamma = dt_gamma;
    c.dt_min = 2 * kSqrt3 / max_steps;                   // raymarching.cu:344
    c.dt_max = 2 * kSqrt3 * (1 << (C - 1)) / H;          // raymarching.cu:345
    c.Hf = (float)H;
    c.inv_Hm1_dummy = 0.f;
    c.C = C;
    c.H = H;
    c.H3 = H * H * H;
    return c;
}

}  // namespace enerf

using namespace enerf;

extern "C" {

int enerf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                             float min_near, float* nears, float* fars, void* stream) {
    if (N == 0) return 0;
    k_near_far_from_aabb<<<ceil_div(N, 256u), 256, 0, as_stream(stream)>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    ENERF_CHECK_LAUNCH("near_far_from_aabb");
    return 0;
}

int enerf_polar_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, void* stream) {
    if (N == 0) return 0;
    k_polar_from_ray<<<ceil_div(N, 256u), 256, 0, as_stream(stream)>>>(rays_o, rays_d, radius, N, coords);
    ENERF_CHECK_LAUNCH("polar_from_ray");
    return 0;
}

int enerf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream) {
    if (N == 0) return 0;
    k_morton3D<<<ceil_div(N, 256u), 256, 0, as_stream(stream)>>>(coords, N, indices);
    ENERF_CHECK_LAUNCH("morton3D");
    return 0;
}

int enerf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream) {
    if (N == 0) return 0;
    k_morton3D_invert<<<ceil_div(N, 256u), 256, 0, as_stream(stream)>>>(indices, N, coords);
    ENERF_CHECK_LAUNCH("morton3D_invert");
    return 0;
}

int enerf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream) {
    if (N == 0) return 0;
    ENERF_REQUIRE(((uintptr_t)bitfield & 3u) == 0, "packbits", "bitfield must be 4-byte aligned");
    const uint32_t n_warps = ceil_div(N, 32u);  // 32 bytes per warp
    k_packbits<<<max(1u, ceil_div(n_warps * 32u, 256u)), 256, 0, as_stream(stream)>>>(grid, N, density_thresh, bitfield);
    ENERF_CHECK_LAUNCH("packbits");
    return 0;
}

int enerf_occupancy_bounds(const uint8_t* grid, uint32_t C, uint32_t H, int32_t* bounds, void* stream) {
    ENERF_REQUIRE(C >= 1 && C <= 16 && H >= 4 && H <= 1024 && (H & (H - 1)) == 0, "occupancy_bounds", "C in [1,16], H a power of two in [4,1024]");
    ENERF_REQUIRE(((uintptr_t)grid & 15u) == 0, "occupancy_bounds", "bitfield must be 16-byte aligned");
    k_occupancy_bounds<<<C, 1024, 0, as_stream(stream)>>>(grid, H, H * H * H / 32u, bounds);
    ENERF_CHECK_LAUNCH("occupancy_bounds");
    return 0;
}

int enerf_march_rays_train_bounded(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                   float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                   const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                                   int32_t* rays, int32_t* counter, uint32_t perturb, const int32_t* occ_bounds, void* stream) {
    if (N == 0) return 0;
    ENERF_REQUIRE(C >= 1 && C <= 16 && H >= 2 && H <= 1024 && max_steps > 0, "march_rays_train", "bad C/H/max_steps");
    const MarchCfg c = make_cfg(bound, dt_gamma, max_steps, C, H);
    k_march_rays_train<<<ceil_div(N, 8u), 256, 0, as_stream(stream)>>>(rays_o, rays_d, grid, c, max_steps, N, M, nears, fars,
                                                                      xyzs, dirs, deltas, rays, counter, perturb, occ_bounds);
    ENERF_CHECK_LAUNCH("march_rays_train");
    return 0;
}
int enerf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                           float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                           const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                           int32_t* rays, int32_t* counter, uint32_t perturb, void* stream) {
    return enerf_march_rays_train_bounded(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter,
                                          perturb, nullptr, stream);
}

#define ENERF_NCH_SWITCH(n_ch, name, CALL)                                         \
    switch (n_ch) {                                                                \
        case 1: { constexpr int NCH = 1; CALL; } break;                            \
        case 2: { constexpr int NCH = 2; CALL; } break;                            \
        case 3: { constexpr int NCH = 3; CALL; } break;                            \
        case 4: { constexpr int NCH = 4; CALL; } break;                            \
        default: ::enerf::set_error("%s: n_ch must be in [1,4]", name); return -2; \
    }

int enerf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                       const int32_t* rays, uint32_t M, uint32_t N, uint32_t n_ch,
                                       float* weights_sum, float* depth, float* image, void* stream) {
    if (N == 0) return 0;
    ENERF_NCH_SWITCH(n_ch, "composite_rays_train_forward",
                     (k_composite_train_fwd<NCH><<<ceil_div(N, 8u), 256, 0, as_stream(stream)>>>(
                         sigmas, rgbs, deltas, rays, M, N, weights_sum, depth, image)));
    ENERF_CHECK_LAUNCH("composite_rays_train_forward");
    return 0;
}

int enerf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                        const float* rgbs, const float* deltas, const int32_t* rays,
                                        const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                        uint32_t n_ch, float* grad_sigmas, float* grad_rgbs, void* stream) {
    if (N == 0) return 0;
    ENERF_NCH_SWITCH(n_ch, "composite_rays_train_backward",
                     (k_composite_train_bwd<NCH><<<ceil_div(N, 8u), 256, 0, as_stream(stream)>>>(
                         grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N,
                         grad_sigmas, grad_rgbs)));
    ENERF_CHECK_LAUNCH("composite_rays_train_backward");
    return 0;
}

int enerf_march_rays_bounded(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                             const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                             uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                             float* dirs, float* deltas, uint32_t perturb, const int32_t* n_alive_dev, const int32_t* occ_bounds, void* stream) {
    if (n_alive == 0 || n_step == 0) return 0;
    ENERF_REQUIRE(C >= 1 && C <= 16 && H >= 2 && H <= 1024 && max_steps > 0, "march_rays", "bad C/H/max_steps");
    const MarchCfg c = make_cfg(bound, dt_gamma, max_steps, C, H);
    k_march_rays<<<ceil_div(n_alive, 8u), 256, 0, as_stream(stream)>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, c,
                                                                     grid, nears, fars, xyzs, dirs, deltas, perturb, n_alive_dev, occ_bounds);
    ENERF_CHECK_LAUNCH("march_rays");
    return 0;
}
int enerf_march_rays_dev(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                         const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                         uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                         float* dirs, float* deltas, uint32_t perturb, const int32_t* n_alive_dev, void* stream) {
    return enerf_march_rays_bounded(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, xyzs,
                                    dirs, deltas, perturb, n_alive_dev, nullptr, stream);
}
int enerf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                     const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                     uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                     float* dirs, float* deltas, uint32_t perturb, void* stream) {
    return enerf_march_rays_dev(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, xyzs, dirs,
                                deltas, perturb, nullptr, stream);
}

int enerf_composite_rays_dev(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t,
                             const float* sigmas, const float* rgbs, const float* deltas, uint32_t n_ch,
                             float* weights_sum, float* depth, float* image, const int32_t* n_alive_dev, void* stream) {
    if (n_alive == 0) return 0;
    // many steps per round: one warp per ray (coalesced, scans); few (the reference's n_step <= 8, and up to a warp's width, where most
    // lanes of a warp-per-ray kernel would idle: measured 138 vs 126 ms per 800x800 frame at n_step = 13): one thread per ray
    if (n_step >= 32) {
        ENERF_NCH_SWITCH(n_ch, "composite_rays",
                         (k_composite_rays_warp<NCH><<<ceil_div(n_alive, 8u), 256, 0, as_stream(stream)>>>(
                             n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, n_alive_dev)));
    } else {
        ENERF_NCH_SWITCH(n_ch, "composite_rays",
                         (k_composite_rays<NCH><<<ceil_div(n_alive, 128u), 128, 0, as_stream(stream)>>>(
                             n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, n_alive_dev)));
    }
    ENERF_CHECK_LAUNCH("composite_rays");
    return 0;
}
int enerf_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t,
                         const float* sigmas, const float* rgbs, const float* deltas, uint32_t n_ch,
                         float* weights_sum, float* depth, float* image, void* stream) {
    return enerf_composite_rays_dev(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, n_ch, weights_sum, depth, image, nullptr, stream);
}

int enerf_compact_rays_dev(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t,
                           const float* rays_t_old, int32_t* alive_counter, const int32_t* n_alive_dev, void* stream) {
    if (n_alive == 0) return 0;
    k_compact_rays<<<ceil_div(n_alive, 256u), 256, 0, as_stream(stream)>>>(n_alive, rays_alive, rays_alive_old, rays_t,
                                                                          rays_t_old, alive_counter, n_alive_dev);
    ENERF_CHECK_LAUNCH("compact_rays");
    return 0;
}
int enerf_compact_rays(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t,
                       const float* rays_t_old, int32_t* alive_counter, void* stream) {
    return enerf_compact_rays_dev(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter, nullptr, stream);
}

}  // extern "C"
