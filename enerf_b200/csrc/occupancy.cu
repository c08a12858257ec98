// Occupancy-grid maintenance of NeRFRenderer (nerf/renderer.py:408-563) as a handful of kernels — SURVEY.md K21 / row a14.
//
// The reference refreshes `density_grid` every 16 training steps (nerf/utils.py:945-947) with ~40 small ATen kernels, a 5-level
// Python loop and three host synchronisations (`torch.nonzero`, `mean().item()`, `step_counter.sum().item()`).  Here:
//   * k_occ_points_full / k_occ_points_partial  generate the jittered query positions directly in Morton order (cell index ->
//     coordinates -> world position of the cascade + jitter), so the density values come back already laid out like the grid;
//   * k_compact_*  an ORDER-PRESERVING stream compaction (what `torch.nonzero(grid > 0)` returns, without the host round trip);
//     also used by the run() path for the `weights > 1e-4` colour mask (renderer.py:236);
//   * k_occ_claim + k_occ_update  EMA-max update (`max(grid*decay, new)` where both are >= 0) with "last writer wins" for duplicate
//     cells of the partial update (the sequential index_put of the reference) and the sum of clamp(grid, 0) for the mean;
//   * k_packbits_mean  threshold = min(mean, density_thresh) read on the device -> bitfield;
//   * k_mark_untrained  cells no training camera sees get density -1 (renderer.py:408-471), one thread per cell, poses in smem.
// Arithmetic follows the reference's fp32 operation order (separate roundings, no contraction) so that a scripted-RNG run of the
// reference's own Python reproduces `density_grid` exactly (tests/golden/make_golden_grid_state.py).
#include "common.cuh"
#include <float.h>

namespace enerf {

static constexpr unsigned kFull = 0xffffffffu;

// world position of cell coordinate c in [0,H) of a cascade, jittered: renderer.py:499-508
//   xyz = 2*c/(H-1) - 1 ; cas_xyz = xyz * (bound - hgs) ; cas_xyz += (u*2 - 1) * hgs      (each op rounded to fp32)
__device__ __forceinline__ float cell_pos(uint32_t c, uint32_t H, float span, float hgs, float u) {
    const float x = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, (float)c), (float)(H - 1)), 1.0f);
    const float jitter = __fmul_rn(__fsub_rn(__fmul_rn(u, 2.0f), 1.0f), hgs);
    return __fadd_rn(__fmul_rn(x, span), jitter);
}

struct CascadeGeom {
    float span, hgs;    // (bound_c - half cell) and half cell = bound_c / H, both computed in double like the Python scalars
};
__device__ __forceinline__ CascadeGeom cascade_geom(uint32_t cas, uint32_t H, float bound) {
    const double b = fmin((double)(1u << cas), (double)bound);
    const double h = b / (double)H;
    CascadeGeom g;
    g.span = (float)(b - h);
    g.hgs = (float)h;
    return g;
}

// full refresh: sample t = cas*H^3 + m is cell with Morton index m of cascade cas (renderer.py:485-515)
__global__ void __launch_bounds__(256)
k_occ_points_full(float* __restrict__ xyzs, uint32_t C, uint32_t H, float bound, const float* __restrict__ noise, uint64_t seed) {
    const uint32_t cells = H * H * H;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= C * cells) return;
    const uint32_t cas = t / cells, m = t - cas * cells;
    const CascadeGeom g = cascade_geom(cas, H, bound);
    float u[3];
    if (noise) {
        u[0] = noise[(size_t)t * 3]; u[1] = noise[(size_t)t * 3 + 1]; u[2] = noise[(size_t)t * 3 + 2];
    } else {
        Pcg32 rng(seed, (uint64_t)t);
        u[0] = rng.next_float(); u[1] = rng.next_float(); u[2] = rng.next_float();
    }
    xyzs[(size_t)t * 3] = cell_pos(compact3(m), H, g.span, g.hgs, u[0]);
    xyzs[(size_t)t * 3 + 1] = cell_pos(compact3(m >> 1), H, g.span, g.hgs, u[1]);
    xyzs[(size_t)t * 3 + 2] = cell_pos(compact3(m >> 2), H, g.span, g.hgs, u[2]);
}

// partial refresh (renderer.py:517-545): per cascade n_pick uniformly random cells followed by n_pick cells drawn (with
// repetition) from the currently occupied ones (grid > 0; `occ_list` = their Morton indices in increasing order).
// sample t = cas*2*n_pick + k.  rand_coords [C,n_pick,3] / rand_occ [C,n_pick] / noise [C*2*n_pick,3]: scripted draws (tests);
// NULL = in-kernel PCG32 streams.  A cascade without occupied cells (the reference raises there) draws random cells instead.
__global__ void __launch_bounds__(256)
k_occ_points_partial(float* __restrict__ xyzs, int32_t* __restrict__ indices, uint32_t n_pick, uint32_t C, uint32_t H, float bound,
                     const int32_t* __restrict__ occ_list, const int32_t* __restrict__ occ_count, const int32_t* __restrict__ rand_coords,
                     const int32_t* __restrict__ rand_occ, const float* __restrict__ noise, uint64_t seed) {
    const uint32_t cells = H * H * H;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= C * 2 * n_pick) return;
    const uint32_t cas = t / (2 * n_pick), k = t - cas * 2 * n_pick;
    const CascadeGeom g = cascade_geom(cas, H, bound);
    Pcg32 rng(seed, (uint64_t)t);
    const uint32_t n_occ = (uint32_t)occ_count[cas];
    uint32_t m;
    if (k < n_pick || n_occ == 0) {
        uint32_t c[3];
        if (rand_coords && k < n_pick) {
            const int32_t* rc = rand_coords + ((size_t)cas * n_pick + k) * 3;
            c[0] = (uint32_t)rc[0]; c[1] = (uint32_t)rc[1]; c[2] = (uint32_t)rc[2];
        } else {
            c[0] = rng.next_uint() % H; c[1] = rng.next_uint() % H; c[2] = rng.next_uint() % H;
        }
        m = morton3(c[0], c[1], c[2]);
    } else {
        const uint32_t r = rand_occ ? (uint32_t)rand_occ[(size_t)cas * n_pick + (k - n_pick)]
                                    : min((uint32_t)(rng.next_float() * (float)n_occ), n_occ - 1);
        m = (uint32_t)occ_list[(size_t)cas * cells + r];
    }
    float u[3];
    if (noise) {
        u[0] = noise[(size_t)t * 3]; u[1] = noise[(size_t)t * 3 + 1]; u[2] = noise[(size_t)t * 3 + 2];
    } else {
        u[0] = rng.next_float(); u[1] = rng.next_float(); u[2] = rng.next_float();
    }
    indices[t] = (int32_t)m;
    xyzs[(size_t)t * 3] = cell_pos(compact3(m), H, g.span, g.hgs, u[0]);
    xyzs[(size_t)t * 3 + 1] = cell_pos(compact3(m >> 1), H, g.span, g.hgs, u[1]);
    xyzs[(size_t)t * 3 + 2] = cell_pos(compact3(m >> 2), H, g.span, g.hgs, u[2]);
}

// ---- order-preserving compaction of {i : values[i] > thresh} ------------------------------------------------------------
// 256 threads x 16 consecutive elements per CTA.  Pass 1 counts per CTA, pass 2 scans the CTA counts (one CTA), pass 3 writes.
static constexpr int kCompPerThread = 16;
static constexpr int kCompPerCta = 256 * kCompPerThread;

// flags of 16 consecutive elements: values[i] > thresh (float input) or mask[i] != 0 (byte input, a torch.bool tensor)
__device__ __forceinline__ uint32_t comp_flags(const float* __restrict__ v, float thresh, uint32_t n, uint32_t first) {
    uint32_t flags = 0;
    if (first + kCompPerThread <= n && ((reinterpret_cast<uintptr_t>(v + first) & 15u) == 0)) {
#pragma unroll
        for (int q = 0; q < kCompPerThread / 4; ++q) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(v + first) + q);
            flags |= (x.x > thresh ? 1u : 0u) << (4 * q) | (x.y > thresh ? 1u : 0u) << (4 * q + 1) | (x.z > thresh ? 1u : 0u) << (4 * q + 2) |
                     (x.w > thresh ? 1u : 0u) << (4 * q + 3);
        }
    } else {
        for (int e = 0; e < kCompPerThread; ++e)
            if (first + e < n && v[first + e] > thresh) flags |= 1u << e;
    }
    return flags;
}
__device__ __forceinline__ uint32_t comp_flags(const uint8_t* __restrict__ v, float, uint32_t n, uint32_t first) {
    uint32_t flags = 0;
    if (first + kCompPerThread <= n && ((reinterpret_cast<uintptr_t>(v + first) & 15u) == 0)) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(v + first));
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int b = 0; b < 4; ++b) flags |= (((w[q] >> (8 * b)) & 0xffu) ? 1u : 0u) << (4 * q + b);
    } else {
        for (int e = 0; e < kCompPerThread; ++e)
            if (first + e < n && v[first + e]) flags |= 1u << e;
    }
    return flags;
}

// CTA-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[8];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const uint32_t o = __shfl_up_sync(kFull, incl, k);
        if (lane >= (uint32_t)k) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t s = warp_sums[w];
        if ((uint32_t)w < warp) before += s;
        all += s;
    }
    *total = all;
    return before + incl - v;
}

template <typename V>
__global__ void __launch_bounds__(256)
k_compact_count(const V* __restrict__ values, float thresh, uint32_t n, int32_t* __restrict__ block_counts) {
    const uint32_t first = blockIdx.x * kCompPerCta + threadIdx.x * kCompPerThread;
    const uint32_t c = (first < n) ? __popc(comp_flags(values, thresh, n, first)) : 0u;
    uint32_t total;
    cta_exclusive_scan(c, &total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (int32_t)total;
}

// exclusive scan of block_counts[0..nb) in place (single CTA of 1024 threads); count[0] = total
__global__ void __launch_bounds__(1024)
k_compact_scan(int32_t* __restrict__ block_counts, uint32_t nb, int32_t* __restrict__ count) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < nb) ? (uint32_t)block_counts[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t o = __shfl_up_sync(kFull, incl, k);
            if (lane >= (uint32_t)k) incl += o;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        uint32_t before = 0, all = 0;
        for (int w = 0; w < 32; ++w) {
            const uint32_t s = warp_sums[w];
            if ((uint32_t)w < warp) before += s;
            all += s;
        }
        const uint32_t carry = carry_s;
        if (i < nb) block_counts[i] = (int32_t)(carry + before + incl - v);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) count[0] = (int32_t)carry_s;
}

template <typename V>
__global__ void __launch_bounds__(256)
k_compact_write(const V* __restrict__ values, float thresh, uint32_t n, const int32_t* __restrict__ block_offsets,
                int32_t* __restrict__ out) {
    const uint32_t first = blockIdx.x * kCompPerCta + threadIdx.x * kCompPerThread;
    const uint32_t flags = (first < n) ? comp_flags(values, thresh, n, first) : 0u;
    uint32_t total;
    uint32_t pos = (uint32_t)block_offsets[blockIdx.x] + cta_exclusive_scan(__popc(flags), &total);
#pragma unroll
    for (int e = 0; e < kCompPerThread; ++e)
        if (flags & (1u << e)) out[pos++] = (int32_t)(first + e);
}

// ---- EMA-max update -----------------------------------------------------------------------------------------------------
// partial update: the sample with the largest index that hit a cell owns it ("last writer wins", as the sequential index_put
// `tmp_grid[cas, indices] = sigmas` of renderer.py:545 does on the CPU; on the GPU the reference's winner is unspecified)
__global__ void __launch_bounds__(256)
k_occ_claim(int32_t* __restrict__ owner, const int32_t* __restrict__ indices, uint32_t per_cascade, uint32_t C, uint32_t cells) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= C * per_cascade) return;
    const uint32_t cas = t / per_cascade;
    atomicMax(owner + (size_t)cas * cells + (uint32_t)indices[t], (int32_t)t);
}

// per cell: new = sigma[owner]*scale (or -1 without a sample); grid = (grid >= 0 && new >= 0) ? max(grid*decay, new) : grid
// (renderer.py:548-549); sum += max(grid, 0) (renderer.py:550).  owner == NULL: sample index = cell index (full refresh).
__global__ void __launch_bounds__(256)
k_occ_update(float* __restrict__ grid, const float* __restrict__ sigmas, int32_t* __restrict__ owner, uint32_t n_cells, float decay, float scale,
             double* __restrict__ sum) {
    float local = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x) {
        int32_t o = (int32_t)i;
        if (owner) {
            o = owner[i];
            if (o >= 0) owner[i] = -1;                      // leave the scratch ready for the next refresh
        }
        float g = grid[i];
        if (o >= 0) {
            const float fresh = __fmul_rn(sigmas[o], scale);
            if (g >= 0.f && fresh >= 0.f) {
                g = fmaxf(__fmul_rn(g, decay), fresh);
                grid[i] = g;
            }
        }
        local += fmaxf(g, 0.f);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) local += __shfl_xor_sync(kFull, local, k);
    __shared__ float ws[8];
    if (lane_id() == 0) ws[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += (double)ws[w];
        atomicAdd(sum, s);
    }
}

// bitfield with the threshold min(mean(clamp(grid, 0)), density_thresh) taken from the device-side sum (renderer.py:550-555);
// bit order as k_packbits (raymarching.cu:283-290).  mean_out[0] = the mean (what `self.mean_density` holds in the reference).
__global__ void __launch_bounds__(256)
k_packbits_mean(const float* __restrict__ grid, uint32_t n_words, const double* __restrict__ sum, uint32_t n_cells, float density_thresh,
                uint32_t* __restrict__ bitfield, float* __restrict__ mean_out) {
    const float mean = (float)(*sum / (double)n_cells);
    const float thresh = fminf(mean, density_thresh);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = lane_id();
    if (blockIdx.x == 0 && threadIdx.x == 0) mean_out[0] = mean;
    const uint32_t word0 = warp * 8;
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t w = word0 + j;
        float v = -FLT_MAX;
        if (w < n_words) v = grid[(size_t)w * 32 + lane];
        const uint32_t bits = __ballot_sync(kFull, v > thresh);
        if (lane == (uint32_t)j) mine = bits;
    }
    if (lane < 8 && word0 + lane < n_words) bitfield[word0 + lane] = mine;
}

// ---- mark_untrained_grid (renderer.py:408-471) --------------------------------------------------------------------------
// thread per (cascade, cell); camera-to-world poses [B, 4, 4] row-major are staged in shared memory 128 at a time.
//   cam = (p - t) @ R            (p = un-jittered cell centre of the cascade)
//   seen = cam.z > 0 && |cam.x| < cx/fx*cam.z + 2*hgs && |cam.y| < cy/fy*cam.z + 2*hgs
// a cell no pose sees gets density -1 and is never updated again (renderer.py:469, :548).
__global__ void __launch_bounds__(256)
k_mark_untrained(float* __restrict__ grid, const float* __restrict__ poses, uint32_t B, float cx_fx, float cy_fy, uint32_t C, uint32_t H,
                 float bound) {
    __shared__ float sp[128 * 12];
    const uint32_t cells = H * H * H;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < C * cells;
    const uint32_t cas = active ? t / cells : 0u, m = active ? t - cas * cells : 0u;
    const CascadeGeom g = cascade_geom(cas, H, bound);
    const float margin = (float)((double)g.hgs * 2.0);
    float p[3];
    p[0] = cell_pos(compact3(m), H, g.span, 0.f, 0.5f);
    p[1] = cell_pos(compact3(m >> 1), H, g.span, 0.f, 0.5f);
    p[2] = cell_pos(compact3(m >> 2), H, g.span, 0.f, 0.5f);
    bool seen = false;
    for (uint32_t base = 0; base < B; base += 128) {
        const uint32_t nb = min(128u, B - base);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nb * 12; i += blockDim.x) {
            const uint32_t b = i / 12, e = i - b * 12;       // rows 0..2 of the 4x4 matrix
            sp[i] = poses[(size_t)(base + b) * 16 + (e / 4) * 4 + (e % 4)];
        }
        __syncthreads();
        if (active && !seen) {
            for (uint32_t b = 0; b < nb; ++b) {
                const float* P = sp + b * 12;                // P[r*4+c]
                const float d0 = __fsub_rn(p[0], P[3]), d1 = __fsub_rn(p[1], P[7]), d2 = __fsub_rn(p[2], P[11]);
                // (d @ R)_j = d0*R[0][j] + d1*R[1][j] + d2*R[2][j]
                const float cxv = __fadd_rn(__fadd_rn(__fmul_rn(d0, P[0]), __fmul_rn(d1, P[4])), __fmul_rn(d2, P[8]));
                const float cyv = __fadd_rn(__fadd_rn(__fmul_rn(d0, P[1]), __fmul_rn(d1, P[5])), __fmul_rn(d2, P[9]));
                const float czv = __fadd_rn(__fadd_rn(__fmul_rn(d0, P[2]), __fmul_rn(d1, P[6])), __fmul_rn(d2, P[10]));
                if (czv > 0.f && fabsf(cxv) < __fadd_rn(__fmul_rn(cx_fx, czv), margin) && fabsf(cyv) < __fadd_rn(__fmul_rn(cy_fy, czv), margin)) {
                    seen = true;
                    break;
                }
            }
        }
    }
    if (active && !seen) grid[t] = -1.0f;
}

// ---- row gather / scatter (32-bit words) and the weighted ray sum of NeRFRenderer.run ------------------------------------
// dst[i] = src[idx[i]] for i < n, zero rows for n <= i < n_pad
__global__ void __launch_bounds__(256)
k_gather_rows(const uint32_t* __restrict__ src, const int32_t* __restrict__ idx, uint32_t n, uint32_t n_pad, uint32_t wpr,
              uint32_t* __restrict__ dst, const int32_t* __restrict__ n_dev) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) {                                   // row count on the device: n = capacity, zero rows up to the next multiple of 128
        n = min(n, (uint32_t)max(*n_dev, 0));
        n_pad = min(n_pad, (n + 127u) / 128u * 128u);
    }
    if (t >= (uint64_t)n_pad * wpr) return;
    const uint32_t i = (uint32_t)(t / wpr), w = (uint32_t)(t - (uint64_t)i * wpr);
    dst[t] = (i < n) ? src[(size_t)idx[i] * wpr + w] : 0u;
}
// dst[idx[i]] = src[i] for i < n (dst pre-initialised by the caller)
__global__ void __launch_bounds__(256)
k_scatter_rows(const uint32_t* __restrict__ src, const int32_t* __restrict__ idx, uint32_t n, uint32_t wpr, uint32_t* __restrict__ dst,
               const int32_t* __restrict__ n_dev) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n = min(n, (uint32_t)max(*n_dev, 0));
    if (t >= (uint64_t)n * wpr) return;
    const uint32_t i = (uint32_t)(t / wpr), w = (uint32_t)(t - (uint64_t)i * wpr);
    dst[(size_t)idx[i] * wpr + w] = src[t];
}

// image[n,c] = sum_t w[n,t] * rgb[n,t,c]  (renderer.py:255), warp per ray
template <int NCH>
__global__ void __launch_bounds__(256)
k_weighted_sum_fwd(const float* __restrict__ w, const float* __restrict__ rgb, uint32_t N, uint32_t T, float* __restrict__ image) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const uint32_t lane = lane_id();
    float acc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[c] = 0.f;
    for (uint32_t i = lane; i < T; i += 32) {
        const float wi = w[(size_t)n * T + i];
#pragma unroll
        for (int c = 0; c < NCH; ++c) acc[c] = __fmaf_rn(wi, rgb[((size_t)n * T + i) * NCH + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) acc[c] += __shfl_xor_sync(kFull, acc[c], k);
        if (lane == 0) image[(size_t)n * NCH + c] = acc[c];
    }
}
// g_w[n,t] = sum_c g_img[n,c]*rgb[n,t,c] ; g_rgb[n,t,c] = w[n,t]*g_img[n,c]
template <int NCH>
__global__ void __launch_bounds__(256)
k_weighted_sum_bwd(const float* __restrict__ g_img, const float* __restrict__ w, const float* __restrict__ rgb, uint64_t total, uint32_t T,
                   float* __restrict__ g_w, float* __restrict__ g_rgb) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const uint64_t n = t / T;
    const float wi = w[t];
    float gw = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const float g = g_img[n * NCH + c];
        gw = __fmaf_rn(g, rgb[t * NCH + c], gw);
        if (g_rgb) g_rgb[t * NCH + c] = wi * g;
    }
    if (g_w) g_w[t] = gw;
}

}  // namespace enerf

using namespace enerf;

extern "C" {

int enerf_occ_points_full(float* xyzs, uint32_t C, uint32_t H, float bound, const float* noise, uint64_t seed, void* stream) {
    ENERF_REQUIRE(C >= 1 && C <= 16 && H >= 2 && H <= 1024, "occ_points_full", "bad C/H");
    const uint32_t n = C * H * H * H;
    k_occ_points_full<<<ceil_div(n, 256u), 256, 0, as_stream(stream)>>>(xyzs, C, H, bound, noise, seed);
    ENERF_CHECK_LAUNCH("occ_points_full");
    return 0;
}

int enerf_occ_points_partial(float* xyzs, int32_t* indices, uint32_t n_pick, uint32_t C, uint32_t H, float bound, const int32_t* occ_list,
                             const int32_t* occ_count, const int32_t* rand_coords, const int32_t* rand_occ, const float* noise, uint64_t seed,
                             void* stream) {
    ENERF_REQUIRE(C >= 1 && C <= 16 && H >= 2 && H <= 1024, "occ_points_partial", "bad C/H");
    ENERF_REQUIRE(occ_list != nullptr && occ_count != nullptr, "occ_points_partial", "occ_list / occ_count must not be NULL");
    if (n_pick == 0) return 0;
    const uint32_t n = C * 2 * n_pick;
    k_occ_points_partial<<<ceil_div(n, 256u), 256, 0, as_stream(stream)>>>(xyzs, indices, n_pick, C, H, bound, occ_list, occ_count, rand_coords,
                                                                         rand_occ, noise, seed);
    ENERF_CHECK_LAUNCH("occ_points_partial");
    return 0;
}

}  // extern "C"

template <typename V>
static int compact_launch(const char* name, const V* values, float thresh, uint32_t n, int32_t* indices, int32_t* count, int32_t* scratch, void* stream) {
    ENERF_REQUIRE(count != nullptr && scratch != nullptr, name, "count / scratch must not be NULL");
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        ENERF_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st), name);
        return 0;
    }
    const uint32_t nb = ceil_div(n, (uint32_t)kCompPerCta);
    k_compact_count<V><<<nb, 256, 0, st>>>(values, thresh, n, scratch);
    ENERF_CHECK_LAUNCH(name);
    k_compact_scan<<<1, 1024, 0, st>>>(scratch, nb, count);
    ENERF_CHECK_LAUNCH(name);
    if (indices) {
        k_compact_write<V><<<nb, 256, 0, st>>>(values, thresh, n, scratch, indices);
        ENERF_CHECK_LAUNCH(name);
    }
    return 0;
}

extern "C" {

int enerf_compact_greater(const float* values, float thresh, uint32_t n, int32_t* indices, int32_t* count, int32_t* scratch, void* stream) {
    return compact_launch<float>("compact_greater", values, thresh, n, indices, count, scratch, stream);
}
int enerf_compact_mask(const uint8_t* mask, uint32_t n, int32_t* indices, int32_t* count, int32_t* scratch, void* stream) {
    return compact_launch<uint8_t>("compact_mask", mask, 0.f, n, indices, count, scratch, stream);
}

int enerf_occ_update(float* density_grid, const float* sigmas, const int32_t* indices, uint32_t per_cascade, uint32_t C, uint32_t H, float decay,
                     float scale, float density_thresh, int32_t* owner, double* sum, uint8_t* bitfield, float* mean_density, void* stream) {
    ENERF_REQUIRE(C >= 1 && C <= 16 && H >= 2 && H <= 1024, "occ_update", "bad C/H");
    ENERF_REQUIRE(sum != nullptr && mean_density != nullptr && bitfield != nullptr, "occ_update", "sum / mean_density / bitfield must not be NULL");
    ENERF_REQUIRE(indices == nullptr || owner != nullptr, "occ_update", "a partial update needs the owner scratch (int32 [C*H^3], filled with -1)");
    ENERF_REQUIRE(((uintptr_t)bitfield & 3u) == 0 && (H * H * H) % 32 == 0, "occ_update", "bitfield must be 4-byte aligned, H^3 a multiple of 32");
    cudaStream_t st = as_stream(stream);
    const uint32_t cells = H * H * H, n_cells = C * cells;
    ENERF_CUDA(cudaMemsetAsync(sum, 0, sizeof(double), st), "occ_update");
    if (indices && per_cascade > 0) {
        k_occ_claim<<<ceil_div(C * per_cascade, 256u), 256, 0, st>>>(owner, indices, per_cascade, C, cells);
        ENERF_CHECK_LAUNCH("occ_update(claim)");
    }
    const uint32_t blocks = min(ceil_div(n_cells, 256u), (uint32_t)num_sms() * 8u);
    k_occ_update<<<blocks, 256, 0, st>>>(density_grid, sigmas, indices ? owner : nullptr, n_cells, decay, scale, sum);
    ENERF_CHECK_LAUNCH("occ_update");
    const uint32_t n_words = n_cells / 32;
    k_packbits_mean<<<ceil_div(ceil_div(n_words, 8u) * 32u, 256u), 256, 0, st>>>(density_grid, n_words, sum, n_cells, density_thresh,
                                                                                reinterpret_cast<uint32_t*>(bitfield), mean_density);
    ENERF_CHECK_LAUNCH("occ_update(packbits)");
    return 0;
}

int enerf_mark_untrained_grid(float* density_grid, const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t C, uint32_t H,
                              float bound, void* stream) {
    ENERF_REQUIRE(C >= 1 && C <= 16 && H >= 2 && H <= 1024, "mark_untrained_grid", "bad C/H");
    const uint32_t n = C * H * H * H;
    // `cx / fx` and `cy / fy` are Python doubles in the reference, rounded to fp32 when they meet the tensor
    k_mark_untrained<<<ceil_div(n, 256u), 256, 0, as_stream(stream)>>>(density_grid, poses, B, (float)((double)cx / (double)fx),
                                                                      (float)((double)cy / (double)fy), C, H, bound);
    ENERF_CHECK_LAUNCH("mark_untrained_grid");
    return 0;
}

int enerf_gather_rows(const void* src, const int32_t* idx, uint32_t n, uint32_t n_pad, uint32_t row_bytes, void* dst, const int32_t* n_dev, void* stream) {
    ENERF_REQUIRE(row_bytes % 4 == 0 && row_bytes > 0 && n_pad >= n, "gather_rows", "row_bytes must be a positive multiple of 4, n_pad >= n");
    if (n_pad == 0) return 0;
    const uint32_t wpr = row_bytes / 4;
    const uint64_t total = (uint64_t)n_pad * wpr;
    k_gather_rows<<<(uint32_t)ceil_div(total, (uint64_t)256), 256, 0, as_stream(stream)>>>((const uint32_t*)src, idx, n, n_pad, wpr, (uint32_t*)dst, n_dev);
    ENERF_CHECK_LAUNCH("gather_rows");
    return 0;
}

int enerf_scatter_rows(const void* src, const int32_t* idx, uint32_t n, uint32_t row_bytes, void* dst, const int32_t* n_dev, void* stream) {
    ENERF_REQUIRE(row_bytes % 4 == 0 && row_bytes > 0, "scatter_rows", "row_bytes must be a positive multiple of 4");
    if (n == 0) return 0;
    const uint32_t wpr = row_bytes / 4;
    const uint64_t total = (uint64_t)n * wpr;
    k_scatter_rows<<<(uint32_t)ceil_div(total, (uint64_t)256), 256, 0, as_stream(stream)>>>((const uint32_t*)src, idx, n, wpr, (uint32_t*)dst, n_dev);
    ENERF_CHECK_LAUNCH("scatter_rows");
    return 0;
}

#define ENERF_WS_SWITCH(n_ch, name, CALL)                                            \
    switch (n_ch) {                                                                  \
        case 1: { constexpr int NCH = 1; CALL; } break;                              \
        case 2: { constexpr int NCH = 2; CALL; } break;                              \
        case 3: { constexpr int NCH = 3; CALL; } break;                              \
        case 4: { constexpr int NCH = 4; CALL; } break;                              \
        default: set_error("%s: n_ch must be in [1,4]", name); return -2;            \
    }

int enerf_weighted_sum_forward(const float* weights, const float* rgbs, uint32_t N, uint32_t T, uint32_t n_ch, float* image, void* stream) {
    if (N == 0) return 0;
    ENERF_WS_SWITCH(n_ch, "weighted_sum_forward", (k_weighted_sum_fwd<NCH><<<ceil_div(N, 8u), 256, 0, as_stream(stream)>>>(weights, rgbs, N, T, image)));
    ENERF_CHECK_LAUNCH("weighted_sum_forward");
    return 0;
}

int enerf_weighted_sum_backward(const float* grad_image, const float* weights, const float* rgbs, uint32_t N, uint32_t T, uint32_t n_ch,
                                float* grad_weights, float* grad_rgbs, void* stream) {
    const uint64_t total = (uint64_t)N * T;
    if (total == 0) return 0;
    ENERF_WS_SWITCH(n_ch, "weighted_sum_backward",
                    (k_weighted_sum_bwd<NCH><<<(uint32_t)ceil_div(total, (uint64_t)256), 256, 0, as_stream(stream)>>>(grad_image, weights, rgbs, total, T,
                                                                                                                 grad_weights, grad_rgbs)));
    ENERF_CHECK_LAUNCH("weighted_sum_backward");
    return 0;
}

}  // extern "C"
